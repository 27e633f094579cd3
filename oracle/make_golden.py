"""Generate tests/golden/*.npz by executing the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # regenerates every fixture

The reference has no tests or golden vectors of its own, so
the fixtures are outputs of the reference's own functions on seeded inputs.
Each fixture stores the inputs too, so the GPU-box tests need nothing else.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def gae_inputs(T, N, seed=0, p_term=0.01, p_tl=0.005):
    """SURVEY.md section 8(d) recipe, in this exact draw order."""
    rs = np.random.RandomState(seed)
    values = rs.randn(T, N, 1)
    rewards = rs.randn(T, N, 1)
    terminals = (rs.rand(T, N, 1) < p_term)
    time_limits = (rs.rand(T, N, 1) < p_tl)
    last_value = rs.randn(N, 1)
    return values, rewards, terminals, time_limits, last_value


def golden_gae(trl):
    from torchrl.replay_buffers import OnPolicyReplayBuffer
    out = {}
    cases = [(128, 8, 0, 0.01, 0.005), (37, 5, 1, 0.2, 0.1), (1, 3, 2, 0.5, 0.5), (256, 33, 3, 0.02, 0.02)]
    for ci, (T, N, seed, pt, ptl) in enumerate(cases):
        values, rewards, terminals, time_limits, last_value = gae_inputs(T, N, seed, pt, ptl)
        for flt in (True, False):
            buf = OnPolicyReplayBuffer(max_replay_buffer_size=T * N, env_nums=N, time_limit_filter=flt)
            buf._values = values.astype(np.float64)
            buf._rewards = rewards.astype(np.float64)
            buf._terminals = terminals.astype(np.float64)
            buf._time_limits = time_limits.astype(np.float64)
            buf.generalized_advantage_estimation(last_value, 0.99, 0.95)
            tag = "c%d_f%d" % (ci, int(flt))
            out[tag + "_gae_advs"] = buf._advs.copy()
            out[tag + "_gae_rets"] = buf._estimate_returns.copy()
            buf.discount_reward(last_value, 0.99)
            out[tag + "_disc_advs"] = buf._advs.copy()
            out[tag + "_disc_rets"] = buf._estimate_returns.copy()
        out["c%d_values" % ci] = values
        out["c%d_rewards" % ci] = rewards
        out["c%d_terminals" % ci] = terminals
        out["c%d_time_limits" % ci] = time_limits
        out["c%d_last_value" % ci] = last_value
    out["ncases"] = np.array(len(cases))
    out["gamma_tau"] = np.array([0.99, 0.95])
    np.savez_compressed(os.path.join(GOLDEN, "gae.npz"), **out)
    print("gae.npz:", len(out), "arrays")


class _Quiet:
    """Logger stand-in for the reference agent (it only calls these)."""
    def add_update_info(self, info): pass
    def add_epoch_info(self, *a, **k): pass
    def log(self, *a): pass
    def finish(self): pass


def golden_hotpath(trl):
    """Seeded inputs -> outputs of the reference's OWN functions for the non-GAE rows of SURVEY.md 8(a):
    K2 Normalizer (env/base_wrapper.py:63-100), K3 TanhNormal.log_prob / entropy (policies/distribution.py:33-79),
    K7/K9 minibatch row order and uniform row indices (replay_buffers/on_policy.py:72-91, base.py:39-51),
    K8 PPO.update on one minibatch (algo/on_policy/ppo.py:41-150; the policy / value network OUTPUTS at the
    moment of the update are stored as the loss kernels' inputs), K10 quantile_regression_loss (algo/utils.py:5-13)."""
    import tempfile
    import torch
    import torchrl.networks as networks
    import torchrl.policies as policies
    from torchrl.algo import PPO
    from torchrl.algo.utils import quantile_regression_loss
    from torchrl.collector.on_policy import VecOnPolicyCollector
    from torchrl.env import get_vec_env
    from torchrl.env.base_wrapper import Normalizer
    from torchrl.policies.distribution import TanhNormal
    from torchrl.replay_buffers.on_policy import OnPolicyReplayBuffer
    from torchrl.replay_buffers.base import BaseReplayBuffer
    out = {}

    # ---- K2: running normaliser over 6 batches of (64, 17)
    rs = np.random.RandomState(11)
    batches = rs.randn(6, 64, 17) * rs.uniform(0.5, 3.0, size=(1, 1, 17)) + 2.0 * rs.randn(1, 1, 17)
    nrm = Normalizer((17,))
    means, vars_, counts, filts = [], [], [], []
    for k in range(6):
        nrm.update_estimate(batches[k])
        means.append(nrm._mean.copy()); vars_.append(nrm._var.copy()); counts.append(nrm._count)
        filts.append(nrm.filt(batches[k]))
    far = batches[0] * 50.0                              # filter-only call (eval mode) far outside: +-10 clip
    out.update(norm_batches=batches, norm_mean=np.array(means), norm_var=np.array(vars_),
               norm_count=np.array(counts), norm_filt=np.array(filts), norm_far=far, norm_far_filt=nrm.filt(far))

    # ---- K3: tanh-Gaussian log-density and entropy (float64 tensors: the formula, not fp32 noise)
    rs = np.random.RandomState(12)
    mean = rs.randn(512, 6)
    log_std = 0.3 * rs.randn(6) - 1.0
    std = np.exp(log_std)[None].repeat(512, 0)
    acts = np.tanh(mean + std * rs.randn(512, 6) * 1.2).astype(np.float32).astype(np.float64)
    dist = TanhNormal(torch.from_numpy(mean), torch.from_numpy(std))
    out.update(tn_mean=mean, tn_log_std=log_std, tn_acts=acts,
               tn_log_prob=dist.log_prob(torch.from_numpy(acts)).numpy(), tn_entropy=dist.entropy().numpy())

    # ---- K7 / K9: row order of one_iteration and row indices of random_batch from the global NumPy stream
    T, N = 24, 4
    buf = OnPolicyReplayBuffer(max_replay_buffer_size=T * N, env_nums=N)
    rows = np.arange(T, dtype=np.float64)[:, None, None].repeat(N, 1)
    buf._obs = rows.copy()
    buf._max_replay_buffer_size = T
    np.random.seed(5)
    order = []
    for _pass in range(2):
        for mb in buf.one_iteration(batch_size=6 * N, sample_key=["obs"], shuffle=True):
            order.append(mb["obs"].reshape(6, N)[:, 0])
    out["iter_rows"] = np.array(order).astype(np.int64)               # (2*T/6, 6) row indices, bit-exact contract
    ring = BaseReplayBuffer(max_replay_buffer_size=40 * N, env_nums=N)
    ring._obs = np.arange(40, dtype=np.float64)[:, None, None].repeat(N, 1)
    ring._size = 33
    np.random.seed(6)
    out["rand_rows"] = np.array([ring.random_batch(5 * N, ["obs"])["obs"].reshape(5, N)[:, 0] for _ in range(4)]
                                ).astype(np.int64)

    # ---- K8: one PPO minibatch update by the reference agent
    for tag, clipped in (("ppo", False), ("ppoc", True)):
        torch.manual_seed(21)
        np.random.seed(21)
        Nenv, B = 4, 256
        params = {"reward_scale": 1, "obs_norm": False}
        env = get_vec_env("SynthHalfCheetah-v0", dict(params), Nenv)
        eval_env = get_vec_env("SynthHalfCheetah-v0", dict(params), Nenv)
        rb = OnPolicyReplayBuffer(env_nums=Nenv, max_replay_buffer_size=8 * Nenv, time_limit_filter=True)
        net = dict(hidden_shapes=[32, 32], append_hidden_shapes=[], base_type=networks.MLPBase,
                   activation_func=torch.nn.Tanh)
        pf = policies.GuassianContPolicyBasicBias(input_shape=17, output_shape=6, tanh_action=True, **net)
        vf = networks.Net(input_shape=(17,), output_shape=1, **net)
        col = VecOnPolicyCollector(vf, env=env, eval_env=eval_env, pf=pf, replay_buffer=rb, device="cpu",
                                   train_render=False, epoch_frames=8 * Nenv, max_episode_frames=999, eval_episodes=1)
        agent = PPO(pf=pf, vf=vf, plr=3e-4, vlr=3e-4, clip_para=0.2, opt_epochs=1, tau=0.95, shuffle=True,
                    entropy_coeff=0.005, env=env, replay_buffer=rb, collector=col, logger=_Quiet(), discount=0.99,
                    num_epochs=10, batch_size=B, gae=True, device="cpu", save_dir=tempfile.mkdtemp(),
                    clipped_value_loss=clipped)
        with torch.no_grad():
            for prm in agent.target_pf.parameters():               # old policy != new policy: ratios leave 1 +- clip
                prm.add_(0.01 * torch.randn_like(prm))
        rs = np.random.RandomState(22)
        obs = rs.randn(B, 17)
        obs_t = torch.Tensor(obs)
        with torch.no_grad():
            mean_t, std_t, log_std_t = pf(obs_t)
        acts = np.tanh(mean_t.numpy() + std_t.numpy() * rs.randn(B, 6)).astype(np.float32).astype(np.float64)
        batch = {"obs": obs, "acts": acts, "advs": 2.0 * rs.randn(B, 1) + 0.3,
                 "values": rs.randn(B, 1), "estimate_returns": rs.randn(B, 1)}
        acts_t = torch.Tensor(batch["acts"])
        with torch.no_grad():
            old_lp = agent.target_pf.update(obs_t, acts_t)["log_prob"]
            v_t = vf(obs_t)
        info = agent.update(batch)
        for k, v in batch.items():
            out["%s_%s" % (tag, k)] = v
        out["%s_mean" % tag] = mean_t.numpy().astype(np.float64)
        out["%s_log_std" % tag] = log_std_t.numpy().astype(np.float64).reshape(-1)
        out["%s_old_logp" % tag] = old_lp.numpy().astype(np.float64)
        out["%s_v" % tag] = v_t.numpy().astype(np.float64)
        keys = sorted(info.keys())
        out["%s_info_keys" % tag] = np.array(keys)
        out["%s_info_vals" % tag] = np.array([info[k] for k in keys], dtype=np.float64)

    # ---- K10: quantile-regression Huber loss (algo/utils.py:5-13) on (B, Q) source / target
    rs = np.random.RandomState(31)
    Bq, Q = 32, 51
    src, tgt = rs.randn(Bq, Q), rs.randn(Bq, Q) * 1.5 + 0.2
    tau = (2 * np.arange(Q) + 1) / (2.0 * Q)
    out.update(qr_source=src, qr_target=tgt, qr_tau=tau,
               qr_loss=np.array(quantile_regression_loss(torch.from_numpy(tau), torch.from_numpy(src),
                                                         torch.from_numpy(tgt)).item()))
    np.savez_compressed(os.path.join(GOLDEN, "hotpath.npz"), **out)
    print("hotpath.npz:", len(out), "arrays")


def main():
    from oracle import reference_loader
    trl = reference_loader.load()
    os.makedirs(GOLDEN, exist_ok=True)
    only = sys.argv[1:]
    for name, fn in sorted(globals().items()):
        if name.startswith("golden_") and callable(fn):
            if only and name[len("golden_"):] not in only:
                continue
            fn(trl)


if __name__ == "__main__":
    main()
