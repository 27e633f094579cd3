"""Host-side vec envs: the product's VecEnv / SubProcVecEnv for real CPU envs and the
inline wrapper chain, checked against the oracle's restatement of the reference classes; the GPU test
drives a full PPO epoch through the pinned-memory device bridge and compares it with the oracle port."""
import numpy as np
import pytest

from oracle import cpu_envs, synth_env

ENV = "SynthHalfCheetah-v0"
PARAM = {"reward_scale": 0.5}


def _make(env_id, env_param):
    """Top-level so that spawned workers can unpickle it: HostEnv over the synthetic CPU env."""
    from torchrl_b200.hostenv import get_single_env
    return get_single_env(env_id, env_param, make=synth_env.make_env)


def _rollout(venv, steps, seed, acts):
    venv.seed(seed)
    out = [np.array(venv.reset())]
    for t in range(steps):
        ob, r, d, info = venv.step(acts[t])
        out += [np.array(ob), np.array(r), np.array(d), np.array(info["time_limit"])]
        if d.any():
            out.append(np.array(venv.partial_reset(d.squeeze(-1))))
    return out


@pytest.mark.parametrize("procs,adt", [(0, np.float64), (2, np.float64), (0, np.float32), (2, np.float32)])
def test_host_vecenv_matches_oracle_vecenv(procs, adt):
    from torchrl_b200.hostenv import VecEnv, SubProcVecEnv
    N, T = 6, 12
    rs = np.random.RandomState(3)
    acts = rs.uniform(-1.3, 1.3, size=(T, N, 6)).astype(adt)
    ref = cpu_envs.InProcVecEnv(N, ENV, PARAM)
    for e in ref.envs:
        e.core._max_episode_steps = 5          # force time-limit episode ends inside the rollout
    want = _rollout(ref, T, 7, acts)
    if procs:
        venv = SubProcVecEnv(procs, N, _make_short, [ENV, PARAM])
    else:
        venv = VecEnv(N, _make_short, [ENV, PARAM])
    try:
        assert venv.observation_space.shape == (17,) and venv.action_space.shape == (6,)
        got = _rollout(venv, T, 7, acts)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert g.shape == w.shape and g.dtype == w.dtype
            np.testing.assert_array_equal(g, w)
        # eval mode switches the reward scale off (RewardShift is train-only)
        venv.eval(); ref.eval()
        r1 = venv.step(acts[0])[1].copy(); r2 = ref.step(acts[0])[1]
        np.testing.assert_array_equal(r1, r2)
    finally:
        venv.close()


def _make_short(env_id, env_param):
    env = _make(env_id, env_param)
    env.env._max_episode_steps = 5
    return env


def test_subproc_seed_is_applied():
    """The reference's SubProcVecEnv.seed is a silent no-op (A.3); here it reseeds the workers' envs."""
    from torchrl_b200.hostenv import SubProcVecEnv
    venv = SubProcVecEnv(2, 4, _make, [ENV, PARAM])
    try:
        venv.seed(1); a = venv.reset().copy()
        venv.seed(2); b = venv.reset().copy()
        venv.seed(1); c = venv.reset().copy()
        assert not np.array_equal(a, b)
        np.testing.assert_array_equal(a, c)
    finally:
        venv.close()


def _make_real(env_id, env_param):
    return _make(env_id, env_param)


@pytest.mark.gpu
@pytest.mark.parametrize("procs,max_frames", [(0, 999), (2, 7)])
def test_ppo_epoch_through_host_bridge_matches_reference_port(procs, max_frames):
    """A REAL host env (here: the synthetic CPU env behind the product's VecEnv / SubProcVecEnv) feeding the
    GPU pipeline through the pinned-memory bridge reproduces the oracle port's PPO epoch -- the same bar as
    tests/test_reference_parity.py holds for the device-resident env."""
    import torch
    import torchrl_b200.networks as networks
    import torchrl_b200.policies as policies
    from oracle import ref_port
    from torchrl_b200.algo import PPO
    from torchrl_b200.collector import VecOnPolicyCollector
    from torchrl_b200.env import HostEnvBridge
    from torchrl_b200.hostenv import VecEnv, SubProcVecEnv
    from torchrl_b200.policies import set_noise_mode
    from torchrl_b200.replay_buffers import OnPolicyReplayBuffer
    from torchrl_b200.utils import NullLogger
    N, T, hidden, rows, oe, seed = 8, 16, (32, 32), 4, 2, 5
    keys = ("obs", "next_obs", "acts", "values", "rewards", "terminals", "time_limits")
    param = {"reward_scale": 1, "obs_norm": True}
    torch.set_num_threads(4)
    penv, pcol, pagent = ref_port.build_ppo(env_nums=N, horizon=T, hidden=hidden, batch_rows=rows, opt_epochs=oe,
                                            seed=seed, max_episode_frames=max_frames)
    pagent.current_epoch = 0
    p_out = pcol.train_one_epoch()
    roll = {k: pagent.buffer.data[k].copy() for k in keys}
    pagent.update_per_epoch()

    def host():
        hp = {"reward_scale": 1}
        return SubProcVecEnv(procs, N, _make_real, [ENV, hp]) if procs else VecEnv(N, _make_real, [ENV, hp])

    dev = torch.device("cuda:0")
    env, eval_env = HostEnvBridge(host(), param, device=dev), HostEnvBridge(host(), param, device=dev)
    set_noise_mode("reference_cpu")
    try:
        env.seed(seed)
        torch.manual_seed(seed)
        np.random.seed(seed)
        buf = OnPolicyReplayBuffer(env_nums=N, max_replay_buffer_size=T * N, time_limit_filter=True)
        net = dict(hidden_shapes=list(hidden), append_hidden_shapes=[], base_type=networks.MLPBase,
                   activation_func=torch.nn.Tanh)
        pf = policies.GuassianContPolicyBasicBias(input_shape=17, output_shape=6, tanh_action=True, **net)
        vf = networks.Net(input_shape=(17,), output_shape=1, **net)
        col = VecOnPolicyCollector(vf, env=env, eval_env=eval_env, pf=pf, replay_buffer=buf, device=dev,
                                   train_render=False, epoch_frames=T * N, max_episode_frames=max_frames,
                                   eval_episodes=1)
        agent = PPO(pf=pf, vf=vf, plr=3e-4, vlr=3e-4, clip_para=0.2, opt_epochs=oe, tau=0.95, shuffle=True,
                    entropy_coeff=0.005, env=env, replay_buffer=buf, collector=col, logger=NullLogger(),
                    discount=0.99, num_epochs=10, batch_size=rows * N, gae=True, device=dev, save_dir=None)
        agent.current_epoch = 0
        out = col.train_one_epoch()
        assert abs(out["train_epoch_reward"] - p_out["train_epoch_reward"]) < 1e-3
        assert len(out["train_rewards"]) == len(p_out["train_rewards"])
        for k in keys:
            got = getattr(buf, "_" + k).cpu().numpy().astype(np.float64)
            np.testing.assert_allclose(got, roll[k].reshape(got.shape), rtol=1e-4, atol=1e-4, err_msg=k)
        agent.update_per_epoch()
        np.testing.assert_allclose(buf._advs.cpu().numpy(), pagent.buffer.data["advs"], rtol=1e-3, atol=2e-4)
        mine = torch.cat([p.detach().reshape(-1) for p in list(agent.pf.mean_params()) +
                          list(agent.vf.parameters())]).cpu().numpy()
        ref = torch.cat([p.detach().reshape(-1) for p in list(pagent.pf.net.parameters()) +
                         list(pagent.vf.parameters())]).numpy()
        np.testing.assert_allclose(mine, ref, atol=2e-4)
        nrm = env._obs_normalizer
        np.testing.assert_allclose(nrm._mean.cpu().numpy(), penv.norm.mean, rtol=1e-5, atol=1e-6)
        # a second epoch + the evaluation loop run through the bridge as well
        col.train_one_epoch()
        ev = col.eval_one_epoch()
        assert len(ev["eval_rewards"]) == N and np.isfinite(ev["eval_traj_length"])
    finally:
        set_noise_mode("philox")
        env.close()
        eval_env.close()
