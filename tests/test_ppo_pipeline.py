"""End-to-end PPO on the device: collector -> GAE -> minibatch updates.

Self-consistency (this file): the CUDA-graph path and the eager path produce identical buffers,
parameters and logged infos from the same seeds; the stored rollout satisfies the collector
semantics of the reference (obs[t+1] == next_obs[t] absent resets, GAE of the stored rollout ==
oracle GAE).  Parity against the reference's own pipeline lives in tests/test_reference_parity.py.
"""
import numpy as np
import pytest


def _build(N=64, T=16, hidden=(32, 32), use_graph=True, seed=0, obs_norm=True, opt_epochs=2, max_frames=999,
           env_id="SynthHalfCheetah-v0", batch_rows=4):
    import torch
    import torchrl_b200.networks as networks
    import torchrl_b200.policies as policies
    from torchrl_b200.algo import PPO
    from torchrl_b200.collector import VecOnPolicyCollector
    from torchrl_b200.env import get_vec_env
    from torchrl_b200.replay_buffers import OnPolicyReplayBuffer
    from torchrl_b200.utils import NullLogger
    dev = torch.device("cuda:0")
    env = get_vec_env(env_id, {"reward_scale": 1, "obs_norm": obs_norm}, N)
    eval_env = get_vec_env(env_id, {"reward_scale": 1, "obs_norm": obs_norm}, N)
    env.seed(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    buf = OnPolicyReplayBuffer(env_nums=N, max_replay_buffer_size=T * N, time_limit_filter=True)
    net = dict(hidden_shapes=list(hidden), append_hidden_shapes=[], base_type=networks.MLPBase,
               activation_func=torch.nn.Tanh)
    pf = policies.GuassianContPolicyBasicBias(input_shape=env.observation_space.shape[0],
                                              output_shape=env.action_space.shape[0], tanh_action=True, **net)
    vf = networks.Net(input_shape=env.observation_space.shape, output_shape=1, **net)
    col = VecOnPolicyCollector(vf, env=env, eval_env=eval_env, pf=pf, replay_buffer=buf, device=dev,
                               train_render=False, epoch_frames=T * N, max_episode_frames=max_frames,
                               eval_episodes=1, use_cuda_graph=use_graph)
    agent = PPO(pf=pf, vf=vf, plr=3e-4, vlr=3e-4, clip_para=0.2, opt_epochs=opt_epochs, tau=0.95, shuffle=True,
                entropy_coeff=0.005, env=env, replay_buffer=buf, collector=col, logger=NullLogger(), discount=0.99,
                num_epochs=10, batch_size=batch_rows * N, gae=True, device=dev, save_dir=None,
                use_cuda_graph=use_graph)
    return agent, col, buf, env


@pytest.mark.gpu
def test_graph_and_eager_paths_agree():
    import torch
    runs = []
    for use_graph in (False, True):
        agent, col, buf, env = _build(use_graph=use_graph)
        infos = []
        for epoch in range(3):
            agent.current_epoch = epoch
            out = col.train_one_epoch()
            agent.update_per_epoch()
            infos.append((out["train_epoch_reward"], agent._last_infos))
        runs.append((agent, buf, infos))
    (a0, b0, i0), (a1, b1, i1) = runs
    for k in ("obs", "next_obs", "acts", "values", "rewards", "terminals", "time_limits", "advs",
              "estimate_returns", "old_logp"):
        x0, x1 = getattr(b0, "_" + k), getattr(b1, "_" + k)
        if x0.dtype == torch.uint8:
            assert torch.equal(x0, x1), k
        else:   # cuBLAS may pick different algorithms under capture: allow fp32 round-off, not more
            torch.testing.assert_close(x0, x1, rtol=1e-4, atol=1e-5, msg=k)
    torch.testing.assert_close(a0.opt.data, a1.opt.data, rtol=1e-4, atol=1e-6)
    for (r0, u0), (r1, u1) in zip(i0, i1):
        assert abs(r0 - r1) <= 1e-4 * max(1.0, abs(r0)) and len(u0) == len(u1) == 2 * 4
        for d0, d1 in zip(u0, u1):
            assert d0.keys() == d1.keys()
            for k in d0:
                assert abs(d0[k] - d1[k]) <= 1e-3 * max(1.0, abs(d0[k])) or (np.isnan(d0[k]) and np.isnan(d1[k])), k


@pytest.mark.gpu
def test_rollout_semantics_and_gae_vs_oracle():
    import torch
    from oracle import ref_numpy as rn
    agent, col, buf, env = _build(N=32, T=24, use_graph=True, max_frames=10)   # forces 2 timeouts per epoch
    agent.current_epoch = 0
    out = col.train_one_epoch()
    obs, nxt = buf._obs.cpu().numpy(), buf._next_obs.cpu().numpy()
    term, tl = buf._terminals.cpu().numpy(), buf._time_limits.cpu().numpy()
    # terminals carry the collector-level timeout (on_policy.py:141), env time_limit never fires here
    exp_term = np.zeros((24, 32, 1), dtype=np.uint8)
    exp_term[9] = 1
    exp_term[19] = 1
    np.testing.assert_array_equal(term, exp_term)
    assert tl.sum() == 0
    for t in range(23):
        if t in (9, 19):
            # quirk A.1: after a partial reset the collector continues from the RAW reset state
            assert np.abs(obs[t + 1]).max() <= 0.1 + 1e-6
        else:
            np.testing.assert_array_equal(obs[t + 1], nxt[t])
    agent.process_epoch_samples()
    with torch.no_grad():
        lv = agent.vf(buf._next_obs[23]).cpu().numpy() * (1 - term[23])
    ea, er = rn.gae(buf._rewards.cpu().numpy(), buf._values.cpu().numpy(), term, tl, lv, 0.99, 0.95, True)
    np.testing.assert_allclose(buf._advs.cpu().numpy(), ea, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(buf._estimate_returns.cpu().numpy(), er, rtol=1e-4, atol=2e-5)
    # bootstrapped reward on timeout rows: r + gamma * V(next_obs)   (on_policy.py:142-143)
    assert out["train_rewards"] == []            # no env-level `done` -> no finished episodes logged
    assert np.isfinite(out["train_epoch_reward"])


@pytest.mark.gpu
def test_training_improves_and_eval_runs():
    """A few epochs of PPO on the synthetic env: losses finite, eval epoch returns one entry per env."""
    agent, col, buf, env = _build(N=128, T=32, hidden=(64, 64), opt_epochs=4, batch_rows=8)
    for epoch in range(4):
        agent.current_epoch = epoch
        col.train_one_epoch()
        agent.update_per_epoch()
        for info in agent._last_infos:
            assert all(np.isfinite(v) for v in info.values()), info
    col.eval_env._max_episode_steps = 50
    ev = col.eval_one_epoch()
    assert len(ev["eval_rewards"]) == 128 and ev["eval_traj_length"] == 50.0
