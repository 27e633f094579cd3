"""GPU product vs the CPU oracle port (itself pinned bit-for-bit to the unmodified reference by
tests/test_oracle_vs_reference.py) on the SAME seeds and the SAME exploration noise.

North-star bar: "within stated fp32 tolerance for GAE/returns/losses; bit-exact for replay
indexing".  Stated tolerances (fp32 device maths vs float64 NumPy buffers / torch-CPU nets):
  rollout tensors after one 16-step epoch ........ rtol 1e-4, atol 1e-4
  advantages / returns ........................... rtol 1e-3, atol 2e-4
  logged losses of the first update pass ......... rtol 2e-3, atol 2e-4
  parameters after one epoch of updates .......... atol 2e-4
The minibatch row order is bit-exact (same np.random.permutation stream, checked via the infos
being comparable at all: a different order would pair different minibatches).
"""
import numpy as np
import pytest


@pytest.mark.gpu
@pytest.mark.parametrize("max_frames,env_id,N,T", [(999, "SynthHalfCheetah-v0", 8, 16), (7, "SynthHalfCheetah-v0", 8, 16),
                                                   (999, "SynthHalfCheetahTerm-v0", 64, 32)])
def test_ppo_epoch_matches_reference_port(max_frames, env_id, N, T):
    """Lock-step episodes, collector time-outs every 7 frames (bootstrap + quirk A.1), and the variant with
    state-dependent early termination: episodes desynchronise, some envs reset on most steps, the extra V(next_obs)
    forward of collector/on_policy.py:132-148 runs on every step."""
    import torch
    from oracle import ref_port
    from torchrl_b200.policies import set_noise_mode
    from tests.test_ppo_pipeline import _build
    hidden, rows, oe, seed = (32, 32), 4, 2, 5
    keys = ("obs", "next_obs", "acts", "values", "rewards", "terminals", "time_limits")
    # ---- CPU oracle port ------------------------------------------------------------------
    torch.set_num_threads(4)
    penv, pcol, pagent = ref_port.build_ppo(env_id=env_id, env_nums=N, horizon=T, hidden=hidden, batch_rows=rows,
                                            opt_epochs=oe, seed=seed, max_episode_frames=max_frames)
    pagent.current_epoch = 0
    p_out = pcol.train_one_epoch()
    if env_id.endswith("Term-v0"):
        assert pagent.buffer.data["terminals"].sum() >= 3, "the variant must actually terminate episodes early"
    roll = {k: pagent.buffer.data[k].copy() for k in keys}
    pagent.update_per_epoch()
    # ---- device product, reference noise --------------------------------------------------
    set_noise_mode("reference_cpu")
    try:
        for use_graph in (False, True):
            agent, col, buf, env = _build(N=N, T=T, hidden=hidden, use_graph=use_graph, seed=seed, opt_epochs=oe,
                                          max_frames=max_frames, batch_rows=rows, env_id=env_id)
            agent.current_epoch = 0
            out = col.train_one_epoch()
            assert abs(out["train_epoch_reward"] - p_out["train_epoch_reward"]) < 1e-3
            for k in keys:
                got = getattr(buf, "_" + k).cpu().numpy().astype(np.float64)
                np.testing.assert_allclose(got, roll[k].reshape(got.shape), rtol=1e-4, atol=1e-4, err_msg=k)
            agent.update_per_epoch()
            np.testing.assert_allclose(buf._advs.cpu().numpy(), pagent.buffer.data["advs"], rtol=1e-3, atol=2e-4)
            np.testing.assert_allclose(buf._estimate_returns.cpu().numpy(), pagent.buffer.data["estimate_returns"],
                                       rtol=1e-3, atol=2e-4)
            infos = agent._last_infos
            assert len(infos) == len(pagent.infos)
            for u in range(T // rows):            # first pass: parameters have barely moved
                for k, v in pagent.infos[u].items():
                    assert abs(infos[u][k] - v) <= 2e-3 * abs(v) + 2e-4, (u, k, infos[u][k], v)
            mine = torch.cat([p.detach().reshape(-1) for p in list(agent.pf.mean_params()) +
                              list(agent.vf.parameters())]).cpu().numpy()
            ref = torch.cat([p.detach().reshape(-1) for p in list(pagent.pf.net.parameters()) +
                             list(pagent.vf.parameters())]).numpy()
            np.testing.assert_allclose(mine, ref, atol=2e-4)
            nrm = env._obs_normalizer
            np.testing.assert_allclose(nrm._mean.cpu().numpy(), penv.norm.mean, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(nrm._var.cpu().numpy(), penv.norm.var, rtol=1e-5, atol=1e-6)
    finally:
        set_noise_mode("philox")
