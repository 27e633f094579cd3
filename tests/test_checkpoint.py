"""Resume path: a run restored from a checkpoint continues exactly like the run that
wrote it -- parameters, Adam state, normaliser, env state, RNG streams and ring contents.
Both runs execute eagerly (no CUDA graphs) so that the comparison is not blurred by cuBLAS choosing
different algorithms under stream capture.""" 
import numpy as np
import pytest


def _epochs(agent, col, first, n):
    out = []
    for e in range(first, first + n):
        agent.current_epoch = e
        out.append(col.train_one_epoch()["train_epoch_reward"])
        agent.update_per_epoch()
    return out


@pytest.mark.gpu
def test_ppo_resume_continues_identically(tmp_path):
    import torch
    from tests.test_ppo_pipeline import _build
    path = str(tmp_path / "ck.pt")
    agent, col, buf, env = _build(N=32, T=16, seed=3, max_frames=11, use_graph=False)
    _epochs(agent, col, 0, 2)
    agent.save_checkpoint(path)
    want_r = _epochs(agent, col, 2, 2)
    want = agent.opt.data.clone()
    want_mean = env._obs_normalizer._mean.clone()
    want_obs = buf._obs.clone()
    # a differently-seeded fresh agent: everything must come from the file
    agent2, col2, buf2, env2 = _build(N=32, T=16, seed=99, max_frames=11, use_graph=False)
    assert agent2.load_checkpoint(path) == 2
    got_r = _epochs(agent2, col2, 2, 2)
    np.testing.assert_allclose(got_r, want_r, rtol=1e-5)
    torch.testing.assert_close(agent2.opt.data, want, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(env2._obs_normalizer._mean, want_mean, rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(buf2._obs, want_obs, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_sac_resume_restores_ring_and_targets(tmp_path):
    import torch
    from tests.test_offpolicy import _build_offpolicy
    path = str(tmp_path / "ck.pt")
    agent, col, buf, env = _build_offpolicy("sac", seed=1, use_graph=False)
    for e in range(3):
        agent.current_epoch = e
        col.train_one_epoch()
        agent.update_per_epoch()
    agent.save_checkpoint(path)
    agent.current_epoch = 3
    col.train_one_epoch()
    agent.update_per_epoch()
    want, want_t, want_top = agent.opt.data.clone(), agent._target_flat.data.clone(), buf._top
    agent2, col2, buf2, env2 = _build_offpolicy("sac", seed=77, use_graph=False)
    assert agent2.load_checkpoint(path) == 3
    assert buf2._top == (want_top - col2.sample_epoch_frames) % buf2._max_replay_buffer_size
    agent2.current_epoch = 3
    col2.train_one_epoch()
    agent2.update_per_epoch()
    assert buf2._top == want_top
    torch.testing.assert_close(agent2.opt.data, want, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(agent2._target_flat.data, want_t, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(buf2._obs, buf._obs, rtol=1e-5, atol=1e-6)
