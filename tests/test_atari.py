"""BASELINE config 4 pieces: synthetic Atari-shaped pixel env (bit-exact vs oracle/synth_atari.py), the uint8
pixel collector, and DQN / QR-DQN (+ prioritised replay) training on it."""
import numpy as np
import pytest

from oracle import synth_atari as oa


def test_oracle_game_rules():
    lat = {"bx": np.array([10, 40]), "by": np.array([73, 73]), "vx": np.array([1, -1]), "vy": np.array([2, 2]),
           "px": np.array([8, 60])}
    new, r, miss = oa.step_latent(lat, np.array([0, 0]))
    assert list(r) == [1, -1] and list(miss) == [False, True]          # env 0 hits the paddle, env 1 misses
    assert new["vy"][0] == -2 and new["by"][0] == 2 * 74 - 75
    fr = oa.render(new)
    assert fr.shape == (2, 84, 84) and fr.dtype == np.uint8 and fr.max() == 255 and (fr == 200).sum() == 2 * 24


@pytest.mark.gpu
def test_atari_env_bit_exact_vs_oracle():
    import torch
    from torchrl_b200.env import get_vec_env
    N = 37
    env = get_vec_env("SynthAtari-v0", {}, N)
    env.seed(9)
    obs = env.reset().cpu().numpy()
    seeds = ((9 * N + np.arange(N)) & 0xFFFFFFFF).astype(np.uint64)
    episodes = np.zeros(N, dtype=np.uint64)
    lat = oa.reset_latent(seeds, episodes)
    np.testing.assert_array_equal(env.latent.cpu().numpy(), np.stack([lat[k] for k in ("bx", "by", "vx", "vy", "px")], 1))
    stack = np.repeat(oa.render(lat)[:, None], 4, axis=1)
    np.testing.assert_array_equal(obs, stack)
    rs = np.random.RandomState(0)
    elapsed = np.zeros(N, dtype=np.int64)
    n_done = 0
    for t in range(300):
        act = rs.randint(0, 6, N)
        lat, r, miss = oa.step_latent(lat, act)
        elapsed += 1
        stack = np.concatenate([stack[:, 1:], oa.render(lat)[:, None]], axis=1)
        o, rew, dn, info = env.step(torch.from_numpy(act).cuda())
        np.testing.assert_array_equal(o.cpu().numpy(), stack)
        np.testing.assert_array_equal(rew.cpu().numpy()[:, 0], r.astype(np.float32))
        np.testing.assert_array_equal(dn.cpu().numpy()[:, 0], miss)
        assert not info["time_limit"].any()
        if miss.any():                                   # partial reset of the finished envs, like the collectors
            n_done += int(miss.sum())
            episodes[miss] += 1
            fresh = oa.reset_latent(seeds[miss], episodes[miss])
            for k in lat:
                lat[k][miss] = fresh[k]
            stack[miss] = np.repeat(oa.render(fresh)[:, None], 4, axis=1)
            elapsed[miss] = 0
            o = env.partial_reset(dn.squeeze(-1))
            np.testing.assert_array_equal(o.cpu().numpy(), stack)
    assert n_done > 0
    f = env.to_float(env.obs)
    np.testing.assert_allclose(f.cpu().numpy(), stack.astype(np.float32) / 255.0, rtol=1e-6)


def _build_pixel(kind, N=16, rows=24, use_graph=True, prioritized=False, dedup=False, max_frames=30):
    import torch
    import torch.nn as nn
    import torchrl_b200.networks as networks
    import torchrl_b200.policies as policies
    from torchrl_b200.algo import DQN, QRDQN
    from torchrl_b200.collector import PixelVecCollector
    from torchrl_b200.env import get_vec_env
    from torchrl_b200.replay_buffers import BaseReplayBuffer, MemoryEfficientReplayBuffer, PrioritizedReplayBuffer
    from torchrl_b200.utils import NullLogger
    dev = torch.device("cuda:0")
    env = get_vec_env("SynthAtari-v0", {}, N)
    eval_env = get_vec_env("SynthAtari-v0", {}, N)
    env.seed(0); torch.manual_seed(0); np.random.seed(0)
    Q = 11 if kind == "qrdqn" else 1
    cls = MemoryEfficientReplayBuffer if dedup else (PrioritizedReplayBuffer if prioritized else BaseReplayBuffer)
    buf = cls(env_nums=N, max_replay_buffer_size=rows * N)
    qf = networks.Net(input_shape=(4, 84, 84), output_shape=6 * Q,
                      hidden_shapes=[[16, [8, 8], [4, 4], [0, 0]], [32, [4, 4], [2, 2], [0, 0]]],
                      append_hidden_shapes=[64], base_type=networks.CNNBase, activation_func=nn.ReLU)
    if kind == "qrdqn":
        pf = policies.EpsilonGreedyQRDQNDiscretePolicy(quantile_num=Q, qf=qf, start_epsilon=0.5, end_epsilon=0.1,
                                                       decay_frames=100, action_shape=6)
    else:
        pf = policies.EpsilonGreedyDQNDiscretePolicy(qf=qf, start_epsilon=0.5, end_epsilon=0.1, decay_frames=100,
                                                     action_shape=6)
    col = PixelVecCollector(env=env, eval_env=eval_env, pf=pf, replay_buffer=buf, device=dev, epoch_frames=8 * N,
                            max_episode_frames=max_frames, use_cuda_graph=use_graph)
    common = dict(qf=qf, pf=pf, qlr=1e-3, optimizer_info={"eps": 1e-4}, env=env, replay_buffer=buf, collector=col,
                  logger=NullLogger(), discount=0.99, batch_size=4 * N, device=dev, save_dir=None, opt_times=4,
                  use_soft_update=False, target_hard_update_period=3, pretrain_epochs=1, num_epochs=2,
                  use_cuda_graph=use_graph)
    agent = QRDQN(quantile_num=Q, **common) if kind == "qrdqn" else DQN(**common)
    return agent, col, buf, env


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["dqn", "qrdqn"])
def test_pixel_collector_and_dqn_training(kind):
    import torch
    agent, col, buf, env = _build_pixel(kind)
    agent.pretrain()
    for epoch in range(2):
        agent.current_epoch = epoch
        col.train_one_epoch()
        agent.update_per_epoch()
        for info in agent._last_infos:
            assert np.isfinite(info["Training/qf_loss"]) and np.isfinite(info["q_s_a"])
    assert buf._obs.dtype == torch.uint8 and buf._obs.shape == (24, 16, 4, 84, 84) and buf._size == 24
    obs, nxt = buf._obs.cpu().numpy(), buf._next_obs.cpu().numpy()
    term = buf._terminals.cpu().numpy()[..., 0].astype(bool)
    for t in range(23):
        # frame-stack property of every stored transition, and continuity where no reset happened
        np.testing.assert_array_equal(nxt[t][:, :3], obs[t][:, 1:])
        cont = ~term[t]
        same = (obs[t + 1][cont] == nxt[t][cont]).reshape(cont.sum(), -1).all(axis=1)
        assert same.mean() > 0.9                      # the rest were reset by the 30-frame collector timeout
    acts = buf._acts.cpu().numpy()
    assert acts.min() >= 0 and acts.max() <= 5 and np.all(acts == np.round(acts))
    assert 0.1 <= agent.pf.epsilon < 0.5             # schedule advanced on the host, fed through a device scalar


@pytest.mark.gpu
def test_qrdqn_with_prioritized_replay_on_pixels():
    """Config 4 end to end: QR-DQN heads + prioritised row sampling with importance weights on uint8 frames."""
    import torch
    agent, col, buf, env = _build_pixel("qrdqn", prioritized=True, use_graph=True)
    agent.pretrain()
    col.train_one_epoch()
    assert torch.all(buf._priorities[:16] == 1.0)            # collector-written rows entered with the max priority
    np.random.seed(3)
    batch = buf.random_batch(4 * 16, agent.sample_key)
    assert batch["obs"].dtype == torch.uint8 and batch["weights"].shape == (64, 1)
    before = buf._priorities.clone()
    agent.update_per_epoch()                                 # prioritised path: sample, weighted loss, new priorities
    assert len(agent._last_infos) == 4 and all(np.isfinite(i["Training/qf_loss"]) for i in agent._last_infos)
    assert not torch.equal(before, buf._priorities) and float(buf._priorities[:16].min()) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("use_graph", [False, True])
def test_frame_deduplicated_ring_reconstructs_every_stack_exactly(use_graph):
    """MemoryEfficientReplayBuffer (one frame of obs + one of next_obs per row, csrc/frames.cu) against the full-stack
    ring filled by an identical run: after the ring has wrapped several times, EVERY row -- including the oldest ones,
    whose older frames the ring has overwritten, and rows right after resets (collector time-outs every 5 frames and
    missed balls) -- gathers to the same float32 stacks, bit for bit; the pixel storage is 4x smaller."""
    import torch
    runs = []
    for dedup in (False, True):
        agent, col, buf, env = _build_pixel("dqn", N=8, rows=10, use_graph=use_graph, dedup=dedup, max_frames=5)
        for _ in range(37):                                  # 3.7 laps of the 10-row ring
            col._step()
        runs.append((buf, env))
    (full, env0), (ded, env1) = runs
    assert full._size == ded._size == 10 and full._top == ded._top == 7
    idx = torch.arange(10, device="cuda")
    a = full.gather_rows(idx, ["obs", "next_obs", "acts", "rewards", "terminals"])
    b = ded.gather_rows(idx, ["obs", "next_obs", "acts", "rewards", "terminals"])
    for k in ("acts", "rewards", "terminals"):
        assert torch.equal(a[k], b[k]), k
    for k in ("obs", "next_obs"):
        assert a[k].dtype == torch.uint8 and b[k].dtype == torch.float32
        want = env0.to_float(a[k].contiguous())
        assert torch.equal(b[k], want), (k, (b[k] != want).float().mean().item())
    assert int(ded._age.max()) == 3 and int(ded._age.min()) == 0      # both fresh and mid-episode rows are present
    stacks = full._obs.numel() + full._next_obs.numel()
    assert ded.stored_frame_bytes() < 0.35 * stacks          # 2 of 8 frames per row (+ a 3-frame history per env)
    # device-position gather (what the captured update graphs use)
    pos = torch.tensor([1], dtype=torch.int32, device="cuda")
    table = torch.tensor([3, 4, 9, 0, 7, 2], dtype=torch.int64, device="cuda")
    c = ded.gather_rows(table, ["obs"], pos_ptr=pos, rows=3)
    want = env0.to_float(full.gather_rows(torch.tensor([0, 7, 2], device="cuda"), ["obs"])["obs"].contiguous())
    assert torch.equal(c["obs"], want)


@pytest.mark.gpu
def test_dqn_trains_on_the_frame_deduplicated_ring():
    agent, col, buf, env = _build_pixel("qrdqn", dedup=True)
    agent.pretrain()
    for epoch in range(2):
        agent.current_epoch = epoch
        col.train_one_epoch()
        agent.update_per_epoch()
        assert all(np.isfinite(i["Training/qf_loss"]) for i in agent._last_infos)
