"""K3 (action sampling), K7/K9/K4 (row gather / ring write / uniform row sampling) parity."""
import numpy as np
import pytest

from oracle import ref_numpy as rn


def test_random_batch_indices_known_answer():
    """SURVEY.md section 8(a) K9 [probed on the reference]: np.random.seed(0), size 976, 4 rows."""
    np.random.seed(0)
    np.testing.assert_array_equal(rn.uniform_row_indices(976, 4096, 1024), [684, 559, 629, 192])


@pytest.mark.gpu
def test_sample_with_given_noise_matches_formula():
    import torch
    from torchrl_b200 import ops
    rs = np.random.RandomState(0)
    M, a = 300, 6
    mean = rs.randn(M, a).astype(np.float32)
    ls = (0.3 * rs.randn(a) - 1.0).astype(np.float32)
    eps = rs.randn(M, a).astype(np.float32)
    out = ops.tanh_gaussian_sample(torch.from_numpy(mean).cuda(), torch.from_numpy(ls).cuda(),
                                   eps=torch.from_numpy(eps).cuda(), want_log_prob=True, want_pre_tanh=True)
    z = mean.astype(np.float64) + np.exp(ls.astype(np.float64)) * eps
    act = np.tanh(z)
    np.testing.assert_allclose(out["pre_tanh"].cpu().numpy(), z, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(out["action"].cpu().numpy(), act, rtol=1e-5, atol=1e-6)
    lp = rn.tanh_normal_log_prob(act, mean, np.exp(ls)[None, :], pre_tanh=z).sum(-1, keepdims=True)
    np.testing.assert_allclose(out["log_prob"].cpu().numpy(), lp, rtol=1e-4, atol=1e-4)
    # per-row log_std (SAC-style policy head)
    ls2 = (0.3 * rs.randn(M, a) - 1.0).astype(np.float32)
    out2 = ops.tanh_gaussian_sample(torch.from_numpy(mean).cuda(), torch.from_numpy(ls2).cuda(),
                                    eps=torch.from_numpy(eps).cuda(), want_log_prob=True)
    z2 = mean + np.exp(ls2.astype(np.float64)) * eps
    np.testing.assert_allclose(out2["action"].cpu().numpy(), np.tanh(z2), rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_philox_noise_statistics_and_determinism():
    import torch
    from torchrl_b200 import ops
    from torchrl_b200.policies.continuous_policy import _DeviceRng
    rng = _DeviceRng()
    torch.manual_seed(11)
    rng.ensure("cuda")
    M, a = 200000, 6
    mean = torch.zeros(M, a, device="cuda")
    ls = torch.zeros(a, device="cuda")
    o1 = ops.tanh_gaussian_sample(mean, ls, tanh_action=False, rng=rng)["action"].clone()
    o1b = ops.tanh_gaussian_sample(mean, ls, tanh_action=False, rng=rng)["action"].clone()
    torch.testing.assert_close(o1, o1b, rtol=0, atol=0)                 # same counter -> same noise
    ops.counter_advance(rng.counter)
    o2 = ops.tanh_gaussian_sample(mean, ls, tanh_action=False, rng=rng)["action"]
    assert not torch.equal(o1, o2)
    x = o1.double()
    assert abs(x.mean().item()) < 5e-3 and abs(x.var().item() - 1.0) < 1e-2
    assert abs((x ** 3).mean().item()) < 2e-2 and abs((x ** 4).mean().item() - 3.0) < 6e-2
    c = torch.corrcoef(x.T)
    assert (c - torch.eye(a, device="cuda", dtype=torch.float64)).abs().max().item() < 1e-2


@pytest.mark.gpu
def test_sample_backward_matches_autograd_of_reference_formula():
    """d(action, log_prob)/d(mean, log_std) vs torch autograd through the reference's expressions
    (distribution.py:33-45, 60-76; continuous_policy.py:109-121)."""
    import torch
    from torchrl_b200.policies import distribution as D
    torch.manual_seed(0)
    M, a = 64, 8
    mean = torch.randn(M, a, device="cuda", dtype=torch.float64)
    ls = (0.2 * torch.randn(M, a, device="cuda", dtype=torch.float64) - 0.5)
    eps = torch.randn(M, a, device="cuda", dtype=torch.float64)
    w_a = torch.randn(M, a, device="cuda", dtype=torch.float64)
    w_l = torch.randn(M, 1, device="cuda", dtype=torch.float64)
    m_ref, l_ref = mean.clone().requires_grad_(), ls.clone().requires_grad_()
    std = l_ref.exp()
    z = m_ref + std * eps
    act = torch.tanh(z)
    lp = (torch.distributions.Normal(m_ref, std).log_prob(z) - torch.log(1 - act * act + 1e-6)).sum(-1, keepdim=True)
    ((act * w_a).sum() + (lp * w_l).sum()).backward()
    m32, l32 = mean.float().requires_grad_(), ls.float().requires_grad_()
    act2, lp2, _ = D._SampleFn.apply(m32, l32, eps.float(), True, True, None)
    ((act2 * w_a.float()).sum() + (lp2 * w_l.float()).sum()).backward()
    torch.testing.assert_close(act2.double(), act.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(lp2.double(), lp.detach(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(m32.grad.double(), m_ref.grad, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(l32.grad.double(), l_ref.grad, rtol=2e-4, atol=2e-4)


@pytest.mark.gpu
def test_row_gather_and_ring_write_bit_exact():
    import torch
    from torchrl_b200.replay_buffers import BaseReplayBuffer
    N, o, a, T = 24, 17, 6, 10
    np.random.seed(0)
    buf = BaseReplayBuffer(T * N, env_nums=N, device="cuda")
    host = {k: [] for k in ("obs", "acts", "rewards", "terminals")}
    rs = np.random.RandomState(3)
    for t in range(T + 3):                                   # wraps the ring
        s = {"obs": rs.randn(N, o).astype(np.float32), "acts": rs.randn(N, a).astype(np.float32),
             "rewards": rs.randn(N, 1).astype(np.float32), "terminals": rs.rand(N, 1) < 0.3}
        buf.add_sample({k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in s.items()})
        for k in host:
            host[k].append(s[k])
    assert buf._top == 3 and buf._size == T and buf.num_steps_can_sample() == T
    assert buf._top_dev.item() == 3 and buf._size_dev.item() == T
    ring = {k: np.stack(host[k][3:T + 3]) for k in host}     # rows 3..12 live at (t % T)
    for k in host:
        stored = getattr(buf, "_" + k).cpu().numpy()
        for t in range(3, T + 3):
            np.testing.assert_array_equal(stored[t % T], host[k][t].astype(stored.dtype))
    # random_batch: same indices as the reference's np.random.randint stream, same gathered bytes
    np.random.seed(42)
    exp_idx = np.random.randint(0, T, 4)
    np.random.seed(42)
    batch = buf.random_batch(4 * N, ["obs", "acts", "rewards", "terminals"])
    for k in host:
        stored = getattr(buf, "_" + k).cpu().numpy()
        np.testing.assert_array_equal(batch[k].cpu().numpy(), rn.gather_rows(stored, exp_idx))
    assert batch["obs"].shape == (4 * N, o) and batch["terminals"].dtype == torch.uint8


@pytest.mark.gpu
def test_vec_stats():
    import torch
    from torchrl_b200 import ops
    x = torch.randn(16384, device="cuda") * 3 + 1
    st = ops.vec_stats(x).cpu().numpy()
    np.testing.assert_allclose(st, [x.mean().item(), x.std().item(), x.max().item(), x.min().item()], rtol=1e-5)


def test_per_oracle_properties():
    """CPU: the NumPy definition of prioritised row sampling behaves like proportional sampling."""
    rs = np.random.RandomState(0)
    prio = rs.rand(500).astype(np.float32) + 0.01
    counts = np.zeros(500)
    for _ in range(400):
        idx, w = rn.per_sample(prio, 500, rs.rand(64), 0.4)
        assert idx.min() >= 0 and idx.max() < 500 and np.all(np.diff(idx) >= 0)     # stratified => sorted
        assert w.max() <= 1.0 + 1e-12
        np.add.at(counts, idx, 1)
    corr = np.corrcoef(counts, prio)[0, 1]
    assert corr > 0.9


@pytest.mark.gpu
@pytest.mark.parametrize("size,b", [(1953, 4), (976, 64), (7, 3), (4096, 256)])
def test_prioritized_sampling_matches_oracle(size, b):
    import torch
    from torchrl_b200 import ops
    rs = np.random.RandomState(size)
    rows = max(size, 8)
    prio = (rs.rand(rows).astype(np.float32) + 1e-3) ** 0.6
    u = rs.rand(b)
    idx, w = ops.per_sample(torch.from_numpy(prio).cuda(), size, torch.from_numpy(u).cuda(), 0.4)
    e_idx, e_w = rn.per_sample(prio, size, u, 0.4)
    np.testing.assert_array_equal(idx.cpu().numpy(), e_idx)          # bit-exact indices
    np.testing.assert_allclose(w.cpu().numpy(), e_w, rtol=1e-5)
    # priority update from TD errors + running maximum
    td = rs.randn(b, 16).astype(np.float32)
    p_dev = torch.from_numpy(prio.copy()).cuda()
    mx = torch.ones(1, device="cuda")
    uniq = np.unique(e_idx, return_index=True)[1]                     # duplicates: last-writer unordered on GPU
    ops.per_update(p_dev, idx, torch.from_numpy(td).cuda(), 0.6, 1e-6, mx)
    p_ref = prio.copy()
    e_max = rn.per_update(p_ref, e_idx, td, 0.6, 1e-6, 1.0)
    sel = e_idx[np.array([k for k in range(b) if (e_idx == e_idx[k]).sum() == 1], dtype=int)] if b > 1 else e_idx
    np.testing.assert_allclose(p_dev.cpu().numpy()[sel], p_ref[sel], rtol=1e-5)
    assert abs(mx.item() - e_max) < 1e-5 * max(1.0, e_max)


@pytest.mark.gpu
def test_prioritized_buffer_api():
    import torch
    from torchrl_b200.replay_buffers import PrioritizedReplayBuffer
    N, o, T = 8, 5, 32
    buf = PrioritizedReplayBuffer(T * N, env_nums=N, device="cuda", alpha=0.6, beta=0.4)
    for t in range(20):
        buf.add_sample({"obs": torch.full((N, o), float(t), device="cuda"),
                        "rewards": torch.zeros(N, 1, device="cuda"),
                        "terminals": torch.zeros(N, 1, dtype=torch.bool, device="cuda")})
    assert torch.all(buf._priorities[:20] == 1.0) and torch.all(buf._priorities[20:] == 0.0)
    np.random.seed(1)
    batch = buf.random_batch(4 * N, ["obs", "rewards"])
    assert batch["obs"].shape == (4 * N, o) and batch["weights"].shape == (4 * N, 1) and batch["indices"].shape == (4,)
    rows = batch["indices"].cpu().numpy()
    np.testing.assert_array_equal(batch["obs"].cpu().numpy()[::N, 0], rows.astype(np.float32))
    assert torch.allclose(batch["weights"], torch.ones_like(batch["weights"]))   # equal priorities -> weights 1
    td = torch.zeros(4 * N, 1, device="cuda")
    td[:N] = 10.0                                                               # first sampled row: large TD error
    buf.update_priorities(batch["indices"], td)
    assert buf._priorities[rows[0]].item() > 3.9 and abs(buf._max_prio.item() - (10 + 1e-6) ** 0.6) < 1e-4
    np.random.seed(2)
    hits = 0
    for _ in range(50):
        hits += int((buf.random_batch(4 * N, ["obs"])["indices"] == int(rows[0])).any().item())
    assert hits >= 25                                                           # high-priority row is drawn often


@pytest.mark.gpu
def test_global_vec_stats_single_rank_matches_vec_stats():
    """The K12 statistics path (raw moments -> gather -> combine) equals the single-launch statistics."""
    import torch
    from torchrl_b200 import ops
    from torchrl_b200.distributed import DataParallelContext
    ctx = DataParallelContext(backend="nccl")
    x = torch.randn(16384, device="cuda") * 2 - 0.5
    out = torch.zeros(4, device="cuda")
    ctx.global_vec_stats(x, out)
    torch.testing.assert_close(out, ops.vec_stats(x), rtol=1e-6, atol=1e-7)
