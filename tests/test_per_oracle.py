"""The prioritised-replay oracle (oracle/ref_numpy.per_sample / per_update) against the published algorithm it
restates: Schaul, Quan, Antonoglou, Silver, "Prioritized Experience Replay", ICLR 2016 (arXiv:1511.05952),
proportional variant.  The reference repository has no prioritised replay, so this is what pins the oracle (the CUDA
kernels are pinned bit-exactly to the oracle by tests/test_sample_gather.py)."""
import numpy as np

from oracle import ref_numpy as rn


def test_sampling_frequencies_follow_eq_1():
    """eq. (1): P(i) = p_i^alpha / sum_k p_k^alpha -- empirical frequencies over many stratified minibatches."""
    rs = np.random.RandomState(0)
    size, b, draws = 64, 8, 40000
    prio = (rs.rand(size).astype(np.float32) + 0.05) ** np.float32(0.6)
    counts = np.zeros(size)
    for _ in range(draws):
        idx, _ = rn.per_sample(prio, size, rs.rand(b), 0.4)
        np.add.at(counts, idx, 1)
    p = prio.astype(np.float64) / prio.astype(np.float64).sum()
    freq = counts / counts.sum()
    # binomial standard error of each frequency
    se = np.sqrt(p * (1 - p) / counts.sum())
    assert np.all(np.abs(freq - p) < 6 * se + 1e-4), np.abs(freq - p).max()


def test_one_draw_per_stratum_appendix_b21():
    """Appendix B.2.1: the range [0, p_total] is divided into b equal ranges, one value sampled in each."""
    rs = np.random.RandomState(1)
    size, b = 100, 10
    prio = rs.rand(size).astype(np.float32) + 0.01
    pre = np.cumsum(prio.astype(np.float64))
    for _ in range(200):
        idx, _ = rn.per_sample(prio, size, rs.rand(b), 0.5)
        lo = np.concatenate([[0.0], pre[:-1]])[idx]
        hi = pre[idx]
        for k in range(b):      # the retrieved row's cumulative interval intersects stratum k
            s0, s1 = k / b * pre[-1], (k + 1) / b * pre[-1]
            assert hi[k] > s0 and lo[k] < s1, (k, lo[k], hi[k], s0, s1)
        assert np.all(np.diff(idx) >= 0)


def test_importance_weights_section_3_4():
    """w_i = (N P(i))^-beta / max_i w_i with max_i w_i attained at the minimum-priority row."""
    rs = np.random.RandomState(2)
    size, b, beta = 50, 16, 0.7
    prio = rs.rand(size).astype(np.float32) + 0.1
    idx, w = rn.per_sample(prio, size, rs.rand(b), beta)
    P = prio.astype(np.float64) / prio.astype(np.float64).sum()
    expect = (size * P[idx]) ** (-beta) / ((size * P.min()) ** (-beta))
    np.testing.assert_allclose(w, expect, rtol=1e-12)
    assert w.max() <= 1.0 + 1e-12


def test_priority_update_algorithm_1():
    """p_i = (|delta_i| + eps)^alpha per sampled row (|delta| = mean over the row's transitions); running max kept."""
    rs = np.random.RandomState(3)
    prio = np.ones(10, dtype=np.float32)
    idx = np.array([2, 7])
    td = rs.randn(2, 5)
    mx = rn.per_update(prio, idx, td, 0.6, 1e-6, 1.0)
    expect = (np.abs(td).mean(axis=1) + 1e-6) ** 0.6
    np.testing.assert_allclose(prio[idx], expect, rtol=1e-6)
    assert mx == max(1.0, float(prio[idx].max()))
