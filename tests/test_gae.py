"""K6 parity: GAE / discounted-return scan.

CPU part: the NumPy oracle vs the golden vectors produced by the unmodified reference
(oracle/make_golden.py).  GPU part: the CUDA kernels through the C ABI vs the oracle /
golden vectors.  Tolerance (stated per north star: "within stated fp32 tolerance for
GAE/returns"): the reference computes in float64, the kernel in fp32 with chunk-level
re-association -> rtol 1e-4, atol 2e-5 on advantages and returns.
"""
import numpy as np
import pytest

from oracle import ref_numpy as rn

RTOL, ATOL = 1e-4, 2e-5
KEYS = ("values", "rewards", "terminals", "time_limits", "last_value")


def _case(g, ci):
    return tuple(g["c%d_%s" % (ci, k)] for k in KEYS)


def test_oracle_matches_reference_golden(golden_dir):
    g = np.load(golden_dir + "/gae.npz")
    gamma, tau = g["gamma_tau"]
    for ci in range(int(g["ncases"])):
        v, r, t, tl, lv = _case(g, ci)
        for f in (0, 1):
            a, ret = rn.gae(r, v, t, tl, lv, gamma, tau, bool(f))
            np.testing.assert_array_equal(a, g["c%d_f%d_gae_advs" % (ci, f)])
            np.testing.assert_array_equal(ret, g["c%d_f%d_gae_rets" % (ci, f)])
            a, ret = rn.discount_return(r, v, t, tl, lv, gamma, bool(f))
            np.testing.assert_array_equal(a, g["c%d_f%d_disc_advs" % (ci, f)])
            np.testing.assert_array_equal(ret, g["c%d_f%d_disc_rets" % (ci, f)])


def _to_dev(v, r, t, tl, lv):
    import torch
    d = "cuda"
    return (torch.tensor(r, dtype=torch.float32, device=d), torch.tensor(v, dtype=torch.float32, device=d),
            torch.tensor(t.astype(np.uint8), device=d), torch.tensor(tl.astype(np.uint8), device=d),
            torch.tensor(lv, dtype=torch.float32, device=d))


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_gae_kernel_vs_golden(golden_dir, variant):
    from torchrl_b200 import ops
    g = np.load(golden_dir + "/gae.npz")
    gamma, tau = (float(x) for x in g["gamma_tau"])
    for ci in range(int(g["ncases"])):
        v, r, t, tl, lv = _case(g, ci)
        R, V, Tm, TL, LV = _to_dev(v, r, t, tl, lv)
        for f in (0, 1):
            a, ret = ops.gae_scan(R, V, Tm, TL, LV, gamma, tau, bool(f), variant=variant)
            np.testing.assert_allclose(a.cpu().numpy(), g["c%d_f%d_gae_advs" % (ci, f)], rtol=RTOL, atol=ATOL)
            np.testing.assert_allclose(ret.cpu().numpy(), g["c%d_f%d_gae_rets" % (ci, f)], rtol=RTOL, atol=ATOL)
            a, ret = ops.discount_return(R, V, Tm, TL, LV, gamma, bool(f), variant=variant)
            np.testing.assert_allclose(a.cpu().numpy(), g["c%d_f%d_disc_advs" % (ci, f)], rtol=RTOL, atol=ATOL)
            np.testing.assert_allclose(ret.cpu().numpy(), g["c%d_f%d_disc_rets" % (ci, f)], rtol=RTOL, atol=ATOL)


@pytest.mark.gpu
@pytest.mark.parametrize("T,N", [(128, 4096), (1000, 1024), (3, 7), (129, 4100), (2048, 64), (5, 40000)])
def test_gae_kernel_vs_oracle_sizes(T, N):
    """Oracle-sized cases incl. ragged T (not a multiple of the chunk) and N (not of 4/32)."""
    from torchrl_b200 import ops
    from oracle.make_golden import gae_inputs
    v, r, t, tl, lv = gae_inputs(T, N, seed=T + N, p_term=0.02, p_tl=0.01)
    R, V, Tm, TL, LV = _to_dev(v, r, t, tl, lv)
    for f in (True, False):
        ea, er = rn.gae(r, v, t, tl, lv, 0.99, 0.95, f)
        da, dr = rn.discount_return(r, v, t, tl, lv, 0.99, f)
        for variant in (1, 2, 3):
            a, ret = ops.gae_scan(R, V, Tm, TL, LV, 0.99, 0.95, f, variant=variant)
            np.testing.assert_allclose(a.cpu().numpy(), ea, rtol=RTOL, atol=ATOL)
            np.testing.assert_allclose(ret.cpu().numpy(), er, rtol=RTOL, atol=ATOL)
            a, ret = ops.discount_return(R, V, Tm, TL, LV, 0.99, f, variant=variant)
            np.testing.assert_allclose(a.cpu().numpy(), da, rtol=RTOL, atol=3e-5)
            np.testing.assert_allclose(ret.cpu().numpy(), dr, rtol=RTOL, atol=3e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("T,N", [(128, 4096), (1000, 1024), (3, 128), (64, 128), (65, 384), (129, 256), (191, 37888)])
def test_gae_tma_kernel_vs_oracle(T, N):
    """The persistent TMA-staged scan (variant 4, csrc/gae_tma.cu): one tile, several tiles, a ragged earliest tile
    (T not a multiple of 64: the box starts below row 0), more env groups than SMs -- GAE and discounted returns,
    filter on and off, against the float64 oracle."""
    from torchrl_b200 import ops
    from oracle.make_golden import gae_inputs
    v, r, t, tl, lv = gae_inputs(T, N, seed=7 * T + N, p_term=0.02, p_tl=0.01)
    R, V, Tm, TL, LV = _to_dev(v, r, t, tl, lv)
    for f in (True, False):
        ea, er = rn.gae(r, v, t, tl, lv, 0.99, 0.95, f)
        da, dr = rn.discount_return(r, v, t, tl, lv, 0.99, f)
        a, ret = ops.gae_scan(R, V, Tm, TL, LV, 0.99, 0.95, f, variant=4)
        np.testing.assert_allclose(a.cpu().numpy(), ea, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(ret.cpu().numpy(), er, rtol=RTOL, atol=ATOL)
        a, ret = ops.discount_return(R, V, Tm, TL, LV, 0.99, f, variant=4)
        np.testing.assert_allclose(a.cpu().numpy(), da, rtol=RTOL, atol=3e-5)
        np.testing.assert_allclose(ret.cpu().numpy(), dr, rtol=RTOL, atol=3e-5)


@pytest.mark.gpu
def test_gae_tma_rejects_ragged_env_count():
    import torch
    from torchrl_b200 import ops
    z = torch.zeros(4, 100, device="cuda")
    f = torch.zeros(4, 100, device="cuda", dtype=torch.uint8)
    with pytest.raises(RuntimeError):
        ops.gae_scan(z, z, f, f, torch.zeros(100, device="cuda"), 0.99, 0.95, True, variant=4)


@pytest.mark.gpu
def test_gae_full_size_properties():
    """BASELINE size (T=128, N=2**20: 2.4 GB of traffic) through size-independent properties:
    (1) the TMA-staged (what variant 1 picks at this size), vectorised, scalar and serial kernels agree;
    (2) linearity: GAE is linear in (rewards, values, last_value) for fixed flags; (3) an all-terminal rollout
    gives adv = r - V."""
    import torch
    from torchrl_b200 import ops
    T, N = 128, 1 << 20
    gen = torch.Generator(device="cuda").manual_seed(0)
    R = torch.randn(T, N, device="cuda", generator=gen)
    V = torch.randn(T, N, device="cuda", generator=gen)
    Tm = (torch.rand(T, N, device="cuda", generator=gen) < 0.01).to(torch.uint8)
    TL = (torch.rand(T, N, device="cuda", generator=gen) < 0.005).to(torch.uint8)
    LV = torch.randn(N, device="cuda", generator=gen)
    a1, r1 = ops.gae_scan(R, V, Tm, TL, LV, 0.99, 0.95, True, variant=1)
    a0, r0 = ops.gae_scan(R, V, Tm, TL, LV, 0.99, 0.95, True, variant=0)
    a3, r3 = ops.gae_scan(R, V, Tm, TL, LV, 0.99, 0.95, True, variant=3)
    a2v, r2v = ops.gae_scan(R, V, Tm, TL, LV, 0.99, 0.95, True, variant=2)
    a4v, r4v = ops.gae_scan(R, V, Tm, TL, LV, 0.99, 0.95, True, variant=4)
    torch.testing.assert_close(a1, a0, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(r1, r0, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(a2v, a3, rtol=0, atol=0)  # same association, different vector width
    torch.testing.assert_close(a1, a4v, rtol=0, atol=0)  # variant 1 IS the TMA kernel at this size
    torch.testing.assert_close(a4v, a3, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(r4v, r3, rtol=RTOL, atol=ATOL)
    a2, _ = ops.gae_scan(2 * R, 2 * V, Tm, TL, 2 * LV, 0.99, 0.95, True)
    torch.testing.assert_close(a2, 2 * a1, rtol=1e-6, atol=1e-6)
    ones = torch.ones_like(Tm)
    a4, r4 = ops.gae_scan(R, V, ones, torch.zeros_like(TL), LV, 0.99, 0.95, True)
    torch.testing.assert_close(a4, R - V, rtol=0, atol=0)
    torch.testing.assert_close(r4, (R - V) + V, rtol=0, atol=0)


@pytest.mark.gpu
def test_gae_empty_and_errors():
    import torch
    from torchrl_b200 import ops
    e = torch.empty(0, 8, device="cuda")
    f = torch.empty(0, 8, device="cuda", dtype=torch.uint8)
    a, r = ops.gae_scan(e, e, f, f, torch.zeros(8, device="cuda"), 0.99, 0.95, True)
    assert a.shape == (0, 8)
    with pytest.raises(TypeError):
        ops.gae_scan(torch.zeros(2, 8, device="cuda", dtype=torch.float64), torch.zeros(2, 8, device="cuda"),
                     torch.zeros(2, 8, device="cuda", dtype=torch.uint8),
                     torch.zeros(2, 8, device="cuda", dtype=torch.uint8), torch.zeros(8, device="cuda"),
                     0.99, 0.95, True)
    with pytest.raises(ValueError):
        ops.gae_scan(torch.zeros(2, 8), torch.zeros(2, 8), torch.zeros(2, 8, dtype=torch.uint8),
                     torch.zeros(2, 8, dtype=torch.uint8), torch.zeros(8), 0.99, 0.95, True)
