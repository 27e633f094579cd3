"""tests/golden/hotpath.npz: outputs of the UNMODIFIED reference's own functions on seeded inputs
(oracle/make_golden.py::golden_hotpath, generated in the build container) for the non-GAE rows of SURVEY.md 8(a).

CPU part (always runs): the NumPy oracle reproduces every fixture -- that is what pins the oracle -- and the
product's host-side index streams (bit-exact contract) match the reference's.
GPU part: the CUDA kernels, called through the C ABI wrappers, against the same fixtures.  Tolerances: the
reference computes the PPO scalars in fp32 torch (fixture noise ~1e-6 relative); kernels are fp32: losses and
statistics rtol 2e-4, normaliser state (fp64 on the device) 1e-9, filtered observations 1e-5.
"""
import numpy as np
import pytest

from oracle import ref_numpy as rn


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(golden_dir + "/hotpath.npz")


def _info(g, tag):
    return {str(k): float(v) for k, v in zip(g[tag + "_info_keys"], g[tag + "_info_vals"])}


# ------------------------------------------------------------------------------------------ CPU: oracle == reference
def test_oracle_normaliser_matches_reference(g):
    nrm = rn.RunningNorm(17)
    for k in range(6):
        nrm.update(g["norm_batches"][k])
        np.testing.assert_allclose(nrm.mean, g["norm_mean"][k], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(nrm.var, g["norm_var"][k], rtol=1e-13, atol=1e-15)
        assert abs(nrm.count - g["norm_count"][k]) < 1e-12
        np.testing.assert_allclose(nrm.filt(g["norm_batches"][k]), g["norm_filt"][k], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(nrm.filt(g["norm_far"]), g["norm_far_filt"], rtol=1e-12, atol=1e-13)
    assert np.abs(g["norm_far_filt"]).max() == 10.0      # the clip is exercised


def test_oracle_tanh_normal_matches_reference(g):
    std = np.exp(g["tn_log_std"])[None]
    lp = rn.tanh_normal_log_prob(g["tn_acts"], g["tn_mean"], std)
    np.testing.assert_allclose(lp, g["tn_log_prob"], rtol=1e-11, atol=1e-11)
    ent = rn.normal_entropy(np.broadcast_to(std, g["tn_mean"].shape))
    np.testing.assert_allclose(ent, g["tn_entropy"], rtol=1e-12, atol=1e-12)


def test_index_streams_are_bit_exact(g):
    """K7 / K9: the minibatch row order and the uniform row indices come from the same global NumPy stream as the
    reference's one_iteration / random_batch."""
    from torchrl_b200.replay_buffers import OnPolicyReplayBuffer
    T, N, b = 24, 4, 6
    np.random.seed(5)
    got = []
    for _ in range(2):
        order = rn.epoch_row_order(T, True)
        got += [order[i:i + b] for i in range(0, T, b)]
    np.testing.assert_array_equal(np.array(got), g["iter_rows"])
    buf = OnPolicyReplayBuffer(env_nums=N, max_replay_buffer_size=T * N)        # product: host-side order only
    np.random.seed(5)
    got = []
    for _ in range(2):
        order = buf.epoch_order(True)
        got += [order[i:i + b] for i in range(0, T, b)]
    np.testing.assert_array_equal(np.array(got), g["iter_rows"])
    np.random.seed(6)
    rows = np.array([rn.uniform_row_indices(33, 5 * N, N) for _ in range(4)])
    np.testing.assert_array_equal(rows, g["rand_rows"])


@pytest.mark.parametrize("tag,clipped", [("ppo", False), ("ppoc", True)])
def test_oracle_ppo_losses_match_reference(g, tag, clipped):
    info = _info(g, tag)
    advs = g[tag + "_advs"]
    assert abs(advs.astype(np.float32).mean() - info["advs/mean"]) < 1e-6
    assert abs(advs.astype(np.float32).std(ddof=1) - info["advs/std"]) < 1e-5
    loss, lp, ratio, _ = rn.ppo_actor_loss(g[tag + "_mean"], g[tag + "_log_std"], g[tag + "_acts"],
                                           g[tag + "_old_logp"], rn.normalize_advantages(advs), 0.2, 0.005)
    assert abs(loss - info["Training/policy_loss"]) < 2e-5 * max(1.0, abs(loss))
    assert abs(lp.mean() - info["logprob/mean"]) < 1e-4 and abs(lp.max() - info["logprob/max"]) < 1e-4
    assert abs(lp.min() - info["logprob/min"]) < 1e-4 and abs(lp.std(ddof=1) - info["logprob/std"]) < 1e-4
    assert abs(ratio.max() - info["ratio/max"]) < 1e-4 * info["ratio/max"]
    assert abs(ratio.min() - info["ratio/min"]) < 1e-4
    assert info["ratio/max"] > 1.2 and info["ratio/min"] < 0.8          # both clip branches are exercised
    vloss = rn.ppo_critic_loss(g[tag + "_v"], g[tag + "_values"], g[tag + "_estimate_returns"], 0.2, clipped)
    assert abs(vloss - info["Training/vf_loss"]) < 2e-6 * max(1.0, vloss)


def test_oracle_quantile_loss_matches_reference(g):
    assert abs(rn.quantile_regression_loss(g["qr_tau"], g["qr_source"], g["qr_target"]) - float(g["qr_loss"])) < 1e-13


# ------------------------------------------------------------------------------------------ GPU: kernels vs golden
@pytest.mark.gpu
def test_obs_norm_kernels_vs_golden(g):
    import torch
    from torchrl_b200.env import DeviceNormalizer
    nrm = DeviceNormalizer((17,), device="cuda")
    for k in range(6):
        x = torch.from_numpy(g["norm_batches"][k].astype(np.float32)).cuda()
        x64 = x.cpu().numpy().astype(np.float64)            # what the fp32 device buffer actually holds
        nrm.update_estimate(x)
        ob = nrm.filt(x).cpu().numpy()
        # state is fp64 on the device; the inputs were rounded to fp32 first: compare with the oracle on the same
        # rounded inputs at 1e-9, and with the reference's float64 fixture at fp32 input round-off
        if k == 0:
            ref = rn.RunningNorm(17)
        ref.update(x64)
        np.testing.assert_allclose(nrm._mean.cpu().numpy(), ref.mean, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(nrm._var.cpu().numpy(), ref.var, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(nrm._mean.cpu().numpy(), g["norm_mean"][k], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(nrm._var.cpu().numpy(), g["norm_var"][k], rtol=1e-5, atol=1e-5)
        assert abs(nrm._count.item() - g["norm_count"][k]) < 1e-9
        np.testing.assert_allclose(ob, g["norm_filt"][k], rtol=1e-4, atol=1e-4)
    far = nrm.filt(torch.from_numpy(g["norm_far"].astype(np.float32)).cuda()).cpu().numpy()
    np.testing.assert_allclose(far, g["norm_far_filt"], rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_log_prob_kernel_vs_golden(g):
    import torch
    from torchrl_b200 import ops
    lp = ops.gaussian_log_prob(torch.from_numpy(g["tn_mean"].astype(np.float32)).cuda(),
                               torch.from_numpy(g["tn_log_std"].astype(np.float32)).cuda(),
                               torch.from_numpy(g["tn_acts"].astype(np.float32)).cuda(), True).cpu().numpy()
    np.testing.assert_allclose(lp.reshape(-1), g["tn_log_prob"].sum(-1), rtol=2e-4, atol=2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,clipped", [("ppo", False), ("ppoc", True)])
def test_ppo_loss_kernels_vs_golden(g, tag, clipped):
    import torch
    from torchrl_b200 import ops
    info = _info(g, tag)
    dev = "cuda"
    f32 = lambda k: torch.from_numpy(g[tag + "_" + k].astype(np.float32)).to(dev)
    B, a = g[tag + "_mean"].shape
    scratch = ops.LossScratch(B, a, dev)
    advs = f32("advs").reshape(-1)
    stats = ops.vec_stats(advs)
    s = stats.cpu().numpy()
    np.testing.assert_allclose(s, [info["advs/mean"], info["advs/std"], info["advs/max"], info["advs/min"]],
                               rtol=1e-5, atol=1e-6)
    g_mean, g_ls, out = ops.ppo_actor_loss(f32("mean"), f32("log_std").contiguous(), f32("acts"),
                                           f32("old_logp").reshape(-1), advs, stats, 0.2, 0.005, True, scratch)
    out = out.cpu().numpy()
    assert abs(out[0] - info["Training/policy_loss"]) < 2e-4 * max(1.0, abs(info["Training/policy_loss"]))
    np.testing.assert_allclose(out[1:5], [info["logprob/mean"], info["logprob/std"], info["logprob/max"],
                                          info["logprob/min"]], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(out[5:7], [info["ratio/max"], info["ratio/min"]], rtol=5e-4, atol=1e-6)
    np.testing.assert_allclose(out[7], info["log_std/mean"], rtol=1e-5, atol=1e-6)
    assert g_mean.shape == (B, a) and torch.isfinite(g_mean).all() and torch.isfinite(g_ls).all()
    sc2 = ops.LossScratch(B, 1, dev)
    _, vinfo = ops.ppo_critic_loss(f32("v").reshape(-1), f32("estimate_returns").reshape(-1),
                                   f32("values").reshape(-1), clipped, 0.2, sc2)
    assert abs(vinfo.item() - info["Training/vf_loss"]) < 1e-5 * max(1.0, info["Training/vf_loss"])


@pytest.mark.gpu
def test_qr_loss_kernel_vs_golden(g):
    """quantile_regression_loss(tau, source, target) through the fused QR-DQN kernel: one action, zero reward, no
    terminal, gamma = 1 make the kernel's target equal the fixture's `target`."""
    import torch
    from torchrl_b200 import ops
    src, tgt = g["qr_source"], g["qr_target"]
    B, Q = src.shape
    sc = ops.OffPolicyScratch(B, "cuda")
    zeros = torch.zeros(B, device="cuda")
    _, info = ops.qr_dqn_loss(torch.from_numpy(src.astype(np.float32)).cuda(),
                              torch.from_numpy(tgt.astype(np.float32)).cuda(), zeros.clone(), zeros.clone(),
                              torch.zeros(B, dtype=torch.uint8, device="cuda"), 1.0, sc, 1, Q)
    assert abs(info[0].item() - float(g["qr_loss"])) < 2e-5 * max(1.0, float(g["qr_loss"]))
