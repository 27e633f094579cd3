"""K8 / K11 parity: fused PPO losses and the flat clip+Adam step vs PyTorch (float64 / fp32) restatements
of the reference code paths, and vs the NumPy oracle.  Tolerances: losses/gradients rtol 2e-4
(fp32 kernel vs fp64 reference); Adam trajectories rtol 1e-5 after 5 steps."""
import numpy as np
import pytest

from oracle import ref_numpy as rn


def _ref_actor(mean, ls, acts, old_lp, advs, clip, ent_coef, tanh_action=True):
    """torch.float64 restatement of PPO.update_actor + GuassianContPolicyBase.update + TanhNormal.log_prob
    (ppo.py:48-67, continuous_policy.py:134-153, distribution.py:33-45)."""
    import torch
    std = ls.exp().expand_as(mean)
    if tanh_action:
        pre = torch.log((1 + acts) / (1 - acts)) / 2
        lp = torch.distributions.Normal(mean, std).log_prob(pre) - torch.log(1 - acts * acts + 1e-6)
    else:
        lp = torch.distributions.Normal(mean, std).log_prob(acts)
    log_probs = lp.sum(-1, keepdim=True)
    ent = torch.distributions.Normal(mean, std).entropy().sum(-1, keepdim=True)
    advs_n = (advs - advs.mean()) / (advs.std() + 1e-5)
    ratio = torch.exp(log_probs - old_lp)
    s1 = ratio * advs_n
    s2 = torch.clamp(ratio, 1 - clip, 1 + clip) * advs_n
    loss = -torch.mean(torch.min(s2, s1)) - ent_coef * ent.mean()
    return loss, log_probs, ratio


@pytest.mark.gpu
@pytest.mark.parametrize("shared_ls", [True, False])
@pytest.mark.parametrize("B", [64, 1000, 16384])
def test_ppo_actor_loss_and_grads(shared_ls, B):
    import torch
    from torchrl_b200 import ops
    torch.manual_seed(B)
    a = 6
    mean = 0.5 * torch.randn(B, a, dtype=torch.float64)
    ls = (0.2 * torch.randn(a if shared_ls else (B * a), dtype=torch.float64) - 1.5)
    ls = ls if shared_ls else ls.reshape(B, a)
    z = mean + ls.exp() * torch.randn(B, a, dtype=torch.float64) * 1.3
    acts = torch.tanh(z).float().double()                   # stored actions are fp32 values
    advs = torch.randn(B, 1, dtype=torch.float64) * 2 + 0.3
    m_r, l_r = mean.clone().requires_grad_(), ls.clone().requires_grad_()
    with torch.no_grad():
        _, old_lp, _ = _ref_actor(mean + 0.05 * torch.randn_like(mean), ls, acts, 0, advs, 0.2, 0.005)
    loss, lp, ratio = _ref_actor(m_r, l_r, acts, old_lp, advs, 0.2, 0.005)
    loss.backward()
    dev = "cuda"
    scratch = ops.LossScratch(B, a, dev)
    stats = ops.vec_stats(advs.float().reshape(-1).to(dev))
    g_mean, g_ls, info = ops.ppo_actor_loss(mean.float().to(dev), ls.float().to(dev).contiguous(), acts.float().to(dev),
                                            old_lp.float().reshape(-1).to(dev), advs.float().reshape(-1).to(dev), stats,
                                            0.2, 0.005, True, scratch)
    info = info.cpu().numpy()
    assert abs(info[0] - loss.item()) < 2e-4 * max(1, abs(loss.item()))
    np.testing.assert_allclose(info[1], lp.mean().item(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(info[2], lp.std().item(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(info[3], lp.max().item(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(info[4], lp.min().item(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(info[5], ratio.max().item(), rtol=5e-4)
    np.testing.assert_allclose(info[6], ratio.min().item(), rtol=5e-4, atol=1e-6)
    np.testing.assert_allclose(info[7], ls.mean().item(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(info[8], ls.std().item(), rtol=1e-4, atol=1e-6)
    scale = m_r.grad.abs().max().item()
    np.testing.assert_allclose(g_mean.cpu().numpy(), m_r.grad.numpy(), rtol=2e-3, atol=2e-4 * scale)
    scale = l_r.grad.abs().max().item()
    np.testing.assert_allclose(g_ls.cpu().numpy(), l_r.grad.numpy(), rtol=2e-3, atol=2e-4 * scale)
    # the NumPy oracle agrees with the torch restatement (pins the oracle itself)
    o_loss, _, _, _ = rn.ppo_actor_loss(mean.numpy(), ls.numpy(), acts.numpy(), old_lp.numpy(),
                                        rn.normalize_advantages(advs.numpy()), 0.2, 0.005)
    assert abs(o_loss - loss.item()) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("shared_ls", [True, False])
def test_ppo_actor_loss_folds_the_log_std_clamp(shared_ls):
    """ls_clamp=(lo, hi): the kernel takes the RAW log-std, applies torch.clamp itself and returns the gradient of the
    raw parameter -- identical to clamping outside and back-propagating through torch.clamp
    (GuassianContPolicyBasicBias.forward, continuous_policy.py:173-188)."""
    import torch
    from torchrl_b200 import ops
    torch.manual_seed(3)
    dev, B, a, lo, hi = "cuda", 3000, 6, -2.0, -1.0
    mean = (0.5 * torch.randn(B, a)).to(dev)
    raw = (0.8 * torch.randn(a if shared_ls else (B, a)) - 1.5).to(dev)       # some entries outside [lo, hi]
    assert ((raw < lo) | (raw > hi)).any() and ((raw >= lo) & (raw <= hi)).any()
    acts = torch.tanh(mean + torch.randn(B, a, device=dev) * 0.3)
    advs = torch.randn(B, device=dev)
    old_lp = torch.randn(B, device=dev) * 0.1 - 3.0
    scratch = ops.LossScratch(B, a, dev)
    stats = ops.vec_stats(advs)
    ls_c = raw.clone().requires_grad_()
    clamped = torch.clamp(ls_c, lo, hi)
    g_mean0, g_ls0, info0 = ops.ppo_actor_loss(mean, clamped.detach().contiguous(), acts, old_lp, advs, stats, 0.2, 0.005,
                                               True, scratch)
    clamped.backward(g_ls0)
    g_mean1, g_ls1, info1 = ops.ppo_actor_loss(mean, raw, acts, old_lp, advs, stats, 0.2, 0.005, True, scratch,
                                               ls_clamp=(lo, hi))
    assert torch.equal(g_mean0, g_mean1)
    assert torch.equal(info0, info1)
    assert torch.equal(g_ls1, ls_c.grad)
    assert (g_ls1[(raw < lo) | (raw > hi)] == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("clipped", [False, True])
def test_ppo_critic_loss_and_grads(clipped):
    import torch
    from torchrl_b200 import ops
    torch.manual_seed(1)
    B = 5000
    v = torch.randn(B, 1, dtype=torch.float64, requires_grad=True)
    old_v = v.detach() + 0.3 * torch.randn(B, 1, dtype=torch.float64)
    ret = torch.randn(B, 1, dtype=torch.float64)
    if clipped:   # ppo.py:100-107
        vc = old_v + (v - old_v).clamp(-0.2, 0.2)
        loss = 0.5 * torch.max((v - ret).pow(2), (vc - ret).pow(2)).mean()
    else:
        loss = torch.nn.functional.mse_loss(v, ret)
    loss.backward()
    scratch = ops.LossScratch(B, 1, "cuda")
    g, info = ops.ppo_critic_loss(v.detach().float().cuda().reshape(-1), ret.float().cuda().reshape(-1),
                                  old_v.float().cuda().reshape(-1), clipped, 0.2, scratch)
    assert abs(info.item() - loss.item()) < 1e-5 * max(1.0, abs(loss.item()))
    np.testing.assert_allclose(g.cpu().numpy(), v.grad.reshape(-1).numpy(), rtol=1e-4, atol=1e-8)
    assert abs(rn.ppo_critic_loss(v.detach().numpy(), old_v.numpy(), ret.numpy(), 0.2, clipped) - loss.item()) < 1e-12


@pytest.mark.gpu
def test_gaussian_log_prob_matches_reference_formula():
    import torch
    from torchrl_b200 import ops
    rs = np.random.RandomState(0)
    B, a = 4096, 6
    mean = rs.randn(B, a).astype(np.float32)
    ls = (0.2 * rs.randn(a) - 2.0).astype(np.float32)
    acts = np.tanh(mean + np.exp(ls) * rs.randn(B, a)).astype(np.float32)
    lp = ops.gaussian_log_prob(torch.from_numpy(mean).cuda(), torch.from_numpy(ls).cuda(),
                               torch.from_numpy(acts).cuda(), True).cpu().numpy()
    exp = rn.tanh_normal_log_prob(acts.astype(np.float64), mean, np.exp(ls.astype(np.float64))[None]).sum(-1)
    np.testing.assert_allclose(lp, exp, rtol=2e-4, atol=2e-4)


@pytest.mark.gpu
def test_flat_adam_matches_torch_adam_with_clipping():
    """Two segments with different lr; clip_grad_norm_(0.5) per segment; Adam(eps=1e-5) (a2c.py:29-39)."""
    import copy
    import torch
    from torchrl_b200.flat import FlatAdam
    torch.manual_seed(0)
    net_a = torch.nn.Sequential(torch.nn.Linear(17, 64), torch.nn.Tanh(), torch.nn.Linear(64, 6)).cuda()
    net_b = torch.nn.Sequential(torch.nn.Linear(17, 32), torch.nn.Tanh(), torch.nn.Linear(32, 1)).cuda()
    ref_a, ref_b = copy.deepcopy(net_a), copy.deepcopy(net_b)
    opt_a = torch.optim.Adam(ref_a.parameters(), lr=3e-4, eps=1e-5)
    opt_b = torch.optim.Adam(ref_b.parameters(), lr=1e-3, eps=1e-5)
    flat = FlatAdam([net_a, net_b], lrs=[3e-4, 1e-3], eps=1e-5, max_norms=[0.5, 0.5])
    x = torch.randn(256, 17, device="cuda")
    for it in range(5):
        scale = 10.0 if it % 2 == 0 else 0.01          # alternate clipped / unclipped steps
        for nets in ((net_a, net_b), (ref_a, ref_b)):
            loss = scale * (nets[0](x).pow(2).mean() + nets[1](x).pow(2).mean())
            loss.backward()
        na = torch.nn.utils.clip_grad_norm_(ref_a.parameters(), 0.5)
        nb = torch.nn.utils.clip_grad_norm_(ref_b.parameters(), 0.5)
        opt_a.step(); opt_b.step(); opt_a.zero_grad(); opt_b.zero_grad()
        flat.step()
        norms = flat.grad_norms().cpu().numpy()
        np.testing.assert_allclose(norms, [na.item(), nb.item()], rtol=1e-5)
        assert float(flat.grad.abs().max()) == 0.0      # zero_grad folded into the step
    for p, q in zip(list(net_a.parameters()) + list(net_b.parameters()),
                    list(ref_a.parameters()) + list(ref_b.parameters())):
        torch.testing.assert_close(p, q, rtol=1e-5, atol=1e-7)
    # segment mask: only net_b steps
    before = flat.seg_slice(0).clone()
    (net_a(x).sum() + net_b(x).sum()).backward()
    flat.step(active_mask=0b10)
    torch.testing.assert_close(flat.seg_slice(0), before, rtol=0, atol=0)


@pytest.mark.gpu
def test_ring_write_advance_is_write_then_advance():
    import torch
    from torchrl_b200 import ops
    dev = "cuda"
    T = 5
    src_a = torch.zeros(1, 40000, device=dev)
    src_b = torch.zeros(1, 7, dtype=torch.float64, device=dev)
    ring_a = torch.zeros(T, 40000, device=dev)
    ring_b = torch.zeros(T, 7, dtype=torch.float64, device=dev)
    plan = ops.RowCopyPlan([src_a, src_b], [ring_a, ring_b], [40000 * 4, 7 * 8])
    pos = torch.zeros(1, dtype=torch.int32, device=dev)
    size = torch.zeros(1, dtype=torch.int32, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)
    for i in range(7):
        src_a.fill_(i + 1.0)
        src_b.fill_(-(i + 1.0))
        ops.ring_write_advance(plan, pos, T, ticket, size_ptr=size)
        assert int(pos) == (i + 1) % T and int(size) == min(i + 1, T) and int(ticket) == 0
        assert float(ring_a[i % T].min()) == float(ring_a[i % T].max()) == i + 1.0
        assert float(ring_b[i % T].max()) == -(i + 1.0)


@pytest.mark.gpu
def test_polyak_update_matches_formula():
    import torch
    from torchrl_b200 import ops
    t = torch.randn(100003, device="cuda")
    s = torch.randn(100003, device="cuda")
    exp = rn.polyak(t.cpu().numpy().astype(np.float64), s.cpu().numpy().astype(np.float64), 0.005)
    ops.polyak_update(t, s, 0.005)
    np.testing.assert_allclose(t.cpu().numpy(), exp, rtol=1e-6, atol=1e-7)
