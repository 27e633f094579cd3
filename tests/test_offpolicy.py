"""K9/K10/K11 parity for the off-policy agents.

Kernel level: every fused loss/target kernel vs the NumPy oracle (oracle/ref_numpy.py).
Agent level: TwinSACQ.update / TD3.update on the device vs the CPU oracle port (pinned bit-for-bit to
the unmodified reference by tests/test_oracle_vs_reference.py) on identical batches, weights and
noise.  Tolerances: losses / logged stats rtol 2e-3 + atol 2e-4; parameters after 4 updates atol 2e-4.
(QR-)DQN: vs a float64 torch restatement of qrdqn.py:36-60 / dqn.py:53-60 (the reference's DQN.update
itself is broken on the collector's action shape).
"""
import numpy as np
import pytest

from oracle import ref_numpy as rn


@pytest.mark.gpu
def test_td_target_and_mse_kernels():
    import torch
    from torchrl_b200 import ops
    rs = np.random.RandomState(0)
    B = 3000
    r, d = rs.randn(B).astype(np.float32), (rs.rand(B) < 0.2)
    q1, q2, lp = (rs.randn(B).astype(np.float32) for _ in range(3))
    la = np.float32(-0.7)
    sc = ops.OffPolicyScratch(B, "cuda")
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    y, info = ops.td_target(T(r), T(d.astype(np.uint8)), T(q1), T(q2), T(lp), T(np.array([la])), 0.99, sc)
    exp = rn.sac_q_target(r.astype(np.float64), d.astype(np.float64), q1, q2, lp, np.exp(np.float64(la)), 0.99)
    np.testing.assert_allclose(y.cpu().numpy(), exp, rtol=1e-5, atol=1e-5)
    assert abs(info.item() - r.mean()) < 1e-5
    y3, _ = ops.td_target(T(r), T(d.astype(np.uint8)), T(q1), T(q2), None, None, 0.99, sc)
    np.testing.assert_allclose(y3.cpu().numpy(), rn.td3_q_target(r, d.astype(np.float64), q1, q2, 0.99), rtol=1e-5,
                               atol=1e-5)
    g1, g2, info = ops.twin_mse_loss(T(q1), T(q2), y3, sc)
    yv = y3.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(info.cpu().numpy(), [np.mean((q1 - yv) ** 2), np.mean((q2 - yv) ** 2)], rtol=1e-5)
    np.testing.assert_allclose(g1.cpu().numpy(), 2 * (q1 - yv) / B, rtol=1e-5, atol=1e-9)
    eps = rs.randn(B, 6).astype(np.float32)
    a = np.tanh(rs.randn(B, 6)).astype(np.float32)
    out = ops.td3_smooth_action(T(a), 0.2, 0.5, eps=T(eps))
    np.testing.assert_allclose(out.cpu().numpy(), rn.td3_smooth_action(a, 0.2 * eps, 0.5), rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_sac_alpha_and_policy_loss_kernels():
    import torch
    from torchrl_b200 import ops
    rs = np.random.RandomState(1)
    B = 2048
    lp, q1, q2 = (rs.randn(B).astype(np.float32) for _ in range(3))
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    sc = ops.OffPolicyScratch(B, "cuda")
    # alpha: compare three consecutive steps against torch.optim.Adam on the reference expression
    la_ref = torch.zeros(1, dtype=torch.float32, requires_grad=True)
    opt = torch.optim.Adam([la_ref], lr=3e-4)
    la, st = torch.zeros(1, device="cuda"), torch.zeros(3, device="cuda")
    for it in range(3):
        lpt = torch.from_numpy(lp + 0.1 * it)
        loss = -(la_ref * (lpt + (-6.0)).detach()).mean()
        opt.zero_grad(); loss.backward(); opt.step()
        info = ops.sac_alpha_step(T(lp + np.float32(0.1 * it)), -6.0, la, st, 3e-4, sc)
        assert abs(info[1].item() - loss.item()) < 1e-5 * max(1, abs(loss.item()))
        assert abs(la.item() - la_ref.item()) < 1e-7
        assert abs(info[0].item() - np.exp(la_ref.item())) < 1e-6
    # policy loss value + gradients
    g_lp, g1, g2, info = ops.sac_policy_loss(T(lp), T(q1), T(q2), la, sc)
    alpha = np.exp(np.float64(la.item()))
    assert abs(info[0].item() - np.mean(alpha * lp - np.minimum(q1, q2))) < 1e-5
    np.testing.assert_allclose(info[1:5].cpu().numpy(), [lp.mean(), lp.std(ddof=1), lp.max(), lp.min()], rtol=1e-5)
    np.testing.assert_allclose(g_lp.cpu().numpy(), np.full(B, alpha / B), rtol=1e-5)
    np.testing.assert_allclose(g1.cpu().numpy(), np.where(q1 < q2, -1.0 / B, 0.0), rtol=1e-6)
    np.testing.assert_allclose(g2.cpu().numpy(), np.where(q2 < q1, -1.0 / B, 0.0), rtol=1e-6)


def _ref_qr(pred, nxt, acts, r, d, gamma, Q):
    """float64 torch restatement of QRDQN.update's loss (qrdqn.py:36-60) with autograd."""
    import torch
    B = pred.shape[0]
    tau = torch.tensor((2 * np.arange(Q) + 1) / (2.0 * Q)).view(1, -1)
    q_pred = pred.view(B, -1, Q)
    q_s_a = q_pred.gather(1, acts.view(B, 1, 1).repeat(1, 1, Q).long()).squeeze(1)
    nq = nxt.view(B, -1, Q)
    a_star = nq.detach().mean(dim=2).max(dim=1, keepdim=True)[1]
    tgt = r + gamma * (1 - d) * nq.gather(1, a_star.unsqueeze(2).repeat(1, 1, Q)).squeeze(1)
    diff = tgt.detach().unsqueeze(-1) - q_s_a.unsqueeze(1)
    hub = torch.where(diff.abs() < 1.0, 0.5 * diff.pow(2), diff.abs() - 0.5)
    loss = (hub * (tau - (diff.detach() < 0).double()).abs()).mean()
    return loss, q_s_a


@pytest.mark.gpu
@pytest.mark.parametrize("B,A,Q", [(32, 6, 200), (7, 3, 5), (64, 18, 51)])
def test_qr_dqn_loss_kernel(B, A, Q):
    import torch
    from torchrl_b200 import ops
    torch.manual_seed(B)
    pred = torch.randn(B, A * Q, dtype=torch.float64, requires_grad=True)
    nxt = torch.randn(B, A * Q, dtype=torch.float64)
    acts = torch.randint(0, A, (B,)).double()
    r = torch.randn(B, 1, dtype=torch.float64)
    d = (torch.rand(B, 1) < 0.2).double()
    loss, q_s_a = _ref_qr(pred, nxt, acts, r, d, 0.99, Q)
    loss.backward()
    sc = ops.OffPolicyScratch(B, "cuda")
    g, info = ops.qr_dqn_loss(pred.detach().float().cuda(), nxt.float().cuda(), acts.float().cuda(),
                              r.float().cuda().reshape(-1), d.to(torch.uint8).cuda().reshape(-1), 0.99, sc, A, Q)
    assert abs(info[0].item() - loss.item()) < 2e-5 * max(1.0, abs(loss.item()))
    assert abs(info[1].item() - q_s_a.mean().item()) < 1e-5
    np.testing.assert_allclose(g.cpu().numpy(), pred.grad.numpy(), rtol=2e-4, atol=1e-9)
    # numpy oracle agrees with the torch restatement
    tq, _ = rn.qrdqn_targets(r.numpy(), d.numpy(), nxt.view(B, A, Q).numpy(), 0.99)
    tau = (2 * np.arange(Q) + 1) / (2.0 * Q)
    assert abs(rn.quantile_regression_loss(tau, q_s_a.detach().numpy(), tq) - loss.item()) < 1e-12


@pytest.mark.gpu
def test_dqn_loss_kernel():
    import torch
    from torchrl_b200 import ops
    torch.manual_seed(0)
    B, A = 100, 6
    pred = torch.randn(B, A, dtype=torch.float64, requires_grad=True)
    nxt = torch.randn(B, A, dtype=torch.float64)
    acts = torch.randint(0, A, (B, 1))
    r = torch.randn(B, 1, dtype=torch.float64)
    d = (torch.rand(B, 1) < 0.2).double()
    q_s_a = pred.gather(-1, acts)                                  # dqn.py:53-54 with (B,1) actions
    tgt = r + 0.99 * (1 - d) * nxt.max(-1, keepdim=True)[0]
    loss = torch.nn.functional.mse_loss(q_s_a, tgt)
    loss.backward()
    sc = ops.OffPolicyScratch(B, "cuda")
    g, info = ops.qr_dqn_loss(pred.detach().float().cuda(), nxt.float().cuda(), acts.float().cuda().reshape(-1),
                              r.float().cuda().reshape(-1), d.to(torch.uint8).cuda().reshape(-1), 0.99, sc, A, 1,
                              mse=True)
    assert abs(info[0].item() - loss.item()) < 1e-5
    np.testing.assert_allclose(g.cpu().numpy(), pred.grad.numpy(), rtol=1e-4, atol=1e-9)
    assert abs(rn.dqn_target(r.numpy(), d.numpy(), nxt.numpy(), 0.99) - tgt.numpy()).max() < 1e-12


class _Env:
    def __init__(self, o, a):
        from torchrl_b200.spaces import Box
        self.action_space = Box(-np.ones(a), np.ones(a))
        self.observation_space = Box(-np.ones(o), np.ones(o))


class _Col:
    epoch_frames = 1


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sac", "td3"])
def test_agent_update_matches_reference_port(kind):
    import torch
    import torch.nn as nn
    import torchrl_b200.networks as networks
    import torchrl_b200.policies as policies
    from oracle import ref_port
    from torchrl_b200.algo import TD3, TwinSACQ
    from torchrl_b200.policies import set_noise_mode
    from tests.test_oracle_vs_reference import _offpolicy_batches
    o, a, hidden, B, seed = 11, 3, (24, 24), 48, 4
    batches = _offpolicy_batches(o, a, B, 4, seed)
    torch.set_num_threads(4)
    # ---- CPU oracle port
    torch.manual_seed(seed)
    if kind == "sac":
        ppf = ref_port.TanhGaussianPolicy(o, a, list(hidden), nn.ReLU, state_dependent_std=True)
    else:
        ppf = ref_port.FixedNoisePolicy(o, a, list(hidden), nn.ReLU, norm_std_explore=0.1, tanh_action=True)
    pq1, pq2 = ref_port.QNet(o + a, 1, list(hidden), nn.ReLU), ref_port.QNet(o + a, 1, list(hidden), nn.ReLU)
    port = ref_port.SACPort(ppf, pq1, pq2, a, std_reg=1e-3, mean_reg=1e-3) if kind == "sac" else \
        ref_port.TD3Port(ppf, pq1, pq2, plr=1e-3, qlr=1e-3)
    torch.manual_seed(100)
    port_infos = [port.update(b) for b in batches]
    # ---- device agent, same weights (same seed + creation order), same CPU noise stream
    torch.manual_seed(seed)
    net = dict(hidden_shapes=list(hidden), append_hidden_shapes=[], base_type=networks.MLPBase,
               activation_func=nn.ReLU)
    if kind == "sac":
        pf = policies.GuassianContPolicy(input_shape=o, output_shape=2 * a, tanh_action=True, **net)
    else:
        pf = policies.FixGuassianContPolicy(input_shape=o, output_shape=a, tanh_action=True, norm_std_explore=0.1, **net)
    qf1 = networks.QNet(input_shape=o + a, output_shape=1, **net)
    qf2 = networks.QNet(input_shape=o + a, output_shape=1, **net)
    common = dict(env=_Env(o, a), replay_buffer=None, collector=_Col(), logger=None, discount=0.99, batch_size=B,
                  device="cuda:0", save_dir=None, tau=0.005, use_soft_update=True, use_cuda_graph=False)

    class _RB:          # update(batch) never touches the buffer; _ub_setup only needs env_nums
        env_nums = 1
    common["replay_buffer"] = _RB()
    set_noise_mode("reference_cpu")
    try:
        if kind == "sac":
            agent = TwinSACQ(pf=pf, qf1=qf1, qf2=qf2, plr=3e-4, qlr=3e-4, policy_std_reg_weight=1e-3,
                             policy_mean_reg_weight=1e-3, **common)
        else:
            agent = TD3(pf=pf, qf1=qf1, qf2=qf2, plr=1e-3, qlr=1e-3, **common)
        torch.manual_seed(100)
        infos = [agent.update(b) for b in batches]
    finally:
        set_noise_mode("philox")
    for u, (mine, ref) in enumerate(zip(infos, port_infos)):
        assert mine.keys() == ref.keys(), (mine.keys(), ref.keys())
        for k, v in ref.items():
            assert abs(mine[k] - v) <= 2e-3 * abs(v) + 2e-4, (u, k, mine[k], v)
    nets = [agent.pf, agent.qf1, agent.qf2] + ([agent.target_qf1, agent.target_qf2] if kind == "sac" else
                                                  [agent.target_pf, agent.target_qf1, agent.target_qf2])
    pnets = [port.pf, port.qf1, port.qf2] + ([port.tqf1, port.tqf2] if kind == "sac" else
                                               [port.tpf, port.tqf1, port.tqf2])
    mine = torch.cat([p.detach().reshape(-1) for n in nets for p in n.parameters()]).cpu().numpy()
    ref = torch.cat([p.detach().reshape(-1) for n in pnets for p in n.parameters()]).numpy()
    np.testing.assert_allclose(mine, ref, atol=2e-4)


def _build_offpolicy(kind, N=32, T_rows=64, use_graph=True, seed=0, env_id="SynthAnt-v0", batch_rows=4, opt_times=6):
    import torch
    import torch.nn as nn
    import torchrl_b200.networks as networks
    import torchrl_b200.policies as policies
    from torchrl_b200.algo import TD3, TwinSACQ
    from torchrl_b200.collector import VecCollector
    from torchrl_b200.env import get_vec_env
    from torchrl_b200.replay_buffers import BaseReplayBuffer
    from torchrl_b200.utils import NullLogger
    dev = torch.device("cuda:0")
    env = get_vec_env(env_id, {"reward_scale": 1, "obs_norm": False}, N)
    eval_env = get_vec_env(env_id, {"reward_scale": 1, "obs_norm": False}, N)
    env.seed(seed); torch.manual_seed(seed); np.random.seed(seed)
    o, a = env.observation_space.shape[0], env.action_space.shape[0]
    buf = BaseReplayBuffer(env_nums=N, max_replay_buffer_size=T_rows * N, time_limit_filter=False)
    net = dict(hidden_shapes=[32, 32], append_hidden_shapes=[], base_type=networks.MLPBase, activation_func=nn.ReLU)
    if kind == "sac":
        pf = policies.GuassianContPolicy(input_shape=o, output_shape=2 * a, tanh_action=True, **net)
    else:
        pf = policies.FixGuassianContPolicy(input_shape=o, output_shape=a, tanh_action=True, norm_std_explore=0.1, **net)
    qf1 = networks.QNet(input_shape=o + a, output_shape=1, **net)
    qf2 = networks.QNet(input_shape=o + a, output_shape=1, **net)
    col = VecCollector(env=env, eval_env=eval_env, pf=pf, replay_buffer=buf, device=dev, epoch_frames=8 * N,
                       max_episode_frames=20, use_cuda_graph=use_graph)
    common = dict(env=env, replay_buffer=buf, collector=col, logger=NullLogger(), discount=0.99,
                  batch_size=batch_rows * N, device=dev, save_dir=None, tau=0.005, use_soft_update=True,
                  opt_times=opt_times, pretrain_epochs=1, num_epochs=3, use_cuda_graph=use_graph)
    if kind == "sac":
        agent = TwinSACQ(pf=pf, qf1=qf1, qf2=qf2, plr=3e-4, qlr=3e-4, policy_std_reg_weight=0,
                         policy_mean_reg_weight=0, **common)
    else:
        agent = TD3(pf=pf, qf1=qf1, qf2=qf2, plr=1e-3, qlr=1e-3, **common)
    return agent, col, buf, env


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sac", "td3"])
def test_offpolicy_pipeline_graph_vs_eager(kind):
    """VecCollector + ring buffer + agent for 3 epochs: CUDA-graph and eager paths agree; ring semantics
    (terminals exclude the collector timeout, collector/base.py:204-213; resets happen on done|timeout)."""
    import torch
    runs = []
    for use_graph in (False, True):
        agent, col, buf, env = _build_offpolicy(kind, use_graph=use_graph)
        agent.pretrain()
        for epoch in range(3):
            agent.current_epoch = epoch
            col.train_one_epoch()
            agent.update_per_epoch()
        runs.append((agent, buf, [dict(i) for i in agent._last_infos]))
    (a0, b0, i0), (a1, b1, i1) = runs
    assert b0._size == b1._size == 32 and b0._top == b1._top == 32
    for k in ("obs", "next_obs", "acts", "rewards", "terminals", "time_limits"):
        x0, x1 = getattr(b0, "_" + k)[:32], getattr(b1, "_" + k)[:32]
        if x0.dtype == torch.uint8:
            assert torch.equal(x0, x1), k
        else:
            torch.testing.assert_close(x0, x1, rtol=1e-4, atol=1e-5, msg=k)
    torch.testing.assert_close(a0.opt.data, a1.opt.data, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(a0._target_flat.data, a1._target_flat.data, rtol=1e-3, atol=1e-5)
    assert len(i0) == len(i1) == 6
    for d0, d1 in zip(i0, i1):
        assert d0.keys() == d1.keys()
        for k in d0:
            assert abs(d0[k] - d1[k]) <= 2e-3 * max(1.0, abs(d0[k])), (k, d0[k], d1[k])
    # collector timeout at 20 frames resets but does not mark terminal; Ant-shaped env may terminate early
    obs, nxt = b0._obs.cpu().numpy(), b0._next_obs.cpu().numpy()
    assert np.isfinite(obs).all() and np.isfinite(nxt).all()


@pytest.mark.gpu
def test_qrdqn_agent_update_runs_and_matches_restatement():
    import torch
    import torch.nn as nn
    import torchrl_b200.networks as networks
    import torchrl_b200.policies as policies
    from torchrl_b200.algo import QRDQN
    from torchrl_b200.spaces import Box, Discrete
    torch.manual_seed(0)
    o, A, Q, B = 10, 4, 11, 40

    class E:
        action_space = Discrete(A)
        observation_space = Box(-np.ones(o), np.ones(o))

    class RB:
        env_nums = 1
    qf = networks.Net(input_shape=o, output_shape=A * Q, hidden_shapes=[32], append_hidden_shapes=[],
                      base_type=networks.MLPBase, activation_func=nn.ReLU)
    pf = policies.EpsilonGreedyQRDQNDiscretePolicy(quantile_num=Q, qf=qf, start_epsilon=0.1, end_epsilon=0.1,
                                                   decay_frames=100, action_shape=A)
    agent = QRDQN(quantile_num=Q, qf=qf, pf=pf, qlr=1e-3, optimizer_info={"eps": 0.0003125}, env=E(),
                  replay_buffer=RB(), collector=_Col(), logger=None, discount=0.99, batch_size=B, device="cuda:0",
                  save_dir=None, use_soft_update=False, target_hard_update_period=2, use_cuda_graph=False)
    rs = np.random.RandomState(0)
    batch = {"obs": rs.randn(B, o), "next_obs": rs.randn(B, o), "acts": rs.randint(0, A, B).astype(np.float64),
             "rewards": rs.randn(B, 1), "terminals": (rs.rand(B, 1) < 0.2).astype(np.float64)}
    with torch.no_grad():
        pred = qf(torch.tensor(batch["obs"], dtype=torch.float32, device="cuda")).double().cpu()
        nxt = agent.target_qf(torch.tensor(batch["next_obs"], dtype=torch.float32, device="cuda")).double().cpu()
    loss, q_s_a = _ref_qr(pred, nxt, torch.tensor(batch["acts"]), torch.tensor(batch["rewards"]),
                          torch.tensor(batch["terminals"]), 0.99, Q)
    before = agent.opt.data.clone()
    info = agent.update(batch)
    assert abs(info["Training/qf_loss"] - loss.item()) < 1e-4 and abs(info["q_s_a"] - q_s_a.mean().item()) < 1e-4
    assert not torch.equal(before, agent.opt.data)
    assert not torch.equal(agent._target_flat.data, agent.opt.data)       # update 1: no hard copy yet
    agent.update(batch)
    assert torch.equal(agent._target_flat.data, agent.opt.data)           # update 2: hard copy (period 2)
    # greedy action for a batch of envs (the reference's .item() version only handles one env, A.5)
    acts = pf.eval_act(torch.randn(5, o, device="cuda"))
    assert acts.shape == (5, 1)


@pytest.mark.gpu
def test_ddpg_update_matches_reference_port():
    """DDPG on the off-policy kernels vs the oracle port (pinned bit-exact to the reference's DDPG.update by
    tests/test_oracle_vs_reference.py): logged scalars and all parameters after 4 updates."""
    import torch
    import torch.nn as nn
    import torchrl_b200.networks as networks
    import torchrl_b200.policies as policies
    from oracle import ref_port
    from torchrl_b200.algo import DDPG
    from tests.test_oracle_vs_reference import _offpolicy_batches
    o, a, hidden, B, seed = 11, 3, (24, 24), 48, 4
    batches = _offpolicy_batches(o, a, B, 4, seed)
    torch.set_num_threads(4)
    torch.manual_seed(seed)
    ppf = ref_port.FixedNoisePolicy(o, a, list(hidden), nn.ReLU, norm_std_explore=0.1, tanh_action=True)
    pq = ref_port.QNet(o + a, 1, list(hidden), nn.ReLU)
    port = ref_port.DDPGPort(ppf, pq, plr=1e-3, qlr=1e-3)
    port_infos = [port.update(b) for b in batches]
    torch.manual_seed(seed)
    net = dict(hidden_shapes=list(hidden), append_hidden_shapes=[], base_type=networks.MLPBase,
               activation_func=nn.ReLU)
    pf = policies.DetContPolicy(input_shape=o, output_shape=a, tanh_action=True, **net)
    qf = networks.QNet(input_shape=o + a, output_shape=1, **net)

    class _RB:
        env_nums = 1
    agent = DDPG(pf=pf, qf=qf, plr=1e-3, qlr=1e-3, env=_Env(o, a), replay_buffer=_RB(), collector=_Col(),
                 logger=None, discount=0.99, batch_size=B, device="cuda:0", save_dir=None, tau=0.005,
                 use_soft_update=True, use_cuda_graph=False)
    infos = [agent.update(b) for b in batches]
    for u, (mine, ref) in enumerate(zip(infos, port_infos)):
        assert mine.keys() == ref.keys(), (mine.keys(), ref.keys())
        for k, v in ref.items():
            assert abs(mine[k] - v) <= 2e-3 * abs(v) + 2e-4, (u, k, mine[k], v)
    mine = torch.cat([p.detach().reshape(-1) for n in (agent.pf, agent.qf, agent.target_pf, agent.target_qf)
                      for p in n.parameters()]).cpu().numpy()
    ref = torch.cat([p.detach().reshape(-1) for n in (port.pf, port.qf, port.tpf, port.tqf)
                     for p in n.parameters()]).numpy()
    np.testing.assert_allclose(mine, ref, atol=2e-4)
