"""Tensor-level wrappers over the C ABI (include/torchrl_b200.h).

Each function takes torch CUDA tensors, checks dtype / contiguity / device, and launches
the sm_100a kernel on torch's *current* stream (so every call is CUDA-graph capturable).
Memory is owned by PyTorch's caching allocator; the library never allocates.  There is no
CPU implementation: CPU tensors raise.
"""
import ctypes

import torch

from . import _lib

F32, F64, U8, I32, I64 = torch.float32, torch.float64, torch.uint8, torch.int32, torch.int64


def _stream():
    return torch.cuda.current_stream().cuda_stream


class CapturedGraph:
    """torch.cuda.CUDAGraph capture of `fn` that also remembers how many library kernels it holds,
    so that replays keep `_lib.launch_count()` truthful."""

    def __init__(self, fn):
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            fn()
        self.launches = _lib.launch_count() - before
        _lib.add_launches(-self.launches)          # capture records, it does not execute

    def replay(self):
        self.graph.replay()
        _lib.add_launches(self.launches)


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (torchrl_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t.data_ptr()


def _opt(t, dtype, name):
    return None if t is None else _chk(t, dtype, name)


def _tn(t):
    """(T, N) sizes of a (T,N) or (T,N,1) tensor."""
    if t.dim() == 3 and t.shape[2] == 1:
        return t.shape[0], t.shape[1]
    if t.dim() == 2:
        return t.shape[0], t.shape[1]
    raise ValueError("expected a (T,N) or (T,N,1) tensor, got %s" % (tuple(t.shape),))


# ------------------------------------------------------------------------------------------ K6
def gae_scan(rewards, values, terminals, time_limits, last_value, gamma, tau, time_limit_filter,
             advs=None, returns=None, variant=1):
    """GAE backward scan.  Mirrors OnPolicyReplayBufferBase.generalized_advantage_estimation
    (/root/reference/torchrl/replay_buffers/on_policy.py:16-44) on (T,N[,1]) device tensors."""
    T, N = _tn(rewards)
    if advs is None:
        advs = torch.empty_like(rewards)
    if returns is None:
        returns = torch.empty_like(rewards)
    assert values.shape == rewards.shape and terminals.shape == rewards.shape and \
        time_limits.shape == rewards.shape and last_value.numel() == N
    _lib.call("trl_gae_scan",
              _chk(rewards, F32, "rewards"), _chk(values, F32, "values"),
              _chk(terminals, U8, "terminals"), _chk(time_limits, U8, "time_limits"),
              _chk(last_value, F32, "last_value"), _chk(advs, F32, "advs"), _chk(returns, F32, "returns"),
              T, N, float(gamma), float(tau), int(bool(time_limit_filter)), int(variant), _stream())
    return advs, returns


def discount_return(rewards, values, terminals, time_limits, last_value, gamma, time_limit_filter,
                    advs=None, returns=None, variant=1):
    """Discounted-reward returns.  Mirrors OnPolicyReplayBufferBase.discount_reward
    (/root/reference/torchrl/replay_buffers/on_policy.py:46-70)."""
    T, N = _tn(rewards)
    if advs is None:
        advs = torch.empty_like(rewards)
    if returns is None:
        returns = torch.empty_like(rewards)
    assert values.shape == rewards.shape and last_value.numel() == N
    _lib.call("trl_discount_return",
              _chk(rewards, F32, "rewards"), _chk(values, F32, "values"),
              _chk(terminals, U8, "terminals"), _chk(time_limits, U8, "time_limits"),
              _chk(last_value, F32, "last_value"), _chk(advs, F32, "advs"), _chk(returns, F32, "returns"),
              T, N, float(gamma), int(bool(time_limit_filter)), int(variant), _stream())
    return advs, returns


# ------------------------------------------------------------------------------------------ K2
def obs_norm_moments(x, sums=None):
    N, o = x.shape
    if sums is None:
        sums = torch.empty(2 * o, dtype=F64, device=x.device)
    _lib.call("trl_obs_norm_moments", _chk(x, F32, "x"), N, o, _chk(sums, F64, "sums"), _stream())
    return sums


def obs_norm_merge(sums, batch_n, mean, var, count):
    o = mean.numel()
    _lib.call("trl_obs_norm_merge", _chk(sums, F64, "sums"), float(batch_n), o, _chk(mean, F64, "mean"),
              _chk(var, F64, "var"), _chk(count, F64, "count"), _stream())


def obs_norm_filt(raw, mean, var, clip=10.0, out=None):
    N, o = raw.shape
    if out is None:
        out = torch.empty_like(raw)
    _lib.call("trl_obs_norm_filt", _chk(raw, F32, "raw"), _chk(mean, F64, "mean"), _chk(var, F64, "var"), N, o,
              float(clip), _chk(out, F32, "out"), _stream())
    return out


# ------------------------------------------------------------------------------------------ K3
def tanh_gaussian_sample(mean, log_std, eps=None, tanh_action=True, want_log_prob=False, want_pre_tanh=False,
                         want_eps=False, noise_scale=1.0, rng=None, nan_flag=None, action_out=None):
    """action = tanh(mean + exp(log_std)*eps) (+ log-prob, pre-tanh).  eps None -> Philox noise keyed by
    (rng.seed, rng.counter).  Mirrors TanhNormal.rsample / log_prob
    (/root/reference/torchrl/policies/distribution.py:60-76, 33-45)."""
    a = mean.shape[-1]
    M = mean.numel() // a
    action = action_out if action_out is not None else torch.empty_like(mean)
    pre = torch.empty_like(mean) if want_pre_tanh else None
    logp = torch.empty(mean.shape[:-1] + (1,), dtype=F32, device=mean.device) if want_log_prob else None
    eps_out = torch.empty_like(mean) if want_eps else None
    ls_stride = 0 if log_std.dim() == 1 else a
    if ls_stride:
        assert log_std.shape == mean.shape
    seed, ctr = (0, None)
    if eps is None:
        if rng is None or rng.counter is None:
            raise ValueError("tanh_gaussian_sample needs either eps or an rng state")
        seed, ctr = rng.seed, rng.counter
    _lib.call("trl_tanh_gaussian_sample", _chk(mean, F32, "mean"), _chk(log_std, F32, "log_std"), ls_stride,
              _opt(eps, F32, "eps"), float(noise_scale), ctypes.c_uint64(seed), _opt(ctr, I64, "rng_counter"), M, a,
              int(bool(tanh_action)), _chk(action, F32, "action"), _opt(pre, F32, "pre_tanh"),
              _opt(logp, F32, "log_prob"), _opt(eps_out, F32, "eps_out"), _opt(nan_flag, I32, "nan_flag"), _stream())
    out = {"action": action}
    if pre is not None:
        out["pre_tanh"] = pre
    if logp is not None:
        out["log_prob"] = logp
    if eps_out is not None:
        out["eps"] = eps_out
    return out


def tanh_gaussian_sample_bwd(action, eps, log_std, g_action, g_logp, tanh_action):
    a = action.shape[-1]
    M = action.numel() // a
    g_mean = torch.empty_like(action)
    g_ls = torch.empty_like(action)
    ls_stride = 0 if log_std.dim() == 1 else a
    ga = g_action.contiguous() if g_action is not None else None
    gl = g_logp.contiguous() if g_logp is not None else None
    _lib.call("trl_tanh_gaussian_sample_bwd", _chk(action, F32, "action"), _chk(eps, F32, "eps"),
              _chk(log_std, F32, "log_std"), ls_stride, _opt(ga, F32, "g_action"), _opt(gl, F32, "g_logp"), M, a,
              int(bool(tanh_action)), _chk(g_mean, F32, "g_mean"), _chk(g_ls, F32, "g_log_std"), _stream())
    return g_mean, g_ls


def counter_advance(counter, t_ptr=None, T=1, size_ptr=None):
    """Device-side `counter += 1` (and optionally t = (t+1) % T, size = min(size+1, T))."""
    _lib.call("trl_step_advance", _opt(t_ptr, I32, "t_ptr"), int(T), _opt(size_ptr, I32, "size_ptr"),
              _opt(counter, I64, "counter"), _stream())


# ------------------------------------------------------------------------------------------ K7/K9/K4
class RowCopyPlan:
    """Pre-built key table for trl_row_gather / trl_ring_write (host arrays of device pointers)."""

    def __init__(self, srcs, dsts, row_bytes):
        n = len(srcs)
        assert n == len(dsts) == len(row_bytes) and 1 <= n <= 8
        self.n = n
        self.src = (ctypes.c_void_p * n)(*[s.data_ptr() for s in srcs])
        self.dst = (ctypes.c_void_p * n)(*[d.data_ptr() for d in dsts])
        self.rb = (ctypes.c_int64 * n)(*[int(b) for b in row_bytes])
        self._keep = (list(srcs), list(dsts))


def row_bytes_of(t):
    """bytes of one time-row of a (T, N, ...) tensor"""
    return t[0].numel() * t.element_size()


def row_gather(plan, idx, rows, pos_ptr=None):
    """dst[k] = src[idx[pos*rows + k]] for every key of the plan; idx is an int64 device tensor."""
    _lib.call("trl_row_gather", plan.n, plan.src, plan.dst, plan.rb, _chk(idx, I64, "idx"),
              _opt(pos_ptr, I32, "pos_ptr"), int(rows), _stream())


def ring_write(plan, row_ptr):
    """dst[*row_ptr] = src[0] for every key of the plan (one time-row)."""
    _lib.call("trl_ring_write", plan.n, plan.src, plan.dst, plan.rb, _chk(row_ptr, I32, "row_ptr"), _stream())


def ring_write_advance(plan, row_ptr, T, ticket, size_ptr=None):
    """ring_write(plan, row_ptr) then row_ptr = (row_ptr + 1) % T [, size = min(size + 1, T)] in one launch.
    ticket: a zero-initialised int32[1] owned by the caller."""
    _lib.call("trl_ring_write_advance", plan.n, plan.src, plan.dst, plan.rb, _chk(row_ptr, I32, "row_ptr"), int(T),
              _opt(size_ptr, I32, "size_ptr"), _chk(ticket, I32, "ticket"), _stream())


def vec_stats(x, out=None):
    """[mean, unbiased std, max, min] of a float vector, on the device."""
    if out is None:
        out = torch.empty(4, dtype=F32, device=x.device)
    _lib.call("trl_vec_stats", _chk(x, F32, "x"), x.numel(), _chk(out, F32, "stats"), _stream())
    return out


# ------------------------------------------------------------------------------------------ K8
class LossScratch:
    """Scratch + ticket for the two-level reductions of the loss kernels (allocated once)."""

    def __init__(self, B, act_dim, device):
        n = int(_lib.load().trl_ppo_actor_scratch_doubles(int(B), int(act_dim)))
        self.actor = torch.zeros(max(n, 1), dtype=F64, device=device)
        self.critic = torch.zeros(max((int(B) + 255) // 256, 1), dtype=F64, device=device)
        self.tickets = torch.zeros(4, dtype=I32, device=device)
        self.B, self.a = int(B), int(act_dim)


def ppo_actor_loss(mean, log_std, actions, old_logp, advs, adv_stats, clip_para, entropy_coeff, tanh_action,
                   scratch, g_mean=None, g_log_std=None, info=None, logp_out=None, stats_pos=None, ls_clamp=None):
    """PPO clipped-surrogate loss value, dL/dmean, dL/dlog_std and logged stats in one launch
    (/root/reference/torchrl/algo/on_policy/ppo.py:41-91).  ls_clamp=(lo, hi): `log_std` is the raw parameter, the
    policy's torch.clamp is applied inside the kernel and g_log_std is the gradient of the raw parameter."""
    ls_lo, ls_hi = (1.0, -1.0) if ls_clamp is None else (float(ls_clamp[0]), float(ls_clamp[1]))
    B, a = mean.shape
    assert scratch.B >= B and scratch.a == a
    ls_stride = 0 if log_std.dim() == 1 else a
    if g_mean is None:
        g_mean = torch.empty_like(mean)
    if g_log_std is None:
        g_log_std = torch.empty_like(log_std)
    if info is None:
        info = torch.zeros(16, dtype=F32, device=mean.device)
    _lib.call("trl_ppo_actor_loss", _chk(mean, F32, "mean"), _chk(log_std, F32, "log_std"), ls_stride,
              _chk(actions, F32, "actions"), _opt(old_logp, F32, "old_logp"), _chk(advs, F32, "advs"),
              _opt(adv_stats, F32, "adv_stats"), _opt(stats_pos, I32, "stats_pos"), B, a, int(bool(tanh_action)),
              float(clip_para),
              float(entropy_coeff), ls_lo, ls_hi, _chk(g_mean, F32, "g_mean"), _chk(g_log_std, F32, "g_log_std"),
              _opt(logp_out, F32, "logp_out"), _chk(info, F32, "info"), _chk(scratch.actor, F64, "scratch"),
              scratch.tickets[0:1].data_ptr(), _stream())
    return g_mean, g_log_std, info


def ppo_critic_loss(values, returns, old_values, clipped, clip_para, scratch, g_values=None, info=None):
    """Critic loss value and dL/dV (/root/reference/torchrl/algo/on_policy/ppo.py:93-122)."""
    B = values.numel()
    if g_values is None:
        g_values = torch.empty_like(values)
    if info is None:
        info = torch.zeros(1, dtype=F32, device=values.device)
    _lib.call("trl_ppo_critic_loss", _chk(values, F32, "values"), _chk(returns, F32, "returns"),
              _opt(old_values, F32, "old_values"), B, int(bool(clipped)), float(clip_para),
              _chk(g_values, F32, "g_values"), _chk(info, F32, "info"), _chk(scratch.critic, F64, "scratch"),
              scratch.tickets[1:2].data_ptr(), _stream())
    return g_values, info


def gaussian_log_prob(mean, log_std, actions, tanh_action, out=None):
    B, a = mean.shape
    if out is None:
        out = torch.empty(B, dtype=F32, device=mean.device)
    ls_stride = 0 if log_std.dim() == 1 else a
    _lib.call("trl_gaussian_log_prob", _chk(mean, F32, "mean"), _chk(log_std, F32, "log_std"), ls_stride,
              _chk(actions, F32, "actions"), B, a, int(bool(tanh_action)), _chk(out, F32, "logp"), _stream())
    return out


def row_group_moments(x, idx, groups, b, out=None):
    """out (groups,4) f64 = sum, sum of squares, max, -min over the rows idx[u*b:(u+1)*b] of x (rows, n)."""
    n = x.numel() // x.shape[0]
    if out is None:
        out = torch.empty(groups, 4, dtype=F64, device=x.device)
    _lib.call("trl_row_group_moments", _chk(x, F32, "x"), _chk(idx, I64, "idx"), int(groups), int(b), n,
              _chk(out, F64, "moments"), _stream())
    return out


def group_stats_from_moments(gathered, world, groups, n_total, out=None):
    """out (groups,4) f32 = mean, unbiased std, max, min per group from (world, groups, 4) raw moments."""
    if out is None:
        out = torch.empty(groups, 4, dtype=F32, device=gathered.device)
    _lib.call("trl_group_stats_from_moments", _chk(gathered, F64, "moments"), int(world), int(groups), float(n_total),
              _chk(out, F32, "stats"), _stream())
    return out


# ------------------------------------------------------------------------------------------ K11
def polyak_update(target_flat, source_flat, tau, planes=None):
    """target <- (1 - tau) target + tau source on flat buffers; planes = (hi, lo) flat TF32 planes of the target
    kept current in the same pass (flat.FlatParams.hi / .lo)."""
    hi, lo = planes if planes is not None else (None, None)
    _lib.call("trl_polyak_update", _chk(target_flat, F32, "target"), _chk(source_flat, F32, "source"),
              target_flat.numel(), float(tau), _opt(hi, F32, "hi"), _opt(lo, F32, "lo"), _stream())


# ------------------------------------------------------------------------------------------ K10
class OffPolicyScratch:
    """Scratch + tickets for the off-policy loss kernels (allocated once per batch size)."""

    def __init__(self, B, device):
        n = int(_lib.load().trl_offpolicy_scratch_doubles(int(B)))
        self.buf = [torch.zeros(max(n, 1), dtype=F64, device=device) for _ in range(5)]
        self.tickets = torch.zeros(8, dtype=I32, device=device)
        self.B = int(B)

    def t(self, i):
        return self.tickets[i:i + 1].data_ptr()


def td_target(rewards, terminals, q1_next, q2_next, logp_next, log_alpha, gamma, scratch, y=None, info=None,
              fixed_alpha=1.0):
    """y = r + (1-d)*gamma*(min(Q1',Q2') - alpha*logpi')  (SAC, twin_sac_q.py:133-139) or the TD3 form
    (td3.py:86-90) when logp_next is None.  info[0] = mean reward."""
    B = rewards.numel()
    if y is None:
        y = torch.empty(B, dtype=F32, device=rewards.device)
    if info is None:
        info = torch.zeros(1, dtype=F32, device=rewards.device)
    _lib.call("trl_td_target", _chk(rewards, F32, "rewards"), _chk(terminals, U8, "terminals"),
              _chk(q1_next, F32, "q1_next"), _opt(q2_next, F32, "q2_next"), _opt(logp_next, F32, "logp_next"),
              _opt(log_alpha, F32, "log_alpha"), float(fixed_alpha), float(gamma), B, _chk(y, F32, "y"),
              _chk(info, F32, "info"), scratch.buf[0].data_ptr(), scratch.t(0), _stream())
    return y, info


def td3_smooth_action(action, sigma, noise_clip, eps=None, rng=None, out=None):
    """clamp(a + clamp(sigma*eps, +-c), +-1)  (td3.py:75-84)."""
    if out is None:
        out = torch.empty_like(action)
    seed, ctr = (0, None)
    if eps is None:
        seed, ctr = rng.seed, rng.counter
    _lib.call("trl_td3_smooth_action", _chk(action, F32, "action"), _opt(eps, F32, "eps"), float(sigma),
              float(noise_clip), ctypes.c_uint64(seed), _opt(ctr, I64, "rng_counter"), action.numel(),
              _chk(out, F32, "out"), _stream())
    return out


def sac_alpha_step(logp, target_entropy, log_alpha, adam_state, lr, scratch, info=None, betas=(0.9, 0.999), eps=1e-8):
    """Temperature loss and its Adam step in one launch (twin_sac_q.py:111-123); info = [alpha, alpha_loss]."""
    if info is None:
        info = torch.zeros(2, dtype=F32, device=logp.device)
    _lib.call("trl_sac_alpha_step", _chk(logp, F32, "logp"), float(target_entropy), _chk(log_alpha, F32, "log_alpha"),
              _chk(adam_state, F32, "adam_state"), float(lr), float(betas[0]), float(betas[1]), float(eps),
              logp.numel(), _chk(info, F32, "info"), scratch.buf[1].data_ptr(), scratch.t(1), _stream())
    return info


def sac_policy_loss(logp, q1, q2, log_alpha, scratch, info=None, fixed_alpha=1.0):
    """mean(alpha*logpi - min(q1,q2)) with gradients wrt its three inputs (twin_sac_q.py:145-153)."""
    B = logp.numel()
    g_lp, g1, g2 = torch.empty_like(logp), torch.empty_like(q1), torch.empty_like(q2)
    if info is None:
        info = torch.zeros(5, dtype=F32, device=logp.device)
    _lib.call("trl_sac_policy_loss", _chk(logp, F32, "logp"), _chk(q1, F32, "q1"), _chk(q2, F32, "q2"),
              _opt(log_alpha, F32, "log_alpha"), float(fixed_alpha), B, _chk(g_lp, F32, "g_logp"),
              _chk(g1, F32, "g_q1"), _chk(g2, F32, "g_q2"), _chk(info, F32, "info"), scratch.buf[2].data_ptr(),
              scratch.t(2), _stream())
    return g_lp, g1, g2, info


def twin_mse_loss(q1, q2, y, scratch, info=None):
    """MSE of one or two critics against y: losses in info[0:2], gradients returned (twin_sac_q.py:142-143)."""
    B = q1.numel()
    g1 = torch.empty_like(q1)
    g2 = torch.empty_like(q2) if q2 is not None else None
    if info is None:
        info = torch.zeros(2, dtype=F32, device=q1.device)
    _lib.call("trl_twin_mse_loss", _chk(q1, F32, "q1"), _opt(q2, F32, "q2"), _chk(y, F32, "y"), B,
              _chk(g1, F32, "g1"), _opt(g2, F32, "g2"), _chk(info, F32, "info"), scratch.buf[3].data_ptr(),
              scratch.t(3), _stream())
    return g1, g2, info


def qr_dqn_loss(pred, nxt, actions, rewards, terminals, gamma, scratch, n_actions, n_quantiles, mse=False, kappa=1.0,
                info=None, weights=None, td_out=None):
    """Fused (QR-)DQN loss: quantile-Huber (qrdqn.py:36-60, utils.py:5-13) or squared TD error (dqn.py:53-60);
    returns d loss / d pred with the shape of pred and info = [loss, mean q_s_a, mean reward]."""
    B = rewards.numel()
    grad = torch.empty_like(pred)
    if info is None:
        info = torch.zeros(3, dtype=F32, device=pred.device)
    _lib.call("trl_qr_dqn_loss", _chk(pred, F32, "pred"), _chk(nxt, F32, "next"), _chk(actions, F32, "actions"),
              _chk(rewards, F32, "rewards"), _chk(terminals, U8, "terminals"), _opt(weights, F32, "weights"), B,
              int(n_actions), int(n_quantiles), float(gamma), float(kappa), int(bool(mse)), _chk(grad, F32, "grad"),
              _opt(td_out, F32, "td_out"), _chk(info, F32, "info"),
              scratch.buf[4].data_ptr(), scratch.t(4), _stream())
    return grad, info


# ------------------------------------------------------------------------------------------ K9 prioritised
def per_sample(prio, size, u, beta, idx=None, weights=None):
    """Stratified proportional row sampling + importance weights (csrc/prioritized.cu; parity unpinned)."""
    b = u.numel()
    if idx is None:
        idx = torch.empty(b, dtype=I64, device=prio.device)
    if weights is None:
        weights = torch.empty(b, dtype=F32, device=prio.device)
    _lib.call("trl_per_sample", _chk(prio, F32, "prio"), int(size), _chk(u, F64, "u"), b, float(beta),
              _chk(idx, I64, "idx"), _chk(weights, F32, "weights"), _stream())
    return idx, weights


def per_update(prio, idx, td, alpha, eps, max_prio):
    """prio[idx_k] = (mean_n |td[k,n]| + eps)^alpha and running max priority."""
    b = idx.numel()
    n = td.numel() // b
    _lib.call("trl_per_update", _chk(prio, F32, "prio"), _chk(idx, I64, "idx"), _chk(td, F32, "td"), b, n,
              float(alpha), float(eps), _chk(max_prio, F32, "max_prio"), _stream())


def per_insert(prio, row_ptr, max_prio):
    _lib.call("trl_per_insert", _chk(prio, F32, "prio"), _chk(row_ptr, I32, "row_ptr"), _chk(max_prio, F32, "max_prio"),
              _stream())


# ------------------------------------------------------------------------------------------ tensor-core GEMM
def gemm_tf32x3_nt(a, b, out=None, splits=1, workspace=None, bias=None, act=0):
    """out (M,256) = a (M,K) @ b (256,K)^T on the tcgen05 tensor cores with 3xTF32 error compensation
    (csrc/gemm_tf32x3.cu).  a, b contiguous fp32, K % (32*splits) == 0."""
    M, K = a.shape
    assert b.shape == (256, K), "B must be (256, K)"
    if out is None:
        out = torch.empty(M, 256, dtype=F32, device=a.device)
    if splits > 1 and workspace is None:
        workspace = torch.empty(splits * M * 256, dtype=F32, device=a.device)
    _lib.call("trl_gemm_tf32x3_nt", _chk(a, F32, "a"), _chk(b, F32, "b"), _chk(out, F32, "out"), M, K, int(splits),
              None if workspace is None else workspace.data_ptr(), _opt(bias, F32, "bias"), int(act), _stream())
    if splits > 1:
        _lib.add_launches(1)      # + the split-K reduction launch
    return out


def gemm_tf32x3_tn(a, b, out=None, splits=1, workspace=None):
    """out (M,256) = a (K,M)^T @ b (K,256): the weight-gradient shape dW = g^T x on the tcgen05 tensor cores
    (3xTF32), operands consumed M/N-major straight from their row-major storage (no transposes)."""
    K, M = a.shape
    assert b.shape == (K, 256), "B must be (K, 256)"
    if out is None:
        out = torch.empty(M, 256, dtype=F32, device=a.device)
    if splits > 1 and workspace is None:
        workspace = torch.empty(splits * M * 256, dtype=F32, device=a.device)
    _lib.call("trl_gemm_tf32x3_tn", _chk(a, F32, "a"), _chk(b, F32, "b"), _chk(out, F32, "out"), M, K, int(splits),
              None if workspace is None else workspace.data_ptr(), _stream())
    if splits > 1:
        _lib.add_launches(1)
    return out


def gemm3_pair(a, b, out=None, planes=None, b_nmajor=False, bias=None, act=0):
    """out (M,256) = act(a (M,K) @ B + bias) on CTA pairs (csrc/gemm_pair.cu, tcgen05 cta_group::2, 3xTF32).
    b_nmajor False: b is (256,K) and B = b^T (Linear forward); True: b is (K,256) and B = b (dgrad).
    planes = (hi, lo): pre-split TF32 planes of b (same shape), else b is split in shared memory."""
    M, K = a.shape
    assert tuple(b.shape) == ((K, 256) if b_nmajor else (256, K)), "B must be (K,256) if b_nmajor else (256,K)"
    if out is None:
        out = torch.empty(M, 256, dtype=F32, device=a.device)
    if planes is not None:
        hi, lo = planes
        assert hi.shape == b.shape and lo.shape == b.shape
        bh, bl = _chk(hi, F32, "b_hi"), _chk(lo, F32, "b_lo")
    else:
        bh, bl = _chk(b, F32, "b"), None
    _lib.call("trl_gemm3_pair", _chk(a, F32, "a"), bh, bl, _chk(out, F32, "out"), M, K, int(bool(b_nmajor)),
              _opt(bias, F32, "bias"), int(act), _stream())
    return out


def gemm3_pair_tn(a, b, out=None, splits=1, workspace=None):
    """out (M,256) = a (K,M)^T @ b (K,256) on CTA pairs: the weight-gradient shape, deterministic split-K."""
    K, M = a.shape
    assert b.shape == (K, 256), "B must be (K, 256)"
    if out is None:
        out = torch.empty(M, 256, dtype=F32, device=a.device)
    if splits > 1 and workspace is None:
        workspace = torch.empty(splits * M * 256, dtype=F32, device=a.device)
    _lib.call("trl_gemm3_pair_tn", _chk(a, F32, "a"), _chk(b, F32, "b"), _chk(out, F32, "out"), M, K, int(splits),
              None if workspace is None else workspace.data_ptr(), _stream())
    if splits > 1:
        _lib.add_launches(1)
    return out


def transpose_f32(x, out=None):
    """out (C,R) = x (R,C)^T (contiguous)."""
    R, C = x.shape
    if out is None:
        out = torch.empty(C, R, dtype=F32, device=x.device)
    _lib.call("trl_transpose_f32", _chk(x, F32, "x"), _chk(out, F32, "out"), R, C, _stream())
    return out
