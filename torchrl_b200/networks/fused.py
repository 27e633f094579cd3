"""Linear (+ activation) layers of MLPBase (/root/reference/torchrl/networks/base.py:24-44) on this library's kernels.

Routing of one Linear layer (default matmul mode "tc3"):
  * 256 output units, reduction length a multiple of 32, >= _TC3_MIN_ROWS rows: the hand-written tcgen05 3xTF32
    GEMM on CTA pairs (csrc/gemm_pair.cu; fp32-faithful) -- forward with bias + activation in the TMEM epilogue,
    dgrad reading the weights N-major (no transpose), wgrad with both operands M/N-major and deterministic split-K.
    Inside a `presplit()` scope the weights come as pre-split TF32 planes kept current by the fused Adam / Polyak
    kernels (flat.FlatParams.hi / .lo); elsewhere the kernel splits them in shared memory.
  * first layer (K = obs_dim <= 24) and output layer (<= 8 units): csrc/skinny.cu (fp32 kernels that stream one
    (M x H) matrix each, weights in registers).  Their weight / bias gradients are "per-CTA slabs, then a slab sum";
    inside `deferred_reduces()` (the fused minibatch body) only the first stages run and ONE launch sums all slabs.
  * anything else: cuBLAS fp32 SIMT + the fused bias/activation epilogues of csrc/mlp_epilogue.cu.
Numerically every route is fp32 arithmetic (3xTF32: 2e-6 relative at K = 256); the bias gradient is a fixed-order
two-level sum.
"""
import os
import weakref

import torch
import torch.nn as nn

from .. import _lib, ops

ACT_CODES = {nn.Tanh: 1, nn.ReLU: 2}
_ENABLED = True
# "tc3"   : 256-wide layers with >= _TC3_MIN_ROWS rows on the hand-written tcgen05 3xTF32 kernel
#           (csrc/gemm_pair.cu, fp32-faithful), skinny first / output layers on csrc/skinny.cu,
#           everything else cuBLAS fp32 SIMT                                                   [default]
# "fp32"  : cuBLAS fp32 SIMT sgemm everywhere
# "tf32x3": error-compensated TF32 through three cuBLAS GEMMs (kept for comparison; no faster than fp32)
_MATMUL_MODE = "tc3"
_TC3_MIN_ROWS = 2048             # one CTA pair per 256 rows: 8 pairs already beat the SIMT sgemm's latency
_TF32X3_MIN_DIM = 64             # layers narrower than this stay on the plain path
# "pair": csrc/gemm_pair.cu (cta_group::2, pre-split weights, 3-stage ring) [default]; "single": csrc/gemm_tf32x3.cu
_GEMM_IMPL = os.environ.get("TORCHRL_B200_GEMM", "pair")


def set_matmul_mode(mode):
    """"tc3" (default): 256-wide layers on the tcgen05 3xTF32 kernel, the rest cuBLAS fp32 SIMT; "fp32": cuBLAS
    fp32 SIMT everywhere; "tf32x3": x@w as x_hi@w_hi + x_lo@w_hi + x_hi@w_lo through three cuBLAS TF32 GEMMs
    (operands split by csrc/mlp_epilogue.cu:split_tf32_kernel; kept for comparison)."""
    global _MATMUL_MODE
    assert mode in ("fp32", "tf32x3", "tc3")
    _MATMUL_MODE = mode


def set_gemm_impl(impl):
    global _GEMM_IMPL
    assert impl in ("pair", "single")
    _GEMM_IMPL = impl


# ---- pre-split weight planes ------------------------------------------------------------------------------------
_FLATS = weakref.WeakSet()       # live flat.FlatParams objects
_PLANES = {}                     # weight address -> (hi, lo) views into the owning FlatParams' planes
_PRESPLIT_DEPTH = 0


def _drop_planes(ptrs):
    for q in ptrs:
        _PLANES.pop(q, None)


def register_flat(flat):
    """Called by flat.FlatParams: remember its TF32 planes per 2-D parameter (dropped when the object dies)."""
    views = flat.plane_views()
    _PLANES.update(views)
    _FLATS.add(flat)
    weakref.finalize(flat, _drop_planes, list(views.keys()))


class presplit:
    """Scope in which the 256-wide layers read their weights from the pre-split planes.  On entry every live flat
    parameter buffer refreshes its planes from the current weights (one small launch each); inside the scope the
    weights may only change through the fused Adam / Polyak kernels or FlatParams.copy_from, which keep the planes
    current.  Used by the collectors' epochs and the agents' update loops; direct layer calls outside a scope split
    the weights in shared memory instead (always correct, 32 KB more conversion per stage)."""

    def __enter__(self):
        global _PRESPLIT_DEPTH
        if _PRESPLIT_DEPTH == 0 and _MATMUL_MODE == "tc3" and _GEMM_IMPL == "pair":
            for f in list(_FLATS):
                f.refresh_split()
        _PRESPLIT_DEPTH += 1
        return self

    def __exit__(self, *a):
        global _PRESPLIT_DEPTH
        _PRESPLIT_DEPTH -= 1


def presplit_scope(fn):
    """Decorator: run the method inside a `presplit()` scope."""
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        with presplit():
            return fn(*a, **k)
    return wrapped


def _planes_of(weight):
    return _PLANES.get(weight.data_ptr()) if _PRESPLIT_DEPTH > 0 else None


def mm_fwd(x, weight, bias=None, act=0):
    """act(x (M,K) @ weight (256,K)^T + bias) on the tensor cores."""
    if _GEMM_IMPL == "pair":
        return ops.gemm3_pair(x, weight, planes=_planes_of(weight), bias=bias, act=act)
    return ops.gemm_tf32x3_nt(x, weight, bias=bias, act=act)


def mm_dgrad(gz, weight):
    """gz (M,256) @ weight (256,256): the weights are read N-major by the pair kernel (no transpose)."""
    if _GEMM_IMPL == "pair":
        return ops.gemm3_pair(gz, weight, planes=_planes_of(weight), b_nmajor=True)
    return ops.gemm_tf32x3_nt(gz, ops.transpose_f32(weight))


def get_matmul_mode():
    return _MATMUL_MODE


def _tc3_ok(rows, n_out, k):
    """Shapes served by the hand-written tcgen05 3xTF32 kernel (csrc/gemm_tf32x3.cu): 256 output columns,
    reduction length a multiple of 32, enough rows to fill the chip (one 128-row tile per CTA)."""
    return _MATMUL_MODE == "tc3" and rows >= _TC3_MIN_ROWS and n_out == 256 and k >= 32 and k % 32 == 0


def split_tf32(t):
    """(hi, lo) of a contiguous fp32 tensor."""
    hi, lo = torch.empty_like(t), torch.empty_like(t)
    _lib.call("trl_split_tf32", ops._chk(t, torch.float32, "x"), t.numel(), hi.data_ptr(), lo.data_ptr(),
              ops._stream())
    return hi, lo


_WGRAD_WS = {}


def _stream_key():
    """Scratch buffers (split-K slabs, reduction partials, tickets) are per CUDA stream: the critic and the actor
    branch of a minibatch run concurrently on two streams (algo/on_policy/a2c.py) and must not share them."""
    return torch.cuda.current_stream().cuda_stream


def wgrad(gz, x, out=None):
    """dW = gz^T @ x  for gz (M,H), x (M,K): split-K batched GEMM + partial sum.

    The direct `mm(gz.t(), x)` lands on a slow cuBLAS `nt` kernel for this shape class (tiny output, K =
    minibatch): measured on B200 at M=16384: (256x256) 104 us -> 52 us, (256x17) 54 -> 19 us, (6x256) 27 -> 17 us
    (gpurun_out/wgrad_probe.txt, scripts/wgrad_probe.py).  Splitting the reduction over S slabs exposes S x more
    CTAs; the S partial products are summed in a fixed order (deterministic)."""
    M, H = gz.shape
    K = x.shape[1]
    if _tc3_ok(M, K, M) and H % 128 == 0 and M % (32 * 64) == 0:
        # tcgen05 3xTF32, operands consumed M/N-major from their row-major storage, deterministic split-K
        key = (H, str(gz.device), _stream_key())
        ws = _WGRAD_WS.get(key)
        if ws is None:
            ws = _WGRAD_WS[key] = torch.empty(64 * H * 256, dtype=torch.float32, device=gz.device)
        if _GEMM_IMPL == "pair" and H % 256 == 0:
            return ops.gemm3_pair_tn(gz, x, out=out, splits=64, workspace=ws)
        return ops.gemm_tf32x3_tn(gz, x, out=out, splits=64, workspace=ws)
    if (_MATMUL_MODE != "fp32" and K <= 24 and H % 32 == 0 and H <= 256 and _skinny_ok(gz)
            and (out is None or out.is_contiguous())):
        return skinny_tn(gz, x, out=out)                 # (H,K) = gz^T x, first-layer weight gradient
    if H < 4 or M < 4096:
        return torch.mm(gz.t(), x, out=out) if out is not None else torch.mm(gz.t(), x)
    S = 16 if min(H, K) >= 128 else 64
    while S > 1 and M % S:
        S //= 2
    if S == 1:
        return torch.mm(gz.t(), x, out=out) if out is not None else torch.mm(gz.t(), x)
    m = M // S
    part = torch.bmm(gz.view(S, m, H).transpose(1, 2), x.view(S, m, K))
    return torch.sum(part, 0, out=out) if out is not None else part.sum(0)


def mm3(a_hi, a_lo, b_hi, b_lo):
    """a@b from pre-split operands: three TF32 tensor-core GEMMs accumulated in fp32."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        out = torch.mm(a_lo, b_hi)
        out.addmm_(a_hi, b_lo)
        out.addmm_(a_hi, b_hi)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out


_DIRECT_GRAD = False
_FORK = None
_FORK_STREAMS = {}


class backward_fork:
    """Scope around ONE `torch.autograd.backward` call of an MLP with a fused tail: the weight-gradient work that nothing
    downstream in that backward pass depends on (output-layer dW / db and the 256 x 256 wgrad GEMM) is issued on a
    companion stream while the dgrad GEMM and the first-layer backward continue on the calling stream; the scope's
    exit joins the two (under graph capture: a parallel branch).  Tensors the companion stream reads are kept alive
    until the join.  Outside such a scope the backward is strictly sequential."""

    def __enter__(self):
        global _FORK
        self.main = torch.cuda.current_stream()
        key = self.main.cuda_stream
        if key not in _FORK_STREAMS:
            _FORK_STREAMS[key] = torch.cuda.Stream(device=self.main.device)
        self.side = _FORK_STREAMS[key]
        self.keep, self.used, self.prev = [], False, _FORK
        _FORK = self
        return self

    def __exit__(self, *a):
        global _FORK
        _FORK = self.prev
        if self.used:
            self.main.wait_stream(self.side)
        self.keep.clear()


_FORK_ENABLED = os.environ.get("TORCHRL_B200_BWD_FORK", "1") == "1"


def _fork_here():
    f = _FORK if _FORK_ENABLED else None
    if f is not None and torch.cuda.current_stream().cuda_stream == f.main.cuda_stream:
        return f
    return None


class direct_grad:
    """Context manager: inside it the fused layers write weight / bias gradients STRAIGHT into the parameters'
    pre-allocated `.grad` storage (views of the agent's flat gradient buffer, zeroed by the fused Adam step) and
    return None to autograd, which removes one AccumulateGrad add-kernel per parameter tensor.  Valid only
    when every parameter receives exactly one gradient contribution per backward and `.grad` is zero on
    entry -- true for the PPO minibatch update, which is the only user."""

    def __enter__(self):
        global _DIRECT_GRAD
        self.prev = _DIRECT_GRAD
        _DIRECT_GRAD = True

    def __exit__(self, *a):
        global _DIRECT_GRAD
        _DIRECT_GRAD = self.prev


def _grad_out(param):
    """The parameter's own `.grad` storage when direct mode is on (else None)."""
    return param.grad if (_DIRECT_GRAD and param.grad is not None) else None


def set_fused_epilogue(flag):
    """Globally enable/disable the fused epilogue path (default on for CUDA inputs)."""
    global _ENABLED
    _ENABLED = bool(flag)


def fused_enabled():
    return _ENABLED


class _Workspace:
    """Per-(M,H) scratch for the backward's column-sum partials + tickets (allocated once, reused)."""
    cache = {}

    @classmethod
    def get(cls, M, H, device):
        key = (int(M), int(H), str(device), _stream_key())
        ws = cls.cache.get(key)
        if ws is None:
            n = int(_lib.load().trl_bias_act_bwd_scratch_floats(int(M), int(H)))
            ws = (torch.empty(max(n, 4), dtype=torch.float32, device=device),
                  torch.zeros((H + 127) // 128, dtype=torch.int32, device=device))
            cls.cache[key] = ws
        return ws


_SKINNY_MIN_ROWS = 1024
_SKINNY = True         # csrc/skinny.cu serves the first (K = obs_dim) and output (N <= 8) layers
_TN_WS = {}


def set_skinny(flag):
    """Route the first (K = obs_dim <= 24) and output (N <= 8) Linear layers through csrc/skinny.cu (default on).
    Measured on B200 at M = 16384, warm L2, launch gap included (scripts/skinny_probe.py, profiles/skinny_probe_r2.txt):
    k_fwd 8.3 us (cuBLAS + epilogue: 27.6 cold), n_fwd 4.9 / 3.1 us (N = 6 / 1), output-layer dgrad fused with the
    activation backward + bias gradient 9.2, output-layer wgrad + bias gradient 8.4, first-layer wgrad + bias gradient
    12.2.  `set_skinny(False)` restores cuBLAS for these layers."""
    global _SKINNY
    _SKINNY = bool(flag)


def _skinny_ok(x):
    return _SKINNY and _ENABLED and x.is_cuda and x.shape[0] >= _SKINNY_MIN_ROWS


def _tn_scratch(M, H, K, device):
    key = (M, H, K, str(device), _stream_key())
    ws = _TN_WS.get(key)
    if ws is None:
        n = int(_lib.load().trl_skinny_tn_scratch_floats(M, H, K))
        ws = _TN_WS[key] = torch.empty(n, dtype=torch.float32, device=device)
    return ws


# ---- deferred second stages ---------------------------------------------------------------------------------------------
# Every skinny weight / bias gradient is "per-CTA slabs, then a small kernel that sums the slabs".  Inside a
# `deferred_reduces()` scope (the fused minibatch body, where gradients go straight into the flat buffer and nobody reads
# them before the optimizer) only the first stages are launched; `flush_reduces()` sums the slabs of ALL pending jobs in
# ONE launch (trl_skinny_reduce_jobs): six few-microsecond launches per PPO minibatch become one.
_DEFER = None           # list of pending jobs while a scope is open (backward runs on autograd threads: module state)


class deferred_reduces:
    def __enter__(self):
        global _DEFER
        self._prev = _DEFER
        _DEFER = []
        return self

    def __exit__(self, *exc):
        global _DEFER
        jobs, _DEFER = _DEFER, self._prev
        if jobs and exc[0] is None:
            raise RuntimeError("deferred_reduces: %d pending jobs; call flush_reduces() inside the scope" % len(jobs))
        return False


def _defer_scratch(kind, M, H, K, device):
    """One scratch buffer per pending job (its slabs must survive until the flush): keyed by the job's position."""
    slot = len(_DEFER)
    key = ("defer", slot, kind, M, H, K, str(device))
    ws = _TN_WS.get(key)
    if ws is None:
        lib = _lib.load()
        n = int(lib.trl_skinny_dgrad_act_scratch_floats(M, H)) if kind == 2 else int(lib.trl_skinny_tn_scratch_floats(M, H, K))
        ws = _TN_WS[key] = torch.empty(n, dtype=torch.float32, device=device)
    return ws


def flush_reduces():
    """Second stages of every job recorded since the scope opened, on the current stream (which must already be ordered
    after the streams the first stages ran on)."""
    import ctypes
    global _DEFER
    jobs = _DEFER
    if not jobs:
        return
    for i in range(0, len(jobs), 8):
        part = jobs[i:i + 8]
        n = len(part)
        vp = ctypes.c_void_p
        _lib.call("trl_skinny_reduce_jobs", n, (ctypes.c_int * n)(*[j[0] for j in part]),
                  (vp * n)(*[j[1].data_ptr() for j in part]),
                  (vp * n)(*[0 if j[2] is None else j[2].data_ptr() for j in part]),
                  (vp * n)(*[0 if j[3] is None else j[3].data_ptr() for j in part]),
                  (ctypes.c_int64 * n)(*[j[4] for j in part]), (ctypes.c_int * n)(*[j[5] for j in part]),
                  (ctypes.c_int * n)(*[j[6] for j in part]), (ctypes.c_int * n)(*[j[7] for j in part]), ops._stream())
    _DEFER = []


def _can_defer(*outs):
    """Deferral is only sound when every output is a slice of the flat gradient buffer (direct_grad): autograd would
    otherwise accumulate a tensor that is not complete yet."""
    return _DEFER is not None and all(o is not None for o in outs)


def skinny_tn(a, b, out=None, colsum=None, out_transposed=False, may_defer=False):
    """out = a^T @ b for a (M,H), b (M,K<=32) [+ colsum = b.sum(0)]: csrc/skinny.cu, two deterministic stages."""
    M, H = a.shape
    K = b.shape[1]
    if may_defer and _can_defer(out) and len(_DEFER) < 64:
        ws = _defer_scratch(0, M, H, K, a.device)
        _lib.call("trl_skinny_tn_partial", ops._chk(a, torch.float32, "a"), ops._chk(b, torch.float32, "b"), M, H, K,
                  int(colsum is not None), ws.data_ptr(), ops._stream())
        _DEFER.append((0, ws, out, colsum, M, H, K, int(bool(out_transposed))))
        return out
    ws = _tn_scratch(M, H, K, a.device)
    if out is None:
        out = torch.empty((K, H) if out_transposed else (H, K), dtype=torch.float32, device=a.device)
    _lib.call("trl_skinny_tn", ops._chk(a, torch.float32, "a"), ops._chk(b, torch.float32, "b"),
              ops._chk(out, torch.float32, "out"), None if colsum is None else ops._chk(colsum, torch.float32, "colsum"),
              M, H, K, int(bool(out_transposed)), ws.data_ptr(), ops._stream())
    _lib.add_launches(1)
    return out


class _LinearAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act):
        tc = (_MATMUL_MODE == "tf32x3" and min(x.shape[0], x.shape[1], weight.shape[0]) >= _TF32X3_MIN_DIM)
        if (_MATMUL_MODE != "fp32" and _skinny_ok(x) and x.shape[1] <= 24 and weight.shape[0] % 4 == 0
                and weight.shape[0] <= 1024 and weight.is_contiguous() and bias.is_contiguous()):
            # skinny first layer: GEMM + bias + activation in one memory-bound launch
            z = torch.empty(x.shape[0], weight.shape[0], dtype=torch.float32, device=x.device)
            _lib.call("trl_skinny_k_fwd", x.data_ptr(), weight.data_ptr(), bias.data_ptr(), z.data_ptr(), x.shape[0],
                      x.shape[1], weight.shape[0], act, ops._stream())
            ctx.save_for_backward(x, weight, z)
            ctx.act, ctx.tc, ctx.params = act, False, (weight, bias)
            return z
        if _tc3_ok(x.shape[0], weight.shape[0], x.shape[1]) and weight.is_contiguous():
            # tcgen05: act(x (M,K) . W (256,K)^T + b), bias + activation fused into the TMEM epilogue
            z = mm_fwd(x, weight, bias=bias, act=act)
            ctx.save_for_backward(x, weight, z)
            ctx.act, ctx.tc, ctx.params = act, False, (weight, bias)
            return z
        elif tc:
            x_hi, x_lo = split_tf32(x)
            w_hi, w_lo = split_tf32(weight)
            z = mm3(x_hi, x_lo, w_hi.t(), w_lo.t())
            ctx.save_for_backward(x_hi, x_lo, w_hi, w_lo, z)
        else:
            z = torch.mm(x, weight.t())
            ctx.save_for_backward(x, weight, z)
        _lib.call("trl_bias_act_fwd", z.data_ptr(), ops._chk(bias, torch.float32, "bias"), z.shape[0], z.shape[1],
                  act, ops._stream())
        ctx.act = act
        ctx.tc = tc
        ctx.params = (weight, bias)        # the Parameter objects themselves (for direct-grad mode)
        return z

    @staticmethod
    def backward(ctx, g):
        if ctx.tc:
            return _LinearAct._backward_tc(ctx, g)
        x, weight, y = ctx.saved_tensors
        g = g if g.is_contiguous() else g.contiguous()
        M, H = y.shape
        w_param, b_param = ctx.params
        db_out, dw_out = _grad_out(b_param), _grad_out(w_param)
        db = db_out if db_out is not None else torch.empty(H, dtype=torch.float32, device=y.device)
        K = x.shape[1]
        if (not ctx.needs_input_grad[0] and _MATMUL_MODE != "fp32" and _skinny_ok(y) and K <= 24 and H % 32 == 0
                and H <= 256 and x.is_contiguous() and (dw_out is None or dw_out.is_contiguous())):
            # first layer (its input needs no gradient): dW and db straight from g and y in ONE pass over the
            # (M, H) matrices; the activation gradient gz is never written to memory
            dw = dw_out if dw_out is not None else torch.empty(H, K, dtype=torch.float32, device=y.device)
            if _can_defer(dw_out, db_out):
                ws = _defer_scratch(1, M, H, K, y.device)
                _lib.call("trl_skinny_act_wgrad_partial", g.data_ptr(), y.data_ptr(), x.data_ptr(), M, H, K, ctx.act,
                          ws.data_ptr(), ops._stream())
                _DEFER.append((1, ws, dw, db, M, H, K, 0))
                return None, None, None, None
            _lib.call("trl_skinny_act_wgrad", g.data_ptr(), y.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(),
                      M, H, K, ctx.act, _tn_scratch(M, H, K, y.device).data_ptr(), ops._stream())
            _lib.add_launches(1)
            return None, None if dw_out is not None else dw, None if db_out is not None else db, None
        gz = torch.empty_like(y)
        scratch, tickets = _Workspace.get(M, H, y.device)
        _lib.call("trl_bias_act_bwd", g.data_ptr(), y.data_ptr(), gz.data_ptr(), db.data_ptr(), M, H, ctx.act,
                  scratch.data_ptr(), tickets.data_ptr(), ops._stream())
        dx = None
        if ctx.needs_input_grad[0]:
            if _tc3_ok(M, weight.shape[1], H) and weight.is_contiguous():
                dx = mm_dgrad(gz, weight)
            else:
                dx = torch.mm(gz, weight)
        dw = wgrad(gz, x, out=dw_out) if ctx.needs_input_grad[1] else None
        return (dx, None if dw_out is not None else dw,
                None if db_out is not None else (db if ctx.needs_input_grad[2] else None), None)


def _backward_tc(ctx, g):
    x_hi, x_lo, w_hi, w_lo, y = ctx.saved_tensors
    g = g if g.is_contiguous() else g.contiguous()
    M, H = y.shape
    gz = torch.empty_like(y)
    db = torch.empty(H, dtype=torch.float32, device=y.device)
    scratch, tickets = _Workspace.get(M, H, y.device)
    _lib.call("trl_bias_act_bwd", g.data_ptr(), y.data_ptr(), gz.data_ptr(), db.data_ptr(), M, H, ctx.act,
              scratch.data_ptr(), tickets.data_ptr(), ops._stream())
    g_hi, g_lo = split_tf32(gz)
    dx = mm3(g_hi, g_lo, w_hi, w_lo) if ctx.needs_input_grad[0] else None
    dw = mm3(g_hi.t(), g_lo.t(), x_hi, x_lo) if ctx.needs_input_grad[1] else None
    return dx, dw, (db if ctx.needs_input_grad[2] else None), None


_LinearAct._backward_tc = staticmethod(_backward_tc)


class _LinearPlain(torch.autograd.Function):
    """Linear without activation (the output layer of Net): cuBLAS forward, split-K wgrad backward."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.params = (weight, bias)
        N, H = weight.shape
        ctx.skinny = (_MATMUL_MODE != "fp32" and _skinny_ok(x) and N <= 8 and H in (128, 256)
                      and weight.is_contiguous())
        if ctx.skinny:
            y = torch.empty(x.shape[0], N, dtype=torch.float32, device=x.device)
            _lib.call("trl_skinny_n_fwd", x.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(), x.shape[0], H, N,
                      ops._stream())
            return y
        return torch.addmm(bias, x, weight.t())

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = g if g.is_contiguous() else g.contiguous()
        w_param, b_param = ctx.params
        db_out, dw_out = _grad_out(b_param), _grad_out(w_param)
        if ctx.skinny:
            N, H = weight.shape
            dx = None
            if ctx.needs_input_grad[0]:
                dx = torch.empty(x.shape[0], H, dtype=torch.float32, device=x.device)
                _lib.call("trl_skinny_n_dgrad", g.data_ptr(), weight.data_ptr(), dx.data_ptr(), x.shape[0], H, N,
                          ops._stream())
            db = db_out if db_out is not None else torch.empty(N, dtype=torch.float32, device=x.device)
            dw = skinny_tn(x, g, out=dw_out, colsum=db, out_transposed=True)     # dW (N,H) = g^T x, db = sum g
            return dx, None if dw_out is not None else dw, None if db_out is not None else db
        dx = torch.mm(g, weight) if ctx.needs_input_grad[0] else None
        dw = wgrad(g, x, out=dw_out) if ctx.needs_input_grad[1] else None
        db = None
        if ctx.needs_input_grad[2]:
            db = torch.sum(g, 0, out=db_out) if db_out is not None else g.sum(0)
        return dx, None if dw_out is not None else dw, None if db_out is not None else db


class _MLPTail(torch.autograd.Function):
    """Last hidden layer + output layer of an MLP head as ONE autograd node:
        y2 = act(x W2^T + b2)   (tcgen05 3xTF32, bias + activation in the TMEM epilogue)
        out = y2 W3^T + b3      (csrc/skinny.cu n_fwd)
    so that the backward can fuse the output-layer dgrad with the activation backward of the hidden layer
    (trl_skinny_n_dgrad_act: gz2 and db2 in one pass, the (M, 256) dgrad matrix is never stored un-activated)."""

    @staticmethod
    def forward(ctx, x, w2, b2, w3, b3, act):
        y2 = mm_fwd(x, w2, bias=b2, act=act)
        M, H = y2.shape
        N = w3.shape[0]
        out = torch.empty(M, N, dtype=torch.float32, device=x.device)
        _lib.call("trl_skinny_n_fwd", y2.data_ptr(), w3.data_ptr(), b3.data_ptr(), out.data_ptr(), M, H, N,
                  ops._stream())
        ctx.save_for_backward(x, w2, y2, w3)
        ctx.act = act
        ctx.params = (w2, b2, w3, b3)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w2, y2, w3 = ctx.saved_tensors
        g = g if g.is_contiguous() else g.contiguous()
        M, H = y2.shape
        N = w3.shape[0]
        dev = y2.device
        w2p, b2p, w3p, b3p = ctx.params
        dw2_out, db2_out, dw3_out, db3_out = _grad_out(w2p), _grad_out(b2p), _grad_out(w3p), _grad_out(b3p)
        db2 = db2_out if db2_out is not None else torch.empty(H, dtype=torch.float32, device=dev)
        db3 = db3_out if db3_out is not None else torch.empty(N, dtype=torch.float32, device=dev)
        gz = torch.empty_like(y2)
        if _can_defer(db2_out):
            ws = _defer_scratch(2, M, H, 0, dev)
            _lib.call("trl_skinny_n_dgrad_act_partial", g.data_ptr(), w3.data_ptr(), y2.data_ptr(), gz.data_ptr(), M, H, N,
                      ctx.act, ws.data_ptr(), ops._stream())
            _DEFER.append((2, ws, None, db2, M, H, 0, 0))
        else:
            key = ("dgrad_act", M, H, str(dev), _stream_key())
            ws = _TN_WS.get(key)
            if ws is None:
                n = int(_lib.load().trl_skinny_dgrad_act_scratch_floats(M, H))
                ws = _TN_WS[key] = torch.empty(n, dtype=torch.float32, device=dev)
            _lib.call("trl_skinny_n_dgrad_act", g.data_ptr(), w3.data_ptr(), y2.data_ptr(), gz.data_ptr(), db2.data_ptr(),
                      M, H, N, ctx.act, ws.data_ptr(), ops._stream())
            _lib.add_launches(1)
        fk = _fork_here()
        if fk is not None and dw2_out is not None and dw3_out is not None:
            # the two weight gradients feed nothing but the optimizer: companion stream (see backward_fork)
            fk.side.wait_stream(fk.main)
            with torch.cuda.stream(fk.side):
                dw3 = skinny_tn(y2, g, out=dw3_out, colsum=db3, out_transposed=True, may_defer=db3_out is not None)
                dw2 = wgrad(gz, x, out=dw2_out)
            fk.keep += [gz, g, y2, x]
            fk.used = True
            dx = mm_dgrad(gz, w2) if ctx.needs_input_grad[0] else None
        else:
            dw3 = skinny_tn(y2, g, out=dw3_out, colsum=db3, out_transposed=True,   # dW3 (N,H) = g^T y2, db3 = sum g
                            may_defer=dw3_out is not None and db3_out is not None)
            dx = mm_dgrad(gz, w2) if ctx.needs_input_grad[0] else None
            dw2 = wgrad(gz, x, out=dw2_out)
        return (dx, None if dw2_out is not None else dw2, None if db2_out is not None else db2,
                None if dw3_out is not None else dw3, None if db3_out is not None else db3, None)


def tail_ok(h, fc, act_module, head):
    """True when `head(act(fc(h)))` can run as one _MLPTail node: 256-wide hidden layer on the tcgen05 kernel, an
    output layer of at most 8 units, enough rows, gradients being recorded."""
    if not (_ENABLED and _SKINNY and h.is_cuda and h.dtype == torch.float32):
        return False
    rows = h.numel() // h.shape[-1]
    return (type(act_module) in ACT_CODES and rows >= max(_TC3_MIN_ROWS, _SKINNY_MIN_ROWS)
            and _tc3_ok(rows, fc.out_features, fc.in_features) and head.in_features == fc.out_features
            and head.out_features <= 8 and fc.bias is not None and head.bias is not None
            and fc.weight.is_contiguous() and head.weight.is_contiguous())


def mlp_tail(h, fc, act_code, head):
    lead = h.shape[:-1]
    h2 = h.reshape(-1, h.shape[-1])
    if not h2.is_contiguous():
        h2 = h2.contiguous()
    out = _MLPTail.apply(h2, fc.weight, fc.bias, head.weight, head.bias, act_code)
    return out.reshape(tuple(lead) + (out.shape[-1],))


def linear_plain(x, fc):
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    out = _LinearPlain.apply(x2, fc.weight, fc.bias)
    return out.reshape(tuple(lead) + (out.shape[-1],))


def linear_act(x, fc, act_code):
    """act(fc(x)) through the fused path; x may have any number of leading dims."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    out = _LinearAct.apply(x2, fc.weight, fc.bias, act_code)
    return out.reshape(tuple(lead) + (out.shape[-1],))


def can_fuse(x, fc, act_module):
    return (_ENABLED and x.is_cuda and x.dtype == torch.float32 and type(act_module) in ACT_CODES
            and fc.out_features % 4 == 0 and fc.bias is not None)
