// gae.cu -- K6: GAE / discounted-return backward scan over a time-major (T, N) rollout.
//
// Replaces the Python `for t in reversed(range(T))` loops of
//   /root/reference/torchrl/replay_buffers/on_policy.py:16-44  (generalized_advantage_estimation)
//   /root/reference/torchrl/replay_buffers/on_policy.py:46-70  (discount_reward)
//
// Both recurrences are affine in the carried quantity x_{t+1}:
//     x_t = a_t + b_t * x_{t+1}
//   GAE  : a_t = m_t*(r_t + nt_t*g*V_{t+1} - V_t), b_t = m_t*g*tau*nt_t, x_T = 0
//          adv_t = x_t, ret_t = x_t + V_t               (m_t = 1-tl_t if filter else 1)
//   DISC : a_t = r_t + tl_t*V_t, b_t = nt_t*g*(1-tl_t)  (filter)  |  a_t = r_t, b_t = nt_t*g
//          x_T = last_value ; adv_t = x_t - V_t, ret_t = x_t
// so the scan is parallelised over TIME as well as over envs: a CTA owns
// (32*VEC envs) x (W*TC timesteps); lane -> env group (coalesced 128B/512B rows),
// warp -> chunk of TC consecutive timesteps.  Every thread issues all of its loads
// up front (TC x 4 arrays, independent), composes its chunk's affine map in
// registers, publishes (a,b) to shared memory, picks up the composition of the
// later chunks as its carry-in, and replays its TC steps from registers -- the data
// are read from HBM exactly once (10 B/elt) and written once (8 B/elt): 18 B/elt.
// HBM-bound; no tensor-core work here.
#include "gae_common.cuh"

namespace trl {

// blockDim = (32, W).  dynamic smem: (2*W + 1) * 32*VEC floats.
template <int MODE, int VEC, int TC>
__global__ void __launch_bounds__(1024) gae_chunked_kernel(const GaeParams p) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x, w = threadIdx.y, W = blockDim.y;
  const int row = 32 * VEC;
  float* sa = smem;                // [W][row]
  float* sb = smem + W * row;      // [W][row]
  float* sc = smem + 2 * W * row;  // [row] super-chunk carry
  const long long env0 = (static_cast<long long>(blockIdx.x) * 32 + lane) * VEC;
  const bool env_ok = env0 < p.N;
  const long long N = p.N, T = p.T;
  const float g = p.gamma, gt = p.gamma_tau;
  const int filter = p.filter;
  const int span = W * TC;

  float carry[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) carry[i] = 0.f;
  if (MODE == MODE_DISC && env_ok) load_f<VEC>(p.last_value + env0, carry);

  for (long long hi = T; hi > 0; hi -= span) {
    const long long t0 = hi - span + static_cast<long long>(w) * TC;  // may be < 0
    float r[TC][VEC], v[TC][VEC], vn[VEC];
    unsigned ft[TC], fl[TC];
    // ---- issue every load of this chunk before any use -----------------------------
#pragma unroll
    for (int k = 0; k < TC; ++k) {
      const long long t = t0 + k;
      if (env_ok && t >= 0) {
        const long long off = t * N + env0;
        load_f<VEC>(p.rewards + off, r[k]);
        load_f<VEC>(p.values + off, v[k]);
        ft[k] = load_flags<VEC>(p.terminals + off);
        fl[k] = load_flags<VEC>(p.time_limits + off);
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) { r[k][i] = 0.f; v[k][i] = 0.f; }
        ft[k] = 0u; fl[k] = 0u;
      }
    }
    {
      // first step of the next chunk; in the ragged head super-chunk (T % span != 0) whole chunks lie at t < 0 and
      // tn can be <= 0: those steps are identities, nothing is read for them (vn stays 0)
      const long long tn = t0 + TC;
#pragma unroll
      for (int i = 0; i < VEC; ++i) vn[i] = 0.f;
      if (env_ok) {
        if (tn >= T) load_f<VEC>(p.last_value + env0, vn);
        else if (tn >= 0) load_f<VEC>(p.values + tn * N + env0, vn);
      }
    }
    // ---- pass 1: per-step coefficients, chunk composition --------------------------
    // a_t overwrites r[k] (dead afterwards); b_t is a function of the flags only and
    // is recomputed in pass 2 instead of being kept live (register budget: 64/thread).
    float ca[VEC], cb[VEC];      // chunk map: x_{t0} = ca + cb * x_{t0+TC}
#pragma unroll
    for (int i = 0; i < VEC; ++i) { ca[i] = 0.f; cb[i] = 1.f; }
#pragma unroll
    for (int k = TC - 1; k >= 0; --k) {
      const bool live = (t0 + k) >= 0;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const unsigned term = (ft[k] >> (8 * i)) & 0xffu, tl = (fl[k] >> (8 * i)) & 0xffu;
        const float vnext = (k == TC - 1) ? vn[i] : v[(k + 1) % TC][i];
        float ak, bk;
        coeffs<MODE>(r[k][i], v[k][i], vnext, term, tl, g, gt, filter, ak, bk);
        if (!live) { ak = 0.f; bk = 1.f; }   // identity for the ragged head (t < 0)
        r[k][i] = ak;
        ca[i] = fmaf(bk, ca[i], ak);
        cb[i] = bk * cb[i];
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      sa[w * row + lane * VEC + i] = ca[i];
      sb[w * row + lane * VEC + i] = cb[i];
    }
    __syncthreads();
    // ---- carry-in: compose the later chunks of this super-chunk onto `carry` ---------
    float x[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) x[i] = carry[i];
    for (int w2 = W - 1; w2 > w; --w2) {
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        x[i] = fmaf(sb[w2 * row + lane * VEC + i], x[i], sa[w2 * row + lane * VEC + i]);
    }
    // ---- pass 2: replay from registers, write outputs -------------------------------
#pragma unroll
    for (int k = TC - 1; k >= 0; --k) {
      const long long t = t0 + k;
      float oa[VEC], orr[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const unsigned term = (ft[k] >> (8 * i)) & 0xffu, tl = (fl[k] >> (8 * i)) & 0xffu;
        float bk = bcoef<MODE>(term, tl, g, gt, filter);
        if (t < 0) bk = 1.f;
        x[i] = fmaf(bk, x[i], r[k][i]);
        if (MODE == MODE_GAE) { oa[i] = x[i]; orr[i] = x[i] + v[k][i]; }
        else { oa[i] = x[i] - v[k][i]; orr[i] = x[i]; }
      }
      if (env_ok && t >= 0) {
        store_f<VEC>(p.advs + t * N + env0, oa);
        store_f<VEC>(p.rets + t * N + env0, orr);
      }
    }
    if (hi > span) {  // another (earlier) super-chunk follows: hand over x at hi-span
      if (w == 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) sc[lane * VEC + i] = x[i];
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < VEC; ++i) carry[i] = sc[lane * VEC + i];
      // next iteration's sa/sb writes are ordered after every read above by this barrier
    }
  }
}

// Reference-shaped kernel: one thread per env, serial over T (validation / comparison).
template <int MODE>
__global__ void gae_serial_kernel(const GaeParams p) {
  const long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  float x = (MODE == MODE_DISC) ? p.last_value[n] : 0.f;
  float vnext = p.last_value[n];
  for (long long t = p.T - 1; t >= 0; --t) {
    const long long off = t * p.N + n;
    const float r = p.rewards[off], v = p.values[off];
    float a, b;
    coeffs<MODE>(r, v, vnext, p.terminals[off], p.time_limits[off], p.gamma, p.gamma_tau, p.filter, a, b);
    x = fmaf(b, x, a);
    if (MODE == MODE_GAE) { p.advs[off] = x; p.rets[off] = x + v; }
    else { p.advs[off] = x - v; p.rets[off] = x; }
    vnext = v;
  }
}

template <int MODE, int VEC, int TC>
static int launch_chunked(const GaeParams& p, cudaStream_t s) {
  const long long chunks = ceil_div<long long>(p.T, TC);
  const int W = static_cast<int>(chunks < 32 ? chunks : 32);
  const long long blocks = ceil_div<long long>(p.N, 32LL * VEC);
  const size_t smem = static_cast<size_t>(2 * W + 1) * 32 * VEC * sizeof(float);
  gae_chunked_kernel<MODE, VEC, TC><<<static_cast<unsigned>(blocks), dim3(32, W), smem, s>>>(p);
  return check_launch("gae_chunked_kernel");
}

template <int MODE>
static int dispatch(const GaeParams& p, int variant, cudaStream_t s) {
  if (variant == 0) {
    const int threads = 128;
    gae_serial_kernel<MODE><<<static_cast<unsigned>(ceil_div<long long>(p.N, threads)), threads, 0, s>>>(p);
    return check_launch("gae_serial_kernel");
  }
  // variant 4 / auto at scale: persistent kernel with TMA-staged tiles (csrc/gae_tma.cu), >= 2 env groups per SM
  if (variant == 4 || (variant == 1 && p.N >= 2LL * 128 * kNumSM)) {
    if (gae_tma_supported(p)) return gae_tma_launch(p, MODE, s);
    if (variant == 4) { set_error("gae: the TMA variant needs N %% 128 == 0 and 16-byte aligned arrays"); return TRL_EUNSUPPORTED; }
  }
  // VEC=4 needs N % 4 == 0 and 16B/4B-aligned bases; use it once it still fills >= 2 waves of CTAs.
  const bool vec_ok = (p.N % 4 == 0) && aligned16(p.rewards) && aligned16(p.values) && aligned16(p.advs) &&
                      aligned16(p.rets) && aligned16(p.last_value) && aligned4(p.terminals) &&
                      aligned4(p.time_limits);
  const bool want_vec = (variant == 2) || (variant == 1 && p.N >= 128LL * 2 * kNumSM);
  if (variant == 3 || !(vec_ok && want_vec)) return launch_chunked<MODE, 1, 4>(p, s);
  return launch_chunked<MODE, 4, 4>(p, s);
}

}  // namespace trl

// C-ABI ------------------------------------------------------------------------------
TRL_API int trl_gae_scan(const float* rewards, const float* values, const uint8_t* terminals,
                            const uint8_t* time_limits, const float* last_value, float* advs, float* returns,
                            int64_t T, int64_t N, float gamma, float tau, int time_limit_filter, int variant,
                            void* stream) {
  using namespace trl;
  TRL_REQUIRE(T >= 0 && N >= 0, "trl_gae_scan: negative size T=%lld N=%lld", (long long)T, (long long)N);
  if (T == 0 || N == 0) return TRL_OK;
  TRL_REQUIRE(rewards && values && terminals && time_limits && last_value && advs && returns,
              "trl_gae_scan: null pointer");
  TRL_REQUIRE(variant >= 0 && variant <= 4, "trl_gae_scan: variant %d not in 0..4", variant);
  GaeParams p{rewards, values, terminals, time_limits, last_value, advs, returns, T, N, gamma, gamma * tau,
              time_limit_filter};
  return dispatch<MODE_GAE>(p, variant, static_cast<cudaStream_t>(stream));
}

TRL_API int trl_discount_return(const float* rewards, const float* values, const uint8_t* terminals,
                                   const uint8_t* time_limits, const float* last_value, float* advs,
                                   float* returns, int64_t T, int64_t N, float gamma, int time_limit_filter,
                                   int variant, void* stream) {
  using namespace trl;
  TRL_REQUIRE(T >= 0 && N >= 0, "trl_discount_return: negative size");
  if (T == 0 || N == 0) return TRL_OK;
  TRL_REQUIRE(rewards && values && terminals && time_limits && last_value && advs && returns,
              "trl_discount_return: null pointer");
  TRL_REQUIRE(variant >= 0 && variant <= 4, "trl_discount_return: variant %d not in 0..4", variant);
  GaeParams p{rewards, values, terminals, time_limits, last_value, advs, returns, T, N, gamma, gamma,
              time_limit_filter};
  return dispatch<MODE_DISC>(p, variant, static_cast<cudaStream_t>(stream));
}
