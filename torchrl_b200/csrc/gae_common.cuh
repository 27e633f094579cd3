// gae_common.cuh -- definitions shared by the two K6 kernels (csrc/gae.cu: register-staged chunked scan;
// csrc/gae_tma.cu: persistent scan with TMA-staged tiles): parameters, vector load/store helpers and the
// per-element affine coefficients of
//   /root/reference/torchrl/replay_buffers/on_policy.py:16-44  (generalized_advantage_estimation)
//   /root/reference/torchrl/replay_buffers/on_policy.py:46-70  (discount_reward)
#pragma once
#include "common.cuh"

namespace trl {

enum { MODE_GAE = 0, MODE_DISC = 1 };

struct GaeParams {
  const float* __restrict__ rewards;       // (T,N)
  const float* __restrict__ values;        // (T,N)
  const uint8_t* __restrict__ terminals;   // (T,N) 0/1
  const uint8_t* __restrict__ time_limits; // (T,N) 0/1
  const float* __restrict__ last_value;    // (N)
  float* __restrict__ advs;                // (T,N)
  float* __restrict__ rets;                // (T,N)
  long long T, N;
  float gamma, gamma_tau;
  int filter;
};

template <int VEC> struct VecF;
template <> struct VecF<1> { using type = float; using flag_t = unsigned char; };
template <> struct VecF<4> { using type = float4; using flag_t = unsigned; };

template <int VEC> __device__ __forceinline__ void load_f(const float* p, float (&o)[VEC]);
template <> __device__ __forceinline__ void load_f<1>(const float* p, float (&o)[1]) { o[0] = ld_stream(p); }
template <> __device__ __forceinline__ void load_f<4>(const float* p, float (&o)[4]) {
  const float4 v = ld_stream(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <int VEC> __device__ __forceinline__ unsigned load_flags(const uint8_t* p);
template <> __device__ __forceinline__ unsigned load_flags<1>(const uint8_t* p) { return ld_stream(p); }
template <> __device__ __forceinline__ unsigned load_flags<4>(const uint8_t* p) {
  return ld_stream(reinterpret_cast<const unsigned*>(p));
}
template <int VEC> __device__ __forceinline__ void store_f(float* p, const float (&o)[VEC]);
template <> __device__ __forceinline__ void store_f<1>(float* p, const float (&o)[1]) { st_stream(p, o[0]); }
template <> __device__ __forceinline__ void store_f<4>(float* p, const float (&o)[4]) {
  st_stream(reinterpret_cast<float4*>(p), make_float4(o[0], o[1], o[2], o[3]));
}

// per-element affine coefficients
template <int MODE>
__device__ __forceinline__ void coeffs(float r, float v, float vnext, unsigned term, unsigned tl, float g, float gt,
                                       int filter, float& a, float& b) {
  const float nt = term ? 0.f : 1.f;
  if (MODE == MODE_GAE) {
    const float m = (filter && tl) ? 0.f : 1.f;
    const float delta = r + nt * g * vnext - v;
    a = m * delta;
    b = m * gt * nt;
  } else {
    if (filter) {
      const float tlf = tl ? 1.f : 0.f;
      a = r + tlf * v;
      b = nt * g * (1.f - tlf);
    } else {
      a = r;
      b = nt * g;
    }
  }
}

// b_t alone (a function of the flags only)
template <int MODE>
__device__ __forceinline__ float bcoef(unsigned term, unsigned tl, float g, float gt, int filter) {
  const float nt = term ? 0.f : 1.f;
  if (MODE == MODE_GAE) return ((filter && tl) ? 0.f : 1.f) * gt * nt;
  return filter ? nt * g * (tl ? 0.f : 1.f) : nt * g;
}

// csrc/gae_tma.cu
bool gae_tma_supported(const GaeParams& p);
int gae_tma_launch(const GaeParams& p, int mode, cudaStream_t st);

}  // namespace trl
