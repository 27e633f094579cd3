// skinny.cu -- the "skinny" Linear layers of the small MLPs: first layer (K = obs_dim, e.g. 17 -> 256) and
// output layer (256 -> act_dim / 1).  K3/K8 support (networks/base.py:24-44, networks/nets.py:13-52).
//
// These products have one tiny dimension: each streams ONE big (M x H) activation matrix (16.8 MB at M = 16384,
// H = 256 => ~3 us at HBM speed) and a few KB of everything else; cuBLAS serves them with generic sgemm / gemv
// kernels at 9-14 us each.  Four HBM-bound fp32 kernels, one pass over the big matrix each, sharing one scheme:
//   * the CTA owns a slab of rows; the skinny operand of the slab is staged in shared memory, zero-padded to a
//     multiple of 4 columns so it is read back with broadcast LDS.128;
//   * the weights a thread needs live in REGISTERS for the whole slab (no per-FMA shared-memory traffic);
//   * the big matrix moves as float4, several independent 16-byte accesses in flight per thread.
//   skinny_k_fwd   : Y (M,H)  = act(X (M,K) . W (H,K)^T + b)                    K <= 24, H % 4 == 0, H <= 1024
//   skinny_tn      : Out (H,K) = A (M,H)^T . B (M,K)  [+ column sums of B]       K <= 24, H % 32 == 0, H <= 256
//   skinny_n_fwd   : Y (M,N)  = X (M,H) . W (N,H)^T + b                          N <= 8, H in {128, 256}
//   skinny_n_dgrad : dX (M,H) = G (M,N) . W (N,H)                                N <= 8, H % 4 == 0, H <= 1024
// and two backward fusions that remove a whole pass over the (M x H) matrix each:
//   skinny_act_wgrad   : dW1 = (G * act'(Y))^T X,  db1 = colsum(G * act'(Y))    (first layer: gz never stored)
//   skinny_n_dgrad_act : gz = (G . W) * act'(Y),  db = colsum(gz)               (output-layer dgrad + act backward)
// Reductions have a fixed combination order (deterministic, run-to-run bit-identical).
#include "common.cuh"

namespace trl {

constexpr int kSkMaxRows = 128;      // rows of the skinny operand staged per CTA (<= 12 KB of shared memory)
constexpr int kSkCtas = 2 * kNumSM;  // target grid: two resident CTAs per SM

__device__ __forceinline__ float sk_tanh(float x) {
  // 1 - 2 / (exp(2x) + 1): two MUFU ops; absolute error < 2e-7 (fp32 round-off of the activations themselves)
  x = fminf(fmaxf(x, -15.f), 15.f);
  const float t = __expf(2.f * x);
  return 1.f - __fdividef(2.f, t + 1.f);
}

__device__ __forceinline__ float sk_act(float x, int act) {
  if (act == 1) return sk_tanh(x);
  if (act == 2) return fmaxf(x, 0.f);
  return x;
}

// derivative of the activation expressed through its OUTPUT y (same convention as mlp_epilogue.cu)
__device__ __forceinline__ float sk_dact(float y, int act) {
  if (act == 1) return 1.f - y * y;
  if (act == 2) return y > 0.f ? 1.f : 0.f;
  return 1.f;
}
__device__ __forceinline__ float4 sk_dact4(float4 g, float4 y, int act) {
  return make_float4(g.x * sk_dact(y.x, act), g.y * sk_dact(y.y, act), g.z * sk_dact(y.z, act), g.w * sk_dact(y.w, act));
}

static inline int sk_rows_per_cta(long long M) {
  long long r = ceil_div<long long>(M, kSkCtas);
  if (r < 8) r = 8;
  if (r > kSkMaxRows) r = kSkMaxRows;
  return static_cast<int>(r);
}

// stage rows [row0, row0+nrows) of a row-major (M x K) matrix into shared memory as [nrows][KP], zero padded
template <int KP>
__device__ __forceinline__ void stage_rows(float* __restrict__ dst, const float* __restrict__ src, long long row0,
                                           int nrows, int K, int tid, int nthr) {
  for (int i = tid; i < nrows * KP; i += nthr) {
    const int r = i / KP, k = i - r * KP;
    dst[i] = (k < K) ? src[(row0 + r) * K + k] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------- skinny_k_fwd
// thread = (column group of 4, row lane); W[4 cols][KP] in registers; per row: KP/4 broadcast LDS.128, 4*KP FMA,
// activation, one 16-byte store (a warp writes 512 contiguous bytes).
template <int KP>
__global__ void __launch_bounds__(256, 2) skinny_k_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                             const float* __restrict__ bias, float* __restrict__ Y,
                                                             long long M, int K, int H, int act, int rows_per_cta) {
  extern __shared__ __align__(16) float sk_smem[];
  const int tid = threadIdx.x;
  const long long row0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  if (row0 >= M) return;
  const int nrows = static_cast<int>(min(static_cast<long long>(rows_per_cta), M - row0));
  stage_rows<KP>(sk_smem, X, row0, nrows, K, tid, 256);
  const int cpg = H >> 2;
  const int RL = 256 / cpg;
  const bool active = tid < RL * cpg;
  const int cg = tid % cpg, rl = tid / cpg;
  float w[4][KP];
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < KP; ++k) w[j][k] = (k < K) ? W[static_cast<long long>(4 * cg + j) * K + k] : 0.f;
    bb = *reinterpret_cast<const float4*>(bias + 4 * cg);
  }
  __syncthreads();
  if (!active) return;
  for (int r = rl; r < nrows; r += RL) {
    const float4* xr = reinterpret_cast<const float4*>(sk_smem + r * KP);
    float4 acc = bb;
#pragma unroll
    for (int q = 0; q < KP / 4; ++q) {
      const float4 xv = xr[q];
      acc.x = fmaf(xv.x, w[0][4 * q], acc.x); acc.y = fmaf(xv.x, w[1][4 * q], acc.y);
      acc.z = fmaf(xv.x, w[2][4 * q], acc.z); acc.w = fmaf(xv.x, w[3][4 * q], acc.w);
      acc.x = fmaf(xv.y, w[0][4 * q + 1], acc.x); acc.y = fmaf(xv.y, w[1][4 * q + 1], acc.y);
      acc.z = fmaf(xv.y, w[2][4 * q + 1], acc.z); acc.w = fmaf(xv.y, w[3][4 * q + 1], acc.w);
      acc.x = fmaf(xv.z, w[0][4 * q + 2], acc.x); acc.y = fmaf(xv.z, w[1][4 * q + 2], acc.y);
      acc.z = fmaf(xv.z, w[2][4 * q + 2], acc.z); acc.w = fmaf(xv.z, w[3][4 * q + 2], acc.w);
      acc.x = fmaf(xv.w, w[0][4 * q + 3], acc.x); acc.y = fmaf(xv.w, w[1][4 * q + 3], acc.y);
      acc.z = fmaf(xv.w, w[2][4 * q + 3], acc.z); acc.w = fmaf(xv.w, w[3][4 * q + 3], acc.w);
    }
    acc.x = sk_act(acc.x, act); acc.y = sk_act(acc.y, act); acc.z = sk_act(acc.z, act); acc.w = sk_act(acc.w, act);
    *reinterpret_cast<float4*>(Y + (row0 + r) * H + 4 * cg) = acc;
  }
}

// ------------------------------------------------------------------------------------------------- skinny_tn
// Out[h][k] = sum_m A[m][h] * B[m][k].  One warp owns 32 columns of A: lane = (row lane 0..3) x (column group of 4),
// so one warp-wide LDG.128 fetches four full 128-byte row segments.  acc[4 cols][KP] in registers; the B slab is
// broadcast from shared memory (16 FMA per LDS.128).  After the slab: butterfly over the 4 row lanes, then the CTA
// writes its partial k-major ([K+1][H], row K = column sums of B); skinny_tn_reduce sums the CTAs in a fixed order.
template <int KP>
__device__ __forceinline__ void tn_fma_row(float (&acc)[4][KP], const float4 a, const float* __restrict__ brow) {
#pragma unroll
  for (int q = 0; q < KP / 4; ++q) {
    const float4 b = reinterpret_cast<const float4*>(brow)[q];
    acc[0][4 * q] = fmaf(a.x, b.x, acc[0][4 * q]); acc[0][4 * q + 1] = fmaf(a.x, b.y, acc[0][4 * q + 1]);
    acc[0][4 * q + 2] = fmaf(a.x, b.z, acc[0][4 * q + 2]); acc[0][4 * q + 3] = fmaf(a.x, b.w, acc[0][4 * q + 3]);
    acc[1][4 * q] = fmaf(a.y, b.x, acc[1][4 * q]); acc[1][4 * q + 1] = fmaf(a.y, b.y, acc[1][4 * q + 1]);
    acc[1][4 * q + 2] = fmaf(a.y, b.z, acc[1][4 * q + 2]); acc[1][4 * q + 3] = fmaf(a.y, b.w, acc[1][4 * q + 3]);
    acc[2][4 * q] = fmaf(a.z, b.x, acc[2][4 * q]); acc[2][4 * q + 1] = fmaf(a.z, b.y, acc[2][4 * q + 1]);
    acc[2][4 * q + 2] = fmaf(a.z, b.z, acc[2][4 * q + 2]); acc[2][4 * q + 3] = fmaf(a.z, b.w, acc[2][4 * q + 3]);
    acc[3][4 * q] = fmaf(a.w, b.x, acc[3][4 * q]); acc[3][4 * q + 1] = fmaf(a.w, b.y, acc[3][4 * q + 1]);
    acc[3][4 * q + 2] = fmaf(a.w, b.z, acc[3][4 * q + 2]); acc[3][4 * q + 3] = fmaf(a.w, b.w, acc[3][4 * q + 3]);
  }
}

// ACT = true: A is not read but formed on the fly as G * act'(Yact) (first-layer backward: the activation
// gradient is never written to memory) and row K of the partial receives the column sums of A (the bias gradient).
template <int KP, bool ACT>
__global__ void __launch_bounds__(256, 2) skinny_tn_kernel(const float* __restrict__ A, const float* __restrict__ Yact,
                                                          const float* __restrict__ B, int want_colsum, int act,
                                                          float* __restrict__ partial, long long M, int H, int K,
                                                          int rows_per_cta) {
  extern __shared__ __align__(16) float sk_smem[];
  const int tid = threadIdx.x, nthr = blockDim.x;        // blockDim.x = 32 * (H / 32)
  const long long row0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const int nrows = static_cast<int>(min(static_cast<long long>(rows_per_cta), M - row0));   // grid never overshoots
  stage_rows<KP>(sk_smem, B, row0, nrows, K, tid, nthr);
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const int cg = lane & 7, rl = lane >> 3;
  const int c0 = warp * 32 + cg * 4;
  float acc[4][KP];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < KP; ++k) acc[j][k] = 0.f;
  float4 asum = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* ap = A + row0 * H + c0;
  const float* yp = ACT ? Yact + row0 * H + c0 : nullptr;
  int r = rl;
  if (!ACT) {
    for (; r + 12 < nrows; r += 16) {                    // four independent 16-byte loads in flight per thread
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(ap + static_cast<long long>(r) * H));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(ap + static_cast<long long>(r + 4) * H));
      const float4 a2 = __ldg(reinterpret_cast<const float4*>(ap + static_cast<long long>(r + 8) * H));
      const float4 a3 = __ldg(reinterpret_cast<const float4*>(ap + static_cast<long long>(r + 12) * H));
      tn_fma_row<KP>(acc, a0, sk_smem + r * KP);
      tn_fma_row<KP>(acc, a1, sk_smem + (r + 4) * KP);
      tn_fma_row<KP>(acc, a2, sk_smem + (r + 8) * KP);
      tn_fma_row<KP>(acc, a3, sk_smem + (r + 12) * KP);
    }
  } else {
    for (; r + 4 < nrows; r += 8) {                      // 2 rows x (G, Y): four 16-byte loads in flight
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(ap + static_cast<long long>(r) * H));
      const float4 y0 = __ldg(reinterpret_cast<const float4*>(yp + static_cast<long long>(r) * H));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(ap + static_cast<long long>(r + 4) * H));
      const float4 y1 = __ldg(reinterpret_cast<const float4*>(yp + static_cast<long long>(r + 4) * H));
      const float4 a0 = sk_dact4(g0, y0, act), a1 = sk_dact4(g1, y1, act);
      asum.x += a0.x; asum.y += a0.y; asum.z += a0.z; asum.w += a0.w;
      asum.x += a1.x; asum.y += a1.y; asum.z += a1.z; asum.w += a1.w;
      tn_fma_row<KP>(acc, a0, sk_smem + r * KP);
      tn_fma_row<KP>(acc, a1, sk_smem + (r + 4) * KP);
    }
  }
  for (; r < nrows; r += 4) {
    float4 a0 = __ldg(reinterpret_cast<const float4*>(ap + static_cast<long long>(r) * H));
    if (ACT) {
      a0 = sk_dact4(a0, __ldg(reinterpret_cast<const float4*>(yp + static_cast<long long>(r) * H)), act);
      asum.x += a0.x; asum.y += a0.y; asum.z += a0.z; asum.w += a0.w;
    }
    tn_fma_row<KP>(acc, a0, sk_smem + r * KP);
  }
  // combine the 4 row lanes (lanes l, l^8, l^16, l^24 hold the same columns): fixed order
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      float v = acc[j][k];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      acc[j][k] = v;
    }
  // partial layout per CTA: [K + 1][H] (k-major); row lane (k & 3) stores column-quad k
  float* pp = partial + static_cast<long long>(blockIdx.x) * (K + 1) * H;
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    if (k < K && (k & 3) == rl)
      *reinterpret_cast<float4*>(pp + static_cast<long long>(k) * H + c0) =
          make_float4(acc[0][k], acc[1][k], acc[2][k], acc[3][k]);
  }
  if (ACT) {
    asum.x += __shfl_xor_sync(0xffffffffu, asum.x, 8); asum.x += __shfl_xor_sync(0xffffffffu, asum.x, 16);
    asum.y += __shfl_xor_sync(0xffffffffu, asum.y, 8); asum.y += __shfl_xor_sync(0xffffffffu, asum.y, 16);
    asum.z += __shfl_xor_sync(0xffffffffu, asum.z, 8); asum.z += __shfl_xor_sync(0xffffffffu, asum.z, 16);
    asum.w += __shfl_xor_sync(0xffffffffu, asum.w, 8); asum.w += __shfl_xor_sync(0xffffffffu, asum.w, 16);
    if (rl == (K & 3)) *reinterpret_cast<float4*>(pp + static_cast<long long>(K) * H + c0) = asum;
  } else if (tid < K) {
    float s = 0.f;
    if (want_colsum)
      for (int rr = 0; rr < nrows; ++rr) s += sk_smem[rr * KP + tid];
    pp[static_cast<long long>(K) * H + tid] = s;          // only the first K entries of row K are meaningful
  }
}

// second stage: e indexes the k-major partial ([K][H] then K column sums).  CTA = 32 elements x 8 groups; group g
// sums partials g, g+8, ... (independent loads), then thread g==0 adds the 8 group sums in order.
__global__ void __launch_bounds__(256) skinny_tn_reduce_kernel(const float* __restrict__ partial, float* __restrict__ Out,
                                                              float* __restrict__ colsum, int n_cs, int nslab, int H,
                                                              int K, int out_transposed) {
  __shared__ float red[8][33];
  const int el = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + el;
  const int n_main = K * H;
  const int n_all = n_main + n_cs;                       // n_cs trailing entries of row K go to colsum[]
  const long long stride = static_cast<long long>(K + 1) * H;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (e < n_all) {
    int sl = g;
    for (; sl + 24 < nslab; sl += 32) {
      s0 += partial[static_cast<long long>(sl) * stride + e];
      s1 += partial[static_cast<long long>(sl + 8) * stride + e];
      s2 += partial[static_cast<long long>(sl + 16) * stride + e];
      s3 += partial[static_cast<long long>(sl + 24) * stride + e];
    }
    for (; sl < nslab; sl += 8) s0 += partial[static_cast<long long>(sl) * stride + e];
  }
  red[g][el] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (g == 0 && e < n_all) {
    float s = red[0][el];
#pragma unroll
    for (int i = 1; i < 8; ++i) s += red[i][el];
    if (e < n_main) {
      const int k = e / H, h = e - k * H;
      if (out_transposed) Out[e] = s;                                  // Out is (K, H)
      else Out[static_cast<long long>(h) * K + k] = s;                 // Out is (H, K)
    } else {
      colsum[e - n_main] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------- skinny_n_fwd
// Y[m][n] = b[n] + sum_h X[m][h] * W[n][h].  One warp per row; lane holds W[n][its 4*HC columns] for all n in
// registers (rows n >= N are zero).  Four rows per iteration (4*HC independent 16-byte loads per lane).  The 8
// per-lane partial sums of a row are reduced with a halving butterfly (9 shuffles instead of 40).
__device__ __forceinline__ float n8_butterfly(float (&v)[8], int lane) {
  // after the call the lanes with (lane & 3) == 0 ... all lanes hold the total of output n = (lane >> 2) & 7
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b4 ? v[i] : v[i + 4];
    const float keep = b4 ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b3 ? v[i] : v[i + 2];
    const float keep = b3 ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const float send = b2 ? v[0] : v[1];
    const float keep = b2 ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];       // output index n = 4*bit4 + 2*bit3 + bit2 of the lane id
}

template <int HC>
__global__ void __launch_bounds__(256) skinny_n_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ Y,
                                                          long long M, int H, int N) {
  const int lane = threadIdx.x & 31;
  const long long gw = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const long long nw = static_cast<long long>(gridDim.x) * 8;
  float4 w[8][HC];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int c = 0; c < HC; ++c)
      w[n][c] = (n < N) ? *reinterpret_cast<const float4*>(W + static_cast<long long>(n) * H + c * 128 + lane * 4)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
  const int n_out = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  const float b_out = (n_out < N) ? bias[n_out] : 0.f;
  for (long long m0 = gw * 4; m0 < M; m0 += nw * 4) {
    float4 x[4][HC];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < HC; ++c)
        x[i][c] = (m0 + i < M) ? __ldg(reinterpret_cast<const float4*>(X + (m0 + i) * H + c * 128 + lane * 4))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < HC; ++c)
          s = fmaf(x[i][c].x, w[n][c].x, fmaf(x[i][c].y, w[n][c].y, fmaf(x[i][c].z, w[n][c].z, fmaf(x[i][c].w, w[n][c].w, s))));
        v[n] = s;
      }
      const float tot = n8_butterfly(v, lane);
      if ((lane & 3) == 0 && n_out < N && m0 + i < M) Y[(m0 + i) * N + n_out] = tot + b_out;
    }
  }
}

// ------------------------------------------------------------------------------------------------- skinny_n_dgrad
// dX[m][h] = sum_n G[m][n] * W[n][h].  thread = (column group of 4, row lane), W[n][4 cols] in registers, the G slab
// ([rows][8], zero padded) broadcast from shared memory: 2 LDS.128 + 32 FMA + one 16-byte store per row.
// ACT = true: the result is multiplied by act'(Yact) before it is stored (gz of the last hidden layer) and the
// per-column sums of the stored values (that layer's bias gradient) go to colpart[cta][H].
template <bool ACT>
__global__ void __launch_bounds__(256, 2) skinny_n_dgrad_kernel(const float* __restrict__ G, const float* __restrict__ W,
                                                               const float* __restrict__ Yact, int act,
                                                               float* __restrict__ dX, float* __restrict__ colpart,
                                                               long long M, int H, int N, int rows_per_cta) {
  extern __shared__ __align__(16) float sk_smem[];
  __shared__ __align__(16) float colred[1024];           // [RL][H], RL * H <= 1024
  const int tid = threadIdx.x;
  const long long row0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const int nrows = static_cast<int>(min(static_cast<long long>(rows_per_cta), M - row0));   // grid never overshoots
  stage_rows<8>(sk_smem, G, row0, nrows, N, tid, 256);
  const int cpg = H >> 2;
  const int RL = 256 / cpg;
  const bool active = tid < RL * cpg;
  const int cg = tid % cpg, rl = tid / cpg;
  float4 w[8];
#pragma unroll
  for (int n = 0; n < 8; ++n)
    w[n] = (active && n < N) ? *reinterpret_cast<const float4*>(W + static_cast<long long>(n) * H + 4 * cg)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
    for (int r = rl; r < nrows; r += RL) {
      float4 yv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ACT) yv = __ldg(reinterpret_cast<const float4*>(Yact + (row0 + r) * H + 4 * cg));
      const float4 g0 = reinterpret_cast<const float4*>(sk_smem + r * 8)[0];
      const float4 g1 = reinterpret_cast<const float4*>(sk_smem + r * 8)[1];
      float4 acc;
      acc.x = g0.x * w[0].x; acc.y = g0.x * w[0].y; acc.z = g0.x * w[0].z; acc.w = g0.x * w[0].w;
      acc.x = fmaf(g0.y, w[1].x, acc.x); acc.y = fmaf(g0.y, w[1].y, acc.y); acc.z = fmaf(g0.y, w[1].z, acc.z); acc.w = fmaf(g0.y, w[1].w, acc.w);
      acc.x = fmaf(g0.z, w[2].x, acc.x); acc.y = fmaf(g0.z, w[2].y, acc.y); acc.z = fmaf(g0.z, w[2].z, acc.z); acc.w = fmaf(g0.z, w[2].w, acc.w);
      acc.x = fmaf(g0.w, w[3].x, acc.x); acc.y = fmaf(g0.w, w[3].y, acc.y); acc.z = fmaf(g0.w, w[3].z, acc.z); acc.w = fmaf(g0.w, w[3].w, acc.w);
      acc.x = fmaf(g1.x, w[4].x, acc.x); acc.y = fmaf(g1.x, w[4].y, acc.y); acc.z = fmaf(g1.x, w[4].z, acc.z); acc.w = fmaf(g1.x, w[4].w, acc.w);
      acc.x = fmaf(g1.y, w[5].x, acc.x); acc.y = fmaf(g1.y, w[5].y, acc.y); acc.z = fmaf(g1.y, w[5].z, acc.z); acc.w = fmaf(g1.y, w[5].w, acc.w);
      acc.x = fmaf(g1.z, w[6].x, acc.x); acc.y = fmaf(g1.z, w[6].y, acc.y); acc.z = fmaf(g1.z, w[6].z, acc.z); acc.w = fmaf(g1.z, w[6].w, acc.w);
      acc.x = fmaf(g1.w, w[7].x, acc.x); acc.y = fmaf(g1.w, w[7].y, acc.y); acc.z = fmaf(g1.w, w[7].z, acc.z); acc.w = fmaf(g1.w, w[7].w, acc.w);
      if (ACT) {
        acc = sk_dact4(acc, yv, act);
        cs.x += acc.x; cs.y += acc.y; cs.z += acc.z; cs.w += acc.w;
      }
      *reinterpret_cast<float4*>(dX + (row0 + r) * H + 4 * cg) = acc;
    }
  }
  if (ACT) {
    if (active) *reinterpret_cast<float4*>(colred + rl * H + 4 * cg) = cs;
    __syncthreads();
    if (tid < H) {
      float s = colred[tid];
      for (int i = 1; i < RL; ++i) s += colred[i * H + tid];
      colpart[static_cast<long long>(blockIdx.x) * H + tid] = s;
    }
  }
}

}  // namespace trl

TRL_API int trl_skinny_k_fwd(const float* X, const float* W, const float* bias, float* Y, int64_t M, int K, int H,
                             int act, void* stream) {
  using namespace trl;
  TRL_REQUIRE(M >= 1 && K >= 1 && K <= 24 && H >= 4 && H % 4 == 0 && H <= 1024,
              "trl_skinny_k_fwd: need 1<=K<=24, H%%4==0, H<=1024 (K=%d H=%d)", K, H);
  TRL_REQUIRE(X && W && bias && Y, "trl_skinny_k_fwd: null pointer");
  TRL_REQUIRE(aligned16(bias) && aligned16(Y), "trl_skinny_k_fwd: bias/Y must be 16-byte aligned");
  const int rows = sk_rows_per_cta(M);
  const unsigned grid = static_cast<unsigned>(ceil_div<long long>(M, rows));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int kp = (K + 3) & ~3;
  const size_t smem = sizeof(float) * rows * kp;
#define TRL_KF(KP) skinny_k_fwd_kernel<KP><<<grid, 256, smem, st>>>(X, W, bias, Y, M, K, H, act, rows)
  switch (kp) {
    case 4: TRL_KF(4); break;
    case 8: TRL_KF(8); break;
    case 12: TRL_KF(12); break;
    case 16: TRL_KF(16); break;
    case 20: TRL_KF(20); break;
    default: TRL_KF(24); break;
  }
#undef TRL_KF
  return check_launch("skinny_k_fwd_kernel");
}

TRL_API int64_t trl_skinny_tn_scratch_floats(int64_t M, int H, int K) {
  const int rows = trl::sk_rows_per_cta(M);
  return trl::ceil_div<long long>(M, rows) * (K + 1) * H;
}

static int launch_skinny_tn(const float* A, const float* Yact, const float* B, float* Out, float* colsum, int64_t M,
                            int H, int K, int out_transposed, int act, bool fused_act, float* scratch, void* stream,
                            const char* who) {
  using namespace trl;
  TRL_REQUIRE(M >= 1 && K >= 1 && K <= 24 && H >= 32 && H % 32 == 0 && H <= 256,
              "%s: need 1<=K<=24, H%%32==0, H<=256 (K=%d H=%d)", who, K, H);
  TRL_REQUIRE(A && B && Out && scratch, "%s: null pointer", who);
  TRL_REQUIRE(aligned16(A) && aligned16(scratch) && (!Yact || aligned16(Yact)), "%s: A/Y/scratch must be 16-byte aligned", who);
  const int rows = sk_rows_per_cta(M);
  const int nslab = static_cast<int>(ceil_div<long long>(M, rows));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int wc = colsum ? 1 : 0;
  const int kp = (K + 3) & ~3;
  const size_t smem = sizeof(float) * rows * kp;
  const unsigned nthr = static_cast<unsigned>(H);            // 32 threads per 32 columns
#define TRL_TN(KP)                                                                                              \
  if (fused_act) skinny_tn_kernel<KP, true><<<nslab, nthr, smem, st>>>(A, Yact, B, wc, act, scratch, M, H, K, rows); \
  else skinny_tn_kernel<KP, false><<<nslab, nthr, smem, st>>>(A, nullptr, B, wc, act, scratch, M, H, K, rows)
  switch (kp) {
    case 4: TRL_TN(4); break;
    case 8: TRL_TN(8); break;
    case 12: TRL_TN(12); break;
    case 16: TRL_TN(16); break;
    case 20: TRL_TN(20); break;
    default: TRL_TN(24); break;
  }
#undef TRL_TN
  int rc = check_launch("skinny_tn_kernel");
  if (rc != TRL_OK) return rc;
  const int n_cs = colsum ? (fused_act ? H : K) : 0;
  const int n_all = K * H + n_cs;
  skinny_tn_reduce_kernel<<<ceil_div(n_all, 32), 256, 0, st>>>(scratch, Out, colsum, n_cs, nslab, H, K, out_transposed);
  return check_launch("skinny_tn_reduce_kernel");
}

// Out = A^T B: A (M,H), B (M,K<=24).  out_transposed=0: Out (H,K); =1: Out (K,H).  colsum: NULL or (K) = column sums of B.
// scratch: trl_skinny_tn_scratch_floats(M,H,K) floats.
TRL_API int trl_skinny_tn(const float* A, const float* B, float* Out, float* colsum, int64_t M, int H, int K,
                          int out_transposed, float* scratch, void* stream) {
  return launch_skinny_tn(A, nullptr, B, Out, colsum, M, H, K, out_transposed, 0, false, scratch, stream, "trl_skinny_tn");
}

// First-layer backward in one pass: with gz = G * act'(Y) (never stored),  dW (H,K) = gz^T X  and  db (H) = colsum(gz).
TRL_API int trl_skinny_act_wgrad(const float* G, const float* Y, const float* X, float* dW, float* db, int64_t M, int H,
                                 int K, int act, float* scratch, void* stream) {
  using namespace trl;
  TRL_REQUIRE(Y && db, "trl_skinny_act_wgrad: null pointer");
  TRL_REQUIRE(act >= 0 && act <= 2, "trl_skinny_act_wgrad: unknown activation %d", act);
  return launch_skinny_tn(G, Y, X, dW, db, M, H, K, 0, act, true, scratch, stream, "trl_skinny_act_wgrad");
}

TRL_API int trl_skinny_n_fwd(const float* X, const float* W, const float* bias, float* Y, int64_t M, int H, int N,
                             void* stream) {
  using namespace trl;
  TRL_REQUIRE(M >= 1 && N >= 1 && N <= 8 && (H == 128 || H == 256),
              "trl_skinny_n_fwd: need N<=8 and H in {128, 256} (N=%d H=%d)", N, H);
  TRL_REQUIRE(X && W && bias && Y, "trl_skinny_n_fwd: null pointer");
  TRL_REQUIRE(aligned16(X) && aligned16(W), "trl_skinny_n_fwd: X/W must be 16-byte aligned");
  long long blocks = ceil_div<long long>(M, 8 * 4);          // 8 warps x 4 rows per iteration
  if (blocks > kNumSM) blocks = kNumSM;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (H == 128) skinny_n_fwd_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, st>>>(X, W, bias, Y, M, H, N);
  else skinny_n_fwd_kernel<2><<<static_cast<unsigned>(blocks), 256, 0, st>>>(X, W, bias, Y, M, H, N);
  return check_launch("skinny_n_fwd_kernel");
}

TRL_API int trl_skinny_n_dgrad(const float* G, const float* W, float* dX, int64_t M, int H, int N, void* stream) {
  using namespace trl;
  TRL_REQUIRE(M >= 1 && N >= 1 && N <= 8 && H >= 4 && H % 4 == 0 && H <= 1024,
              "trl_skinny_n_dgrad: need N<=8, H%%4==0, H<=1024");
  TRL_REQUIRE(G && W && dX, "trl_skinny_n_dgrad: null pointer");
  TRL_REQUIRE(aligned16(W) && aligned16(dX), "trl_skinny_n_dgrad: W/dX must be 16-byte aligned");
  const int rows = sk_rows_per_cta(M);
  const unsigned grid = static_cast<unsigned>(ceil_div<long long>(M, rows));
  skinny_n_dgrad_kernel<false><<<grid, 256, sizeof(float) * rows * 8, static_cast<cudaStream_t>(stream)>>>(
      G, W, nullptr, 0, dX, nullptr, M, H, N, rows);
  return check_launch("skinny_n_dgrad_kernel");
}

TRL_API int64_t trl_skinny_dgrad_act_scratch_floats(int64_t M, int H) {
  return trl::ceil_div<long long>(M, trl::sk_rows_per_cta(M)) * H;
}

// Output-layer dgrad fused with the previous layer's activation backward:
//   gz (M,H) = (G (M,N) . W (N,H)) * act'(Y (M,H)),   db (H) = colsum(gz).   scratch: ..._scratch_floats(M,H) floats.
TRL_API int trl_skinny_n_dgrad_act(const float* G, const float* W, const float* Y, float* gz, float* db, int64_t M,
                                   int H, int N, int act, float* scratch, void* stream) {
  using namespace trl;
  TRL_REQUIRE(M >= 1 && N >= 1 && N <= 8 && H >= 4 && H % 4 == 0 && H <= 1024,
              "trl_skinny_n_dgrad_act: need N<=8, H%%4==0, H<=1024");
  TRL_REQUIRE(G && W && Y && gz && db && scratch, "trl_skinny_n_dgrad_act: null pointer");
  TRL_REQUIRE(act >= 0 && act <= 2, "trl_skinny_n_dgrad_act: unknown activation %d", act);
  TRL_REQUIRE(aligned16(W) && aligned16(gz) && aligned16(Y), "trl_skinny_n_dgrad_act: W/Y/gz must be 16-byte aligned");
  const int rows = sk_rows_per_cta(M);
  const int nslab = static_cast<int>(ceil_div<long long>(M, rows));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  skinny_n_dgrad_kernel<true><<<nslab, 256, sizeof(float) * rows * 8, st>>>(G, W, Y, act, gz, scratch, M, H, N, rows);
  int rc = check_launch("skinny_n_dgrad_kernel");
  if (rc != TRL_OK) return rc;
  // column sums: partial [nslab][H] viewed as a K = 0 "tn" partial (stride H, all H entries are colsum entries)
  skinny_tn_reduce_kernel<<<ceil_div(H, 32), 256, 0, st>>>(scratch, db, db, H, nslab, H, 0, 0);
  return check_launch("skinny_tn_reduce_kernel");
}
