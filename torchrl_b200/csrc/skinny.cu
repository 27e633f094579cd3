// skinny.cu -- the "skinny" Linear layers of the small MLPs: first layer (K = obs_dim, e.g. 17 -> 256) and
// output layer (256 -> act_dim / 1).  K3/K8 support (networks/base.py:24-44, networks/nets.py:13-52).
//
// These GEMMs have one tiny dimension; cuBLAS serves them with generic sgemm / gemv kernels at 10-20 us each
// although they only stream one (M x 256) activation matrix (16.8 MB at M = 16384 => ~4 us at HBM/L2 speed).
// Four memory-bound kernels, each reading / writing the big matrix exactly once, bias / activation fused:
//   skinny_k_fwd   : Y (M,H)  = act(X (M,K) . W (H,K)^T + b)                    K <= 128
//   skinny_tn      : Out (H,K) = A (M,H)^T . B (M,K)  [+ column sums of B]       K <= 32   (both wgrads)
//   skinny_n_fwd   : Y (M,N)  = X (M,H) . W (N,H)^T + b                          N <= 8
//   skinny_n_dgrad : dX (M,H) = G (M,N) . W (N,H)                                N <= 8
// All fp32 FFMA; reductions are two-level with a fixed combination order (deterministic).
#include "common.cuh"

namespace trl {

__device__ __forceinline__ float sk_act(float x, int act) {
  if (act == 1) return tanhf(x);
  if (act == 2) return fmaxf(x, 0.f);
  return x;
}

// ------------------------------------------------------------------------------------------------- skinny_k_fwd
// CTA: 256 threads = 64 column groups (x4 columns) x 4 rows per pass; kKfRows rows per CTA.
// dynamic smem: Wt[K][H] (transposed weights) + Xs[kKfRows][K]
constexpr int kKfRows = 32;

__global__ void __launch_bounds__(256) skinny_k_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ Y,
                                                          long long M, int K, int H, int act) {
  extern __shared__ float sm[];
  float* Wt = sm;              // [K][H]
  float* Xs = sm + K * H;      // [kKfRows][K]
  const int tid = threadIdx.x;
  for (int i = tid; i < H * K; i += 256) {
    const int h = i / K, k = i - h * K;
    Wt[k * H + h] = W[i];
  }
  const long long row0 = static_cast<long long>(blockIdx.x) * kKfRows;
  const int nrows = static_cast<int>(min(static_cast<long long>(kKfRows), M - row0));
  for (int i = tid; i < nrows * K; i += 256) Xs[i] = X[row0 * K + i];
  __syncthreads();
  const int cg = tid & 63, rr = tid >> 6;
  for (int c0 = cg * 4; c0 < H; c0 += 256) {
    const float4 bb = *reinterpret_cast<const float4*>(bias + c0);
    for (int r = rr; r < nrows; r += 4) {
      float4 acc = bb;
      const float* xr = Xs + r * K;
      for (int k = 0; k < K; ++k) {
        const float xv = xr[k];
        const float4 w = *reinterpret_cast<const float4*>(Wt + k * H + c0);
        acc.x = fmaf(xv, w.x, acc.x); acc.y = fmaf(xv, w.y, acc.y);
        acc.z = fmaf(xv, w.z, acc.z); acc.w = fmaf(xv, w.w, acc.w);
      }
      acc.x = sk_act(acc.x, act); acc.y = sk_act(acc.y, act); acc.z = sk_act(acc.z, act); acc.w = sk_act(acc.w, act);
      *reinterpret_cast<float4*>(Y + (row0 + r) * H + c0) = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------- skinny_tn
// Out[h][k] = sum_m A[m][h] * B[m][k]  (A: M x H, H a multiple of 256 per grid.y block; B: M x K, K <= 32)
// CTA (blockIdx.x = row slab, blockIdx.y = 256-column block of A): thread h accumulates K values over the slab
// -> partial[(slab, h, k)]; skinny_tn_reduce_kernel then sums the slabs in a fixed order.
// Optional: colsum[k] = sum_m B[m][k] (bias gradient of the output layer), computed by column block 0.
constexpr int kTnRows = 128;
constexpr int kTnMaxK = 32;

template <int KMAX>
__global__ void __launch_bounds__(256) skinny_tn_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       int want_colsum, float* __restrict__ partial, long long M, int H,
                                                       int K) {
  __shared__ float Bs[kTnRows * kTnMaxK];
  const int tid = threadIdx.x;
  const int h = blockIdx.y * 256 + tid;
  const long long row0 = static_cast<long long>(blockIdx.x) * kTnRows;
  const int nrows = static_cast<int>(min(static_cast<long long>(kTnRows), M - row0));
  for (int i = tid; i < nrows * K; i += 256) Bs[i] = B[row0 * K + i];
  __syncthreads();
  float acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
  if (h < H) {
#pragma unroll 4
    for (int r = 0; r < nrows; ++r) {
      const float a = A[(row0 + r) * H + h];
      const float* br = Bs + r * K;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) acc[k] = fmaf(a, br[k], acc[k]);
    }
  }
  // partial layout: [slab][H + 1][K]  (row H holds the column sums of B)
  float* pp = partial + static_cast<long long>(blockIdx.x) * (H + 1) * K;
  if (h < H) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) pp[h * K + k] = acc[k];
  }
  if (want_colsum && blockIdx.y == 0 && tid < K) {
    float s = 0.f;
    for (int r = 0; r < nrows; ++r) s += Bs[r * K + tid];
    pp[H * K + tid] = s;
  }
}

// second stage: sum the row-slab partials in fixed order.  partial: [nslab][(H+1)*K]; element e = h*K + k.
__global__ void __launch_bounds__(256) skinny_tn_reduce_kernel(const float* __restrict__ partial, float* __restrict__ Out,
                                                              float* __restrict__ colsum, int nslab, int H, int K,
                                                              int out_transposed) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_main = H * K, n_all = (H + 1) * K;
  if (e >= (colsum ? n_all : n_main)) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int sl = 0;
  for (; sl + 3 < nslab; sl += 4) {
    s0 += partial[static_cast<long long>(sl) * n_all + e];
    s1 += partial[static_cast<long long>(sl + 1) * n_all + e];
    s2 += partial[static_cast<long long>(sl + 2) * n_all + e];
    s3 += partial[static_cast<long long>(sl + 3) * n_all + e];
  }
  for (; sl < nslab; ++sl) s0 += partial[static_cast<long long>(sl) * n_all + e];
  const float s = (s0 + s1) + (s2 + s3);
  if (e < n_main) {
    const int h = e / K, k = e - h * K;
    if (out_transposed) Out[static_cast<long long>(k) * H + h] = s;   // Out is (K, H)
    else Out[e] = s;                                                   // Out is (H, K)
  } else {
    colsum[e - n_main] = s;
  }
}

// ------------------------------------------------------------------------------------------------- skinny_n_fwd
// Y[m][n] = b[n] + sum_h X[m][h] * W[n][h];  one warp per row, N <= 8, H % 4 == 0.
constexpr int kNMax = 8;

__global__ void __launch_bounds__(256) skinny_n_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ Y,
                                                          long long M, int H, int N) {
  extern __shared__ float sm[];          // W [N][H]
  for (int i = threadIdx.x; i < N * H; i += blockDim.x) sm[i] = W[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (long long m = static_cast<long long>(blockIdx.x) * wpb + wib; m < M; m += static_cast<long long>(gridDim.x) * wpb) {
    float acc[kNMax];
#pragma unroll
    for (int n = 0; n < kNMax; ++n) acc[n] = 0.f;
    const float* xr = X + m * H;
    for (int c = lane * 4; c < H; c += 128) {
      const float4 xv = *reinterpret_cast<const float4*>(xr + c);
#pragma unroll
      for (int n = 0; n < kNMax; ++n) {
        if (n < N) {
          const float4 w = *reinterpret_cast<const float4*>(sm + n * H + c);
          acc[n] = fmaf(xv.x, w.x, fmaf(xv.y, w.y, fmaf(xv.z, w.z, fmaf(xv.w, w.w, acc[n]))));
        }
      }
    }
#pragma unroll
    for (int n = 0; n < kNMax; ++n)
      if (n < N) acc[n] = warp_sum(acc[n]);
    if (lane == 0) {
#pragma unroll
      for (int n = 0; n < kNMax; ++n)
        if (n < N) Y[m * N + n] = acc[n] + bias[n];
    }
  }
}

// ------------------------------------------------------------------------------------------------- skinny_n_dgrad
// dX[m][h] = sum_n G[m][n] * W[n][h];  thread per (m, 4 columns), N <= 8, H % 4 == 0.
__global__ void __launch_bounds__(256) skinny_n_dgrad_kernel(const float* __restrict__ G, const float* __restrict__ W,
                                                            float* __restrict__ dX, long long M, int H, int N) {
  extern __shared__ float sm[];          // W [N][H]
  for (int i = threadIdx.x; i < N * H; i += blockDim.x) sm[i] = W[i];
  __syncthreads();
  const int hq = H / 4;
  const long long total = M * hq;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / hq;
    const int c = static_cast<int>(i - m * hq) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int n = 0; n < kNMax; ++n) {
      if (n < N) {
        const float g = G[m * N + n];
        const float4 w = *reinterpret_cast<const float4*>(sm + n * H + c);
        acc.x = fmaf(g, w.x, acc.x); acc.y = fmaf(g, w.y, acc.y); acc.z = fmaf(g, w.z, acc.z); acc.w = fmaf(g, w.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(dX + m * H + c) = acc;
  }
}

}  // namespace trl

TRL_API int trl_skinny_k_fwd(const float* X, const float* W, const float* bias, float* Y, int64_t M, int K, int H,
                             int act, void* stream) {
  using namespace trl;
  TRL_REQUIRE(M >= 1 && K >= 1 && K <= 128 && H >= 4 && H % 4 == 0, "trl_skinny_k_fwd: need 1<=K<=128, H%%4==0 (K=%d H=%d)", K, H);
  TRL_REQUIRE(X && W && bias && Y, "trl_skinny_k_fwd: null pointer");
  TRL_REQUIRE(aligned16(W) && aligned16(bias) && aligned16(Y), "trl_skinny_k_fwd: W/bias/Y must be 16-byte aligned");
  const size_t smem = sizeof(float) * (static_cast<size_t>(K) * H + static_cast<size_t>(kKfRows) * K);
  TRL_REQUIRE(smem <= 200 * 1024, "trl_skinny_k_fwd: K*H too large for shared memory");
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    const cudaError_t e = cudaFuncSetAttribute(skinny_k_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr = smem;
  }
  skinny_k_fwd_kernel<<<static_cast<unsigned>(ceil_div<long long>(M, kKfRows)), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      X, W, bias, Y, M, K, H, act);
  return check_launch("skinny_k_fwd_kernel");
}

TRL_API int64_t trl_skinny_tn_scratch_floats(int64_t M, int H, int K) {
  return trl::ceil_div<long long>(M, trl::kTnRows) * (H + 1) * K;
}

// Out = A^T B: A (M,H), B (M,K<=32).  out_transposed=0: Out (H,K); =1: Out (K,H).  colsum: NULL or (K) = column sums of B.
// scratch: trl_skinny_tn_scratch_floats(M,H,K) floats.
TRL_API int trl_skinny_tn(const float* A, const float* B, float* Out, float* colsum, int64_t M, int H, int K,
                          int out_transposed, float* scratch, void* stream) {
  using namespace trl;
  TRL_REQUIRE(M >= 1 && H >= 1 && K >= 1 && K <= kTnMaxK, "trl_skinny_tn: need 1<=K<=32 (K=%d)", K);
  TRL_REQUIRE(A && B && Out && scratch, "trl_skinny_tn: null pointer");
  const int nslab = static_cast<int>(ceil_div<long long>(M, kTnRows));
  const dim3 grid(static_cast<unsigned>(nslab), static_cast<unsigned>(ceil_div(H, 256)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int wc = colsum ? 1 : 0;
  if (K <= 8) skinny_tn_kernel<8><<<grid, 256, 0, st>>>(A, B, wc, scratch, M, H, K);
  else if (K <= 16) skinny_tn_kernel<16><<<grid, 256, 0, st>>>(A, B, wc, scratch, M, H, K);
  else if (K <= 24) skinny_tn_kernel<24><<<grid, 256, 0, st>>>(A, B, wc, scratch, M, H, K);
  else skinny_tn_kernel<32><<<grid, 256, 0, st>>>(A, B, wc, scratch, M, H, K);
  int rc = check_launch("skinny_tn_kernel");
  if (rc != TRL_OK) return rc;
  const int n_out = (H + (colsum ? 1 : 0)) * K;
  skinny_tn_reduce_kernel<<<ceil_div(n_out, 256), 256, 0, st>>>(scratch, Out, colsum, nslab, H, K, out_transposed);
  return check_launch("skinny_tn_reduce_kernel");
}

TRL_API int trl_skinny_n_fwd(const float* X, const float* W, const float* bias, float* Y, int64_t M, int H, int N,
                             void* stream) {
  using namespace trl;
  TRL_REQUIRE(M >= 1 && N >= 1 && N <= kNMax && H >= 4 && H % 4 == 0, "trl_skinny_n_fwd: need N<=8, H%%4==0 (N=%d H=%d)", N, H);
  TRL_REQUIRE(X && W && bias && Y, "trl_skinny_n_fwd: null pointer");
  TRL_REQUIRE(aligned16(X) && aligned16(W), "trl_skinny_n_fwd: X/W must be 16-byte aligned");
  long long blocks = ceil_div<long long>(M, 8);
  if (blocks > 8LL * kNumSM) blocks = 8LL * kNumSM;
  skinny_n_fwd_kernel<<<static_cast<unsigned>(blocks), 256, sizeof(float) * N * H, static_cast<cudaStream_t>(stream)>>>(
      X, W, bias, Y, M, H, N);
  return check_launch("skinny_n_fwd_kernel");
}

TRL_API int trl_skinny_n_dgrad(const float* G, const float* W, float* dX, int64_t M, int H, int N, void* stream) {
  using namespace trl;
  TRL_REQUIRE(M >= 1 && N >= 1 && N <= kNMax && H >= 4 && H % 4 == 0, "trl_skinny_n_dgrad: need N<=8, H%%4==0");
  TRL_REQUIRE(G && W && dX, "trl_skinny_n_dgrad: null pointer");
  TRL_REQUIRE(aligned16(W) && aligned16(dX), "trl_skinny_n_dgrad: W/dX must be 16-byte aligned");
  long long blocks = ceil_div<long long>(M * (H / 4), 256);
  if (blocks > 8LL * kNumSM) blocks = 8LL * kNumSM;
  skinny_n_dgrad_kernel<<<static_cast<unsigned>(blocks), 256, sizeof(float) * N * H, static_cast<cudaStream_t>(stream)>>>(
      G, W, dX, M, H, N);
  return check_launch("skinny_n_dgrad_kernel");
}
