// gemm_tf32x3_pair.cu -- WORK IN PROGRESS, NOT PART OF libtorchrl_b200.so (torchrl_b200/build.py does not compile
// this directory).  Draft of the next step named in DESIGN.md section 8: the fp32-faithful 3xTF32 GEMM of
// csrc/gemm_tf32x3.cu on CTA PAIRS (tcgen05 cta_group::2).  It has been cross-compiled for sm_100a (ptxas accepts
// every instruction form) but has NEVER RUN on hardware: build and check it with scripts/gemm_pair_probe.py under a
// short `timeout` before anything else uses it.
//
//   C[M x 256] = act(A[M x K] * B[256 x K]^T + bias)         A, B row-major (K-major operands), M % 256 == 0
//
// Why pairs: the single-CTA kernel is shared-memory-bandwidth bound (per 32-deep K block and CTA: 144 KB of MMA
// operand reads + 144 KB of hi/lo conversion traffic + 48 KB of TMA writes against 1.5 k cycles of tensor work;
// ncu: tensor pipe 23.5 %).  With cta_group::2 one MMA covers 256 x 256: each CTA of the pair stages and converts
// its own 128 rows of A and only HALF of B (128 of the 256 output columns), the tensor cores of both SMs read both
// halves -> per CTA and K block: 96 KB operand reads + 96 KB conversion + 32 KB TMA = 224 KB, under the tensor time.
//
// Roles per CTA (192 threads, identical in both CTAs so that shared-memory offsets match):
//   warp 0      TMA producer : A box (32 x 128 rows) at this CTA's rows, B box (32 x 128 rows) at rows rank*128
//                              -> LOCAL full[s]                        (plain cta_group::1 loads: the consumers of
//                              the raw tiles are this CTA's own converter warps, not the tensor core)
//   warps 2..5  converters   : raw -> (hi in place, lo twin) for the local A tile and the local half of B, then one
//                              arrival per warp on the LEADER's conv[s] (remote arrive through mapa for rank 1)
//   warp 1      MMA issuer   : rank 0 only: waits conv[s] (8 arrivals), issues 12 tcgen05.mma.cta_group::2 per stage,
//                              commits with multicast to empty[s] of BOTH CTAs; last commit -> tmem_full of both
//   warps 2..5  epilogue     : each CTA drains its own 128 accumulator rows from its own TMEM
#include "../common.cuh"
#include <cuda.h>

namespace trlwip {

using namespace trl;

constexpr int kBM = 128, kBNHalf = 128, kBN = 256, kBK = 32;
constexpr int kStages = 3;
constexpr int kUmmaK = 8;
constexpr int kABytes = kBM * kBK * 4;                  // 16 KB
constexpr int kBBytes = kBNHalf * kBK * 4;              // 16 KB (half of B)
constexpr int kStageBytes = 2 * (kABytes + kBBytes);    // 64 KB: A hi | A lo | B hi | B lo
constexpr int kThreads = 192;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 remote;\n\t"
      "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [remote];\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// cluster-scope acquire: the barrier may have been completed by arrivals / commits of the peer CTA
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// K-major SWIZZLE_128B canonical layout (8-row atoms of 1024 B), same as csrc/gemm_tf32x3.cu
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::tf32, fp32 accumulate, K-major A and B, M = 256 (two CTAs x 128 rows), N = 256
__device__ __forceinline__ uint32_t umma_idesc_tf32_256x256() {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(kBN >> 3) << 17) |
         (static_cast<uint32_t>((2 * kBM) >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of this thread -> one arrival on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void split4(const float4 v, float4& h, float4& l) {
  unsigned u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.x)); h.x = __uint_as_float(u); l.x = v.x - h.x;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.y)); h.y = __uint_as_float(u); l.y = v.y - h.y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.z)); h.z = __uint_as_float(u); l.z = v.z - h.z;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.w)); h.w = __uint_as_float(u); l.w = v.w - h.w;
}

struct PairParams {
  const float* __restrict__ bias;
  int act;
  float* __restrict__ C;
  long long M;
  int k_blocks;
  int ldc;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  auto a_hi = [&](int s) { return smem + s * kStageBytes; };
  auto a_lo = [&](int s) { return smem + s * kStageBytes + kABytes; };
  auto b_hi = [&](int s) { return smem + s * kStageBytes + 2 * kABytes; };
  auto b_lo = [&](int s) { return smem + s * kStageBytes + 2 * kABytes + kBBytes; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;                  // [kStages] local TMA -> local converters
  uint64_t* conv = bars + kStages;        // [kStages] converters of both CTAs -> MMA (used in the leader only)
  uint64_t* empty = bars + 2 * kStages;   // [kStages] MMA (multicast commit) -> local TMA
  uint64_t* tmem_full = bars + 3 * kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int m_blk = blockIdx.x;           // this CTA's 128 output rows
  const int nkb = p.k_blocks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&conv[s], 8);           // 4 converter warps x 2 CTAs
        mbar_init(&empty[s], 1);
      }
      mbar_init(tmem_full, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // the same warp of both CTAs allocates: 256 fp32 accumulator columns at the same address in both SMs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                     // barriers of the peer are initialised before anyone arrives on them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], kABytes + kBBytes);
        tma_load_2d(a_hi(s), &map_a, &full[s], kb * kBK, m_blk * kBM);
        tma_load_2d(b_hi(s), &map_b, &full[s], kb * kBK, static_cast<int>(rank) * kBNHalf);
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc_tf32_256x256();
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&conv[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t da_hi = umma_desc_k_sw128(smem_u32(a_hi(s))), da_lo = umma_desc_k_sw128(smem_u32(a_lo(s)));
        const uint64_t db_hi = umma_desc_k_sw128(smem_u32(b_hi(s))), db_lo = umma_desc_k_sw128(smem_u32(b_lo(s)));
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k) {
          const uint64_t adv = static_cast<uint64_t>((k * kUmmaK * 4) >> 4);
          umma_tf32_pair(tmem_base, da_lo + adv, db_hi + adv, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_tf32_pair(tmem_base, da_hi + adv, db_lo + adv, idesc, 1u);
          umma_tf32_pair(tmem_base, da_hi + adv, db_hi + adv, idesc, 1u);
        }
        umma_commit_multicast(&empty[s], 0b11);          // both CTAs may refill their stage s
      }
      umma_commit_multicast(tmem_full, 0b11);            // both accumulator halves are complete
    }
  } else {
    const int ct = threadIdx.x - 64;                     // 0..127
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      mbar_wait(&full[s], ph);
      float4* ah = reinterpret_cast<float4*>(a_hi(s));
      float4* al = reinterpret_cast<float4*>(a_lo(s));
      float4* bh = reinterpret_cast<float4*>(b_hi(s));
      float4* bl = reinterpret_cast<float4*>(b_lo(s));
#pragma unroll
      for (int i = 0; i < kABytes / 16 / 128; ++i) {
        const int c = ct + i * 128;
        float4 h, l;
        split4(ah[c], h, l);
        ah[c] = h;
        al[c] = l;
        split4(bh[c], h, l);
        bh[c] = h;
        bl[c] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive_local(&conv[s]);
        else mbar_arrive_remote(&conv[s], 0);
      }
    }
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int quad = warp & 3;
    const long long row = static_cast<long long>(m_blk) * kBM + quad * 32 + lane;
    float* crow = p.C + row * p.ldc;
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      uint32_t r[32];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < p.M) {
        float4* dst = reinterpret_cast<float4*>(crow + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                 __uint_as_float(r[4 * j + 3]));
          if (p.bias) {
            const float4 b = *reinterpret_cast<const float4*>(p.bias + c * 32 + 4 * j);
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            if (p.act == 1) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
            else if (p.act == 2) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          }
          dst[j] = v;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                     // nobody frees TMEM / exits while the peer still reads or is being read
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t K, uint32_t box_rows) {
  static PFN_encodeTiled enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return false;
    enc = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  const cuuint64_t gdim[2] = {K, rows};
  const cuuint64_t gstride[1] = {K * sizeof(float)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace trlwip

// C (M x 256) = act(A (M x K) . B (256 x K)^T [+ bias]); M % 256 == 0, K % 32 == 0, 16-byte aligned pointers.
extern "C" __attribute__((visibility("default"))) int trl_wip_gemm_tf32x3_nt_pair(const float* A, const float* B, float* C,
                                                                                  int64_t M, int64_t K, const float* bias,
                                                                                  int act, void* stream) {
  using namespace trlwip;
  if (M < 256 || M % 256 != 0 || K < kBK || K % kBK != 0 || !A || !B || !C || act < 0 || act > 2) return 1;
  CUtensorMap map_a, map_b;
  if (!make_map(&map_a, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), kBM) ||
      !make_map(&map_b, B, static_cast<uint64_t>(kBN), static_cast<uint64_t>(K), kBNHalf))
    return 2;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_tf32x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess)
      return 3;
    attr_set = true;
  }
  PairParams p{bias, act, C, M, static_cast<int>(K / kBK), kBN};
  gemm_tf32x3_pair_kernel<<<static_cast<unsigned>(M / kBM), kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(
      map_a, map_b, p);
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}
