// EXPERIMENTAL -- NOT part of libeadgan.so, NOT validated on hardware (written at the end of round 1 when the GPU
// budget was spent; it assembles for sm_100a).  First thing to run in round 2:  python tools/try_2cta.py
//
// bf16 GEMM  C[m, n] (fp32) = A[m, k] . B[n, k]^T  with cta_group::2 tiles: a CTA PAIR (cluster of 2, two SMs of one
// TPC) computes a 256 x 256 tile with ONE tcgen05.mma.cta_group::2 stream issued by the leader CTA.  Each CTA
// loads its own 128 A rows and HALF of the B rows (128 of 256): 32 KB per 64-wide k block per SM instead of the 48 KB
// of the cta_group::1 kernels in tc_conv.cu -- the per-SM L2 traffic that caps them.
//
// Protocol (DeepGEMM / CUTLASS sm100 pattern):
//   * both CTAs: tcgen05.alloc.cta_group::2; barriers initialised locally; cluster barrier before use.
//   * producer thread of EACH CTA: waits its own empty[s]; TMA-loads A and its B half into its own shared memory with
//     .cta_group::2 loads whose completion is signalled on the LEADER's full[s] (address mapped with mapa to rank 0).
//     full[s] expects 2 arrivals: the leader's arrive.expect_tx(bytes of BOTH CTAs) and the peer's remote arrive.
//   * leader's MMA thread: waits full[s], issues 4 x tcgen05.mma.cta_group::2 (M 256, N 256, K 16), then
//     tcgen05.commit.cta_group::2 ... multicast::cluster with mask 0b11 -> arrives on empty[s] of BOTH CTAs;
//     after the last k block the same multicast commit signals tmem_full of both CTAs.
//   * epilogue warps of each CTA read their own 128 TMEM lanes x 256 columns and store their 128 rows of C.
//   * cluster barrier, then tcgen05.dealloc.cta_group::2.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx_local(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n\t"
      "@P bra.uni WAIT_DONE;\n\tbra.uni WAIT_LOOP;\n\tWAIT_DONE:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 2-SM TMA load: data lands in THIS CTA's shared memory, the transaction bytes complete on `mbar_cluster_addr`
// (a shared::cluster address: the leader's full barrier)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint32_t mbar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(mbar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
      "r"(accumulate) : "memory");
}
// arrive (once all previously issued MMAs are done) on the barrier at this shared-memory OFFSET in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
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
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int BLOCK_K = 64, STAGES = 6;
constexpr int A_BYTES = 128 * BLOCK_K * 2, B_BYTES = 128 * BLOCK_K * 2, STAGE_BYTES = A_BYTES + B_BYTES;   // per CTA
constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* __restrict__ C,
                 int M, int N, int nkb, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;                     // one 256 x 256 tile per CTA pair
  const int m0 = (pair / n_tiles) * 256 + (int)rank * 128, n0 = (pair % n_tiles) * 256;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
      mbar_init(tmem_full_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm(tmem_ptr, 256);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();          // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        const uint32_t leader_full = mapa(smem_u32(&full_bar[stage]), 0);
        if (leader) mbar_expect_tx_local(&full_bar[stage], 2 * STAGE_BYTES);
        else mbar_arrive_remote(leader_full);
        tma_load_2d_2sm(sa, &map_a, leader_full, kb * BLOCK_K, m0);
        tma_load_2d_2sm(sa + A_BYTES, &map_b, leader_full, kb * BLOCK_K, n0 + (int)rank * 128);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = make_idesc(256, 256);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            umma_bf16_2sm(tmem_base, make_desc(sa + k * 32, 16, 1024), make_desc(sb + k * 32, 16, 1024), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_multicast(&empty_bar[stage], 3);
          if (kb == nkb - 1) umma_commit_multicast(tmem_full_bar, 3);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = m0 + quad * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
      if (row < M) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + c0 + j < N) C[(int64_t)row * N + n0 + c0 + j] = v[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();          // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
  const cuuint64_t gd[2] = {cols, rows}, gs[1] = {cols * 2};
  const cuuint32_t bx[2] = {64, 128}, es[2] = {1, 1};
  return reinterpret_cast<EncodeTiledFn>(p)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gd, gs, bx, es,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -2;
}

}  // namespace

// C[m, n] fp32 = A[m, k] bf16 . B[n, k]^T bf16;  k a multiple of 64.  Returns 0, or a negative code.
extern "C" int eadgan_x_gemm_2cta(const void* a_bf16, const void* b_bf16, float* c_f32, int m, int n, int k, void* stream) {
  if (!a_bf16 || !b_bf16 || !c_f32 || m <= 0 || n <= 0 || k <= 0 || k % 64) return -10;
  CUtensorMap ma, mb;
  if (int e = make_map(&ma, a_bf16, (uint64_t)m, (uint64_t)k)) return e;
  if (int e = make_map(&mb, b_bf16, (uint64_t)n, (uint64_t)k)) return e;
  if (cudaFuncSetAttribute(gemm_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return -3;
  const int m_tiles = (m + 255) / 256, n_tiles = (n + 255) / 256;
  gemm_2cta_kernel<<<2 * m_tiles * n_tiles, 192, SMEM, (cudaStream_t)stream>>>(ma, mb, c_f32, m, n, k / 64, n_tiles);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}
