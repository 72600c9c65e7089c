// EXPERIMENTAL micro-benchmark (not part of libeadgan.so): issue rate of tcgen05.mma.kind::f16 with cta_group::1 and
// cta_group::2, operands resident in shared memory (no TMA in the loop), N = 256 or 128.  Prints clocks per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 mma_rate.cu -o mma_rate && ./mma_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n\t@P bra.uni WD;\n\tbra.uni WL;\n\tWD:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((16 >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

template <int CG, int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * (16384 + 32768));
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < 4 * (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_ptr;
  if (warp == 0 && lane == 0 && rank == 0) {
    constexpr uint32_t idesc = make_idesc(128 * CG, N);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t sa = smem_u32(smem) + (r & 3) * (16384 + 32768);   // 4 "stages", 4 k-slices each
      const uint32_t sb = sa + 16384;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = make_desc(sa + k * 32), db = make_desc(sb + k * 32);
        if (CG == 1)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                       ::"r"(tmem + (uint32_t)((r & 1) * 256)), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                       ::"r"(tmem + (uint32_t)((r & 1) * 256)), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
      }
    }
    if (CG == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(bar)), "h"((uint16_t)1) : "memory");
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

template <int CG, int N>
void run(int grid, int reps) {
  const int smem = 4 * (16384 + 32768) + 1024 + 64;
  cudaFuncSetAttribute(mma_rate_kernel<CG, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out;
  cudaMalloc(&out, 8);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = CG == 2 ? 1 : 0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) {
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, mma_rate_kernel<CG, N>, reps, out);
    cudaEventRecord(e1);
    cudaError_t er = cudaDeviceSynchronize();
    if (er != cudaSuccess) { printf("cg%d N%d: %s\n", CG, N, cudaGetErrorString(er)); return; }
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long clk; cudaMemcpy(&clk, out, 8, cudaMemcpyDeviceToHost);
  const double macs = (double)grid / CG * reps * 4.0 * (128.0 * CG) * N * 16;
  printf("cta_group::%d M=%d N=%d grid=%3d: %7.1f clk per MMA (CTA 0), kernel %.3f ms -> %.0f TFLOP/s chip\n", CG, 128 * CG, N, grid,
         (double)clk / (reps * 4.0), ms, 2.0 * macs / (ms * 1e-3) / 1e12);
  cudaFree(out);
}

int main() {
  const int reps = 20000;
  run<1, 256>(1, reps); run<2, 256>(2, reps); run<1, 128>(1, reps); run<2, 128>(2, reps);
  run<1, 256>(148, reps); run<2, 256>(148, reps); run<1, 128>(148, reps); run<2, 128>(148, reps);
  return 0;
}
