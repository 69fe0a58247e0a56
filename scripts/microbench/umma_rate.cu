// umma_rate.cu — issue-rate microbenchmark for tcgen05.mma.kind::tf32 (cta_group::1, M=128): how many cycles
// does one MMA cost as a function of N, operand form (A from smem "SS" / from TMEM "TS") and accumulator
// dependence?  One CTA per SM (grid = 148), one issuing thread, operands are whatever is in shared memory
// (values irrelevant), 512 back-to-back MMAs between two clock64() reads closed by a tcgen05.commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench/umma_rate scripts/microbench/umma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // K-major SWIZZLE_64B, rows of 64 B
  return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

template <bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int n_acc, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {  // warp-uniform role; one elected lane issues (the pattern CUTLASS uses)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a_desc = make_desc(base), b_desc = make_desc(base + 16384);
    const uint32_t d0 = tmem, d1 = tmem + (uint32_t)((n_acc > 1) ? N : 0);
    uint32_t elected = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(elected));
    const long long t0 = clock64();
    if (elected) {
#pragma unroll 1
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t d = (u & 1) ? d1 : d0;
          if (TS) {
            const uint32_t a_t = tmem + 480u + 8u * u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(b_desc),
                         "r"(idesc), "r"(1u) : "memory");
          } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc + 2 * (u & 1)),
                         "l"(b_desc + 2 * (u & 1)), "r"(idesc), "r"(1u) : "memory");
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0,1,0,p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0 && elected) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// cta_group::2 variant: a cluster pair, rank 0 issues M=256 MMAs (128 rows per CTA); B is read half from each
// CTA's shared memory.  Same measurement otherwise.
template <bool TS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate2_kernel(int N, int n_acc, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);
    const uint64_t a_desc = make_desc(base), b_desc = make_desc(base + 16384);
    const uint32_t d0 = tmem, d1 = tmem + (uint32_t)((n_acc > 1) ? N : 0);
    uint32_t elected = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(elected));
    const long long t0 = clock64();
    if (elected) {
#pragma unroll 1
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t d = (u & 1) ? d1 : d0;
          if (TS) {
            const uint32_t a_t = tmem + 480u + 8u * u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(b_desc),
                         "r"(idesc), "r"(1u) : "memory");
          } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc + 2 * (u & 1)),
                         "l"(b_desc + 2 * (u & 1)), "r"(idesc), "r"(1u) : "memory");
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0,1,0,p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0 && elected) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 1024 + 65536;
  cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 512;
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {32, 64, 128, 256})
      for (int n_acc : {1, 2}) {
        if (n_acc * N > 448) continue;
        for (int grid : {1, 148}) {
          long long h = 0;
          for (int rep = 0; rep < 2; ++rep) {
            if (ts) rate_kernel<true><<<grid, 128, smem>>>(N, n_acc, iters, d);
            else rate_kernel<false><<<grid, 128, smem>>>(N, n_acc, iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          printf("%s N=%3d accumulators=%d grid=%3d : %.1f cycles per MMA (floor formula %d)\n", ts ? "TS" : "SS", N, n_acc,
                 grid, (double)h / iters, 128 * N / 256);
        }
      }
  cudaFuncSetAttribute(rate2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {64, 128, 256})
      for (int n_acc : {1, 2}) {
        if (n_acc * N > 448) continue;
        for (int grid : {2, 148}) {
          long long h = 0;
          for (int rep = 0; rep < 2; ++rep) {
            if (ts) rate2_kernel<true><<<grid, 128, smem>>>(N, n_acc, iters, d);
            else rate2_kernel<false><<<grid, 128, smem>>>(N, n_acc, iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          printf("cta_group::2 %s N=%3d accumulators=%d grid=%3d : %.1f cycles per M=256 MMA (one-SM floor formula %d)\n",
                 ts ? "TS" : "SS", N, n_acc, grid, (double)h / iters, 128 * N / 256);
        }
      }
  return 0;
}
