// tcgen05.mma kind::i8, M = 128, K = 32: cycles per MMA against the N of the instruction.
// Question behind it: k_pairs_tensor uses 40 x 30 frame tiles (N = 96 per weight
// class, 480 of the 512 TMEM columns), so the accumulators cannot be double-buffered and the drain
// of a tile serialises with the MMAs of the next.  Tiles of 40 x 15 frames (N = 48, 240 columns) could
// be double-buffered -- if the tensor pipe keeps its rate at N = 48, where the A operand (4 KB of
// shared memory per MMA) is re-read twice as often per accumulated product.
// One CTA per SM, one thread issues `iters` back-to-back MMAs on arbitrary shared-memory data.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mma_n tools/ubench_mma_n.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, SBO = 1024 B
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int N, int DISTINCT>
__global__ void __launch_bounds__(128, 1) k(long long* cyc, int iters) {
  constexpr int NACC = 512 / N < 5 ? 512 / N : 5;  // accumulators cycled through (the kernel has 5 weight classes)
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t base = smem_u32(smem);
    const long long t0 = clock64();
    // 4 operand slices (the K = 32 steps of a 128-byte swizzled row, as the kernel walks them) x nacc accumulators
    // per trip, everything but the trip counter a compile-time constant so that the descriptors sit in uniform
    // registers ahead of time (a loop that derives them per MMA is bound by that chain at ~68 cycles per MMA)
    const uint64_t A0 = make_desc(base), B0 = make_desc(base + 49152);
    for (int i = 0; i < iters; i += 4 * NACC) {
#pragma unroll
      for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) mma_i8(tb + (uint32_t)(a * N), A0 + (uint64_t)(sl * 2 * DISTINCT), B0 + (uint64_t)(sl * 2 * DISTINCT), idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    cyc[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}
template <int N, int DISTINCT> void run(long long* cyc, int sms) {
  const int smem = 96 * 1024;
  constexpr int NACC = 512 / N < 5 ? 512 / N : 5;
  const int iters = 4 * NACC * 2000;
  cudaFuncSetAttribute(k<N, DISTINCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, DISTINCT><<<sms, 128, smem>>>(cyc, 4 * NACC); cudaDeviceSynchronize();
  k<N, DISTINCT><<<sms, 128, smem>>>(cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("N=%3d operand_slices=%d accumulators=%d: %7.2f cycles per MMA (tensor-rate ideal %5.1f) -> %.3f of rate %s\n", N, DISTINCT ? 4 : 1, NACC,
         (double)h / iters, N / 2.0, (N / 2.0) / ((double)h / iters), e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* cyc; cudaMalloc(&cyc, sizeof(long long) * sms);
  run<16, 1>(cyc, sms); run<32, 1>(cyc, sms); run<48, 1>(cyc, sms); run<64, 1>(cyc, sms); run<96, 1>(cyc, sms);
  run<128, 1>(cyc, sms); run<192, 1>(cyc, sms); run<256, 1>(cyc, sms);
  run<48, 0>(cyc, sms); run<96, 0>(cyc, sms); run<128, 0>(cyc, sms);
  return 0;
}
