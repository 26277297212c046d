// Does CUDA-core fp64 (or fp32) math slow down while the tensor pipe runs tcgen05 int8 MMAs?
// One CTA per SM: warp 0 (optionally) issues back-to-back UTCIMMA on arbitrary smem data,
// warps 1..12 run DFMA / FFMA chains.  Reports math throughput with and without the MMA stream.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int OP>
__global__ void __launch_bounds__(416, 1) k(double* out, long long* cyc, int iters, int mma_iters) {
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
  long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0 && mma_iters > 0) {
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(192 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t A = make_desc(smem_u32(smem)), B = make_desc(smem_u32(smem) + 32768);
      for (int i = 0; i < mma_iters; ++i) {
        mma_i8(tb, A, B, idesc, 1); mma_i8(tb + 192, A, B, idesc, 1);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t done = 0;
      while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      cyc[blockIdx.x * 2 + 0] = clock64() - t0;
    }
  } else {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    float f0 = threadIdx.x * 1e-3f, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
    const double c = 1.0000001, d = 1e-9;
    for (int i = 0; i < iters; ++i) {
      if (OP == 0) { a0 = fma(a0, c, d); a1 = fma(a1, c, d); a2 = fma(a2, c, d); a3 = fma(a3, c, d); }
      else if (OP == 1) { f0 = fmaf(f0, 1.0000001f, 1e-9f); f1 = fmaf(f1, 1.0000001f, 1e-9f); f2 = fmaf(f2, 1.0000001f, 1e-9f); f3 = fmaf(f3, 1.0000001f, 1e-9f); }
      else if (OP == 2) { f0 = __shfl_sync(0xffffffffu, f0, (lane + 1) & 31); f1 = __shfl_sync(0xffffffffu, f1, (lane + 2) & 31); f2 = __shfl_sync(0xffffffffu, f2, (lane + 3) & 31); f3 = __shfl_sync(0xffffffffu, f3, (lane + 4) & 31); }
      else if (OP == 3) { f0 = __double2float_rn(a0 + f0); f1 = __double2float_rn(a1 + f1); f2 = __double2float_rn(a2 + f2); f3 = __double2float_rn(a3 + f3); }
      else if (OP == 4) { f0 = f0 + f1; f1 = f1 - f2; f2 = f2 + f3; f3 = f3 - f0; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + f0 + f1 + f2 + f3;
    if (threadIdx.x == 32) cyc[blockIdx.x * 2 + 1] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}
template <int OP> void run(const char* name, int iters, int mma_iters) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; long long* cyc; cudaMalloc(&out, sizeof(double) * sms * 416); cudaMalloc(&cyc, sizeof(long long) * sms * 2);
  cudaMemset(cyc, 0, sizeof(long long) * sms * 2);
  cudaFuncSetAttribute(k<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  k<OP><<<sms, 416, 96 * 1024>>>(out, cyc, 10, 1); cudaDeviceSynchronize();
  k<OP><<<sms, 416, 96 * 1024>>>(out, cyc, iters, mma_iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double math_ops = 12.0 * 32 * iters * 4;
  printf("%-6s math iters=%6d mma=%6d | math warps: %9lld cyc -> %.1f lanes/clk/SM | mma warp: %9lld cyc (ideal %d) %s\n", name, iters, mma_iters,
         h[1], h[1] ? math_ops / h[1] : 0.0, h[0], mma_iters * 2 * 96, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("DFMA", 20000, 0);      // fp64 alone
  run<0>("DFMA", 20000, 20000);  // fp64 + MMA stream of comparable duration
  run<0>("DFMA", 10, 20000);     // MMA alone
  run<1>("FFMA", 40000, 0);
  run<1>("FFMA", 40000, 20000);
  run<1>("FFMA", 80000, 20000);
  run<2>("SHFL", 10000, 0);
  run<2>("SHFL", 10000, 20000);
  run<3>("F2F+DADD", 10000, 0);
  run<3>("F2F+DADD", 10000, 20000);
  run<4>("FADD", 40000, 0);
  run<4>("FADD", 80000, 20000);
  return 0;
}
