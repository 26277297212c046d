// Does tcgen05.ld.32x32b accept ANY start column (not a multiple of 16 / 32), and do the .x1 / .x2 forms work at odd
// columns?  k_pairs_tensor wants 33-column groups (11 frames x 3 axes) so that a column tile holds 32 frames in its
// 96 columns instead of 30.  Writes column index * 1000 + lane with tcgen05.st, reads back at awkward offsets.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_tmem_ld tools/ubench_tmem_ld.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(int* bad, long long* cyc) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 512; c0 += 16) {
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) v[c] = (uint32_t)((c0 + c) * 1000 + warp * 32 + lane);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(tl + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  int nbad = 0;
  const int starts[8] = {0, 33, 66, 1, 7, 129, 337, 495};
  for (int s = 0; s < 8; ++s) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tl + starts[s]) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 16; ++c) if (r[c] != (uint32_t)((starts[s] + c) * 1000 + warp * 32 + lane)) ++nbad;
    uint32_t r1, r2a, r2b;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r1) : "r"(tl + starts[s] + 16) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r2a), "=r"(r2b) : "r"(tl + starts[s] + 15) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (r1 != (uint32_t)((starts[s] + 16) * 1000 + warp * 32 + lane)) ++nbad;
    if (r2a != (uint32_t)((starts[s] + 15) * 1000 + warp * 32 + lane) || r2b != (uint32_t)((starts[s] + 16) * 1000 + warp * 32 + lane)) ++nbad;
  }
  // timing: aligned x16 + x16 against x16 + x16 + x1 at an odd start (per warp, 200 rounds over 5 "classes")
  long long t[2];
  for (int mode = 0; mode < 2; ++mode) {
    const uint32_t base = tl + (mode ? 33u : 32u);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < 200; ++it)
      for (int cls = 0; cls < 5; ++cls) {
        uint32_t r[16], q[16], e = 0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(base + cls * 96) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]),
                       "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
                     : "r"(base + cls * 96 + 16) : "memory");
        if (mode) asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(e) : "r"(base + cls * 96 + 32) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 16; ++c) acc += r[c] ^ q[c];
        acc += e;
      }
    t[mode] = clock64() - t0;
    if (acc == 0x12345678u) ++nbad;
  }
  atomicAdd(bad, nbad);
  if (threadIdx.x == 0) { cyc[0] = t[0]; cyc[1] = t[1]; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}
int main() {
  int* bad; long long* cyc;
  cudaMalloc(&bad, 4); cudaMalloc(&cyc, 16); cudaMemset(bad, 0, 4);
  k<<<1, 128>>>(bad, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  int hb; long long hc[2];
  cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
  printf("tcgen05.ld.32x32b at unaligned start columns (0,33,66,1,7,129,337,495), x16 / x1 / x2: mismatches = %d  (%s)\n", hb, cudaGetErrorString(e));
  printf("cycles per 5-class drain of one warp group (4 warps): 32 aligned columns %.1f, 33 columns from column 33 %.1f\n", hc[0] / 200.0, hc[1] / 200.0);
  return hb != 0;
}
