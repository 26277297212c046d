// Measured dense int8 tensor-pipe peak of the device the roofline is quoted against
// ("builder adds its own measured INT8 dense MMA peak from a tcgen05
// micro-benchmark in the same run").  Every SM pair issues back-to-back
// tcgen05.mma.cta_group::2.kind::i8 with M = 256, N = 256, K = 32 on operands that stay in
// shared memory (pseudo-random bytes, so the multipliers toggle like real data), two accumulators
// filling all 512 TMEM columns, for long enough (seconds) that the figure is the SUSTAINED rate under
// the board's power cap, not a burst.  No global-memory traffic: this is the ceiling of the pipe
// itself -- a real contraction also pays for its operand feed and epilogue.
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace loosb200 {
namespace {

__device__ __forceinline__ uint32_t pk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t pk_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void pk_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void pk_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int PK_A_BYTES = 128 * 128;  // 128 rows x 128 B (four K = 32 slices side by side)
constexpr int PK_B_BYTES = 128 * 128;  // this CTA's half of the N = 256 operand
constexpr int PK_SMEM = PK_A_BYTES + PK_B_BYTES + 1024;

// trips x 8 MMAs per CTA pair
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_int8_peak(int trips, uint32_t seed) {
  extern __shared__ unsigned char pk_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)pk_raw + 1023) & ~(uintptr_t)1023);
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  // operand bytes: a cheap hash, different per CTA
  for (int i = threadIdx.x; i < (PK_A_BYTES + PK_B_BYTES) / 4; i += blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u + seed + blockIdx.x * 40503u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    reinterpret_cast<uint32_t*>(smem)[i] = x;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pk_smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes visible to the MMA's reads
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(pk_smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  pk_cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  if (crank == 0 && warp == 0 && lane == 0) {
    // D = S32, A = B = signed int8, K-major, M = 256 across the pair, N = 256
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t base = pk_smem_u32(smem);
    const uint64_t A0 = pk_desc(base), B0 = pk_desc(base + PK_A_BYTES);
    for (int t = 0; t < trips; ++t) {
#pragma unroll
      for (int acc = 0; acc < 2; ++acc)
#pragma unroll
        for (int sl = 0; sl < 4; ++sl)  // a K = 32 step inside the 128-byte swizzled row: +32 B = +2 in the descriptor
          pk_mma(tb + (uint32_t)(acc * 256), A0 + (uint64_t)(sl * 2), B0 + (uint64_t)(sl * 2), idesc);
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(pk_smem_u32(&bar)), "h"((uint16_t)0x3) : "memory");
  }
  if (warp == 0 && lane == 0) {  // both CTAs: the leader's commit arrives on the barrier of each
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(pk_smem_u32(&bar)), "r"(0u) : "memory");
      if (!done) __nanosleep(256);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  pk_cluster_sync();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tb));
  }
}

}  // namespace
}  // namespace loosb200

using namespace loosb200;

// seconds: how long the measured launch should run (>= 2 for a sustained figure).  tops_out: int8
// tera-operations per second (2 per multiply-add) over all SMs; ms_out: duration of the measured launch.
extern "C" int loosb200_measure_int8_peak(int device, double seconds, double* tops_out, double* ms_out) {
  if (!tops_out) return LOOSB200_ERR_INVALID;
  if (device < 0 || device >= loosb200_device_count()) return LOOSB200_ERR_NO_DEVICE;
  LB_CUDA(cudaSetDevice(device));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const unsigned grid = (unsigned)(sms / 2) * 2;
  LB_CUDA(cudaFuncSetAttribute(k_int8_peak, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM));
  cudaEvent_t e0, e1;
  LB_CUDA(cudaEventCreate(&e0));
  LB_CUDA(cudaEventCreate(&e1));
  auto timed = [&](int trips, float* ms) -> cudaError_t {
    cudaEventRecord(e0, 0);
    k_int8_peak<<<grid, 128, PK_SMEM, 0>>>(trips, 12345u);
    cudaEventRecord(e1, 0);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms, e0, e1);
    return e;
  };
  float ms = 0.f;
  cudaError_t e = timed(2000, &ms);               // warm-up
  int trips = 200000;                             // ~0.1 s calibration
  if (e == cudaSuccess) e = timed(trips, &ms);
  if (e == cudaSuccess && seconds > 0.0) {
    double want = seconds * 1e3 / (ms > 0.f ? ms : 1.f) * trips;
    if (want > 2.0e9) want = 2.0e9;
    if (want < 1000.0) want = 1000.0;
    trips = (int)want;
    e = timed(trips, &ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (e != cudaSuccess) return cuda_fail(e, "int8 peak micro-benchmark", __FILE__, __LINE__);
  const double macs = 256.0 * 256.0 * 32.0 * 8.0 * (double)trips * (double)(grid / 2);
  *tops_out = 2.0 * macs / (ms * 1e-3) / 1e12;
  if (ms_out) *ms_out = ms;
  return LOOSB200_OK;
}
