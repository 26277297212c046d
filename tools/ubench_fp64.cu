// Micro-benchmark: sustained throughput of the CUDA-core pipes the epilogue uses
// (DFMA / DADD / FFMA / MUFU.RCP / I2F.F64) with ample ILP and occupancy.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  float f0 = threadIdx.x * 1e-3f, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
  const double c = 1.0000001, d = 1e-9;
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) { a0 = fma(a0, c, d); a1 = fma(a1, c, d); a2 = fma(a2, c, d); a3 = fma(a3, c, d); a4 = fma(a4, c, d); a5 = fma(a5, c, d); a6 = fma(a6, c, d); a7 = fma(a7, c, d); }
    if (OP == 1) { a0 += d; a1 += d; a2 += d; a3 += d; a4 += d; a5 += d; a6 += d; a7 += d; }
    if (OP == 2) { f0 = fmaf(f0, 1.0000001f, 1e-9f); f1 = fmaf(f1, 1.0000001f, 1e-9f); f2 = fmaf(f2, 1.0000001f, 1e-9f); f3 = fmaf(f3, 1.0000001f, 1e-9f); f4 = fmaf(f4, 1.0000001f, 1e-9f); f5 = fmaf(f5, 1.0000001f, 1e-9f); f6 = fmaf(f6, 1.0000001f, 1e-9f); f7 = fmaf(f7, 1.0000001f, 1e-9f); }
    if (OP == 3) { f0 = __frcp_rn(f0 + 1.f); f1 = __frcp_rn(f1 + 1.f); f2 = __frcp_rn(f2 + 1.f); f3 = __frcp_rn(f3 + 1.f); f4 = __frcp_rn(f4 + 1.f); f5 = __frcp_rn(f5 + 1.f); f6 = __frcp_rn(f6 + 1.f); f7 = __frcp_rn(f7 + 1.f); }
    if (OP == 4) { a0 = (double)(float)a0 + d; a1 = (double)(float)a1 + d; a2 = (double)(float)a2 + d; a3 = (double)(float)a3 + d; a4 = (double)(float)a4 + d; a5 = (double)(float)a5 + d; a6 = (double)(float)a6 + d; a7 = (double)(float)a7 + d; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
}
template <int OP> void run(const char* name, int warps_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
  const int iters = 20000, threads = warps_per_sm * 32;
  k<OP><<<sms, threads>>>(out, 100); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<OP><<<sms, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)sms * threads * iters * 8;
  printf("%-10s warps/SM=%2d  %.3f ms  %.2f Tops/s  = %.1f lanes/clk/SM @1.965GHz\n", name, warps_per_sm, ms, ops / ms / 1e9, ops / (ms * 1e-3) / sms / 1.965e9);
  cudaFree(out);
}
int main() {
  for (int w : {4, 12, 32}) { run<0>("DFMA", w); run<1>("DADD", w); run<2>("FFMA", w); run<3>("MUFU.RCP", w); run<4>("F2F pair", w); }
  return 0;
}
