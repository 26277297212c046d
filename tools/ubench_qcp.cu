// Micro-benchmark of the per-pair solver phases (qcp.cuh) at the epilogue's occupancy.
#include <cstdio>
#include <cuda_runtime.h>
#include "../loos_b200/csrc/qcp.cuh"
using namespace loosb200;
template <int MODE>
__global__ void k(const double* in, double* out, int iters) {
  double R[9];
  for (int i = 0; i < 9; ++i) R[i] = in[i] * (1.0 + 1e-6 * threadIdx.x);
  double e0 = in[9];
  double acc = 0.0;
  for (int it = 0; it < iters; ++it) {
    R[0] += 1e-7 * acc;  // serialise iterations
    if (MODE == 0) {  // poly only
      QcpPoly p = qcp_poly(R, e0);
      acc += p.c0 + p.c1 + p.c2 + p.c3 + p.t0;
    } else if (MODE == 1) {  // poly + newton(6 fixed) + polish
      bool ok; QcpPoly p = qcp_poly(R, e0);
      float t = p.t0; const float f0 = (float)p.c0, f1 = (float)p.c1, f2 = (float)p.c2, f3 = (float)p.c3;
      for (int k2 = 0; k2 < 6; ++k2) qcp_newton_step(f0, f1, f2, f3, t);
      float st=0; acc += qcp_polish(p, t, st, &ok) + (ok ? 0.0 : 1.0);
    } else if (MODE == 2) {  // 4 interleaved
      QcpPoly p[4]; float t[4]; bool ok;
#pragma unroll
      for (int r = 0; r < 4; ++r) { double Rr[9]; for (int i = 0; i < 9; ++i) Rr[i] = R[i] * (1.0 + 1e-3 * r); p[r] = qcp_poly(Rr, e0); t[r] = p[r].t0; }
      for (int k2 = 0; k2 < 6; ++k2)
#pragma unroll
        for (int r = 0; r < 4; ++r) qcp_newton_step((float)p[r].c0, (float)p[r].c1, (float)p[r].c2, (float)p[r].c3, t[r]);
#pragma unroll
      for (int r = 0; r < 4; ++r) acc += qcp_polish(p[r], t[r], 0.f, &ok);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, int per_iter, const double* din) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 384);
  const int iters = 2000;
  k<MODE><<<sms, 384>>>(din, out, 10); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<sms, 384>>>(din, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * 1.965e9;
  printf("%-28s %.3f ms : %.0f cycles per warp-solve-slot (12 warps/SM) ; 1200 solves/SM would take %.0f cycles\n", name, ms,
         cyc / iters / per_iter, cyc / iters / per_iter * (1200.0 / 384.0));
}
int main() {
  double h[10] = {900.1, 12.0, -7.0, 5.0, 700.2, 11.0, -3.0, 9.0, 400.3, 2100.0};
  double* d; cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("poly", 1, d); run<1>("poly+newton6+polish", 1, d); run<2>("4x interleaved full", 4, d);
  return 0;
}
