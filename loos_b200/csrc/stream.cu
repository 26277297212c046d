// Single-reference streaming path (rmsd2ref and iterative alignment).
//
//   k_cov     3x3 covariance of every frame with the centred reference (alignment::kabsch,
//             src/alignment.cpp:217-233: centre both, dgemm); frame centroids come from staging
//   k_solve   one frame per thread: QCP root of Horn's K, eigenvector -> rotation Z,
//             4x4 W = T(vc) Z T(-uc); optionally the minimal RMSD from the eigenvalue
//   k_resid   x' = W x over the rmsd selection (AtomicGroup::applyTransform,
//             src/AG_numerical.cpp:363-370) and sum |x' - y|^2 (AtomicGroup::rmsd, :301-318)
//   k_accum   sum of x' into the running average (iterativeAlignment / averageStructure,
//             src/alignment.cpp:379-393, src/ensembles.cpp:108-118), k_finish_iter closes an iteration
// HBM-bound: 12 N_align (+ 12 N_rmsd if a different selection) bytes per frame and pass.
#include <string.h>

#include <algorithm>
#include <vector>

#include <stdlib.h>
#include "common.cuh"
#include "qcp.cuh"

namespace loosb200 {

namespace {

constexpr int ST = 128;  // threads per CTA

// Sum of the transformed frames, sum_f (W_f x_f), for iterativeAlignment's running average
// (src/alignment.cpp:385-397) and averageStructure (src/ensembles.cpp:91-121).  The grid is
// (frame walkers) x (atom chunks): a CTA walks frames f = blockIdx.x, blockIdx.x + gridDim.x, ...
// over ITS chunk of at most ACC_CHUNK atoms and keeps the partial sum of that chunk in shared
// memory; a thread owns the float4 groups tid, tid + ST, ... of the chunk, so there is no
// synchronisation inside the frame loop and no atomic until the final flush.  Selections of any
// size work (rmsd2ref's default --rmsd is every heavy atom): more atoms, more chunks.  W_f
// ([F][16], from k_solve) is read straight from global memory (one broadcast line per frame).
constexpr uint32_t ACC_CHUNK = 8192;  // atoms per chunk: 3 x 8192 doubles = 192 KB of shared memory
struct AccumArgs {
  const float* raw; size_t Npad; uint32_t N;  // frame planes [F][3][Npad]
  const double* xf;                           // [F][16]
  uint32_t F;
  uint32_t chunk4;                            // float4 groups per atom chunk (blockIdx.y)
  double* avg;                                // [3][Npad] global accumulators (zeroed by the caller)
};

__global__ void __launch_bounds__(ST) k_accum(const AccumArgs a) {
  extern __shared__ double acc_s[];  // [3][4 * chunk4]
  const int tid = threadIdx.x;
  const uint32_t n4 = (uint32_t)(a.Npad / 4);
  const uint32_t g_lo = blockIdx.y * a.chunk4, g_hi = min(n4, g_lo + a.chunk4);
  const uint32_t clen = 4 * a.chunk4;  // doubles per plane of the chunk
  for (uint32_t i = tid; i < 3 * clen; i += ST) acc_s[i] = 0.0;
  __syncthreads();  // the zeroing and the flush stride over all three planes, the frame loop does not
  double2* ax = reinterpret_cast<double2*>(acc_s);
  double2* ay = reinterpret_cast<double2*>(acc_s + clen);
  double2* az = reinterpret_cast<double2*>(acc_s + 2 * clen);
  for (uint32_t f = blockIdx.x; f < a.F; f += gridDim.x) {
    double W[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) W[k] = __ldg(a.xf + (size_t)f * 16 + k);
    const float4* px = reinterpret_cast<const float4*>(a.raw + (size_t)f * 3 * a.Npad);
    const float4* py = px + n4, *pz = py + n4;
    // two groups per trip, all six loads issued before the first use
    for (uint32_t g0 = g_lo + tid; g0 < g_hi; g0 += 2 * ST) {
      const uint32_t g1 = g0 + ST;
      const bool two = g1 < g_hi;
      const float4 X0 = __ldcs(px + g0), Y0 = __ldcs(py + g0), Z0 = __ldcs(pz + g0);
      float4 X1 = X0, Y1 = Y0, Z1 = Z0;
      if (two) { X1 = __ldcs(px + g1); Y1 = __ldcs(py + g1); Z1 = __ldcs(pz + g1); }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        const uint32_t g = h ? g1 : g0;
        const float4 X = h ? X1 : X0, Y = h ? Y1 : Y0, Z = h ? Z1 : Z0;
        const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
        double tx[4], ty[4], tz[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool real = 4 * g + k < a.N;  // pad atoms stay zero
          const double x = xs[k], y = ys[k], z = zs[k];
          tx[k] = real ? W[0] * x + W[1] * y + W[2] * z + W[3] : 0.0;
          ty[k] = real ? W[4] * x + W[5] * y + W[6] * z + W[7] : 0.0;
          tz[k] = real ? W[8] * x + W[9] * y + W[10] * z + W[11] : 0.0;
        }
        const uint32_t l = g - g_lo;  // group inside the chunk
        double2 u;
        u = ax[2 * l]; u.x += tx[0]; u.y += tx[1]; ax[2 * l] = u;
        u = ax[2 * l + 1]; u.x += tx[2]; u.y += tx[3]; ax[2 * l + 1] = u;
        u = ay[2 * l]; u.x += ty[0]; u.y += ty[1]; ay[2 * l] = u;
        u = ay[2 * l + 1]; u.x += ty[2]; u.y += ty[3]; ay[2 * l + 1] = u;
        u = az[2 * l]; u.x += tz[0]; u.y += tz[1]; az[2 * l] = u;
        u = az[2 * l + 1]; u.x += tz[2]; u.y += tz[3]; az[2 * l + 1] = u;
      }
    }
  }
  __syncthreads();
  const uint32_t here = 4 * (g_hi > g_lo ? g_hi - g_lo : 0);  // doubles of this chunk that exist
  for (uint32_t i = tid; i < 3 * clen; i += ST) {
    const uint32_t plane = i / clen, within = i - plane * clen;
    if (within < here && acc_s[i] != 0.0) atomicAdd(&a.avg[(size_t)plane * a.Npad + 4 * g_lo + within], acc_s[i]);
  }
}

// ---------------------------------------------------------------------------------------
// rmsd2ref, target mode, and the final per-frame rmsd of the iterative mode, as three
// kernels instead of one warp-per-frame loop (covariance -> solve -> residuals):
//   * a fused kernel keeps (#warps x frames per warp x 24 KB) of frames alive between its two
//     passes -- 114 MB at full occupancy, more than L2 holds -- so the second pass re-read HBM
//     (ncu: dram read 2x the frame bytes, L2 hit rate 31 %, long-scoreboard stalls), and the
//     serial 4x4 eigen-solve (~7 K cycles, all 32 lanes redundant) sat inside every warp's loop;
//   * split, both streaming kernels are plain one-pass reductions with ld.cs loads and high
//     occupancy, and the solves run one frame per THREAD;
//   * when the rmsd selection and target are the align selection and target, the minimal RMSD is
//     sqrt((Ga + Gb - 2 lambda_max) / N) and the residual pass is not needed at all.
constexpr int SW_THREADS = 256;

__device__ __forceinline__ double wsum_all(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int SW_FPW = 2;  // frames per warp: every load of the fp64 reference from L1 serves two frames

struct CovArgs {
  const float* raw; size_t Npad;  // frame planes [F][3][Npad]
  const double* ref;              // [3][Npad] centred reference
  uint32_t F;
  double* cov;                    // [F][9] out: sum_k x_p v_q (uncentred x; corrected in k_solve)
};

__global__ void __launch_bounds__(SW_THREADS, 2) k_cov(const CovArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * SW_THREADS + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * SW_THREADS) >> 5;
  const double2* rx = reinterpret_cast<const double2*>(a.ref);
  const double2* ry = reinterpret_cast<const double2*>(a.ref + a.Npad);
  const double2* rz = reinterpret_cast<const double2*>(a.ref + 2 * a.Npad);
  for (uint32_t f0 = warp_global * SW_FPW; f0 < a.F; f0 += nwarps * SW_FPW) {
    const int nf = (int)(a.F - f0 < (uint32_t)SW_FPW ? a.F - f0 : (uint32_t)SW_FPW);
    double s[SW_FPW][9];
#pragma unroll
    for (int ff = 0; ff < SW_FPW; ++ff)
#pragma unroll
      for (int k = 0; k < 9; ++k) s[ff][k] = 0.0;
    // explicit two-deep pipeline: the frame loads of group i4 + 32 are issued before group i4 is
    // consumed (the kernel is latency-bound on HBM: ncu long-scoreboard 14 per issue without it)
    const uint32_t n4 = (uint32_t)(a.Npad / 4);
    const float4* fb[SW_FPW];
#pragma unroll
    for (int ff = 0; ff < SW_FPW; ++ff)
      fb[ff] = reinterpret_cast<const float4*>(a.raw + (size_t)(f0 + (ff < nf ? ff : 0)) * 3 * a.Npad);
    float4 cur[SW_FPW][3], nxt[SW_FPW][3];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ff = 0; ff < SW_FPW; ++ff)
#pragma unroll
      for (int p = 0; p < 3; ++p) cur[ff][p] = lane < n4 ? __ldcs(fb[ff] + p * (a.Npad / 4) + lane) : zero4;
    for (uint32_t i4 = lane; i4 < n4; i4 += 32) {  // pad atoms are zero in frame and reference
      const uint32_t i4n = i4 + 32;
#pragma unroll
      for (int ff = 0; ff < SW_FPW; ++ff)
#pragma unroll
        for (int p = 0; p < 3; ++p) nxt[ff][p] = i4n < n4 ? __ldcs(fb[ff] + p * (a.Npad / 4) + i4n) : zero4;
      const double2 vx0 = __ldg(rx + 2 * i4), vx1 = __ldg(rx + 2 * i4 + 1);
      const double2 vy0 = __ldg(ry + 2 * i4), vy1 = __ldg(ry + 2 * i4 + 1);
      const double2 vz0 = __ldg(rz + 2 * i4), vz1 = __ldg(rz + 2 * i4 + 1);
      const double vxs[4] = {vx0.x, vx0.y, vx1.x, vx1.y}, vys[4] = {vy0.x, vy0.y, vy1.x, vy1.y},
                   vzs[4] = {vz0.x, vz0.y, vz1.x, vz1.y};
#pragma unroll
      for (int ff = 0; ff < SW_FPW; ++ff) {
        if (ff < nf) {
          const float4 X = cur[ff][0], Y = cur[ff][1], Z = cur[ff][2];
          const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const double x = xs[k], y = ys[k], z = zs[k];
            s[ff][0] = fma(x, vxs[k], s[ff][0]); s[ff][1] = fma(x, vys[k], s[ff][1]); s[ff][2] = fma(x, vzs[k], s[ff][2]);
            s[ff][3] = fma(y, vxs[k], s[ff][3]); s[ff][4] = fma(y, vys[k], s[ff][4]); s[ff][5] = fma(y, vzs[k], s[ff][5]);
            s[ff][6] = fma(z, vxs[k], s[ff][6]); s[ff][7] = fma(z, vys[k], s[ff][7]); s[ff][8] = fma(z, vzs[k], s[ff][8]);
          }
        }
      }
#pragma unroll
      for (int ff = 0; ff < SW_FPW; ++ff)
#pragma unroll
        for (int p = 0; p < 3; ++p) cur[ff][p] = nxt[ff][p];
    }
#pragma unroll
    for (int ff = 0; ff < SW_FPW; ++ff) {
#pragma unroll
      for (int k = 0; k < 9; ++k) s[ff][k] = wsum_all(s[ff][k]);
      if (ff < nf && lane < 9) {
        double v = s[ff][0];
#pragma unroll
        for (int k = 1; k < 9; ++k) v = lane == k ? s[ff][k] : v;
        a.cov[(size_t)(f0 + ff) * 9 + lane] = v;
      }
    }
  }
}

// One frame per thread: centroid correction, QCP root + eigenvector, W = T(vc) Z T(-uc)
// (alignment::kabsch, src/alignment.cpp:217-233).  rmsd_out (optional): the minimal RMSD from the
// eigenvalue, for the case where the residual pass would recompute exactly that.
struct SolveArgs {
  const double* cov; const double* cen; const double* gd;  // per frame
  double ref_c[3], ref_sum[3], ref_g;
  // the same seven scalars in device memory ([0..2] centroid, [3..5] sum, [7] sum of squares), when they
  // are produced on the device (iterative alignment: k_finish_iter writes the next target's): no host round trip
  const double* ref_scalars_dev;
  uint32_t F, Na;
  int identity;
  double* xf;        // [F][16] or null
  double* rmsd_out;  // [F] or null
};

__global__ void __launch_bounds__(128) k_solve(const SolveArgs a_in) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= a_in.F) return;
  SolveArgs a = a_in;
  if (a.ref_scalars_dev) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { a.ref_c[k] = __ldg(a.ref_scalars_dev + k); a.ref_sum[k] = __ldg(a.ref_scalars_dev + 3 + k); }
    a.ref_g = __ldg(a.ref_scalars_dev + 7);
  }
  double W[12];
  if (a.identity) {
#pragma unroll
    for (int k = 0; k < 12; ++k) W[k] = (k % 5 == 0) ? 1.0 : 0.0;
  } else {
    const double uc[3] = {a.cen[3 * (size_t)f], a.cen[3 * (size_t)f + 1], a.cen[3 * (size_t)f + 2]};
    double R[9];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = 0; q < 3; ++q) R[3 * p + q] = a.cov[(size_t)f * 9 + 3 * p + q] - uc[p] * a.ref_sum[q];
    // the frame's own centred sum of squares: with E0 = (ga + gb)/2 the quartic's root is <= 1
    const double ga = a.gd[f];
    double Z[9], resid = 0.0;
    if (!qcp_rotation_fast(R, ga, a.ref_g, Z, &resid)) {  // degenerate: Jacobi
      double lam = 0.0;
      qcp_rotation(R, Z, &lam);
      resid = ga + a.ref_g - 2.0 * lam;
    }
    if (a.rmsd_out) a.rmsd_out[f] = sqrt(fabs(resid) / (double)a.Na);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      W[4 * i] = Z[3 * i]; W[4 * i + 1] = Z[3 * i + 1]; W[4 * i + 2] = Z[3 * i + 2];
      W[4 * i + 3] = a.ref_c[i] - (Z[3 * i] * uc[0] + Z[3 * i + 1] * uc[1] + Z[3 * i + 2] * uc[2]);
    }
  }
  if (a.xf) {
    double* o = a.xf + (size_t)f * 16;
#pragma unroll
    for (int k = 0; k < 12; ++k) o[k] = W[k];
    o[12] = o[13] = o[14] = 0.0; o[15] = 1.0;
  }
}

// Residual pass (AtomicGroup::applyTransform + rmsd, AG_numerical.cpp:363-370,301-318): explicit
// residuals, no cancellation.  With both structures shifted to their centroids every term is
// a few tens of Angstrom at most, where float32 resolves 2e-6 A -- the resolution of the float32
// trajectory itself -- so the residuals are formed in fp32 (reference carried as float-float,
// frame constants prepared in fp64) and only their squares are summed in fp64:
//   r - (Z x + t) = (r - rc) - Z (x - uc) - kappa,   kappa = t + Z uc - rc.
// Measured error against the fp64 form < 1e-6 A (tests/test_gpu_rmsd2ref.py).
struct ResidArgs {
  const float* raw; size_t Npad; uint32_t Nr;  // rmsd selection planes
  const float2* ref_ff;  // [3][Npad] reference minus its centroid, float-float (k_ref_split)
  const double* ref_c;   // [3] device: that centroid
  const double* cen;     // [F][3] frame centroids (rmsd selection)
  const double* xf;      // [F][16]
  uint32_t F;
  double* out_rmsd;
};

__global__ void __launch_bounds__(SW_THREADS, 2) k_resid(const ResidArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * SW_THREADS + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * SW_THREADS) >> 5;
  const float4* rx = reinterpret_cast<const float4*>(a.ref_ff);
  const float4* ry = reinterpret_cast<const float4*>(a.ref_ff + a.Npad);
  const float4* rz = reinterpret_cast<const float4*>(a.ref_ff + 2 * a.Npad);
  for (uint32_t f0 = warp_global * SW_FPW; f0 < a.F; f0 += nwarps * SW_FPW) {
    const int nf = (int)(a.F - f0 < (uint32_t)SW_FPW ? a.F - f0 : (uint32_t)SW_FPW);
    float Zf[SW_FPW][9], kap[SW_FPW][3], ucf[SW_FPW][3];
    double d[SW_FPW];
#pragma unroll
    for (int ff = 0; ff < SW_FPW; ++ff) {
      d[ff] = 0.0;
      const size_t f = f0 + (ff < nf ? ff : 0);
      const double* W = a.xf + f * 16;
#pragma unroll
      for (int p = 0; p < 3; ++p) ucf[ff][p] = (float)a.cen[3 * f + p];
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        const double w0 = W[4 * p], w1 = W[4 * p + 1], w2 = W[4 * p + 2], w3 = W[4 * p + 3];
        Zf[ff][3 * p] = (float)w0; Zf[ff][3 * p + 1] = (float)w1; Zf[ff][3 * p + 2] = (float)w2;
        kap[ff][p] = (float)(w3 + w0 * (double)ucf[ff][0] + w1 * (double)ucf[ff][1] + w2 * (double)ucf[ff][2] - a.ref_c[p]);
      }
    }
    // same two-deep pipeline of the frame loads as k_cov
    const uint32_t n4 = (uint32_t)(a.Npad / 4);
    const float4* fb[SW_FPW];
#pragma unroll
    for (int ff = 0; ff < SW_FPW; ++ff)
      fb[ff] = reinterpret_cast<const float4*>(a.raw + (size_t)(f0 + (ff < nf ? ff : 0)) * 3 * a.Npad);
    float4 cur[SW_FPW][3], nxt[SW_FPW][3];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ff = 0; ff < SW_FPW; ++ff)
#pragma unroll
      for (int p = 0; p < 3; ++p) cur[ff][p] = lane < n4 ? __ldcs(fb[ff] + p * (a.Npad / 4) + lane) : zero4;
    for (uint32_t i4 = lane; i4 < n4; i4 += 32) {
      const uint32_t i4n = i4 + 32;
#pragma unroll
      for (int ff = 0; ff < SW_FPW; ++ff)
#pragma unroll
        for (int p = 0; p < 3; ++p) nxt[ff][p] = i4n < n4 ? __ldcs(fb[ff] + p * (a.Npad / 4) + i4n) : zero4;
      const float4 t0 = __ldg(rx + 2 * i4), t1 = __ldg(rx + 2 * i4 + 1);
      const float4 u0 = __ldg(ry + 2 * i4), u1 = __ldg(ry + 2 * i4 + 1);
      const float4 w0 = __ldg(rz + 2 * i4), w1 = __ldg(rz + 2 * i4 + 1);
      const float vxh[4] = {t0.x, t0.z, t1.x, t1.z}, vxl[4] = {t0.y, t0.w, t1.y, t1.w};
      const float vyh[4] = {u0.x, u0.z, u1.x, u1.z}, vyl[4] = {u0.y, u0.w, u1.y, u1.w};
      const float vzh[4] = {w0.x, w0.z, w1.x, w1.z}, vzl[4] = {w0.y, w0.w, w1.y, w1.w};
#pragma unroll
      for (int ff = 0; ff < SW_FPW; ++ff) {
        if (ff < nf) {
          const float4 X = cur[ff][0], Y = cur[ff][1], Z = cur[ff][2];
          const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
          float part = 0.0f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (4 * i4 + k < a.Nr) {  // pad atoms would contribute |kappa|^2
              const float x = xs[k] - ucf[ff][0], y = ys[k] - ucf[ff][1], z = zs[k] - ucf[ff][2];
              const float tx = fmaf(Zf[ff][0], x, fmaf(Zf[ff][1], y, fmaf(Zf[ff][2], z, kap[ff][0])));
              const float ty = fmaf(Zf[ff][3], x, fmaf(Zf[ff][4], y, fmaf(Zf[ff][5], z, kap[ff][1])));
              const float tz = fmaf(Zf[ff][6], x, fmaf(Zf[ff][7], y, fmaf(Zf[ff][8], z, kap[ff][2])));
              const float ex = (vxh[k] - tx) + vxl[k], ey = (vyh[k] - ty) + vyl[k], ez = (vzh[k] - tz) + vzl[k];
              part = fmaf(ex, ex, fmaf(ey, ey, fmaf(ez, ez, part)));
            }
          }
          d[ff] += (double)part;
        }
      }
#pragma unroll
      for (int ff = 0; ff < SW_FPW; ++ff)
#pragma unroll
        for (int p = 0; p < 3; ++p) cur[ff][p] = nxt[ff][p];
    }
#pragma unroll
    for (int ff = 0; ff < SW_FPW; ++ff) {
      const double dd = wsum_all(d[ff]);
      if (ff < nf && lane == 0) a.out_rmsd[f0 + ff] = sqrt(dd / a.Nr);
    }
  }
}

// Centroid of an [3][Npad] fp64 structure and its centred float-float copy (one CTA).
__global__ void k_ref_split(const double* __restrict__ ref, uint32_t n, size_t Npad, float2* __restrict__ out,
                            double* __restrict__ cen) {
  __shared__ double red[3 * 32];
  __shared__ double c_s[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int ax = 0; ax < 3; ++ax) {
    double v = 0.0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) v += ref[ax * Npad + i];
    v = wsum_all(v);
    if (lane == 0) red[ax * 32 + warp] = v;
  }
  __syncthreads();
  if (warp < 3) {
    double v = lane < (int)(blockDim.x >> 5) ? red[warp * 32 + lane] : 0.0;
    v = wsum_all(v);
    if (lane == 0) { c_s[warp] = v / (double)n; cen[warp] = v / (double)n; }
  }
  __syncthreads();
  for (int ax = 0; ax < 3; ++ax)
    for (uint32_t i = threadIdx.x; i < Npad; i += blockDim.x) {
      const double v = i < n ? ref[ax * Npad + i] - c_s[ax] : 0.0;
      const float h = (float)v;
      out[ax * Npad + i] = make_float2(h, (float)(v - (double)h));
    }
}

// avg /= F; rms = sqrt(sum |avg - target|^2 / n) (AtomicGroup::rmsd); then the new
// target := avg, stored centred with its centroid (kabsch centres it anyway).
// target_c holds the uncentred target as [3][Npad] = tgt_centred + centroid.
__global__ void k_finish_iter(double* avg, double* tgt_centred, double* tgt_cen, double* tgt_sum, double* tgt_g,
                              uint32_t n, size_t Npad, double invF, double* rms_out) {
  __shared__ double red[4 * 32];
  double s[4] = {0, 0, 0, 0};
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    for (int ax = 0; ax < 3; ++ax) {
      const double v = avg[ax * Npad + i] * invF;
      avg[ax * Npad + i] = v;
      const double t = tgt_centred[ax * Npad + i] + tgt_cen[ax];
      s[ax] += v;
      s[3] += (v - t) * (v - t);
    }
  // block reduce (blockDim = 1024)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < 4; ++k) {
    double v = s[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k * 32 + warp] = v;
  }
  __syncthreads();
  for (int k = 0; k < 4; ++k) {
    double v = (lane < (int)(blockDim.x >> 5)) ? red[k * 32 + lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    s[k] = v;
  }
  __syncthreads();
  const double c[3] = {s[0] / n, s[1] / n, s[2] / n};
  double cs[4] = {0, 0, 0, 0};
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    for (int ax = 0; ax < 3; ++ax) {
      const double v = avg[ax * Npad + i] - c[ax];
      tgt_centred[ax * Npad + i] = v;
      cs[ax] += v;
      cs[3] = fma(v, v, cs[3]);
    }
  for (int k = 0; k < 4; ++k) {
    double v = cs[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k * 32 + warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 3; ++k) {
      double v = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[k * 32 + w];
      tgt_sum[k] = v;
      tgt_cen[k] = c[k];
    }
    {
      double v = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[3 * 32 + w];
      tgt_g[0] = v;
    }
    rms_out[0] = sqrt(s[3] / n);
  }
}

__global__ void k_scale(double* v, size_t n, double f) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] *= f;
}

// host: interleaved 3N doubles -> [3][Npad] SoA (optionally centred like
// alignment::centerAtOrigin, src/alignment.cpp:90-109)
void to_soa(const double* xyz, size_t n, size_t Npad, bool centre, std::vector<double>& soa,
            double c[3], double csum[3]) {
  soa.assign(3 * Npad, 0.0);
  c[0] = c[1] = c[2] = 0.0;
  if (centre) {
    for (size_t i = 0; i < n; ++i) { c[0] += xyz[3 * i]; c[1] += xyz[3 * i + 1]; c[2] += xyz[3 * i + 2]; }
    for (int k = 0; k < 3; ++k) c[k] = 3 * c[k] / (double)(3 * n);
  }
  csum[0] = csum[1] = csum[2] = 0.0;
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      const double v = xyz[3 * i + k] - c[k];
      soa[k * Npad + i] = v;
      csum[k] += v;
    }
}

// Launch k_accum over all frames and all atoms of a [F][3][Npad] plane set: atom chunks of at most
// ACC_CHUNK atoms (equal sizes), and per chunk as many frame walkers as fill the device.
int launch_accum(AccumArgs aa, cudaStream_t st) {
  const uint32_t n4 = (uint32_t)(aa.Npad / 4);
  const uint32_t nchunks = std::max<uint32_t>(1, (n4 * 4 + ACC_CHUNK - 1) / ACC_CHUNK);
  aa.chunk4 = std::max<uint32_t>(1, (n4 + nchunks - 1) / nchunks);
  const int smem = (int)(3 * 4 * aa.chunk4 * sizeof(double));
  LB_CUDA(cudaFuncSetAttribute(k_accum, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_accum, ST, (size_t)smem);
  if (per_sm < 1) per_sm = 1;
  uint32_t walkers = std::max<uint32_t>(1, (uint32_t)(sms * per_sm) / nchunks);
  if (walkers > aa.F) walkers = aa.F;
  k_accum<<<dim3(walkers, nchunks), ST, smem, st>>>(aa);
  return LOOSB200_OK;
}

template <class K>
int grid_warp(uint32_t F, K kernel) {  // persistent: as many CTAs as fit, each warp walks frames
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SW_THREADS, 0);
  if (per_sm < 1) per_sm = 1;
  const uint32_t per_cta = (SW_THREADS / 32) * SW_FPW;
  const uint32_t need = (F + per_cta - 1) / per_cta;
  const uint32_t cap = (uint32_t)(sms * per_sm);
  return (int)(need < cap ? need : cap);
}

}  // namespace
}  // namespace loosb200

using namespace loosb200;

static int check_pair(const loosb200_frames* align, const loosb200_frames* rs) {
  if (!align) return LOOSB200_ERR_INVALID;
  if (rs && (rs->F != align->F || rs->device != align->device)) return LOOSB200_ERR_SIZE_MISMATCH;
  return LOOSB200_OK;
}

extern "C" int loosb200_rmsd2ref(loosb200_frames* align, loosb200_frames* rmsd_set,
                                 const double* ref_align, const double* ref_rmsd, double* out_rmsd,
                                 double* out_xforms) {
  int rc = check_pair(align, rmsd_set);
  if (rc) return rc;
  if (!out_rmsd) return LOOSB200_ERR_INVALID;
  if (!ref_align && !ref_rmsd) return LOOSB200_ERR_INVALID;
  if (rmsd_set && rmsd_set != align && !ref_rmsd) return LOOSB200_ERR_INVALID;
  loosb200_frames* rs = rmsd_set ? rmsd_set : align;
  if (!ref_rmsd) ref_rmsd = ref_align;
  LB_CUDA(cudaSetDevice(align->device));
  cudaStream_t st = align->stream;
  const size_t F = align->F;

  std::vector<double> soaA, soaR;
  double c[3] = {0, 0, 0}, cs[3] = {0, 0, 0}, c0[3], cs0[3];
  if (ref_align) to_soa(ref_align, align->N, align->Npad, true, soaA, c, cs);
  to_soa(ref_rmsd, rs->N, rs->Npad, false, soaR, c0, cs0);

  // When the rmsd selection and target ARE the align selection and target, the rmsd after
  // superposition is the minimal RMSD sqrt((Ga + Gb - 2 lambda_max) / N): no residual pass.
  // LOOSB200_STREAM_EXPLICIT=1 forces the explicit form (tests compare the two).
  const bool same = ref_align && rs == align && ref_rmsd == ref_align && !getenv("LOOSB200_STREAM_EXPLICIT");
  const bool need_xf = out_xforms || !same;

  double *dA = nullptr, *dR = nullptr, *dOut = nullptr, *dXf = nullptr, *dRc = nullptr, *dCov = nullptr;
  float2* dRff = nullptr;
  bool ok = dev_alloc_t(&dOut, sizeof(double) * F) == cudaSuccess;
  if (ok && !same)
    ok = dev_alloc_t(&dR, sizeof(double) * soaR.size()) == cudaSuccess &&
         dev_alloc_t(&dRff, sizeof(float2) * soaR.size()) == cudaSuccess &&
         dev_alloc_t(&dRc, sizeof(double) * 3) == cudaSuccess;
  if (ok && ref_align)
    ok = dev_alloc_t(&dA, sizeof(double) * soaA.size()) == cudaSuccess &&
         dev_alloc_t(&dCov, sizeof(double) * 9 * F) == cudaSuccess;
  if (ok && need_xf) ok = dev_alloc_t(&dXf, sizeof(double) * 16 * F) == cudaSuccess;
  auto cleanup = [&]() {
    dev_release(dA); dev_release(dR); dev_release(dRff); dev_release(dRc); dev_release(dOut); dev_release(dXf); dev_release(dCov);
  };
  if (!ok) { cudaGetLastError(); cleanup(); return LOOSB200_ERR_NOMEM; }
  cudaEventRecord(align->timing.ev[0], st);
  if (ref_align) cudaMemcpyAsync(dA, soaA.data(), sizeof(double) * soaA.size(), cudaMemcpyHostToDevice, st);
  if (!same) {
    cudaMemcpyAsync(dR, soaR.data(), sizeof(double) * soaR.size(), cudaMemcpyHostToDevice, st);
    k_ref_split<<<1, 1024, 0, st>>>(dR, (uint32_t)rs->N, rs->Npad, dRff, dRc);
  }
  cudaEventRecord(align->timing.ev[1], st);
  int launches = 0;
  if (ref_align) {
    CovArgs ca;
    ca.raw = align->raw; ca.Npad = align->Npad; ca.ref = dA; ca.F = (uint32_t)F; ca.cov = dCov;
    k_cov<<<grid_warp((uint32_t)F, k_cov), SW_THREADS, 0, st>>>(ca);
    ++launches;
  }
  SolveArgs sa;
  memset(&sa, 0, sizeof(sa));
  sa.cov = dCov; sa.cen = align->centroid; sa.gd = align->gd;
  for (int k = 0; k < 3; ++k) { sa.ref_c[k] = c[k]; sa.ref_sum[k] = cs[k]; }
  for (double v : soaA) sa.ref_g += v * v;
  sa.F = (uint32_t)F; sa.Na = (uint32_t)align->N;
  sa.identity = ref_align ? 0 : 1;
  sa.xf = dXf; sa.rmsd_out = same ? dOut : nullptr;
  k_solve<<<(unsigned)((F + 127) / 128), 128, 0, st>>>(sa);
  ++launches;
  if (!same) {
    ResidArgs ra;
    ra.raw = rs->raw; ra.Npad = rs->Npad; ra.Nr = (uint32_t)rs->N;
    ra.ref_ff = dRff; ra.ref_c = dRc; ra.cen = rs->centroid; ra.xf = dXf; ra.F = (uint32_t)F; ra.out_rmsd = dOut;
    k_resid<<<grid_warp((uint32_t)F, k_resid), SW_THREADS, 0, st>>>(ra);
    ++launches;
  }
  cudaEventRecord(align->timing.ev[2], st);
  cudaMemcpyAsync(out_rmsd, dOut, sizeof(double) * F, cudaMemcpyDeviceToHost, st);
  if (out_xforms) cudaMemcpyAsync(out_xforms, dXf, sizeof(double) * 16 * F, cudaMemcpyDeviceToHost, st);
  cudaEventRecord(align->timing.ev[3], st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  cleanup();
  if (e != cudaSuccess) return cuda_fail(e, "rmsd2ref kernels", __FILE__, __LINE__);
  align->timing.launches = launches; align->timing.valid = true; align->timing.kev_used = 0;
  return LOOSB200_OK;
}

extern "C" int loosb200_align_accumulate(loosb200_frames* align, loosb200_frames* accum_set, const double* target,
                                         double* sum_out, double* out_xforms) {
  int rc = check_pair(align, accum_set);
  if (rc) return rc;
  if (!target || !sum_out) return LOOSB200_ERR_INVALID;
  loosb200_frames* as = accum_set ? accum_set : align;
  LB_CUDA(cudaSetDevice(align->device));
  cudaStream_t st = align->stream;
  const size_t F = align->F, NpS = as->Npad;

  std::vector<double> soaA;
  double c[3], cs[3];
  to_soa(target, align->N, align->Npad, true, soaA, c, cs);
  double *dA = nullptr, *dCov = nullptr, *dXf = nullptr, *dSum = nullptr;
  auto cleanup = [&]() { dev_release(dA); dev_release(dCov); dev_release(dXf); dev_release(dSum); };
  const bool ok = dev_alloc_t(&dA, sizeof(double) * soaA.size()) == cudaSuccess &&
                  dev_alloc_t(&dCov, sizeof(double) * 9 * F) == cudaSuccess &&
                  dev_alloc_t(&dXf, sizeof(double) * 16 * F) == cudaSuccess &&
                  dev_alloc_t(&dSum, sizeof(double) * 3 * NpS) == cudaSuccess;
  if (!ok) { cudaGetLastError(); cleanup(); return LOOSB200_ERR_NOMEM; }
  cudaEventRecord(align->timing.ev[0], st);
  cudaMemcpyAsync(dA, soaA.data(), sizeof(double) * soaA.size(), cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(dSum, 0, sizeof(double) * 3 * NpS, st);
  cudaEventRecord(align->timing.ev[1], st);
  CovArgs ca;
  ca.raw = align->raw; ca.Npad = align->Npad; ca.ref = dA; ca.F = (uint32_t)F; ca.cov = dCov;
  k_cov<<<grid_warp((uint32_t)F, k_cov), SW_THREADS, 0, st>>>(ca);
  SolveArgs sa;
  memset(&sa, 0, sizeof(sa));
  sa.cov = dCov; sa.cen = align->centroid; sa.gd = align->gd;
  for (int k = 0; k < 3; ++k) { sa.ref_c[k] = c[k]; sa.ref_sum[k] = cs[k]; }
  for (double v : soaA) sa.ref_g += v * v;
  sa.F = (uint32_t)F; sa.Na = (uint32_t)align->N; sa.xf = dXf;
  k_solve<<<(unsigned)((F + 127) / 128), 128, 0, st>>>(sa);
  AccumArgs aa;
  aa.raw = as->raw; aa.Npad = NpS; aa.N = (uint32_t)as->N; aa.xf = dXf; aa.F = (uint32_t)F; aa.avg = dSum;
  launch_accum(aa, st);
  cudaEventRecord(align->timing.ev[2], st);
  std::vector<double> sum_h(3 * NpS);
  cudaMemcpyAsync(sum_h.data(), dSum, sizeof(double) * 3 * NpS, cudaMemcpyDeviceToHost, st);
  if (out_xforms) cudaMemcpyAsync(out_xforms, dXf, sizeof(double) * 16 * F, cudaMemcpyDeviceToHost, st);
  cudaEventRecord(align->timing.ev[3], st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  cleanup();
  if (e != cudaSuccess) return cuda_fail(e, "align_accumulate", __FILE__, __LINE__);
  for (size_t i = 0; i < as->N; ++i)
    for (int ax = 0; ax < 3; ++ax) sum_out[3 * i + ax] = sum_h[ax * NpS + i];
  align->timing.launches = 3; align->timing.valid = true; align->timing.kev_used = 0;
  return LOOSB200_OK;
}

extern "C" int loosb200_rmsd2ref_iterative(loosb200_frames* align, loosb200_frames* rmsd_set,
                                           double tol, int maxiter, double* out_rmsd,
                                           double* out_xforms, double* out_avg, int* iters,
                                           double* final_rms) {
  int rc = check_pair(align, rmsd_set);
  if (rc) return rc;
  if (!out_rmsd) return LOOSB200_ERR_INVALID;
  loosb200_frames* rs = rmsd_set ? rmsd_set : align;
  LB_CUDA(cudaSetDevice(align->device));
  cudaStream_t st = align->stream;
  const size_t F = align->F, Na = align->N, NpA = align->Npad, Nr = rs->N, NpR = rs->Npad;

  double *dT = nullptr, *dAvg = nullptr, *dAvgR = nullptr, *dXf = nullptr, *dOut = nullptr, *dScal = nullptr, *dRc = nullptr, *dCov = nullptr;
  float2* dRff = nullptr;
  bool ok = dev_alloc_t(&dT, sizeof(double) * 3 * NpA) == cudaSuccess &&
            dev_alloc_t(&dAvg, sizeof(double) * 3 * NpA) == cudaSuccess &&
            dev_alloc_t(&dAvgR, sizeof(double) * 3 * NpR) == cudaSuccess &&
            dev_alloc_t(&dRff, sizeof(float2) * 3 * NpR) == cudaSuccess &&
            dev_alloc_t(&dRc, sizeof(double) * 3) == cudaSuccess &&
            dev_alloc_t(&dXf, sizeof(double) * 16 * F) == cudaSuccess &&
            dev_alloc_t(&dOut, sizeof(double) * F) == cudaSuccess &&
            dev_alloc_t(&dCov, sizeof(double) * 9 * F) == cudaSuccess &&
            dev_alloc_t(&dScal, sizeof(double) * 8) == cudaSuccess;
  auto cleanup = [&]() {
    dev_release(dT); dev_release(dAvg); dev_release(dAvgR); dev_release(dRff); dev_release(dRc); dev_release(dXf); dev_release(dOut); dev_release(dScal); dev_release(dCov);
  };
  if (!ok) { cudaGetLastError(); cleanup(); return LOOSB200_ERR_NOMEM; }
  double* tgt_cen = dScal; double* tgt_sum = dScal + 3; double* rms_dev = dScal + 6;

  // target_0 = first frame, centred (alignment.cpp:372-373): fetch its planes
  std::vector<float> f0(3 * NpA);
  {
    const cudaError_t ce = cudaMemcpy(f0.data(), align->raw, sizeof(float) * 3 * NpA, cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) { cleanup(); return cuda_fail(ce, "fetching the first frame", __FILE__, __LINE__); }
  }
  std::vector<double> t0(3 * NpA, 0.0);
  double c[3] = {0, 0, 0}, cs[3] = {0, 0, 0};
  for (int ax = 0; ax < 3; ++ax) {
    for (size_t i = 0; i < Na; ++i) c[ax] += (double)f0[ax * NpA + i];
    c[ax] /= (double)Na;
    for (size_t i = 0; i < Na; ++i) { t0[ax * NpA + i] = (double)f0[ax * NpA + i] - c[ax]; cs[ax] += t0[ax * NpA + i]; }
  }
  // the centred first frame IS the target (centroid 0)
  double g0 = 0.0;
  for (double v : t0) g0 += v * v;
  double scal[8] = {0, 0, 0, cs[0], cs[1], cs[2], 0, g0};
  cudaEventRecord(align->timing.ev[0], st);
  cudaMemcpyAsync(dT, t0.data(), sizeof(double) * 3 * NpA, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dScal, scal, sizeof(scal), cudaMemcpyHostToDevice, st);
  cudaEventRecord(align->timing.ev[1], st);

  int iter = 0, launches = 0;
  double rms = 0.0;
  cudaError_t e = cudaSuccess;
  do {
    // one iteration = covariance of every frame with the current target (k_cov), one solve per
    // thread (k_solve -> transforms), transformed frames summed into the new average (k_accum)
    cudaMemsetAsync(dAvg, 0, sizeof(double) * 3 * NpA, st);
    CovArgs ca;
    ca.raw = align->raw; ca.Npad = NpA; ca.ref = dT; ca.F = (uint32_t)F; ca.cov = dCov;
    k_cov<<<grid_warp((uint32_t)F, k_cov), SW_THREADS, 0, st>>>(ca);
    SolveArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.cov = dCov; sa.cen = align->centroid; sa.gd = align->gd;
    sa.ref_scalars_dev = dScal;  // the target's centroid, sum and G as k_finish_iter left them on the device
    sa.F = (uint32_t)F; sa.Na = (uint32_t)Na; sa.xf = dXf;
    k_solve<<<(unsigned)((F + 127) / 128), 128, 0, st>>>(sa);
    AccumArgs aa;
    aa.raw = align->raw; aa.Npad = NpA; aa.N = (uint32_t)Na; aa.xf = dXf; aa.F = (uint32_t)F; aa.avg = dAvg;
    launch_accum(aa, st);
    k_finish_iter<<<1, 1024, 0, st>>>(dAvg, dT, tgt_cen, tgt_sum, dScal + 7, (uint32_t)Na, NpA, 1.0 / (double)F, rms_dev);
    launches += 4;
    cudaMemcpyAsync(&rms, rms_dev, sizeof(double), cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
    ++iter;
  } while (rms > tol && iter <= maxiter);  // alignment.cpp:399
  if (e == cudaSuccess) {
    // averageStructure(subset, xforms, traj, indices): ensembles.cpp:91-121
    cudaMemsetAsync(dAvgR, 0, sizeof(double) * 3 * NpR, st);
    AccumArgs ar;
    ar.raw = rs->raw; ar.Npad = NpR; ar.N = (uint32_t)Nr; ar.xf = dXf; ar.F = (uint32_t)F; ar.avg = dAvgR;
    launch_accum(ar, st);
    k_scale<<<(unsigned)((3 * NpR + 255) / 256), 256, 0, st>>>(dAvgR, 3 * NpR, 1.0 / (double)F);
    // per-frame rmsd to the average: rmsd2ref.cpp:238-245
    k_ref_split<<<1, 1024, 0, st>>>(dAvgR, (uint32_t)Nr, NpR, dRff, dRc);
    ResidArgs ra;
    ra.raw = rs->raw; ra.Npad = NpR; ra.Nr = (uint32_t)Nr;
    ra.ref_ff = dRff; ra.ref_c = dRc; ra.cen = rs->centroid; ra.xf = dXf; ra.F = (uint32_t)F; ra.out_rmsd = dOut;
    k_resid<<<grid_warp((uint32_t)F, k_resid), SW_THREADS, 0, st>>>(ra);
    launches += 4;
    cudaEventRecord(align->timing.ev[2], st);
    cudaMemcpyAsync(out_rmsd, dOut, sizeof(double) * F, cudaMemcpyDeviceToHost, st);
    if (out_xforms) cudaMemcpyAsync(out_xforms, dXf, sizeof(double) * 16 * F, cudaMemcpyDeviceToHost, st);
    std::vector<double> avg_h;
    if (out_avg) {
      avg_h.resize(3 * NpR);
      cudaMemcpyAsync(avg_h.data(), dAvgR, sizeof(double) * 3 * NpR, cudaMemcpyDeviceToHost, st);
    }
    cudaEventRecord(align->timing.ev[3], st);
    e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess && out_avg)
      for (size_t i = 0; i < Nr; ++i)
        for (int ax = 0; ax < 3; ++ax) out_avg[3 * i + ax] = avg_h[ax * NpR + i];
  }
  cleanup();
  if (e != cudaSuccess) return cuda_fail(e, "rmsd2ref_iterative", __FILE__, __LINE__);
  if (iters) *iters = iter;
  if (final_rms) *final_rms = rms;
  align->timing.launches = launches; align->timing.valid = true; align->timing.kev_used = 0;
  return LOOSB200_OK;
}
