// Single-reference streaming path (rmsd2ref and iterative alignment).
//
// One CTA walks frames; per frame it makes at most two passes over the frame's
// fp32 planes (the second one hits L1/L2):
//   pass 1  centroid + 3x3 covariance against the centred reference
//           (alignment::kabsch, src/alignment.cpp:217-233: centre both, dgemm)
//           -> Jacobi eigen-solve of Horn's K -> rotation Z, 4x4 W = T(vc) Z T(-uc)
//   pass 2  x' = W x over the rmsd selection (AtomicGroup::applyTransform,
//           src/AG_numerical.cpp:363-370) and either sum |x' - y|^2
//           (AtomicGroup::rmsd, :301-318) or accumulation of x' into the running
//           average (iterativeAlignment / averageStructure,
//           src/alignment.cpp:379-393, src/ensembles.cpp:108-118).
// HBM-bound: 12 N_align (+ 12 N_rmsd if a different selection) bytes per frame.
#include <string.h>

#include <vector>

#include "common.cuh"
#include "qcp.cuh"

namespace loosb200 {

namespace {

constexpr int ST = 128;  // threads per CTA

enum : int { M_COMPUTE_W = 1, M_ACCUM = 2, M_RMSD = 4, M_IDENTITY = 8 };

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>
__device__ __forceinline__ void bsum(double (&v)[NV], double* red /* NV*4 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = wsum(v[i]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * 4 + warp] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = red[i * 4] + red[i * 4 + 1] + red[i * 4 + 2] + red[i * 4 + 3];
}

struct StreamArgs {
  const float* rawA; size_t NpadA; uint32_t Na;   // align selection planes
  const float* rawR; size_t NpadR; uint32_t Nr;   // rmsd selection planes
  const double* refA;   // [3][NpadA] centred reference (align selection)
  double refA_c[3];     // its centroid before centring
  double refA_sum[3];   // sum of the centred reference (~0; kept for exactness)
  double refA_g;        // sum of squares of the centred reference
  const double* refR;   // [3][NpadR] uncentred reference (rmsd selection)
  double* xf;           // [F][16] in/out
  double* out_rmsd;     // [F]
  double* avg;          // [3][Npad of the accumulated set] global accumulators
  uint32_t F;
  int mode;
  int accum_align_set;  // M_ACCUM over the align planes (iterative) or the rmsd planes
};

__global__ void __launch_bounds__(ST) k_stream(const StreamArgs a) {
  extern __shared__ double acc_s[];  // [3][Npad] when M_ACCUM
  __shared__ double red[12 * 4];
  __shared__ double Ws[12];
  const int tid = threadIdx.x;
  const bool accum = a.mode & M_ACCUM;
  const float* rawS = a.accum_align_set ? a.rawA : a.rawR;
  const size_t NpadS = a.accum_align_set ? a.NpadA : a.NpadR;
  const uint32_t Ns = a.accum_align_set ? a.Na : a.Nr;
  if (accum)
    for (size_t i = tid; i < 3 * NpadS; i += ST) acc_s[i] = 0.0;

  for (uint32_t f = blockIdx.x; f < a.F; f += gridDim.x) {
    if (a.mode & M_COMPUTE_W) {
      const float4* px = reinterpret_cast<const float4*>(a.rawA + ((size_t)f * 3 + 0) * a.NpadA);
      const float4* py = reinterpret_cast<const float4*>(a.rawA + ((size_t)f * 3 + 1) * a.NpadA);
      const float4* pz = reinterpret_cast<const float4*>(a.rawA + ((size_t)f * 3 + 2) * a.NpadA);
      const double* rx = a.refA, *ry = a.refA + a.NpadA, *rz = a.refA + 2 * a.NpadA;
      double s[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) s[k] = 0.0;
      for (uint32_t i4 = tid; i4 < a.NpadA / 4; i4 += ST) {
        const float4 X = __ldg(px + i4), Y = __ldg(py + i4), Z = __ldg(pz + i4);
        const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double x = xs[k], y = ys[k], z = zs[k];
          const double vx = rx[4 * i4 + k], vy = ry[4 * i4 + k], vz = rz[4 * i4 + k];
          s[0] += x; s[1] += y; s[2] += z;
          s[3] = fma(x, vx, s[3]); s[4] = fma(x, vy, s[4]); s[5] = fma(x, vz, s[5]);
          s[6] = fma(y, vx, s[6]); s[7] = fma(y, vy, s[7]); s[8] = fma(y, vz, s[8]);
          s[9] = fma(z, vx, s[9]); s[10] = fma(z, vy, s[10]); s[11] = fma(z, vz, s[11]);
        }
      }
      bsum<12>(s, red);
      if (tid == 0) {
        const double uc[3] = {s[0] / a.Na, s[1] / a.Na, s[2] / a.Na};
        double R[9];
        // sum (x - uc) v^T = sum x v^T - uc (sum v)^T
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int q = 0; q < 3; ++q) R[3 * p + q] = s[3 + 3 * p + q] - uc[p] * a.refA_sum[q];
        double Z[9];
        qcp_rotation(R, Z, 0);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          Ws[4 * i] = Z[3 * i]; Ws[4 * i + 1] = Z[3 * i + 1]; Ws[4 * i + 2] = Z[3 * i + 2];
          Ws[4 * i + 3] = a.refA_c[i] - (Z[3 * i] * uc[0] + Z[3 * i + 1] * uc[1] + Z[3 * i + 2] * uc[2]);
        }
        if (a.xf) {
          double* o = a.xf + (size_t)f * 16;
#pragma unroll
          for (int k = 0; k < 12; ++k) o[k] = Ws[k];
          o[12] = o[13] = o[14] = 0.0; o[15] = 1.0;
        }
      }
      __syncthreads();
    } else {
      if (tid < 12) {
        if (a.mode & M_IDENTITY) Ws[tid] = (tid % 5 == 0) ? 1.0 : 0.0;
        else Ws[tid] = a.xf[(size_t)f * 16 + tid];
      }
      if ((a.mode & M_IDENTITY) && a.xf && tid < 16) a.xf[(size_t)f * 16 + tid] = (tid % 5 == 0) ? 1.0 : 0.0;
      __syncthreads();
    }
    double W[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) W[k] = Ws[k];

    if (accum) {
      const float* px = rawS + ((size_t)f * 3 + 0) * NpadS;
      const float* py = px + NpadS, *pz = py + NpadS;
      for (uint32_t i = tid; i < Ns; i += ST) {  // each thread owns its atoms: no atomics
        const double x = px[i], y = py[i], z = pz[i];
        acc_s[i] += W[0] * x + W[1] * y + W[2] * z + W[3];
        acc_s[NpadS + i] += W[4] * x + W[5] * y + W[6] * z + W[7];
        acc_s[2 * NpadS + i] += W[8] * x + W[9] * y + W[10] * z + W[11];
      }
    }
    if (a.mode & M_RMSD) {
      const float* px = a.rawR + ((size_t)f * 3 + 0) * a.NpadR;
      const float* py = px + a.NpadR, *pz = py + a.NpadR;
      const double* rx = a.refR, *ry = rx + a.NpadR, *rz = ry + a.NpadR;
      double d[1] = {0.0};
      for (uint32_t i = tid; i < a.Nr; i += ST) {
        const double x = px[i], y = py[i], z = pz[i];
        const double ex = rx[i] - (W[0] * x + W[1] * y + W[2] * z + W[3]);
        const double ey = ry[i] - (W[4] * x + W[5] * y + W[6] * z + W[7]);
        const double ez = rz[i] - (W[8] * x + W[9] * y + W[10] * z + W[11]);
        d[0] += ex * ex + ey * ey + ez * ez;
      }
      bsum<1>(d, red);
      if (tid == 0) a.out_rmsd[f] = sqrt(d[0] / a.Nr);
    }
    __syncthreads();
  }
  if (accum)
    for (size_t i = tid; i < 3 * NpadS; i += ST)
      if (acc_s[i] != 0.0) atomicAdd(&a.avg[i], acc_s[i]);
}

// Warp-per-frame variant for the modes without accumulation (rmsd2ref with a target,
// final per-frame RMSD of the iterative mode): one warp owns a frame, so the warps of
// an SM keep hundreds of KB of loads in flight, nothing is block-synchronised, and the
// 3x3 solve (quartic + adjugate eigenvector, redundantly in every lane) is ~300 fp64
// instructions instead of a Jacobi sweep in local memory.  This is the HBM-bound
// kernel of the path: 12 N_align (+ 12 N_rmsd) bytes per frame.
constexpr int SW_THREADS = 256;

__device__ __forceinline__ double wsum_all(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(SW_THREADS) k_stream_warp(const StreamArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * SW_THREADS + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * SW_THREADS) >> 5;
  for (uint32_t f = warp_global; f < a.F; f += nwarps) {
    double W[12];
    if (a.mode & M_COMPUTE_W) {
      const float4* px = reinterpret_cast<const float4*>(a.rawA + ((size_t)f * 3 + 0) * a.NpadA);
      const float4* py = reinterpret_cast<const float4*>(a.rawA + ((size_t)f * 3 + 1) * a.NpadA);
      const float4* pz = reinterpret_cast<const float4*>(a.rawA + ((size_t)f * 3 + 2) * a.NpadA);
      const double2* rx = reinterpret_cast<const double2*>(a.refA);
      const double2* ry = reinterpret_cast<const double2*>(a.refA + a.NpadA);
      const double2* rz = reinterpret_cast<const double2*>(a.refA + 2 * a.NpadA);
      double s[13];
#pragma unroll
      for (int k = 0; k < 13; ++k) s[k] = 0.0;
      for (uint32_t i4 = lane; i4 < a.NpadA / 4; i4 += 32) {
        const float4 X = __ldcs(px + i4), Y = __ldcs(py + i4), Z = __ldcs(pz + i4);  // streamed once
        const double2 vx0 = __ldg(rx + 2 * i4), vx1 = __ldg(rx + 2 * i4 + 1);
        const double2 vy0 = __ldg(ry + 2 * i4), vy1 = __ldg(ry + 2 * i4 + 1);
        const double2 vz0 = __ldg(rz + 2 * i4), vz1 = __ldg(rz + 2 * i4 + 1);
        const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
        const double vxs[4] = {vx0.x, vx0.y, vx1.x, vx1.y}, vys[4] = {vy0.x, vy0.y, vy1.x, vy1.y},
                     vzs[4] = {vz0.x, vz0.y, vz1.x, vz1.y};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double x = xs[k], y = ys[k], z = zs[k];
          s[0] += x; s[1] += y; s[2] += z;
          s[3] = fma(x, vxs[k], s[3]); s[4] = fma(x, vys[k], s[4]); s[5] = fma(x, vzs[k], s[5]);
          s[6] = fma(y, vxs[k], s[6]); s[7] = fma(y, vys[k], s[7]); s[8] = fma(y, vzs[k], s[8]);
          s[9] = fma(z, vxs[k], s[9]); s[10] = fma(z, vys[k], s[10]); s[11] = fma(z, vzs[k], s[11]);
          s[12] = fma(x, x, fma(y, y, fma(z, z, s[12])));
        }
      }
#pragma unroll
      for (int k = 0; k < 13; ++k) s[k] = wsum_all(s[k]);
      const double uc[3] = {s[0] / a.Na, s[1] / a.Na, s[2] / a.Na};
      double R[9];
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) R[3 * p + q] = s[3 + 3 * p + q] - uc[p] * a.refA_sum[q];
      // the frame's own centred sum of squares: with E0 = (ga + gb)/2 the quartic's root is <= 1
      const double ga = s[12] - (double)a.Na * (uc[0] * uc[0] + uc[1] * uc[1] + uc[2] * uc[2]);
      double Z[9];
      if (!qcp_rotation_fast(R, ga, a.refA_g, Z)) qcp_rotation(R, Z, 0);  // degenerate: Jacobi
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        W[4 * i] = Z[3 * i]; W[4 * i + 1] = Z[3 * i + 1]; W[4 * i + 2] = Z[3 * i + 2];
        W[4 * i + 3] = a.refA_c[i] - (Z[3 * i] * uc[0] + Z[3 * i + 1] * uc[1] + Z[3 * i + 2] * uc[2]);
      }
      if (a.xf && lane == 0) {
        double* o = a.xf + (size_t)f * 16;
#pragma unroll
        for (int k = 0; k < 12; ++k) o[k] = W[k];
        o[12] = o[13] = o[14] = 0.0; o[15] = 1.0;
      }
    } else if (a.mode & M_IDENTITY) {
#pragma unroll
      for (int k = 0; k < 12; ++k) W[k] = (k % 5 == 0) ? 1.0 : 0.0;
      if (a.xf && lane < 16) a.xf[(size_t)f * 16 + lane] = (lane % 5 == 0) ? 1.0 : 0.0;
    } else {
#pragma unroll
      for (int k = 0; k < 12; ++k) W[k] = a.xf[(size_t)f * 16 + k];
    }
    if (a.mode & M_RMSD) {
      const float* px = a.rawR + ((size_t)f * 3 + 0) * a.NpadR;
      const float* py = px + a.NpadR, *pz = py + a.NpadR;
      const double* rx = a.refR, *ry = rx + a.NpadR, *rz = ry + a.NpadR;
      double d = 0.0;
      for (uint32_t i = lane; i < a.Nr; i += 32) {  // second pass: L1/L2 hits when the selections coincide
        const double x = px[i], y = py[i], z = pz[i];
        const double ex = rx[i] - (W[0] * x + W[1] * y + W[2] * z + W[3]);
        const double ey = ry[i] - (W[4] * x + W[5] * y + W[6] * z + W[7]);
        const double ez = rz[i] - (W[8] * x + W[9] * y + W[10] * z + W[11]);
        d += ex * ex + ey * ey + ez * ez;
      }
      d = wsum_all(d);
      if (lane == 0) a.out_rmsd[f] = sqrt(d / a.Nr);
    }
  }
}

// avg /= F; rms = sqrt(sum |avg - target|^2 / n) (AtomicGroup::rmsd); then the new
// target := avg, stored centred with its centroid (kabsch centres it anyway).
// target_c holds the uncentred target as [3][Npad] = tgt_centred + centroid.
__global__ void k_finish_iter(double* avg, double* tgt_centred, double* tgt_cen, double* tgt_sum,
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
  double cs[3] = {0, 0, 0};
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    for (int ax = 0; ax < 3; ++ax) {
      const double v = avg[ax * Npad + i] - c[ax];
      tgt_centred[ax * Npad + i] = v;
      cs[ax] += v;
    }
  for (int k = 0; k < 3; ++k) {
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

int grid_for(uint32_t F, bool accum) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int g = accum ? sms * 2 : sms * 16;
  return (int)((F < (uint32_t)g) ? F : (uint32_t)g);
}

int grid_warp(uint32_t F) {  // persistent: as many CTAs as fit, each warp walks frames
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stream_warp, SW_THREADS, 0);
  if (per_sm < 1) per_sm = 1;
  const uint32_t need = (F + SW_THREADS / 32 - 1) / (SW_THREADS / 32);
  const uint32_t cap = (uint32_t)(sms * per_sm);
  return (int)(need < cap ? need : cap);
}

int smem_for_accum(size_t Npad) { return (int)(3 * Npad * sizeof(double)); }

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

  double *dA = nullptr, *dR = nullptr, *dOut = nullptr, *dXf = nullptr;
  bool ok = cudaMalloc(&dR, sizeof(double) * soaR.size()) == cudaSuccess &&
            cudaMalloc(&dOut, sizeof(double) * F) == cudaSuccess;
  if (ok && ref_align) ok = cudaMalloc(&dA, sizeof(double) * soaA.size()) == cudaSuccess;
  if (ok && out_xforms) ok = cudaMalloc(&dXf, sizeof(double) * 16 * F) == cudaSuccess;
  if (!ok) { cudaGetLastError(); cudaFree(dA); cudaFree(dR); cudaFree(dOut); cudaFree(dXf); return LOOSB200_ERR_NOMEM; }
  cudaEventRecord(align->timing.ev[0], st);
  if (ref_align) cudaMemcpyAsync(dA, soaA.data(), sizeof(double) * soaA.size(), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dR, soaR.data(), sizeof(double) * soaR.size(), cudaMemcpyHostToDevice, st);

  StreamArgs a;
  memset(&a, 0, sizeof(a));
  a.rawA = align->raw; a.NpadA = align->Npad; a.Na = (uint32_t)align->N;
  a.rawR = rs->raw; a.NpadR = rs->Npad; a.Nr = (uint32_t)rs->N;
  a.refA = dA; a.refR = dR;
  for (int k = 0; k < 3; ++k) { a.refA_c[k] = c[k]; a.refA_sum[k] = cs[k]; }
  a.refA_g = 0.0;
  for (double v : soaA) a.refA_g += v * v;
  a.xf = dXf; a.out_rmsd = dOut; a.F = (uint32_t)F;
  a.mode = (ref_align ? M_COMPUTE_W : M_IDENTITY) | M_RMSD;
  cudaEventRecord(align->timing.ev[1], st);
  k_stream_warp<<<grid_warp((uint32_t)F), SW_THREADS, 0, st>>>(a);
  cudaEventRecord(align->timing.ev[2], st);
  cudaMemcpyAsync(out_rmsd, dOut, sizeof(double) * F, cudaMemcpyDeviceToHost, st);
  if (out_xforms) cudaMemcpyAsync(out_xforms, dXf, sizeof(double) * 16 * F, cudaMemcpyDeviceToHost, st);
  cudaEventRecord(align->timing.ev[3], st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(dA); cudaFree(dR); cudaFree(dOut); cudaFree(dXf);
  if (e != cudaSuccess) return cuda_fail(e, "k_stream", __FILE__, __LINE__);
  align->timing.launches = 1; align->timing.valid = true; align->timing.kev_used = 0;
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
  const int smemA = smem_for_accum(NpA), smemR = smem_for_accum(NpR);
  if (smemA > 200 * 1024 || smemR > 200 * 1024) return LOOSB200_ERR_UNSUPPORTED;
  LB_CUDA(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               smemA > smemR ? smemA : smemR));

  double *dT = nullptr, *dAvg = nullptr, *dAvgR = nullptr, *dXf = nullptr, *dOut = nullptr, *dScal = nullptr;
  bool ok = cudaMalloc(&dT, sizeof(double) * 3 * NpA) == cudaSuccess &&
            cudaMalloc(&dAvg, sizeof(double) * 3 * NpA) == cudaSuccess &&
            cudaMalloc(&dAvgR, sizeof(double) * 3 * NpR) == cudaSuccess &&
            cudaMalloc(&dXf, sizeof(double) * 16 * F) == cudaSuccess &&
            cudaMalloc(&dOut, sizeof(double) * F) == cudaSuccess &&
            cudaMalloc(&dScal, sizeof(double) * 8) == cudaSuccess;
  auto cleanup = [&]() { cudaFree(dT); cudaFree(dAvg); cudaFree(dAvgR); cudaFree(dXf); cudaFree(dOut); cudaFree(dScal); };
  if (!ok) { cudaGetLastError(); cleanup(); return LOOSB200_ERR_NOMEM; }
  double* tgt_cen = dScal; double* tgt_sum = dScal + 3; double* rms_dev = dScal + 6;

  // target_0 = first frame, centred (alignment.cpp:372-373): fetch its planes
  std::vector<float> f0(3 * NpA);
  LB_CUDA(cudaMemcpy(f0.data(), align->raw, sizeof(float) * 3 * NpA, cudaMemcpyDeviceToHost));
  std::vector<double> t0(3 * NpA, 0.0);
  double c[3] = {0, 0, 0}, cs[3] = {0, 0, 0};
  for (int ax = 0; ax < 3; ++ax) {
    for (size_t i = 0; i < Na; ++i) c[ax] += (double)f0[ax * NpA + i];
    c[ax] /= (double)Na;
    for (size_t i = 0; i < Na; ++i) { t0[ax * NpA + i] = (double)f0[ax * NpA + i] - c[ax]; cs[ax] += t0[ax * NpA + i]; }
  }
  // the centred first frame IS the target (centroid 0)
  double scal[8] = {0, 0, 0, cs[0], cs[1], cs[2], 0, 0};
  cudaEventRecord(align->timing.ev[0], st);
  cudaMemcpyAsync(dT, t0.data(), sizeof(double) * 3 * NpA, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dScal, scal, sizeof(scal), cudaMemcpyHostToDevice, st);
  cudaEventRecord(align->timing.ev[1], st);

  StreamArgs a;
  memset(&a, 0, sizeof(a));
  a.rawA = align->raw; a.NpadA = NpA; a.Na = (uint32_t)Na;
  a.rawR = rs->raw; a.NpadR = NpR; a.Nr = (uint32_t)Nr;
  a.refA = dT; a.xf = dXf; a.out_rmsd = dOut; a.F = (uint32_t)F;
  int iter = 0, launches = 0;
  double rms = 0.0;
  cudaError_t e = cudaSuccess;
  do {
    cudaMemsetAsync(dAvg, 0, sizeof(double) * 3 * NpA, st);
    cudaMemcpyAsync(scal, dScal, sizeof(scal), cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
    for (int k = 0; k < 3; ++k) { a.refA_c[k] = scal[k]; a.refA_sum[k] = scal[3 + k]; }
    a.mode = M_COMPUTE_W | M_ACCUM; a.accum_align_set = 1; a.avg = dAvg;
    k_stream<<<grid_for((uint32_t)F, true), ST, smemA, st>>>(a);
    k_finish_iter<<<1, 1024, 0, st>>>(dAvg, dT, tgt_cen, tgt_sum, (uint32_t)Na, NpA, 1.0 / (double)F, rms_dev);
    launches += 2;
    cudaMemcpyAsync(&rms, rms_dev, sizeof(double), cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
    ++iter;
  } while (rms > tol && iter <= maxiter);  // alignment.cpp:399
  if (e == cudaSuccess) {
    // averageStructure(subset, xforms, traj, indices): ensembles.cpp:91-121
    cudaMemsetAsync(dAvgR, 0, sizeof(double) * 3 * NpR, st);
    a.mode = M_ACCUM; a.accum_align_set = 0; a.avg = dAvgR;
    k_stream<<<grid_for((uint32_t)F, true), ST, smemR, st>>>(a);
    k_scale<<<(unsigned)((3 * NpR + 255) / 256), 256, 0, st>>>(dAvgR, 3 * NpR, 1.0 / (double)F);
    // per-frame rmsd to the average: rmsd2ref.cpp:238-245
    a.mode = M_RMSD; a.refR = dAvgR; a.avg = nullptr;
    k_stream_warp<<<grid_warp((uint32_t)F), SW_THREADS, 0, st>>>(a);
    launches += 3;
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
