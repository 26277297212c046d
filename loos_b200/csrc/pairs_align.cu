// All-to-all RMSD with DIFFERENT alignment and RMSD selections
// (Packages/Clustering/rmsds-align.py:236-256).
//
// For frame i of trajectory 1 and frame j of trajectory 2 the reference script does
//   W = ref_align.alignOnto(align_selection_2)      (AtomicGroup::superposition, src/AG_linalg.cpp:188-197)
//   ref_target.applyTransform(W)                    (src/AG_numerical.cpp:363-370)
//   rmsd_selection_2.rmsd(ref_target)               (src/AG_numerical.cpp:301-318)
// where ref_align / ref_target are copies of frame i's align / rmsd selections that keep the
// transforms of the earlier j (a composition of optimal superpositions is the optimal
// superposition).  With ca_i, cb_j the centroids of the two ALIGN selections, x' = x - ca_i and
// y' = y - cb_j the rmsd-selection atoms about them, and Z_ij the rotation that superposes the
// align selection of i onto that of j,
//   d_ij^2 = ( sum |x'|^2 + sum |y'|^2 - 2 sum_ab Z_ab C_ba ) / N_r ,   C_ab = sum_k x'_a y'_b .
// So a pair needs two 3x3 contractions -- R over the align selection (it gives Z through the
// quartic / adjugate solve of qcp.cuh) and C over the rmsd selection -- and nothing per atom
// after that.  Two engines:
//   fp64    both contractions on the fp64 CUDA cores, a 40 x 40 tile of pairs per CTA and a 2 x 2 block
//           of pairs per thread, as pairs_fp64.cu does for the one-selection case (k_pairs_align below;
//           roofline: fp64 pipe, 18 (N_a + N_r) flop per pair; agrees with the reference to 1e-8 A)
//   tensor  both contractions on the int8 tensor path of pairs_tensor.cu (exact integer sums of the
//           24-bit fixed-point coordinates): pass 1 over the align selections leaves the rotation of
//           every pair as a unit quaternion, pass 2 over the rmsd selections -- quantised about the
//           ALIGN centroids -- turns it into d_ij (rmsds_align_tensor below; error = the coordinate grid,
//           ~1e-6 A)
// The result is not symmetric and is kept in double: the script prints Python doubles.
#include <limits.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "qcp.cuh"

namespace loosb200 {

namespace {

constexpr int TF = TILE_FRAMES;  // 40
constexpr int KC = 16;           // atoms per smem chunk
constexpr int TFP = 42;          // row pitch of the smem tiles in doubles: 16-byte aligned pairs, and the k-fastest fill
                                 // hits 8 distinct bank pairs instead of 2 (pitch 40: 16-way conflicts on every store)
constexpr int NT = 416;          // 13 warps; threads 0..399 own a 2x2 block of pairs

// g[f] = sum over the set's atoms of |x - cen[f]|^2 with cen taken from ANOTHER set (the align
// selection of the same frame).  One warp per frame.
__global__ void __launch_bounds__(128)
k_g_about(const float* __restrict__ raw, size_t Npad, uint32_t N, uint32_t F,
          const double* __restrict__ cen, double* __restrict__ g) {
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t f = blockIdx.x * 4 + (threadIdx.x >> 5); f < F; f += gridDim.x * 4) {
    double s = 0.0;
    for (int ax = 0; ax < 3; ++ax) {
      const double c = cen[(size_t)f * 3 + ax];
      const float* p = raw + ((size_t)f * 3 + ax) * Npad;
      for (uint32_t k = lane; k < N; k += 32) {
        const double v = (double)p[k] - c;
        s = fma(v, v, s);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) g[f] = s;
  }
}

// maxabs[f] = max over the set's atoms and axes of |x - cen[f]| with cen taken from ANOTHER set: what
// k_unit needs to put the rmsd selection on a fixed-point grid about the align centroid.  A non-finite
// coordinate marks the frame with +Inf, as k_stage does.  One warp per frame.
__global__ void __launch_bounds__(128)
k_maxabs_about(const float* __restrict__ raw, size_t Npad, uint32_t N, uint32_t F,
               const double* __restrict__ cen, float* __restrict__ maxabs) {
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t f = blockIdx.x * 4 + (threadIdx.x >> 5); f < F; f += gridDim.x * 4) {
    float m = 0.f;
    bool bad = false;
    for (int ax = 0; ax < 3; ++ax) {
      const double c = cen[(size_t)f * 3 + ax];
      const float* p = raw + ((size_t)f * 3 + ax) * Npad;
      for (uint32_t k = lane; k < N; k += 32) {
        const float v = fabsf((float)((double)p[k] - c));
        bad = bad || !(v <= 3.0e38f);
        m = fmaxf(m, v);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) maxabs[f] = bad ? __int_as_float(0x7f800000) : m * 1.0000002f;  // fp32 rounding of |x - c|
  }
}

struct AlignArgs {
  const float *alnA, *alnB, *rmsA, *rmsB;  // raw planes of the four staged sets
  size_t padAlnA, padAlnB, padRmsA, padRmsB;
  const double *cenA, *cenB;               // centroids of the ALIGN sets
  const double *gdA, *gdB;                 // sum of squares of the centred align sets
  const double *gxA, *gxB;                 // rmsd-set sum of squares about the align centroid
  uint32_t FA, FB, Na, Nr;
  uint32_t row0;                           // first frame of this band of rows
  uint32_t band_rows;                      // frames in the band (= leading dimension of out)
  double* out;                             // band, column-major: out[j * band_rows + (i - row0)]
};

// 13 warps put 4 on one SM sub-partition (16,384 registers): 128 registers per thread is the most that
// launches (144 and 152 fail with cudaErrorLaunchOutOfResources), which is what __launch_bounds__ settles on
__global__ void __launch_bounds__(NT)
k_pairs_align(const AlignArgs a) {
  __shared__ __align__(16) double As[3][KC][TFP];
  __shared__ __align__(16) double Bs[3][KC][TFP];
  __shared__ double cA[TF][3], cB[TF][3];

  const uint32_t I0 = a.row0 + blockIdx.y * TF, J0 = blockIdx.x * TF;
  const int tid = threadIdx.x;
  const bool worker = tid < TF * TF / 4;
  const int ty = tid / 20, tx = tid % 20;

  for (int t = tid; t < TF * 3; t += NT) {
    const uint32_t fa = I0 + t / 3, fb = J0 + t / 3;
    cA[t / 3][t % 3] = fa < a.FA ? a.cenA[(size_t)fa * 3 + t % 3] : 0.0;
    cB[t / 3][t % 3] = fb < a.FB ? a.cenB[(size_t)fb * 3 + t % 3] : 0.0;
  }
  __syncthreads();

  double acc[2][2][9];
  double quat[2][2][4];  // the rotations of phase 0, kept as unit quaternions while phase 1 accumulates

  // phase 0: R over the align selection -> Z;  phase 1: C over the rmsd selection
  for (int phase = 0; phase < 2; ++phase) {
    const float* rawA = phase ? a.rmsA : a.alnA;
    const float* rawB = phase ? a.rmsB : a.alnB;
    const size_t padA = phase ? a.padRmsA : a.padAlnA, padB = phase ? a.padRmsB : a.padAlnB;
    const uint32_t N = phase ? a.Nr : a.Na;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[p][q][k] = 0.0;

    // the chunk after the one being multiplied is already on its way from global memory (registers):
    // with one CTA per SM nothing else would hide that latency.  A chunk is 120 rows (frame, axis) of
    // 16 atoms per operand = 2 x 480 float4; item = tid + r * NT names (operand, row, quarter).
    constexpr int ITEMS = 2 * TF * 3 * (KC / 4);       // 960
    constexpr int PF = (ITEMS + NT - 1) / NT;          // 3 float4 per thread and chunk
    float4 pv[PF];
    auto prefetch = [&](uint32_t k0) {
#pragma unroll
      for (int r = 0; r < PF; ++r) {
        const int item = tid + r * NT;
        const int op = item / (ITEMS / 2), within = item % (ITEMS / 2), row = within >> 2, qd = within & 3;
        const uint32_t fr = (op ? J0 : I0) + row / 3, at = k0 + 4 * qd;
        // Npad is a multiple of 4, so a quarter that starts below N lies inside the row
        const bool in = item < ITEMS && at < N && fr < (op ? a.FB : a.FA);
        const float* src = (op ? rawB : rawA) + ((size_t)fr * 3 + row % 3) * (op ? padB : padA) + at;
        pv[r] = in ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    prefetch(0);
    for (uint32_t k0 = 0; k0 < N; k0 += KC) {
#pragma unroll
      for (int r = 0; r < PF; ++r) {
        const int item = tid + r * NT;
        if (item < ITEMS) {
          const int op = item / (ITEMS / 2), within = item % (ITEMS / 2), row = within >> 2, qd = within & 3;
          const int f = row / 3, ax = row % 3;
          const bool fr_ok = (op ? J0 : I0) + f < (op ? a.FB : a.FA);
          const double c = op ? cB[f][ax] : cA[f][ax];
          double (*dst)[TFP] = op ? Bs[ax] : As[ax];
          const float v[4] = {pv[r].x, pv[r].y, pv[r].z, pv[r].w};
#pragma unroll
          for (int e = 0; e < 4; ++e)  // padded atoms and frames beyond the sets contribute exact zeros
            dst[4 * qd + e][f] = fr_ok && k0 + 4 * qd + e < N ? (double)v[e] - c : 0.0;
        }
      }
      if (k0 + KC < N) prefetch(k0 + KC);
      __syncthreads();
      if (worker) {
#pragma unroll 4
        for (int k = 0; k < KC; ++k) {
          double a0[3], a1[3], b0[3], b1[3];
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {  // frames 2 ty, 2 ty + 1 and 2 tx, 2 tx + 1: one 16-byte load each
            const double2 av = *reinterpret_cast<const double2*>(&As[ax][k][2 * ty]);
            const double2 bv = *reinterpret_cast<const double2*>(&Bs[ax][k][2 * tx]);
            a0[ax] = av.x; a1[ax] = av.y;
            b0[ax] = bv.x; b1[ax] = bv.y;
          }
#pragma unroll
          for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              acc[0][0][3 * p + q] = fma(a0[p], b0[q], acc[0][0][3 * p + q]);
              acc[0][1][3 * p + q] = fma(a0[p], b1[q], acc[0][1][3 * p + q]);
              acc[1][0][3 * p + q] = fma(a1[p], b0[q], acc[1][0][3 * p + q]);
              acc[1][1][3 * p + q] = fma(a1[p], b1[q], acc[1][1][3 * p + q]);
            }
        }
      }
      __syncthreads();
    }

    if (!worker) continue;
    if (phase == 0) {
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint32_t i = I0 + 2 * ty + p, j = J0 + 2 * tx + q;
          if (i >= a.FA || j >= a.FB) continue;
          // Z maps the align selection of i onto that of j; the Jacobi solve takes the
          // (near-)degenerate cases -- collinear / planar selections, fewer than 3 atoms
          if (!qcp_rotation_fast(acc[p][q], a.gdA[i], a.gdB[j], nullptr, nullptr, quat[p][q])) qcp_rotation(acc[p][q], nullptr, nullptr, quat[p][q]);
        }
    } else {
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint32_t i = I0 + 2 * ty + p, j = J0 + 2 * tx + q;
          if (i >= a.FA || j >= a.FB || i >= a.row0 + a.band_rows) continue;
          const double* C = acc[p][q];
          double z[9];
          qcp_quat_to_rot(quat[p][q], z);
          double tr = 0.0;  // sum_k (Z x'_k) . y'_k = sum_ab Z_ab C_ba
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) tr = fma(z[3 * r + c], C[3 * c + r], tr);
          const double ss = a.gxA[i] + a.gxB[j] - 2.0 * tr;
          a.out[(size_t)j * a.band_rows + (i - a.row0)] = sqrt(fmax(ss, 0.0) / (double)a.Nr);
        }
    }
  }
}

}  // namespace
}  // namespace loosb200

// ---- tensor engine ----------------------------------------------------------------------------------
namespace loosb200 {
namespace {

// The rmsd selection of a trajectory seen from its ALIGN centroids: shares the fp32 planes of the staged
// rmsd set, borrows the centroids of the align set, owns the digit planes derived from the two.
struct AboutView {
  loosb200_frames f;
  bool owns = false;
  int build(const loosb200_frames* rmsd, const loosb200_frames* align, cudaStream_t st) {
    f.device = rmsd->device; f.F = rmsd->F; f.N = rmsd->N; f.Npad = rmsd->Npad;
    f.raw = rmsd->raw; f.centroid = align->centroid; f.gd = nullptr; f.stream = st;
    if (dev_alloc_t(&f.maxabs, sizeof(float) * f.F) != cudaSuccess) { cudaGetLastError(); return LOOSB200_ERR_NOMEM; }
    owns = true;
    k_maxabs_about<<<(unsigned)std::min<size_t>((f.F + 3) / 4, 148 * 16), 128, 0, st>>>(f.raw, f.Npad, (uint32_t)f.N, (uint32_t)f.F,
                                                                                  f.centroid, f.maxabs);
    LB_CUDA(cudaGetLastError());
    return LOOSB200_OK;
  }
  ~AboutView() {
    if (!owns) return;
    cudaStreamSynchronize(f.stream);  // an early error return must not hand buffers of running kernels back to the cache
    cudaGetLastError();
    dev_release(f.maxabs); dev_release(f.q8); dev_release(f.q8b); dev_release(f.gq); dev_release(f.unit_log2_dev);
    f.raw = nullptr; f.centroid = nullptr;
  }
};

// both sets on one fixed-point grid (the coarser of the two)
int common_grid(loosb200_frames* a, loosb200_frames* b) {
  int rc = ensure_quantised(a, INT_MIN);
  if (rc || a == b) return rc;
  if ((rc = ensure_quantised(b, INT_MIN))) return rc;
  const int e = std::max(a->unit_log2, b->unit_log2);
  if ((rc = ensure_quantised(a, e))) return rc;
  return ensure_quantised(b, e);
}

int rmsds_align_tensor(loosb200_frames* align1, loosb200_frames* rmsd1, loosb200_frames* align2,
                       loosb200_frames* rmsd2, double* out, size_t ld) {
  NvtxRange nvtx("loosb200 rmsds-align (tensor engine)");
  cudaStream_t st = align1->stream;
  const size_t FA = align1->F, FB = align2->F;
  cudaEventRecord(align1->timing.ev[0], st);
  cudaEventRecord(align1->timing.ev[1], st);  // the fixed-point images of the four sets count as this engine's work
  int rc = common_grid(align1, align2);
  if (rc) return rc;
  // pass-2 operands: the align sets themselves when the rmsd selection IS the align selection
  AboutView v1, v2;
  loosb200_frames* x1 = align1;
  loosb200_frames* x2 = align2;
  if (rmsd1 != align1) { if ((rc = v1.build(rmsd1, align1, st))) return rc; x1 = &v1.f; }
  if (rmsd2 == rmsd1 && align2 == align1) x2 = x1;
  else if (rmsd2 != align2) { if ((rc = v2.build(rmsd2, align2, st))) return rc; x2 = &v2.f; }
  if ((x1 != align1 || x2 != align2) && (rc = common_grid(x1, x2))) return rc;

  const size_t row_tiles = (FA + TF - 1) / TF;
  // a band holds 32 B of quaternion and 8 B of result per pair: ~2 GiB at most
  size_t band_tiles = std::max<size_t>(1, (size_t(2) << 30) / ((size_t)40 * TF * std::max<size_t>(FB, 1)));
  if (const char* env = getenv("LOOSB200_ALIGN_BAND_TILES")) band_tiles = std::max<long>(1, atol(env));
  band_tiles = std::min(band_tiles, row_tiles);
  const size_t ldq = band_tiles * TF;

  std::vector<uint32_t> rows(row_tiles);
  for (size_t t = 0; t < row_tiles; ++t) rows[t] = (uint32_t)t;
  uint32_t* d_rows = nullptr;
  double *d_quat = nullptr, *d_band = nullptr;
  std::vector<void*> scratch;
  auto cleanup = [&]() {
    dev_release(d_rows); dev_release(d_quat); dev_release(d_band);
    for (void* p : scratch) dev_release(p);
  };
  if (dev_alloc_t(&d_rows, sizeof(uint32_t) * row_tiles) != cudaSuccess ||
      dev_alloc_t(&d_quat, sizeof(double) * 4 * ldq * FB) != cudaSuccess ||
      dev_alloc_t(&d_band, sizeof(double) * ldq * FB) != cudaSuccess) {
    cudaGetLastError(); cleanup(); return LOOSB200_ERR_NOMEM;
  }
  cudaMemcpyAsync(d_rows, rows.data(), sizeof(uint32_t) * row_tiles, cudaMemcpyHostToDevice, st);
  int launches = 0;
  const uint32_t col_tiles = (uint32_t)((FB + COL_TILE_FRAMES - 1) / COL_TILE_FRAMES);
  for (size_t t0 = 0; t0 < row_tiles && rc == LOOSB200_OK; t0 += band_tiles) {
    const size_t nt = std::min(band_tiles, row_tiles - t0);
    PairJob job;
    job.self = false; job.symmetric = false; job.out = nullptr; job.ld = 0;
    job.rows_host = rows.data() + t0; job.row_tiles_dev = d_rows + t0; job.tile_prefix_dev = nullptr;
    job.n_row_tiles = (uint32_t)nt; job.n_tiles = (uint64_t)nt * col_tiles;
    job.packed_rows = true; job.stats_dev = nullptr; job.scratch = &scratch;
    job.quat = d_quat; job.out_d = d_band; job.ldq = ldq;
    int l = 0;
    job.a = align1; job.b = align2; job.kind = PairJob::Rotation;
    rc = launch_pairs_tensor(job, st, &l);
    launches += l;
    if (rc) break;
    job.a = x1; job.b = x2; job.kind = PairJob::Trace;
    rc = launch_pairs_tensor(job, st, &l);
    launches += l;
    if (rc) break;
    cudaEventRecord(align1->timing.ev[2], st);  // re-recorded per band, as the fp64 engine does
    const size_t row0 = t0 * TF, band_rows = std::min<size_t>(nt * TF, FA - row0);
    cudaError_t e = cudaMemcpy2DAsync(out + row0, ld * sizeof(double), d_band, ldq * sizeof(double),
                                      band_rows * sizeof(double), FB, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && t0 + band_tiles < row_tiles) e = cudaStreamSynchronize(st);  // the band buffers are reused
    if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "rmsds_align (tensor): band copy", __FILE__, __LINE__); }
  }
  cudaEventRecord(align1->timing.ev[3], st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  cleanup();
  if (rc) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "rmsds_align (tensor) kernels", __FILE__, __LINE__);
  align1->timing.launches = launches; align1->timing.valid = true; align1->timing.kev_used = 0;
  return LOOSB200_OK;
}

}  // namespace
}  // namespace loosb200

using namespace loosb200;

extern "C" int loosb200_rmsds_align_engine(loosb200_frames* align1, loosb200_frames* rmsd1, loosb200_frames* align2,
                                           loosb200_frames* rmsd2, double* out, size_t ld, int engine) {
  if (!align1 || !align2 || !out) return LOOSB200_ERR_INVALID;
  if (engine != LOOSB200_ENGINE_AUTO && engine != LOOSB200_ENGINE_TENSOR && engine != LOOSB200_ENGINE_FP64) return LOOSB200_ERR_INVALID;
  if (!rmsd1) rmsd1 = align1;
  if (!rmsd2) rmsd2 = align2;
  const loosb200_frames* all[4] = {align1, rmsd1, align2, rmsd2};
  for (const loosb200_frames* h : all)
    if (h->stager || !h->raw) { set_last_error("rmsds_align: a frame set is not staged"); return LOOSB200_ERR_INVALID; }
  if (rmsd1->F != align1->F || rmsd2->F != align2->F) { set_last_error("rmsds_align: align and rmsd sets of one trajectory differ in frame count"); return LOOSB200_ERR_SIZE_MISMATCH; }
  // AtomicGroup::superposition / ::rmsd throw on unequal sizes (src/AG_linalg.cpp:189-192, src/AG_numerical.cpp:303-304)
  if (align1->N != align2->N || rmsd1->N != rmsd2->N) { set_last_error("rmsds_align: selections of the two trajectories differ in size"); return LOOSB200_ERR_SIZE_MISMATCH; }
  for (const loosb200_frames* h : all)
    if (h->device != align1->device) { set_last_error("rmsds_align: frame sets live on different devices"); return LOOSB200_ERR_INVALID; }
  const size_t FA = align1->F, FB = align2->F;
  if (FA == 0 || FB == 0) return LOOSB200_OK;
  if (ld < FA) return LOOSB200_ERR_INVALID;
  if (FA > 0xffffff00ull || FB > 0xffffff00ull || (FB + TF - 1) / TF > 0x7fffffffull) return LOOSB200_ERR_UNSUPPORTED;
  LB_CUDA(cudaSetDevice(align1->device));
  cudaStream_t st = align1->stream;
  // the frame sets were staged on their own streams: wait for all of them
  for (const loosb200_frames* h : all)
    if (h->stream != st) {
      cudaError_t e = cudaStreamSynchronize(h->stream);
      if (e != cudaSuccess) return cuda_fail(e, "rmsds_align: staging", __FILE__, __LINE__);
    }
  if (engine != LOOSB200_ENGINE_FP64) {
    const bool can = tensor_path_available() && align1->N <= TENSOR_MAX_ATOMS && rmsd1->N <= TENSOR_MAX_ATOMS;
    if (!can && engine == LOOSB200_ENGINE_TENSOR) return LOOSB200_ERR_UNSUPPORTED;
    if (can) {
      const int rc = rmsds_align_tensor(align1, rmsd1, align2, rmsd2, out, ld);
      // NaN / Inf coordinates have no fixed-point image; the fp64 engine propagates them like the reference
      if (!(rc == LOOSB200_ERR_NUMERICAL && engine == LOOSB200_ENGINE_AUTO)) return rc;
    }
  }

  // bands of whole row tiles, at most ~1 GiB of doubles each (and 65535 tiles: gridDim.y)
  size_t band_tiles = std::max<size_t>(1, (size_t(1) << 30) / (sizeof(double) * TF * FB));
  band_tiles = std::min<size_t>(band_tiles, 65535);
  if (const char* env = getenv("LOOSB200_ALIGN_BAND_TILES")) band_tiles = std::max<long>(1, atol(env));  // tests: force several bands
  const size_t row_tiles = (FA + TF - 1) / TF;
  band_tiles = std::min(band_tiles, row_tiles);
  const size_t band_rows_max = std::min(FA, band_tiles * TF);

  double *gxA = nullptr, *gxB = nullptr, *d_band = nullptr;
  bool ok = dev_alloc_t(&gxA, sizeof(double) * FA) == cudaSuccess &&
            dev_alloc_t(&gxB, sizeof(double) * FB) == cudaSuccess &&
            dev_alloc_t(&d_band, sizeof(double) * band_rows_max * FB) == cudaSuccess;
  auto cleanup = [&]() { dev_release(gxA); dev_release(gxB); dev_release(d_band); };
  if (!ok) { cudaGetLastError(); cleanup(); return LOOSB200_ERR_NOMEM; }

  cudaEventRecord(align1->timing.ev[0], st);
  int launches = 0;
  k_g_about<<<(unsigned)std::min<size_t>((FA + 3) / 4, 148 * 16), 128, 0, st>>>(rmsd1->raw, rmsd1->Npad, (uint32_t)rmsd1->N, (uint32_t)FA, align1->centroid, gxA);
  k_g_about<<<(unsigned)std::min<size_t>((FB + 3) / 4, 148 * 16), 128, 0, st>>>(rmsd2->raw, rmsd2->Npad, (uint32_t)rmsd2->N, (uint32_t)FB, align2->centroid, gxB);
  launches += 2;
  cudaEventRecord(align1->timing.ev[1], st);

  AlignArgs a;
  a.alnA = align1->raw; a.alnB = align2->raw; a.rmsA = rmsd1->raw; a.rmsB = rmsd2->raw;
  a.padAlnA = align1->Npad; a.padAlnB = align2->Npad; a.padRmsA = rmsd1->Npad; a.padRmsB = rmsd2->Npad;
  a.cenA = align1->centroid; a.cenB = align2->centroid;
  a.gdA = align1->gd; a.gdB = align2->gd;
  a.gxA = gxA; a.gxB = gxB;
  a.FA = (uint32_t)FA; a.FB = (uint32_t)FB; a.Na = (uint32_t)align1->N; a.Nr = (uint32_t)rmsd1->N;
  a.out = d_band;
  cudaError_t e = cudaSuccess;
  for (size_t t0 = 0; t0 < row_tiles && e == cudaSuccess; t0 += band_tiles) {
    const size_t nt = std::min(band_tiles, row_tiles - t0);
    a.row0 = (uint32_t)(t0 * TF);
    a.band_rows = (uint32_t)std::min<size_t>(nt * TF, FA - a.row0);
    const dim3 grid((unsigned)((FB + TF - 1) / TF), (unsigned)nt);
    k_pairs_align<<<grid, NT, 0, st>>>(a);
    ++launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) break;
    cudaEventRecord(align1->timing.ev[2], st);  // re-recorded per band: the kernels (and, with several bands, the copies between them)
    // band (column-major, ld = band_rows) -> rows [row0, row0 + band_rows) of the caller's matrix
    e = cudaMemcpy2DAsync(out + a.row0, ld * sizeof(double), d_band, (size_t)a.band_rows * sizeof(double),
                          (size_t)a.band_rows * sizeof(double), FB, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && t0 + band_tiles < row_tiles) e = cudaStreamSynchronize(st);  // the band buffer is reused
  }
  cudaEventRecord(align1->timing.ev[3], st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  cleanup();
  if (e != cudaSuccess) return cuda_fail(e, "rmsds_align kernels", __FILE__, __LINE__);
  align1->timing.launches = launches; align1->timing.valid = true; align1->timing.kev_used = 0;
  return LOOSB200_OK;
}

extern "C" int loosb200_rmsds_align(loosb200_frames* align1, loosb200_frames* rmsd1, loosb200_frames* align2,
                                    loosb200_frames* rmsd2, double* out, size_t ld) {
  return loosb200_rmsds_align_engine(align1, rmsd1, align2, rmsd2, out, ld, LOOSB200_ENGINE_FP64);
}
