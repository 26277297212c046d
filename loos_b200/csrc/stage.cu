// Frame staging: the device-side counterpart of readCoords() + centerAtOrigin()
// (src/ensembles.cpp:253-296, src/alignment.cpp:90-109).
//
//   k_stage     gathers the selected atoms of each frame (Trajectory::
//               updateGroupCoords, src/dcd.cpp:413-425) into fp32 SoA planes and
//               computes the fp64 centroid, centred sum of squares and max |x-c|.
//   k_unit      picks the power-of-two fixed-point unit from the global max.
//   k_quantise  centres in fp64, rounds to the 24-bit fixed-point grid, splits
//               into three signed 8-bit digits and writes them as K-major operand
//               planes for tcgen05, once with the "a" row layout (row = 32*(f/10) +
//               3*(f%10) + axis: a frame's axes in one warp's TMEM lanes) and once dense
//               (row = 3*f + axis: the "b" operand), together with the exact integer
//               sum of squares G.
//
// All three are HBM-bound byte movers: one CTA per frame, coalesced reads of the
// source frame, coalesced writes of the planes.
#include <limits.h>
#include <string.h>

#include <new>

#include "common.cuh"

namespace loosb200 {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum `NV` doubles across the block; result valid in every thread.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* >= NV*32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[i * 32 + warp] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double t = (lane < nwarp) ? smem[i * 32 + lane] : 0.0;
    v[i] = warp_sum(t);
  }
}

template <int LAYOUT>
__global__ void __launch_bounds__(256)
k_stage(const float* __restrict__ src, size_t natoms, const uint32_t* __restrict__ sel, int nsel,
        size_t Npad, float* __restrict__ raw, double* __restrict__ centroid,
        double* __restrict__ gd, float* __restrict__ maxabs) {
  __shared__ double red[4 * 32];
  __shared__ float redf[32];
  const size_t f = blockIdx.x;
  const float* s = src + f * natoms * 3;
  float* px = raw + (f * 3 + 0) * Npad;
  float* py = raw + (f * 3 + 1) * Npad;
  float* pz = raw + (f * 3 + 2) * Npad;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < (int)Npad; i += blockDim.x) {
    float x = 0.f, y = 0.f, z = 0.f;
    if (i < nsel) {
      const size_t a = sel ? sel[i] : (size_t)i;
      if (LAYOUT == LOOSB200_LAYOUT_XYZ) {
        x = s[3 * a]; y = s[3 * a + 1]; z = s[3 * a + 2];
      } else {
        x = s[a]; y = s[natoms + a]; z = s[2 * natoms + a];
      }
      acc[0] += (double)x; acc[1] += (double)y; acc[2] += (double)z;
    }
    px[i] = x; py[i] = y; pz[i] = z;
  }
  block_sum<3>(acc, red);
  const double cx = acc[0] / nsel, cy = acc[1] / nsel, cz = acc[2] / nsel;
  // second pass over this thread's own writes (same thread, same addresses)
  double g[1] = {0.0};
  float m = 0.f;
  for (int i = threadIdx.x; i < nsel; i += blockDim.x) {
    const double dx = (double)px[i] - cx, dy = (double)py[i] - cy, dz = (double)pz[i] - cz;
    g[0] += dx * dx + dy * dy + dz * dz;
    m = fmaxf(m, fmaxf(fabsf((float)dx), fmaxf(fabsf((float)dy), fabsf((float)dz))));
  }
  block_sum<1>(g, red);
  m = warp_max(m);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = 0.f;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) mm = fmaxf(mm, redf[w]);
    centroid[f * 3 + 0] = cx; centroid[f * 3 + 1] = cy; centroid[f * 3 + 2] = cz;
    gd[f] = g[0];
    // a NaN / Inf coordinate makes the centred sum of squares non-finite (fmaxf alone would drop
    // the NaN): mark the frame so that k_unit refuses the fixed-point path for this set
    maxabs[f] = (g[0] <= 1.7e308) ? mm : __int_as_float(0x7f800000);
  }
}

// unit = 2^e with max|x-c| / unit <= 127 * 65793 - 2: three balanced digits in [-128, 127] hold
// -128 * 65793 .. 127 * 65793 = 8355711 (NOT 2^23 - 1: the carry of the lower digits pushes a value
// above that into a top digit of +128); 2 units of headroom cover the rounding to the grid and
// the fp32 rounding of maxabs.  force != INT_MIN overrides (cross mode: common unit).
// e_out[1]: 0 fine, 1 forced unit too fine for these coordinates, 2 non-finite / absurd coordinates.
__global__ void k_unit(const float* __restrict__ maxabs, size_t F, int force, int* __restrict__ e_out) {
  __shared__ float redf[32];
  float m = 0.f;
  for (size_t i = threadIdx.x; i < F; i += blockDim.x) m = fmaxf(m, maxabs[i]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, redf[w]);
    int e = -20;  // never finer than 2^-20 A (float32 inputs carry no more)
    const double lim = (double)Q_MAX - 2.0;
    while (e < 40 && (double)m / exp2((double)e) > lim) ++e;
    const bool absurd = !((double)m / exp2((double)e) <= lim);  // Inf (NaN frames are marked Inf) or > 2^63 A
    if (force != INT_MIN && force > e) e = force;
    e_out[0] = e;
    e_out[1] = absurd ? 2 : ((force != INT_MIN && force < e) ? 1 : 0);  // forced unit too fine: range error
  }
}

__global__ void __launch_bounds__(256)
k_quantise(const float* __restrict__ raw, const double* __restrict__ centroid, size_t N,
           size_t Npad, size_t rows_total, size_t rows_total_b, size_t Kpad,
           const int* __restrict__ unit_log2, int8_t* __restrict__ q8, int8_t* __restrict__ q8b,
           double* __restrict__ gq) {
  __shared__ long long red[32];
  const size_t f = blockIdx.x;
  const double inv_unit = exp2(-(double)unit_log2[0]);
  const size_t row0 = (f / FRAMES_PER_BLOCK) * ROWS_PER_BLOCK + (f % FRAMES_PER_BLOCK) * 3;
  const size_t plane = rows_total * Kpad, plane_b = rows_total_b * Kpad;
  long long g = 0;
  // four atoms per thread and trip: one 16-byte load, six 4-byte stores (digit planes are Kpad-pitched,
  // Kpad a multiple of 128; rows of `raw` are Npad-pitched, Npad a multiple of 4).  Byte-wise stores made
  // this kernel 1.5 ms at C4 (3 GB of traffic).
  const int n4 = (int)((N + 3) / 4);
  for (int axis = 0; axis < 3; ++axis) {
    const float4* p = reinterpret_cast<const float4*>(raw + (f * 3 + axis) * Npad);
    const double c = centroid[f * 3 + axis];
    uint32_t* d0 = reinterpret_cast<uint32_t*>(q8 + (row0 + axis) * Kpad);
    uint32_t* b0 = reinterpret_cast<uint32_t*>(q8b + (f * 3 + axis) * Kpad);
    for (int i4 = threadIdx.x; i4 < n4; i4 += blockDim.x) {
      const float4 v4 = p[i4];
      const float vs[4] = {v4.x, v4.y, v4.z, v4.w};
      uint32_t w0 = 0, w1 = 0, w2 = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if ((size_t)(4 * i4 + k) < N) {  // pad atoms keep zero digits and do not count in G
          const long long q = __double2ll_rn(((double)vs[k] - c) * inv_unit);
          g += q * q;
          // balanced base-256 digits: q = x0 + 256 x1 + 65536 x2, each in [-128, 127]
          const int x0 = (int)(int8_t)(q & 0xff);
          const long long q1 = (q - x0) >> 8;
          const int x1 = (int)(int8_t)(q1 & 0xff);
          long long x2 = (q1 - x1) >> 8;
          x2 = x2 > 127 ? 127 : (x2 < -128 ? -128 : x2);  // unreachable below k_unit's limit; never wrap
          w0 |= (uint32_t)(uint8_t)x0 << (8 * k);
          w1 |= (uint32_t)(uint8_t)x1 << (8 * k);
          w2 |= (uint32_t)(uint8_t)(int)x2 << (8 * k);
        }
      }
      d0[i4] = w0;
      d0[plane / 4 + i4] = w1;
      d0[2 * (plane / 4) + i4] = w2;
      b0[i4] = w0;
      b0[plane_b / 4 + i4] = w1;
      b0[2 * (plane_b / 4) + i4] = w2;
    }
  }
  // exact integer block sum
  for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = g;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    gq[f] = (double)t;
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}

int ensure_quantised(loosb200_frames* h, int force_unit_log2) {
  if (h->quantised && (force_unit_log2 == INT_MIN || force_unit_log2 == h->unit_log2))
    return LOOSB200_OK;
  if (h->N > TENSOR_MAX_ATOMS) return LOOSB200_ERR_UNSUPPORTED;
  NvtxRange nvtx("loosb200 quantise (fixed-point digit planes)");
  LB_CUDA(cudaSetDevice(h->device));
  const size_t nblocks = (h->F + FRAMES_PER_BLOCK - 1) / FRAMES_PER_BLOCK;
  const size_t ntiles = (nblocks + BLOCKS_PER_TILE - 1) / BLOCKS_PER_TILE;
  h->rows_total = ntiles * TILE_ROWS + TILE_ROWS;
  // dense "b" rows, a whole number of 96-row column tiles (plus one of slack): no TMA box leaves the tensor
  h->rows_total_b = ((h->F + COL_TILE_FRAMES - 1) / COL_TILE_FRAMES + 1) * COL_TILE_ROWS;
  h->Kpad = ((h->N + K_CHUNK - 1) / K_CHUNK) * K_CHUNK;
  const size_t bytes = (size_t)NUM_SLICES * h->rows_total * h->Kpad;
  const size_t bytes_b = (size_t)NUM_SLICES * h->rows_total_b * h->Kpad;
  if (!h->q8) {
    if (dev_alloc_t(&h->q8, bytes) != cudaSuccess) { cudaGetLastError(); return LOOSB200_ERR_NOMEM; }
    if (dev_alloc_t(&h->q8b, bytes_b) != cudaSuccess) { cudaGetLastError(); return LOOSB200_ERR_NOMEM; }
    if (dev_alloc_t(&h->gq, sizeof(double) * h->F) != cudaSuccess) { cudaGetLastError(); return LOOSB200_ERR_NOMEM; }
    if (dev_alloc_t(&h->unit_log2_dev, 2 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return LOOSB200_ERR_NOMEM; }
  }
  LB_CUDA(cudaMemsetAsync(h->q8, 0, bytes, h->stream));
  LB_CUDA(cudaMemsetAsync(h->q8b, 0, bytes_b, h->stream));
  k_unit<<<1, 1024, 0, h->stream>>>(h->maxabs, h->F, force_unit_log2, h->unit_log2_dev);
  k_quantise<<<(unsigned)h->F, 256, 0, h->stream>>>(h->raw, h->centroid, h->N, h->Npad, h->rows_total,
                                                    h->rows_total_b, h->Kpad, h->unit_log2_dev, h->q8, h->q8b, h->gq);
  LB_CUDA(cudaGetLastError());
  int e[2];
  LB_CUDA(cudaMemcpyAsync(e, h->unit_log2_dev, sizeof(e), cudaMemcpyDeviceToHost, h->stream));
  LB_CUDA(cudaStreamSynchronize(h->stream));
  if (e[1] == 2) return LOOSB200_ERR_NUMERICAL;  // NaN / Inf coordinates: the fp64 engine propagates them
  if (e[1]) return LOOSB200_ERR_RANGE;
  h->unit_log2 = e[0];
  h->quantised = true;
  h->timing.launches += 2;

  // 3-D tensor map (atom, row, slice); one TMA box = 128 atoms x 128 rows x 3 slices.
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  if (!enc) { set_last_error("cuTensorMapEncodeTiled not available"); return LOOSB200_ERR_CUDA; }
  cuuint32_t estr[3] = {1, 1, 1};
  for (int which = 0; which < 4; ++which) {
    const size_t nrows = which == 0 ? h->rows_total : h->rows_total_b;
    cuuint64_t dims[3] = {(cuuint64_t)h->Kpad, (cuuint64_t)nrows, (cuuint64_t)NUM_SLICES};
    cuuint64_t strides[2] = {(cuuint64_t)h->Kpad, (cuuint64_t)(nrows * h->Kpad)};
    const cuuint32_t rows[4] = {TILE_ROWS, COL_TILE_ROWS, COL_TILE_ROWS, COL_TILE_ROWS / 2};
    cuuint32_t box[3] = {(cuuint32_t)K_CHUNK, rows[which], (cuuint32_t)(which < 2 ? NUM_SLICES : 1)};
    CUtensorMap* dst = which == 0 ? &h->tmapA : which == 1 ? &h->tmapB : which == 2 ? &h->tmapB96 : &h->tmapB48;
    CUresult r = enc(dst, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, which == 0 ? h->q8 : h->q8b, dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
      return LOOSB200_ERR_CUDA;
    }
  }
  h->tmap_valid = true;
  return LOOSB200_OK;
}

// ---- host side of staging -------------------------------------------------

int alloc_frames(loosb200_frames** out, size_t F, size_t N, int device) {
  loosb200_frames* h = new (std::nothrow) loosb200_frames();
  if (!h) return LOOSB200_ERR_NOMEM;
  h->device = device; h->F = F; h->N = N;
  h->Npad = (N + 3) & ~(size_t)3;
  cudaError_t e;
  if ((e = cudaSetDevice(device)) != cudaSuccess) { delete h; return cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__); }
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) { delete h; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
  bool ok = dev_alloc_t(&h->raw, sizeof(float) * F * 3 * h->Npad) == cudaSuccess &&
            dev_alloc_t(&h->centroid, sizeof(double) * F * 3) == cudaSuccess &&
            dev_alloc_t(&h->gd, sizeof(double) * F) == cudaSuccess &&
            dev_alloc_t(&h->maxabs, sizeof(float) * F) == cudaSuccess;
  for (int i = 0; i < 4 && ok; ++i) ok = cudaEventCreate(&h->timing.ev[i]) == cudaSuccess;
  if (!ok) { cudaGetLastError(); loosb200_free(h); return LOOSB200_ERR_NOMEM; }
  *out = h;
  return LOOSB200_OK;
}

static void launch_stage(loosb200_frames* h, const float* dev_src, size_t f0, size_t nf, size_t natoms,
                         const uint32_t* sel_dev, int layout, cudaStream_t st) {
  float* raw = h->raw + f0 * 3 * h->Npad;
  if (layout == LOOSB200_LAYOUT_XYZ)
    k_stage<LOOSB200_LAYOUT_XYZ><<<(unsigned)nf, 256, 0, st>>>(dev_src, natoms, sel_dev, (int)h->N, h->Npad, raw,
                                                               h->centroid + f0 * 3, h->gd + f0, h->maxabs + f0);
  else
    k_stage<LOOSB200_LAYOUT_PLANES><<<(unsigned)nf, 256, 0, st>>>(dev_src, natoms, sel_dev, (int)h->N, h->Npad, raw,
                                                                  h->centroid + f0 * 3, h->gd + f0, h->maxabs + f0);
}

static int check_stage_args(loosb200_frames** out, const void* xyz, size_t F, size_t natoms,
                            const uint32_t* sel, size_t nsel, int layout, int device) {
  if (!out || !xyz || F == 0 || natoms == 0) return LOOSB200_ERR_INVALID;
  if (layout != LOOSB200_LAYOUT_XYZ && layout != LOOSB200_LAYOUT_PLANES) return LOOSB200_ERR_INVALID;
  if (sel && nsel == 0) return LOOSB200_ERR_INVALID;
  if (F > 0x7fffffffu || natoms > 0x3fffffffu) return LOOSB200_ERR_INVALID;
  if (device < 0 || device >= loosb200_device_count()) return LOOSB200_ERR_NO_DEVICE;
  if (sel)
    for (size_t i = 0; i < nsel; ++i)
      if (sel[i] >= natoms) return LOOSB200_ERR_RANGE;
  return LOOSB200_OK;
}

static int upload_sel(const uint32_t* sel, size_t nsel, uint32_t** dev, cudaStream_t st) {
  *dev = nullptr;
  if (!sel) return LOOSB200_OK;
  if (dev_alloc_t(dev, sizeof(uint32_t) * nsel) != cudaSuccess) { cudaGetLastError(); return LOOSB200_ERR_NOMEM; }
  LB_CUDA(cudaMemcpyAsync(*dev, sel, sizeof(uint32_t) * nsel, cudaMemcpyHostToDevice, st));
  return LOOSB200_OK;
}

}  // namespace loosb200

using namespace loosb200;

extern "C" int loosb200_stage_frames_device(loosb200_frames** out, const float* dev_xyz, size_t F,
                                            size_t natoms, const uint32_t* sel, size_t nsel,
                                            int layout, int device) {
  int rc = check_stage_args(out, dev_xyz, F, natoms, sel, nsel, layout, device);
  if (rc) return rc;
  loosb200_frames* h = nullptr;
  if ((rc = alloc_frames(&h, F, sel ? nsel : natoms, device))) return rc;
  uint32_t* sel_dev = nullptr;
  if ((rc = upload_sel(sel, nsel, &sel_dev, h->stream))) { loosb200_free(h); return rc; }
  cudaEventRecord(h->timing.ev[0], h->stream);
  launch_stage(h, dev_xyz, 0, F, natoms, sel_dev, layout, h->stream);
  cudaEventRecord(h->timing.ev[1], h->stream);
  cudaEventRecord(h->timing.ev[2], h->stream);
  cudaEventRecord(h->timing.ev[3], h->stream);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (sel_dev) dev_release(sel_dev);
  if (e != cudaSuccess) { loosb200_free(h); return cuda_fail(e, "k_stage", __FILE__, __LINE__); }
  h->timing.launches = 1; h->timing.valid = true;
  *out = h;
  return LOOSB200_OK;
}

// Host frames -> device: two pinned bounce buffers (allocated on first use: a pinned caller
// buffer is copied from directly) and two device chunk buffers;
// the CPU (trajectory decode + memcpy) fills bounce[k] while the copy engine drains
// bounce[k^1] on a side stream and the SMs gather the chunk before that.
namespace loosb200 {
struct Stager {
  size_t natoms = 0, F_total = 0, done = 0, chunk_frames = 0, frame_bytes = 0;
  int layout = 0;
  uint32_t* sel_dev = nullptr;
  float* bounce[2] = {nullptr, nullptr};
  float* dchunk[2] = {nullptr, nullptr};
  cudaEvent_t copied[2] = {nullptr, nullptr}, staged[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;
  int k = 0;
  bool started = false;
  ~Stager() {
    for (int i = 0; i < 2; ++i) {
      if (bounce[i]) host_release(bounce[i]);
      if (dchunk[i]) dev_release(dchunk[i]);
      if (copied[i]) cudaEventDestroy(copied[i]);
      if (staged[i]) cudaEventDestroy(staged[i]);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (sel_dev) dev_release(sel_dev);
    cudaGetLastError();
  }
};
}  // namespace loosb200

static int stager_push(loosb200_frames* h, Stager* s, const float* src, size_t nf, bool src_pinned) {
  // slot k is free once the gather kernel that read dchunk[k] (two pushes ago) has finished
  cudaEventSynchronize(s->staged[s->k]);
  const float* from = src;
  if (!src_pinned) {
    if (!s->bounce[s->k] &&
        host_alloc(reinterpret_cast<void**>(&s->bounce[s->k]), s->chunk_frames * s->frame_bytes) != cudaSuccess) {
      cudaGetLastError();
      return LOOSB200_ERR_NOMEM;
    }
    memcpy(s->bounce[s->k], src, nf * s->frame_bytes);
    from = s->bounce[s->k];
  }
  if (!s->started) { cudaEventRecord(h->timing.ev[0], h->stream); s->started = true; }
  cudaMemcpyAsync(s->dchunk[s->k], from, nf * s->frame_bytes, cudaMemcpyHostToDevice, s->copy_stream);
  cudaEventRecord(s->copied[s->k], s->copy_stream);
  cudaStreamWaitEvent(h->stream, s->copied[s->k], 0);
  launch_stage(h, s->dchunk[s->k], s->done, nf, s->natoms, s->sel_dev, s->layout, h->stream);
  cudaEventRecord(s->staged[s->k], h->stream);
  h->timing.launches++;
  s->done += nf;
  s->k ^= 1;
  return cudaGetLastError() == cudaSuccess ? LOOSB200_OK : LOOSB200_ERR_CUDA;
}

extern "C" int loosb200_stage_begin(loosb200_frames** out, size_t F, size_t natoms, const uint32_t* sel,
                                    size_t nsel, int layout, int device) {
  static const float dummy = 0.f;
  int rc = check_stage_args(out, &dummy, F, natoms, sel, nsel, layout, device);
  if (rc) return rc;
  loosb200_frames* h = nullptr;
  if ((rc = alloc_frames(&h, F, sel ? nsel : natoms, device))) return rc;
  Stager* s = new (std::nothrow) Stager();
  if (!s) { loosb200_free(h); return LOOSB200_ERR_NOMEM; }
  h->stager = s;
  s->natoms = natoms; s->F_total = F; s->layout = layout;
  s->frame_bytes = natoms * 3 * sizeof(float);
  s->chunk_frames = (size_t)(64u << 20) / s->frame_bytes;  // ~64 MiB per chunk
  if (s->chunk_frames < 1) s->chunk_frames = 1;
  if (s->chunk_frames > F) s->chunk_frames = F;
  if ((rc = upload_sel(sel, nsel, &s->sel_dev, h->stream))) { loosb200_free(h); return rc; }
  bool ok = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int k = 0; k < 2 && ok; ++k)
    ok = dev_alloc_t(&s->dchunk[k], s->chunk_frames * s->frame_bytes) == cudaSuccess &&
         cudaEventCreateWithFlags(&s->copied[k], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&s->staged[k], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { cudaGetLastError(); loosb200_free(h); return LOOSB200_ERR_NOMEM; }
  *out = h;
  return LOOSB200_OK;
}

static int stage_append_impl(loosb200_frames* h, const float* host_xyz, size_t nframes, bool allow_direct) {
  if (!h || !h->stager || !host_xyz) return LOOSB200_ERR_INVALID;
  Stager* s = h->stager;
  if (s->done + nframes > s->F_total) return LOOSB200_ERR_RANGE;
  LB_CUDA(cudaSetDevice(h->device));
  bool pinned = false;
  if (allow_direct) {  // one-shot staging from a pinned caller buffer: no bounce copy
    cudaPointerAttributes attr;
    pinned = cudaPointerGetAttributes(&attr, host_xyz) == cudaSuccess &&
             (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
  }
  for (size_t f0 = 0; f0 < nframes; f0 += s->chunk_frames) {
    const size_t nf = nframes - f0 < s->chunk_frames ? nframes - f0 : s->chunk_frames;
    int rc = stager_push(h, s, host_xyz + f0 * s->natoms * 3, nf, pinned);
    if (rc) return rc;
  }
  return LOOSB200_OK;
}

extern "C" int loosb200_stage_append(loosb200_frames* h, const float* host_xyz, size_t nframes) {
  return stage_append_impl(h, host_xyz, nframes, false);
}

extern "C" int loosb200_stage_end(loosb200_frames* h) {
  if (!h || !h->stager) return LOOSB200_ERR_INVALID;
  Stager* s = h->stager;
  LB_CUDA(cudaSetDevice(h->device));
  const bool complete = s->done == s->F_total;
  cudaEventRecord(h->timing.ev[1], h->stream);
  cudaEventRecord(h->timing.ev[2], h->stream);
  cudaEventRecord(h->timing.ev[3], h->stream);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  delete s;
  h->stager = nullptr;
  if (e != cudaSuccess) return cuda_fail(e, "stage_end", __FILE__, __LINE__);
  if (!complete) return LOOSB200_ERR_INVALID;  // fewer frames appended than announced
  h->timing.valid = true;
  return LOOSB200_OK;
}

extern "C" int loosb200_stage_frames(loosb200_frames** out, const float* host_xyz, size_t F,
                                     size_t natoms, const uint32_t* sel, size_t nsel, int layout,
                                     int device) {
  if (!host_xyz) return LOOSB200_ERR_INVALID;
  NvtxRange nvtx("loosb200 stage frames (H2D + gather + centre)");
  loosb200_frames* h = nullptr;
  int rc = loosb200_stage_begin(&h, F, natoms, sel, nsel, layout, device);
  if (rc) return rc;
  if ((rc = stage_append_impl(h, host_xyz, F, true)) || (rc = loosb200_stage_end(h))) { loosb200_free(h); return rc; }
  *out = h;
  return LOOSB200_OK;
}

extern "C" void loosb200_free(loosb200_frames* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->stager) { delete h->stager; h->stager = nullptr; }
  dev_release(h->raw); dev_release(h->centroid); dev_release(h->gd); dev_release(h->maxabs);
  dev_release(h->q8); dev_release(h->q8b); dev_release(h->gq); dev_release(h->unit_log2_dev);
  for (int i = 0; i < 4; ++i) if (h->timing.ev[i]) cudaEventDestroy(h->timing.ev[i]);
  for (cudaEvent_t e : h->timing.kev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h;
}

extern "C" int loosb200_frames_info(const loosb200_frames* h, size_t* F, size_t* N, int* device) {
  if (!h) return LOOSB200_ERR_INVALID;
  if (F) *F = h->F;
  if (N) *N = h->N;
  if (device) *device = h->device;
  return LOOSB200_OK;
}

extern "C" int loosb200_get_centroids(const loosb200_frames* h, double* out_F3) {
  if (!h || !out_F3) return LOOSB200_ERR_INVALID;
  LB_CUDA(cudaSetDevice(h->device));
  LB_CUDA(cudaMemcpy(out_F3, h->centroid, sizeof(double) * 3 * h->F, cudaMemcpyDeviceToHost));
  return LOOSB200_OK;
}

extern "C" int loosb200_get_unit_log2(loosb200_frames* h, int* e) {
  if (!h || !e) return LOOSB200_ERR_INVALID;
  int rc = ensure_quantised(h, INT_MIN);
  if (rc) return rc;
  *e = h->unit_log2;
  return LOOSB200_OK;
}
