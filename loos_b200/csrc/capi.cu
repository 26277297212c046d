// C-ABI glue: argument checking, tile-space partitioning, band pipelining of the
// result copy, error reporting.  See include/loosb200.h for the contract.
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "common.cuh"

namespace loosb200 {

static thread_local std::string g_last_error;

namespace {
struct DevCache {
  std::mutex mu;
  std::map<std::pair<int, size_t>, std::vector<void*>> free_blocks;  // (device, bytes) -> blocks
  std::map<void*, std::pair<int, size_t>> live;                     // block -> (device, bytes)
  size_t cached_bytes = 0;
};
DevCache& cache() { static DevCache c; return c; }
constexpr size_t CACHE_MAX_PER_KEY = 64;               // blocks kept per (device, size)
// Never sit on more than this (all devices together).  LOOSB200_CACHE_GB overrides the 96 GB default --
// 0 turns the cache off -- for processes that share their GPUs with another allocator (a framework's
// caching allocator, NCCL): those cannot see what is parked here, so either cap it or call
// loosb200_trim() before handing the device over.  Blocks are only ever released after the stream
// that used them has been synchronised (every call site does so before dev_release), which is what
// makes a cached block safe to hand to a handle on another stream.
size_t cache_max_total() {
  static const size_t v = [] {
    const char* e = getenv("LOOSB200_CACHE_GB");
    const double gb = e ? atof(e) : 96.0;
    return gb <= 0.0 ? (size_t)0 : (size_t)(gb * 1073741824.0);
  }();
  return v;
}
}  // namespace

cudaError_t dev_alloc(void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 1;
  int dev = 0;
  cudaGetDevice(&dev);
  DevCache& c = cache();
  {
    std::lock_guard<std::mutex> g(c.mu);
    auto it = c.free_blocks.find(std::make_pair(dev, bytes));
    if (it != c.free_blocks.end() && !it->second.empty()) {
      *p = it->second.back();
      it->second.pop_back();
      c.cached_bytes -= bytes;
      c.live[*p] = std::make_pair(dev, bytes);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {  // give the cached blocks back and retry once
    cudaGetLastError();
    loosb200_trim();
    e = cudaMalloc(p, bytes);
  }
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> g(c.mu);
    c.live[*p] = std::make_pair(dev, bytes);
  }
  return e;
}

void dev_release(void* p) {
  if (!p) return;
  DevCache& c = cache();
  std::pair<int, size_t> info(-1, 0);
  {
    std::lock_guard<std::mutex> g(c.mu);
    auto it = c.live.find(p);
    if (it != c.live.end()) { info = it->second; c.live.erase(it); }
    if (info.first >= 0 && c.cached_bytes + info.second <= cache_max_total() &&
        c.free_blocks[info].size() < CACHE_MAX_PER_KEY) {
      c.free_blocks[info].push_back(p);
      c.cached_bytes += info.second;
      return;
    }
  }
  cudaFree(p);
}

// Pinned host blocks (the stager's bounce buffers) go through the same cache under the
// pseudo-device -1: pinning 2 x 64 MiB costs ~100 ms per call otherwise.
cudaError_t host_alloc(void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 1;
  DevCache& c = cache();
  {
    std::lock_guard<std::mutex> g(c.mu);
    auto it = c.free_blocks.find(std::make_pair(-1, bytes));
    if (it != c.free_blocks.end() && !it->second.empty()) {
      *p = it->second.back();
      it->second.pop_back();
      c.live[*p] = std::make_pair(-1, bytes);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMallocHost(p, bytes);
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> g(c.mu);
    c.live[*p] = std::make_pair(-1, bytes);
  }
  return e;
}

void host_release(void* p) {
  if (!p) return;
  DevCache& c = cache();
  {
    std::lock_guard<std::mutex> g(c.mu);
    auto it = c.live.find(p);
    if (it != c.live.end()) {
      const std::pair<int, size_t> info = it->second;
      c.live.erase(it);
      if (c.free_blocks[info].size() < 4) { c.free_blocks[info].push_back(p); return; }
    }
  }
  cudaFreeHost(p);
}

void set_last_error(const std::string& s) { g_last_error = s; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  g_last_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " in " + what +
                 " (" + file + ":" + std::to_string(line) + ")";
  cudaGetLastError();
  return e == cudaErrorMemoryAllocation ? LOOSB200_ERR_NOMEM : LOOSB200_ERR_CUDA;
}

// Static split of the row tiles of the lower triangle over `nparts` parts.  Row
// tile t costs t+1 tiles, so tiles are dealt from the heaviest down in
// boustrophedon order (0..P-1, P-1..0, ...): every part gets the same number of
// row tiles (+-1) and the same number of tiles to within one row tile.
static std::vector<uint32_t> part_row_tiles(size_t F, int part, int nparts) {
  const uint32_t T = (uint32_t)((F + TILE_FRAMES - 1) / TILE_FRAMES);
  std::vector<uint32_t> rows;
  for (uint32_t k = 0; k < T; ++k) {
    const uint32_t t = T - 1 - k;
    const uint32_t round = k / nparts, pos = k % nparts;
    const uint32_t owner = (round & 1) ? (nparts - 1 - pos) : pos;
    if ((int)owner == part) rows.push_back(t);
  }
  std::sort(rows.begin(), rows.end());
  return rows;
}

struct Bands {
  std::vector<uint32_t> rows;    // all row tiles of the job, ascending
  std::vector<uint32_t> prefix;  // per band: local prefix arrays, concatenated
  struct Band { uint32_t r0, r1, p0; uint64_t ntiles; };  // rows[r0..r1), prefix offset
  std::vector<Band> bands;       // in launch order (heaviest rows first)
};

// Cut `rows` (ascending) into <= maxbands contiguous bands of about equal tile
// count, ordered from the last rows to the first so that, in symmetric mode, the
// columns of a band are complete as soon as the band has been computed.
// equal_rows: bands of equal HEIGHT instead (column-block delivery, see run_pairs).
static Bands make_bands(const std::vector<uint32_t>& rows, bool self, uint32_t col_tiles, int maxbands, bool equal_rows = false) {
  Bands B;
  B.rows = rows;
  uint64_t total = 0;
  for (uint32_t t : rows) total += self ? (uint64_t)t + 1 : col_tiles;
  int nb = (int)std::min<uint64_t>((uint64_t)std::max(1, maxbands), std::max<uint64_t>(1, rows.size()));
  const uint64_t per = (total + nb - 1) / nb;
  const uint32_t rows_per = (uint32_t)((rows.size() + nb - 1) / nb);
  uint32_t r1 = (uint32_t)rows.size();
  while (r1 > 0) {
    uint64_t acc = 0;
    uint32_t r0 = r1;
    while (r0 > 0 && (equal_rows ? r1 - r0 < rows_per : (acc < per || B.bands.size() + 1 == (size_t)nb))) {
      --r0;
      acc += self ? (uint64_t)rows[r0] + 1 : col_tiles;
    }
    Bands::Band b;
    b.r0 = r0; b.r1 = r1; b.p0 = (uint32_t)B.prefix.size(); b.ntiles = acc;
    uint32_t run = 0;
    for (uint32_t r = r0; r < r1; ++r) {
      B.prefix.push_back(run);
      run += self ? rows[r] + 1 : col_tiles;
    }
    B.prefix.push_back(run);
    B.bands.push_back(b);
    r1 = r0;
  }
  return B;
}

static int pick_engine(int engine, const loosb200_frames* a) {
  if (engine == LOOSB200_ENGINE_FP64) return LOOSB200_ENGINE_FP64;
  const bool can = tensor_path_available() && a->N <= TENSOR_MAX_ATOMS;
  if (engine == LOOSB200_ENGINE_TENSOR) return can ? LOOSB200_ENGINE_TENSOR : LOOSB200_ERR_UNSUPPORTED;
  if (engine == LOOSB200_ENGINE_AUTO) return can ? LOOSB200_ENGINE_TENSOR : LOOSB200_ENGINE_FP64;
  return LOOSB200_ERR_INVALID;
}

// Equal-work cut of the row tiles [0, T) into `nparts` CONTIGUOUS ranges: row tile t of the self
// matrix costs (t + 1) tiles, of a rectangle a constant.  Contiguous ranges (instead of the
// interleaved deal above) make every part's share a few wide rectangles of the matrix, which is
// what the copy engines want when the parts are gathered into one matrix over PCIe or NVLink.
// `weights` (optional, one per part) makes the shares proportional to them instead of equal: the
// host-output path sizes a device's share by what its PCIe link delivers (multi.cu).
std::vector<uint32_t> contiguous_cuts(size_t F, int nparts, bool self, const double* weights) {
  const uint32_t T = (uint32_t)((F + TILE_FRAMES - 1) / TILE_FRAMES);
  std::vector<uint32_t> cuts(nparts + 1, T);
  cuts[0] = 0;
  const double total = self ? 0.5 * (double)T * (T + 1) : (double)T;
  double wsum = 0.0;
  for (int p = 0; p < nparts; ++p) wsum += weights ? std::max(weights[p], 0.0) : 1.0;
  if (!(wsum > 0.0)) { weights = nullptr; wsum = nparts; }
  uint32_t t = 0;
  double wacc = 0.0;
  for (int p = 1; p < nparts; ++p) {
    wacc += weights ? std::max(weights[p - 1], 0.0) : 1.0;
    const double want = total * wacc / wsum;
    while (t < T && (self ? 0.5 * (double)t * (t + 1) : (double)t) < want) ++t;
    cuts[p] = t;
  }
  return cuts;
}

// The one driver behind every all-to-all entry point.
//   okind None    stats only, the matrix never leaves the SMs
//         Device  `out` is the matrix (or the packed rows) in memory the kernel can store to: the
//                 handle's own HBM, or a peer's (NVLink stores) -- written directly, one launch
//         Copy    `out` is the matrix somewhere the copy engines reach (pinned or pageable host memory,
//                 another GPU's HBM): the job runs as up to 16 bands, heaviest rows first, every band
//                 into its own compact strips, which are copied out behind the next band's kernel
int run_pairs(loosb200_frames* a, loosb200_frames* b, bool self, bool symmetric, bool packed,
              const std::vector<uint32_t>& rows, float* out, OutKind okind, size_t ld,
              loosb200_stats* stats, int engine, bool canon) {
  NvtxRange nvtx_call(self ? "loosb200 all-to-all (self)" : "loosb200 all-to-all (cross)");
  LB_CUDA(cudaSetDevice(a->device));
  const bool engine_auto = engine == LOOSB200_ENGINE_AUTO;
  engine = pick_engine(engine, a);
  if (engine < 0) return engine;
  if (engine == LOOSB200_ENGINE_TENSOR) {
    int rc;
    if (self) {
      rc = ensure_quantised(a, INT_MIN);
    } else {
      // both sets must share one fixed-point unit: quantise each, then re-quantise
      // the finer one on the coarser grid
      if (!(rc = ensure_quantised(a, INT_MIN)) && !(rc = ensure_quantised(b, INT_MIN))) {
        const int e = std::max(a->unit_log2, b->unit_log2);
        if (!(rc = ensure_quantised(a, e))) rc = ensure_quantised(b, e);
      }
    }
    // NaN / Inf coordinates have no fixed-point image; the fp64 engine propagates them like the reference
    if (rc == LOOSB200_ERR_NUMERICAL && engine_auto) engine = LOOSB200_ENGINE_FP64;
    else if (rc) return rc;
  }
  cudaStream_t st = a->stream;
  const size_t FA = a->F, FB = b->F;
  const uint32_t col_tiles = (uint32_t)((FB + TILE_FRAMES - 1) / TILE_FRAMES);
  const bool copy = okind == OutKind::Copy;
  const bool strips = copy && !packed;  // compact per-band strips (rows must be contiguous row tiles)
  if (strips)
    for (size_t r = 1; r < rows.size(); ++r)
      if (rows[r] != rows[r - 1] + 1) return LOOSB200_ERR_INVALID;
  const size_t out_rows = packed ? rows.size() * TILE_FRAMES : 0;  // packed: rows x ld (row-major)

  // Column-block delivery: the WHOLE symmetric matrix from one device.  Bands run from the last rows up, so
  // once the band of rows [A, B) is done the COLUMNS [A, B) of the symmetric matrix are final (rows below B
  // came from the earlier bands' direct stores, rows above A from this band's mirror stores): the kernels
  // write into a full F x F bounce matrix and every band ships its columns as ONE contiguous block of
  // (B - A) x F floats -- the link then runs at its plain-copy rate, which the 2-D strip copies (runs of
  // B - A floats) do not reach (C4, same box: call 785 -> 714 ms = 56 GB/s, profiles/bench/r2_e2e_column_blocks.txt).  Bands of equal height, so that the copies,
  // not the shrinking kernels, pace the pipeline; only the first band's kernel is exposed.
  const bool colmode = strips && self && symmetric && out && !rows.empty() && rows[0] == 0 &&
                       rows.size() == (FA + TILE_FRAMES - 1) / TILE_FRAMES &&
                       !(getenv("LOOSB200_STRIPS") && atoi(getenv("LOOSB200_STRIPS")));
  int maxbands = copy ? (colmode ? 64 : 16) : 1;
  if (copy) if (const char* e = getenv("LOOSB200_BANDS")) maxbands = std::min(256, std::max(1, atoi(e)));
  Bands B = make_bands(rows, self, col_tiles, maxbands, colmode);
  // band k covers matrix rows [bA[k], bB[k]); strip mode: its row strip S_k and column strip T_k
  std::vector<size_t> bA(B.bands.size()), bB(B.bands.size()), offS(B.bands.size()), offT(B.bands.size());
  size_t strip_floats = 0;
  for (size_t k = 0; k < B.bands.size(); ++k) {
    const Bands::Band& bd = B.bands[k];
    bA[k] = (size_t)B.rows[bd.r0] * TILE_FRAMES;
    bB[k] = std::min(FA, ((size_t)B.rows[bd.r1 - 1] + 1) * TILE_FRAMES);
    if (bB[k] < bA[k]) bB[k] = bA[k];
    const size_t h = bB[k] - bA[k];
    offS[k] = strip_floats; strip_floats += h * (self ? bB[k] : FB);
    offT[k] = strip_floats; strip_floats += (self && symmetric) ? h * bA[k] : 0;
  }

  uint32_t *d_rows = nullptr, *d_prefix = nullptr;
  float* d_out = nullptr;
  double* d_stats = nullptr;
  // the strips of consecutive bands leave on alternating streams: two copies in flight keep the link
  // busy across the gaps between 2-D transfers (LOOSB200_COPY_STREAMS=1: one stream)
  cudaStream_t copy_streams[2] = {nullptr, nullptr};
  const int ncopy = (getenv("LOOSB200_COPY_STREAMS") && atoi(getenv("LOOSB200_COPY_STREAMS")) == 1) ? 1 : 2;
  std::vector<void*> scratch;  // per-launch device scratch of the engines, released after the final sync
  std::vector<cudaEvent_t> evs(B.bands.size(), nullptr);
  auto cleanup = [&]() {
    dev_release(d_rows); dev_release(d_prefix); dev_release(d_stats);
    for (void* p : scratch) dev_release(p);
    if (copy) dev_release(d_out);
    for (cudaStream_t cs : copy_streams) if (cs) cudaStreamDestroy(cs);
    for (auto e : evs) if (e) cudaEventDestroy(e);
  };
  bool ok = dev_alloc_t(&d_rows, sizeof(uint32_t) * std::max<size_t>(1, B.rows.size())) == cudaSuccess &&
            dev_alloc_t(&d_prefix, sizeof(uint32_t) * std::max<size_t>(1, B.prefix.size())) == cudaSuccess &&
            dev_alloc_t(&d_stats, 2 * sizeof(double)) == cudaSuccess;
  if (ok && copy) {
    const size_t n = packed ? out_rows * FB : (colmode ? FA * FA : strip_floats);
    ok = dev_alloc_t(&d_out, sizeof(float) * std::max<size_t>(1, n)) == cudaSuccess;
    for (int c = 0; c < ncopy && ok; ++c) ok = cudaStreamCreateWithFlags(&copy_streams[c], cudaStreamNonBlocking) == cudaSuccess;
    for (auto& e : evs) if (ok) ok = cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
  } else if (okind == OutKind::Device) {
    d_out = out;
  }
  if (!ok) { cudaGetLastError(); cleanup(); return LOOSB200_ERR_NOMEM; }

  cudaEventRecord(a->timing.ev[0], st);
  cudaMemcpyAsync(d_rows, B.rows.data(), sizeof(uint32_t) * B.rows.size(), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_prefix, B.prefix.data(), sizeof(uint32_t) * B.prefix.size(), cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(d_stats, 0, 2 * sizeof(double), st);
  cudaEventRecord(a->timing.ev[1], st);
  int launches = 0, rc = LOOSB200_OK;
  for (size_t k = 0; k < B.bands.size() && rc == LOOSB200_OK; ++k) {
    const Bands::Band& bd = B.bands[k];
    PairJob job;
    job.a = a; job.b = b; job.self = self; job.symmetric = symmetric;
    job.rows_host = B.rows.data() + bd.r0;
    job.row_tiles_dev = d_rows + bd.r0;
    job.tile_prefix_dev = d_prefix + bd.p0;
    job.n_row_tiles = bd.r1 - bd.r0;
    job.n_tiles = bd.ntiles;
    job.packed_rows = packed;
    job.canon = canon;
    if (!d_out) {
      job.out = nullptr; job.ld = ld;
    } else if (colmode) {
      job.out = d_out; job.ld = FA;  // the plain F x F matrix, direct and mirror stores
    } else if (strips) {
      job.out = d_out + offS[k]; job.ld = bB[k] - bA[k]; job.row_off = bA[k];
      job.out_t = d_out + offT[k]; job.ld_t = bA[k];
    } else if (packed) {
      job.ld = copy ? FB : ld;
      job.out = d_out + (size_t)bd.r0 * TILE_FRAMES * job.ld;
    } else {
      job.out = d_out; job.ld = ld;
    }
    job.stats_dev = stats ? d_stats : nullptr;
    job.scratch = &scratch;
    while (a->timing.kev.size() < 2 * (k + 1)) { cudaEvent_t e = nullptr; cudaEventCreate(&e); a->timing.kev.push_back(e); }
    job.ev_k0 = a->timing.kev[2 * k]; job.ev_k1 = a->timing.kev[2 * k + 1];
    a->timing.kev_used = (int)(2 * (k + 1));
    if (engine == LOOSB200_ENGINE_TENSOR) {
      int l = 0;
      rc = launch_pairs_tensor(job, st, &l);
      launches += l;
    } else {
      rc = launch_pairs_fp64(job, st);
      launches += 1;
    }
    if (rc == LOOSB200_OK && copy) {
      cudaStream_t copy_stream = copy_streams[k % ncopy];
      cudaEventRecord(evs[k], st);
      cudaStreamWaitEvent(copy_stream, evs[k], 0);
      if (packed) {
        // rows of the band are contiguous in the packed buffer; only columns j < i exist
        const size_t r_lo = (size_t)bd.r0 * TILE_FRAMES;
        size_t nr = (size_t)(bd.r1 - bd.r0) * TILE_FRAMES;
        // canon: the packed rows of a rectangle over one set ARE the columns of its symmetric matrix; the
        // caller's matrix has exactly FA of them (the packed buffers of the part calls are padded to tiles)
        if (canon) nr = std::min(nr, FA - std::min(FA, (size_t)B.rows[bd.r0] * TILE_FRAMES));
        const size_t width = self ? std::min(FB, ((size_t)B.rows[bd.r1 - 1] + 1) * TILE_FRAMES) : FB;
        cudaMemcpy2DAsync(out + r_lo * ld, ld * sizeof(float), d_out + r_lo * FB, FB * sizeof(float),
                          width * sizeof(float), nr, cudaMemcpyDefault, copy_stream);
      } else if (colmode) {
        if (bB[k] > bA[k])
          cudaMemcpy2DAsync(out + bA[k] * ld, ld * sizeof(float), d_out + bA[k] * FA, FA * sizeof(float),
                            FA * sizeof(float), bB[k] - bA[k], cudaMemcpyDefault, copy_stream);
      } else if (bB[k] > bA[k]) {
        // Everything that involves the band's rows [A, B) is final once the band is done: the row
        // strip rows [A,B) x cols [0,B) (all columns in rectangle mode) and, for the symmetric matrix,
        // its mirror cols [A,B) x rows [0,A).  Both go out now, so the copy volume of a band is
        // proportional to its work and only the last (lightest) band's copy is exposed.
        const size_t A = bA[k], Bk = bB[k], h = Bk - A;
        cudaMemcpy2DAsync(out + A, ld * sizeof(float), d_out + offS[k], h * sizeof(float),
                          h * sizeof(float), self ? Bk : FB, cudaMemcpyDefault, copy_stream);
        if (self && symmetric && A > 0)
          cudaMemcpy2DAsync(out + A * ld, ld * sizeof(float), d_out + offT[k], A * sizeof(float),
                            A * sizeof(float), h, cudaMemcpyDefault, copy_stream);
      }
    }
  }
  cudaEventRecord(a->timing.ev[2], st);
  double hs[2] = {0.0, 0.0};
  if (stats) cudaMemcpyAsync(hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, st);
  cudaEventRecord(a->timing.ev[3], st);
  NvtxRange nvtx_wait("loosb200 contract + copy (wait)");
  cudaError_t e = cudaStreamSynchronize(st);
  for (cudaStream_t cs : copy_streams) if (e == cudaSuccess && cs) e = cudaStreamSynchronize(cs);
  if (e == cudaSuccess) e = cudaGetLastError();
  cleanup();
  if (rc != LOOSB200_OK) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "all-to-all kernel", __FILE__, __LINE__);
  if (stats) {
    stats->sum = hs[0];
    stats->max = hs[1];
    uint64_t cnt = 0;
    if (self) {
      for (uint32_t t : rows) {
        const uint64_t i0 = (uint64_t)t * TILE_FRAMES, i1 = std::min<uint64_t>(FA, i0 + TILE_FRAMES);
        cnt += (i1 * (i1 - 1) - i0 * (i0 ? i0 - 1 : 0)) / 2;
      }
    } else {
      for (uint32_t t : rows) {
        const uint64_t i0 = (uint64_t)t * TILE_FRAMES, i1 = std::min<uint64_t>(FA, i0 + TILE_FRAMES);
        if (i1 <= i0) continue;
        if (canon) cnt += (i1 * (i1 - 1) - i0 * (i0 ? i0 - 1 : 0)) / 2;  // only j < i counts, see PairJob::canon
        else cnt += (i1 - i0) * (uint64_t)FB;
      }
    }
    stats->count = cnt;
  }
  a->timing.launches = launches; a->timing.valid = true;
  return LOOSB200_OK;
}

std::vector<uint32_t> row_tile_range(uint32_t t0, uint32_t t1) {
  std::vector<uint32_t> r;
  for (uint32_t t = t0; t < t1; ++t) r.push_back(t);
  return r;
}

}  // namespace loosb200

using namespace loosb200;

extern "C" void loosb200_trim(void) {
  DevCache& c = cache();
  std::lock_guard<std::mutex> g(c.mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& kv : c.free_blocks) {
    if (kv.first.first < 0) {
      for (void* p : kv.second) cudaFreeHost(p);
    } else {
      cudaSetDevice(kv.first.first);
      for (void* p : kv.second) cudaFree(p);
    }
    kv.second.clear();
  }
  c.cached_bytes = 0;
  cudaSetDevice(cur);
  cudaGetLastError();
}

extern "C" const char* loosb200_version(void) { return "loosb200 0.1 (sm_100a)"; }

extern "C" const char* loosb200_last_error(void) { return g_last_error.c_str(); }

extern "C" const char* loosb200_strerror(int code) {
  switch (code) {
    case LOOSB200_OK: return "success";
    case LOOSB200_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (there is no CPU fallback)";
    case LOOSB200_ERR_CUDA: return "CUDA call failed";
    case LOOSB200_ERR_INVALID: return "invalid argument";
    case LOOSB200_ERR_SIZE_MISMATCH: return "groups/frames of differing sizes";
    case LOOSB200_ERR_RANGE: return "index or coordinate out of range";
    case LOOSB200_ERR_NOMEM: return "out of memory";
    case LOOSB200_ERR_NUMERICAL: return "numerical failure in the superposition solver";
    case LOOSB200_ERR_UNSUPPORTED: return "request not supported by this build";
    default: return "unknown error";
  }
}

extern "C" int loosb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int usable = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10)
      usable = d + 1;  // devices are addressed by CUDA ordinal
  }
  cudaGetLastError();
  return usable;
}

static int check_self(loosb200_frames* h, float* out, size_t ld) {
  if (!h) return LOOSB200_ERR_INVALID;
  if (out && ld < h->F) return LOOSB200_ERR_INVALID;
  return LOOSB200_OK;
}

static std::vector<uint32_t> all_rows(size_t F) { return part_row_tiles(F, 0, 1); }

extern "C" int loosb200_rmsds_self(loosb200_frames* h, float* out, size_t ld, loosb200_stats* stats, int engine) {
  int rc = check_self(h, out, ld);
  if (rc) return rc;
  return run_pairs(h, h, true, true, false, all_rows(h->F), out, out ? OutKind::Copy : OutKind::None, ld, stats, engine);
}

extern "C" int loosb200_rmsds_self_device(loosb200_frames* h, float* out_dev, size_t ld, loosb200_stats* stats, int engine) {
  int rc = check_self(h, out_dev, ld);
  if (rc) return rc;
  return run_pairs(h, h, true, true, false, all_rows(h->F), out_dev, out_dev ? OutKind::Device : OutKind::None, ld, stats, engine);
}

static int check_cross(loosb200_frames* a, loosb200_frames* b, float* out, size_t ld) {
  if (!a || !b) return LOOSB200_ERR_INVALID;
  if (a->N != b->N) return LOOSB200_ERR_SIZE_MISMATCH;
  if (a->device != b->device) return LOOSB200_ERR_INVALID;
  if (out && ld < a->F) return LOOSB200_ERR_INVALID;
  return LOOSB200_OK;
}

extern "C" int loosb200_rmsds_cross(loosb200_frames* a, loosb200_frames* b, float* out, size_t ld, loosb200_stats* stats, int engine) {
  int rc = check_cross(a, b, out, ld);
  if (rc) return rc;
  return run_pairs(a, b, false, false, false, all_rows(a->F), out, out ? OutKind::Copy : OutKind::None, ld, stats, engine);
}

extern "C" int loosb200_rmsds_cross_device(loosb200_frames* a, loosb200_frames* b, float* out_dev, size_t ld, loosb200_stats* stats, int engine) {
  int rc = check_cross(a, b, out_dev, ld);
  if (rc) return rc;
  return run_pairs(a, b, false, false, false, all_rows(a->F), out_dev, out_dev ? OutKind::Device : OutKind::None, ld, stats, engine);
}

// matrix in HBM -> text through the sink (text.cu)
static int pairs_text(loosb200_frames* a, loosb200_frames* b, bool self, unsigned precision, loosb200_sink sink,
                      void* user, loosb200_stats* stats, int engine) {
  if (!sink) return LOOSB200_ERR_INVALID;
  LB_CUDA(cudaSetDevice(a->device));
  const size_t m = a->F, n = b->F;
  // The whole float matrix in HBM when it fits (one pass over the triangle, rows read back by symmetry);
  // otherwise bands of rows computed as rectangles against every column -- memory is then bounded by
  // the band, at twice the arithmetic of the triangle plus one stats-only pass.
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const size_t full_bytes = sizeof(float) * std::max<size_t>(1, m * n) * (self ? 1 : 2);  // + transposed copy
  size_t band_rows = 0;
  if (const char* env = getenv("LOOSB200_TEXT_BAND_ROWS")) band_rows = (size_t)std::max(0, atoi(env));
  else if (full_bytes + ((size_t)2 << 30) > free_b) band_rows = std::max<size_t>(TILE_FRAMES, ((size_t)8 << 30) / (sizeof(float) * n));
  if (band_rows == 0) {
    float* d_M = nullptr;
    if (dev_alloc_t(&d_M, sizeof(float) * std::max<size_t>(1, m * n)) != cudaSuccess) { cudaGetLastError(); return LOOSB200_ERR_NOMEM; }
    int rc = run_pairs(a, b, self, self, false, all_rows(a->F), d_M, OutKind::Device, m, stats, engine);
    if (rc == LOOSB200_OK) rc = stream_matrix_text(d_M, m, n, m, self, precision, sink, user, a->stream);
    dev_release(d_M);
    return rc;
  }
  band_rows = (band_rows + TILE_FRAMES - 1) / TILE_FRAMES * TILE_FRAMES;
  int rc = LOOSB200_OK;
  if (stats && self) rc = run_pairs(a, b, true, false, false, all_rows(a->F), nullptr, OutKind::None, m, stats, engine);
  if (stats && !self) { stats->sum = 0.0; stats->max = 0.0; stats->count = 0; }
  if (rc == LOOSB200_OK) rc = matrix_text_header(m, n, sink, user);
  float* d_band = nullptr;
  if (rc == LOOSB200_OK && dev_alloc_t(&d_band, sizeof(float) * band_rows * n) != cudaSuccess) { cudaGetLastError(); rc = LOOSB200_ERR_NOMEM; }
  for (size_t R0 = 0; R0 < m && rc == LOOSB200_OK; R0 += band_rows) {
    const size_t R1 = std::min(m, R0 + band_rows);
    std::vector<uint32_t> tiles;
    for (size_t t = R0 / TILE_FRAMES; t * TILE_FRAMES < R1; ++t) tiles.push_back((uint32_t)t);
    loosb200_stats bst;
    rc = run_pairs(a, b, false, false, true, tiles, d_band, OutKind::Device, n, (stats && !self) ? &bst : nullptr, engine, self);
    if (rc != LOOSB200_OK) break;
    if (stats && !self) { stats->sum += bst.sum; stats->max = std::max(stats->max, bst.max); stats->count += bst.count; }
    if (self) zero_band_diagonal(d_band, n, R0, (uint32_t)(R1 - R0), a->stream);
    rc = stream_rows_text(d_band, n, R1 - R0, n, precision, sink, user, a->stream);
  }
  dev_release(d_band);
  return rc;
}

// ---- the same text from several GPUs ------------------------------------------------------------
// Bands of rows (about one 256 MB batch of text each) are dealt round-robin to the devices; every
// device computes its band as a rectangle against all columns, formats it and copies the text to
// the host over its own PCIe link, and the bands are handed to the sink in row order through a
// turnstile.  Each device is at most one band ahead of the sink, so n bands are in flight.
namespace {
struct Turnstile {
  std::mutex mu;
  std::condition_variable cv;
  size_t next = 0;
  bool failed = false;
  loosb200_sink sink = nullptr;
  void* user = nullptr;
};
struct TurnUser { Turnstile* t; size_t band; };
int turnstile_sink(void* u, const char* data, size_t n) {
  TurnUser* tu = static_cast<TurnUser*>(u);
  Turnstile* t = tu->t;
  {
    std::unique_lock<std::mutex> lk(t->mu);
    t->cv.wait(lk, [&] { return t->next == tu->band || t->failed; });
    if (t->failed) return 1;
  }
  // only the band whose turn it is gets here, and the calls of one band are sequential
  const int r = t->sink(t->user, data, n);
  if (r) { std::lock_guard<std::mutex> lk(t->mu); t->failed = true; t->cv.notify_all(); }
  return r;
}
}  // namespace

static int pairs_text_multi(loosb200_frames* const* a, loosb200_frames* const* b, int ndev, bool self,
                            unsigned precision, loosb200_sink sink, void* user, loosb200_stats* stats, int engine) {
  if (!sink) return LOOSB200_ERR_INVALID;
  if (precision > 9) return LOOSB200_ERR_UNSUPPORTED;
  const size_t m = a[0]->F, n = b[0]->F;
  const unsigned prec = precision == 0 ? 1 : precision;
  const size_t worst_row = n * (size_t)(prec + 7) + 1;
  size_t band_rows = std::max<size_t>(1, ((size_t)256 << 20) / worst_row);
  if (const char* env = getenv("LOOSB200_TEXT_BAND_ROWS")) { const int v = atoi(env); if (v > 0) band_rows = (size_t)v; }
  band_rows = std::max<size_t>(TILE_FRAMES, band_rows / TILE_FRAMES * TILE_FRAMES);
  const size_t nbands = (m + band_rows - 1) / band_rows;
  int rc = LOOSB200_OK;
  // statistics: the triangle (self) split over the devices, no output; rectangles report their own
  if (stats && self) {
    std::vector<loosb200_stats> ps(ndev);
    const std::vector<uint32_t> cuts = contiguous_cuts(m, ndev, true);
    std::vector<int> rcs(ndev, LOOSB200_OK);
    std::vector<std::thread> th;
    for (int d = 0; d < ndev; ++d)
      th.emplace_back([&, d]() {
        ps[d].sum = 0.0; ps[d].max = 0.0; ps[d].count = 0;
        const std::vector<uint32_t> rows = row_tile_range(cuts[d], cuts[d + 1]);
        if (!rows.empty()) rcs[d] = run_pairs(a[d], a[d], true, false, false, rows, nullptr, OutKind::None, m, &ps[d], engine);
      });
    for (auto& t : th) t.join();
    stats->sum = 0.0; stats->max = 0.0; stats->count = 0;
    for (int d = 0; d < ndev; ++d) {
      if (rcs[d]) return rcs[d];
      stats->sum += ps[d].sum; stats->max = std::max(stats->max, ps[d].max); stats->count += ps[d].count;
    }
  }
  if ((rc = matrix_text_header(m, n, sink, user))) return rc;
  Turnstile ts;
  ts.sink = sink; ts.user = user;
  std::vector<int> rcs(ndev, LOOSB200_OK);
  std::vector<loosb200_stats> ps(ndev);
  std::vector<std::thread> th;
  for (int d = 0; d < ndev; ++d)
    th.emplace_back([&, d]() {
      ps[d].sum = 0.0; ps[d].max = 0.0; ps[d].count = 0;
      int r = LOOSB200_OK;
      float* d_band = nullptr;
      if (cudaSetDevice(a[d]->device) != cudaSuccess || dev_alloc_t(&d_band, sizeof(float) * band_rows * n) != cudaSuccess) {
        cudaGetLastError();
        r = LOOSB200_ERR_NOMEM;
      }
      for (size_t k = (size_t)d; k < nbands; k += (size_t)ndev) {
        const size_t R0 = k * band_rows, R1 = std::min(m, R0 + band_rows);
        if (r == LOOSB200_OK) {
          const std::vector<uint32_t> tiles = row_tile_range((uint32_t)(R0 / TILE_FRAMES), (uint32_t)((R1 + TILE_FRAMES - 1) / TILE_FRAMES));
          loosb200_stats bst;
          r = run_pairs(a[d], b[d], false, false, true, tiles, d_band, OutKind::Device, n, (stats && !self) ? &bst : nullptr, engine, self);
          if (r == LOOSB200_OK && stats && !self) { ps[d].sum += bst.sum; ps[d].max = std::max(ps[d].max, bst.max); ps[d].count += bst.count; }
        }
        if (r == LOOSB200_OK) {
          if (self) zero_band_diagonal(d_band, n, R0, (uint32_t)(R1 - R0), a[d]->stream);
          TurnUser tu{&ts, k};
          r = stream_rows_text(d_band, n, R1 - R0, n, precision, turnstile_sink, &tu, a[d]->stream);
        }
        std::lock_guard<std::mutex> lk(ts.mu);
        if (r != LOOSB200_OK) ts.failed = true;
        else ts.next = k + 1;
        ts.cv.notify_all();
        if (r != LOOSB200_OK) break;
      }
      dev_release(d_band);
      rcs[d] = r;
    });
  for (auto& t : th) t.join();
  for (int d = 0; d < ndev; ++d) if (rcs[d]) return rcs[d];
  if (stats && !self) {
    stats->sum = 0.0; stats->max = 0.0; stats->count = 0;
    for (int d = 0; d < ndev; ++d) { stats->sum += ps[d].sum; stats->max = std::max(stats->max, ps[d].max); stats->count += ps[d].count; }
  }
  return LOOSB200_OK;
}

static int check_text_replicas(loosb200_frames* const* r, int n) {
  if (!r || n < 1) return LOOSB200_ERR_INVALID;
  for (int d = 0; d < n; ++d) {
    if (!r[d] || r[d]->stager) return LOOSB200_ERR_INVALID;
    if (r[d]->F != r[0]->F || r[d]->N != r[0]->N) return LOOSB200_ERR_SIZE_MISMATCH;
  }
  return LOOSB200_OK;
}

extern "C" int loosb200_rmsds_self_multi_text(loosb200_frames* const* replicas, int n, unsigned precision,
                                              loosb200_sink sink, void* user, loosb200_stats* stats, int engine) {
  int rc = check_text_replicas(replicas, n);
  if (rc) return rc;
  if (n == 1) return pairs_text(replicas[0], replicas[0], true, precision, sink, user, stats, engine);
  return pairs_text_multi(replicas, replicas, n, true, precision, sink, user, stats, engine);
}

extern "C" int loosb200_rmsds_cross_multi_text(loosb200_frames* const* a, loosb200_frames* const* b, int n,
                                               unsigned precision, loosb200_sink sink, void* user,
                                               loosb200_stats* stats, int engine) {
  int rc = check_text_replicas(a, n);
  if (!rc) rc = check_text_replicas(b, n);
  if (rc) return rc;
  for (int d = 0; d < n; ++d) {
    if (a[d]->device != b[d]->device) return LOOSB200_ERR_INVALID;
    if (a[d]->N != b[d]->N) return LOOSB200_ERR_SIZE_MISMATCH;
  }
  if (n == 1) return pairs_text(a[0], b[0], false, precision, sink, user, stats, engine);
  return pairs_text_multi(a, b, n, false, precision, sink, user, stats, engine);
}

extern "C" int loosb200_rmsds_self_text(loosb200_frames* h, unsigned precision, loosb200_sink sink, void* user,
                                        loosb200_stats* stats, int engine) {
  if (!h) return LOOSB200_ERR_INVALID;
  return pairs_text(h, h, true, precision, sink, user, stats, engine);
}

extern "C" int loosb200_rmsds_cross_text(loosb200_frames* a, loosb200_frames* b, unsigned precision,
                                         loosb200_sink sink, void* user, loosb200_stats* stats, int engine) {
  int rc = check_cross(a, b, nullptr, 0);
  if (rc) return rc;
  return pairs_text(a, b, false, precision, sink, user, stats, engine);
}

extern "C" int loosb200_part_rows(size_t F, int part, int nparts, uint32_t* rows_or_null, size_t* nrows) {
  if (F == 0 || nparts < 1 || part < 0 || part >= nparts) return LOOSB200_ERR_INVALID;
  std::vector<uint32_t> rt = part_row_tiles(F, part, nparts);
  size_t n = 0;
  for (uint32_t t : rt)
    for (uint32_t k = 0; k < (uint32_t)TILE_FRAMES; ++k) {
      const size_t f = (size_t)t * TILE_FRAMES + k;
      // the packed buffer always has TILE_FRAMES rows per row tile; rows beyond F are padding
      if (rows_or_null) rows_or_null[n] = (uint32_t)(f < F ? f : 0xffffffffu);
      ++n;
    }
  if (nrows) *nrows = n;
  return LOOSB200_OK;
}

static int check_part(loosb200_frames* h, int part, int nparts, float* out, size_t ld) {
  if (!h || nparts < 1 || part < 0 || part >= nparts) return LOOSB200_ERR_INVALID;
  if (out && ld < h->F) return LOOSB200_ERR_INVALID;
  return LOOSB200_OK;
}

extern "C" int loosb200_rmsds_self_part(loosb200_frames* h, int part, int nparts, float* out, size_t ld, loosb200_stats* stats, int engine) {
  int rc = check_part(h, part, nparts, out, ld);
  if (rc) return rc;
  return run_pairs(h, h, true, false, true, part_row_tiles(h->F, part, nparts), out, out ? OutKind::Copy : OutKind::None, ld, stats, engine);
}

extern "C" int loosb200_rmsds_self_part_device(loosb200_frames* h, int part, int nparts, float* out_dev, size_t ld, loosb200_stats* stats, int engine) {
  int rc = check_part(h, part, nparts, out_dev, ld);
  if (rc) return rc;
  return run_pairs(h, h, true, false, true, part_row_tiles(h->F, part, nparts), out_dev, out_dev ? OutKind::Device : OutKind::None, ld, stats, engine);
}

extern "C" int loosb200_last_timing(const loosb200_frames* h, double* ms_main, double* ms_other) {
  if (!h) return LOOSB200_ERR_INVALID;
  if (!h->timing.valid) return LOOSB200_ERR_INVALID;
  float t_all = 0.f, t_main = 0.f;
  LB_CUDA(cudaEventElapsedTime(&t_all, h->timing.ev[0], h->timing.ev[3]));
  LB_CUDA(cudaEventElapsedTime(&t_main, h->timing.ev[1], h->timing.ev[2]));
  if (h->timing.kev_used > 0) {  // precise: events right around every contraction launch
    t_main = 0.f;
    for (int k = 0; k + 1 < h->timing.kev_used; k += 2) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, h->timing.kev[k], h->timing.kev[k + 1]) == cudaSuccess) t_main += t;
    }
    cudaGetLastError();
  }
  if (ms_main) *ms_main = t_main;
  if (ms_other) *ms_other = t_all - t_main;
  return h->timing.launches;
}
