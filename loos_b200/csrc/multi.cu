// Several GPUs in one process: the reference fills ONE matrix from np worker threads
// (Master / Threader, Tools/rmsds.cpp:219-285,387-419; multi-rmsds.cpp:395-417); here the workers are
// the GPUs of one box.  Frame pairs are independent, so there is no exchange step while computing:
// every device holds a replica of the staged frames (loosb200_replicate: one H2D upload, then
// NVLink copies), computes a CONTIGUOUS equal-work range of row tiles of the lower triangle, and
// delivers its part of the matrix itself:
//   host output     every device copies its row strips and their mirror images straight into the
//                   caller's matrix over ITS OWN PCIe link, band by band behind its kernel;
//   device output   the matrix is gathered in the HBM of replicas[0]'s device over NVLink: either the
//                   other devices' kernels store their tiles there directly (peer stores, the
//                   transfer rides under the contraction tile by tile), or they fill local strips
//                   that the copy engines push band by band (LOOSB200_GATHER=copy).
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"

namespace loosb200 {

// peer access src -> dst (kernels and copy engines of `dev` may touch memory of `peer`)
static bool enable_peer(int dev, int peer) {
  if (dev == peer) return true;
  int can = 0;
  if (cudaDeviceCanAccessPeer(&can, dev, peer) != cudaSuccess || !can) { cudaGetLastError(); return false; }
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(dev);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  cudaSetDevice(cur);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return true; }
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return true;
}

static int check_replicas(loosb200_frames* const* r, int n) {
  if (!r || n < 1) return LOOSB200_ERR_INVALID;
  for (int d = 0; d < n; ++d) {
    if (!r[d] || r[d]->stager) return LOOSB200_ERR_INVALID;
    if (r[d]->F != r[0]->F || r[d]->N != r[0]->N) return LOOSB200_ERR_SIZE_MISMATCH;
    // (two replicas on one device are allowed: wasteful, but it lets a one-GPU box run this path)
  }
  return LOOSB200_OK;
}

// Run fn(d) for d = 0..n-1 on n host threads (each binds its device inside run_pairs); first error wins.
template <class Fn>
static int on_all_devices(int n, Fn fn) {
  std::vector<int> rc(n, LOOSB200_OK);
  std::vector<std::string> msg(n);
  std::vector<std::thread> th;
  for (int d = 1; d < n; ++d)
    th.emplace_back([&, d]() { rc[d] = fn(d); if (rc[d]) msg[d] = loosb200_last_error(); });
  rc[0] = fn(0);
  for (auto& t : th) t.join();
  for (int d = 0; d < n; ++d)
    if (rc[d]) { if (d) set_last_error(msg[d]); return rc[d]; }
  return LOOSB200_OK;
}

static void merge_stats(loosb200_stats* total, const std::vector<loosb200_stats>& parts) {
  if (!total) return;
  total->sum = 0.0; total->max = 0.0; total->count = 0;
  for (const loosb200_stats& s : parts) {
    total->sum += s.sum;
    total->max = std::max(total->max, s.max);
    total->count += s.count;
  }
}

// Device-to-host rate of every device's PCIe link with ALL of them copying at once, measured once per
// process and device list (2 x 64 MiB per device into pinned memory, ~20 ms).  On the 8-GPU boxes this was
// developed on the links are far from equal (10.8 ... 22.6 GB/s per device with eight copying, 122 GB/s
// in total: profiles/bench/r2_pinned_d2h_8gpu.jsonl), and with equal shares of the matrix the slowest
// link sets the time of the host-output call.  Returns false when anything fails (equal shares then).
static bool link_rates(loosb200_frames* const* a, int n, std::vector<double>& gbs) {
  static std::mutex mu;
  static std::vector<int> cached_devs;
  static std::vector<double> cached;
  std::vector<int> devs(n);
  for (int d = 0; d < n; ++d) devs[d] = a[d]->device;
  std::lock_guard<std::mutex> lk(mu);
  if (devs == cached_devs) { gbs = cached; return true; }
  NvtxRange nvtx("loosb200 probe of the devices' PCIe links");
  constexpr size_t BYTES = (size_t)64 << 20;
  std::vector<void*> dbuf(n, nullptr), hbuf(n, nullptr);
  std::vector<double> rate(n, 0.0);
  std::atomic<int> ready(0), failed(0);
  std::vector<std::thread> th;
  for (int d = 0; d < n; ++d)
    th.emplace_back([&, d]() {
      bool ok = cudaSetDevice(devs[d]) == cudaSuccess && dev_alloc(&dbuf[d], BYTES) == cudaSuccess &&
                host_alloc(&hbuf[d], BYTES) == cudaSuccess;
      cudaStream_t st = nullptr;
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      ok = ok && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess;
      if (ok) ok = cudaMemcpyAsync(hbuf[d], dbuf[d], BYTES / 8, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                   cudaStreamSynchronize(st) == cudaSuccess;  // warm-up
      if (!ok) failed.fetch_add(1);
      ready.fetch_add(1);
      while (ready.load() < n) std::this_thread::yield();  // all links start together
      if (ok && failed.load() == 0) {
        cudaEventRecord(e0, st);
        for (int r = 0; r < 2; ++r) cudaMemcpyAsync(hbuf[d], dbuf[d], BYTES, cudaMemcpyDeviceToHost, st);
        cudaEventRecord(e1, st);
        float ms = 0.f;
        if (cudaStreamSynchronize(st) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess && ms > 0.f)
          rate[d] = 2.0 * (double)BYTES / (ms * 1e-3) / 1e9;
        else failed.fetch_add(1);
      }
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
      if (st) cudaStreamDestroy(st);
      if (dbuf[d]) dev_release(dbuf[d]);
      if (hbuf[d]) host_release(hbuf[d]);
      cudaGetLastError();
    });
  for (auto& t : th) t.join();
  if (failed.load()) return false;
  cached_devs = devs; cached = rate; gbs = rate;
  if (getenv("LOOSB200_VERBOSE")) {
    fprintf(stderr, "[loosb200] device-to-host GB/s per device, all copying:");
    for (double r : rate) fprintf(stderr, " %.1f", r);
    fprintf(stderr, "\n");
  }
  return true;
}

// All parts of one all-to-all job.  kind Copy: out is the caller's matrix (host memory, or the HBM of
// any device): every device delivers its strips there.  kind Device: out lives in the HBM of
// a[0]'s device; that device writes its share directly, the others by peer stores or strip copies.
static int pairs_multi(loosb200_frames* const* a, loosb200_frames* const* b, int n, bool self, float* out,
                       OutKind kind, size_t ld, loosb200_stats* stats, int engine) {
  const size_t F = a[0]->F;
  // Shares of the tile rows: equal work, except when every device ships its part to HOST memory over its
  // own PCIe link -- then a share costs max(compute, copy) per pair, and the copy term differs by device.
  // (Tried and dropped: every device computing the mirror-canonical RECTANGLE of its rows against all
  // columns and shipping finished, contiguous column blocks of the symmetric matrix -- twice the
  // arithmetic, no short runs in the copies.  What limits eight devices copying at once is the host
  // side of the links, not the run length: 405 ms against 383 ms for the strips on 8 GPUs, 637 against
  // 624 ms on 2, profiles/bench/r2_e2e_modes_{2,8}gpu.txt.)
  std::vector<double> weights;
  if (kind == OutKind::Copy && n > 1 && out && !(getenv("LOOSB200_EQUAL_SHARES") && atoi(getenv("LOOSB200_EQUAL_SHARES")))) {
    cudaPointerAttributes pa;
    const bool host = cudaPointerGetAttributes(&pa, out) != cudaSuccess || pa.type == cudaMemoryTypeHost ||
                      pa.type == cudaMemoryTypeUnregistered;
    cudaGetLastError();
    std::vector<double> gbs;
    if (host && link_rates(a, n, gbs)) {
      // s per pair on one B200: tensor engine 1.25e10 pairs/s at N = 1000, fp64 engine 9.5 TFLOP/s of 18 N flop
      const bool fp64 = engine == LOOSB200_ENGINE_FP64 || a[0]->N > TENSOR_MAX_ATOMS;
      const double compute = (fp64 ? 1.9e-9 : 8e-11) * (double)std::max<size_t>(a[0]->N, 256) / 1000.0;
      const double bytes = self ? 8.0 : 4.0;  // entry + mirror image
      for (int d = 0; d < n; ++d) weights.push_back(1.0 / std::max(compute, bytes / (std::max(gbs[d], 0.1) * 1e9)));
    }
  }
  const std::vector<uint32_t> cuts = contiguous_cuts(F, n, self, weights.empty() ? nullptr : weights.data());
  bool peer_stores = false;
  if (kind == OutKind::Device) {
    const char* g = getenv("LOOSB200_GATHER");
    peer_stores = !(g && strcmp(g, "copy") == 0);
    for (int d = 1; d < n; ++d)
      if (!enable_peer(a[d]->device, a[0]->device)) return LOOSB200_ERR_UNSUPPORTED;
  }
  std::vector<loosb200_stats> ps(n);
  int rc = on_all_devices(n, [&](int d) -> int {
    const std::vector<uint32_t> rows = row_tile_range(cuts[d], cuts[d + 1]);
    loosb200_stats* st = stats ? &ps[d] : nullptr;
    if (st) { st->sum = 0.0; st->max = 0.0; st->count = 0; }
    if (rows.empty()) return LOOSB200_OK;
    OutKind k = kind;
    if (kind == OutKind::Device && d > 0 && !peer_stores) k = OutKind::Copy;
    return run_pairs(a[d], b ? b[d] : a[d], self, self, false, rows, out, k, ld, st, engine);
  });
  if (rc == LOOSB200_OK) merge_stats(stats, ps);
  return rc;
}

}  // namespace loosb200

using namespace loosb200;

extern "C" int loosb200_replicate(const loosb200_frames* src, int device, loosb200_frames** out) {
  if (!src || !out || src->stager) return LOOSB200_ERR_INVALID;
  if (device < 0 || device >= loosb200_device_count()) return LOOSB200_ERR_NO_DEVICE;
  NvtxRange nvtx("loosb200 replicate (device to device)");
  // an empty staged set on `device`, then plane-for-plane copies (peer-to-peer over NVLink when the
  // devices allow it, through the host otherwise); the digit planes are re-derived there on first use
  // from the same fp32 planes and centroids, hence bit-identical
  loosb200_frames* h = nullptr;
  int rc = alloc_frames(&h, src->F, src->N, device);
  if (rc) return rc;
  enable_peer(device, src->device);
  enable_peer(src->device, device);
  LB_CUDA(cudaSetDevice(device));
  cudaStream_t st = h->stream;
  cudaError_t e = cudaMemcpyPeerAsync(h->raw, device, src->raw, src->device, sizeof(float) * src->F * 3 * src->Npad, st);
  if (e == cudaSuccess) e = cudaMemcpyPeerAsync(h->centroid, device, src->centroid, src->device, sizeof(double) * 3 * src->F, st);
  if (e == cudaSuccess) e = cudaMemcpyPeerAsync(h->gd, device, src->gd, src->device, sizeof(double) * src->F, st);
  if (e == cudaSuccess) e = cudaMemcpyPeerAsync(h->maxabs, device, src->maxabs, src->device, sizeof(float) * src->F, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { loosb200_free(h); return cuda_fail(e, "replicate", __FILE__, __LINE__); }
  h->timing.valid = false;
  *out = h;
  return LOOSB200_OK;
}

// Host frames -> a replica on every device, with the upload itself spread over the devices: device d
// stages frames [F d/n, F (d+1)/n) from the caller's buffer over ITS OWN PCIe link (gather of the
// selection, centroids, G: the ordinary staging path), then every device collects the other slices
// device-to-device.  The host buffer crosses PCIe once in total, n links wide.
extern "C" int loosb200_stage_frames_multi(loosb200_frames** out, const int* devices, int n, const float* host_xyz,
                                           size_t F, size_t natoms, const uint32_t* sel, size_t nsel, int layout) {
  if (!out || !devices || n < 1 || !host_xyz) return LOOSB200_ERR_INVALID;
  for (int d = 0; d < n; ++d) out[d] = nullptr;
  if (n == 1 || F < (size_t)n * 64) {  // nothing to gain from splitting a small upload
    int rc = loosb200_stage_frames(&out[0], host_xyz, F, natoms, sel, nsel, layout, devices[0]);
    for (int d = 1; d < n && rc == LOOSB200_OK; ++d) rc = loosb200_replicate(out[0], devices[d], &out[d]);
    if (rc) for (int d = 0; d < n; ++d) { loosb200_free(out[d]); out[d] = nullptr; }
    return rc;
  }
  NvtxRange nvtx("loosb200 stage frames on several devices");
  std::vector<size_t> off(n + 1);
  for (int d = 0; d <= n; ++d) off[d] = F * (size_t)d / (size_t)n;
  std::vector<loosb200_frames*> part(n, nullptr);
  int rc = on_all_devices(n, [&](int d) -> int {
    return loosb200_stage_frames(&part[d], host_xyz + off[d] * natoms * 3, off[d + 1] - off[d], natoms, sel, nsel, layout, devices[d]);
  });
  if (rc == LOOSB200_OK)
    for (int d = 0; d < n; ++d)
      for (int s = 0; s < n; ++s) enable_peer(devices[d], devices[s]);
  for (int d = 0; d < n && rc == LOOSB200_OK; ++d) rc = alloc_frames(&out[d], F, part[0]->N, devices[d]);
  cudaError_t e = cudaSuccess;
  if (rc == LOOSB200_OK) {
    for (int d = 0; d < n && e == cudaSuccess; ++d) {  // every device pulls all slices (its own included) on its stream
      loosb200_frames* h = out[d];
      cudaSetDevice(h->device);
      for (int k = 0; k < n && e == cudaSuccess; ++k) {
        const int s = (d + k) % n;  // start with the local slice, then walk the peers in a different order per device
        const loosb200_frames* p = part[s];
        const size_t f0 = off[s], nf = off[s + 1] - off[s];
        e = cudaMemcpyPeerAsync(h->raw + f0 * 3 * h->Npad, h->device, p->raw, p->device, sizeof(float) * nf * 3 * h->Npad, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(h->centroid + f0 * 3, h->device, p->centroid, p->device, sizeof(double) * 3 * nf, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(h->gd + f0, h->device, p->gd, p->device, sizeof(double) * nf, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(h->maxabs + f0, h->device, p->maxabs, p->device, sizeof(float) * nf, h->stream);
      }
    }
    for (int d = 0; d < n; ++d) {
      cudaSetDevice(out[d]->device);
      const cudaError_t es = cudaStreamSynchronize(out[d]->stream);
      if (e == cudaSuccess) e = es;
    }
  }
  for (int d = 0; d < n; ++d) loosb200_free(part[d]);
  if (rc == LOOSB200_OK && e != cudaSuccess) rc = cuda_fail(e, "stage_frames_multi", __FILE__, __LINE__);
  if (rc) for (int d = 0; d < n; ++d) { loosb200_free(out[d]); out[d] = nullptr; }
  return rc;
}

extern "C" int loosb200_rmsds_self_multi(loosb200_frames* const* replicas, int n, float* out, size_t ld,
                                         loosb200_stats* stats, int engine) {
  int rc = check_replicas(replicas, n);
  if (rc) return rc;
  if (out && ld < replicas[0]->F) return LOOSB200_ERR_INVALID;
  return pairs_multi(replicas, nullptr, n, true, out, out ? OutKind::Copy : OutKind::None, ld, stats, engine);
}

extern "C" int loosb200_rmsds_self_multi_device(loosb200_frames* const* replicas, int n, float* out_dev, size_t ld,
                                                loosb200_stats* stats, int engine) {
  int rc = check_replicas(replicas, n);
  if (rc) return rc;
  if (out_dev && ld < replicas[0]->F) return LOOSB200_ERR_INVALID;
  return pairs_multi(replicas, nullptr, n, true, out_dev, out_dev ? OutKind::Device : OutKind::None, ld, stats, engine);
}

extern "C" int loosb200_rmsds_cross_multi(loosb200_frames* const* a, loosb200_frames* const* b, int n, float* out,
                                          size_t ld, loosb200_stats* stats, int engine) {
  int rc = check_replicas(a, n);
  if (!rc) rc = check_replicas(b, n);
  if (rc) return rc;
  for (int d = 0; d < n; ++d) {
    if (a[d]->device != b[d]->device) return LOOSB200_ERR_INVALID;
    if (a[d]->N != b[d]->N) return LOOSB200_ERR_SIZE_MISMATCH;
  }
  if (out && ld < a[0]->F) return LOOSB200_ERR_INVALID;
  return pairs_multi(a, b, n, false, out, out ? OutKind::Copy : OutKind::None, ld, stats, engine);
}

extern "C" void* loosb200_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0) bytes = 1;
  // portable: pinned for every device of the process, so each GPU can copy its strips into it
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}

extern "C" void loosb200_host_free(void* p) {
  if (p) { cudaFreeHost(p); cudaGetLastError(); }
}

extern "C" int loosb200_share_cuts(size_t F, int nparts, int self, const double* weights, uint32_t* cuts_out) {
  if (F == 0 || nparts < 1 || !cuts_out) return LOOSB200_ERR_INVALID;
  const std::vector<uint32_t> c = contiguous_cuts(F, nparts, self != 0, weights);
  for (int p = 0; p <= nparts; ++p) cuts_out[p] = c[p];
  return LOOSB200_OK;
}
