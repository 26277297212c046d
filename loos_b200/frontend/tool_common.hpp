// Shared pieces of the three tool mains.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <exception>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include "../../include/loosb200.h"
#include "frontend.hpp"

namespace lbfe {

inline void check(int rc, const char* what) {
  if (rc != LOOSB200_OK) {
    std::string msg = std::string("Error- ") + what + ": " + loosb200_strerror(rc);
    const char* extra = loosb200_last_error();
    if (rc == LOOSB200_ERR_CUDA && extra && *extra) msg += std::string(" (") + extra + ")";
    throw Error(msg);
  }
}

inline void addBasicOptions(Options& o) {  // BasicOptions, src/OptionsFramework.cpp:30-37
  o.add("help", 'h', "", "Produce this message", true);
  o.add("fullhelp", 0, "", "More detailed help", true);
  o.add("verbosity", 'v', "0", "Verbosity of output (if available)");
  o.add("config", 0, "", "Options config file");
  o.add("gpu", 0, "0", "CUDA device to use (extension; the reference has no such option)");
  o.add("gpus", 0, "1", "Number of GPUs the matrix is split over, starting at --gpu (extension)");
  o.add("engine", 0, "auto", "auto|tensor|fp64 contraction engine (extension)");
}

// --gpus: how many devices work on the matrix (the GPU counterpart of --threads)
inline int gpusOf(const Options& o) {
  const unsigned long n = o.uns("gpus");
  const int have = loosb200_device_count();
  if (n < 1 || (have > 0 && n > (unsigned long)have))
    throw Error("the argument ('" + o.str("gpus") + "') for option '--gpus' is invalid: " + std::to_string(have) + " usable device(s)");
  return (int)n;
}

inline int engineOf(const Options& o) {
  const std::string e = o.str("engine");
  if (e == "tensor") return LOOSB200_ENGINE_TENSOR;
  if (e == "fp64") return LOOSB200_ENGINE_FP64;
  if (e == "auto") return LOOSB200_ENGINE_AUTO;
  throw Error("the argument ('" + e + "') for option '--engine' is invalid");
}

// loosb200_sink writing to stdout: the matrix text arrives already formatted from the GPU
inline int stdoutSink(void*, const char* data, size_t n) { return fwrite(data, 1, n, stdout) == n ? 0 : 1; }
// The GPU formats the text unless the float matrix itself is needed on the host (--binary, or
// rms-overlap's fractional statistics), the precision exceeds what it handles exactly, or
// LOOSB200_HOST_TEXT=1 asks for the host writer.
inline bool gpuText(bool need_host_matrix, unsigned long precision) {
  const char* env = getenv("LOOSB200_HOST_TEXT");
  if (!need_host_matrix && precision > 9) {
    static bool told = false;  // a float carries 9 significant digits; beyond that the host writer pads like printf
    if (!told) std::cerr << "Note- precision " << precision << " exceeds the 9 digits the GPU formatter writes exactly; the matrix is copied to the host and formatted there (slower)\n";
    told = true;
  }
  return !need_host_matrix && precision <= 9 && !(env && atoi(env) != 0);
}

struct Frames {  // RAII around loosb200_frames
  loosb200_frames* h = nullptr;
  ~Frames() { if (h) loosb200_free(h); }
};

// A staged set and its copies on the other devices of the job; h[0] is the original (not owned).
struct Replicas {
  std::vector<loosb200_frames*> h;
  Replicas() {}
  Replicas(const Replicas&) = delete;
  Replicas& operator=(const Replicas&) = delete;
  ~Replicas() { for (size_t k = 1; k < h.size(); ++k) if (h[k]) loosb200_free(h[k]); }
  void build(const Frames& first, int first_device, int ngpus) {
    h.assign(1, first.h);
    const int have = std::max(1, loosb200_device_count());
    for (int k = 1; k < ngpus; ++k) {
      loosb200_frames* r = nullptr;
      check(loosb200_replicate(first.h, (first_device + k) % have, &r), "replicating frames");
      h.push_back(r);
    }
  }
};

// The F x F float matrix on the host, when it is needed there at all (--binary, -p > 9, fractional
// statistics): page-locked for every device so that the result copies run at the PCIe rate
// (40 GB at 100,000 frames), pageable only if the host refuses to pin that much.
struct HostMatrix {
  float* p = nullptr;
  bool pinned = false;
  HostMatrix() {}
  HostMatrix(const HostMatrix&) = delete;
  HostMatrix& operator=(const HostMatrix&) = delete;
  ~HostMatrix() { release(); }
  void release() {
    if (p) { if (pinned) loosb200_host_free(p); else free(p); }
    p = nullptr;
  }
  void allocate(size_t count) {
    release();
    p = static_cast<float*>(loosb200_host_alloc(std::max<size_t>(1, count) * sizeof(float)));
    pinned = p != nullptr;
    if (!p) p = static_cast<float*>(calloc(std::max<size_t>(1, count), sizeof(float)));
    if (!p) throw Error("Error- cannot allocate the " + std::to_string(count * sizeof(float) >> 20) + " MB result matrix");
  }
  float* data() const { return p; }
};

// The reference warns before it fills two thirds of the machine with the coordinate cache and the
// matrix (checkMemoryUsage, Tools/rmsds.cpp:58-60,467-485,516-518).  Here the coordinates live in
// HBM, so the estimate is the host matrix alone (zero when the text is formatted on the GPU).
inline void warnHostMemory(size_t used_bytes, long verbosity) {
  const long pagesize = sysconf(_SC_PAGESIZE), pages = sysconf(_SC_PHYS_PAGES);  // loos::availableMemory(), src/utils.cpp:433-439
  if (pagesize <= 0 || pages <= 0) return;
  const double mem = (double)pagesize * (double)pages;
  char line[256];
  if (verbosity > 2) {
    snprintf(line, sizeof(line), "Memory: available=%ld GB, estimated used=%.2f MB\n", (long)(mem / (double)(1ul << 30)), (double)used_bytes / (double)(1ul << 20));
    std::cerr << line;
  }
  if ((double)used_bytes / mem >= 0.66) {
    snprintf(line, sizeof(line), "***WARNING***\nThe estimated memory used is %.1f%% (%zu MB) of your total memory (%ld GB).\n",
             100.0 * (double)used_bytes / mem, used_bytes >> 20, (long)(mem / (double)(1ul << 30)));
    std::cerr << line << "If your machine starts swapping, try subsampling the trajectories\n";
  }
}

// readCoords (src/ensembles.cpp:253-296) without the host-side coordinate cache: frames
// are decoded in batches (planes X|Y|Z per frame, the DCD record order) and appended to
// one or more staged sets; while batch k+1 is being decoded, batch k is on its way to
// the GPU (pinned bounce buffer -> async H2D on a side stream -> gather kernel).
// `read(j, dst)` must put frame j of the list (3 * natoms floats) at dst.
// Reader threads for the staging loop (LOOSB200_READER_THREADS, default: the host's threads, at most 8).
inline unsigned readerThreads() {
  if (const char* env = getenv("LOOSB200_READER_THREADS")) return (unsigned)std::max(1, atoi(env));
  return std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
}

template <class ReadFrame>
inline void stageStreaming(std::vector<Frames*> sets, const std::vector<const std::vector<uint32_t>*>& sels,
                           size_t F, size_t natoms, int device, ReadFrame read, std::vector<float>* keep = nullptr) {
  for (size_t k = 0; k < sets.size(); ++k)
    check(loosb200_stage_begin(&sets[k]->h, F, natoms, sels[k]->data(), sels[k]->size(), LOOSB200_LAYOUT_PLANES, device),
          "staging frames");
  const size_t batch = std::max<size_t>(1, std::min<size_t>(F, (16u << 20) / (natoms * 12)));
  std::vector<float> buf(batch * 3 * natoms);
  if (keep) keep->resize(F * 3 * natoms);
  const unsigned nthreads = readerThreads();
  for (size_t j0 = 0; j0 < F; j0 += batch) {
    const size_t nb = std::min(batch, F - j0);
    // the frames of a batch are decoded by several threads (Trajectory::readFramePlanes is re-entrant):
    // XTC decompression and NetCDF byte swapping are CPU work, and even DCD reads gain from parallel preads
    const unsigned nt = (unsigned)std::min<size_t>(nthreads, nb);
    if (nt <= 1) {
      for (size_t j = 0; j < nb; ++j) read(j0 + j, buf.data() + j * 3 * natoms);
    } else {
      std::vector<std::thread> pool;
      std::vector<std::exception_ptr> errs(nt);
      for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([&, t]() {
          try {
            for (size_t j = t; j < nb; j += nt) read(j0 + j, buf.data() + j * 3 * natoms);
          } catch (...) { errs[t] = std::current_exception(); }
        });
      for (std::thread& th : pool) th.join();
      for (const std::exception_ptr& e : errs) if (e) std::rethrow_exception(e);
    }
    if (keep) std::copy(buf.begin(), buf.begin() + nb * 3 * natoms, keep->begin() + j0 * 3 * natoms);
    for (size_t k = 0; k < sets.size(); ++k) check(loosb200_stage_append(sets[k]->h, buf.data(), nb), "staging frames");
  }
  for (size_t k = 0; k < sets.size(); ++k) check(loosb200_stage_end(sets[k]->h), "staging frames");
}

inline void stageTrajectory(Frames& f, Trajectory& traj, const std::vector<uint32_t>& indices, size_t natoms,
                            const std::vector<uint32_t>& sel, int device) {
  stageStreaming({&f}, {&sel}, indices.size(), natoms, device,
                 [&](size_t j, float* dst) { traj.readFramePlanes(indices[j], dst); });
}

}  // namespace lbfe
