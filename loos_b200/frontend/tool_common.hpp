// Shared pieces of the three tool mains.
#pragma once
#include <cstdio>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

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
  o.add("gpu", 0, "0", "CUDA device to use (extension; the reference has no such option)");
  o.add("engine", 0, "auto", "auto|tensor|fp64 contraction engine (extension)");
}

inline int engineOf(const Options& o) {
  const std::string e = o.str("engine");
  if (e == "tensor") return LOOSB200_ENGINE_TENSOR;
  if (e == "fp64") return LOOSB200_ENGINE_FP64;
  if (e == "auto") return LOOSB200_ENGINE_AUTO;
  throw Error("the argument ('" + e + "') for option '--engine' is invalid");
}

// readCoords (src/ensembles.cpp:253-296): the listed frames of one trajectory as
// planes [F][3][natoms]; the device gathers the selection.
inline std::vector<float> readFrames(DCD& traj, const std::vector<uint32_t>& indices) {
  const size_t na = traj.natoms();
  std::vector<float> buf(indices.size() * 3 * na);
  for (size_t j = 0; j < indices.size(); ++j) traj.readFramePlanes(indices[j], buf.data() + j * 3 * na);
  return buf;
}

struct Frames {  // RAII around loosb200_frames
  loosb200_frames* h = nullptr;
  ~Frames() { if (h) loosb200_free(h); }
};

inline void stage(Frames& f, const std::vector<float>& planes, size_t F, size_t natoms,
                  const std::vector<uint32_t>& sel, int device) {
  check(loosb200_stage_frames(&f.h, planes.data(), F, natoms, sel.data(), sel.size(), LOOSB200_LAYOUT_PLANES, device),
        "staging frames");
}

}  // namespace lbfe
