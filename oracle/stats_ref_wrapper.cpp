// TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by or called from the product path.
//
// The REFERENCE'S OWN statistics lines of the tools, compiled from the sources where they lie:
//   * Tools/rmsds.cpp:425-456        showStatsHalf, showStatsWhole   (multi-rmsds has the same two)
//   * Tools/rms-overlap.cpp:458-521  showStats, showFractionalStats  (with its brace-less `if`)
// They are free functions on a RealMatrix (src/MatrixImpl.hpp, included in place) that print through boost::format to
// cerr / cout; the shims of oracle/shim/boost stand in for Boost, and the text is captured by swapping the stream
// buffers for the duration of a call.
#define LOOS_SHIM_REAL_FORMAT
#include <loos_defs.hpp>          // oracle/shim
#include <boost/format.hpp>       // oracle/shim
#include <boost/tuple/tuple.hpp>  // oracle/shim
#define LOOS_MATRIX_HPP           // src/Matrix.hpp's guard: MatrixUtils / MatrixOps stay out
#include <MatrixImpl.hpp>         // /root/reference/src

#include <cstring>
#include <iostream>
#include <sstream>

namespace loos { typedef Math::Matrix<float, Math::ColMajor> RealMatrix; }
using namespace std;
using namespace loos;

namespace rmsds_tool {
#include "_ref/rmsds_stats.inc"      // showStatsHalf, showStatsWhole
}
namespace overlap_tool {
#include "_ref/overlap_stats.inc"    // showStats, showFractionalStats
}

namespace {
RealMatrix fill(const float* M, int m, int n, long ld) {
  RealMatrix R(m, n);
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < n; ++i) R(j, i) = M[(size_t)i * ld + j];
  return R;
}
template <class F> long capture(F f, char* out, long cap) {  // what the call writes to cerr and cout, concatenated
  std::ostringstream e, o;
  std::streambuf* pe = std::cerr.rdbuf(e.rdbuf());
  std::streambuf* po = std::cout.rdbuf(o.rdbuf());
  try { f(); } catch (...) { std::cerr.rdbuf(pe); std::cout.rdbuf(po); return -1; }
  std::cerr.rdbuf(pe);
  std::cout.rdbuf(po);
  const std::string s = e.str() + o.str();
  if ((long)s.size() + 1 > cap) return -2;
  memcpy(out, s.c_str(), s.size() + 1);
  return (long)s.size();
}
}  // namespace

extern "C" {
long ref_stats_half(const float* M, int n, long ld, char* out, long cap) {
  const RealMatrix R = fill(M, n, n, ld);
  return capture([&]() { rmsds_tool::showStatsHalf(R); }, out, cap);
}
long ref_stats_whole(const float* M, int m, int n, long ld, char* out, long cap) {
  const RealMatrix R = fill(M, m, n, ld);
  return capture([&]() { rmsds_tool::showStatsWhole(R); }, out, cap);
}
long ref_overlap_stats(const float* M, int m, int n, long ld, char* out, long cap) {
  const RealMatrix R = fill(M, m, n, ld);
  return capture([&]() { overlap_tool::showStats(R); }, out, cap);
}
long ref_overlap_fractional(const float* M, int m, int n, long ld, float cutoff, int is_noop, char* out, long cap) {
  const RealMatrix R = fill(M, m, n, ld);
  return capture([&]() { overlap_tool::showFractionalStats(R, cutoff, is_noop != 0); }, out, cap);
}
}  // extern "C"
