// TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by or called from the product path.
//
// The REFERENCE'S OWN matrix text output and input: src/MatrixImpl.hpp (operator<<, :208-222), src/MatrixOrder.hpp,
// src/MatrixStorage.hpp and src/MatrixRead.hpp (readAsciiMatrix, :102-143) are header-only and are included
// from where they lie (-I/root/reference/src), behind shims for boost::format / shared_array / unordered_map.
// src/Matrix.hpp itself (which adds MatrixUtils / MatrixOps and with them BLAS and more of Boost) is kept out by
// pre-defining its include guard.  RealMatrix is the typedef of src/loos_defs.hpp:113.
#define LOOS_SHIM_REAL_FORMAT
#include <loos_defs.hpp>      // oracle/shim
#include <boost/format.hpp>   // oracle/shim
#define LOOS_MATRIX_HPP       // src/Matrix.hpp's guard
#include <MatrixImpl.hpp>     // /root/reference/src
#include <MatrixRead.hpp>     // /root/reference/src

#include <cstring>
#include <iomanip>
#include <sstream>

namespace loos { typedef Math::Matrix<float, Math::ColMajor> RealMatrix; }

extern "C" {

// `cout << setprecision(p) << M` (Tools/rmsds.cpp:566-567) for a column-major float matrix with leading dimension ld
long ref_matrix_text(const float* M, int m, int n, long ld, int precision, char* out, long cap) {
  try {
    loos::RealMatrix R(m, n);
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < n; ++i) R(j, i) = M[(size_t)i * ld + j];
    std::ostringstream os;
    os << std::setprecision(precision) << R;
    const std::string s = os.str();
    if ((long)s.size() > cap) return -2;
    memcpy(out, s.data(), s.size());
    return (long)s.size();
  } catch (const std::exception&) { return -1; }
}

// readAsciiMatrix<float, ColMajor, SharedArray>(istream): out column-major m x n; returns 0, -1 on a read error
int ref_matrix_read(const char* text, long len, float* out, long cap, int* m, int* n, char* err, int errcap) {
  try {
    std::istringstream is(std::string(text, (size_t)len));
    loos::RealMatrix R = loos::readAsciiMatrix<float, loos::Math::ColMajor, loos::Math::SharedArray>(is);
    *m = (int)R.rows(); *n = (int)R.cols();
    if ((long)R.rows() * R.cols() > cap) return -2;
    for (unsigned j = 0; j < R.rows(); ++j)
      for (unsigned i = 0; i < R.cols(); ++i) out[(size_t)i * R.rows() + j] = R(j, i);
    return 0;
  } catch (const std::exception& e) {
    if (err && errcap > 0) { strncpy(err, e.what(), (size_t)errcap - 1); err[errcap - 1] = 0; }
    return -1;
  }
}

}  // extern "C"
