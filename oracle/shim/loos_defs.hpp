// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the reference's
// src/loos_defs.hpp so that the reference's own Coord.hpp, Matrix44.hpp,
// XForm.{hpp,cpp} and the arithmetic core of src/alignment.cpp compile without
// Boost.  Mirrors the typedefs at /root/reference/src/loos_defs.hpp:51-55,82,
// 113-122 and the Fortran prototypes at :60-79.  Nothing here is product code.
#if !defined(LOOS_DEFS_HPP)
#define LOOS_DEFS_HPP

#include <sys/types.h>
#include <cmath>
#include <vector>
#include <string>
#include <tuple>
#include <iostream>
#include <stdexcept>

typedef unsigned int uint;
typedef unsigned long ulong;
typedef int f77int;

extern "C" {
void dgesvj_(char*, char*, char*, f77int*, f77int*, double*, f77int*, double*,
             f77int*, double*, f77int*, double*, f77int*, f77int*);
void dgemm_(const char*, const char*, const f77int*, const f77int*, const f77int*,
            const double*, const double*, const f77int*, const double*, const f77int*,
            const double*, double*, const f77int*);
double dlamch_(const char*);
void openblas_set_num_threads(int);
}

namespace loos {
typedef double greal;
typedef long gint;
template <class T> class Coord;
typedef Coord<double> GCoord;
template <class T> class Matrix44;
typedef Matrix44<double> GMatrix;

struct NumericalError : public std::runtime_error {
  NumericalError(const std::string& s, const int info)
      : std::runtime_error(s + " (info=" + std::to_string(info) + ")") {}
};
}  // namespace loos

// boost::format is only used for a warning message on the hot path
// (alignment.cpp:56-58); swallow it.  (The matrix wrapper needs the header line it formats and brings
// shim/boost/format.hpp instead: LOOS_SHIM_REAL_FORMAT.)
#if !defined(LOOS_SHIM_REAL_FORMAT)
namespace boost {
struct fmtstub {
  template <class T> fmtstub& operator%(const T&) { return *this; }
};
inline fmtstub format(const char*) { return fmtstub(); }
inline std::ostream& operator<<(std::ostream& o, const fmtstub&) { return o; }
}  // namespace boost
#endif

#endif
