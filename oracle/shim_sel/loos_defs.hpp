// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for src/loos_defs.hpp for compiling the reference's own selection
// scanner / grammar / kernel (scanner.cc, grammar.cc, Kernel*.cpp, Atom.cpp) without Boost.
// Mirrors the typedefs at /root/reference/src/loos_defs.hpp:51-55,113-123,143,167.
#if !defined(LOOS_DEFS_HPP)
#define LOOS_DEFS_HPP
#include <sys/types.h>
#include <cmath>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
typedef unsigned int uint;
typedef unsigned long ulong;
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }
namespace loos {
typedef double greal;
typedef long gint;
template <class T> class Coord;
typedef Coord<double> GCoord;
class Atom;
class AtomicGroup;
typedef boost::shared_ptr<Atom> pAtom;
}  // namespace loos
#endif
