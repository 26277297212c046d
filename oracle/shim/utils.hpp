// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the reference's src/utils.hpp (which needs Boost) so
// that the reference's own src/xdr.hpp compiles.  The one function xdr.hpp takes from it, loos::swab, is
// NOT restated: oracle/Makefile extracts src/utils.hpp:301-318 into _ref/utils_swab.inc, included below.
// The exception types mirror src/exceptions.hpp:40-92,180-230 as far as the extracted XTC code throws them.
#if !defined(LOOS_UTILS_HPP)
#define LOOS_UTILS_HPP
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#include <loos_defs.hpp>

namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }

namespace loos {
struct LOOSError : public std::runtime_error {
  explicit LOOSError(const std::string& s) : std::runtime_error(s) {}
};
struct XDRDataSizeError : public LOOSError {
  XDRDataSizeError() : LOOSError("XDR data size error") {}
};
struct FileReadError : public LOOSError {
  FileReadError(const std::string& fname, const std::string& msg) : LOOSError(fname + ": " + msg) {}
};
struct FileWriteError : public LOOSError {
  FileWriteError(const std::string& fname, const std::string& msg) : LOOSError(fname + ": " + msg) {}
};
#include "_ref/utils_swab.inc"  // verbatim reference lines (template <typename T> T swab(const T&))
}  // namespace loos
#endif
