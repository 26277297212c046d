// TEST INFRASTRUCTURE ONLY (oracle/): the sliver of boost::format that the reference's Matrix headers use --
// "%d" / "%s" / "%u" directives filled in order by operator% and streamed with operator<< -- so that
// src/MatrixImpl.hpp and src/MatrixRead.hpp compile without Boost.  (Not the swallow-everything stub of
// shim/loos_defs.hpp: here the header line "# m n (0)" must really be produced.)
#if !defined(LOOS_SHIM_BOOST_FORMAT_HPP)
#define LOOS_SHIM_BOOST_FORMAT_HPP
#include <ostream>
#include <sstream>
#include <string>
#include <vector>
namespace boost {
class format {
 public:
  explicit format(const char* f) : fmt_(f) {}
  explicit format(const std::string& f) : fmt_(f) {}
  template <class T> format& operator%(const T& v) {
    std::ostringstream os;
    os << v;
    args_.push_back(os.str());
    return *this;
  }
  std::string str() const {
    std::string out;
    size_t a = 0;
    for (size_t i = 0; i < fmt_.size(); ++i) {
      if (fmt_[i] != '%') { out += fmt_[i]; continue; }
      if (i + 1 < fmt_.size() && fmt_[i + 1] == '%') { out += '%'; ++i; continue; }
      size_t j = i + 1;
      while (j < fmt_.size() && !isalpha((unsigned char)fmt_[j])) ++j;  // flags / width / precision are not used by the callers
      if (a < args_.size()) out += args_[a++];
      i = j;
    }
    return out;
  }
 private:
  std::string fmt_;
  std::vector<std::string> args_;
};
inline std::ostream& operator<<(std::ostream& o, const format& f) { return o << f.str(); }
}  // namespace boost
#endif
