// TEST INFRASTRUCTURE ONLY (oracle/): the sliver of boost::format that the reference code compiled here uses --
// printf-style directives ("%d", "%s", "%u", "%.4f", "%8.3e" ...) filled in order by operator% and streamed with
// operator<< -- so that src/MatrixImpl.hpp, src/MatrixRead.hpp and the statistics functions of the tools compile
// without Boost.  A directive sets the stream state the way boost::format does (f: fixed, e: scientific, g: general;
// precision; width; '-' left, '0' zero fill, '+' showpos) and the argument is then streamed with operator<<.
// (Not the swallow-everything stub of shim/loos_defs.hpp: here the text must really be produced.)
#if !defined(LOOS_SHIM_BOOST_FORMAT_HPP)
#define LOOS_SHIM_BOOST_FORMAT_HPP
#include <cctype>
#include <cstdlib>
#include <iomanip>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>
namespace boost {
class format {
  struct Piece { std::string text; bool directive; std::string flags; int width, prec; char type; };
 public:
  explicit format(const char* f) { parse(f); }
  explicit format(const std::string& f) { parse(f); }
  template <class T> format& operator%(const T& v) {
    while (next_ < pieces_.size() && !pieces_[next_].directive) ++next_;
    if (next_ < pieces_.size()) {
      const Piece& p = pieces_[next_];
      std::ostringstream os;
      if (p.type == 'f' || p.type == 'F') os.setf(std::ios::fixed, std::ios::floatfield);
      else if (p.type == 'e' || p.type == 'E') os.setf(std::ios::scientific, std::ios::floatfield);
      if (p.prec >= 0) os.precision(p.prec);
      if (p.width > 0) os.width(p.width);
      if (p.flags.find('-') != std::string::npos) os.setf(std::ios::left, std::ios::adjustfield);
      if (p.flags.find('0') != std::string::npos) os.fill('0');
      if (p.flags.find('+') != std::string::npos) os.setf(std::ios::showpos);
      os << v;
      pieces_[next_].text = os.str();
      ++next_;
    }
    return *this;
  }
  std::string str() const {
    std::string out;
    for (const Piece& p : pieces_) out += p.text;
    return out;
  }
 private:
  void parse(const std::string& f) {
    std::string lit;
    for (size_t i = 0; i < f.size(); ++i) {
      if (f[i] != '%') { lit += f[i]; continue; }
      if (i + 1 < f.size() && f[i + 1] == '%') { lit += '%'; ++i; continue; }
      if (!lit.empty()) { pieces_.push_back(Piece{lit, false, "", 0, -1, 0}); lit.clear(); }
      Piece p{"", true, "", 0, -1, 's'};
      size_t j = i + 1;
      while (j < f.size() && std::string("-+0 #").find(f[j]) != std::string::npos) p.flags += f[j++];
      while (j < f.size() && isdigit((unsigned char)f[j])) p.width = p.width * 10 + (f[j++] - '0');
      if (j < f.size() && f[j] == '.') { ++j; p.prec = 0; while (j < f.size() && isdigit((unsigned char)f[j])) p.prec = p.prec * 10 + (f[j++] - '0'); }
      while (j < f.size() && (f[j] == 'l' || f[j] == 'h' || f[j] == 'z')) ++j;  // length modifiers mean nothing here
      if (j < f.size()) p.type = f[j];
      pieces_.push_back(p);
      i = j;
    }
    if (!lit.empty()) pieces_.push_back(Piece{lit, false, "", 0, -1, 0});
  }
  std::vector<Piece> pieces_;
  size_t next_ = 0;
};
inline std::ostream& operator<<(std::ostream& o, const format& f) { return o << f.str(); }
}  // namespace boost
#endif
