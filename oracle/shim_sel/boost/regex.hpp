// TEST INFRASTRUCTURE ONLY (oracle/): boost::regex as src/KernelActions.{hpp,cpp} use it (perl | icase, regex_search
// with and without a match object), on top of std::regex (ECMAScript, icase).  The native front-end uses std::regex
// with the same flags, so for the `=~` and `->` operators this build checks everything AROUND the regular expression
// (operand order, value types, the number extracted), not the dialect itself.
#if !defined(LOOS_SHIM_BOOST_REGEX_HPP)
#define LOOS_SHIM_BOOST_REGEX_HPP
#include <regex>
#include <string>
namespace boost {
class regex {
 public:
  enum flag_type { perl = 1, icase = 2 };
  regex() {}
  regex(const std::string& s, int flags = perl)
      : re_(s, (flags & icase) ? std::regex::ECMAScript | std::regex::icase : std::regex::ECMAScript) {}
  const std::regex& std_regex() const { return re_; }
 private:
  std::regex re_;
};
typedef std::smatch smatch;
inline bool regex_search(const std::string& s, const regex& r) { return std::regex_search(s, r.std_regex()); }
inline bool regex_search(const std::string& s, smatch& m, const regex& r) { return std::regex_search(s, m, r.std_regex()); }
}  // namespace boost
#endif
