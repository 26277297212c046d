// TEST INFRASTRUCTURE ONLY (oracle/): stub for the reference's utils_random.hpp,
// which Coord.hpp:30 includes for Coord<T>::random() (Coord.hpp:478-490).
// That member is never instantiated by the oracle; the names merely have to parse.
#if !defined(UTILS_RANDOM_HPP)
#define UTILS_RANDOM_HPP
#include <loos_defs.hpp>
namespace boost {
struct mt19937 {};
template <class T = double> struct uniform_on_sphere {
  explicit uniform_on_sphere(int) {}
};
template <class E, class D> struct variate_generator {
  variate_generator(E, D) {}
  std::vector<double> operator()() { return std::vector<double>(3, 0.0); }
};
}  // namespace boost
namespace loos {
typedef boost::mt19937 base_generator_type;
inline base_generator_type& rng_singleton(void) {
  static base_generator_type g;
  return g;
}
}  // namespace loos
#endif
