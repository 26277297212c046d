// TEST INFRASTRUCTURE ONLY (oracle/): boost::shared_array as far as src/MatrixStorage.hpp uses it.
#if !defined(LOOS_SHIM_BOOST_SHARED_ARRAY_HPP)
#define LOOS_SHIM_BOOST_SHARED_ARRAY_HPP
#include <cstddef>
#include <memory>
namespace boost {
template <class T> class shared_array {
 public:
  shared_array() {}
  explicit shared_array(T* p) : p_(p, std::default_delete<T[]>()) {}
  T& operator[](std::ptrdiff_t i) const { return p_.get()[i]; }
  T* get() const { return p_.get(); }
  void reset() { p_.reset(); }
  void reset(T* p) { p_.reset(p, std::default_delete<T[]>()); }
 private:
  std::shared_ptr<T> p_;
};
}  // namespace boost
#endif
