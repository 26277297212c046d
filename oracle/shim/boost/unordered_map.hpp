// TEST INFRASTRUCTURE ONLY (oracle/): boost::unordered_map -> std::unordered_map (src/MatrixStorage.hpp, sparse policy).
#if !defined(LOOS_SHIM_BOOST_UNORDERED_MAP_HPP)
#define LOOS_SHIM_BOOST_UNORDERED_MAP_HPP
#include <unordered_map>
namespace boost { using std::unordered_map; }
#endif
