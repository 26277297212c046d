// TEST INFRASTRUCTURE ONLY (oracle/): boost::tuple -> std::tuple alias, enough
// for SVDTupleVec at /root/reference/src/alignment.hpp:44.
#pragma once
#include <tuple>
namespace boost {
template <class... A> using tuple = std::tuple<A...>;
template <int I, class T> auto get(T&& t) -> decltype(std::get<I>(t)) { return std::get<I>(t); }
}  // namespace boost
