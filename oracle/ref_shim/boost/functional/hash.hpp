/* Minimal stand-in for <boost/functional/hash.hpp> so the unmodified reference
 * sources compile in this image (boost is not installed). Only DecoderHash /
 * raster hashes use it; it never influences a decoded pixel. Test scaffolding. */
#pragma once
#include <cstddef>
#include <functional>
namespace boost {
template <class T> inline void hash_combine(std::size_t& seed, const T& v) {
  seed ^= std::hash<T>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}
template <class It> inline void hash_range(std::size_t& seed, It first, It last) {
  for (; first != last; ++first) hash_combine(seed, *first);
}
template <class It> inline std::size_t hash_range(It first, It last) {
  std::size_t seed = 0; hash_range(seed, first, last); return seed;
}
}  // namespace boost
