// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for btllib's <btllib/bloom_filter.hpp>.
// Only needed to parse overloads that are dead code on the GoldRush-Path hot path
// (goldrush_path/MIBFConstructSupport.hpp:111-132,285-320).
#ifndef GRB_SHIM_BTLLIB_BLOOM_FILTER_HPP
#define GRB_SHIM_BTLLIB_BLOOM_FILTER_HPP
#include <cstdint>
#include <vector>
namespace btllib {
class BloomFilter
{
public:
  bool contains(const std::vector<uint64_t>&) const { return false; }
  bool contains(const uint64_t*) const { return false; }
};
} // namespace btllib
#endif
