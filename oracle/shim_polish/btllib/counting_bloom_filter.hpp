// TEST INFRASTRUCTURE ONLY (oracle/).  Stand-in for btllib's <btllib/counting_bloom_filter.hpp>: the
// one method the GoldPolish targeted-Bloom-filter builder calls on KmerCountingBloomFilter8
// (subprojects/goldpolish/src/utils.cpp:115-117).  btllib is third party, not vendored, not
// installable offline: the semantics below are RECALLED from btllib >= 1.4 and PARITY UNPINNED:
//   counters are uint8, one per byte of the filter; counter of hash h = array[h % bytes];
//   contains(hashes) = the minimum of the hash_num counters;
//   insert(hashes, min) increments only the counters that hold that minimum (conservative update)
//   and saturates at 255;
//   insert_thresh_contains(hashes, t): count = contains; if count < t: insert, count + 1 is returned;
//   else count is returned (the element's count after the conditional insert).
#ifndef GRB_SHIM_POLISH_BTLLIB_CBF_HPP
#define GRB_SHIM_POLISH_BTLLIB_CBF_HPP
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

namespace btllib {

class KmerCountingBloomFilter8
{
public:
  KmerCountingBloomFilter8(size_t bytes, unsigned hash_num, unsigned k)
    : array_(bytes, 0), h_(hash_num), k_(k)
  {
  }
  uint8_t contains(const uint64_t* hashes) const
  {
    uint8_t m = array_[hashes[0] % array_.size()];
    for (unsigned i = 1; i < h_; ++i) {
      const uint8_t c = array_[hashes[i] % array_.size()];
      m = c < m ? c : m;
    }
    return m;
  }
  void insert(const uint64_t* hashes, uint8_t min_val)
  {
    if (min_val == std::numeric_limits<uint8_t>::max()) {
      return;
    }
    for (unsigned i = 0; i < h_; ++i) {
      uint8_t& c = array_[hashes[i] % array_.size()];
      if (c == min_val) {
        c = (uint8_t)(min_val + 1);
      }
    }
  }
  uint8_t insert_thresh_contains(const uint64_t* hashes, uint8_t threshold)
  {
    const uint8_t count = contains(hashes);
    if (count < threshold) {
      insert(hashes, count);
      return (uint8_t)(count + 1);
    }
    return count;
  }
  unsigned get_k() const { return k_; }
  const std::vector<uint8_t>& raw() const { return array_; }

private:
  std::vector<uint8_t> array_;
  unsigned h_, k_;
};

} // namespace btllib
#endif
