// TEST INFRASTRUCTURE ONLY (oracle/).  Stand-in for btllib's <btllib/bloom_filter.hpp>: what the
// GoldPolish targeted-Bloom-filter builder calls on KmerBloomFilter (utils.cpp:118,
// goldpolish_targeted_bfs.cpp:140-142).  RECALLED from btllib, PARITY UNPINNED: bit of hash h =
// h % (bytes * 8), stored LSB first in byte (h % bits) / 8.  save() writes the raw bit array only
// (btllib prepends a TOML header; the file format is outside what is compared).
#ifndef GRB_SHIM_POLISH_BTLLIB_BF_HPP
#define GRB_SHIM_POLISH_BTLLIB_BF_HPP
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace btllib {

class KmerBloomFilter
{
public:
  KmerBloomFilter(size_t bytes, unsigned hash_num, unsigned k)
    : array_(bytes, 0), h_(hash_num), k_(k)
  {
  }
  void insert(const uint64_t* hashes)
  {
    const uint64_t bits = (uint64_t)array_.size() * 8;
    for (unsigned i = 0; i < h_; ++i) {
      const uint64_t pos = hashes[i] % bits;
      array_[pos / 8] |= (uint8_t)(1u << (pos % 8));
    }
  }
  void save(const std::string& path)
  {
    if (FILE* f = fopen(path.c_str(), "wb")) {
      fwrite(array_.data(), 1, array_.size(), f);
      fclose(f);
    }
  }
  unsigned get_k() const { return k_; }
  const std::vector<uint8_t>& raw() const { return array_; }

private:
  std::vector<uint8_t> array_;
  unsigned h_, k_;
};

} // namespace btllib
#endif
