// TEST INFRASTRUCTURE ONLY.  C entry point around the REFERENCE's own fill_bfs
// (subprojects/goldpolish/src/utils.cpp:96-123, compiled unmodified from /root/reference against
// the stand-in btllib headers in shim_polish/), set up the way serve_batch sets its filters up
// (goldpolish_targeted_bfs.cpp:68-77) and called read by read in the given order (:131-135).
// Built into oracle/_ref/libgoldpolish_ref.so by oracle/Makefile when /root/reference is present.
#include "utils.hpp"

#include <cstring>

extern "C" int
grbp_ref_fill(const char* seqs, const uint64_t* off, const uint32_t* thresholds, uint64_t n_reads,
              unsigned hash_num, const unsigned* k_values, unsigned n_k, size_t cbf_bytes, size_t bf_bytes,
              uint8_t* out_bfs)
{
  std::vector<unsigned> ks(k_values, k_values + n_k);
  std::vector<std::unique_ptr<btllib::KmerCountingBloomFilter8>> cbfs;
  std::vector<std::unique_ptr<btllib::KmerBloomFilter>> bfs;
  for (const auto k : ks) {
    cbfs.push_back(std::unique_ptr<btllib::KmerCountingBloomFilter8>(
      new btllib::KmerCountingBloomFilter8(cbf_bytes, hash_num, k)));
    bfs.push_back(
      std::unique_ptr<btllib::KmerBloomFilter>(new btllib::KmerBloomFilter(bf_bytes, hash_num, k)));
  }
  for (uint64_t r = 0; r < n_reads; ++r) {
    fill_bfs(seqs + off[r], off[r + 1] - off[r], hash_num, ks, thresholds[r], cbfs, bfs);
  }
  for (unsigned i = 0; i < n_k; ++i) {
    memcpy(out_bfs + (size_t)i * bf_bytes, bfs[i]->raw().data(), bf_bytes);
  }
  return 0;
}
