// TEST INFRASTRUCTURE ONLY.  CPU restatement (flat arrays, no btllib types) of what the GoldPolish
// targeted-Bloom-filter builder computes per batch -- SURVEY.md 8(f4):
//   serve_batch          subprojects/goldpolish/src/goldpolish_targeted_bfs.cpp:53-149
//   k-mer threshold      :43-51  (mappings_bases_to_kmer_threshold)
//   fill_bfs             subprojects/goldpolish/src/utils.cpp:96-123
// and of the btllib pieces it binds, which are NOT in the tree (parity unpinned there, see
// oracle/shim_polish/btllib/*.hpp for what is stated in-tree and what is recalled):
//   NtHash               k-mer ntHash, hash_num hashes (ntedit/lib/nthash.hpp:24-28,100-191,262-300)
//   KmerCountingBloomFilter8::insert_thresh_contains, KmerBloomFilter::insert
// Checked against the reference's own fill_bfs (oracle/_ref/libgoldpolish_ref.so) by
// tests/test_polish.py.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <tuple>
#include <vector>

namespace {

const uint64_t kSeed[4] = { 0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL, 0x20323ed082572324ULL,
                            0x295549f54be24456ULL };

inline int
code(unsigned char c)
{
  switch (c & 0xDF) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
  }
}

// split rotation by d: the upper 31 and the lower 33 bits rotate left independently (rol1 +
// swapbits033, nthash.hpp:66-92)
inline uint64_t
srol(uint64_t x, unsigned d)
{
  const uint64_t hi = x >> 33, lo = x & 0x1FFFFFFFFULL;
  const unsigned dh = d % 31, dl = d % 33;
  const uint64_t h2 = dh ? ((hi << dh) | (hi >> (31 - dh))) & 0x7FFFFFFFULL : hi;
  const uint64_t l2 = dl ? ((lo << dl) | (lo >> (33 - dl))) & 0x1FFFFFFFFULL : lo;
  return (h2 << 33) | l2;
}

// canonical hash of the k-mer at seq[p, p + k), from scratch (NTC64, nthash.hpp:172-178)
uint64_t
kmer_hash(const char* seq, size_t p, unsigned k)
{
  uint64_t f = 0, r = 0;
  for (unsigned j = 0; j < k; ++j) {
    f ^= srol(kSeed[code((unsigned char)seq[p + j])], k - 1 - j);
    r ^= srol(kSeed[3 - code((unsigned char)seq[p + j])], j);
  }
  return f + r;
}

} // namespace

extern "C" {

// goldpolish_targeted_bfs.cpp:43-51
int
grbo_polish_kmer_threshold(uint64_t mappings_bases)
{
  const double a = 4.66943, b = 2.11391e-07;
  const int t = int(std::round(a + double(mappings_bases) * b));
  return std::min(t, 13);
}

// goldpolish_targeted_bfs.cpp:88-127: which mapped reads of one target are used, in which order,
// and the k-mer threshold they are inserted with.  order_out[n_mappings]; returns the number used.
uint32_t
grbo_polish_plan_target(uint64_t target_len, double subsample_max_per_10kbp, uint32_t n_mappings,
                        const char* const* ids, const double* phred_avg, const uint64_t* lens,
                        uint32_t* order_out, int32_t* kmer_threshold)
{
  const size_t num_max = size_t(double(target_len) * subsample_max_per_10kbp / 10'000.0);
  const size_t adjusted = std::min<size_t>(n_mappings, num_max);
  std::vector<std::tuple<std::string, size_t, uint32_t>> v;
  for (uint32_t i = 0; i < n_mappings; ++i) {
    v.emplace_back(ids[i], size_t(phred_avg[i]), i); // the tuple holds the Phred average as size_t (:104-108)
  }
  std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) {
    return (std::get<1>(a) > std::get<1>(b)) ||
           (std::get<1>(a) == std::get<1>(b) && std::get<0>(a) < std::get<0>(b));
  });
  uint64_t bases = 0;
  for (size_t i = 0; i < v.size(); ++i) {
    order_out[i] = std::get<2>(v[i]);
    if (i < adjusted) {
      bases += lens[std::get<2>(v[i])];
    }
  }
  *kmer_threshold = grbo_polish_kmer_threshold(bases);
  return (uint32_t)adjusted;
}

// fill_bfs over reads [0, n_reads) in order (utils.cpp:96-123), filters set up as in serve_batch
// (goldpolish_targeted_bfs.cpp:68-77); out_bfs[n_k * bf_bytes]
int
grbo_polish_fill(const char* seqs, const uint64_t* off, const uint32_t* thresholds, uint64_t n_reads,
                 unsigned hash_num, const unsigned* k_values, unsigned n_k, size_t cbf_bytes, size_t bf_bytes,
                 uint8_t* out_bfs)
{
  memset(out_bfs, 0, (size_t)n_k * bf_bytes);
  std::vector<uint64_t> h(hash_num);
  for (unsigned ki = 0; ki < n_k; ++ki) {
    const unsigned k = k_values[ki];
    std::vector<uint8_t> cbf(cbf_bytes, 0);
    uint8_t* bf = out_bfs + (size_t)ki * bf_bytes;
    const uint64_t bf_bits = (uint64_t)bf_bytes * 8;
    for (uint64_t r = 0; r < n_reads; ++r) {
      if (thresholds[r] < 4) {
        return -1; // utils.cpp:105-107
      }
      const unsigned thr = thresholds[r] - 2 + ki; // utils.cpp:108,121: one more per k value
      const char* seq = seqs + off[r];
      const size_t len = off[r + 1] - off[r];
      size_t run = 0; // valid characters ending at i
      for (size_t i = 0; i < len; ++i) {
        run = code((unsigned char)seq[i]) < 0 ? 0 : run + 1;
        if (run < k) {
          continue;
        }
        const uint64_t base = kmer_hash(seq, i + 1 - k, k);
        h[0] = base;
        for (unsigned j = 1; j < hash_num; ++j) {
          uint64_t t = base * (j ^ k * 0x90b45d39fb6da1faULL);
          t ^= t >> 27;
          h[j] = t;
        }
        uint8_t count = 255;
        for (unsigned j = 0; j < hash_num; ++j) {
          count = std::min(count, cbf[h[j] % cbf_bytes]);
        }
        uint8_t after = count;
        if (count < (uint8_t)std::min(thr, 255u)) {
          for (unsigned j = 0; j < hash_num; ++j) {
            uint8_t& c = cbf[h[j] % cbf_bytes];
            if (c == count) {
              c = (uint8_t)(count + 1);
            }
          }
          after = (uint8_t)(count + 1);
        }
        if (after >= thr) {
          for (unsigned j = 0; j < hash_num; ++j) {
            const uint64_t pos = h[j] % bf_bits;
            bf[pos / 8] |= (uint8_t)(1u << (pos % 8));
          }
        }
      }
    }
  }
  return 0;
}

} // extern "C"
