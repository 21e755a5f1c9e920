// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for btllib's <btllib/nthash.hpp>.
//
// btllib (>= 1.6.2 per /root/reference/requirements.txt:8) is a third-party dependency that is
// NOT vendored under /root/reference and cannot be installed offline. This header restates the
// published spaced-seed ntHash that the reference binds at
//   goldrush_path/multiLensfrHashIterator.hpp:39-41,54,60   (SeedNtHash ctor / roll / hashes)
// so that the reference's own sources compile unmodified into oracle/_ref/.
//
// Algorithm restated (btllib nthash_seed / ntHash2 "ntmsm64"):
//   per-base 64-bit seeds A/C/G/T        = constants also stated in-tree at
//       subprojects/goldpolish/subprojects/ntedit/lib/nthash.hpp:24-28
//   split rotation "srol" (upper 31 bits and lower 33 bits rotate independently)
//       = rol1 + swapbits033 at nthash.hpp:66-92
//   forward  hash = XOR over care positions p of srol^(k-1-p)(seed[base_p])
//   reverse  hash = XOR over care positions p of srol^(p)(seed[complement(base_p)])
//       = the per-position masks at nthash.hpp:529-563 (NTMS64)
//   canonical     = forward + reverse            (nthash.hpp:172-191 "fhVal + rhVal")
// PARITY UNPINNED at this boundary: equality with the real btllib binary cannot be checked
// offline; the hash is isolated here (oracle/_ref), in oracle/grb_oracle.cpp and in
// goldrush_b200/csrc/nthash.cuh so all three can be corrected together.
//
// Rolling is O(#blocks) per step (table driven) so that the CPU baseline is not handicapped.
#ifndef GRB_SHIM_BTLLIB_NTHASH_HPP
#define GRB_SHIM_BTLLIB_NTHASH_HPP

// The reference relies on btllib's headers pulling these in transitively
// (std::minstd_rand in MIBloomFilter.hpp:222, std::max_element in read_hashing.cpp:112).
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

namespace btllib {

namespace shim_detail {

static const uint64_t kSeedA = 0x3c8bfbb395c60474ULL;
static const uint64_t kSeedC = 0x3193c18562a02b4cULL;
static const uint64_t kSeedG = 0x20323ed082572324ULL;
static const uint64_t kSeedT = 0x295549f54be24456ULL;

// 0..3 = A,C,G,T ; 4 = anything else
inline unsigned
base_code(unsigned char c)
{
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}

inline uint64_t
srol1(uint64_t x)
{
  const uint64_t carry = ((x & 0x8000000000000000ULL) >> 30) | ((x & 0x100000000ULL) >> 32);
  return ((x << 1) & 0xFFFFFFFDFFFFFFFFULL) | carry;
}

inline uint64_t
sror1(uint64_t x)
{
  const uint64_t carry = ((x & 0x200000000ULL) << 30) | ((x & 1ULL) << 32);
  return ((x >> 1) & 0xFFFFFFFEFFFFFFFFULL) | carry;
}

// rot_tab[base][r] = srol^r(seed[base]); period of the split rotation is lcm(31,33) = 1023.
struct RotTables
{
  std::vector<std::array<uint64_t, 4>> fwd; // by rotation
  explicit RotTables(unsigned max_rot)
    : fwd(max_rot + 1)
  {
    uint64_t cur[4] = { kSeedA, kSeedC, kSeedG, kSeedT };
    for (unsigned r = 0; r <= max_rot; ++r) {
      for (unsigned b = 0; b < 4; ++b) {
        fwd[r][b] = cur[b];
        cur[b] = srol1(cur[b]);
      }
    }
  }
};

} // namespace shim_detail

class SeedNtHash
{
public:
  SeedNtHash(const std::string& seq,
             const std::vector<std::string>& seeds,
             unsigned hashes_per_seed,
             unsigned k,
             size_t pos = 0)
    : m_seq(seq.data())
    , m_len(seq.size())
    , m_k(k)
    , m_pos(pos)
    , m_tabs(std::make_shared<shim_detail::RotTables>(k))
  {
    if (hashes_per_seed != 1 || seeds.size() != 1) {
      std::cerr << "SeedNtHash stand-in: only one seed / one hash per seed is supported"
                << std::endl;
      std::exit(EXIT_FAILURE);
    }
    if (seeds[0].size() != k) {
      std::cerr << "SeedNtHash: spaced seed string length (" << seeds[0].size()
                << ") not equal to k=" << k << std::endl;
      std::exit(EXIT_FAILURE);
    }
    if (m_len < m_k) {
      std::cerr << "SeedNtHash: sequence length (" << m_len << ") is smaller than k (" << m_k
                << ")" << std::endl;
      std::exit(EXIT_FAILURE);
    }
    const std::string& s = seeds[0];
    for (unsigned i = 0; i < k; ++i) {
      if (s[i] == '1') {
        m_care.push_back(i);
        if (i == 0 || s[i - 1] != '1') {
          m_block_start.push_back(i);
        }
        if (i + 1 == k || s[i + 1] != '1') {
          m_block_end.push_back(i + 1);
        }
      }
    }
  }

  bool roll()
  {
    if (!m_initialized) {
      return init();
    }
    if (m_pos >= m_len - m_k) {
      return false;
    }
    if (shim_detail::base_code((unsigned char)m_seq[m_pos + m_k]) > 3) {
      m_pos += m_k;
      return init();
    }
    const auto& tab = m_tabs->fwd;
    uint64_t f = m_fwd;
    uint64_t r = m_rev;
    const size_t nb = m_block_start.size();
    for (size_t b = 0; b < nb; ++b) {
      const unsigned s = m_block_start[b];
      const unsigned c = shim_detail::base_code((unsigned char)m_seq[m_pos + s]);
      f ^= tab[m_k - 1 - s][c];
      r ^= tab[s][3 - c];
    }
    f = shim_detail::srol1(f);
    r = shim_detail::sror1(r);
    for (size_t b = 0; b < nb; ++b) {
      const unsigned e = m_block_end[b];
      const unsigned c = shim_detail::base_code((unsigned char)m_seq[m_pos + e]);
      f ^= tab[m_k - e][c];
      r ^= tab[e - 1][3 - c];
    }
    m_fwd = f;
    m_rev = r;
    m_hash = f + r;
    ++m_pos;
    return true;
  }

  const uint64_t* hashes() const { return &m_hash; }
  size_t get_pos() const { return m_pos; }

private:
  // first window at or after m_pos whose span holds only A/C/G/T
  bool init()
  {
    while (m_pos + m_k <= m_len) {
      bool ok = true;
      for (unsigned i = m_k; i-- > 0;) {
        if (shim_detail::base_code((unsigned char)m_seq[m_pos + i]) > 3) {
          m_pos += i + 1;
          ok = false;
          break;
        }
      }
      if (!ok) {
        continue;
      }
      const auto& tab = m_tabs->fwd;
      uint64_t f = 0;
      uint64_t r = 0;
      for (unsigned p : m_care) {
        const unsigned c = shim_detail::base_code((unsigned char)m_seq[m_pos + p]);
        f ^= tab[m_k - 1 - p][c];
        r ^= tab[p][3 - c];
      }
      m_fwd = f;
      m_rev = r;
      m_hash = f + r;
      m_initialized = true;
      return true;
    }
    return false;
  }

  const char* m_seq;
  size_t m_len;
  unsigned m_k;
  size_t m_pos;
  bool m_initialized = false;
  uint64_t m_fwd = 0;
  uint64_t m_rev = 0;
  uint64_t m_hash = 0;
  std::vector<unsigned> m_care;
  std::vector<unsigned> m_block_start;
  std::vector<unsigned> m_block_end;
  std::shared_ptr<shim_detail::RotTables> m_tabs;
};

} // namespace btllib

#endif
