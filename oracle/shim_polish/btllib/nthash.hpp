// TEST INFRASTRUCTURE ONLY (oracle/).  Stand-in for btllib's <btllib/nthash.hpp>, the part the
// GoldPolish targeted-Bloom-filter builder binds: btllib::NtHash(seq, seq_len, hash_num, k),
// roll(), hashes()  (subprojects/goldpolish/src/utils.cpp:113-118).
//
// btllib is a third-party dependency that is not vendored and cannot be installed offline.  The
// arithmetic restated here is the one stated IN the reference tree, in the ntHash copy vendored
// with ntEdit (subprojects/goldpolish/subprojects/ntedit/lib/nthash.hpp):
//   seeds A / C / G / T                         :24-28
//   forward / reverse hash of the first k-mer   NTF64 / NTR64, :100-119 (rol1 + swapbits033 per base)
//   rolling update                               :122-131, :143-152
//   canonical = forward + reverse                NTC64, :172-191
//   extra hashes t = base * (i ^ k * multiSeed); t ^= t >> multiShift     NTMC64, :18-21,262-300
// What is NOT in the tree and is recalled from btllib: that roll() skips every k-mer holding a
// character outside ACGTacgt, and that hashes()[0] is the canonical hash.  PARITY UNPINNED there.
#ifndef GRB_SHIM_POLISH_BTLLIB_NTHASH_HPP
#define GRB_SHIM_POLISH_BTLLIB_NTHASH_HPP
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace btllib {

namespace polish_shim {
static const uint64_t kSeed[4] = { 0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL, 0x20323ed082572324ULL,
                                   0x295549f54be24456ULL };
static const uint64_t kMultiSeed = 0x90b45d39fb6da1faULL;
static const int kMultiShift = 27;
inline int
code(unsigned char c)
{
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}
// rol1 + swapbits033: the upper 31 and the lower 33 bits rotate left by one, independently
inline uint64_t
srol(uint64_t x)
{
  const uint64_t m = ((x & 0x8000000000000000ULL) >> 30) | ((x & 0x100000000ULL) >> 32);
  return ((x << 1) & 0xFFFFFFFDFFFFFFFFULL) | m;
}
inline uint64_t
srol(uint64_t x, unsigned d)
{
  const uint64_t hi = x >> 33, lo = x & 0x1FFFFFFFFULL;
  const unsigned dh = d % 31, dl = d % 33;
  const uint64_t h2 = dh ? ((hi << dh) | (hi >> (31 - dh))) & 0x7FFFFFFFULL : hi;
  const uint64_t l2 = dl ? ((lo << dl) | (lo >> (33 - dl))) & 0x1FFFFFFFFULL : lo;
  return (h2 << 33) | l2;
}
inline uint64_t
sror(uint64_t x)
{
  const uint64_t m = ((x & 0x200000000ULL) << 30) | ((x & 1ULL) << 32);
  return ((x >> 1) & 0xFFFFFFFEFFFFFFFFULL) | m;
}
} // namespace polish_shim

class NtHash
{
public:
  NtHash(const char* seq, size_t seq_len, unsigned hash_num, unsigned k)
    : seq_(seq), len_(seq_len), h_(hash_num), k_(k), hashes_(hash_num, 0)
  {
  }
  // next k-mer without a character outside ACGTacgt; false at the end of the sequence
  bool roll()
  {
    using namespace polish_shim;
    if (k_ == 0 || len_ < k_) {
      return false;
    }
    while (true) {
      if (!valid_) {
        // (re)start: first window at or after pos_ that holds only ACGT
        size_t start = started_ ? pos_ + 1 : 0;
        started_ = true;
        while (true) {
          if (start + k_ > len_) {
            return false;
          }
          size_t bad = k_;
          for (size_t j = k_; j-- > 0;) {
            if (code((unsigned char)seq_[start + j]) < 0) {
              bad = j;
              break;
            }
          }
          if (bad == k_) {
            break;
          }
          start += bad + 1;
        }
        pos_ = start;
        fh_ = 0;
        rh_ = 0;
        for (unsigned i = 0; i < k_; ++i) {
          fh_ = srol(fh_) ^ kSeed[code((unsigned char)seq_[pos_ + i])];
          rh_ = srol(rh_) ^ kSeed[3 - code((unsigned char)seq_[pos_ + k_ - 1 - i])];
        }
        valid_ = true;
      } else {
        if (pos_ + k_ >= len_) {
          return false;
        }
        const int in = code((unsigned char)seq_[pos_ + k_]);
        if (in < 0) {
          pos_ = pos_ + k_; // every window touching the bad character is skipped
          valid_ = false;
          continue;
        }
        const int out = code((unsigned char)seq_[pos_]);
        fh_ = srol(fh_) ^ kSeed[in] ^ srol(kSeed[out], k_);
        rh_ = sror(rh_ ^ kSeed[3 - out] ^ srol(kSeed[3 - in], k_));
        ++pos_;
      }
      const uint64_t base = fh_ + rh_;
      hashes_[0] = base;
      for (unsigned i = 1; i < h_; ++i) {
        uint64_t t = base * (i ^ k_ * kMultiSeed);
        t ^= t >> kMultiShift;
        hashes_[i] = t;
      }
      return true;
    }
  }
  const uint64_t* hashes() const { return hashes_.data(); }
  size_t get_pos() const { return pos_; }

private:
  const char* seq_;
  size_t len_;
  unsigned h_, k_;
  std::vector<uint64_t> hashes_;
  size_t pos_ = 0;
  uint64_t fh_ = 0, rh_ = 0;
  bool valid_ = false, started_ = false;
};

} // namespace btllib
#endif
