// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for sdsl-lite's <sdsl/bit_vector_il.hpp> and the
// interleaved rank support, so the reference sources compile unmodified into oracle/_ref/.
// Surface used by the reference: goldrush_path/MIBFConstructSupport.hpp:165-170 (conversion from
// bit_vector, rank_support_il<1>(&bv)), goldrush_path/MIBloomFilter.hpp:465-491,538-546
// (operator[], size(), rank(pos) = number of set bits in [0, pos)).
// Like sdsl, one cumulative 64-bit count is interleaved in front of every 512-bit block.
#ifndef GRB_SHIM_SDSL_BIT_VECTOR_IL_HPP
#define GRB_SHIM_SDSL_BIT_VECTOR_IL_HPP

#include "int_vector.hpp"

#include <string>

namespace sdsl {

template<uint8_t t_b>
class rank_support_il;

template<uint32_t t_bs = 512>
class bit_vector_il
{
public:
  static_assert(t_bs == 512, "stand-in implements the 512-bit block size the reference uses");
  bit_vector_il() = default;
  explicit bit_vector_il(const bit_vector& bv)
    : m_size(bv.size())
  {
    const size_t blocks = (m_size + 511) / 512 + 1;
    m_data.assign(blocks * 9, 0);
    uint64_t cum = 0;
    const uint64_t* w = bv.data();
    const size_t nwords = (m_size + 63) / 64;
    for (size_t b = 0; b < blocks; ++b) {
      m_data[b * 9] = cum;
      for (size_t j = 0; j < 8; ++j) {
        const size_t wi = b * 8 + j;
        uint64_t v = wi < nwords ? w[wi] : 0;
        if (wi + 1 == nwords && (m_size & 63)) {
          v &= (~0ULL) >> (64 - (m_size & 63));
        }
        m_data[b * 9 + 1 + j] = v;
        cum += (uint64_t)__builtin_popcountll(v);
      }
    }
  }

  size_t size() const { return m_size; }
  bool operator[](size_t i) const
  {
    return (m_data[(i >> 9) * 9 + 1 + ((i >> 6) & 7)] >> (i & 63)) & 1;
  }
  uint64_t rank1(size_t i) const
  {
    const uint64_t* blk = &m_data[(i >> 9) * 9];
    uint64_t r = blk[0];
    const size_t wj = (i >> 6) & 7;
    for (size_t j = 0; j < wj; ++j) {
      r += (uint64_t)__builtin_popcountll(blk[1 + j]);
    }
    if (i & 63) {
      r += (uint64_t)__builtin_popcountll(blk[1 + wj] & ((1ULL << (i & 63)) - 1));
    }
    return r;
  }
  bool store_to_file(const std::string&) const { return false; }

private:
  size_t m_size = 0;
  std::vector<uint64_t> m_data;
};

template<uint8_t t_b = 1>
class rank_support_il
{
public:
  static_assert(t_b == 1, "stand-in implements rank of 1-bits only");
  rank_support_il() = default;
  explicit rank_support_il(const bit_vector_il<512>* bv)
    : m_bv(bv)
  {}
  uint64_t operator()(size_t i) const { return m_bv->rank1(i); }
  uint64_t rank(size_t i) const { return m_bv->rank1(i); }

private:
  const bit_vector_il<512>* m_bv = nullptr;
};

// free function used only by the reference's development-only MIBloomFilter::store
// (goldrush_path/MIBloomFilter.hpp:152); never called on the hot path.
template<class V>
inline bool
store_to_file(const V&, const std::string&)
{
  return false;
}

} // namespace sdsl

#endif
