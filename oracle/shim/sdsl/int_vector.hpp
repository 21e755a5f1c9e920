// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for sdsl-lite's <sdsl/int_vector.hpp>, so the
// reference sources compile unmodified into oracle/_ref/. Surface used by the reference:
//   goldrush_path/MIBFConstructSupport.hpp:63,83,141-143,168 (bit_vector(n), size(), data()).
// sdsl::bit_vector(n) is n zero bits stored LSB-first in 64-bit words.
#ifndef GRB_SHIM_SDSL_INT_VECTOR_HPP
#define GRB_SHIM_SDSL_INT_VECTOR_HPP

#include <cstddef>
#include <cstdint>
#include <vector>

namespace sdsl {

class bit_vector
{
public:
  bit_vector() = default;
  explicit bit_vector(size_t n)
    : m_size(n)
    , m_words((n + 63) / 64 + 1, 0)
  {}
  size_t size() const { return m_size; }
  uint64_t* data() { return m_words.data(); }
  const uint64_t* data() const { return m_words.data(); }
  bool operator[](size_t i) const { return (m_words[i >> 6] >> (i & 63)) & 1; }

private:
  size_t m_size = 0;
  std::vector<uint64_t> m_words;
};

} // namespace sdsl

#endif
