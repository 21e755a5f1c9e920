// TEST INFRASTRUCTURE ONLY (oracle/).  Stand-in for <btllib/util.hpp>: what the GoldPolish input
// readers call (seqindex.cpp:30-31,44,51; mappings.cpp:22-25).  RECALLED from btllib, PARITY
// UNPINNED: split() cuts at every occurrence of the delimiter (only token 0 is used by the callers),
// calc_phred_avg() is the arithmetic mean of (character - 33) over [start_pos, start_pos + len),
// len == 0 meaning "to the end".
#ifndef GRB_SHIM_POLISH_UTIL_HPP
#define GRB_SHIM_POLISH_UTIL_HPP
#include "status.hpp"
// the reference relies on btllib's headers pulling these in (std::ceil mappings.cpp:257,
// std::stringstream :150, std::tuple seqindex.hpp:47)
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>
namespace btllib {
inline std::vector<std::string>
split(const std::string& s, const std::string& delim)
{
  std::vector<std::string> tokens;
  size_t pos1 = 0, pos2 = 0;
  while ((pos2 = s.find(delim, pos2)) != std::string::npos) {
    tokens.push_back(s.substr(pos1, pos2 - pos1));
    pos2 += delim.size();
    pos1 = pos2;
  }
  tokens.push_back(s.substr(pos1));
  return tokens;
}
inline bool
endswith(const std::string& s, const std::string& suffix)
{
  return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}
inline double
calc_phred_avg(const std::string& qual, const size_t start_pos = 0, size_t len = 0)
{
  if (len == 0) {
    len = qual.size() - start_pos;
  }
  check_error(start_pos + len > qual.size(), "calc_phred_avg: start_pos + len > qual.size()");
  size_t phred_sum = 0;
  for (size_t i = start_pos; i < start_pos + len; ++i) {
    phred_sum += (size_t)(unsigned char)qual[i] - 33;
  }
  return (double)phred_sum / (double)len;
}
} // namespace btllib
#endif
