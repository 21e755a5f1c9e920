// TEST INFRASTRUCTURE ONLY — see grb_oracle.h.  CPU restatement of GoldRush-Path's read-selection
// loop on flat arrays.  Citations are file:line under the reference tree (goldrush_path/ unless
// another directory is named).  No line of the reference is copied: the control flow is restated
// from its observable semantics and pinned against oracle/_ref (the unmodified reference sources).
#include "grb_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <getopt.h>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <unordered_set>
#include <vector>

#if _OPENMP
#include <omp.h>
#endif

namespace {

// ----------------------------------------------------------------------------------------------
// Spaced-seed ntHash.  btllib::SeedNtHash is third-party and absent from the reference tree; the
// call sites are multiLensfrHashIterator.hpp:39-41,54,60.  Constants and rotations follow the
// in-tree statement of ntHash: subprojects/goldpolish/subprojects/ntedit/lib/nthash.hpp:24-28
// (seeds), :66-92 (rol1 + swapbits033 = rotate the upper 31 and lower 33 bits separately),
// :529-563 (forward care position p contributes the base seed rotated by k-1-p, reverse the
// complement seed rotated by p), :172-191 (canonical = forward + reverse).
// Computed directly per window (no rolling) so that it is independent of the rolling stand-in used
// to build oracle/_ref.
// ----------------------------------------------------------------------------------------------
const uint64_t kSeed[4] = { 0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL, 0x20323ed082572324ULL,
                            0x295549f54be24456ULL };

inline uint64_t
split_rotl(uint64_t x, unsigned r)
{
  const uint64_t hi = x >> 33;               // 31 bits
  const uint64_t lo = x & 0x1FFFFFFFFULL;    // 33 bits
  const unsigned rh = r % 31, rl = r % 33;
  const uint64_t nh = rh ? (((hi << rh) | (hi >> (31 - rh))) & 0x7FFFFFFFULL) : hi;
  const uint64_t nl = rl ? (((lo << rl) | (lo >> (33 - rl))) & 0x1FFFFFFFFULL) : lo;
  return (nh << 33) | nl;
}

inline int
base_code(unsigned char c)
{
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}

struct SeedTable
{
  unsigned span = 0;
  std::vector<unsigned> care;
  std::vector<uint64_t> fwd; // [care_idx*4 + base]
  std::vector<uint64_t> rev; // [care_idx*4 + base]  (already complemented)
  explicit SeedTable(const std::string& s)
  {
    span = (unsigned)s.size();
    for (unsigned i = 0; i < span; ++i) {
      if (s[i] == '1') {
        care.push_back(i);
      }
    }
    fwd.resize(care.size() * 4);
    rev.resize(care.size() * 4);
    for (size_t j = 0; j < care.size(); ++j) {
      for (unsigned b = 0; b < 4; ++b) {
        fwd[j * 4 + b] = split_rotl(kSeed[b], span - 1 - care[j]);
        rev[j * 4 + b] = split_rotl(kSeed[3 - b], care[j]);
      }
    }
  }
  inline uint64_t hash(const uint8_t* codes) const
  {
    uint64_t f = 0, r = 0;
    for (size_t j = 0; j < care.size(); ++j) {
      const unsigned b = codes[care[j]];
      f ^= fwd[j * 4 + b];
      r ^= rev[j * 4 + b];
    }
    return f + r;
  }
};

// One btllib::SeedNtHash stream: positions of successive valid windows.  A window is valid when
// its whole span holds A/C/G/T only (restated btllib behaviour, unpinned; only reachable through
// --ntcard because passes 1-2 drop reads with other bytes, goldrush_path.cpp:293-301).
struct Stream
{
  const SeedTable* t;
  const uint8_t* codes; // 0..3, 4 = other
  size_t len;
  // next_bad[i] = smallest j >= i with codes[j] > 3 (len if none)
  const uint32_t* next_bad;
  size_t pos = 0;
  bool started = false;
  uint64_t value = 0;
  bool seek()
  {
    while (pos + t->span <= len) {
      const size_t nb = next_bad ? next_bad[pos] : len;
      if (nb >= pos + t->span) {
        value = t->hash(codes + pos);
        return true;
      }
      pos = nb + 1;
    }
    return false;
  }
  bool roll()
  {
    if (!started) {
      started = true;
      return seek();
    }
    if (pos + t->span >= len) {
      return false;
    }
    ++pos;
    return seek();
  }
};

// multiLensfrHashIterator.hpp:29-68: h streams of different spans advance together; a stream that
// cannot advance keeps its last value; iteration ends when no stream advanced.
template<typename F>
size_t
for_each_frame(const std::vector<SeedTable>& seeds, const uint8_t* codes, size_t len,
               const uint32_t* next_bad, F&& emit)
{
  const size_t h = seeds.size();
  std::vector<Stream> st(h);
  std::vector<uint64_t> cur(h, 0);
  for (size_t i = 0; i < h; ++i) {
    st[i].t = &seeds[i];
    st[i].codes = codes;
    st[i].len = len;
    st[i].next_bad = next_bad;
    st[i].roll();
    cur[i] = st[i].value;
  }
  size_t frames = 0;
  while (true) {
    emit(cur.data());
    ++frames;
    bool update = false;
    for (size_t i = 0; i < h; ++i) {
      if (st[i].roll()) {
        update = true;
        cur[i] = st[i].value;
      }
    }
    if (!update) {
      break;
    }
  }
  return frames;
}

void
encode(const char* seq, size_t n, std::vector<uint8_t>& codes, std::vector<uint32_t>& next_bad,
       bool& any_bad)
{
  codes.resize(n);
  any_bad = false;
  for (size_t i = 0; i < n; ++i) {
    const int c = base_code((unsigned char)seq[i]);
    codes[i] = c < 0 ? 4 : (uint8_t)c;
    any_bad |= c < 0;
  }
  next_bad.clear();
  if (any_bad) {
    next_bad.resize(n + 1);
    next_bad[n] = (uint32_t)n;
    for (size_t i = n; i-- > 0;) {
      next_bad[i] = codes[i] > 3 ? (uint32_t)i : next_bad[i + 1];
    }
  }
}

// ----------------------------------------------------------------------------------------------
// calc_phred_average.cpp:8-43 / :45-58
// ----------------------------------------------------------------------------------------------
inline double
delog(char q)
{
  const int phred_score = (int)(q - 33);
  return pow(10.0, -phred_score / 10.0);
}

void
phred_average(const char* qual, size_t n, uint32_t& avg, uint32_t& delta, double* sums)
{
  double total = 0.0, first = 0.0;
  for (size_t i = 0; i < n; ++i) {
    total += delog(qual[i]);
    if (i == n / 2 - 1) {
      first = total;
    }
  }
  if (sums) {
    sums[0] = first;
    sums[1] = total;
  }
  double second = total - first;
  second = second / (n * 0.5);
  const double first_avg = first / (n * 0.5);
  avg = (uint32_t)(-10 * log10(total / n));
  delta = (uint32_t)abs((int32_t)(-10 * log10(first_avg)) - (int32_t)(-10 * log10(second)));
}

// ----------------------------------------------------------------------------------------------
// The filter: MIBFConstructSupport (bit vector, counts) + MIBloomFilter (rank -> ID).
// ----------------------------------------------------------------------------------------------
struct Filter
{
  uint64_t bits = 0;
  unsigned h = 0;
  std::vector<uint64_t> words;   // sdsl::bit_vector layout
  std::vector<uint64_t> cum;     // set bits before each 64-bit word
  std::vector<uint32_t> data;    // MIBloomFilter::m_data
  std::vector<uint32_t> counts;  // MIBFConstructSupport::m_counts
  bool ready = false;

  // MIBFConstructSupport.hpp:134-147
  inline void set_bit(uint64_t hash)
  {
    const uint64_t pos = hash % bits;
    __sync_fetch_and_or(&words[pos >> 6], (uint64_t)1 << (pos & 63));
  }
  // MIBFConstructSupport.hpp:165-181, MIBloomFilter.hpp:165-184,538-546
  uint64_t setup()
  {
    cum.resize(words.size() + 1);
    uint64_t c = 0;
    for (size_t i = 0; i < words.size(); ++i) {
      cum[i] = c;
      c += (uint64_t)__builtin_popcountll(words[i]);
    }
    cum[words.size()] = c;
    data.assign(c, 0);
    counts.assign(c, 0);
    ready = true;
    return c;
  }
  inline bool bit(uint64_t pos) const { return (words[pos >> 6] >> (pos & 63)) & 1; }
  inline uint64_t rank(uint64_t pos) const
  {
    const uint64_t w = words[pos >> 6];
    const unsigned b = pos & 63;
    return cum[pos >> 6] + (b ? (uint64_t)__builtin_popcountll(w & (((uint64_t)1 << b) - 1)) : 0);
  }
};

const uint32_t kSatMask = 1u << 31;       // MIBloomFilter.hpp:38
const uint32_t kAntiMask = ~kSatMask;     // :39

struct TileVote
{
  uint32_t best_id = 0;
  uint32_t best_count = 0;
  std::vector<std::pair<uint32_t, uint32_t>> cands; // (id, count) with count > 2
};

// goldrush_path.cpp:547-626 for one tile; hashes[frame*h + p]
void
vote_tile(const Filter& f, const uint64_t* hashes, size_t frames, TileVote& out,
          uint64_t* counters)
{
  const unsigned h = f.h;
  std::vector<uint32_t> seen;
  seen.reserve(frames * h);
  uint64_t hits = 0, misses = 0;
  uint64_t ranks[16];
  uint32_t uniq[16];
  for (size_t fr = 0; fr < frames; ++fr) {
    const uint64_t* hv = hashes + fr * h;
    bool all = true; // MIBloomFilter.hpp:465-476 atRank: fail fast on the first clear bit
    for (unsigned p = 0; p < h; ++p) {
      const uint64_t pos = hv[p] % f.bits;
      if (!f.bit(pos)) {
        all = false;
        break;
      }
      ranks[p] = f.rank(pos);
    }
    if (!all) {
      continue;
    }
    unsigned nu = 0;
    for (unsigned p = 0; p < h; ++p) {
      uint32_t v = f.data[ranks[p]];
      if (v > kSatMask) {
        v &= kAntiMask; // goldrush_path.cpp:574-583
      }
      if (v == 0) {
        ++misses;
        continue;
      }
      ++hits;
      bool dup = false; // std::set per frame, :570,583,592
      for (unsigned j = 0; j < nu; ++j) {
        dup |= uniq[j] == v;
      }
      if (!dup) {
        uniq[nu++] = v;
      }
    }
    for (unsigned j = 0; j < nu; ++j) {
      seen.push_back(uniq[j]);
    }
  }
  std::sort(seen.begin(), seen.end());
  out.best_id = 0;
  out.best_count = 0;
  out.cands.clear();
  for (size_t i = 0; i < seen.size();) {
    size_t j = i;
    while (j < seen.size() && seen[j] == seen[i]) {
      ++j;
    }
    const uint32_t c = (uint32_t)(j - i);
    if (c > out.best_count) { // ascending ids + strict '>' = smallest id among ties (:610-615)
      out.best_count = c;
      out.best_id = seen[i];
    }
    if (c > 2) { // :616
      out.cands.emplace_back(seen[i], c);
    }
    i = j;
  }
  std::stable_sort(out.cands.begin(), out.cands.end(),
                   [](const auto& a, const auto& b) { return a.second > b.second; }); // :622
  if (counters) {
    __sync_fetch_and_add(&counters[0], (uint64_t)frames);
    __sync_fetch_and_add(&counters[1], hits);
    __sync_fetch_and_add(&counters[2], misses);
  }
}

// MIBFConstructSupport.hpp:247-283 on the hashes of tiles [start,end) laid end to end
void
insert_mibf(Filter& f, const uint64_t* hashes, size_t n, uint32_t id)
{
  std::vector<uint64_t> ranks(n);
  for (size_t i = 0; i < n; ++i) {
    ranks[i] = f.rank(hashes[i] % f.bits); // getRankPos: no bit test (MIBloomFilter.hpp:488-491)
  }
  std::sort(ranks.begin(), ranks.end());
  ranks.erase(std::unique(ranks.begin(), ranks.end()), ranks.end());
  for (const uint64_t rank : ranks) {
    const uint32_t count = ++f.counts[rank];
    const uint32_t seed32 = (uint32_t)(rank ^ (uint64_t)id); // std::hash<uint32_t> of a uint64 (:274-276)
    if (seed32 % count == count - 1) {
      uint32_t v = id; // setData, MIBloomFilter.hpp:593-602
      if (f.data[rank] > kSatMask) {
        v |= kSatMask;
      }
      f.data[rank] = v;
    }
  }
}

// goldrush_path.cpp:628-889
size_t
smooth_tiles(size_t n, uint32_t* id, uint8_t* as, const std::vector<TileVote>& votes,
             uint64_t threshold)
{
  for (size_t i = 0; i < n; ++i) { // :628-634
    if (!votes[i].cands.empty() && votes[i].cands[0].second > threshold) {
      as[i] = 1;
    }
  }
  if (n >= 3) {
    auto adopt = [&](size_t i, uint32_t nb) { // :649-659
      if (id[i] != nb) {
        for (const auto& c : votes[i].cands) {
          if (c.first == nb) {
            id[i] = nb;
            as[i] = c.second > threshold ? 1 : 0;
          }
        }
      }
    };
    for (size_t i = 1; i < n; ++i) { // :646-661
      adopt(i, id[i - 1]);
    }
    for (size_t i = n - 1; i-- > 0;) { // :667-682
      adopt(i, id[i + 1]);
    }
    auto fill = [&](size_t i) { // :695-709 ; all +-1 in uint32 arithmetic
      if (as[i]) {
        return;
      }
      const uint32_t c = id[i], p = id[i - 1], q = id[i + 1];
      const bool pa = as[i - 1], qa = as[i + 1];
      if ((c == p && pa) || (c == q && qa)) {
        as[i] = 1;
      } else if ((c == (uint32_t)(p + 1) && pa) || (c == (uint32_t)(q + 1) && qa)) {
        as[i] = 1;
      } else if ((c == (uint32_t)(p - 1) && pa) || (c == (uint32_t)(q - 1) && qa)) {
        as[i] = 1;
      } else if (p == q && pa && qa) {
        as[i] = as[i - 1];
        id[i] = p;
      }
    };
    for (size_t i = 1; i + 1 < n; ++i) { // :688-710
      fill(i);
    }
    for (size_t i = n - 2; i >= 1; --i) { // :712-734
      fill(i);
    }
    { // :739-766 bridge unassigned runs whose flanks agree
      std::vector<std::pair<size_t, size_t>> runs;
      size_t s = 0;
      for (size_t i = 1; i + 1 < n; ++i) {
        if (!as[i] && as[i - 1]) {
          s = i;
        } else if (as[i] && !as[i - 1]) {
          runs.emplace_back(s, i - 1);
        }
      }
      for (const auto& r : runs) {
        if (r.first == 0 || r.second == n - 1) {
          continue;
        }
        const uint32_t left = id[r.first - 1], right = id[r.second + 1];
        if (left == right || left == (uint32_t)(right + 1) || left == (uint32_t)(right - 1)) {
          for (size_t i = r.first; i <= r.second; ++i) {
            as[i] = 1;
            id[i] = left;
          }
        }
      }
    }
    // :771-793 isolated assigned tiles, forward then backward
    for (size_t i = 2; i + 2 < n; ++i) {
      if (as[i] && !as[i - 1] && !as[i + 1]) {
        as[i] = 0;
      }
    }
    for (size_t i = n - 3; i >= 2 && i < n; --i) {
      if (as[i] && !as[i - 1] && !as[i + 1]) {
        as[i] = 0;
      }
    }
    { // :799-822 per-id gap fill, ids visited in ascending order, index lists taken up front
      std::map<uint32_t, std::vector<uint32_t>> where;
      for (size_t i = 0; i < n; ++i) {
        if (as[i]) {
          where[id[i]].push_back((uint32_t)i);
        }
      }
      for (auto& kv : where) {
        const auto& idx = kv.second;
        for (size_t j = 1; j < idx.size(); ++j) {
          if (idx[j] > idx[j - 1] + 1) {
            const uint32_t v = id[idx[j - 1]];
            for (size_t t = idx[j - 1] + 1; t <= idx[j]; ++t) {
              id[t] = v;
            }
          }
        }
      }
    }
    { // :827-838 end tiles; this block alone compares in size_t (no 32-bit wrap)
      const size_t last = id[n - 1], last2 = id[n - 2], first = id[0], first2 = id[1];
      if (last == last2 || last == last2 + 1 || last == last2 - 1) {
        as[n - 1] = 1;
      }
      if (first == first2 || first == first2 + 1 || first == first2 - 1) {
        as[0] = 1;
      }
    }
    for (size_t i = 1; i + 1 < n; ++i) { // :840-850
      const uint32_t c = id[i], p = id[i - 1], q = id[i + 1];
      if (c != q && c != (uint32_t)(q - 1) && c != (uint32_t)(q + 1) && c != p &&
          c != (uint32_t)(p - 1) && c != (uint32_t)(p + 1)) {
        as[i] = 0;
      }
    }
    { // :856-877 assigned runs of at most 5 tiles
      std::vector<std::pair<size_t, size_t>> runs;
      size_t s = 0;
      for (size_t i = 1; i + 1 < n; ++i) {
        if (as[i] && !as[i - 1]) {
          s = i;
        } else if (!as[i] && as[i - 1]) {
          runs.emplace_back(s, i - 1);
        }
      }
      for (const auto& r : runs) {
        if (r.second - r.first + 1 <= 5) {
          for (size_t i = r.first; i <= r.second; ++i) {
            as[i] = 0;
          }
        }
      }
    }
  }
  size_t assigned = 0; // :883-889
  for (size_t i = 0; i < n; ++i) {
    assigned += as[i] ? 1 : 0;
  }
  return assigned;
}

// goldrush_path.cpp:195-233
void
find_longest_stretch(const uint8_t* as, size_t n, int64_t& ls, int64_t& le)
{
  size_t start = 0, end = 0, cur = 0, best = 0;
  ls = 0;
  le = 0;
  for (size_t i = 1; i + 1 < n; ++i) {
    const bool a = as[i], p = as[i - 1];
    if (!a && p) {
      start = i;
      cur = 1;
    } else if (!a && !p && i + 1 != n - 1) {
      ++cur;
    } else if (a && !p) {
      end = i - 1;
      if (best < cur) {
        best = cur;
        ls = (int64_t)start;
        le = (int64_t)end;
      }
    } else if (i + 1 == n - 1 && end < start) {
      end = i;
      ++cur;
      if (best < cur) {
        best = cur;
        ls = (int64_t)start;
        le = (int64_t)end;
      }
    }
  }
}

// goldrush_path.cpp:341-527.  The "two neighbouring ids" clauses (:394-398, :430-434, :475-478,
// :514-517) can never fire (they need the top count < 2, hence a pair sum of 2, to exceed 3).
bool
eval_flanks(int64_t ls, int64_t le, const uint32_t* id, size_t n, uint64_t& trim_start,
            uint64_t& trim_end)
{
  auto top_count = [&](int64_t lo, int64_t hi) -> size_t { // inclusive tile range
    std::map<uint32_t, size_t> m;
    size_t best = 0;
    for (int64_t i = lo; i <= hi; ++i) {
      best = std::max(best, ++m[id[i]]);
    }
    return best;
  };
  trim_start = ls != 0 ? (uint64_t)(ls - 1) : (uint64_t)ls;
  trim_end = (uint64_t)(le + 1);
  bool good = false;
  if (n < 15) {
    bool gl = false, gr = false;
    if (ls - 1 >= 0 && top_count(0, ls - 1) >= 2) {
      gl = true;
    }
    if (trim_start == 0) {
      gl = true;
    }
    if (le + 1 < (int64_t)n && top_count(le + 1, (int64_t)n - 1) >= 2) {
      gr = true;
    }
    if (trim_end == n - 1) {
      gr = true;
    }
    good = gl && gr;
  } else {
    if (ls - 5 >= 1) {
      if (top_count(ls - 5, ls - 1) >= 2) {
        good = true;
      }
    } else {
      good = true;
      trim_start = 0;
    }
    if (le + 5 < (int64_t)n - 1) {
      if (top_count(le + 1, le + 5) >= 2) {
        good = true;
      }
    } else {
      good = true;
      trim_end = n - 1;
    }
  }
  return good;
}

// ----------------------------------------------------------------------------------------------
// FASTQ reading (btllib::SeqReader as used at goldrush_path.cpp:87,246, read_hashing.cpp:89,
// ntcard.hpp:200): four-line records, id = header up to the first blank, sequence upper-cased.
// ----------------------------------------------------------------------------------------------
struct Read
{
  std::string id;
  std::string seq;
  std::string qual;
};

bool
load_fastq(const std::string& path, std::vector<Read>& reads, bool& is_fastq)
{
  std::ifstream in(path, std::ios::binary);
  if (!in) {
    return false;
  }
  std::string all((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  is_fastq = !all.empty() && all[0] == '@';
  if (!is_fastq) {
    return true;
  }
  size_t p = 0;
  auto next_line = [&](size_t& b, size_t& e) -> bool {
    if (p >= all.size()) {
      return false;
    }
    b = p;
    const void* nl = memchr(all.data() + p, '\n', all.size() - p);
    e = nl ? (size_t)((const char*)nl - all.data()) : all.size();
    p = e + 1;
    while (e > b && isspace((unsigned char)all[e - 1])) {
      --e;
    }
    return true;
  };
  while (true) {
    size_t hb, he, sb, se, pb, pe, qb, qe;
    do {
      if (!next_line(hb, he)) {
        return true;
      }
    } while (he == hb);
    if (!next_line(sb, se) || !next_line(pb, pe) || !next_line(qb, qe)) {
      return true;
    }
    Read r;
    size_t ws = hb + 1;
    while (ws < he && all[ws] != ' ' && all[ws] != '\t') {
      ++ws;
    }
    r.id.assign(all, hb + 1, ws - hb - 1);
    r.seq.assign(all, sb, se - sb);
    for (auto& c : r.seq) {
      c = (char)toupper((unsigned char)c);
    }
    r.qual.assign(all, qb, qe - qb);
    if (r.seq.empty()) {
      return true; // btllib: an empty record ends iteration
    }
    reads.push_back(std::move(r));
  }
}

// ----------------------------------------------------------------------------------------------
// ntcard.hpp:81-154,156-274
// ----------------------------------------------------------------------------------------------
uint64_t
ntcard_estimate(const std::vector<Read>& reads, uint64_t file_bytes,
                const std::vector<SeedTable>& seeds, uint64_t* per_pattern)
{
  const unsigned rBits = 27;
  const unsigned sBits = file_bytes < 50000000000ULL ? 7 : 11; // :182-183
  const uint64_t rBuck = (uint64_t)1 << rBits;
  const uint64_t sMask = (((uint64_t)1) << (sBits - 1)) - 1; // :188
  const size_t h = seeds.size();
  std::vector<std::vector<uint16_t>> counters(h, std::vector<uint16_t>(2 * rBuck, 0));
  std::vector<uint8_t> codes;
  std::vector<uint32_t> next_bad;
  for (const auto& r : reads) {
    bool any_bad;
    encode(r.seq.data(), r.seq.size(), codes, next_bad, any_bad);
    if (r.seq.size() < seeds.back().span) {
      std::cerr << "SeedNtHash: sequence length (" << r.seq.size() << ") is smaller than k ("
                << seeds.back().span << ")" << std::endl;
      exit(EXIT_FAILURE);
    }
    for_each_frame(seeds, codes.data(), codes.size(), any_bad ? next_bad.data() : nullptr,
                   [&](const uint64_t* hv) {
                     for (size_t i = 0; i < h; ++i) { // :103-110, :81-94
                       const uint64_t v = hv[i];
                       uint64_t ind = 2;
                       if ((v >> (63 - sBits)) == 1) {
                         ind = 0;
                       }
                       if ((v >> (64 - sBits)) == sMask) {
                         ind = 1;
                       }
                       if (ind < 2) {
                         ++counters[i][ind * rBuck + (v & (rBuck - 1))];
                       }
                     }
                   });
  }
  uint64_t total = 0;
  for (size_t i = 0; i < h; ++i) { // :114-139: only F0 (histArray[1]) is consumed, :265-270
    double zeros = 0;
    for (unsigned t = 0; t < 2; ++t) {
      uint64_t z = 0;
      for (uint64_t j = 0; j < rBuck; ++j) {
        z += counters[i][t * rBuck + j] == 0;
      }
      zeros += (double)z;
    }
    const double pMean0 = zeros / (1.0 * 2);
    const double F0Mean =
      (double)(ssize_t)((rBits * log(2) - log(pMean0)) * 1.0 * ((size_t)1 << (sBits + rBits)));
    const uint64_t f0 = (uint64_t)(size_t)F0Mean;
    if (per_pattern) {
      per_pattern[i] = f0;
    }
    total += f0;
  }
  return total;
}

// ----------------------------------------------------------------------------------------------
// Options (opt.hpp:9-47, opt.cpp:5-32,90-217)
// ----------------------------------------------------------------------------------------------
struct Opt
{
  size_t assigned_max = 1, unassigned_min = 5, tile_length = 1000;
  uint64_t hash_universe = 0, genome_size = 0;
  size_t kmer_size = 0, weight = 0, min_length = 20000, hash_num = 3;
  double occupancy = 0.1, ratio = 0.9;
  size_t jobs = 48, block_size = 10, max_paths = 1, threshold = 10;
  uint32_t phred_min = 0, phred_delta = 5;
  std::string prefix_file = "goldrush_out", input, seed_preset, filter_file;
  int help = 0, ntcard = 0, silver_path = 0, verbose = 0, debug = 0;
};

std::vector<std::string>
make_seed_pattern(const std::string& preset, unsigned k, unsigned weight, unsigned h, bool log)
{
  std::string left, right;
  if (preset.empty()) { // spaced_seeds.cpp:18-46
    srand(123);
    if (log) {
      std::cerr << "Designing base symmetrical spaced seed\nUsing:\nspan: " << k
                << "\nweight: " << weight << std::endl;
    }
    std::vector<unsigned> half(k / 2, 0);
    half[0] = 1;
    size_t ones = 0;
    while (ones != weight / 2) {
      for (size_t i = 1; i < k / 2; ++i) {
        half[i] = rand() % 2;
      }
      ones = (size_t)std::count(half.begin(), half.end(), 1u);
    }
    for (unsigned v : half) {
      left += v ? '1' : '0';
    }
    right.assign(left.rbegin(), left.rend());
  } else { // :47-61
    if (log) {
      std::cerr << "Using preset spaced seed\nwith:\n\tspan: " << preset.size() << "\n\tweight: "
                << std::count(preset.begin(), preset.end(), '1') << std::endl;
    }
    left = preset.substr(0, preset.size() / 2);
    right = preset.substr(preset.size() / 2, preset.size() / 2);
  }
  std::vector<std::string> out;
  for (unsigned i = 0; i < h; ++i) { // :63-66
    out.push_back(left + std::string(i, '0') + right);
  }
  return out;
}

uint64_t
calc_optimal_size(uint64_t entries, unsigned hash_num, double occupancy)
{
  const size_t v = size_t(-double(entries) * double(hash_num) / log(1.0 - occupancy));
  return v + (64 - v % 64);
}

uint64_t
default_hash_universe(uint64_t weight, uint64_t genome_size, uint64_t hash_num)
{
  const size_t base = std::min((uint64_t)(pow((uint8_t)4, weight)), (uint64_t)2 * genome_size);
  const float coeff = 0.5f; // goldrush_path.cpp:1115: a float, so the product is a float
  return (uint64_t)(base * coeff * hash_num);
}

struct PathLog // goldrush_path.cpp:41-51
{
  uint64_t valid_reads = 0, total_tiles = 0, assigned = 0, unassigned = 0, queries = 0, hits = 0,
           misses = 0, num_reads_in_path = 0;
  double phred_sum = 0;
};

void
log_path_stat(uint64_t curr_path, const PathLog& l, uint64_t inserted_bases)
{
  std::cerr << "Visited " << l.valid_reads << " reads to generate " << curr_path
            << " silver paths\n"
            << "Saw: " << l.total_tiles << " tiles to generate " << curr_path << " silver paths\n"
            << "Assigned: " << l.assigned << " tiles to generate " << curr_path
            << " silver paths\n"
            << "Unassigned: " << l.unassigned << " tiles to generate " << curr_path
            << " silver paths\n"
            << "Total queries: " << l.queries << " to generate " << curr_path << " silver paths\n"
            << "Total hits: " << l.hits << " to generate " << curr_path << " silver paths\n"
            << "Total misses: " << l.misses << " to generate " << curr_path << " silver paths\n"
            << "Num reads: " << l.num_reads_in_path << " in silver path " << curr_path << "\n";
  const uint32_t avg = (uint32_t)(-10 * log10(l.phred_sum / inserted_bases));
  std::cerr << "Average Phred: " << avg << " in silver path " << curr_path << std::endl;
}

int
run_path(Opt& opt)
{
#if _OPENMP
  omp_set_num_threads((int)opt.jobs);
#endif
  const auto seed_strings = make_seed_pattern(opt.seed_preset, (unsigned)opt.kmer_size,
                                              (unsigned)opt.weight, (unsigned)opt.hash_num, true);
  std::vector<SeedTable> seeds;
  for (const auto& s : seed_strings) {
    seeds.emplace_back(s);
  }
  const unsigned h = (unsigned)seeds.size();
  const unsigned k = (unsigned)opt.kmer_size;
  const unsigned max_span = seeds.back().span;

  std::vector<Read> reads;
  bool is_fastq = false;
  if (!load_fastq(opt.input, reads, is_fastq)) {
    std::cerr << "cannot open " << opt.input << std::endl;
    return 1;
  }
  uint64_t file_bytes = 0;
  {
    std::ifstream in(opt.input, std::ifstream::ate | std::ifstream::binary);
    file_bytes = (uint64_t)in.tellg();
  }

  if (opt.hash_universe == 0) { // goldrush_path.cpp:1109-1123
    if (opt.ntcard) {
      std::cerr << "Calculating expected entries" << std::endl;
      std::vector<uint64_t> per(h);
      opt.hash_universe = ntcard_estimate(reads, file_bytes, seeds, per.data());
      for (unsigned i = 0; i < h; ++i) {
        std::cerr << "Expected entries for seed pattern " << seed_strings[i] << " : " << per[i]
                  << std::endl;
      }
      std::cerr << "Total expected entries for seed patterns: " << opt.hash_universe << std::endl;
    } else {
      opt.hash_universe = default_hash_universe(opt.weight, opt.genome_size, opt.hash_num);
    }
  }

  if (opt.phred_min == 0) { // :79-107; sample = first 50000 reads >= min_length in file order
    std::cerr << "Calculating minimum phred score via median" << std::endl;
    const size_t cap = 50000;
    std::vector<uint32_t> scores(cap, 0);
    size_t n = 0;
    for (const auto& r : reads) {
      if (r.seq.size() < opt.min_length) {
        continue;
      }
      if (n >= cap) {
        ++n; // the reference's counter runs one past the cap before it stops (:93-96, one thread)
        break;
      }
      uint32_t avg, delta;
      phred_average(r.qual.data(), r.qual.size(), avg, delta, nullptr);
      scores[n++] = avg;
    }
    std::sort(scores.begin(), scores.end(), std::greater<uint32_t>());
    opt.phred_min = std::max<uint32_t>(10, scores[n / 2]);
    if (opt.verbose) {
      std::cerr << "Minimum phred score calculated with median: " << opt.phred_min << std::endl;
    }
  }

  std::cerr << "Calculating "
            << (opt.silver_path ? std::to_string(opt.max_paths) + " silver path(s)"
                                : std::string("the golden path"))
            << "\nUsing:\n\ttile length: " << opt.tile_length
            << "\n\tblock size: " << opt.block_size << "\n\tseed patterns: " << opt.hash_num
            << "\n\tthreshold: " << opt.threshold << "\n\tbase seed pattern: " << seed_strings[0]
            << "\n\tminimum unassigned tiles: " << opt.unassigned_min
            << "\n\tmaximum assigned tiles: " << opt.assigned_max
            << "\n\texpected hash space: " << opt.hash_universe
            << "\n\tminimum average phred quality score: " << opt.phred_min
            << "\n\tmaximum average phred delta between first and second half of read: "
            << opt.phred_delta << "\n\toccupancy: " << opt.occupancy << "\n\tjobs: " << opt.jobs
            << std::endl;

  std::unordered_set<std::string> filter_out; // :1163-1172
  if (!opt.filter_file.empty()) {
    std::cerr << "Using only reads not found in: " << opt.filter_file << std::endl;
    std::ifstream fin(opt.filter_file);
    std::string name;
    while (fin >> name) {
      filter_out.insert(name);
    }
  }

  std::ofstream out(opt.silver_path ? opt.prefix_file + "_1.fq" : opt.prefix_file + ".fa");

  double t0 = omp_get_wtime();
  std::cerr << "allocating bit vector" << std::endl;
  Filter f;
  f.bits = calc_optimal_size(opt.hash_universe, 1, opt.occupancy); // :1183-1184
  f.h = h;
  std::cerr << "m_filterSize: " << f.bits << std::endl;
  f.words.assign((f.bits + 63) / 64 + 1, 0);
  std::cerr << "finished allocating bit vector\nin " << std::fixed << t0 - t0 << "\n";
  std::cerr << "opening: " << opt.input << std::endl;

  // ---- pass 1, goldrush_path.cpp:235-339 ----
  std::cerr << "inserting bit vector" << std::endl;
  t0 = omp_get_wtime();
  if (!is_fastq) {
    std::cerr << "Gold Path requires fastq format" << std::endl;
    return 1;
  }
  const size_t nreads = reads.size();
  std::vector<uint8_t> state(nreads, 0); // 0 pass, 1 short, 2 phred/delta, 3 bases
  size_t passed = 0, by_phred = 0, by_delta = 0, by_len = 0, by_bases = 0;
  bool too_short_for_seed = false;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : passed, by_phred, by_delta, by_len, by_bases)
  for (size_t r = 0; r < nreads; ++r) {
    const Read& rd = reads[r];
    if (rd.seq.size() < opt.min_length) {
      state[r] = 1;
      ++by_len;
      continue;
    }
    uint32_t avg, delta;
    phred_average(rd.qual.data(), rd.qual.size(), avg, delta, nullptr);
    if (avg < opt.phred_min || delta >= opt.phred_delta) {
      by_phred += avg < opt.phred_min;
      by_delta += delta >= opt.phred_delta;
      state[r] = 2;
      continue;
    }
    if (rd.seq.find_first_not_of("ACGTacgt") != std::string::npos) {
      ++by_bases;
      state[r] = 3;
      continue;
    }
    ++passed;
    if (rd.seq.size() < max_span) {
      too_short_for_seed = true;
      continue;
    }
    std::vector<uint8_t> codes;
    std::vector<uint32_t> nb;
    bool any_bad;
    encode(rd.seq.data(), rd.seq.size(), codes, nb, any_bad);
    for_each_frame(seeds, codes.data(), codes.size(), nullptr, [&](const uint64_t* hv) {
      for (unsigned p = 0; p < h; ++p) {
        f.set_bit(hv[p]);
      }
    });
  }
  if (too_short_for_seed) {
    std::cerr << "SeedNtHash: sequence length is smaller than k" << std::endl;
    return 1;
  }
  for (size_t r = 0; r < nreads; ++r) {
    if (state[r] >= 2) {
      filter_out.insert(reads[r].id); // :287-299
    }
  }
  if (opt.verbose) {
    std::cerr << "num_passed_reads: " << passed << "\nnum_reads: " << nreads
              << "\nnum_reads - num_passed_reads: " << nreads - passed
              << "\nnum_reads - num_passed_reads / num_reads: "
              << floor((double)(nreads - passed) / nreads)
              << "\nnum_reads_skipped_by_phred: " << by_phred
              << "\nnum_reads_skipped_by_delta: " << by_delta
              << "\nnum_reads_skipped_by_length: " << by_len
              << "\nnum_reads_skipped_by_invalid_bases: " << by_bases
              << "\nTotal reads skipped: " << by_phred + by_delta + by_len + by_bases << std::endl;
  }
  if (passed == 0) {
    std::cerr << "Error: no reads passed the Phred score and min length requirements\n"
              << "Try again with a lower Phred threshold or lower min length" << std::endl;
    return 1;
  }
  std::cerr << "finished inserting bit vector\nin " << omp_get_wtime() - t0 << "\n";

  f.setup(); // :1203-1205
  // MIBloomFilter.hpp:180: the filter's constructor asserts that seed pattern 0 spans -k; an odd -k
  // gives a span of k - 1 (spaced_seeds.cpp:28,58-60) and the reference build (meson default:
  // asserts on) aborts here, after pass 1
  if (seed_strings[0].size() != opt.kmer_size) {
    std::cerr << "Assertion `m_sseeds[0].size() == kmerSize' failed." << std::endl;
    abort();
  }

  // ---- pass 2, goldrush_path.cpp:1207-1256 + process_read :892-1094 ----
  std::cerr << "assigning tiles" << std::endl;
  t0 = omp_get_wtime();
  uint64_t inserted_bases = 0;
  const uint64_t target_bases = (uint64_t)(opt.ratio * opt.genome_size);
  uint64_t curr_path = 1;
  uint32_t id = 1, ids_inserted = 0;
  PathLog lg;
  const size_t T = opt.tile_length, B = opt.block_size;
  const char first_char = opt.silver_path ? '@' : '>';
  bool finished = false;

  auto tick = [&]() {
    ++id;
    if (id % 10000 == 0) {
      std::cerr << "processed " << id << " reads" << std::endl;
    }
  };
  auto silver_check = [&]() { // :156-187
    if (target_bases < inserted_bases) {
      if (opt.verbose) {
        log_path_stat(curr_path, lg, inserted_bases);
      }
      ++curr_path;
      if (opt.max_paths < curr_path) {
        finished = true; // exit(0)
        return;
      }
      inserted_bases = 0;
      lg.num_reads_in_path = 0;
      lg.phred_sum = 0;
      std::fill(f.counts.begin(), f.counts.end(), 0u);
      std::fill(f.data.begin(), f.data.end(), 0u);
      out.close();
      out.open(opt.prefix_file + "_" + std::to_string(curr_path) + ".fq");
      ids_inserted = 0;
    }
  };

  std::vector<uint8_t> codes;
  std::vector<uint32_t> nb;
  for (size_t r = 0; r < nreads && !finished; ++r) {
    const Read& rd = reads[r];
    if (rd.seq.size() < opt.min_length) { // :907-918
      tick();
      continue;
    }
    if (!filter_out.empty() && filter_out.count(rd.id)) { // :919-932
      tick();
      continue;
    }
    const size_t len = rd.seq.size();
    const size_t num_tiles = len / T;
    lg.total_tiles += num_tiles;

    // read_hashing.cpp:43-55: tile i covers substr(i*T, T + k - 1)
    bool any_bad;
    encode(rd.seq.data(), len, codes, nb, any_bad);
    std::vector<std::vector<uint64_t>> tile_hashes(num_tiles);
    bool short_tile = false;
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t i = 0; i < num_tiles; ++i) {
      const size_t tb = i * T;
      const size_t tl = std::min(T + k - 1, len - tb);
      if (tl < max_span) {
        short_tile = true;
        continue;
      }
      auto& hv = tile_hashes[i];
      hv.reserve((tl - k + 1) * h);
      for_each_frame(seeds, codes.data() + tb, tl, any_bad ? nb.data() + tb : nullptr,
                     [&](const uint64_t* v) { hv.insert(hv.end(), v, v + h); });
    }
    if (short_tile) {
      std::cerr << "SeedNtHash: sequence length is smaller than k" << std::endl;
      return 1;
    }

    std::vector<uint32_t> tid(num_tiles, 0);
    std::vector<uint8_t> tas(num_tiles, 0);
    std::vector<TileVote> votes(num_tiles);
    uint64_t counters[3] = { 0, 0, 0 };
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t i = 0; i < num_tiles; ++i) {
      vote_tile(f, tile_hashes[i].data(), tile_hashes[i].size() / h, votes[i], counters);
      tid[i] = votes[i].best_id;
    }
    lg.queries += counters[0];
    lg.hits += counters[1];
    lg.misses += counters[2];
    const size_t n_as = smooth_tiles(num_tiles, tid.data(), tas.data(), votes, opt.threshold);
    const size_t n_un = num_tiles - n_as;
    lg.assigned += n_as;
    lg.unassigned += n_un;

    auto insert_range = [&](size_t a, size_t b, uint32_t the_id) { // tiles [a,b)
      std::vector<uint64_t> flat;
      for (size_t i = a; i < b; ++i) {
        flat.insert(flat.end(), tile_hashes[i].begin(), tile_hashes[i].end());
      }
      insert_mibf(f, flat.data(), flat.size(), the_id);
    };

    if (n_un >= opt.unassigned_min && n_as <= opt.assigned_max) { // :968-1011
      ++ids_inserted;
      for (size_t bs = 0; bs < num_tiles; bs += B) {
        const size_t be = std::min(bs + B, num_tiles);
        insert_range(bs, be, ids_inserted + uint32_t(bs / B));
      }
      ids_inserted = ids_inserted + uint32_t(len / (T * B));
      out << first_char << rd.id << "_untrimmed\n" << rd.seq << std::endl;
      inserted_bases += len;
      ++lg.num_reads_in_path;
      lg.phred_sum += grbo_sum_phred(rd.qual.data(), rd.qual.size());
      if (opt.silver_path) {
        out << "+\n" << rd.qual << std::endl;
        silver_check();
        if (finished) {
          break;
        }
      }
    } else {
      if (n_as == num_tiles) { // :1013-1023
        ++lg.valid_reads;
        tick();
        continue;
      }
      int64_t ls, le;
      find_longest_stretch(tas.data(), num_tiles, ls, le);
      uint64_t ts, te;
      if (eval_flanks(ls, le, tid.data(), num_tiles, ts, te)) { // :1035-1079
        ++ids_inserted;
        for (size_t bs = ts; bs <= te; bs += B) {
          const size_t be = std::min(bs + B - 1, (size_t)te);
          insert_range(bs, be + 1, ids_inserted + uint32_t((bs - ts + 1) / B));
        }
        ids_inserted = ids_inserted + uint32_t((te - ts) / B);
        const size_t end_pos = (te == num_tiles - 1) ? std::string::npos : (te - ts + 1) * T;
        const std::string new_seq = rd.seq.substr(ts * T, end_pos);
        const std::string new_qual = rd.qual.substr(ts * T, end_pos);
        inserted_bases += new_seq.size();
        out << first_char << rd.id << "_trimmed\n" << new_seq << std::endl;
        ++lg.num_reads_in_path;
        lg.phred_sum += grbo_sum_phred(new_qual.data(), new_qual.size());
        if (opt.silver_path) {
          out << "+\n" << new_qual << std::endl;
          silver_check();
          if (finished) {
            break;
          }
        }
      }
    }
    ++lg.valid_reads;
    tick();
  }
  if (finished) {
    return 0;
  }
  if (opt.silver_path && opt.max_paths > curr_path) { // :1257-1264
    std::cerr << "WARNING: Expected " << opt.max_paths << " silver paths, but only " << curr_path
              << " generated.\nPossible reasons include:\n"
              << "\t- Input reads sorted by chromosome/position\n"
              << "\t- Genome size set too large\n";
  }
  if (opt.verbose) {
    log_path_stat(curr_path, lg, inserted_bases);
  }
  std::cerr << "assigned\nin " << omp_get_wtime() - t0 << "\n";
  return 0;
}

} // namespace

extern "C" {

int
grbo_make_seed_pattern(const char* preset, unsigned k, unsigned w, unsigned h, char** out)
{
  const auto v = make_seed_pattern(preset ? preset : "", k, w, h, false);
  for (unsigned i = 0; i < h; ++i) {
    strcpy(out[i], v[i].c_str());
  }
  return 0;
}

uint64_t
grbo_calc_optimal_size(uint64_t entries, unsigned hash_num, double occupancy)
{
  return calc_optimal_size(entries, hash_num, occupancy);
}

uint64_t
grbo_default_hash_universe(uint64_t weight, uint64_t genome_size, uint64_t hash_num)
{
  return default_hash_universe(weight, genome_size, hash_num);
}

void
grbo_calc_phred_average(const char* qual, size_t n, uint32_t* avg, uint32_t* delta, double* sums)
{
  phred_average(qual, n, *avg, *delta, sums);
}

double
grbo_sum_phred(const char* qual, size_t n)
{
  double s = 0;
  for (size_t i = 0; i < n; ++i) {
    s += delog(qual[i]);
  }
  return s;
}

size_t
grbo_hash_sequence(const char* seq, size_t n, const char* const* seeds, unsigned h, uint64_t* out,
                   size_t out_cap)
{
  std::vector<SeedTable> st;
  for (unsigned i = 0; i < h; ++i) {
    st.emplace_back(std::string(seeds[i]));
  }
  if (n < st.back().span) {
    return 0;
  }
  std::vector<uint8_t> codes;
  std::vector<uint32_t> nb;
  bool any_bad;
  encode(seq, n, codes, nb, any_bad);
  size_t w = 0;
  return for_each_frame(st, codes.data(), n, any_bad ? nb.data() : nullptr,
                        [&](const uint64_t* v) {
                          for (unsigned p = 0; p < h; ++p) {
                            if (w < out_cap) {
                              out[w] = v[p];
                            }
                            ++w;
                          }
                        });
}

struct grbo_filter
{
  Filter f;
};

grbo_filter*
grbo_filter_new(uint64_t filter_bits, unsigned h)
{
  auto* g = new grbo_filter;
  g->f.bits = filter_bits;
  g->f.h = h;
  g->f.words.assign((filter_bits + 63) / 64 + 1, 0);
  return g;
}

void
grbo_filter_free(grbo_filter* f)
{
  delete f;
}

void
grbo_filter_insert_bv(grbo_filter* g, const uint64_t* hashes, size_t n)
{
  for (size_t i = 0; i < n; ++i) {
    g->f.set_bit(hashes[i]);
  }
}

uint64_t
grbo_filter_setup(grbo_filter* g)
{
  return g->f.setup();
}

const uint64_t*
grbo_filter_words(const grbo_filter* g, uint64_t* n_words)
{
  *n_words = (g->f.bits + 63) / 64;
  return g->f.words.data();
}

uint64_t
grbo_filter_rank(const grbo_filter* g, uint64_t pos, int* bit)
{
  if (bit) {
    *bit = g->f.bit(pos);
  }
  return g->f.rank(pos);
}

uint32_t
grbo_filter_get_id(const grbo_filter* g, uint64_t rank)
{
  return g->f.data[rank];
}

uint32_t
grbo_filter_get_count(const grbo_filter* g, uint64_t rank)
{
  return g->f.counts[rank];
}

void
grbo_filter_set(grbo_filter* g, uint64_t rank, uint32_t id, uint32_t count)
{
  g->f.data[rank] = id;
  g->f.counts[rank] = count;
}

void
grbo_filter_reset_ids(grbo_filter* g)
{
  std::fill(g->f.data.begin(), g->f.data.end(), 0u);
  std::fill(g->f.counts.begin(), g->f.counts.end(), 0u);
}

uint32_t
grbo_query_tile(const grbo_filter* g, const uint64_t* hashes, size_t frames, uint32_t* best_id,
                uint32_t* best_count, uint32_t* cand_ids, uint32_t* cand_counts, uint32_t cand_cap,
                uint64_t* counters)
{
  TileVote v;
  vote_tile(g->f, hashes, frames, v, counters);
  *best_id = v.best_id;
  *best_count = v.best_count;
  for (size_t i = 0; i < v.cands.size() && i < cand_cap; ++i) {
    cand_ids[i] = v.cands[i].first;
    cand_counts[i] = v.cands[i].second;
  }
  return (uint32_t)v.cands.size();
}

void
grbo_insert_mibf(grbo_filter* g, const uint64_t* hashes, size_t n, uint32_t id)
{
  insert_mibf(g->f, hashes, n, id);
}

size_t
grbo_smooth_tiles(size_t num_tiles, uint32_t* ids, uint8_t* assigned, const uint32_t* cand_off,
                  const uint32_t* cand_ids, const uint32_t* cand_counts, uint64_t threshold)
{
  std::vector<TileVote> votes(num_tiles);
  for (size_t i = 0; i < num_tiles; ++i) {
    for (uint32_t j = cand_off[i]; j < cand_off[i + 1]; ++j) {
      votes[i].cands.emplace_back(cand_ids[j], cand_counts[j]);
    }
    std::stable_sort(votes[i].cands.begin(), votes[i].cands.end(),
                     [](const auto& a, const auto& b) { return a.second > b.second; });
  }
  return smooth_tiles(num_tiles, ids, assigned, votes, threshold);
}

void
grbo_find_longest_stretch(const uint8_t* assigned, size_t n, int64_t* start, int64_t* end)
{
  find_longest_stretch(assigned, n, *start, *end);
}

int
grbo_eval_flanks(int64_t ls, int64_t le, const uint32_t* ids, size_t n, uint64_t* trim_start,
                 uint64_t* trim_end)
{
  return eval_flanks(ls, le, ids, n, *trim_start, *trim_end) ? 1 : 0;
}

uint64_t
grbo_ntcard_sized(const char* fastq_path, const char* const* seed_strs, unsigned h, uint64_t file_bytes,
                  uint64_t* per_pattern)
{
  std::vector<Read> reads;
  bool is_fastq;
  load_fastq(fastq_path, reads, is_fastq);
  std::ifstream in(fastq_path, std::ifstream::ate | std::ifstream::binary);
  const uint64_t bytes = file_bytes ? file_bytes : (uint64_t)in.tellg();
  std::vector<SeedTable> seeds;
  for (unsigned i = 0; i < h; ++i) {
    seeds.emplace_back(std::string(seed_strs[i]));
  }
  return ntcard_estimate(reads, bytes, seeds, per_pattern);
}

uint64_t
grbo_ntcard(const char* fastq_path, const char* const* seed_strs, unsigned h, uint64_t* per_pattern)
{
  return grbo_ntcard_sized(fastq_path, seed_strs, h, 0, per_pattern);
}

int
grbo_main(int argc, char** argv)
{
  Opt opt;
  static int f_debug, f_verbose, f_silver, f_help, f_ntcard;
  f_debug = f_verbose = f_silver = f_help = f_ntcard = 0;
  static const struct option longopts[] = { { "debug", no_argument, &f_debug, 1 },
                                            { "verbose", no_argument, &f_verbose, 1 },
                                            { "silver_path", no_argument, &f_silver, 1 },
                                            { "help", no_argument, &f_help, 1 },
                                            { "ntcard", no_argument, &f_ntcard, 1 },
                                            { nullptr, 0, nullptr, 0 } };
  optind = 1;
  int c, idx = 0;
  char* end = nullptr;
  while ((c = getopt_long(argc, argv, "a:b:d:f:g:h:i:j:k:m:M:o:r:s:t:u:w:x:p:P:H:", longopts,
                          &idx)) != -1) {
    switch (c) {
      case 0: break;
      case 'a': opt.assigned_max = strtoul(optarg, &end, 10); break;
      case 'b': opt.block_size = strtoul(optarg, &end, 10); break;
      case 'd': opt.phred_delta = (uint32_t)strtoul(optarg, &end, 10); break;
      case 'f': opt.filter_file = optarg; break;
      case 'H': opt.hash_universe = strtoull(optarg, &end, 10); break;
      case 'h': opt.hash_num = strtoul(optarg, &end, 10); break;
      case 'i': opt.input = optarg; break;
      case 'j': opt.jobs = strtoul(optarg, &end, 10); break;
      case 'k': opt.kmer_size = strtoul(optarg, &end, 10); break;
      case 'm': opt.min_length = strtoul(optarg, &end, 10); break;
      case 'M': opt.max_paths = strtoul(optarg, &end, 10); break;
      case 'o': opt.occupancy = strtod(optarg, &end); break;
      case 'r': opt.ratio = strtod(optarg, &end); break;
      case 'p': opt.prefix_file = optarg; break;
      case 'P': opt.phred_min = (uint32_t)strtoul(optarg, &end, 10); break;
      case 's': opt.seed_preset = optarg; break;
      case 't': opt.tile_length = strtoul(optarg, &end, 10); break;
      case 'g': opt.genome_size = (uint64_t)strtod(optarg, &end); break;
      case 'u': opt.unassigned_min = strtoul(optarg, &end, 10); break;
      case 'w': opt.weight = strtoul(optarg, &end, 10); break;
      case 'x': opt.threshold = strtoul(optarg, &end, 10); break;
      default: return 1;
    }
  }
  opt.debug = f_debug;
  opt.verbose = f_verbose;
  opt.silver_path = f_silver;
  opt.help = f_help;
  opt.ntcard = f_ntcard;
  if (opt.help) {
    std::cout << "goldrush-path-oracle: CPU oracle, same options as goldrush-path\n";
    return 0;
  }
  if (!opt.kmer_size) {
    std::cerr << "span of spaced seed cannot be 0" << std::endl;
    return 1;
  }
  if (!opt.weight) {
    std::cerr << "weight of spaced seed cannot be 0" << std::endl;
    return 1;
  }
  if (opt.genome_size == 0) {
    std::cerr << "genome size cannot be 0" << std::endl;
    return 1;
  }
  if (!opt.seed_preset.empty()) {
    if (opt.kmer_size != opt.seed_preset.size()) {
      std::cerr << "seed preset must be the same size of k" << std::endl;
      return 1;
    }
    uint8_t ones = 0;
    for (char ch : opt.seed_preset) {
      ones += ch == '1';
    }
    if (opt.weight != ones) {
      std::cerr << "seed preset must have the same weight as w" << std::endl;
      return 1;
    }
  }
  return run_path(opt);
}

} // extern "C"
