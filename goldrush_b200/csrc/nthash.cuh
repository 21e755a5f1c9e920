// Spaced-seed ntHash on the device, one value per (position, pattern), computed directly from the
// 2-bit packed read (no rolling state to carry between threads).
//
// Replaces btllib::SeedNtHash as driven by multiLensfrHashIterator
// (goldrush_path/multiLensfrHashIterator.hpp:29-68).  Definition (restated in oracle/grb_oracle.cpp
// and oracle/shim/btllib/nthash.hpp; constants stated in-tree at
// subprojects/goldpolish/subprojects/ntedit/lib/nthash.hpp:24-28,66-92,529-563,172-191):
//   fwd = XOR over care positions p of srol^(span-1-p)(seed[base_p])
//   rev = XOR over care positions p of srol^(p)(seed[3 - base_p])
//   hash = fwd + rev
// where srol rotates the upper 31 and the lower 33 bits independently.
//
// GoldRush's patterns are always  left + i zeros + right  (spaced_seeds.cpp:63-66), so with
// L = |left| care positions split into a left set (offsets < L) and a right set, and by linearity
// of srol over XOR:
//   fwd_i(f) = srol^i(FL(f)) ^ FR(f + i)      FL(f) = XOR_left  srol^(k-1-p)(seed[b(f+p)])
//   rev_i(f) = RL(f) ^ srol^i(RR(f + i))      FR(g) = XOR_right srol^(k-1-q)(seed[b(g+q)])   (q in pattern 0)
// so the four half hashes are computed once per position and shared by all h patterns.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define GRB_MAX_PATTERNS 8
#define GRB_MAX_WEIGHT 64
#define GRB_MAX_SPAN 64

struct GrbSeedTables
{
  // half-hash tables of pattern 0: [care index][base]; care indices [0, n_left) are the left half
  uint64_t fwd[GRB_MAX_WEIGHT][4];
  uint64_t rev[GRB_MAX_WEIGHT][4];
  uint8_t care[GRB_MAX_WEIGHT]; // offsets in pattern 0
  uint32_t n_care;
  uint32_t n_left;  // care positions in the left half
  uint32_t half;    // |left| = first offset of the right half in pattern 0
  uint32_t k;       // span of pattern 0
  uint32_t h;       // patterns; span of pattern i = k + i
};

__host__ __device__ __forceinline__ uint64_t
grb_srol(uint64_t x, unsigned r)
{
  // r < 31
  if (r == 0) {
    return x;
  }
  const uint64_t hi = x >> 33;
  const uint64_t lo = x & 0x1FFFFFFFFULL;
  const uint64_t nh = ((hi << r) | (hi >> (31 - r))) & 0x7FFFFFFFULL;
  const uint64_t nl = ((lo << r) | (lo >> (33 - r))) & 0x1FFFFFFFFULL;
  return (nh << 33) | nl;
}

// general rotation amount (host-side table construction)
inline uint64_t
grb_srol_any(uint64_t x, unsigned r)
{
  const uint64_t hi = x >> 33;
  const uint64_t lo = x & 0x1FFFFFFFFULL;
  const unsigned rh = r % 31, rl = r % 33;
  const uint64_t nh = rh ? (((hi << rh) | (hi >> (31 - rh))) & 0x7FFFFFFFULL) : hi;
  const uint64_t nl = rl ? (((lo << rl) | (lo >> (33 - rl))) & 0x1FFFFFFFFULL) : lo;
  return (nh << 33) | nl;
}

// 64 bases starting at absolute base index `pos` of a packed array (32 bases per word, LSB first).
// The array must be readable two words past the word holding `pos`.
struct GrbWindow
{
  uint64_t lo, hi;
  __device__ __forceinline__ unsigned base(unsigned i) const
  {
    return (unsigned)((i < 32 ? (lo >> (2 * i)) : (hi >> (2 * (i - 32)))) & 3u);
  }
};

template<typename WordLoader>
__device__ __forceinline__ GrbWindow
grb_window(WordLoader&& word, uint64_t pos)
{
  const uint64_t wi = pos >> 5;
  const unsigned s = (unsigned)(pos & 31) * 2;
  const uint64_t w0 = word(wi), w1 = word(wi + 1), w2 = word(wi + 2);
  GrbWindow w;
  if (s == 0) {
    w.lo = w0;
    w.hi = w1;
  } else {
    w.lo = (w0 >> s) | (w1 << (64 - s));
    w.hi = (w1 >> s) | (w2 << (64 - s));
  }
  return w;
}

struct GrbHalf
{
  uint64_t fl, fr, rl, rr;
};

// Half hashes of pattern 0 anchored at the window start: the left set reads bases at its own
// offsets, the right set at (offset - half) relative to a window that starts at f + half.
// To keep one window per thread, both are evaluated on the window that starts at position g:
//   left(g)  uses bases g + care[j]            (j <  n_left)
//   right(g) uses bases g + care[j] - half     (j >= n_left)   == FR/RR of frame (g - half)
__device__ __forceinline__ GrbHalf
grb_half_hashes(const GrbSeedTables& t, const GrbWindow& w)
{
  GrbHalf o{ 0, 0, 0, 0 };
  const unsigned nl = t.n_left, nc = t.n_care, half = t.half;
#pragma unroll 4
  for (unsigned j = 0; j < nl; ++j) {
    const unsigned b = w.base(t.care[j]);
    o.fl ^= t.fwd[j][b];
    o.rl ^= t.rev[j][b];
  }
#pragma unroll 4
  for (unsigned j = nl; j < nc; ++j) {
    const unsigned b = w.base(t.care[j] - half);
    o.fr ^= t.fwd[j][b];
    o.rr ^= t.rev[j][b];
  }
  return o;
}

// hash of pattern i at frame f from the left halves at f and the right halves at f + half + i
__host__ __device__ __forceinline__ uint64_t
grb_combine(unsigned i, uint64_t fl, uint64_t rl, uint64_t fr, uint64_t rr)
{
  return (grb_srol(fl, i) ^ fr) + (rl ^ grb_srol(rr, i));
}

// ---- grouped half-hash tables ---------------------------------------------------------------------
// The per-care-position evaluation above costs about 20 instructions per care position (variable
// 64-bit shifts to pull one base out, two table reads, two 64-bit XORs): ncu showed the fill kernel
// issue-bound on it.  Both halves only look at window offsets [0, half), so the window is cut into
// groups of 4 consecutive bases = one byte of the packed window, and a 256-entry table per group
// holds the XOR of the contributions of every care position inside the group for each of the 256
// base combinations (don't-care offsets simply do not contribute).  A half hash is then
// ceil(half / 4) byte extractions and 16-byte table reads:
//   L[g][v] = { fl, rl } of the left  care positions with offset in [4g, 4g + 4)
//   R[g][v] = { fr, rr } of the right care positions with offset - half in [4g, 4g + 4)
// Built on the host from GrbSeedTables (engine.cu, build_group_tables).
#define GRB_MAX_GROUPS 8 // half <= 32 bases

// the 32 bases starting at absolute base index `pos` (LSB first)
template<typename WordLoader>
__host__ __device__ __forceinline__ uint64_t
grb_lo64(WordLoader&& word, uint64_t pos)
{
  const uint64_t wi = pos >> 5;
  const unsigned s = (unsigned)(pos & 31) * 2;
  const uint64_t w0 = word(wi);
  return s == 0 ? w0 : ((w0 >> s) | (word(wi + 1) << (64 - s)));
}

// out.x ^= first halves, out.y ^= second halves of the table entries picked by the window bytes
__host__ __device__ __forceinline__ ulonglong2
grb_group_half(const ulonglong2* __restrict__ tab, unsigned ng, uint64_t lo)
{
  ulonglong2 o = make_ulonglong2(0, 0);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (unsigned g = 0; g < GRB_MAX_GROUPS; ++g) {
    if (g < ng) {
      const unsigned v = (unsigned)(lo >> (8 * g)) & 0xFFu;
      const ulonglong2 e = tab[g * 256 + v];
      o.x ^= e.x;
      o.y ^= e.y;
    }
  }
  return o;
}

// Straightforward evaluation of one pattern at one position (used where the sharing above does
// not pay, and by the parity export): window starts at the frame position.
__device__ __forceinline__ uint64_t
grb_hash_direct(const GrbSeedTables& t, unsigned i, const GrbWindow& w)
{
  uint64_t fl = 0, rl = 0, fr = 0, rr = 0;
  const unsigned nl = t.n_left, nc = t.n_care;
  for (unsigned j = 0; j < nl; ++j) {
    const unsigned b = w.base(t.care[j]);
    fl ^= t.fwd[j][b];
    rl ^= t.rev[j][b];
  }
  for (unsigned j = nl; j < nc; ++j) {
    const unsigned b = w.base(t.care[j] + i);
    fr ^= t.fwd[j][b];
    rr ^= t.rev[j][b];
  }
  return grb_combine(i, fl, rl, fr, rr);
}
