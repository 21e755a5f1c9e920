// (f4) GoldPolish targeted Bloom filters: the per-(batch, k) work of the builder, written once for
// host and device.
//
// Replaces, for one batch of mapped reads and one k value:
//   fill_bfs                                  subprojects/goldpolish/src/utils.cpp:96-123
//   the filter set-up of serve_batch          subprojects/goldpolish/src/goldpolish_targeted_bfs.cpp:68-77
// and the btllib pieces they bind (third party, not in the tree): NtHash over the read
// (arithmetic as stated in-tree in ntedit/lib/nthash.hpp:24-28,100-191,262-300; k-mers holding a
// character outside ACGTacgt are skipped), KmerCountingBloomFilter8::insert_thresh_contains
// (conservative-update 8-bit counters) and KmerBloomFilter::insert.  oracle/shim_polish/btllib/
// says which of these semantics are recalled rather than read: parity is unpinned there.
//
// Why one sequential job per (batch, k): with conservative update a k-mer's count after an insert
// depends on the counters its hashes share with EARLIER k-mers, so which k-mers reach the
// threshold depends on the order of the inserts.  The reference keeps that order inside a batch
// (one OpenMP task per batch, goldpolish_targeted_bfs.cpp:181-196) and runs batches side by side;
// so does this: a job owns its counting filter and its Bloom filter, needs no atomics, and the GPU's
// parallelism is the thousands of (batch, k) jobs in flight.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define GRB_PHD __host__ __device__ __forceinline__
#else
#define GRB_PHD inline
#endif

struct GrbPolishJob
{
  const char* seqs;        // all mapped reads of the call, back to back
  const uint64_t* off;     // [n_reads + 1] read r = seqs[off[r], off[r + 1])
  const uint32_t* thr;     // [n_reads] k-mer threshold of the read's target (goldpolish_targeted_bfs.cpp:124-127)
  uint64_t first, last;    // the batch's reads, in serve_batch's order
  uint32_t k, k_index, hash_num;
  uint8_t* cbf;            // cbf_bytes counters, zeroed
  uint64_t cbf_bytes, cbf_inv;
  uint8_t* bf;             // bf_bytes bytes, zeroed
  uint64_t bf_bits, bf_inv;
};

GRB_PHD int
grb_p_code(unsigned char c)
{
  c &= 0xDF;
  return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
}

GRB_PHD uint64_t
grb_p_seed(int code) // nthash.hpp:24-28
{
  return code == 0 ? 0x3c8bfbb395c60474ULL
                   : code == 1 ? 0x3193c18562a02b4cULL : code == 2 ? 0x20323ed082572324ULL : 0x295549f54be24456ULL;
}

// rol1 + swapbits033 (nthash.hpp:66-92): the upper 31 and the lower 33 bits rotate independently
GRB_PHD uint64_t
grb_p_srol1(uint64_t x)
{
  const uint64_t m = ((x & 0x8000000000000000ULL) >> 30) | ((x & 0x100000000ULL) >> 32);
  return ((x << 1) & 0xFFFFFFFDFFFFFFFFULL) | m;
}

GRB_PHD uint64_t
grb_p_sror1(uint64_t x)
{
  const uint64_t m = ((x & 0x200000000ULL) << 30) | ((x & 1ULL) << 32);
  return ((x >> 1) & 0xFFFFFFFEFFFFFFFFULL) | m;
}

GRB_PHD uint64_t
grb_p_srol(uint64_t x, unsigned d)
{
  const uint64_t hi = x >> 33, lo = x & 0x1FFFFFFFFULL;
  const unsigned dh = d % 31, dl = d % 33;
  const uint64_t h2 = dh ? ((hi << dh) | (hi >> (31 - dh))) & 0x7FFFFFFFULL : hi;
  const uint64_t l2 = dl ? ((lo << dl) | (lo >> (33 - dl))) & 0x1FFFFFFFFULL : lo;
  return (h2 << 33) | l2;
}

GRB_PHD uint64_t
grb_p_mulhi(uint64_t a, uint64_t b)
{
#ifdef __CUDA_ARCH__
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// x % m with inv = floor(2^64 / m) (exact: one conditional correction)
GRB_PHD uint64_t
grb_p_mod(uint64_t x, uint64_t m, uint64_t inv)
{
  uint64_t r = x - grb_p_mulhi(x, inv) * m;
  return r >= m ? r - m : r;
}

// Runs one (batch, k) job to the end.  Returns 0, or -1 if a read carries a threshold below 4
// (utils.cpp:105-107).
GRB_PHD int
grb_polish_run(const GrbPolishJob& j)
{
  const unsigned k = j.k, h = j.hash_num;
  if (k == 0 || h == 0 || h > 8) {
    return -1;
  }
  for (uint64_t r = j.first; r < j.last; ++r) {
    if (j.thr[r] < 4) {
      return -1;
    }
    const unsigned thr = j.thr[r] - 2 + j.k_index; // utils.cpp:108,121: one more per k value
    const unsigned thr8 = thr > 255 ? 255 : thr;
    const char* seq = j.seqs + j.off[r];
    const uint64_t len = j.off[r + 1] - j.off[r];
    const uint64_t out_seed_rot[4] = { grb_p_srol(grb_p_seed(0), k), grb_p_srol(grb_p_seed(1), k),
                                       grb_p_srol(grb_p_seed(2), k), grb_p_srol(grb_p_seed(3), k) };
    uint64_t fh = 0, rh = 0, run = 0; // run = valid characters ending at i (capped at k)
    for (uint64_t i = 0; i < len; ++i) {
      const int in = grb_p_code((unsigned char)seq[i]);
      if (in < 0) {
        run = 0;
        fh = rh = 0;
        continue;
      }
      if (run < k) {
        // still filling the first window after the start or a bad character (NTF64 / NTR64 from
        // scratch, nthash.hpp:100-119: the forward hash shifts the new base in, the reverse hash
        // XORs the complement rotated by its position)
        fh = grb_p_srol1(fh) ^ grb_p_seed(in);
        rh ^= grb_p_srol(grb_p_seed(3 - in), (unsigned)run);
        ++run;
        if (run < k) {
          continue;
        }
      } else {
        const int out = grb_p_code((unsigned char)seq[i - k]);
        fh = grb_p_srol1(fh) ^ grb_p_seed(in) ^ out_seed_rot[out];                      // nthash.hpp:122-131
        rh = grb_p_sror1(rh ^ grb_p_seed(3 - out) ^ out_seed_rot[3 - in]);              // nthash.hpp:143-152
      }
      const uint64_t base = fh + rh; // NTC64, nthash.hpp:172-191
      uint64_t idx[8];
      idx[0] = base;
      for (unsigned q = 1; q < h; ++q) { // NTMC64, nthash.hpp:262-300
        uint64_t t = base * (q ^ k * 0x90b45d39fb6da1faULL);
        t ^= t >> 27;
        idx[q] = t;
      }
      // KmerCountingBloomFilter8::insert_thresh_contains (recalled, oracle/shim_polish)
      uint8_t count = 255;
      uint64_t at[8];
      for (unsigned q = 0; q < h; ++q) {
        at[q] = grb_p_mod(idx[q], j.cbf_bytes, j.cbf_inv);
        const uint8_t c = j.cbf[at[q]];
        count = c < count ? c : count;
      }
      unsigned after = count;
      if (count < thr8) {
        for (unsigned q = 0; q < h; ++q) {
          if (j.cbf[at[q]] == count) {
            j.cbf[at[q]] = (uint8_t)(count + 1);
          }
        }
        after = count + 1u;
      }
      if (after >= thr) { // utils.cpp:115-119
        for (unsigned q = 0; q < h; ++q) {
          const uint64_t pos = grb_p_mod(idx[q], j.bf_bits, j.bf_inv);
          j.bf[pos >> 3] |= (uint8_t)(1u << (pos & 7));
        }
      }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Table-driven hashing for the warp kernel (kernels_polish.cuh): the read is 2-bit packed (32 bases
// per 64-bit word, base i of a word at bits 2i, 2i + 1; a parallel mask word holds the characters
// outside ACGTacgt), a k-mer's 2k bits are cut out of the packed words, and its forward / reverse
// hashes are XORs of one table entry per group of 4 bases:
//   T[g][v].x = XOR over i < 4, 4g + i < k of srol^(k - 1 - (4g + i))(seed[b_i])        (NTF64)
//   T[g][v].y = XOR over i < 4, 4g + i < k of srol^(4g + i)(seed[3 - b_i])              (NTR64)
// with b_i = (v >> 2i) & 3: ceil(k / 4) lookups instead of k rotate-and-xor steps per strand.
// ---------------------------------------------------------------------------------------------
#define GRB_P_MAX_K 64
#define GRB_P_GROUPS (GRB_P_MAX_K / 4)

struct GrbPolishPair
{
  uint64_t x, y;
};

// host: the table of one k, GRB_P_GROUPS * 256 entries (groups beyond ceil(k / 4) are zero)
inline void
grb_p_build_table(unsigned k, GrbPolishPair* T)
{
  for (unsigned g = 0; g < GRB_P_GROUPS; ++g) {
    for (unsigned v = 0; v < 256; ++v) {
      uint64_t f = 0, r = 0;
      for (unsigned i = 0; i < 4; ++i) {
        const unsigned pos = 4 * g + i;
        if (pos < k) {
          const int b = (int)((v >> (2 * i)) & 3u);
          f ^= grb_p_srol(grb_p_seed(b), k - 1 - pos);
          r ^= grb_p_srol(grb_p_seed(3 - b), pos);
        }
      }
      T[g * 256 + v] = GrbPolishPair{ f, r };
    }
  }
}

// 32 characters -> packed codes + mask of the characters outside ACGTacgt (n < 32 at the read's end)
GRB_PHD void
grb_p_pack32(const char* s, unsigned n, uint64_t* codes, uint32_t* bad)
{
  uint64_t c = 0;
  uint32_t m = 0;
  for (unsigned i = 0; i < n; ++i) {
    const int b = grb_p_code((unsigned char)s[i]);
    c |= (uint64_t)(b < 0 ? 0 : b) << (2 * i);
    m |= (b < 0 ? 1u : 0u) << i;
  }
  *codes = c;
  *bad = m;
}

// k-mer starting at local position q of a packed segment: false if it holds a bad character, else
// its canonical hash (fh + rh)
GRB_PHD bool
grb_p_hash_packed(const uint64_t* codes, const uint32_t* bad, unsigned q, unsigned k, const GrbPolishPair* T,
                  uint64_t* base)
{
  const unsigned w = q >> 5, o = q & 31;
  // k mask bits from bit o of bad[w..w+2]
  const uint64_t m01 = (uint64_t)bad[w] | ((uint64_t)bad[w + 1] << 32);
  uint64_t mk = m01 >> o;
  if (o) {
    mk |= (uint64_t)bad[w + 2] << (64 - o);
  }
  if (k < 64) {
    mk &= (1ull << k) - 1ull;
  }
  if (mk) {
    return false;
  }
  // 2k code bits from bit 2o of codes[w..w+2]
  uint64_t lo = codes[w] >> (2 * o), hi = codes[w + 1] >> (2 * o);
  if (o) {
    lo |= codes[w + 1] << (64 - 2 * o);
    hi |= codes[w + 2] << (64 - 2 * o);
  }
  uint64_t fh = 0, rh = 0;
  const unsigned groups = (k + 3) / 4;
  for (unsigned g = 0; g < groups; ++g) {
    const unsigned v = (unsigned)((g < 8 ? lo >> (8 * g) : hi >> (8 * (g - 8))) & 0xFFu);
    const GrbPolishPair t = T[g * 256 + v];
    fh ^= t.x;
    rh ^= t.y;
  }
  *base = fh + rh;
  return true;
}
