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
