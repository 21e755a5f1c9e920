// (f4) GoldPolish targeted Bloom filters on the device: one thread per (batch, k) job of
// polish_core.h, thousands of jobs per launch.  See polish_core.h for what it replaces and why a
// job is sequential.
#pragma once
#include "polish_core.h"
#include <cuda_runtime.h>

struct GrbPolishWave
{
  const char* seqs;
  const uint64_t* off;
  const uint32_t* thr;
  const uint64_t* batch_first; // [n_batches + 1]
  const uint32_t* k_values;    // [n_k]
  uint32_t n_k, hash_num;
  uint32_t job0, n_jobs;       // jobs [job0, job0 + n_jobs) of the call: job = batch * n_k + k_index
  uint8_t* cbf;                // n_jobs counting filters of cbf_bytes each, zeroed
  uint8_t* bf;                 // n_jobs Bloom filters of bf_bytes each, zeroed
  uint64_t cbf_bytes, cbf_inv, bf_bytes, bf_inv;
  int* status;
};

__global__ void __launch_bounds__(32)
k_polish_fill(GrbPolishWave w)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= w.n_jobs) {
    return;
  }
  const uint32_t job = w.job0 + t;
  const uint32_t batch = job / w.n_k, ki = job - batch * w.n_k;
  GrbPolishJob j;
  j.seqs = w.seqs;
  j.off = w.off;
  j.thr = w.thr;
  j.first = w.batch_first[batch];
  j.last = w.batch_first[batch + 1];
  j.k = w.k_values[ki];
  j.k_index = ki;
  j.hash_num = w.hash_num;
  j.cbf = w.cbf + (uint64_t)t * w.cbf_bytes;
  j.cbf_bytes = w.cbf_bytes;
  j.cbf_inv = w.cbf_inv;
  j.bf = w.bf + (uint64_t)t * w.bf_bytes;
  j.bf_bits = w.bf_bytes * 8;
  j.bf_inv = w.bf_inv;
  if (grb_polish_run(j) != 0) {
    atomicExch(w.status, -1);
  }
}

// ---------------------------------------------------------------------------------------------
// One WARP per (batch, k) job: 32 consecutive k-mers of a read at a time.
//
// The order dependence of the counting filter (polish_core.h) only bites when two k-mers of the
// group share a counter: if the 32 x hash_num counter indices of a group are pairwise distinct
// ACROSS lanes, each k-mer's count and update depend on counters nobody else in the group touches,
// so the 32 updates commute and are applied at once; if any index is shared (identical k-mers in a
// low-complexity stretch, or a chance collision: 128^2 / 2 / 10^7 = 0.08 % of the groups at the
// reference's filter size) the group is replayed in lane order.  Groups follow one another in
// order, reads in order: the result is the sequential one.  Exact detection: the group's indices
// go into a 256-entry open-addressing table in shared memory, (index, lane) per entry.
// Bloom-filter bits are ORs and commute; they are set with 32-bit atomics because two lanes may
// hit one word.
// ---------------------------------------------------------------------------------------------
#define GRB_PW_TAB 256 // per-warp conflict table entries (128 indices at hash_num = 4)
#define GRB_PW_WARPS 4

__device__ __forceinline__ void
grb_pw_update(const GrbPolishWave& w, uint8_t* cbf, uint32_t* bf32, unsigned h, const uint64_t* at,
              const uint64_t* idx, unsigned thr, unsigned thr8, uint64_t bf_bits)
{
  uint8_t count = 255;
  for (unsigned q = 0; q < h; ++q) {
    const uint8_t c = cbf[at[q]];
    count = c < count ? c : count;
  }
  unsigned after = count;
  if (count < thr8) {
    for (unsigned q = 0; q < h; ++q) {
      if (cbf[at[q]] == count) {
        cbf[at[q]] = (uint8_t)(count + 1);
      }
    }
    after = count + 1u;
  }
  if (after >= thr) {
    for (unsigned q = 0; q < h; ++q) {
      const uint64_t pos = grb_p_mod(idx[q], bf_bits, w.bf_inv);
      atomicOr(&bf32[pos >> 5], 1u << (pos & 31)); // little-endian: bit (pos & 7) of byte pos >> 3
    }
  }
}

// A read is handled in segments of GRB_PW_SEG k-mer positions: the warp packs the segment's bases
// (2 bits each + a mask of the characters outside ACGTacgt) into shared memory once, and every
// k-mer of the segment is hashed from the packed words through the per-k tables of polish_core.h.
#define GRB_PW_SEG 1024
#define GRB_PW_SEG_WORDS ((GRB_PW_SEG + GRB_P_MAX_K) / 32 + 3)

__global__ void __launch_bounds__(32 * GRB_PW_WARPS)
k_polish_fill_warp(GrbPolishWave w, const GrbPolishPair* __restrict__ tables)
{
  __shared__ unsigned long long s_tab[GRB_PW_WARPS][GRB_PW_TAB];
  __shared__ uint64_t s_codes[GRB_PW_WARPS][GRB_PW_SEG_WORDS];
  __shared__ uint32_t s_bad[GRB_PW_WARPS][GRB_PW_SEG_WORDS];
  const unsigned lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const uint32_t t = blockIdx.x * GRB_PW_WARPS + wi;
  if (t >= w.n_jobs) {
    return;
  }
  unsigned long long* tab = s_tab[wi];
  uint64_t* codes = s_codes[wi];
  uint32_t* bad = s_bad[wi];
  const uint32_t job = w.job0 + t;
  const uint32_t batch = job / w.n_k, ki = job - batch * w.n_k;
  const unsigned k = w.k_values[ki], h = w.hash_num;
  const GrbPolishPair* T = tables + (size_t)ki * GRB_P_GROUPS * 256;
  uint8_t* cbf = w.cbf + (uint64_t)t * w.cbf_bytes;
  uint32_t* bf32 = reinterpret_cast<uint32_t*>(w.bf + (uint64_t)t * w.bf_bytes);
  const uint64_t bf_bits = w.bf_bytes * 8;
  const uint64_t first = w.batch_first[batch], last = w.batch_first[batch + 1];
  for (uint64_t r = first; r < last; ++r) {
    const unsigned thr_in = w.thr[r];
    if (thr_in < 4) {
      if (lane == 0) {
        atomicExch(w.status, -1);
      }
      return;
    }
    const unsigned thr = thr_in - 2 + ki, thr8 = thr > 255 ? 255 : thr;
    const char* seq = w.seqs + w.off[r];
    const uint64_t len = w.off[r + 1] - w.off[r];
    if (len < k) {
      continue;
    }
    const uint64_t n_pos = len - k + 1;
    for (uint64_t seg0 = 0; seg0 < n_pos; seg0 += GRB_PW_SEG) {
      // ---- pack the bases this segment's k-mers cover ----
      const uint64_t seg_pos = n_pos - seg0 < GRB_PW_SEG ? n_pos - seg0 : GRB_PW_SEG;
      const uint64_t seg_bases = seg_pos + k - 1;
      __syncwarp();
      for (unsigned wd = lane; wd < GRB_PW_SEG_WORDS; wd += 32) {
        const uint64_t b0 = (uint64_t)wd * 32;
        uint64_t c = 0;
        uint32_t m = 0;
        if (b0 < seg_bases) {
          const uint64_t n = seg_bases - b0 < 32 ? seg_bases - b0 : 32;
          grb_p_pack32(seq + seg0 + b0, (unsigned)n, &c, &m);
        }
        codes[wd] = c;
        bad[wd] = m;
      }
      __syncwarp();
      for (uint64_t g0 = 0; g0 < seg_pos; g0 += 32) {
        const uint64_t q = g0 + lane;
        uint64_t base = 0;
        const bool valid = q < seg_pos && grb_p_hash_packed(codes, bad, (unsigned)q, k, T, &base);
        uint64_t idx[8], at[8];
        if (valid) {
          idx[0] = base;
          for (unsigned u = 1; u < h; ++u) {
            uint64_t x = base * (u ^ k * 0x90b45d39fb6da1faULL);
            x ^= x >> 27;
            idx[u] = x;
          }
          for (unsigned u = 0; u < h; ++u) {
            at[u] = grb_p_mod(idx[u], w.cbf_bytes, w.cbf_inv);
          }
        }
        // ---- do two lanes of the group share a counter? ----
        for (unsigned i = lane; i < GRB_PW_TAB; i += 32) {
          tab[i] = ~0ull;
        }
        __syncwarp();
        bool conflict = false;
        if (valid) {
          for (unsigned u = 0; u < h; ++u) {
            const unsigned long long mine = ((unsigned long long)at[u] << 8) | lane;
            unsigned slot = (unsigned)((at[u] * 0x9E3779B97F4A7C15ull) >> 56) & (GRB_PW_TAB - 1);
            for (unsigned tries = 0; tries < GRB_PW_TAB; ++tries) {
              const unsigned long long old = atomicCAS(&tab[slot], ~0ull, mine);
              if (old == ~0ull) {
                break;
              }
              if ((old >> 8) == at[u]) {
                conflict = conflict || (unsigned)(old & 0xFF) != lane;
                break;
              }
              slot = (slot + 1) & (GRB_PW_TAB - 1);
            }
          }
        }
        const bool any_conflict = __any_sync(0xffffffffu, conflict) || h * 32 > GRB_PW_TAB / 2;
        if (!any_conflict) {
          if (valid) {
            grb_pw_update(w, cbf, bf32, h, at, idx, thr, thr8, bf_bits);
          }
          __syncwarp();
        } else {
          for (unsigned l = 0; l < 32; ++l) { // lane order = k-mer order
            if (l == lane && valid) {
              grb_pw_update(w, cbf, bf32, h, at, idx, thr, thr8, bf_bits);
            }
            __syncwarp();
          }
        }
      }
    }
  }
}
