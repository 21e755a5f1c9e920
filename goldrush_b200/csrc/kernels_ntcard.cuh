// K5 — ntCard-style distinct-seed estimate that sizes the filter (--ntcard).
// Replaces stRead / ntComp / compEst of goldrush_path/ntcard.hpp:81-154 as driven by getHist
// (:156-246): every read, whole and unfiltered, every pattern; a hash is sampled when its top
// bits match one of two fixed prefixes and bumps a counter indexed by its low 27 bits.
//
// The reference counters are uint16_t and wrap; only "how many counters read zero" reaches the
// estimate (ntcard.hpp:127,138-139), so the device keeps 32-bit counters and tests
// (count & 0xFFFF) == 0 at the end, which is exact as long as no bucket receives 2^32 hits.
// The stale-tail repeats of multiLensfrHashIterator (the last hash of an exhausted stream is
// re-counted while other streams still roll, ntcard.hpp:103-110) are added by k_ntcard_tail.
#pragma once
#include "common.cuh"
#include "kernels_filter.cuh"

__device__ __forceinline__ void
grb_ntcard_sample(uint64_t hv, unsigned pattern, unsigned sBits, unsigned rBits,
                  uint32_t* __restrict__ counters, uint32_t times)
{
  const uint64_t rBuck = 1ull << rBits;
  const uint64_t sMask = (1ull << (sBits - 1)) - 1;
  unsigned ind = 2;
  if ((hv >> (63 - sBits)) == 1) {
    ind = 0;
  }
  if ((hv >> (64 - sBits)) == sMask) {
    ind = 1;
  }
  if (ind < 2) {
    atomicAdd(&counters[((uint64_t)pattern * 2 + ind) * rBuck + (hv & (rBuck - 1))], times);
  }
}

// true when no base of [pos, pos + span) was a non-ACGT byte
__device__ __forceinline__ bool
grb_window_clean(const uint32_t* __restrict__ nmask, uint64_t word_off, uint32_t pos, uint32_t span)
{
  // 1 bit per base, 32 bases per word, same indexing as the packed bases
  const uint32_t w0 = pos >> 5, w1 = (pos + span - 1) >> 5;
  for (uint32_t w = w0; w <= w1; ++w) {
    uint32_t m = nmask[word_off + w];
    if (w == w0) {
      m &= 0xFFFFFFFFu << (pos & 31);
    }
    if (w == w1) {
      const uint32_t e = (pos + span - 1) & 31;
      m &= e == 31 ? 0xFFFFFFFFu : ((1u << (e + 1)) - 1u);
    }
    if (m) {
      return false;
    }
  }
  return true;
}

__global__ void __launch_bounds__(256)
k_ntcard_count(GrbReadsDev reads, const GrbSeedTables* __restrict__ seeds_g,
               const uint32_t* __restrict__ chunk_read, const uint64_t* __restrict__ chunk_first,
               uint64_t n_chunks, uint32_t* __restrict__ counters, uint32_t* __restrict__ valid,
               unsigned sBits, unsigned rBits)
{
  __shared__ GrbSeedTables st;
  __shared__ uint64_t sw[GRB_FILL_CHUNK / 32 + 8];
  for (unsigned i = threadIdx.x; i < sizeof(GrbSeedTables) / 8; i += blockDim.x) {
    reinterpret_cast<uint64_t*>(&st)[i] = reinterpret_cast<const uint64_t*>(seeds_g)[i];
  }
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const uint32_t r = chunk_read[c];
    const uint32_t len = reads.len[r];
    const bool dirty = reads.flags[r] & 4u;
    const uint32_t p0 = (uint32_t)(c - chunk_first[r]) * GRB_FILL_CHUNK;
    const uint64_t w_read = reads.word_off[r];
    const uint32_t w_first = p0 >> 5;
    const uint32_t w_total = (len + 31) / 32;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < GRB_FILL_CHUNK / 32 + 8; i += blockDim.x) {
      sw[i] = (w_first + i < w_total) ? reads.bases[w_read + w_first + i] : 0ull;
    }
    __syncthreads();
    const unsigned k = st.k, h = st.h;
    for (unsigned j = threadIdx.x; j < GRB_FILL_CHUNK; j += 256) {
      const uint32_t pos = p0 + j;
      const GrbWindow w = grb_window([&](uint64_t wi) { return sw[wi]; }, (uint64_t)j);
      for (unsigned i = 0; i < h; ++i) {
        if ((uint64_t)pos + k + i > len) {
          continue;
        }
        if (dirty) {
          if (!grb_window_clean(reads.nmask, w_read, pos, k + i)) {
            continue;
          }
          atomicAdd(&valid[(uint64_t)r * h + i], 1u);
        }
        grb_ntcard_sample(grb_hash_direct(st, i, w), i, sBits, rBits, counters, 1u);
      }
    }
  }
}

// one thread per (read, pattern): the repeats of the last valid window of an exhausted stream
__global__ void
k_ntcard_tail(GrbReadsDev reads, const GrbSeedTables* __restrict__ seeds_g, uint64_t n_reads,
              uint32_t* __restrict__ counters, const uint32_t* __restrict__ valid, unsigned sBits,
              unsigned rBits)
{
  const GrbSeedTables& st = *seeds_g;
  const unsigned h = st.h, k = st.k;
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_reads * h) {
    return;
  }
  const uint64_t r = gid / h;
  const unsigned i = (unsigned)(gid - r * h);
  const uint32_t len = reads.len[r];
  const bool dirty = reads.flags[r] & 4u;
  uint32_t frames = 1, mine = 0;
  for (unsigned j = 0; j < h; ++j) {
    const uint32_t v = dirty ? valid[r * h + j] : len - (k + j) + 1;
    frames = v > frames ? v : frames;
    if (j == i) {
      mine = v;
    }
  }
  if (mine == 0 || frames == mine) {
    return; // an uninitialised stream reports hash 0, which is never sampled
  }
  const uint64_t w_read = reads.word_off[r];
  uint32_t pos = len - (k + i);
  if (dirty) {
    while (!grb_window_clean(reads.nmask, w_read, pos, k + i)) {
      --pos;
    }
  }
  const GrbWindow w = grb_window([&](uint64_t wi) { return reads.bases[w_read + wi]; }, pos);
  grb_ntcard_sample(grb_hash_direct(st, i, w), i, sBits, rBits, counters, frames - mine);
}

__global__ void __launch_bounds__(256)
k_ntcard_zeros(const uint32_t* __restrict__ counters, uint64_t n_tables, uint64_t rBuck,
               unsigned long long* __restrict__ zeros)
{
  for (uint64_t t = 0; t < n_tables; ++t) {
    unsigned long long z = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rBuck;
         i += (uint64_t)gridDim.x * blockDim.x) {
      z += (counters[t * rBuck + i] & 0xFFFFu) == 0u ? 1ull : 0ull;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      z += __shfl_xor_sync(0xffffffffu, z, d);
    }
    if ((threadIdx.x & 31) == 0 && z) {
      atomicAdd(&zeros[t], z);
    }
  }
}
