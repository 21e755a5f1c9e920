// K2 + K4a (hash whole reads into the bit vector), K4b (rank build), and the small kernels behind
// the parity exports.  Replaces MIBFConstructSupport::insertBV / setup
// (goldrush_path/MIBFConstructSupport.hpp:134-147,165-170) as driven by fill_bit_vector
// (goldrush_path/goldrush_path.cpp:302-305).
#pragma once
#include "common.cuh"
#include "kernels_decode.cuh"

#define GRB_FILL_CHUNK 2048 // read positions per CTA in the fill kernel (256 threads x 8)

__device__ __forceinline__ void
grb_set_bit(const GrbFilterDev& f, uint64_t hash)
{
  const uint64_t pos = grb_fastmod(hash, f.bits, f.inv);
  const uint64_t blk = grb_div3(pos >> 6);
  const unsigned r = (unsigned)(pos - blk * GRB_BLK_BITS);
  atomicOr(reinterpret_cast<unsigned long long*>(f.blocks + blk * 4 + 1 + (r >> 6)),
           1ull << (r & 63));
}

// Each CTA hashes GRB_FILL_CHUNK consecutive positions of one read for all h patterns.
//   chunk_read[c]  read index of chunk c;  chunk_first[read] = first chunk of that read.
// Stale-tail frames repeat the last valid hash of the longer patterns
// (multiLensfrHashIterator.hpp:49-68): setting the same bit twice is idempotent, so each
// (position, pattern) is hashed exactly once here.
__global__ void __launch_bounds__(256)
k_fill_bits(GrbReadsDev reads, const GrbSeedTables* __restrict__ seeds_g, GrbFilterDev filt,
            const uint32_t* __restrict__ chunk_read, const uint64_t* __restrict__ chunk_first,
            uint64_t n_chunks)
{
  __shared__ GrbSeedTables st;
  __shared__ uint64_t sw[GRB_FILL_CHUNK / 32 + 8];
  __shared__ uint64_t s_fr[GRB_FILL_CHUNK + GRB_MAX_SPAN];
  __shared__ uint64_t s_rr[GRB_FILL_CHUNK + GRB_MAX_SPAN];
  for (unsigned i = threadIdx.x; i < sizeof(GrbSeedTables) / 8; i += blockDim.x) {
    reinterpret_cast<uint64_t*>(&st)[i] = reinterpret_cast<const uint64_t*>(seeds_g)[i];
  }
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const uint32_t r = chunk_read[c];
    const uint32_t len = reads.len[r];
    const uint32_t p0 = (uint32_t)(c - chunk_first[r]) * GRB_FILL_CHUNK; // first position
    const uint64_t w_read = reads.word_off[r];
    const uint32_t w_first = p0 >> 5;
    const uint32_t w_total = (len + 31) / 32;
    __syncthreads();
    // positions p0 .. p0 + CHUNK + span need words w_first .. w_first + CHUNK/32 + 3
    for (unsigned i = threadIdx.x; i < GRB_FILL_CHUNK / 32 + 8; i += blockDim.x) {
      sw[i] = (w_first + i < w_total) ? reads.bases[w_read + w_first + i] : 0ull;
    }
    __syncthreads();
    const unsigned half = st.half, k = st.k, h = st.h;
    // pass A: half hashes for positions [p0, p0 + CHUNK + half + h)
    const unsigned n_half = GRB_FILL_CHUNK + half + h;
    uint64_t fl[GRB_FILL_CHUNK / 256], rl[GRB_FILL_CHUNK / 256];
#pragma unroll
    for (unsigned it = 0; it < GRB_FILL_CHUNK / 256; ++it) {
      const unsigned j = it * 256 + threadIdx.x;
      const GrbWindow w =
        grb_window([&](uint64_t wi) { return sw[wi]; }, (uint64_t)(p0 & 31) + j);
      const GrbHalf hh = grb_half_hashes(st, w);
      fl[it] = hh.fl;
      rl[it] = hh.rl;
      s_fr[j] = hh.fr;
      s_rr[j] = hh.rr;
    }
    for (unsigned j = GRB_FILL_CHUNK + threadIdx.x; j < n_half; j += 256) {
      const GrbWindow w =
        grb_window([&](uint64_t wi) { return sw[wi]; }, (uint64_t)(p0 & 31) + j);
      const GrbHalf hh = grb_half_hashes(st, w);
      s_fr[j] = hh.fr;
      s_rr[j] = hh.rr;
    }
    __syncthreads();
    // pass B: combine and set bits
#pragma unroll
    for (unsigned it = 0; it < GRB_FILL_CHUNK / 256; ++it) {
      const unsigned j = it * 256 + threadIdx.x;
      const uint64_t pos = (uint64_t)p0 + j;
      for (unsigned i = 0; i < h; ++i) {
        if (pos + k + i <= len) {
          const uint64_t hv = grb_combine(i, fl[it], rl[it], s_fr[j + half + i], s_rr[j + half + i]);
          grb_set_bit(filt, hv);
        }
      }
    }
  }
}

// ---- rank build: per-block popcounts -> exclusive scan -> word 0 of every block ----
#define GRB_RANK_ITEMS 8

__global__ void __launch_bounds__(256)
k_rank_partial(const uint64_t* __restrict__ blocks, uint64_t n_blocks,
               uint32_t* __restrict__ partial)
{
  const uint64_t base = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * GRB_RANK_ITEMS;
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < GRB_RANK_ITEMS; ++i) {
    const uint64_t b = base + i;
    if (b < n_blocks) {
      const ulonglong2* p = reinterpret_cast<const ulonglong2*>(blocks + b * 4);
      const ulonglong2 x = p[0], y = p[1];
      c += __popcll(x.y) + __popcll(y.x) + __popcll(y.y);
    }
  }
  uint32_t total;
  grb_block_excl_scan<256>(c, &total);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = total;
  }
}

__global__ void __launch_bounds__(256)
k_rank_write(uint64_t* __restrict__ blocks, uint64_t n_blocks,
             const uint64_t* __restrict__ partial_off)
{
  const uint64_t base = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * GRB_RANK_ITEMS;
  uint32_t pc[GRB_RANK_ITEMS];
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < GRB_RANK_ITEMS; ++i) {
    const uint64_t b = base + i;
    pc[i] = 0;
    if (b < n_blocks) {
      const ulonglong2* p = reinterpret_cast<const ulonglong2*>(blocks + b * 4);
      const ulonglong2 x = p[0], y = p[1];
      pc[i] = __popcll(x.y) + __popcll(y.x) + __popcll(y.y);
    }
    c += pc[i];
  }
  uint32_t total;
  uint64_t run = partial_off[blockIdx.x] + grb_block_excl_scan<256>(c, &total);
#pragma unroll
  for (int i = 0; i < GRB_RANK_ITEMS; ++i) {
    const uint64_t b = base + i;
    if (b < n_blocks) {
      blocks[b * 4] = run;
      run += pc[i];
    }
  }
}

// dst |= src over 8-byte words (multi-GPU OR-reduce of partial bit vectors).  Word 0 of each
// block is still zero at that stage, so OR-ing whole blocks is exact.
__global__ void
k_or_words(uint64_t* __restrict__ dst, const uint64_t* __restrict__ src, uint64_t n)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    dst[i] |= src[i];
  }
}

// ---- parity exports ----
// interleaved blocks -> plain LSB-first words (sdsl::bit_vector layout)
__global__ void
k_export_plain(const uint64_t* __restrict__ blocks, uint64_t n_words, uint64_t* __restrict__ out)
{
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words;
       w += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t blk = w / 3;
    out[w] = blocks[blk * 4 + 1 + (w - blk * 3)];
  }
}

__global__ void
k_import_plain(uint64_t* __restrict__ blocks, uint64_t n_words, const uint64_t* __restrict__ in)
{
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words;
       w += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t blk = w / 3;
    blocks[blk * 4 + 1 + (w - blk * 3)] = in[w];
  }
}

__global__ void
k_rank_query(GrbFilterDev f, const uint64_t* __restrict__ pos, uint64_t n,
             uint64_t* __restrict__ rank, uint8_t* __restrict__ bit)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    bool b;
    uint64_t r;
    grb_probe_block(f, pos[i], b, r);
    rank[i] = r;
    bit[i] = b ? 1 : 0;
  }
}

__global__ void
k_get_slots(const GrbSlot* __restrict__ slots, const uint64_t* __restrict__ rank, uint64_t n,
            uint32_t* __restrict__ ids, uint32_t* __restrict__ counts)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const GrbSlot s = slots[rank[i]];
    ids[i] = s.id;
    counts[i] = s.count;
  }
}

__global__ void
k_set_slots(GrbSlot* __restrict__ slots, const uint64_t* __restrict__ rank, uint64_t n,
            const uint32_t* __restrict__ ids, const uint32_t* __restrict__ counts)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    slots[rank[i]] = GrbSlot{ ids[i], counts[i], ids[i], 0u };
  }
}

// multiLensfrHashIterator over one packed sequence: out[frame*h + p], stale tail included
__global__ void
k_hash_sequence(const uint64_t* __restrict__ words, uint32_t len,
                const GrbSeedTables* __restrict__ seeds, uint64_t* __restrict__ out)
{
  const GrbSeedTables& st = *seeds;
  const uint32_t frames = len - st.k + 1;
  for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < frames;
       f += gridDim.x * blockDim.x) {
    for (unsigned i = 0; i < st.h; ++i) {
      const uint32_t n_i = len - (st.k + i) + 1;
      const uint32_t p = f < n_i ? f : n_i - 1;
      const GrbWindow w = grb_window([&](uint64_t wi) { return words[wi]; }, p);
      out[(uint64_t)f * st.h + i] = grb_hash_direct(st, i, w);
    }
  }
}
