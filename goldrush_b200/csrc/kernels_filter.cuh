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

// ---- partitioned fill: the random read-modify-writes of pass 1 made L2 resident -------------------
// k_fill_bits above is bound by random atomics over the whole bit vector (measured on B200: about
// 25 G atomicOr/s once the footprint leaves L2, 190 G/s inside it, profiles/sector_roofline_r01.json).
// Pass 1 has no order dependence (MIBFConstructSupport.hpp:134-147 is a commutative OR), so the
// positions of a round of reads are first bucketed by filter partition (2^pshift bits, about 22 MB
// of blocks: a fraction of the 126 MB L2) and then applied partition by partition:
//   k_fill_part   CTA per 2048 read positions: hash, bucket in shared memory, append each bucket
//                 to its partition's list with one global reservation per (CTA, partition) and
//                 coalesced 4-byte offsets                               [streaming writes]
//   k_fill_apply  partition-major sweep over the lists: atomicOr into blocks that stay in L2
// A list that overflows its capacity falls back to the direct atomicOr (always correct).
#define GRB_PART_H 4         // patterns the register-resident bucketing is unrolled for
#define GRB_APPLY_TILE 8192u // list entries per CTA step of k_fill_apply

struct GrbFillPart
{
  uint32_t* lists;   // [n_part * cap] offsets within the partition
  uint32_t* cursor;  // [n_part] entries reserved (may exceed cap: the excess went the direct way);
                     // [n_part] = ticket counter of k_fill_apply
  uint32_t n_part;
  uint32_t pshift;   // partition = pos >> pshift
  uint32_t cap;      // list capacity, entries
  uint32_t pad;
};

__device__ __forceinline__ void
grb_set_bit_pos(const GrbFilterDev& f, uint64_t pos)
{
  const uint64_t blk = grb_div3(pos >> 6);
  const unsigned r = (unsigned)(pos - blk * GRB_BLK_BITS);
  atomicOr(reinterpret_cast<unsigned long long*>(f.blocks + blk * 4 + 1 + (r >> 6)),
           1ull << (r & 63));
}

// Dynamic shared memory: ulonglong2 gL[ng * 256] | gR[ng * 256] | uint32 stage[SUB * h] |
// sdst[SUB * h] | hist[n_part] | sbase[n_part + 1] | gbase[n_part]
// A CTA walks its 2048-position chunk in sub-chunks of SUB positions (BS threads, SUB / BS
// positions per thread), so that two or three CTAs fit on an SM and overlap each other's barriers.
template<int BS, int SUB, int MINB>
__global__ void __launch_bounds__(BS, MINB)
k_fill_part(GrbReadsDev reads, const GrbSeedTables* __restrict__ seeds_g,
            const ulonglong2* __restrict__ gtab, uint32_t ng, GrbFilterDev filt,
            const uint32_t* __restrict__ chunk_read, const uint64_t* __restrict__ chunk_first,
            uint64_t chunk0, uint64_t n_chunks, GrbFillPart fp)
{
  constexpr int PER = SUB / BS;
  static_assert(GRB_FILL_CHUNK % SUB == 0 && SUB % BS == 0, "sub-chunks tile the chunk");
  __shared__ uint64_t sw[SUB / 32 + 8];
  __shared__ uint64_t s_fr[SUB + GRB_MAX_SPAN];
  __shared__ uint64_t s_rr[SUB + GRB_MAX_SPAN];
  __shared__ uint32_t s_total;
  extern __shared__ __align__(16) unsigned char fill_dyn[];
  ulonglong2* gL = reinterpret_cast<ulonglong2*>(fill_dyn);
  ulonglong2* gR = gL + ng * 256;
  for (unsigned i = threadIdx.x; i < 2 * ng * 256; i += BS) {
    gL[i] = gtab[i];
  }
  const unsigned half = seeds_g->half, k = seeds_g->k, h = seeds_g->h;
  const uint32_t P = fp.n_part;
  uint32_t* stage = reinterpret_cast<uint32_t*>(gR + ng * 256);
  uint32_t* sdst = stage + SUB * h;
  uint32_t* hist = sdst + SUB * h;
  uint32_t* sbase = hist + P;
  uint32_t* gbase = sbase + P + 1;
  const uint32_t omask = (1u << fp.pshift) - 1u;
  constexpr uint64_t SUBS = GRB_FILL_CHUNK / SUB;
  __syncthreads();

  for (uint64_t cs = (chunk0 * SUBS) + blockIdx.x; cs < (chunk0 + n_chunks) * SUBS; cs += gridDim.x) {
    const uint64_t c = cs / SUBS;
    const uint32_t r = chunk_read[c];
    const uint32_t len = reads.len[r];
    const uint32_t p0 =
      (uint32_t)(c - chunk_first[r]) * GRB_FILL_CHUNK + (uint32_t)(cs - c * SUBS) * SUB;
    if (p0 + k > len) { // nothing of this sub-chunk is a valid position (uniform over the CTA)
      continue;
    }
    const uint64_t w_read = reads.word_off[r];
    const uint32_t w_first = p0 >> 5;
    const uint32_t w_total = (len + 31) / 32;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < SUB / 32 + 8; i += BS) {
      sw[i] = (w_first + i < w_total) ? reads.bases[w_read + w_first + i] : 0ull;
    }
    for (unsigned i = threadIdx.x; i < P; i += BS) {
      hist[i] = 0;
    }
    __syncthreads();
    const unsigned n_half = SUB + half + h;
    uint64_t fl[PER], rl[PER];
#pragma unroll
    for (int it = 0; it < PER; ++it) {
      const unsigned j = it * BS + threadIdx.x;
      const uint64_t lo = grb_lo64([&](uint64_t wi) { return sw[wi]; }, (uint64_t)(p0 & 31) + j);
      const ulonglong2 l = grb_group_half(gL, ng, lo);
      const ulonglong2 rt = grb_group_half(gR, ng, lo);
      fl[it] = l.x;
      rl[it] = l.y;
      s_fr[j] = rt.x;
      s_rr[j] = rt.y;
    }
    for (unsigned j = SUB + threadIdx.x; j < n_half; j += BS) {
      const uint64_t lo = grb_lo64([&](uint64_t wi) { return sw[wi]; }, (uint64_t)(p0 & 31) + j);
      const ulonglong2 rt = grb_group_half(gR, ng, lo);
      s_fr[j] = rt.x;
      s_rr[j] = rt.y;
    }
    __syncthreads();
    // bucket: partition, offset and the rank inside this CTA's bucket stay in registers
    uint32_t off[PER][GRB_PART_H], plr[PER][GRB_PART_H]; // plr = partition << 16 | rank in bucket
#pragma unroll
    for (int it = 0; it < PER; ++it) {
      const unsigned j = it * BS + threadIdx.x;
      const uint64_t pos_in_read = (uint64_t)p0 + j;
#pragma unroll
      for (int i = 0; i < GRB_PART_H; ++i) {
        plr[it][i] = 0xFFFFFFFFu;
        if (i < (int)h && pos_in_read + k + i <= len) {
          const uint64_t hv = grb_combine(i, fl[it], rl[it], s_fr[j + half + i], s_rr[j + half + i]);
          const uint64_t pos = grb_fastmod(hv, filt.bits, filt.inv);
          const uint32_t p = (uint32_t)(pos >> fp.pshift);
          off[it][i] = (uint32_t)pos & omask;
          plr[it][i] = (p << 16) | atomicAdd(&hist[p], 1u); // at most SUB * 4 entries per bucket
        }
      }
    }
    __syncthreads();
    // local exclusive scan of the bucket sizes + one global reservation per non-empty bucket
    {
      uint32_t run = 0;
      for (uint32_t base = 0; base < P; base += BS) { // P > BS only for the 512-thread CTAs
        const uint32_t q = base + threadIdx.x;
        const uint32_t mine = q < P ? hist[q] : 0u;
        uint32_t total;
        const uint32_t ex = grb_block_excl_scan<BS>(mine, &total);
        if (q < P) {
          sbase[q] = run + ex;
          gbase[q] = mine ? atomicAdd(&fp.cursor[q], mine) : 0u;
        }
        run += total;
      }
      if (threadIdx.x == 0) {
        sbase[P] = run;
        s_total = run;
      }
    }
    __syncthreads();
    // scatter into the staging area, bucket by bucket, together with each entry's destination in
    // the partition lists (so that the copy-out below is a plain coalesced loop); entries past a
    // list's capacity take the direct atomicOr instead
#pragma unroll
    for (int it = 0; it < PER; ++it) {
#pragma unroll
      for (int i = 0; i < GRB_PART_H; ++i) {
        if (plr[it][i] != 0xFFFFFFFFu) {
          const uint32_t p = plr[it][i] >> 16, lr = plr[it][i] & 0xFFFFu;
          const uint32_t g = gbase[p] + lr;
          const uint32_t at = sbase[p] + lr;
          if (g < fp.cap) {
            stage[at] = off[it][i];
            sdst[at] = p * fp.cap + g;
          } else {
            sdst[at] = 0xFFFFFFFFu;
            grb_set_bit_pos(filt, ((uint64_t)p << fp.pshift) | off[it][i]);
          }
        }
      }
    }
    __syncthreads();
    const uint32_t total = s_total;
    for (uint32_t idx = threadIdx.x; idx < total; idx += BS) {
      const uint32_t d = sdst[idx];
      if (d != 0xFFFFFFFFu) {
        fp.lists[d] = stage[idx];
      }
    }
  }
}

// Partition-major sweep.  List tiles are handed out in (partition, offset) order through one global
// ticket counter (fp.cursor[n_part]), so at any moment the whole GPU works inside one or two
// partitions and their blocks stay in L2 (a static grid-stride let fast CTAs run partitions ahead:
// ncu showed 30 GB of DRAM traffic per 2 GB of lists).  Dynamic shared memory: uint32 pre[n_part + 1].
__global__ void __launch_bounds__(256)
k_fill_apply(GrbFilterDev filt, GrbFillPart fp)
{
  extern __shared__ uint32_t pre[]; // tiles before partition p
  __shared__ uint32_t s_ticket;
  const uint32_t P = fp.n_part;
  {
    uint32_t run = 0;
    for (uint32_t base = 0; base < P; base += 256) {
      const uint32_t q = base + threadIdx.x;
      const uint32_t n = q < P ? min(__ldg(&fp.cursor[q]), fp.cap) : 0u;
      const uint32_t tiles = (n + GRB_APPLY_TILE - 1) / GRB_APPLY_TILE;
      uint32_t total;
      const uint32_t ex = grb_block_excl_scan<256>(tiles, &total);
      if (q < P) {
        pre[q] = run + ex;
      }
      run += total;
    }
    if (threadIdx.x == 0) {
      pre[P] = run;
    }
  }
  __syncthreads();
  const uint32_t n_tiles = pre[P];
  for (;;) {
    if (threadIdx.x == 0) {
      s_ticket = atomicAdd(&fp.cursor[P], 1u);
    }
    __syncthreads();
    const uint32_t t = s_ticket;
    __syncthreads();
    if (t >= n_tiles) {
      break;
    }
    uint32_t lo = 0, hi = P; // last p with pre[p] <= t
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (pre[mid] <= t) {
        lo = mid;
      } else {
        hi = mid;
      }
    }
    const uint32_t p = lo;
    const uint32_t i0 = (t - pre[p]) * GRB_APPLY_TILE;
    const uint32_t n = min(__ldg(&fp.cursor[p]), fp.cap);
    const uint32_t i1 = min(i0 + GRB_APPLY_TILE, n);
    const uint32_t* list = fp.lists + (uint64_t)p * fp.cap;
    const uint64_t base = (uint64_t)p << fp.pshift;
    // 8 independent streaming loads, then 8 fire-and-forget atomics (RED) per thread and step
    for (uint32_t i = i0 + threadIdx.x; i < i1; i += 256 * 8) {
      uint32_t o[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        o[u] = (i + u * 256 < i1) ? __ldcs(&list[i + u * 256]) : 0xFFFFFFFFu;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (o[u] != 0xFFFFFFFFu) {
          grb_set_bit_pos(filt, base | o[u]);
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
    slots[rank[i]] = GrbSlot{ ids[i], counts[i] };
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
