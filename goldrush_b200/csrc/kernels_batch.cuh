// The batch engine of the pass-2 loop: speculative query of a whole batch of reads against the
// filter as it stands at the batch start, then an ordered commit that re-validates, read by read
// in file order, every frame whose ID slots were rewritten by an earlier read of the same batch.
//
// Replaces the same reference code as kernels_select.cuh (read_hashing.cpp:7-75,
// goldrush_path.cpp:529-890 calc_num_assigned_tiles, :892-1094 process_read, :156-187
// silver_path_check, MIBFConstructSupport.hpp:247-283 insertMIBF) and produces the same decisions:
// the reference's loop-carried dependence (every query sees every earlier insert,
// goldrush_path.cpp:1229-1256) is preserved exactly, not approximately.
//
//   k_batch_begin      new batch serial number (state.epoch)
//   k_spec_query       CTA per (read, tile): hash, probe, per-tile vote table -> compact (id,count)
//                      list, arg-max, per-probe rank stash                        [whole GPU]
//   per read, in order:
//     k_commit_check   CTA per tile: find the probes whose slot was rewritten by an earlier read of
//                      this batch (L2-resident hashed bitmap as prefilter, then the slot's epoch
//                      tag), and for each such frame move its votes from the old IDs (slot.id0) to
//                      the new ones
//     k_commit_decide  one CTA: count matrix of the arg-max IDs, smoothing, decision, bookkeeping
//     k_insert_collect / k_insert_apply (kernels_select.cuh)  reservoir insert; records id0 and
//                      marks the bitmap for every slot whose ID changed
//
// A slot whose epoch tag equals the current batch's serial number was rewritten by an earlier read
// of this batch and held slot.id0 when the batch started; any other slot still holds what the
// speculative query saw.  (The ID value itself cannot tell: with block size 1 a trimmed read hands
// out one ID beyond ids_inserted, goldrush_path.cpp:1040-1053, which the next read reuses.)
#pragma once
#include "common.cuh"
#include "decide.cuh"
#include "kernels_select.cuh"

#define GRB_STASH_NOFRAME (1ull << 63) // on pattern 0's rank: the frame failed the bit test

struct GrbBatchDev
{
  const uint64_t* read_idx;   // [nb] store index of batch read b
  const uint32_t* tile_first; // [nb + 1] first batch tile of read b
  const uint32_t* tile_read;  // [n_bt] b of each batch tile
  uint32_t nb, n_bt;
  uint64_t* stash;      // [n_bt * tile_len * h] rank of every probe
  uint32_t* vt_n;       // [n_bt] entries of the compact vote table
  uint32_t* vt_id;      // [n_bt * vt_cap]
  uint32_t* vt_cnt;     // [n_bt * vt_cap]
  uint32_t* best_id;    // [n_bt]
  uint32_t* best_count; // [n_bt]
  uint32_t* tile_hits;  // [n_bt]
  uint32_t* tile_miss;  // [n_bt]
  uint32_t vt_cap;
  uint32_t dirty_mask;  // bits of the hashed bitmap - 1
  uint32_t* dirty_bits;
  uint32_t* cmat;       // global spill of the count matrix for reads with very many tiles
};

__device__ __forceinline__ uint32_t
grb_norm_id(uint32_t v)
{
  return v > GRB_SAT_MASK ? (v & ~GRB_SAT_MASK) : v; // goldrush_path.cpp:574-583
}

__global__ void
k_batch_begin(GrbSelState* __restrict__ state)
{
  if (!state->halt) {
    state->epoch += 1u;
    state->batch_inserts = 0;
  }
}

// arg-max over a shared-memory vote table (ties -> smallest id) + compaction to global memory.
// Must be called by all threads of the CTA; s_n / s_best are shared scalars zeroed beforehand.
template<int BS>
__device__ __forceinline__ void
grb_compact_table(const uint32_t* keys, const uint32_t* counts, uint32_t table_size, uint32_t* s_n,
                  unsigned long long* s_best, uint32_t* __restrict__ out_id,
                  uint32_t* __restrict__ out_cnt, uint32_t cap)
{
  unsigned long long best = 0;
  for (unsigned i = threadIdx.x; i < table_size; i += BS) {
    const uint32_t c = counts[i];
    if (c) {
      const uint32_t id = keys[i];
      const unsigned long long key = ((unsigned long long)c << 32) | (0xFFFFFFFFu - id);
      best = key > best ? key : best;
      const uint32_t at = atomicAdd(s_n, 1u);
      if (at < cap) {
        out_id[at] = id;
        out_cnt[at] = c;
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
    best = o > best ? o : best;
  }
  if ((threadIdx.x & 31) == 0 && best) {
    atomicMax(s_best, best);
  }
}

__device__ __forceinline__ void
grb_vote_add(uint32_t* keys, uint32_t* counts, uint32_t tmask, uint32_t id, uint32_t delta)
{
  uint32_t slot = grb_mix32(id) & tmask;
  while (true) {
    const uint32_t old = atomicCAS(&keys[slot], 0u, id);
    if (old == 0u || old == id) {
      atomicAdd(&counts[slot], delta); // delta may be (uint32_t)-1: counts are exact mod 2^32
      return;
    }
    slot = (slot + 1) & tmask;
  }
}

// One CTA per batch tile (grid-strided).  Dynamic shared memory:
//   GrbSeedTables | uint64 sw[sw_words] | uint32 keys[table_size] | uint32 counts[table_size]
template<int BS>
__global__ void __launch_bounds__(BS)
k_spec_query(GrbReadsDev reads, const GrbSeedTables* __restrict__ seeds_g, GrbFilterDev filt,
             GrbSelParams prm, GrbBatchDev bd, const GrbSelState* __restrict__ state)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GrbSeedTables& st = *reinterpret_cast<GrbSeedTables*>(smem_raw);
  uint64_t* sw = reinterpret_cast<uint64_t*>(smem_raw + sizeof(GrbSeedTables));
  uint32_t* keys = reinterpret_cast<uint32_t*>(sw + prm.sw_words);
  uint32_t* counts = keys + prm.table_size;
  __shared__ uint32_t s_n, s_hits, s_miss;
  __shared__ unsigned long long s_best;

  if (state->halt) {
    return;
  }
  for (unsigned i = threadIdx.x; i < sizeof(GrbSeedTables) / 8; i += BS) {
    reinterpret_cast<uint64_t*>(&st)[i] = reinterpret_cast<const uint64_t*>(seeds_g)[i];
  }
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h;
  const uint32_t tmask = prm.table_size - 1;

  for (uint32_t bt = blockIdx.x; bt < bd.n_bt; bt += gridDim.x) {
    const uint32_t b = bd.tile_read[bt];
    const uint32_t t = bt - bd.tile_first[b];
    const uint64_t read_idx = bd.read_idx[b];
    const uint32_t len = reads.len[read_idx];
    const uint64_t w_read = reads.word_off[read_idx];
    const uint32_t w_total = (len + 31) / 32;
    const uint32_t tl = grb_tile_bases(len, t, T, k);
    const uint32_t frames = tl - k + 1;
    const uint32_t p0 = t * T;
    const uint32_t w_first = p0 >> 5;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < prm.sw_words; i += BS) {
      sw[i] = (w_first + i < w_total) ? reads.bases[w_read + w_first + i] : 0ull;
    }
    for (unsigned i = threadIdx.x; i < prm.table_size; i += BS) {
      keys[i] = 0;
      counts[i] = 0;
    }
    if (threadIdx.x == 0) {
      s_n = 0;
      s_best = 0;
      s_hits = 0;
      s_miss = 0;
    }
    __syncthreads();
    uint32_t my_hits = 0, my_miss = 0;
    uint64_t* stash = bd.stash + (uint64_t)bt * T * h;
    for (uint32_t f = threadIdx.x; f < frames; f += BS) {
      uint64_t rank[GRB_MAX_PATTERNS];
      bool all = true;
#pragma unroll
      for (unsigned i = 0; i < GRB_MAX_PATTERNS; ++i) {
        if (i < h) {
          const uint32_t n_i = tl - (k + i) + 1; // valid positions of pattern i in this tile
          const uint32_t p = f < n_i ? f : n_i - 1; // stale tail keeps the last value
          const GrbWindow w =
            grb_window([&](uint64_t wi) { return sw[wi]; }, (uint64_t)(p0 & 31) + p);
          const uint64_t hv = grb_hash_direct(st, i, w);
          bool bit;
          grb_probe_block(filt, grb_fastmod(hv, filt.bits, filt.inv), bit, rank[i]);
          all &= bit;
        }
      }
#pragma unroll
      for (unsigned i = 0; i < GRB_MAX_PATTERNS; ++i) {
        if (i < h) {
          stash[(uint64_t)f * h + i] = (i == 0 && !all) ? (rank[i] | GRB_STASH_NOFRAME) : rank[i];
        }
      }
      if (!all) { // MIBloomFilter::atRank fails on the first clear bit: the frame counts nothing
        continue;
      }
      uint32_t ids[GRB_MAX_PATTERNS];
#pragma unroll
      for (unsigned i = 0; i < GRB_MAX_PATTERNS; ++i) {
        if (i < h) {
          ids[i] = grb_norm_id(__ldcg(&filt.slots[rank[i]].id));
        }
      }
#pragma unroll
      for (unsigned i = 0; i < GRB_MAX_PATTERNS; ++i) {
        if (i < h) {
          const uint32_t v = ids[i];
          if (v == 0) {
            ++my_miss;
            continue;
          }
          ++my_hits;
          bool dup = false; // an id counts once per frame (std::set, goldrush_path.cpp:570)
#pragma unroll
          for (unsigned j = 0; j < GRB_MAX_PATTERNS; ++j) {
            if (j < i && ids[j] == v) {
              dup = true;
            }
          }
          if (!dup) {
            grb_vote_add(keys, counts, tmask, v, 1u);
          }
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      my_hits += __shfl_xor_sync(0xffffffffu, my_hits, d);
      my_miss += __shfl_xor_sync(0xffffffffu, my_miss, d);
    }
    if ((threadIdx.x & 31) == 0) {
      if (my_hits) {
        atomicAdd(&s_hits, my_hits);
      }
      if (my_miss) {
        atomicAdd(&s_miss, my_miss);
      }
    }
    __syncthreads();
    grb_compact_table<BS>(keys, counts, prm.table_size, &s_n, &s_best,
                          bd.vt_id + (uint64_t)bt * bd.vt_cap, bd.vt_cnt + (uint64_t)bt * bd.vt_cap,
                          bd.vt_cap);
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long bb = s_best;
      bd.best_count[bt] = (uint32_t)(bb >> 32);
      bd.best_id[bt] = bb ? 0xFFFFFFFFu - (uint32_t)(bb & 0xFFFFFFFFu) : 0u;
      bd.vt_n[bt] = s_n;
      bd.tile_hits[bt] = s_hits;
      bd.tile_miss[bt] = s_miss;
    }
  }
}

// Ordered commit, step 1: one CTA per tile of batch read b.  Dynamic shared memory:
//   uint32 fbits[fb_words] | uint32 keys[table_size] | uint32 counts[table_size]
template<int BS>
__global__ void __launch_bounds__(BS)
k_commit_check(GrbReadsDev reads, GrbFilterDev filt, GrbSelParams prm, GrbBatchDev bd,
               const GrbSelState* __restrict__ state, uint32_t b, uint32_t fb_words)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* fbits = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* keys = fbits + fb_words;
  uint32_t* counts = keys + prm.table_size;
  __shared__ uint32_t s_dirty, s_n;
  __shared__ int s_dhits; // change of the tile's hit count (misses change by the opposite)
  __shared__ unsigned long long s_best;

  if (state->halt) {
    return;
  }
  if (state->batch_inserts == 0) {
    return; // nothing inserted in this batch so far: the speculative votes stand
  }
  const uint32_t epoch = state->epoch;
  const uint64_t read_idx = bd.read_idx[b];
  const uint32_t len = reads.len[read_idx];
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h;
  const uint32_t n_tiles = len / T;
  const uint32_t tmask = prm.table_size - 1;

  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const uint32_t bt = bd.tile_first[b] + t;
    const uint32_t tl = grb_tile_bases(len, t, T, k);
    const uint32_t frames = tl - k + 1;
    const uint64_t* stash = bd.stash + (uint64_t)bt * T * h;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < fb_words; i += BS) {
      fbits[i] = 0;
    }
    if (threadIdx.x == 0) {
      s_dirty = 0;
      s_n = 0;
      s_dhits = 0;
      s_best = 0;
    }
    __syncthreads();
    // ---- which frames probe a slot rewritten by an earlier read of this batch ----
    const uint32_t n_probe = frames * h;
    for (uint32_t idx = threadIdx.x; idx < n_probe; idx += BS) {
      const uint64_t raw = stash[idx];
      const uint32_t f = idx / h;
      if (idx - f * h == 0 && (raw & GRB_STASH_NOFRAME)) {
        continue;
      }
      const uint64_t r = raw & ~GRB_STASH_NOFRAME;
      const uint32_t hb = (uint32_t)r & bd.dirty_mask;
      if ((__ldcg(&bd.dirty_bits[hb >> 5]) >> (hb & 31)) & 1u) {
        if (__ldcg(&filt.slots[r].epoch) == epoch) {
          atomicOr(&fbits[f >> 5], 1u << (f & 31));
          s_dirty = 1;
        }
      }
    }
    __syncthreads();
    if (!s_dirty) {
      continue;
    }
    // ---- load the tile's vote table, move the votes of the dirty frames, write it back ----
    for (unsigned i = threadIdx.x; i < prm.table_size; i += BS) {
      keys[i] = 0;
      counts[i] = 0;
    }
    __syncthreads();
    uint32_t* vid = bd.vt_id + (uint64_t)bt * bd.vt_cap;
    uint32_t* vcnt = bd.vt_cnt + (uint64_t)bt * bd.vt_cap;
    const uint32_t n_old = bd.vt_n[bt];
    for (uint32_t i = threadIdx.x; i < n_old; i += BS) {
      grb_vote_add(keys, counts, tmask, vid[i], vcnt[i]);
    }
    int dh = 0;
    for (uint32_t f = threadIdx.x; f < frames; f += BS) {
      if (!((fbits[f >> 5] >> (f & 31)) & 1u)) {
        continue;
      }
      if (stash[(uint64_t)f * h] & GRB_STASH_NOFRAME) {
        continue; // a frame that failed the bit test never voted (another pattern set the flag)
      }
      uint32_t oldv[GRB_MAX_PATTERNS], newv[GRB_MAX_PATTERNS];
#pragma unroll
      for (unsigned i = 0; i < GRB_MAX_PATTERNS; ++i) {
        if (i < h) {
          const uint64_t r = stash[(uint64_t)f * h + i] & ~GRB_STASH_NOFRAME;
          const uint4 s = __ldcg(reinterpret_cast<const uint4*>(&filt.slots[r]));
          const uint32_t nv = grb_norm_id(s.x);
          newv[i] = nv;
          oldv[i] = s.w == epoch ? grb_norm_id(s.z) : nv;
          dh += (nv != 0) - (oldv[i] != 0);
        }
      }
#pragma unroll
      for (unsigned i = 0; i < GRB_MAX_PATTERNS; ++i) {
        if (i < h) {
          // old id leaves the frame's set unless it is still one of the new ids
          const uint32_t o = oldv[i];
          bool first = o != 0, stays = false;
          const uint32_t n = newv[i];
          bool nfirst = n != 0, was = false;
#pragma unroll
          for (unsigned j = 0; j < GRB_MAX_PATTERNS; ++j) {
            if (j < h) {
              if (j < i && oldv[j] == o) {
                first = false;
              }
              if (newv[j] == o) {
                stays = true;
              }
              if (j < i && newv[j] == n) {
                nfirst = false;
              }
              if (oldv[j] == n) {
                was = true;
              }
            }
          }
          if (first && !stays) {
            grb_vote_add(keys, counts, tmask, o, 0xFFFFFFFFu);
          }
          if (nfirst && !was) {
            grb_vote_add(keys, counts, tmask, n, 1u);
          }
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      dh += __shfl_xor_sync(0xffffffffu, dh, d);
    }
    if ((threadIdx.x & 31) == 0 && dh) {
      atomicAdd(&s_dhits, dh);
    }
    __syncthreads();
    grb_compact_table<BS>(keys, counts, prm.table_size, &s_n, &s_best, vid, vcnt, bd.vt_cap);
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long bb = s_best;
      bd.best_count[bt] = (uint32_t)(bb >> 32);
      bd.best_id[bt] = bb ? 0xFFFFFFFFu - (uint32_t)(bb & 0xFFFFFFFFu) : 0u;
      bd.vt_n[bt] = s_n;
      bd.tile_hits[bt] = (uint32_t)((int)bd.tile_hits[bt] + s_dhits);
      bd.tile_miss[bt] = (uint32_t)((int)bd.tile_miss[bt] - s_dhits);
    }
  }
}

// Votes of one read as a dense matrix over the distinct arg-max ids of its tiles: every id the
// smoothing passes ask about is the arg-max of some tile (goldrush_path.cpp:646-682 only ever
// propagates neighbours' ids), so count[i][u] for those ids is all that is needed.
struct GrbMatrixVotes
{
  const uint32_t* best_id_;
  const uint32_t* best_count_;
  const uint32_t* cmat;  // [n * nu] count of uniq[u] in tile i if > 2, else 0
  const uint32_t* ukeys; // open-addressing map id -> u
  const uint32_t* uvals; // 0xFFFFFFFF = empty
  uint32_t nu, umask;
  __device__ __forceinline__ uint32_t best_id(uint32_t i) const { return best_id_[i]; }
  __device__ __forceinline__ uint32_t best_count(uint32_t i) const { return best_count_[i]; }
  __device__ __forceinline__ uint32_t lookup(uint32_t id) const
  {
    uint32_t s = grb_mix32(id) & umask;
    while (true) {
      const uint32_t u = uvals[s];
      if (u == 0xFFFFFFFFu) {
        return u;
      }
      if (ukeys[s] == id) {
        return u;
      }
      s = (s + 1) & umask;
    }
  }
  __device__ __forceinline__ uint32_t cand_count(uint32_t i, uint32_t id) const
  {
    const uint32_t u = lookup(id);
    return u == 0xFFFFFFFFu ? 0u : cmat[(uint64_t)i * nu + u];
  }
};

// Ordered commit, step 2: one CTA.  Dynamic shared memory (n = tiles of the read, n <= n_cap):
//   uint32 best_id[n] best_cnt[n] root[n] uidx[n] tile_id[n] snap[n+2] | uint32 ukeys[us] uvals[us]
//   | uint8 tile_as[n] (padded) | uint32 cmat[n * nu]   (cmat spills to bd.cmat when cm_smem == 0)
template<int BS>
__global__ void __launch_bounds__(BS)
k_commit_decide(GrbReadsDev reads, GrbSelParams prm, GrbBatchDev bd, GrbReadPlan* __restrict__ plan_out,
                GrbSelState* __restrict__ state, grb_decision* __restrict__ decisions, uint32_t b,
                uint64_t dec_idx, uint32_t n_cap, uint32_t us, uint32_t cm_smem)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_best_id = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* s_best_cnt = s_best_id + n_cap;
  uint32_t* s_root = s_best_cnt + n_cap;
  uint32_t* s_uidx = s_root + n_cap;
  uint32_t* s_tile_id = s_uidx + n_cap;
  uint32_t* s_snap = s_tile_id + n_cap;
  uint32_t* s_ukeys = s_snap + n_cap + 2;
  uint32_t* s_uvals = s_ukeys + us;
  uint8_t* s_tile_as = reinterpret_cast<uint8_t*>(s_uvals + us);
  uint32_t* s_cmat = reinterpret_cast<uint32_t*>(s_tile_as + ((n_cap + 15) / 16) * 16);
  __shared__ uint32_t s_nu;
  __shared__ unsigned long long s_hits, s_miss, s_queries;

  if (state->halt) {
    return;
  }
  const uint64_t read_idx = bd.read_idx[b];
  const uint32_t len = reads.len[read_idx];
  const uint32_t T = prm.tile_len, k = prm.k;
  const uint32_t n = len / T;
  const uint32_t bt0 = bd.tile_first[b];
  const uint32_t umask = us - 1;

  if (threadIdx.x == 0) {
    s_hits = 0;
    s_miss = 0;
    s_queries = 0;
  }
  for (unsigned i = threadIdx.x; i < us; i += BS) {
    s_uvals[i] = 0xFFFFFFFFu;
  }
  __syncthreads();
  unsigned long long my_h = 0, my_m = 0, my_q = 0;
  for (uint32_t i = threadIdx.x; i < n; i += BS) {
    s_best_id[i] = bd.best_id[bt0 + i];
    s_best_cnt[i] = bd.best_count[bt0 + i];
    my_h += bd.tile_hits[bt0 + i];
    my_m += bd.tile_miss[bt0 + i];
    my_q += grb_tile_bases(len, i, T, k) - k + 1;
  }
  if (my_q) {
    atomicAdd(&s_hits, my_h);
    atomicAdd(&s_miss, my_m);
    atomicAdd(&s_queries, my_q);
  }
  __syncthreads();
  // distinct arg-max ids: root[i] = first tile with the same id
  for (uint32_t i = threadIdx.x; i < n; i += BS) {
    const uint32_t v = s_best_id[i];
    uint32_t j = 0;
    while (s_best_id[j] != v) {
      ++j;
    }
    s_root[i] = j;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t nu = 0;
    for (uint32_t i = 0; i < n; ++i) {
      if (s_root[i] == i) {
        const uint32_t v = s_best_id[i];
        uint32_t s = grb_mix32(v) & umask;
        while (s_uvals[s] != 0xFFFFFFFFu) {
          s = (s + 1) & umask;
        }
        s_ukeys[s] = v;
        s_uvals[s] = nu;
        s_uidx[i] = nu++;
      }
    }
    s_nu = nu;
  }
  __syncthreads();
  const uint32_t nu = s_nu;
  uint32_t* cmat = cm_smem ? s_cmat : bd.cmat;
  for (uint32_t i = threadIdx.x; i < n * nu; i += BS) {
    cmat[i] = 0;
  }
  __syncthreads();
  GrbMatrixVotes v{ s_best_id, s_best_cnt, cmat, s_ukeys, s_uvals, nu, umask };
  {
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = warp; i < n; i += BS / 32) {
      const uint32_t ne = bd.vt_n[bt0 + i];
      const uint32_t* vid = bd.vt_id + (uint64_t)(bt0 + i) * bd.vt_cap;
      const uint32_t* vcnt = bd.vt_cnt + (uint64_t)(bt0 + i) * bd.vt_cap;
      for (uint32_t e = lane; e < ne; e += 32) {
        const uint32_t c = vcnt[e];
        if (c > 2) { // the reference's candidate list holds ids with count > 2 (:616)
          const uint32_t u = v.lookup(vid[e]);
          if (u != 0xFFFFFFFFu) {
            cmat[(uint64_t)i * nu + u] = c;
          }
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x != 0) {
    return;
  }
  const uint32_t n_as = grb_smooth_tiles(n, v, prm.threshold, s_tile_id, s_tile_as, s_snap);
  GrbSelState s = *state;
  s.cur.queries += s_queries;
  s.cur.hits += s_hits;
  s.cur.misses += s_miss;
  s.cur.total_tiles += n;
  s.cur.assigned_tiles += n_as;
  s.cur.unassigned_tiles += n - n_as;
  GrbReadPlan plan;
  grb_plan_read(n, n_as, len, prm.tile_len, prm.block_size, prm.unassigned_min, prm.assigned_max,
                s_tile_id, s_tile_as, &s.ids_inserted, &plan);
  grb_decision d;
  d.verdict = plan.verdict;
  d.pad[0] = d.pad[1] = d.pad[2] = 0;
  d.path = (uint32_t)s.curr_path;
  d.trim_start = plan.trim_start;
  d.trim_end = plan.trim_end;
  d.num_tiles = n;
  d.num_assigned = n_as;
  decisions[dec_idx] = d;
  if (plan.verdict == GRB_UNTRIMMED || plan.verdict == GRB_TRIMMED) {
    s.cur.inserted_bases += plan.out_bases;
    s.cur.num_reads_in_path += 1;
    s.batch_inserts += 1;
    if (prm.silver && prm.target_bases < s.cur.inserted_bases) { // silver_path_check, :167-186
      s.snap = s.cur;
      s.snap.rollover_read = read_idx;
      s.n_snap = 1;
      s.curr_path += 1;
      s.halt = 1;
      s.halt_read = read_idx;
      if (prm.max_paths < s.curr_path) {
        s.finished = 1;
      } else {
        s.cur.inserted_bases = 0;
        s.cur.num_reads_in_path = 0;
        s.cur.phred_sum_in_path = 0;
        s.ids_inserted = 0;
      }
      plan.n_blocks = 0; // every ID and count is wiped right after this insert: skip it
    }
  }
  if (!s.finished) {
    s.cur.valid_reads += 1;
  }
  *state = s;
  *plan_out = plan;
}
