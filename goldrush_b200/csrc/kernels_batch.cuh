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
//   k_spec_cmat        CTA per read: distinct arg-max ids of its tiles + their count matrix
//                      and the smoothing passes on the speculative votes
//   k_spec_dedupe      per read: distinct ranks of its probes -> mask of the tiles probing them
//   k_commit_batch     ONE persistent cooperative launch per batch; for each read, in file order:
//     check    CTA per tile: find the probes whose slot was rewritten by an earlier read of this
//              batch (L2-resident hashed bitmap as prefilter, then the slot's epoch tag), and for
//              each such frame move its votes from the old IDs (slot.id0) to the new ones; refresh
//              the tile's arg-max and its row of the count matrix            -- grid barrier --
//     decide   every CTA, redundantly: decision + bookkeeping from the smoothing result (CTA 0
//              re-runs the smoothing first when one of its inputs changed)
//     insert   all CTAs: reservoir insert over the read's de-duplicated ranks; records id0 and
//              marks the bitmap for every slot whose ID changed               -- grid barrier --
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
  uint32_t* cmat;       // global spill of the slow path's count matrix (very long reads)
  // per read: distinct arg-max ids of its tiles (uq[tile_first[b] + u], u < nu[b]) and the
  // count matrix cm[cm_off[b] + i * nu[b] + u] = votes of uq[u] in tile i if > 2, else 0
  uint32_t* uq;           // [n_bt]
  uint32_t* nu;           // [nb]
  uint32_t* u_changed;    // [nb] a re-validated tile's arg-max left uq: decide recomputes both
  uint32_t* cm;
  const uint64_t* cm_off; // [nb]
  // smoothing result on the speculative votes (valid for the commit unless in_changed[b])
  uint32_t* sp_tile_id;   // [n_bt]
  uint8_t* sp_tile_as;    // [n_bt]
  uint32_t* sp_n_as;      // [nb]
  uint32_t* in_changed;   // [nb] re-validation changed an input of the smoothing passes
  // decision of read b up to the ID counter: first_id is relative to ids_inserted (1 when the read
  // inserts), sp_adv[b] is what the read adds to ids_inserted (goldrush_path.cpp:982-994,1040-1053)
  GrbReadPlan* sp_plan;   // [nb]
  uint32_t* sp_adv;       // [nb]
  uint32_t* rd_hits;      // [nb] per-read totals of tile_hits / tile_miss / frames
  uint32_t* rd_miss;      // [nb]
  uint32_t* rd_queries;   // [nb]
  // per-read de-duplicated insert table, built speculatively: rank -> mask of the read's tiles
  // that probe it (reads of at most 64 tiles; longer reads use k_insert_collect at commit time)
  uint64_t* dd_key;
  uint64_t* dd_mask;
  const uint64_t* dd_off;  // [nb] first entry of read b's table
  const uint32_t* dd_size; // [nb] entries (power of two), 0 = no table
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

// id -> u map over the distinct arg-max ids of one read, in shared memory
struct GrbUMap
{
  const uint32_t* keys;
  const uint32_t* vals; // 0xFFFFFFFF = empty
  uint32_t mask;
  GRB_HD uint32_t lookup(uint32_t id) const
  {
    uint32_t s = grb_mix32(id) & mask;
    while (true) {
      const uint32_t u = vals[s];
      if (u == 0xFFFFFFFFu || keys[s] == id) {
        return u;
      }
      s = (s + 1) & mask;
    }
  }
};

__device__ __forceinline__ void
grb_umap_insert(uint32_t* keys, uint32_t* vals, uint32_t mask, uint32_t id, uint32_t u)
{
  uint32_t s = grb_mix32(id) & mask;
  while (vals[s] != 0xFFFFFFFFu) {
    s = (s + 1) & mask;
  }
  keys[s] = id;
  vals[s] = u;
}

// concurrent version (distinct ids, one per thread); lookups only after a barrier
__device__ __forceinline__ void
grb_umap_insert_par(uint32_t* keys, uint32_t* vals, uint32_t mask, uint32_t id, uint32_t u)
{
  uint32_t s = grb_mix32(id) & mask;
  while (atomicCAS(&vals[s], 0xFFFFFFFFu, u) != 0xFFFFFFFFu) {
    s = (s + 1) & mask;
  }
  keys[s] = id;
}

// Votes of one read as a dense matrix over the distinct arg-max ids of its tiles: every id the
// smoothing passes ask about is the arg-max of some tile (goldrush_path.cpp:646-682 only ever
// propagates neighbours' ids), so count[i][u] for those ids is all that is needed.
struct GrbMatrixVotes
{
  const uint32_t* best_id_;
  const uint32_t* best_count_;
  const uint32_t* cmat; // [n * nu] count of uniq[u] in tile i if > 2, else 0
  GrbUMap umap;
  uint32_t nu;
  GRB_HD uint32_t best_id(uint32_t i) const { return best_id_[i]; }
  GRB_HD uint32_t best_count(uint32_t i) const { return best_count_[i]; }
  GRB_HD uint32_t cand_count(uint32_t i, uint32_t id) const
  {
    const uint32_t u = umap.lookup(id);
    return u == 0xFFFFFFFFu ? 0u : cmat[(uint64_t)i * nu + u];
  }
};

// Distinct arg-max ids of the n tiles whose arg-max ids are best[0..n) and their count matrix,
// by the whole CTA.  root[n], ukeys/uvals[us] are shared scratch; uq_out[n] receives the ids.
// Returns nu (to all threads).  cmat[n * nu] is filled from the tiles' compact vote tables.
template<int BS>
__device__ __forceinline__ uint32_t
grb_build_cmat(const GrbBatchDev& bd, uint32_t bt0, uint32_t n, const uint32_t* best, uint32_t* root,
               uint32_t* ukeys, uint32_t* uvals, uint32_t us, uint32_t* uq_out, uint32_t* cmat)
{
  __shared__ uint32_t s_nu;
  for (unsigned i = threadIdx.x; i < us; i += BS) {
    uvals[i] = 0xFFFFFFFFu;
  }
  for (uint32_t i = threadIdx.x; i < n; i += BS) { // root[i] = first tile with the same id
    const uint32_t v = best[i];
    uint32_t j = 0;
    while (best[j] != v) {
      ++j;
    }
    root[i] = j;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t nu = 0;
    for (uint32_t i = 0; i < n; ++i) {
      if (root[i] == i) {
        grb_umap_insert(ukeys, uvals, us - 1, best[i], nu);
        uq_out[nu++] = best[i];
      }
    }
    s_nu = nu;
  }
  __syncthreads();
  const uint32_t nu = s_nu;
  for (uint32_t i = threadIdx.x; i < n * nu; i += BS) {
    cmat[i] = 0;
  }
  __syncthreads();
  const GrbUMap um{ ukeys, uvals, us - 1 };
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = warp; i < n; i += BS / 32) {
    const uint32_t ne = bd.vt_n[bt0 + i];
    const uint32_t* vid = bd.vt_id + (uint64_t)(bt0 + i) * bd.vt_cap;
    const uint32_t* vcnt = bd.vt_cnt + (uint64_t)(bt0 + i) * bd.vt_cap;
    for (uint32_t e = lane; e < ne; e += 32) {
      const uint32_t c = vcnt[e];
      if (c > 2) { // the reference's candidate list holds ids with count > 2 (:616)
        const uint32_t u = um.lookup(vid[e]);
        if (u != 0xFFFFFFFFu) {
          cmat[(uint64_t)i * nu + u] = c;
        }
      }
    }
  }
  __syncthreads();
  return nu;
}

// After the speculative query: one CTA per read of the batch builds the read's count matrix and
// runs the smoothing passes on the speculative votes.  Dynamic shared memory:
//   uint32 best[n_cap] bcnt[n_cap] root[n_cap] tile_id[n_cap] snap[n_cap+2] ukeys[us] uvals[us]
//   | uint8 tile_as[n_cap]
template<int BS>
__global__ void __launch_bounds__(BS)
k_spec_cmat(GrbReadsDev reads, GrbSelParams prm, GrbBatchDev bd, const GrbSelState* __restrict__ state,
            uint32_t n_cap, uint32_t us)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_best = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* s_bcnt = s_best + n_cap;
  uint32_t* s_root = s_bcnt + n_cap;
  uint32_t* s_tile_id = s_root + n_cap;
  uint32_t* s_snap = s_tile_id + n_cap;
  uint32_t* s_ukeys = s_snap + n_cap + 2;
  uint32_t* s_uvals = s_ukeys + us;
  uint8_t* s_tile_as = reinterpret_cast<uint8_t*>(s_uvals + us);
  if (state->halt) {
    return;
  }
  for (uint32_t b = blockIdx.x; b < bd.nb; b += gridDim.x) {
    const uint32_t bt0 = bd.tile_first[b];
    const uint32_t n = bd.tile_first[b + 1] - bt0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += BS) {
      s_best[i] = bd.best_id[bt0 + i];
      s_bcnt[i] = bd.best_count[bt0 + i];
    }
    __syncthreads();
    const uint32_t* cm = bd.cm + bd.cm_off[b];
    const uint32_t nu = grb_build_cmat<BS>(bd, bt0, n, s_best, s_root, s_ukeys, s_uvals, us,
                                           bd.uq + bt0, bd.cm + bd.cm_off[b]);
    if (threadIdx.x == 0) {
      bd.nu[b] = nu;
      bd.u_changed[b] = 0;
      bd.in_changed[b] = 0;
      __threadfence_block();
      const GrbMatrixVotes v{ s_best, s_bcnt, cm, GrbUMap{ s_ukeys, s_uvals, us - 1 }, nu };
      const uint32_t n_as = grb_smooth_tiles(n, v, prm.threshold, s_tile_id, s_tile_as, s_snap);
      bd.sp_n_as[b] = n_as;
      uint32_t rel = 0;
      GrbReadPlan plan;
      grb_plan_read(n, n_as, reads.len[bd.read_idx[b]], prm.tile_len, prm.block_size,
                    prm.unassigned_min, prm.assigned_max, s_tile_id, s_tile_as, &rel, &plan);
      bd.sp_plan[b] = plan;
      bd.sp_adv[b] = rel;
      bd.rd_hits[b] = 0;
      bd.rd_miss[b] = 0;
      bd.rd_queries[b] = 0;
    }
    __syncthreads();
    {
      const uint32_t len = reads.len[bd.read_idx[b]];
      uint32_t my_h = 0, my_m = 0, my_q = 0;
      for (uint32_t i = threadIdx.x; i < n; i += BS) {
        bd.sp_tile_id[bt0 + i] = s_tile_id[i];
        bd.sp_tile_as[bt0 + i] = s_tile_as[i];
        my_h += bd.tile_hits[bt0 + i];
        my_m += bd.tile_miss[bt0 + i];
        my_q += grb_tile_bases(len, i, prm.tile_len, prm.k) - prm.k + 1;
      }
      if (my_q) {
        atomicAdd(&bd.rd_hits[b], my_h);
        atomicAdd(&bd.rd_miss[b], my_m);
        atomicAdd(&bd.rd_queries[b], my_q);
      }
    }
  }
}

// Speculative de-duplication of every read's insert: rank -> mask of the read's tiles probing it
// (the dense_hash_set of MIBFConstructSupport.hpp:255-270, per read instead of per b-tile block;
// k_insert_apply_pre folds the tile mask into insert calls once the decision is known).
__global__ void __launch_bounds__(256)
k_spec_dedupe(GrbReadsDev reads, GrbSelParams prm, GrbBatchDev bd, const GrbSelState* __restrict__ state)
{
  if (state->halt) {
    return;
  }
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h;
  const uint32_t per_tile = T * h;
  for (uint32_t bt = blockIdx.x; bt < bd.n_bt; bt += gridDim.x) {
    const uint32_t b = bd.tile_read[bt];
    const uint32_t size = bd.dd_size[b];
    if (size == 0) {
      continue;
    }
    const uint32_t t = bt - bd.tile_first[b];
    const uint32_t len = reads.len[bd.read_idx[b]];
    const uint32_t tl = grb_tile_bases(len, t, T, k);
    const uint64_t* stash = bd.stash + (uint64_t)bt * per_tile;
    uint64_t* keys = bd.dd_key + bd.dd_off[b];
    uint64_t* masks = bd.dd_mask + bd.dd_off[b];
    const uint64_t m = size - 1;
    for (uint32_t idx = threadIdx.x; idx < per_tile; idx += blockDim.x) {
      const uint32_t f = idx / h, p = idx - f * h;
      if (tl < k + p || f >= tl - (k + p) + 1) {
        continue; // stale-tail repeat of the last valid position: same rank, already registered
      }
      const uint64_t key = stash[idx] & ~GRB_STASH_NOFRAME;
      uint64_t slot = grb_mix64(key) & m;
      while (true) {
        const unsigned long long old =
          atomicCAS(reinterpret_cast<unsigned long long*>(&keys[slot]), GRB_EMPTY_KEY, key);
        if (old == GRB_EMPTY_KEY || old == key) {
          atomicOr(reinterpret_cast<unsigned long long*>(&masks[slot]), 1ull << t);
          break;
        }
        slot = (slot + 1) & m;
      }
    }
  }
}

#define GRB_DELTA_FOUND 0x80000000u

// every load of data another CTA wrote earlier in the same launch goes to L2
template<typename T>
__device__ __forceinline__ T
grb_ld(const T* p)
{
  return __ldcg(p);
}

// Grid-wide barrier of the persistent commit kernel (cooperative launch: all CTAs are resident).
__device__ __forceinline__ void
grb_grid_barrier(unsigned long long* ctr, unsigned long long target)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1ull);
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// Shared-memory carve-up of k_commit_batch:
//   uint32 fbits[fb_words] dkeys[table_size] dvals[table_size] ukeys[us] uvals[us] oldrow[us]
//   | uint32 best_id[n_cap] best_cnt[n_cap] root[n_cap] uq[n_cap] tile_id[n_cap] snap[n_cap + 2]
//   | uint8 tile_as[n_cap] (padded to 16) | uint32 cmat[n_cap * n_cap] when cm_smem
struct GrbCommitSmem
{
  // check
  uint32_t* fbits;  // [fb_words]
  uint32_t* dkeys;  // [table_size]
  uint32_t* dvals;  // [table_size]
  uint32_t* ukeys;  // [us]
  uint32_t* uvals;  // [us]
  uint32_t* oldrow; // [us]
  // decide (n_cap tiles)
  uint32_t* best_id;
  uint32_t* best_cnt;
  uint32_t* root;
  uint32_t* uq;
  uint32_t* tile_id;
  uint32_t* snap;
  uint8_t* tile_as;
  uint32_t* cmat; // [n_cap * n_cap] when cm_smem
};

// check + re-validation of the tiles of batch read b owned by this CTA (tiles cta, cta + n_cta, ..)
template<int BS>
__device__ __forceinline__ void
grb_commit_check(const GrbReadsDev& reads, const GrbFilterDev& filt, const GrbSelParams& prm,
                 const GrbBatchDev& bd, const GrbCommitSmem& sm, uint32_t b, uint32_t epoch,
                 uint32_t cta, uint32_t n_cta, uint32_t fb_words, uint32_t us)
{
  __shared__ uint32_t s_dirty, s_new, s_chg;
  __shared__ int s_dhits; // change of the tile's hit count (misses change by the opposite)
  __shared__ unsigned long long s_best;
  uint32_t* fbits = sm.fbits;
  uint32_t* dkeys = sm.dkeys;
  uint32_t* dvals = sm.dvals;
  uint32_t* ukeys = sm.ukeys;
  uint32_t* uvals = sm.uvals;
  uint32_t* oldrow = sm.oldrow;

  const uint64_t read_idx = bd.read_idx[b];
  const uint32_t len = reads.len[read_idx];
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h;
  const uint32_t n_tiles = len / T;
  const uint32_t bt0 = bd.tile_first[b];
  const uint32_t nu = bd.nu[b];
  const uint32_t* uq = bd.uq + bt0;
  uint32_t* cm = bd.cm + bd.cm_off[b];

  for (uint32_t t = cta; t < n_tiles; t += n_cta) {
    const uint32_t bt = bt0 + t;
    const uint32_t tl = grb_tile_bases(len, t, T, k);
    const uint32_t frames = tl - k + 1;
    const uint64_t* stash = bd.stash + (uint64_t)bt * T * h;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < fb_words; i += BS) {
      fbits[i] = 0;
    }
    if (threadIdx.x == 0) {
      s_dirty = 0;
      s_new = 0;
      s_chg = 0;
      s_dhits = 0;
      s_best = 0;
    }
    __syncthreads();
    // ---- which frames probe a slot rewritten by an earlier read of this batch ----
    const uint32_t n_probe = frames * h;
    for (uint32_t base = 0; base < n_probe; base += 4 * BS) {
      uint64_t raw[4];
      uint32_t word[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t idx = base + u * BS + threadIdx.x;
        raw[u] = idx < n_probe ? __ldcs(&stash[idx]) : ~0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t hb = (uint32_t)raw[u] & bd.dirty_mask;
        word[u] = raw[u] != ~0ull ? __ldcg(&bd.dirty_bits[hb >> 5]) >> (hb & 31) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (word[u] & 1u) {
          const uint32_t idx = base + u * BS + threadIdx.x;
          const uint32_t f = idx / h;
          const uint64_t r = raw[u] & ~GRB_STASH_NOFRAME;
          if (__ldcg(&filt.slots[r].epoch) == epoch) {
            if (!((atomicOr(&fbits[f >> 5], 1u << (f & 31)) >> (f & 31)) & 1u)) {
              atomicAdd(&s_dirty, 1u); // number of dirty frames
            }
          }
        }
      }
    }
    __syncthreads();
    if (!s_dirty) {
      continue;
    }
    // ---- net vote change per id over the dirty frames (table sized to their number) ----
    uint32_t dsize = 256;
    while (dsize < 4u * h * s_dirty && dsize < prm.table_size) {
      dsize <<= 1;
    }
    const uint32_t tmask = dsize - 1;
    for (unsigned i = threadIdx.x; i < dsize; i += BS) {
      dkeys[i] = 0;
      dvals[i] = 0;
    }
    for (unsigned i = threadIdx.x; i < us; i += BS) {
      uvals[i] = 0xFFFFFFFFu;
    }
    __syncthreads();
    for (uint32_t u = threadIdx.x; u < nu; u += BS) {
      grb_umap_insert_par(ukeys, uvals, us - 1, uq[u], u);
      oldrow[u] = cm[(uint64_t)t * nu + u];
      cm[(uint64_t)t * nu + u] = 0;
    }
    int dh = 0;
    for (uint32_t f = threadIdx.x; f < frames; f += BS) {
      if (!((fbits[f >> 5] >> (f & 31)) & 1u)) {
        continue;
      }
      if (stash[(uint64_t)f * h] & GRB_STASH_NOFRAME) {
        continue; // a frame that failed the bit test never voted
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
          // an id leaves the frame's set unless it is still one of the new ids, and joins it
          // unless it was one of the old ids (the set is what counts, goldrush_path.cpp:570-604)
          const uint32_t o = oldv[i], n = newv[i];
          bool first = o != 0, stays = false, nfirst = n != 0, was = false;
#pragma unroll
          for (unsigned j = 0; j < GRB_MAX_PATTERNS; ++j) {
            if (j < h) {
              first = first && !(j < i && oldv[j] == o);
              stays = stays || newv[j] == o;
              nfirst = nfirst && !(j < i && newv[j] == n);
              was = was || oldv[j] == n;
            }
          }
          if (first && !stays) {
            grb_vote_add(dkeys, dvals, tmask, o, 0xFFFFFFFFu);
          }
          if (nfirst && !was) {
            grb_vote_add(dkeys, dvals, tmask, n, 1u);
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
    // ---- one pass over the tile's vote table: apply the changes, new arg-max, matrix row ----
    const GrbUMap um{ ukeys, uvals, us - 1 };
    uint32_t* vid = bd.vt_id + (uint64_t)bt * bd.vt_cap;
    uint32_t* vcnt = bd.vt_cnt + (uint64_t)bt * bd.vt_cap;
    const uint32_t n_old = bd.vt_n[bt];
    unsigned long long best = 0;
    for (uint32_t i = threadIdx.x; i < n_old; i += BS) {
      const uint32_t id = vid[i];
      uint32_t c = vcnt[i];
      uint32_t slot = grb_mix32(id) & tmask;
      while (dkeys[slot] != 0u) {
        if (dkeys[slot] == id) {
          const uint32_t d = dvals[slot];
          dvals[slot] = GRB_DELTA_FOUND;
          if (d) {
            c += d;
            vcnt[i] = c;
          }
          break;
        }
        slot = (slot + 1) & tmask;
      }
      if (c) {
        const unsigned long long key = ((unsigned long long)c << 32) | (0xFFFFFFFFu - id);
        best = key > best ? key : best;
        if (c > 2) {
          const uint32_t u = um.lookup(id);
          if (u != 0xFFFFFFFFu) {
            cm[(uint64_t)t * nu + u] = c;
          }
        }
      }
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < dsize; i += BS) { // ids the tile had not seen
      const uint32_t id = dkeys[i];
      const uint32_t c = dvals[i];
      if (id != 0u && c != GRB_DELTA_FOUND && c != 0u) {
        const uint32_t at = n_old + atomicAdd(&s_new, 1u);
        if (at < bd.vt_cap) {
          vid[at] = id;
          vcnt[at] = c;
        }
        const unsigned long long key = ((unsigned long long)c << 32) | (0xFFFFFFFFu - id);
        best = key > best ? key : best;
        if (c > 2) {
          const uint32_t u = um.lookup(id);
          if (u != 0xFFFFFFFFu) {
            cm[(uint64_t)t * nu + u] = c;
          }
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
      best = o > best ? o : best;
    }
    if ((threadIdx.x & 31) == 0 && best) {
      atomicMax(&s_best, best);
    }
    __syncthreads();
    // did any input of the smoothing passes change?  They read a count c only as c != 0 and
    // c > threshold (goldrush_path.cpp:628-682), and the arg-max count as > max(2, threshold).
    for (uint32_t u = threadIdx.x; u < nu; u += BS) {
      const uint32_t a = oldrow[u], c = cm[(uint64_t)t * nu + u];
      if ((a != 0) != (c != 0) || (a > prm.threshold) != (c > prm.threshold)) {
        s_chg = 1;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long bb = s_best;
      const uint32_t bid = bb ? 0xFFFFFFFFu - (uint32_t)(bb & 0xFFFFFFFFu) : 0u;
      const uint32_t bc = (uint32_t)(bb >> 32), oc = bd.best_count[bt];
      if (s_chg || bid != bd.best_id[bt] ||
          (bc > 2 && bc > prm.threshold) != (oc > 2 && oc > prm.threshold)) {
        bd.in_changed[b] = 1;
      }
      bd.best_count[bt] = bc;
      bd.best_id[bt] = bid;
      bd.vt_n[bt] = n_old + s_new;
      if (s_dhits) {
        atomicAdd(&bd.rd_hits[b], (uint32_t)s_dhits);
        atomicAdd(&bd.rd_miss[b], (uint32_t)(-s_dhits));
      }
      if (um.lookup(bid) == 0xFFFFFFFFu) {
        bd.u_changed[b] = 1; // the matrix columns no longer cover every arg-max id
      }
    }
  }
}

// CTA 0 only: smoothing passes over the re-validated votes of read b; result into bd.sp_*
template<int BS>
__device__ __forceinline__ void
grb_commit_resmooth(const GrbSelParams& prm, const GrbBatchDev& bd, const GrbCommitSmem& sm,
                    uint32_t b, uint32_t n, uint32_t len, uint32_t us, uint32_t cm_smem)
{
  const uint32_t bt0 = bd.tile_first[b];
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n; i += BS) {
    sm.best_id[i] = grb_ld(&bd.best_id[bt0 + i]);
    sm.best_cnt[i] = grb_ld(&bd.best_count[bt0 + i]);
  }
  __syncthreads();
  uint32_t nu;
  const uint32_t* cmat;
  if (grb_ld(&bd.u_changed[b])) {
    // a re-validated tile's arg-max is not a matrix column: rebuild columns and matrix from the
    // (updated) vote tables
    uint32_t* cm_out = cm_smem ? sm.cmat : bd.cmat;
    nu = grb_build_cmat<BS>(bd, bt0, n, sm.best_id, sm.root, sm.ukeys, sm.uvals, us, sm.uq, cm_out);
    cmat = cm_out;
  } else {
    nu = bd.nu[b];
    for (unsigned i = threadIdx.x; i < us; i += BS) {
      sm.uvals[i] = 0xFFFFFFFFu;
    }
    const uint32_t* cm_g = bd.cm + bd.cm_off[b];
    if (cm_smem) {
      for (uint32_t i = threadIdx.x; i < n * nu; i += BS) {
        sm.cmat[i] = grb_ld(&cm_g[i]);
      }
      cmat = sm.cmat;
    } else {
      cmat = cm_g;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (uint32_t u = 0; u < nu; ++u) {
        grb_umap_insert(sm.ukeys, sm.uvals, us - 1, bd.uq[bt0 + u], u);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const GrbMatrixVotes v{ sm.best_id, sm.best_cnt, cmat, GrbUMap{ sm.ukeys, sm.uvals, us - 1 }, nu };
    const uint32_t n_as = grb_smooth_tiles(n, v, prm.threshold, sm.tile_id, sm.tile_as, sm.snap);
    bd.sp_n_as[b] = n_as;
    uint32_t rel = 0;
    GrbReadPlan plan;
    grb_plan_read(n, n_as, len, prm.tile_len, prm.block_size, prm.unassigned_min, prm.assigned_max,
                  sm.tile_id, sm.tile_as, &rel, &plan);
    bd.sp_plan[b] = plan;
    bd.sp_adv[b] = rel;
  }
  __syncthreads();
}

// reservoir insert of one distinct rank for the insert calls (blocks) in tile mask m
__device__ __forceinline__ void
grb_apply_rank(const GrbFilterDev& filt, const GrbBatchDev& bd, const GrbReadPlan& plan, uint64_t key,
               uint64_t m, uint32_t B, uint32_t epoch)
{
  // another CTA may have rewritten the slot earlier in this launch: read it from L2
  const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(&filt.slots[key]));
  GrbSlot s{ raw.x, raw.y, raw.z, raw.w };
  const uint32_t orig = s.id;
  uint32_t last_j = 0xFFFFFFFFu;
  while (m) {
    const uint32_t t = __ffsll((long long)m) - 1;
    m &= m - 1;
    const uint32_t j = (t - plan.trim_start) / B;
    if (j == last_j) {
      continue; // same insert call: a rank counts once per call (MIBFConstructSupport.hpp:255-270)
    }
    last_j = j;
    const uint32_t id = plan.first_id + j + plan.id_bump;
    const uint32_t count = ++s.count;
    if ((uint32_t)(key ^ (uint64_t)id) % count == count - 1) { // MIBFConstructSupport.hpp:274-282
      s.id = s.id > GRB_SAT_MASK ? (id | GRB_SAT_MASK) : id;   // MIBloomFilter.hpp:593-602
    }
  }
  if (s.id != orig) {
    if (s.epoch != epoch) {
      s.id0 = orig;
      s.epoch = epoch;
    }
    const uint32_t hb = (uint32_t)key & bd.dirty_mask;
    atomicOr(&bd.dirty_bits[hb >> 5], 1u << (hb & 31));
  }
  filt.slots[key] = s;
}

// The ordered commit of one batch: ONE cooperative launch, n_cta = gridDim.x resident CTAs.
// Every CTA keeps its own copy of the loop state and takes the same decisions; CTA 0 publishes
// them.  barrier_ctr must be zero at launch.  dec_idx[b] = index of read b in `decisions`.
template<int BS>
__global__ void __launch_bounds__(BS, 1)
k_commit_batch(GrbReadsDev reads, GrbFilterDev filt, GrbSelParams prm, GrbBatchDev bd,
               GrbSelScratch sc, GrbSelState* __restrict__ state_g, grb_decision* __restrict__ decisions,
               const uint64_t* __restrict__ dec_idx, unsigned long long* barrier_ctr,
               uint32_t fb_words, uint32_t us, uint32_t n_cap, uint32_t cm_smem)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GrbCommitSmem sm;
  sm.fbits = reinterpret_cast<uint32_t*>(smem_raw);
  sm.dkeys = sm.fbits + fb_words;
  sm.dvals = sm.dkeys + prm.table_size;
  sm.ukeys = sm.dvals + prm.table_size;
  sm.uvals = sm.ukeys + us;
  sm.oldrow = sm.uvals + us;
  sm.best_id = sm.oldrow + us;
  sm.best_cnt = sm.best_id + n_cap;
  sm.root = sm.best_cnt + n_cap;
  sm.uq = sm.root + n_cap;
  sm.tile_id = sm.uq + n_cap;
  sm.snap = sm.tile_id + n_cap;
  sm.tile_as = reinterpret_cast<uint8_t*>(sm.snap + n_cap + 2);
  sm.cmat = reinterpret_cast<uint32_t*>(sm.tile_as + ((n_cap + 15) / 16) * 16);
  __shared__ GrbSelState st;
  __shared__ GrbReadPlan s_plan;
  __shared__ uint32_t s_redo;

  const uint32_t cta = blockIdx.x, n_cta = gridDim.x;
  if (threadIdx.x == 0) {
    st = *state_g;
  }
  __syncthreads();
  if (st.halt) {
    return;
  }
  unsigned long long phase = 0;
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h;
  const uint32_t B = (uint32_t)prm.block_size;

  for (uint32_t b = 0; b < bd.nb; ++b) {
    const uint64_t read_idx = bd.read_idx[b];
    const uint32_t len = reads.len[read_idx];
    const uint32_t n = len / T;
    const uint32_t bt0 = bd.tile_first[b];
    // ---- check ----
    long long t0 = clock64(), t1;
#define GRB_TICK(slot)                                                                             \
  t1 = clock64();                                                                                  \
  if (threadIdx.x == 0) {                                                                          \
    st.prof[slot] += (unsigned long long)(t1 - t0);                                                \
  }                                                                                                \
  t0 = t1;
    if (st.batch_inserts != 0) {
      grb_commit_check<BS>(reads, filt, prm, bd, sm, b, st.epoch, cta, n_cta, fb_words, us);
      __syncthreads();
      GRB_TICK(0)
      grb_grid_barrier(barrier_ctr, ++phase * n_cta);
      if (threadIdx.x == 0) {
        s_redo = grb_ld(&bd.u_changed[b]) | grb_ld(&bd.in_changed[b]);
        st.prof[9] += 1;
      }
      __syncthreads();
      GRB_TICK(1)
      if (s_redo) { // uniform over the grid: every CTA reads the same flags
        if (cta == 0) {
          grb_commit_resmooth<BS>(prm, bd, sm, b, n, len, us, cm_smem);
        }
        grb_grid_barrier(barrier_ctr, ++phase * n_cta);
        if (threadIdx.x == 0) {
          st.prof[7] += 1;
        }
        GRB_TICK(2)
      }
    }
    // ---- decide (every CTA, same inputs, same result) ----
    if (threadIdx.x == 0) {
      const uint32_t n_as = grb_ld(&bd.sp_n_as[b]);
      GrbSelState& s = st;
      s.cur.queries += grb_ld(&bd.rd_queries[b]);
      s.cur.hits += grb_ld(&bd.rd_hits[b]);
      s.cur.misses += grb_ld(&bd.rd_miss[b]);
      s.cur.total_tiles += n;
      s.cur.assigned_tiles += n_as;
      s.cur.unassigned_tiles += n - n_as;
      GrbReadPlan plan;
      {
        const uint4 lo4 = grb_ld(reinterpret_cast<const uint4*>(&bd.sp_plan[b]));
        const uint4 hi4 = grb_ld(reinterpret_cast<const uint4*>(&bd.sp_plan[b]) + 1);
        uint4* pw = reinterpret_cast<uint4*>(&plan);
        pw[0] = lo4;
        pw[1] = hi4;
      }
      if (plan.verdict == GRB_UNTRIMMED || plan.verdict == GRB_TRIMMED) {
        plan.first_id += s.ids_inserted;
        s.ids_inserted += grb_ld(&bd.sp_adv[b]);
      }
      if (cta == 0) {
        grb_decision d;
        d.verdict = plan.verdict;
        d.pad[0] = d.pad[1] = d.pad[2] = 0;
        d.path = (uint32_t)s.curr_path;
        d.trim_start = plan.trim_start;
        d.trim_end = plan.trim_end;
        d.num_tiles = n;
        d.num_assigned = n_as;
        decisions[dec_idx[b]] = d;
      }
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
      s.prof[6] += 1;
      s.prof[8] += plan.n_blocks != 0;
      s_plan = plan;
    }
    __syncthreads();
    GRB_TICK(3)
    if (st.halt) {
      break; // path rollover: the host resets the ID slots and resumes after this read
    }
    // ---- insert ----
    const GrbReadPlan plan = s_plan;
    if (plan.n_blocks != 0 && (plan.verdict == GRB_UNTRIMMED || plan.verdict == GRB_TRIMMED)) {
      const uint32_t size = bd.dd_size[b];
      if (size) {
        const uint64_t* keys = bd.dd_key + bd.dd_off[b];
        const uint64_t* masks = bd.dd_mask + bd.dd_off[b];
        const uint64_t range = (plan.trim_end >= 63 ? ~0ull : ((2ull << plan.trim_end) - 1ull)) &
                               ~((1ull << plan.trim_start) - 1ull);
        for (uint32_t i = cta * BS + threadIdx.x; i < size; i += n_cta * BS) {
          const uint64_t key = __ldcs(&keys[i]);
          if (key == GRB_EMPTY_KEY) {
            continue;
          }
          const uint64_t m = __ldcs(&masks[i]) & range;
          if (m) {
            grb_apply_rank(filt, bd, plan, key, m, B, st.epoch);
          }
        }
      } else {
        // reads of more than 64 tiles: de-duplicate now, 64 insert calls per round
        const uint64_t* stash = bd.stash + (uint64_t)bt0 * T * h;
        const uint64_t per_tile = (uint64_t)T * h;
        const uint32_t rounds = (plan.n_blocks + 63) / 64;
        for (uint32_t round = 0; round < rounds; ++round) {
          const uint64_t first_tile = plan.trim_start + 64ull * round * B;
          uint64_t last_tile = first_tile + 64ull * B - 1;
          if (last_tile > plan.trim_end) {
            last_tile = plan.trim_end;
          }
          const uint64_t total = (last_tile - first_tile + 1) * per_tile;
          uint64_t tab = 1;
          while (tab < 2 * total) {
            tab <<= 1;
          }
          const uint64_t mask = tab - 1;
          for (uint64_t idx = (uint64_t)cta * BS + threadIdx.x; idx < total;
               idx += (uint64_t)n_cta * BS) {
            const uint64_t trel = idx / per_tile;
            const uint32_t rem = (uint32_t)(idx - trel * per_tile);
            const uint32_t f = rem / h, p = rem - f * h;
            const uint32_t t = (uint32_t)(first_tile + trel);
            const uint32_t tl = grb_tile_bases(len, t, T, k);
            if (tl < k + p || f >= tl - (k + p) + 1) {
              continue;
            }
            const uint64_t key = stash[((uint64_t)t * T + f) * h + p] & ~GRB_STASH_NOFRAME;
            const uint32_t j = (uint32_t)((t - plan.trim_start) / B) - 64u * round;
            uint64_t slot = grb_mix64(key) & mask;
            while (true) {
              const unsigned long long old = atomicCAS(
                reinterpret_cast<unsigned long long*>(&sc.tab_key[slot]), GRB_EMPTY_KEY, key);
              if (old == GRB_EMPTY_KEY || old == key) {
                atomicOr(reinterpret_cast<unsigned long long*>(&sc.tab_mask[slot]), 1ull << j);
                break;
              }
              slot = (slot + 1) & mask;
            }
          }
          grb_grid_barrier(barrier_ctr, ++phase * n_cta);
          GrbReadPlan rp = plan; // block j of this round is insert call 64 * round + j
          rp.first_id = plan.first_id + 64u * round;
          rp.trim_start = 0;
          for (uint64_t i = (uint64_t)cta * BS + threadIdx.x; i < tab; i += (uint64_t)n_cta * BS) {
            const uint64_t key = grb_ld(&sc.tab_key[i]);
            if (key == GRB_EMPTY_KEY) {
              continue;
            }
            const uint64_t m = grb_ld(&sc.tab_mask[i]);
            // the mask holds insert-call bits here, not tile bits: block size 1 maps them 1:1
            grb_apply_rank(filt, bd, rp, key, m, 1u, st.epoch);
            sc.tab_key[i] = GRB_EMPTY_KEY;
            sc.tab_mask[i] = 0;
          }
          if (round + 1 < rounds) {
            grb_grid_barrier(barrier_ctr, ++phase * n_cta);
          }
        }
      }
    }
    __syncthreads();
    GRB_TICK(4)
    grb_grid_barrier(barrier_ctr, ++phase * n_cta);
    GRB_TICK(5)
  }
#undef GRB_TICK
  if (cta == 0 && threadIdx.x == 0) {
    *state_g = st;
  }
}
