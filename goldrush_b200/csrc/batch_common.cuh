// Shared pieces of the batch engine of the pass-2 loop (kernels_query.cuh: speculative query of a
// whole batch against the filter as it stands when the batch starts; kernels_commit.cuh: the
// ordered commit that re-validates, in file order, whatever two reads of one batch can share).
//
// Replaces the same reference code as the serial engine in kernels_select.cuh
// (read_hashing.cpp:7-75, goldrush_path.cpp:529-890 calc_num_assigned_tiles, :892-1094
// process_read, :156-187 silver_path_check, MIBFConstructSupport.hpp:247-283 insertMIBF) and
// produces the same decisions: the reference's loop-carried dependence (every query sees every
// earlier insert, goldrush_path.cpp:1229-1256) is preserved exactly, not approximately.
#pragma once
#include "common.cuh"
#include "decide.cuh"
#include "kernels_select.cuh"

#define GRB_STASH_NOFRAME (1ull << 63) // on pattern 0's rank: the frame failed the bit test
#define GRB_ST_SHARED (1ull << 62) // stash entry: low 32 bits = conflict index, not a rank
#define GRB_IX_EMPTY 0xFFFFFFFFFFFFFFFFull
#define GRB_FR_DEAD 0xFFFFFFFFu // frame record: the frame failed the bit test and never votes

struct GrbBatchDev
{
  const uint64_t* read_idx;   // [nb] store index of batch read b
  const uint32_t* tile_first; // [nb + 1] first batch tile of read b
  const uint32_t* tile_read;  // [n_bt] b of each batch tile
  uint32_t nb, n_bt;
  uint64_t* stash;      // [n_bt * tile_frames * h] rank of every probe
  uint32_t* best_id;    // [n_bt]
  uint32_t* best_count; // [n_bt]
  uint32_t* tile_hits;  // [n_bt]
  uint32_t* tile_miss;  // [n_bt]
  // per read: distinct arg-max ids of its tiles (uq[tile_first[b] + u], u < nu[b]) and the
  // count matrix cm[cm_off[b] + i * nu[b] + u] = votes of uq[u] in tile i if > 2, else 0
  // (global memory only for reads of more than 160 tiles)
  uint32_t* uq;           // [n_bt]
  uint32_t* nu;           // [nb]
  uint32_t* cm;
  const uint64_t* cm_off; // [nb]
  // smoothing result on the speculative votes
  uint32_t* sp_n_as;      // [nb]
  // decision of read b up to the ID counter: first_id is relative to ids_inserted (1 when the read
  // inserts), sp_adv[b] is what the read adds to ids_inserted (goldrush_path.cpp:982-994,1040-1053)
  GrbReadPlan* sp_plan;   // [nb]
  uint32_t* sp_adv;       // [nb]
  uint32_t* rd_hits;      // [nb] per-read totals of tile_hits / tile_miss / frames
  uint32_t* rd_miss;      // [nb]
  uint32_t* rd_queries;   // [nb]
};

// per-tile vote tables of a batch in global memory (written by k2_query, read by the commit)
struct GrbB2
{
  uint32_t* vk; // [n_bt * table_size] vote-table ids (0 = empty)
  uint32_t* vc; // [n_bt * table_size] vote-table counts
  uint32_t table_size;
  uint32_t pad;
};

__device__ __forceinline__ uint32_t
grb_norm_id(uint32_t v)
{
  return v > GRB_SAT_MASK ? (v & ~GRB_SAT_MASK) : v; // goldrush_path.cpp:574-583
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

__device__ __forceinline__ uint32_t
grb2_vote_get(const uint32_t* __restrict__ vk, const uint32_t* __restrict__ vc, uint32_t mask,
              uint32_t id)
{
  uint32_t slot = grb_mix32(id) & mask;
  for (uint32_t tries = 0; tries <= mask; ++tries) {
    const uint32_t kk = __ldcg(&vk[slot]);
    if (kk == 0u) {
      return 0u;
    }
    if (kk == id) {
      return __ldcg(&vc[slot]);
    }
    slot = (slot + 1) & mask;
  }
  return 0u;
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

// (tile, frame, pattern) of a stash index, and whether it is a valid (non-stale) position of its
// pattern: multiLensfrHashIterator.hpp:49-68 repeats the last value of an exhausted pattern, and
// insertMIBF de-duplicates it away (MIBFConstructSupport.hpp:255-270)
struct GrbProbeAt
{
  uint32_t bt, b, t, f, p, tl;
  bool valid;
};

__device__ __forceinline__ GrbProbeAt
grb2_probe_at(const GrbReadsDev& reads, const GrbSelParams& prm, const GrbBatchDev& bd, uint32_t idx)
{
  GrbProbeAt a;
  const uint32_t T = prm.tile_len, h = prm.h, k = prm.k;
  const uint32_t per_tile = prm.tile_frames * h;
  a.bt = idx / per_tile;
  const uint32_t rem = idx - a.bt * per_tile;
  a.f = rem / h;
  a.p = rem - a.f * h;
  a.b = bd.tile_read[a.bt];
  a.t = a.bt - bd.tile_first[a.b];
  a.tl = grb_tile_bases(reads.len[bd.read_idx[a.b]], a.t, T, prm.kmer);
  a.valid = a.tl >= k + a.p && a.f < a.tl - (k + a.p) + 1;
  return a;
}

__device__ __forceinline__ unsigned long long
grb2_pack_best(uint32_t c, uint32_t id)
{
  return c ? (((unsigned long long)c << 32) | (0xFFFFFFFFu - id)) : 0ull;
}

// reservoir insert (MIBFConstructSupport.hpp:274-282, MIBloomFilter.hpp:593-602) of one insert
// call into a {id, count} pair held in registers
__device__ __forceinline__ void
grb2_reservoir(uint64_t rank, uint32_t id, uint32_t& cur_id, uint32_t& cur_count)
{
  const uint32_t count = ++cur_count;
  if ((uint32_t)(rank ^ (uint64_t)id) % count == count - 1) {
    cur_id = cur_id > GRB_SAT_MASK ? (id | GRB_SAT_MASK) : id;
  }
}

// The ordered commit of one batch by ONE CTA.  dec_idx[b] = index of read b in `decisions`.
