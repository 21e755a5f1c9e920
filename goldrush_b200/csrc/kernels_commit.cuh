// Batch engine of the pass-2 loop: the ordered commit as a parallel fixed point.
//
// Replaces the same reference code as kernels_select.cuh (goldrush_path.cpp:529-890
// calc_num_assigned_tiles, :892-1094 process_read, :156-187 silver_path_check,
// MIBFConstructSupport.hpp:247-283 insertMIBF) and produces the same decisions: every query sees
// every earlier insert (goldrush_path.cpp:1229-1256), exactly.
//
// A rank probed by one valid probe of the batch is private (its reservoir
// update is applied after the commit, k3_bulk) and a rank probed twice or more is shared.  The
// members (probes) of a shared rank are sorted by (read, tile).  Given a PLAN for every read of the
// batch (insert or not, tile range, ids), the history of every shared rank follows by walking its
// members in order -- independently per rank -- and with it the id every member sees when its read
// is queried.  From those ids every read's votes, smoothing and plan follow -- independently per
// read.  The plans assumed at the start are the speculative ones; an iteration
//     E  per shared rank: walk members in read order under the assumed plans -> id seen by each
//     F  per read: vote deltas of its conflict frames, re-smoothing if an input changed -> new plan
//     G  in read order: ids / path bookkeeping, first read whose new plan differs from the assumed
// makes every read up to and including the first mismatch final (its inputs depended on final
// plans only), and the loop ends when no assumed plan is contradicted: a fixed point that equals
// the serial order by induction over the reads.  One persistent cooperative launch per batch
// (k3_fix) runs the iterations with grid barriers; typically 1-3 are needed.
#pragma once
#include "common.cuh"
#include "decide.cuh"
#include "kernels_select.cuh"
#include "batch_common.cuh"
#include "kernels_query.cuh"

struct __align__(16) GrbShared3
{
  uint64_t rank;
  uint32_t id0;    // raw slot id / count when the batch started
  uint32_t count0;
  uint32_t off;    // first member in m_key / m_seen
  uint32_t n;      // members
  uint32_t id;     // raw id / count after the batch (last walk)
  uint32_t count;
};

struct GrbFixCtl
{
  uint32_t final_upto; // reads [0, final_upto) have final plans
  uint32_t n_commit;   // reads [0, n_commit) are committed by this batch (path rollover cuts it)
  uint32_t first_ins;  // first read whose assumed plan inserts (reads up to it see no change)
  uint32_t iter;
};

struct GrbB3
{
  uint32_t* vk; // [n_bt * table_size] vote-table ids (0 = empty), speculative, read only here
  uint32_t* vc;
  uint32_t table_size;
  uint32_t d_cap;             // per-CTA global delta table entries (power of two)
  uint32_t* bm;               // rank bit map of the batch, bm_mask + 1 bits
  uint32_t bm_mask;
  uint32_t* cand;             // probes whose bit was already set
  unsigned long long* ix_tab; // exact set of the candidates' ranks (GRB_IX_EMPTY = free)
  uint64_t ix_mask;           // capacity - 1; the part in use follows the candidate count
  uint32_t* ix_cnt;   // per set slot: probes of the batch with that rank
  uint32_t* ix_sidx;  // per set slot: shared-table index (count >= 2)
  uint32_t* counters; // GRB_CTR_*
  uint32_t* t_probe;  // probes found in the set, and their set slot
  uint32_t* t_slot;
  uint32_t* c_probe;  // per conflict: probe (stash index), shared index, member position
  uint32_t* c_sidx;
  uint32_t* c_pos;
  GrbShared3* shared;
  uint32_t* m_fill; // [shared] scatter cursor
  uint32_t* m_key;  // [members] read << 16 | tile, sorted within a shared rank
  uint32_t* m_ci;   // [members] conflict index
  uint32_t* m_seen; // [members] raw id the member's read sees under the assumed plans
  uint32_t* fbits;  // [n_bt * tile_frames bits] frame already listed
  uint32_t* fl_n;   // [nb] conflict frames of read b
  uint32_t* fl;     // frame list, read b's region starts at tile_first[b] * tile_frames: tile << 20 | frame
  uint32_t* fr;     // frame records, (2 + 2h) words: tf, shared mask, old id[h], member pos[h]
  GrbReadPlan* np;  // [nb] new plan (ids relative), its id advance, assigned tiles, hit delta
  uint32_t* np_adv;
  uint32_t* np_nas;
  int32_t* np_dh;
  GrbReadPlan* plan_out;        // [nb] assumed / final plans with absolute ids
  GrbFixCtl* ctl;
  unsigned long long* d_keys;   // [n_cta * d_cap] global delta tables (heavy reads only)
  int32_t* d_vals;
  uint32_t* cmat_g;             // [n_cta * n_cap^2] when the count matrix does not fit shared memory
  unsigned long long* barrier;  // grid barrier counter, zero at launch
  uint32_t* dbg;                // GRB_FIX_DEBUG: [0] cursor, then 4-word records per re-validated read
  uint32_t dbg_cap;
};

// ---------------------------------------------------------------------------------------------
// batch index and conflict bookkeeping
//
// Which ranks are probed more than once in this batch?  A 64-bit CAS table over all ~10^7 probes is
// larger than L2 and ran at a third of the device's atomic rate, so the question is answered in
// three L2-resident steps:
//   k3_mark     every valid probe sets bit mix(rank) of a 2^28-bit map (32 MB); a probe that finds
//               its bit already set is a CANDIDATE (a later duplicate, or a false positive)
//   k3_dupset   the candidates' ranks enter a small exact hash set (a few MB)
//   k3_members  every valid probe looks its rank up in the set; the probes found are counted per
//               rank and listed.  A rank counted twice or more is shared, the others were false
//               positives of the bit map and stay private.
//   k3_open     one shared entry per rank counted >= 2
//   k3_conf     the listed probes of shared ranks become the conflict list (as before)
// ---------------------------------------------------------------------------------------------
#define GRB_FRAME_SPLIT 8 // CTAs sharing one read's conflict-frame list in k3_frames
#define GRB_CTR_CONF 0   // conflicts
#define GRB_CTR_SHARED 1 // shared ranks
#define GRB_CTR_MEMBER 2 // member cursor
#define GRB_CTR_CAND 3   // candidates of k3_mark
#define GRB_CTR_LISTED 4 // probes listed by k3_members

// start of a batch: everything the index and the commit expect zeroed, in one launch (it was eight
// memsets and a one-thread kernel: launch gaps were 5 % of the step)
__global__ void __launch_bounds__(256)
k3_reset(GrbB3 b3, GrbSelState* __restrict__ state, uint32_t bm_words, uint32_t fb_words, uint32_t nb)
{
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  uint4* bm4 = reinterpret_cast<uint4*>(b3.bm);
  for (uint32_t i = tid; i < bm_words / 4; i += nth) {
    bm4[i] = z;
  }
  for (uint32_t i = tid; i < fb_words; i += nth) {
    b3.fbits[i] = 0u;
  }
  uint4* po = reinterpret_cast<uint4*>(b3.plan_out);
  for (uint32_t i = tid; i < nb * (uint32_t)(sizeof(GrbReadPlan) / 16); i += nth) {
    po[i] = z;
  }
  for (uint32_t i = tid; i < nb; i += nth) {
    b3.fl_n[i] = 0u;
  }
  if (tid < 8) {
    b3.counters[tid] = 0u;
  }
  if (tid == 8) {
    *reinterpret_cast<uint4*>(b3.ctl) = z;
    *b3.barrier = 0ull;
    if (!state->halt) {
      state->batch_inserts = 0;
    }
  }
}

// append `item` (when `take`) to a global list with one atomic per CTA round
__device__ __forceinline__ void
grb3_block_append(bool take, uint32_t item, uint32_t item2, uint32_t* list, uint32_t* list2,
                  uint32_t* counter, uint32_t* s_n, uint32_t* s_base)
{
  const unsigned lane = threadIdx.x & 31;
  const unsigned m = __ballot_sync(0xffffffffu, take);
  uint32_t wbase = 0;
  if (lane == 0 && m) {
    wbase = atomicAdd(s_n, (uint32_t)__popc(m));
  }
  wbase = __shfl_sync(0xffffffffu, wbase, 0);
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t n = *s_n;
    *s_base = n ? atomicAdd(counter, n) : 0u;
    *s_n = 0;
  }
  __syncthreads();
  if (take) {
    const uint32_t at = *s_base + wbase + __popc(m & ((1u << lane) - 1u));
    list[at] = item;
    if (list2) {
      list2[at] = item2;
    }
  }
}

// counter += 1 for every calling lane, one atomic per group of lanes that name the same counter;
// returns the caller's slot.  Call from divergent code: the group is whoever is here right now.
__device__ __forceinline__ uint32_t
grb3_agg_inc(uint32_t* counter)
{
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, (unsigned long long)counter);
  const unsigned lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) {
    base = atomicAdd(counter, (uint32_t)__popc(peers));
  }
  base = __shfl_sync(peers, base, leader);
  return base + __popc(peers & ((1u << lane) - 1u));
}

__device__ __forceinline__ uint32_t
grb3_set_mask(uint32_t n_cand, uint32_t cap_mask)
{
  uint32_t want = 1024;
  while (want < 2u * n_cand && want - 1 < cap_mask) {
    want <<= 1;
  }
  return (want - 1) & cap_mask;
}

// Dynamic shared memory: uint32 stage[T * h].  The candidates of a tile are collected in shared
// memory and appended with one reservation per tile (the per-256-probe block append spent as long
// in its two barriers as in the atomics: ncu, 21 barrier-stall cycles per issue).
__global__ void __launch_bounds__(256)
k3_mark(GrbReadsDev reads, GrbSelParams prm, GrbBatchDev bd, GrbB3 b3,
        const GrbSelState* __restrict__ state)
{
  extern __shared__ uint32_t k3_stage[];
  __shared__ uint32_t s_n, s_base;
  if (state->halt) {
    return;
  }
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h;
  const uint32_t per_tile = prm.tile_frames * h;
  for (uint32_t bt = blockIdx.x; bt < bd.n_bt; bt += gridDim.x) {
    const uint32_t b = bd.tile_read[bt];
    const uint32_t t = bt - bd.tile_first[b];
    const uint32_t tl = grb_tile_bases(reads.len[bd.read_idx[b]], t, T, prm.kmer);
    if (threadIdx.x == 0) {
      s_n = 0;
    }
    __syncthreads();
    for (uint32_t rem = threadIdx.x; rem < per_tile; rem += blockDim.x) {
      const uint32_t f = rem / h, p = rem - f * h;
      if (tl >= k + p && f < tl - (k + p) + 1) {
        const uint32_t idx = bt * per_tile + rem;
        const uint64_t rank = __ldcs(&bd.stash[idx]) & ~GRB_STASH_NOFRAME;
        const uint32_t bit = (uint32_t)(grb_mix64(rank) >> 20) & b3.bm_mask;
        const uint32_t m = 1u << (bit & 31);
        if (atomicOr(&b3.bm[bit >> 5], m) & m) {
          k3_stage[atomicAdd(&s_n, 1u)] = idx;
        }
      }
    }
    __syncthreads();
    const uint32_t n = s_n;
    if (threadIdx.x == 0 && n) {
      s_base = atomicAdd(&b3.counters[GRB_CTR_CAND], n);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      b3.cand[s_base + i] = k3_stage[i];
    }
  }
}

// clears the part of the exact set this batch will use (its size follows the candidate count)
__global__ void __launch_bounds__(256)
k3_set_clear(GrbB3 b3, const GrbSelState* __restrict__ state)
{
  if (state->halt) {
    return;
  }
  const uint32_t mask = grb3_set_mask(b3.counters[GRB_CTR_CAND], (uint32_t)b3.ix_mask);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= mask; i += gridDim.x * blockDim.x) {
    b3.ix_tab[i] = GRB_IX_EMPTY;
    b3.ix_cnt[i] = 0;
  }
}

__global__ void __launch_bounds__(256)
k3_dupset(GrbBatchDev bd, GrbB3 b3, const GrbSelState* __restrict__ state)
{
  if (state->halt) {
    return;
  }
  const uint32_t n_cand = b3.counters[GRB_CTR_CAND];
  const uint32_t mask = grb3_set_mask(n_cand, (uint32_t)b3.ix_mask);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cand; i += gridDim.x * blockDim.x) {
    const uint64_t rank = bd.stash[b3.cand[i]] & ~GRB_STASH_NOFRAME;
    uint32_t slot = (uint32_t)grb_mix64(rank) & mask;
    while (true) {
      const unsigned long long old = atomicCAS(&b3.ix_tab[slot], GRB_IX_EMPTY, rank);
      if (old == GRB_IX_EMPTY || old == rank) {
        break;
      }
      slot = (slot + 1) & mask;
    }
  }
}

// Dynamic shared memory: uint32 probe[T * h] | slot[T * h] (per-tile staging, as in k3_mark)
__global__ void __launch_bounds__(256)
k3_members(GrbReadsDev reads, GrbSelParams prm, GrbBatchDev bd, GrbB3 b3,
           const GrbSelState* __restrict__ state)
{
  extern __shared__ uint32_t k3_stage[];
  __shared__ uint32_t s_n, s_base;
  if (state->halt) {
    return;
  }
  const uint32_t n_cand = b3.counters[GRB_CTR_CAND];
  if (n_cand == 0) {
    return;
  }
  const uint32_t mask = grb3_set_mask(n_cand, (uint32_t)b3.ix_mask);
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h;
  const uint32_t per_tile = prm.tile_frames * h;
  uint32_t* st_probe = k3_stage;
  uint32_t* st_slot = k3_stage + per_tile;
  for (uint32_t bt = blockIdx.x; bt < bd.n_bt; bt += gridDim.x) {
    const uint32_t b = bd.tile_read[bt];
    const uint32_t t = bt - bd.tile_first[b];
    const uint32_t tl = grb_tile_bases(reads.len[bd.read_idx[b]], t, T, prm.kmer);
    if (threadIdx.x == 0) {
      s_n = 0;
    }
    __syncthreads();
    for (uint32_t rem = threadIdx.x; rem < per_tile; rem += blockDim.x) {
      const uint32_t f = rem / h, p = rem - f * h;
      if (tl >= k + p && f < tl - (k + p) + 1) {
        const uint32_t idx = bt * per_tile + rem;
        const uint64_t rank = __ldcs(&bd.stash[idx]) & ~GRB_STASH_NOFRAME;
        uint32_t slot = (uint32_t)grb_mix64(rank) & mask;
        bool take = false;
        while (true) {
          const unsigned long long key = __ldcg(&b3.ix_tab[slot]);
          if (key == rank) {
            take = true;
            break;
          }
          if (key == GRB_IX_EMPTY) {
            break;
          }
          slot = (slot + 1) & mask;
        }
        if (take) {
          atomicAdd(&b3.ix_cnt[slot], 1u);
          const uint32_t at = atomicAdd(&s_n, 1u);
          st_probe[at] = idx;
          st_slot[at] = slot;
        }
      }
    }
    __syncthreads();
    const uint32_t n = s_n;
    if (threadIdx.x == 0 && n) {
      s_base = atomicAdd(&b3.counters[GRB_CTR_LISTED], n);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      b3.t_probe[s_base + i] = st_probe[i];
      b3.t_slot[s_base + i] = st_slot[i];
    }
  }
}

__global__ void __launch_bounds__(256)
k3_open(GrbFilterDev filt, GrbB3 b3, const GrbSelState* __restrict__ state)
{
  if (state->halt) {
    return;
  }
  const uint32_t n_cand = b3.counters[GRB_CTR_CAND];
  if (n_cand == 0) {
    return;
  }
  const uint32_t mask = grb3_set_mask(n_cand, (uint32_t)b3.ix_mask);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= mask; i += gridDim.x * blockDim.x) {
    const uint32_t n = b3.ix_cnt[i];
    if (n < 2) {
      continue;
    }
    const uint64_t rank = b3.ix_tab[i];
    const uint32_t sidx = grb3_agg_inc(&b3.counters[GRB_CTR_SHARED]);
    b3.ix_sidx[i] = sidx;
    const uint2 raw = __ldcg(reinterpret_cast<const uint2*>(&filt.slots[rank]));
    GrbShared3 e;
    e.rank = rank;
    e.id0 = raw.x;
    e.count0 = raw.y;
    e.off = 0;
    e.n = n;
    e.id = raw.x;
    e.count = raw.y;
    b3.shared[sidx] = e;
    b3.m_fill[sidx] = 0;
  }
}

// Per listed probe of a shared rank: it becomes a conflict; mark its stash entries (stale-tail
// repeats included) with its conflict index, and list the frames it votes in, once each.
__global__ void __launch_bounds__(256)
k3_conf(GrbReadsDev reads, GrbSelParams prm, GrbBatchDev bd, GrbB3 b3,
        const GrbSelState* __restrict__ state)
{
  if (state->halt) {
    return;
  }
  const uint32_t T = prm.tile_frames, k = prm.k, h = prm.h; // T: per-tile frame stride
  const uint32_t n_listed = b3.counters[GRB_CTR_LISTED];
  for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < n_listed; li += gridDim.x * blockDim.x) {
    const uint32_t slot = b3.t_slot[li];
    if (b3.ix_cnt[slot] < 2) {
      continue; // false positive of the bit map: the rank is private
    }
    const uint32_t sidx = b3.ix_sidx[slot];
    const uint32_t probe = b3.t_probe[li];
    const uint32_t ci = grb3_agg_inc(&b3.counters[GRB_CTR_CONF]);
    b3.c_probe[ci] = probe;
    b3.c_sidx[ci] = sidx;
    const GrbProbeAt a = grb2_probe_at(reads, prm, bd, probe);
    const uint32_t frames = a.tl - k + 1;
    const uint32_t n_p = a.tl - (k + a.p) + 1;
    const uint32_t f_hi = (a.f == n_p - 1) ? frames - 1 : a.f;
    for (uint32_t ff = a.f; ff <= f_hi; ++ff) {
      uint64_t* e = bd.stash + ((uint64_t)a.bt * T + ff) * h + a.p;
      *e = (*e & GRB_STASH_NOFRAME) | GRB_ST_SHARED | ci;
      const uint32_t g = a.bt * T + ff;
      const uint32_t bit = 1u << (g & 31);
      if (!(atomicOr(&b3.fbits[g >> 5], bit) & bit)) {
        const uint32_t at = grb3_agg_inc(&b3.fl_n[a.b]);
        b3.fl[(uint64_t)bd.tile_first[a.b] * T + at] = (a.t << 20) | ff;
      }
    }
  }
}

// member segments: any order of the segments will do
__global__ void __launch_bounds__(256)
k3_seg(GrbB3 b3, const GrbSelState* __restrict__ state)
{
  if (state->halt) {
    return;
  }
  const uint32_t n_shared = b3.counters[1];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_shared; i += gridDim.x * blockDim.x) {
    b3.shared[i].off = atomicAdd(&b3.counters[2], b3.shared[i].n);
  }
}

__global__ void __launch_bounds__(256)
k3_scatter(GrbSelParams prm, GrbBatchDev bd, GrbB3 b3, const GrbSelState* __restrict__ state)
{
  if (state->halt) {
    return;
  }
  const uint32_t per_tile = prm.tile_frames * prm.h;
  const uint32_t n_conf = b3.counters[0];
  for (uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x; ci < n_conf; ci += gridDim.x * blockDim.x) {
    const uint32_t sidx = b3.c_sidx[ci];
    const uint32_t bt = b3.c_probe[ci] / per_tile;
    const uint32_t b = bd.tile_read[bt];
    const uint32_t pos = b3.shared[sidx].off + atomicAdd(&b3.m_fill[sidx], 1u);
    b3.m_key[pos] = (b << 16) | (bt - bd.tile_first[b]);
    b3.m_ci[pos] = ci;
  }
}

// sort every member segment by (read, tile); then every conflict learns its member position
__global__ void __launch_bounds__(256)
k3_sort(GrbB3 b3, const GrbSelState* __restrict__ state)
{
  if (state->halt) {
    return;
  }
  const uint32_t n_shared = b3.counters[1];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_shared; i += gridDim.x * blockDim.x) {
    const uint32_t off = b3.shared[i].off, n = b3.shared[i].n;
    uint32_t* key = b3.m_key + off;
    uint32_t* val = b3.m_ci + off;
    if (n <= 24) {
      for (uint32_t a = 1; a < n; ++a) { // insertion sort
        const uint32_t kk = key[a], vv = val[a];
        uint32_t j = a;
        while (j > 0 && key[j - 1] > kk) {
          key[j] = key[j - 1];
          val[j] = val[j - 1];
          --j;
        }
        key[j] = kk;
        val[j] = vv;
      }
    } else { // heap sort: a rank shared by many probes (a repeat) must not cost n^2
      auto sift = [&](uint32_t root, uint32_t end) {
        while (true) {
          uint32_t child = 2 * root + 1;
          if (child >= end) {
            break;
          }
          if (child + 1 < end && key[child] < key[child + 1]) {
            ++child;
          }
          if (key[root] >= key[child]) {
            break;
          }
          const uint32_t tk = key[root], tv = val[root];
          key[root] = key[child];
          val[root] = val[child];
          key[child] = tk;
          val[child] = tv;
          root = child;
        }
      };
      for (uint32_t s = n / 2; s-- > 0;) {
        sift(s, n);
      }
      for (uint32_t end = n - 1; end > 0; --end) {
        const uint32_t tk = key[0], tv = val[0];
        key[0] = key[end];
        val[0] = val[end];
        key[end] = tk;
        val[end] = tv;
        sift(0, end);
      }
    }
    for (uint32_t a = 0; a < n; ++a) {
      b3.c_pos[val[a]] = off + a;
    }
  }
}

// per listed frame: its record {tile << 20 | frame, mask of shared patterns, per pattern the
// normalised id at batch start, per pattern the member position (shared patterns only)}
__global__ void __launch_bounds__(256)
k3_frames(GrbFilterDev filt, GrbSelParams prm, GrbBatchDev bd, GrbB3 b3,
          const GrbSelState* __restrict__ state)
{
  if (state->halt) {
    return;
  }
  const uint32_t T = prm.tile_frames, h = prm.h; // T: per-tile frame stride
  const uint32_t stride = 2 + 2 * h;
  // a read's frame list is shared by GRB_FRAME_SPLIT CTAs: one CTA per read left the GPU at 18 %
  // of its warps on records that are random slot reads (ncu)
  for (uint32_t u = blockIdx.x; u < bd.nb * GRB_FRAME_SPLIT; u += gridDim.x) {
    const uint32_t b = u / GRB_FRAME_SPLIT, part = u - b * GRB_FRAME_SPLIT;
    const uint32_t nfr = b3.fl_n[b];
    const uint32_t bt0 = bd.tile_first[b];
    const uint64_t off = (uint64_t)bt0 * T;
    for (uint32_t i = part * blockDim.x + threadIdx.x; i < nfr; i += blockDim.x * GRB_FRAME_SPLIT) {
      const uint32_t tf = b3.fl[off + i];
      const uint32_t t = tf >> 20, f = tf & 0xFFFFFu;
      const uint64_t* e = bd.stash + ((uint64_t)(bt0 + t) * T + f) * h;
      uint32_t* rec = b3.fr + (off + i) * stride;
      uint32_t smask = 0;
      const bool dead = (e[0] & GRB_STASH_NOFRAME) != 0;
      for (uint32_t p = 0; p < h; ++p) {
        const uint64_t v = e[p];
        uint32_t old = 0, pos = 0;
        if (!dead) {
          if (v & GRB_ST_SHARED) {
            const uint32_t ci = (uint32_t)v;
            smask |= 1u << p;
            old = grb_norm_id(b3.shared[b3.c_sidx[ci]].id0);
            pos = b3.c_pos[ci];
          } else {
            old = grb_norm_id(__ldcg(&filt.slots[v & ~GRB_STASH_NOFRAME].id));
          }
        }
        rec[2 + p] = old;
        rec[2 + h + p] = pos;
      }
      rec[0] = tf;
      rec[1] = dead ? GRB_FR_DEAD : smask;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the fixed-point kernel
// ---------------------------------------------------------------------------------------------
#define GRB_DK_EMPTY 0xFFFFFFFFFFFFFFFFull

// delta table: (tile << 32 | id) -> signed vote change; shared memory (light reads) or the CTA's
// global scratch (heavy reads, one tile at a time)
struct GrbDelta
{
  unsigned long long* keys;
  int32_t* vals;
  uint32_t mask;
  __device__ __forceinline__ bool add(uint32_t t, uint32_t id, int32_t d) const
  {
    const unsigned long long key = ((unsigned long long)t << 32) | id;
    uint32_t slot = (grb_mix32(id) + t * 0x9E3779B1u) & mask;
    for (uint32_t tries = 0; tries <= mask; ++tries) {
      const unsigned long long old = atomicCAS(&keys[slot], GRB_DK_EMPTY, key);
      if (old == GRB_DK_EMPTY || old == key) {
        atomicAdd(&vals[slot], d);
        return true;
      }
      slot = (slot + 1) & mask;
    }
    return false;
  }
  __device__ __forceinline__ int32_t get(uint32_t t, uint32_t id) const
  {
    const unsigned long long key = ((unsigned long long)t << 32) | id;
    uint32_t slot = (grb_mix32(id) + t * 0x9E3779B1u) & mask;
    for (uint32_t tries = 0; tries <= mask; ++tries) {
      const unsigned long long kk = keys[slot];
      if (kk == GRB_DK_EMPTY) {
        return 0;
      }
      if (kk == key) {
        return vals[slot];
      }
      slot = (slot + 1) & mask;
    }
    return 0;
  }
};

struct GrbFixSmem
{
  unsigned long long* dk; // [dc] delta keys   (G overlays its plan arrays here)
  int32_t* dv;            // [dc]
  unsigned long long* nbest; // [n_cap]
  uint32_t* ukeys;        // [us]
  uint32_t* uvals;        // [us]
  uint32_t* best_id;      // [n_cap] speculative arg-max, then the re-validated one
  uint32_t* best_cnt;
  uint32_t* root;
  uint32_t* uq;
  uint32_t* tile_id;
  uint32_t* snap;         // [n_cap + 2]
  uint32_t* rescan;       // [n_cap]
  uint8_t* tile_as;       // [n_cap] padded to 16
  uint32_t* cmat;         // [n_cap * n_cap] when cm_smem
};

__device__ __forceinline__ bool
grb3_plan_differs(const GrbReadPlan& a, const GrbReadPlan& b)
{
  const bool ia = a.verdict == GRB_UNTRIMMED || a.verdict == GRB_TRIMMED;
  const bool ib = b.verdict == GRB_UNTRIMMED || b.verdict == GRB_TRIMMED;
  if (a.verdict != b.verdict) {
    return true;
  }
  if (!ia && !ib) {
    return false;
  }
  return a.trim_start != b.trim_start || a.trim_end != b.trim_end || a.first_id != b.first_id ||
         a.id_bump != b.id_bump || a.n_blocks != b.n_blocks;
}


__device__ __forceinline__ GrbReadPlan
grb3_ld_plan(const GrbReadPlan* p)
{
  GrbReadPlan r;
  const uint4 a = __ldcg(reinterpret_cast<const uint4*>(p));
  const uint4 b = __ldcg(reinterpret_cast<const uint4*>(p) + 1);
  uint4* w = reinterpret_cast<uint4*>(&r);
  w[0] = a;
  w[1] = b;
  return r;
}

__device__ __forceinline__ void
grb3_st_plan(GrbReadPlan* p, const GrbReadPlan& r)
{
  const uint4* w = reinterpret_cast<const uint4*>(&r);
  __stcg(reinterpret_cast<uint4*>(p), w[0]);
  __stcg(reinterpret_cast<uint4*>(p) + 1, w[1]);
}

__device__ __forceinline__ bool
grb3_inserts(const GrbReadPlan& p)
{
  return (p.verdict == GRB_UNTRIMMED || p.verdict == GRB_TRIMMED) && p.n_blocks != 0;
}

// E: history of every shared rank under the assumed plans (MIBFConstructSupport.hpp:271-282 per
// insert call, calls in read / block order; a rank counts once per call, :255-270)
__device__ __forceinline__ void
grb3_walk(const GrbB3& b3, uint32_t B, uint32_t n_commit)
{
  const uint32_t n_shared = __ldcg(&b3.counters[1]);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_shared; i += gridDim.x * blockDim.x) {
    GrbShared3* e = &b3.shared[i];
    const uint4 lo = *reinterpret_cast<const uint4*>(e);        // rank, id0, count0
    const uint2 seg = *(reinterpret_cast<const uint2*>(e) + 2); // off, n
    const uint64_t rank = ((uint64_t)lo.y << 32) | lo.x;
    uint32_t cur_id = lo.z, cur_count = lo.w;
    const uint32_t* key = b3.m_key + seg.x;
    uint32_t* seen = b3.m_seen + seg.x;
    uint32_t a = 0;
    while (a < seg.y) {
      const uint32_t b = key[a] >> 16;
      uint32_t z = a;
      while (z < seg.y && (key[z] >> 16) == b) {
        __stcg(&seen[z], cur_id);
        ++z;
      }
      if (b < n_commit) {
        const GrbReadPlan plan = grb3_ld_plan(&b3.plan_out[b]);
        if (grb3_inserts(plan)) {
          uint32_t last_j = 0xFFFFFFFFu;
          for (uint32_t q = a; q < z; ++q) {
            const uint32_t t = key[q] & 0xFFFFu;
            if (t < plan.trim_start || t > plan.trim_end) {
              continue;
            }
            const uint32_t j = (t - plan.trim_start) / B;
            if (j != last_j) {
              grb2_reservoir(rank, plan.first_id + j + plan.id_bump, cur_id, cur_count);
              last_j = j;
            }
          }
        }
      }
      a = z;
    }
    e->id = cur_id;
    e->count = cur_count;
  }
}

// F: one read under the ids its conflict frames see now.  Whole CTA.
template<int BS>
__device__ __forceinline__ uint32_t
grb3_read(const GrbReadsDev& reads, const GrbSelParams& prm, const GrbBatchDev& bd, const GrbB3& b3,
          const GrbFixSmem& sm, uint32_t b, uint32_t us, uint32_t dc, uint32_t cm_smem, uint32_t n_cap)
{
  __shared__ uint32_t s_changed, s_uchg, s_over, s_any_rescan;
  __shared__ int s_dh;
  constexpr uint32_t ALL = 0xFFFFFFFFu;
  const uint32_t T = prm.tile_len, h = prm.h;
  const uint32_t vts = b3.table_size, vmask = vts - 1;
  const uint32_t stride = 2 + 2 * h;
  const uint32_t thr_hi = prm.threshold > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)prm.threshold;
  const uint32_t len = reads.len[bd.read_idx[b]];
  const uint32_t n = len / T;
  const uint32_t bt0 = bd.tile_first[b];
  const uint32_t nfr = b3.fl_n[b];
  const uint32_t nu = bd.nu[b];
  const uint32_t* fr = b3.fr + (uint64_t)bt0 * prm.tile_frames * stride;
  uint32_t* cmat = cm_smem ? sm.cmat : b3.cmat_g + (uint64_t)blockIdx.x * n_cap * n_cap;

  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n; i += BS) {
    const uint32_t bi = bd.best_id[bt0 + i], bc = bd.best_count[bt0 + i];
    sm.best_id[i] = bi;
    sm.best_cnt[i] = bc;
    sm.nbest[i] = grb2_pack_best(bc, bi);
    sm.rescan[i] = 0;
  }
  for (unsigned i = threadIdx.x; i < us; i += BS) {
    sm.uvals[i] = 0xFFFFFFFFu;
  }
  if (threadIdx.x == 0) {
    s_changed = 0;
    s_uchg = 0;
    s_over = 0;
    s_any_rescan = 0;
    s_dh = 0;
  }
  __syncthreads();
  for (uint32_t u = threadIdx.x; u < nu; u += BS) {
    const uint32_t id = bd.uq[bt0 + u];
    sm.uq[u] = id;
    grb_umap_insert_par(sm.ukeys, sm.uvals, us - 1, id, u);
  }
  const GrbUMap um{ sm.ukeys, sm.uvals, us - 1 };
  const GrbDelta d_light{ sm.dk, sm.dv, dc - 1 };
  const GrbDelta d_heavy{ b3.d_keys + (uint64_t)blockIdx.x * b3.d_cap,
                          b3.d_vals + (uint64_t)blockIdx.x * b3.d_cap, b3.d_cap - 1 };

  // vote deltas of the frames of tile `tsel` (ALL: every tile) whose ids changed, into D
  auto clear = [&](const GrbDelta& D) {
    for (uint32_t i = threadIdx.x; i <= D.mask; i += BS) {
      D.keys[i] = GRB_DK_EMPTY;
      D.vals[i] = 0;
    }
  };
  auto pass_a = [&](uint32_t tsel, const GrbDelta& D, bool count_dh) {
    int dh = 0;
    bool ok = true;
    for (uint32_t i = threadIdx.x; i < nfr; i += BS) {
      const uint32_t* rec = fr + (uint64_t)i * stride;
      const uint32_t smask = __ldg(&rec[1]);
      if (smask == GRB_FR_DEAD) {
        continue;
      }
      const uint32_t t = __ldg(&rec[0]) >> 20;
      if (tsel != ALL && t != tsel) {
        continue;
      }
      uint32_t oldv[GRB_MAX_PATTERNS], newv[GRB_MAX_PATTERNS];
      bool any = false;
#pragma unroll
      for (unsigned p = 0; p < GRB_MAX_PATTERNS; ++p) {
        if (p < h) {
          oldv[p] = __ldg(&rec[2 + p]);
          newv[p] = oldv[p];
          if ((smask >> p) & 1u) {
            newv[p] = grb_norm_id(__ldcg(&b3.m_seen[__ldg(&rec[2 + h + p])]));
            any = any || newv[p] != oldv[p];
          }
        }
      }
      if (!any) {
        continue;
      }
#pragma unroll
      for (unsigned p = 0; p < GRB_MAX_PATTERNS; ++p) {
        if (p < h) {
          // an id leaves the frame's set unless it is still one of the new ids, and joins it
          // unless it was one of the old ids (the set is what counts, goldrush_path.cpp:570-604)
          const uint32_t o = oldv[p], nw = newv[p];
          dh += (nw != 0) - (o != 0);
          bool first = o != 0, stays = false, nfirst = nw != 0, was = false;
#pragma unroll
          for (unsigned j = 0; j < GRB_MAX_PATTERNS; ++j) {
            if (j < h) {
              first = first && !(j < p && oldv[j] == o);
              stays = stays || newv[j] == o;
              nfirst = nfirst && !(j < p && newv[j] == nw);
              was = was || oldv[j] == nw;
            }
          }
          if (first && !stays) {
            ok = D.add(t, o, -1) && ok;
          }
          if (nfirst && !was) {
            ok = D.add(t, nw, 1) && ok;
          }
        }
      }
    }
    if (!ok) {
      s_over = 1;
    }
    if (count_dh) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        dh += __shfl_xor_sync(0xffffffffu, dh, d);
      }
      if ((threadIdx.x & 31) == 0 && dh) {
        atomicAdd(&s_dh, dh);
      }
    }
  };
  // every (tile, id) whose count moved: new count -> arg-max candidates and smoothing inputs
  auto pass_b = [&](const GrbDelta& D) {
    for (uint32_t i = threadIdx.x; i <= D.mask; i += BS) {
      const unsigned long long key = D.keys[i];
      const int32_t dv = D.vals[i];
      if (key == GRB_DK_EMPTY || dv == 0) {
        continue;
      }
      const uint32_t t = (uint32_t)(key >> 32), id = (uint32_t)key;
      const uint32_t a = grb2_vote_get(b3.vk + (uint64_t)(bt0 + t) * vts, b3.vc + (uint64_t)(bt0 + t) * vts,
                                       vmask, id);
      const uint32_t c = a + (uint32_t)dv;
      // the smoothing passes read a candidate count c only as c > 2 and c > threshold
      // (goldrush_path.cpp:616,628-682)
      if (((a > 2) != (c > 2) || (a > thr_hi) != (c > thr_hi)) && um.lookup(id) != 0xFFFFFFFFu) {
        s_changed = 1;
      }
      if (id == sm.best_id[t] && c < sm.best_cnt[t]) {
        sm.rescan[t] = 1; // the arg-max lost votes: any id of the tile may lead now
        s_any_rescan = 1;
      } else if (c) {
        atomicMax(&sm.nbest[t], grb2_pack_best(c, id));
      }
    }
  };
  // full arg-max of tile t over base counts + deltas
  auto rescan_tile = [&](uint32_t t, const GrbDelta& D) {
    if (threadIdx.x == 0) {
      sm.nbest[t] = 0;
    }
    __syncthreads();
    const uint32_t* vk = b3.vk + (uint64_t)(bt0 + t) * vts;
    const uint32_t* vc = b3.vc + (uint64_t)(bt0 + t) * vts;
    unsigned long long best = 0;
    for (uint32_t i = threadIdx.x; i < vts; i += BS) {
      const uint32_t id = __ldg(&vk[i]);
      if (id) {
        const uint32_t c = __ldg(&vc[i]) + (uint32_t)D.get(t, id);
        const unsigned long long key = grb2_pack_best(c, id);
        best = key > best ? key : best;
      }
    }
    for (uint32_t i = threadIdx.x; i <= D.mask; i += BS) {
      const unsigned long long key = D.keys[i];
      if (key == GRB_DK_EMPTY || (uint32_t)(key >> 32) != t || D.vals[i] <= 0) {
        continue;
      }
      const uint32_t id = (uint32_t)key;
      const uint32_t c = grb2_vote_get(vk, vc, vmask, id) + (uint32_t)D.vals[i];
      const unsigned long long pk = grb2_pack_best(c, id);
      best = pk > best ? pk : best;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
      best = o > best ? o : best;
    }
    if ((threadIdx.x & 31) == 0 && best) {
      atomicMax(&sm.nbest[t], best);
    }
    __syncthreads();
  };

  // ---- deltas -> new arg-max per tile ----
  clear(d_light);
  __syncthreads();
  pass_a(ALL, d_light, true);
  __syncthreads();
  const bool heavy = s_over != 0;
  if (!heavy) {
    pass_b(d_light);
    __syncthreads();
    if (s_any_rescan) {
      // full arg-max of every flagged tile over base counts + deltas, all flagged tiles at once:
      // the vote tables of the flagged tiles are walked as one flat range by the whole CTA, four
      // independent loads in flight per thread (every dependent round trip to L2 costs ~700
      // cycles: one tile after the other with two barriers each was what reads overlapping an
      // inserted read spent 36 000 cycles on), then the ids only the delta table knows are added
      __shared__ uint32_t s_nflag;
      if (threadIdx.x == 0) {
        uint32_t nf = 0;
        for (uint32_t t = 0; t < n; ++t) {
          if (sm.rescan[t]) {
            sm.nbest[t] = 0;
            sm.root[nf++] = t;
          }
        }
        s_nflag = nf;
      }
      __syncthreads();
      const uint32_t total = s_nflag * vts; // vts is a power of two >= 32: a warp never straddles tiles
      for (uint32_t base = threadIdx.x; base < total; base += 4 * BS) {
        uint32_t ids[4], cs[4], ts[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t idx = base + u * BS;
          ids[u] = 0;
          cs[u] = 0;
          ts[u] = 0;
          if (idx < total) {
            ts[u] = sm.root[idx / vts];
            const uint64_t at = (uint64_t)(bt0 + ts[u]) * vts + (idx & vmask);
            ids[u] = __ldg(&b3.vk[at]);
            cs[u] = __ldg(&b3.vc[at]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          unsigned long long best = 0;
          if (ids[u]) {
            best = grb2_pack_best(cs[u] + (uint32_t)d_light.get(ts[u], ids[u]), ids[u]);
          }
          if (__any_sync(0xffffffffu, base + u * BS < total)) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
              const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
              best = o > best ? o : best;
            }
            if ((threadIdx.x & 31) == 0 && best) {
              atomicMax(&sm.nbest[ts[u]], best);
            }
          }
        }
      }
      for (uint32_t i = threadIdx.x; i <= d_light.mask; i += BS) {
        const unsigned long long key = d_light.keys[i];
        if (key == GRB_DK_EMPTY || d_light.vals[i] <= 0) {
          continue;
        }
        const uint32_t t = (uint32_t)(key >> 32), id = (uint32_t)key;
        if (!sm.rescan[t]) {
          continue;
        }
        const uint32_t c = grb2_vote_get(b3.vk + (uint64_t)(bt0 + t) * vts, b3.vc + (uint64_t)(bt0 + t) * vts,
                                         vmask, id) + (uint32_t)d_light.vals[i];
        const unsigned long long pk = grb2_pack_best(c, id);
        if (pk) {
          atomicMax(&sm.nbest[t], pk);
        }
      }
      __syncthreads();
    }
  } else {
    // too many distinct (tile, id) pairs for shared memory: one tile at a time in global scratch
    for (uint32_t t = 0; t < n; ++t) {
      __syncthreads();
      if (threadIdx.x == 0) {
        s_over = 0;
      }
      clear(d_heavy);
      __syncthreads();
      pass_a(t, d_heavy, false);
      __syncthreads();
      pass_b(d_heavy);
      __syncthreads();
      if (sm.rescan[t]) {
        rescan_tile(t, d_heavy);
      }
    }
  }
  __syncthreads();
  // new arg-max per tile; did an input of the smoothing passes change?  They read the arg-max
  // count only as > max(2, threshold) (goldrush_path.cpp:616,630).
  for (uint32_t i = threadIdx.x; i < n; i += BS) {
    const unsigned long long nb = sm.nbest[i];
    const uint32_t nc = (uint32_t)(nb >> 32);
    const uint32_t nid = nb ? 0xFFFFFFFFu - (uint32_t)(nb & 0xFFFFFFFFu) : 0u;
    const uint32_t oc = sm.best_cnt[i], oid = sm.best_id[i];
    if (nid != oid || (nc > 2 && nc > prm.threshold) != (oc > 2 && oc > prm.threshold)) {
      s_changed = 1;
    }
    if (nid != oid && um.lookup(nid) == 0xFFFFFFFFu) {
      s_uchg = 1; // the matrix columns no longer cover every arg-max id
    }
    sm.best_id[i] = nid;
    sm.best_cnt[i] = nc;
  }
  __syncthreads();
  if (s_changed) {
    // ---- re-smoothing on the re-validated votes ----
    uint32_t nu2 = nu;
    if (s_uchg) {
      nu2 = grb2_build_uq<BS>(n, sm.best_id, sm.root, sm.ukeys, sm.uvals, us, sm.uq);
    }
    if (!heavy) {
      for (uint32_t idx = threadIdx.x; idx < n * nu2; idx += BS) {
        const uint32_t i = idx / nu2, u = idx - i * nu2;
        const uint32_t id = sm.uq[u];
        const uint32_t c = grb2_vote_get(b3.vk + (uint64_t)(bt0 + i) * vts,
                                         b3.vc + (uint64_t)(bt0 + i) * vts, vmask, id) +
                           (uint32_t)d_light.get(i, id);
        cmat[idx] = c > 2 ? c : 0u;
      }
    } else {
      for (uint32_t t = 0; t < n; ++t) {
        __syncthreads();
        clear(d_heavy);
        __syncthreads();
        pass_a(t, d_heavy, false);
        __syncthreads();
        for (uint32_t u = threadIdx.x; u < nu2; u += BS) {
          const uint32_t id = sm.uq[u];
          const uint32_t c = grb2_vote_get(b3.vk + (uint64_t)(bt0 + t) * vts,
                                           b3.vc + (uint64_t)(bt0 + t) * vts, vmask, id) +
                             (uint32_t)d_heavy.get(t, id);
          cmat[(uint64_t)t * nu2 + u] = c > 2 ? c : 0u;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const GrbMatrixVotes v{ sm.best_id, sm.best_cnt, cmat, GrbUMap{ sm.ukeys, sm.uvals, us - 1 }, nu2 };
      const uint32_t n_as = grb_smooth_tiles(n, v, prm.threshold, sm.tile_id, sm.tile_as, sm.snap);
      uint32_t rel = 0;
      GrbReadPlan plan;
      grb_plan_read(n, n_as, len, prm.tile_len, prm.block_size, prm.unassigned_min, prm.assigned_max,
                    sm.tile_id, sm.tile_as, &rel, &plan);
      grb3_st_plan(&b3.np[b], plan);
      __stcg(&b3.np_adv[b], rel);
      __stcg(&b3.np_nas[b], n_as);
    }
  } else if (threadIdx.x == 0) {
    grb3_st_plan(&b3.np[b], bd.sp_plan[b]);
    __stcg(&b3.np_adv[b], bd.sp_adv[b]);
    __stcg(&b3.np_nas[b], bd.sp_n_as[b]);
  }
  if (threadIdx.x == 0) {
    __stcg(&b3.np_dh[b], s_dh);
  }
  __syncthreads();
  return (heavy ? 1u : 0u) | (s_changed ? 2u : 0u) | (s_uchg ? 4u : 0u) | (s_any_rescan ? 8u : 0u);
}

// One persistent cooperative launch per batch.  dec_idx[b] = index of read b in `decisions`.
// Dynamic shared memory (GrbFixSmem): uint64 dk[dc] | int32 dv[dc] | uint64 nbest[n_cap] | uint32
// ukeys[us] uvals[us] best_id best_cnt root uq tile_id [n_cap each] snap[n_cap + 2] rescan[n_cap]
// | uint8 tile_as[n_cap] (padded to 16) | uint32 cmat[n_cap * n_cap] when cm_smem.
// prof[] (GrbSelState): 0 walk, 2 reads, 3 order scan, 4 final (SM cycles of CTA 0); 1 plans that
// differed from the assumed ones, 5 conflict frames, 6 reads, 7 iterations, 8 inserts, 9 batches.
template<int BS>
__global__ void __launch_bounds__(BS, 512 / BS)
k3_fix(GrbReadsDev reads, GrbSelParams prm, GrbBatchDev bd, GrbB3 b3, GrbSelState* __restrict__ state_g,
       grb_decision* __restrict__ decisions, const uint64_t* __restrict__ dec_idx, uint32_t us,
       uint32_t n_cap, uint32_t dc, uint32_t cm_smem)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GrbFixSmem sm;
  sm.dk = reinterpret_cast<unsigned long long*>(smem_raw);
  sm.dv = reinterpret_cast<int32_t*>(sm.dk + dc);
  sm.nbest = reinterpret_cast<unsigned long long*>(sm.dv + dc);
  sm.ukeys = reinterpret_cast<uint32_t*>(sm.nbest + n_cap);
  sm.uvals = sm.ukeys + us;
  sm.best_id = sm.uvals + us;
  sm.best_cnt = sm.best_id + n_cap;
  sm.root = sm.best_cnt + n_cap;
  sm.uq = sm.root + n_cap;
  sm.tile_id = sm.uq + n_cap;
  sm.snap = sm.tile_id + n_cap;
  sm.rescan = sm.snap + n_cap + 2;
  sm.tile_as = reinterpret_cast<uint8_t*>(sm.rescan + n_cap);
  sm.cmat = reinterpret_cast<uint32_t*>(sm.tile_as + ((n_cap + 15) / 16) * 16);
  // G and the final pass (CTA 0) overlay the delta table: plans new / old, id advances
  GrbReadPlan* g_np = reinterpret_cast<GrbReadPlan*>(smem_raw);
  GrbReadPlan* g_old = g_np + bd.nb;
  uint32_t* g_a = reinterpret_cast<uint32_t*>(g_old + bd.nb); // [nb] adv, then n_as
  int32_t* g_b = reinterpret_cast<int32_t*>(g_a + bd.nb);     // [nb] dh
  uint32_t* g_c = reinterpret_cast<uint32_t*>(g_b + bd.nb);   // [3 * nb] queries, hits, misses
  __shared__ GrbSelState st;
  __shared__ uint32_t s_final_upto, s_n_commit, s_first_ins, s_mismatch, s_rolled, s_ids_end;

  const uint32_t cta = blockIdx.x, n_cta = gridDim.x;
  const uint32_t nb = bd.nb;
  const uint32_t B = (uint32_t)prm.block_size;
  if (threadIdx.x == 0) {
    st = *state_g;
  }
  __syncthreads();
  if (st.halt) {
    return;
  }
  unsigned long long phase = 0;
  long long t0 = clock64(), t1;
#define GRB_TICK(slot)                                                                             \
  t1 = clock64();                                                                                  \
  if (threadIdx.x == 0) {                                                                          \
    st.prof[slot] += (unsigned long long)(t1 - t0);                                                \
  }                                                                                                \
  t0 = t1;

  // G: read order.  iter 0 turns the speculative plans into the first assumption.
  auto order_scan = [&](uint32_t iter, uint32_t final_upto) {
    for (uint32_t b = threadIdx.x; b < nb; b += BS) {
      if (iter == 0) {
        g_np[b] = bd.sp_plan[b];
        g_a[b] = bd.sp_adv[b];
        grb3_st_plan(&b3.np[b], g_np[b]);
        __stcg(&b3.np_adv[b], g_a[b]);
        __stcg(&b3.np_nas[b], bd.sp_n_as[b]);
        __stcg(&b3.np_dh[b], 0);
      } else {
        g_np[b] = grb3_ld_plan(&b3.np[b]);
        g_a[b] = __ldcg(&b3.np_adv[b]);
        g_old[b] = grb3_ld_plan(&b3.plan_out[b]);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t ids = st.ids_inserted;
      uint64_t bases = st.cur.inserted_bases;
      uint32_t mismatch = 0xFFFFFFFFu, n_commit = nb, first_ins = 0xFFFFFFFFu, rolled = 0;
      for (uint32_t b = 0; b < nb; ++b) {
        GrbReadPlan p = g_np[b];
        const bool ins = p.verdict == GRB_UNTRIMMED || p.verdict == GRB_TRIMMED;
        if (ins) {
          p.first_id += ids;
          ids += g_a[b];
          bases += p.out_bases;
          if (prm.silver && prm.target_bases < bases) { // silver_path_check, :167-186
            p.n_blocks = 0; // every ID and count is wiped right after this insert: skip it
            rolled = 1;
          }
        }
        if (iter != 0 && b >= final_upto && mismatch == 0xFFFFFFFFu && grb3_plan_differs(p, g_old[b])) {
          mismatch = b;
        }
        g_np[b] = p;
        if (first_ins == 0xFFFFFFFFu && grb3_inserts(p)) {
          first_ins = b;
        }
        if (rolled) {
          n_commit = b + 1;
          break;
        }
      }
      uint32_t fu = iter == 0 ? 1u : (mismatch == 0xFFFFFFFFu ? n_commit : mismatch + 1);
      if (fu > n_commit) {
        fu = n_commit;
      }
      s_final_upto = fu;
      s_n_commit = n_commit;
      s_first_ins = first_ins;
      s_mismatch = mismatch;
      s_rolled = rolled;
      s_ids_end = ids;
      if (mismatch != 0xFFFFFFFFu) {
        st.prof[1] += 1;
      }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nb; b += BS) {
      GrbReadPlan p = g_np[b];
      if (b >= s_n_commit) {
        p = GrbReadPlan{};
      }
      grb3_st_plan(&b3.plan_out[b], p);
    }
    if (threadIdx.x == 0) {
      GrbFixCtl c{ s_final_upto, s_n_commit, s_first_ins, iter + 1 };
      __stcg(reinterpret_cast<uint4*>(b3.ctl), *reinterpret_cast<const uint4*>(&c));
    }
    __syncthreads();
  };

  if (cta == 0) {
    order_scan(0, 0);
  }
  GRB_TICK(3)
  grb_grid_barrier(b3.barrier, ++phase * n_cta);
  uint32_t iters = 0;
  while (true) {
    const uint4 craw = __ldcg(reinterpret_cast<const uint4*>(b3.ctl));
    const GrbFixCtl ctl{ craw.x, craw.y, craw.z, craw.w };
    grb3_walk(b3, B, ctl.n_commit);
    GRB_TICK(0)
    grb_grid_barrier(b3.barrier, ++phase * n_cta);
    if (ctl.final_upto >= ctl.n_commit) {
      break;
    }
    ++iters;
    for (uint32_t b = ctl.final_upto + cta; b < nb; b += n_cta) {
      if (ctl.first_ins == 0xFFFFFFFFu || b <= ctl.first_ins || __ldcg(&b3.fl_n[b]) == 0) {
        // no earlier read inserts, or nothing of this read is shared: the speculative plan holds
        if (threadIdx.x == 0) {
          grb3_st_plan(&b3.np[b], bd.sp_plan[b]);
          __stcg(&b3.np_adv[b], bd.sp_adv[b]);
          __stcg(&b3.np_nas[b], bd.sp_n_as[b]);
          __stcg(&b3.np_dh[b], 0);
        }
        continue;
      }
      const long long c0 = clock64();
      const uint32_t fl = grb3_read<BS>(reads, prm, bd, b3, sm, b, us, dc, cm_smem, n_cap);
      if (b3.dbg && threadIdx.x == 0) {
        const uint32_t at = atomicAdd(&b3.dbg[0], 1u);
        if (at < b3.dbg_cap) {
          uint32_t* rec = b3.dbg + 4 + 4 * (size_t)at;
          rec[0] = b | (iters << 12) | (fl << 16) | ((uint32_t)st.prof[9] << 20);
          rec[1] = __ldcg(&b3.fl_n[b]);
          rec[2] = (uint32_t)(clock64() - c0);
          rec[3] = cta;
        }
      }
    }
    GRB_TICK(2)
    grb_grid_barrier(b3.barrier, ++phase * n_cta);
    if (cta == 0) {
      order_scan(ctl.iter, ctl.final_upto);
    }
    GRB_TICK(3)
    grb_grid_barrier(b3.barrier, ++phase * n_cta);
  }
  if (cta != 0) {
    return;
  }
  // ---- final: decisions and loop state of reads [0, n_commit), in order ----
  const uint4 craw = __ldcg(reinterpret_cast<const uint4*>(b3.ctl));
  const uint32_t n_commit = craw.y;
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < n_commit; b += BS) {
    g_np[b] = grb3_ld_plan(&b3.plan_out[b]);
    g_a[b] = __ldcg(&b3.np_nas[b]);
    g_b[b] = __ldcg(&b3.np_dh[b]);
    g_c[3 * b] = bd.rd_queries[b];
    g_c[3 * b + 1] = bd.rd_hits[b];
    g_c[3 * b + 2] = bd.rd_miss[b];
  }
  __syncthreads();
  const uint32_t path = (uint32_t)st.curr_path;
  for (uint32_t b = threadIdx.x; b < n_commit; b += BS) {
    const GrbReadPlan p = g_np[b];
    grb_decision d;
    d.verdict = p.verdict;
    d.pad[0] = d.pad[1] = d.pad[2] = 0;
    d.path = path;
    d.trim_start = p.trim_start;
    d.trim_end = p.trim_end;
    d.num_tiles = bd.tile_first[b + 1] - bd.tile_first[b];
    d.num_assigned = g_a[b];
    decisions[dec_idx[b]] = d;
  }
  // The counters of the committed reads are sums; the one order-dependent event, the path rollover
  // (silver_path_check, goldrush_path.cpp:167-186), can only happen at the last committed read: the
  // order scan cut the batch there (s_rolled).
  __shared__ unsigned long long s_acc[9];
  if (threadIdx.x < 9) {
    s_acc[threadIdx.x] = 0;
  }
  __syncthreads();
  {
    unsigned long long a[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    for (uint32_t b = threadIdx.x; b < n_commit; b += BS) {
      const GrbReadPlan p = g_np[b];
      const uint32_t n = bd.tile_first[b + 1] - bd.tile_first[b];
      const uint32_t n_as = g_a[b];
      const int dh = g_b[b];
      a[0] += g_c[3 * b];
      a[1] += (unsigned long long)((long long)g_c[3 * b + 1] + dh);
      a[2] += (unsigned long long)((long long)g_c[3 * b + 2] - dh);
      a[3] += n;
      a[4] += n_as;
      a[5] += __ldcg(&b3.fl_n[b]);
      if (p.verdict == GRB_UNTRIMMED || p.verdict == GRB_TRIMMED) {
        a[6] += p.out_bases;
        a[7] += 1;
        a[8] += p.n_blocks != 0 ? 1 : 0;
      }
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        a[q] += __shfl_xor_sync(0xffffffffu, a[q], d);
      }
      if ((threadIdx.x & 31) == 0 && a[q]) {
        atomicAdd(&s_acc[q], a[q]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    GrbSelState& s = st;
    s.cur.queries += s_acc[0];
    s.cur.hits += s_acc[1];
    s.cur.misses += s_acc[2];
    s.cur.total_tiles += s_acc[3];
    s.cur.assigned_tiles += s_acc[4];
    s.cur.unassigned_tiles += s_acc[3] - s_acc[4];
    s.cur.inserted_bases += s_acc[6];
    s.cur.num_reads_in_path += s_acc[7];
    s.batch_inserts += (uint32_t)s_acc[7];
    s.prof[8] += s_acc[8];
    s.cur.valid_reads += n_commit;
    if (n_commit && prm.silver && prm.target_bases < s.cur.inserted_bases) {
      const uint32_t b = n_commit - 1; // the read whose insertion closed the path
      const uint64_t read_idx = bd.read_idx[b];
      s.snap = s.cur;
      s.snap.valid_reads -= 1; // the reference counts the closing read after the snapshot (:1089)
      s.snap.rollover_read = read_idx;
      s.n_snap = 1;
      s.curr_path += 1;
      s.halt = 1;
      s.halt_read = read_idx;
      if (prm.max_paths < s.curr_path) {
        s.finished = 1;
        s.cur.valid_reads -= 1; // exit(0) inside silver_path_check: the closing read is never counted
      } else {
        s.cur.inserted_bases = 0;
        s.cur.num_reads_in_path = 0;
        s.cur.phred_sum_in_path = 0;
      }
    }
    s.ids_inserted = (s.halt == 1 && !s.finished) ? 0u : s_ids_end;
    s.prof[5] += s_acc[5];
    s.prof[6] += n_commit;
    s.prof[7] += iters;
    s.prof[9] += 1;
  }
  __syncthreads();
  GRB_TICK(4)
#undef GRB_TICK
  if (threadIdx.x == 0) {
    *state_g = st;
  }
}

// After the fixed point: reservoir insert of the private ranks of every inserted read (any order:
// no other probe of the batch touches them), then the shared ranks go back to their ID slots.
__global__ void __launch_bounds__(256)
k3_bulk(GrbReadsDev reads, GrbFilterDev filt, GrbSelParams prm, GrbBatchDev bd, GrbB3 b3,
        const GrbSelState* __restrict__ state)
{
  if (state->halt == 1 && !state->finished) {
    return; // path rollover inside this batch: every ID and count is wiped next
  }
  if (b3.ctl->iter == 0) {
    return; // the batch never ran (an earlier batch of the chunk halted)
  }
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h;
  const uint32_t B = (uint32_t)prm.block_size;
  const uint32_t per_tile = prm.tile_frames * h;
  for (uint32_t bt = blockIdx.x; bt < bd.n_bt; bt += gridDim.x) {
    const uint32_t b = bd.tile_read[bt];
    const GrbReadPlan plan = b3.plan_out[b];
    if (!grb3_inserts(plan)) {
      continue;
    }
    const uint32_t t = bt - bd.tile_first[b];
    if (t < plan.trim_start || t > plan.trim_end) {
      continue;
    }
    const uint32_t id = plan.first_id + (t - plan.trim_start) / B + plan.id_bump;
    const uint32_t tl = grb_tile_bases(reads.len[bd.read_idx[b]], t, T, prm.kmer);
    const uint64_t* stash = bd.stash + (uint64_t)bt * per_tile;
    for (uint32_t rem = threadIdx.x; rem < per_tile; rem += blockDim.x) {
      const uint32_t f = rem / h, p = rem - f * h;
      if (tl < k + p || f >= tl - (k + p) + 1) {
        continue;
      }
      const uint64_t raw = __ldcs(&stash[rem]);
      if (raw & GRB_ST_SHARED) {
        continue;
      }
      const uint64_t rank = raw & ~GRB_STASH_NOFRAME;
      uint2* slot = reinterpret_cast<uint2*>(&filt.slots[rank]);
      uint2 s = __ldcg(slot);
      grb2_reservoir(rank, id, s.x, s.y);
      *slot = s;
    }
  }
  const uint32_t n_shared = b3.counters[1];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_shared; i += gridDim.x * blockDim.x) {
    const GrbShared3 e = b3.shared[i];
    *reinterpret_cast<uint2*>(&filt.slots[e.rank]) = make_uint2(e.id, e.count);
  }
}
