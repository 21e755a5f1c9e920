// Batch engine of the pass-2 loop, second generation: the ordered commit touches only what two
// reads of one batch can possibly share.
//
// Replaces the same reference code as kernels_select.cuh (read_hashing.cpp:7-75,
// goldrush_path.cpp:529-890 calc_num_assigned_tiles, :892-1094 process_read, :156-187
// silver_path_check, MIBFConstructSupport.hpp:247-283 insertMIBF) and produces the same decisions:
// every query sees every earlier insert (goldrush_path.cpp:1229-1256), exactly.
//
// A rank (ID slot) probed by ONE valid probe of the whole batch is *private*: no other read of the
// batch can see what an insert does to it, so its reservoir update (MIBFConstructSupport.hpp:
// 271-282) commutes with everything else in the batch and is applied after the ordered commit, by
// the whole GPU (k2_bulk).  A rank probed twice or more is *shared*: its {id, count} pair moves to
// a compact per-batch table (GrbShared, L2 resident), the frames that probe it are listed per read,
// and only those frames are re-validated, in read order, by the commit kernel.
//
//   k2_query    CTA per (read, tile): hash, probe, vote -> per-tile hash table of (id, count) in
//               global memory, arg-max, rank stash                              [whole GPU]
//   k2_cmat     CTA per read: distinct arg-max ids, their count matrix, smoothing + plan on the
//               speculative votes                                               [whole GPU]
//   k2_index    thread per valid probe: rank -> first probe (open addressing, one 64-bit CAS);
//               second and later probes of a rank go to the conflict list       [whole GPU]
//   k2_conf     thread per conflict: member chain of its shared entry, stash mark, frame list
//   k2_conf2    per-read lists for the commit: one entry per (shared rank, read) and one record per
//               conflict frame holding its ids at batch start / shared-entry references
//   k2_commit   ONE CTA, reads in file order: re-validate the read's conflict frames against the
//               shared table (vote deltas, arg-max, smoothing inputs), re-smooth if an input
//               changed, decide, reservoir-insert into the read's shared entries
//   k2_bulk     CTA per tile of every inserted read: reservoir insert of the private ranks;
//               then the shared entries are written back to the ID slots        [whole GPU]
#pragma once
#include "common.cuh"
#include "decide.cuh"
#include "kernels_select.cuh"
#include "kernels_batch.cuh"

#define GRB_ST_SHARED (1ull << 62) // stash entry: low 32 bits = shared-table index, not a rank
#define GRB_IX_EMPTY 0xFFFFFFFFFFFFFFFFull
#define GRB_IX_FLAG 1ull
#define GRB_NIL 0xFFFFFFFFu
#define GRB_FR_DEAD 0xFFFFFFFFu // frame record: the frame failed the bit test and never votes

struct __align__(16) GrbShared
{
  uint64_t rank;
  uint32_t id0;   // raw slot id when the batch started
  uint32_t id;    // raw current id
  uint32_t count; // current count
  uint32_t head;  // member chain (conflict-list index), GRB_NIL terminated
  uint32_t pad[2];
};

struct GrbB2
{
  uint32_t* vk; // [n_bt * table_size] vote-table ids (0 = empty)
  uint32_t* vc; // [n_bt * table_size] vote-table counts
  uint32_t table_size;
  uint32_t pad;
  unsigned long long* ix_tab; // batch index: rank << 27 | probe << 1 | shared flag
  uint64_t ix_mask;
  uint32_t* ix_sidx;  // per index slot: shared-table index once flagged
  uint32_t* counters; // [0] conflicts, [1] shared entries
  uint32_t* c_slot;   // per conflict: index slot, probe (stash index), next member, shared index
  uint32_t* c_probe;
  uint32_t* c_next;
  uint32_t* c_sidx;
  GrbShared* shared;
  uint32_t* fbits; // [n_bt * T bits] frame already listed
  uint32_t* fl_n;  // [nb] conflict frames of read b
  uint32_t* fl;    // frame list, read b's region starts at tile_first[b] * T: tile << 20 | frame
  uint32_t* fr;    // frame records, (2 + h) words each, same indexing as fl
  uint32_t* rl_n;  // [nb] shared ranks of read b
  uint2* rl;       // read b's region starts at tile_first[b] * T * h: {shared index, tile | multi << 31}
  GrbReadPlan* plan_out; // [nb] committed plans with absolute ids (verdict 0 = not committed)
};

__device__ __forceinline__ uint64_t
grb_ix_pack(uint64_t rank, uint32_t probe)
{
  return (rank << 27) | ((uint64_t)probe << 1);
}

// ---------------------------------------------------------------------------------------------
// per-tile vote tables in global memory
// ---------------------------------------------------------------------------------------------
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

// count[id] += delta; returns the count before, or 0xFFFFFFFF if the table is full
__device__ __forceinline__ uint32_t
grb2_vote_add(uint32_t* vk, uint32_t* vc, uint32_t mask, uint32_t id, uint32_t delta)
{
  uint32_t slot = grb_mix32(id) & mask;
  for (uint32_t tries = 0; tries <= mask; ++tries) {
    const uint32_t old = atomicCAS(&vk[slot], 0u, id);
    if (old == 0u || old == id) {
      return atomicAdd(&vc[slot], delta);
    }
    slot = (slot + 1) & mask;
  }
  return 0xFFFFFFFFu;
}

// One CTA per batch tile (grid-strided).  Dynamic shared memory:
//   ulonglong2 gL[ng * 256] | gR[ng * 256] | uint64 sw[sw_words] | uint32 keys[table_size] |
//   uint32 counts[table_size]
// Hashing goes through the grouped half-hash tables (nthash.cuh): ceil(half / 4) 16-byte reads per
// half hash instead of one read and one base extraction per care position, which brings the kernel
// under 64 registers so that two CTAs share an SM and one CTA's table set-up / write-out overlaps
// the other's probes.
template<int BS>
__global__ void __launch_bounds__(BS, 2)
k2_query(GrbReadsDev reads, const GrbSeedTables* __restrict__ seeds_g,
         const ulonglong2* __restrict__ gtab, uint32_t ng, GrbFilterDev filt,
         GrbSelParams prm, GrbBatchDev bd, GrbB2 b2, const GrbSelState* __restrict__ state,
         uint32_t bt_lo, uint32_t bt_hi)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ulonglong2* gL = reinterpret_cast<ulonglong2*>(smem_raw);
  ulonglong2* gR = gL + ng * 256;
  uint64_t* sw = reinterpret_cast<uint64_t*>(gR + ng * 256);
  uint32_t* keys = reinterpret_cast<uint32_t*>(sw + prm.sw_words);
  uint32_t* counts = keys + prm.table_size;
  __shared__ uint32_t s_hits, s_miss;
  __shared__ unsigned long long s_best;

  if (state->halt) {
    return;
  }
  for (unsigned i = threadIdx.x; i < 2 * ng * 256; i += BS) {
    gL[i] = gtab[i];
  }
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h, half = seeds_g->half;
  const uint32_t tmask = prm.table_size - 1;

  // tiles [bt_lo, bt_hi) of the batch: the whole batch on one GPU, this rank's share on several
  // (the per-tile outputs are all-gathered afterwards, comm.cuh)
  for (uint32_t bt = bt_lo + blockIdx.x; bt < bt_hi; bt += gridDim.x) {
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
      ulonglong2 lh = make_ulonglong2(0, 0); // left halves { fl, rl } at position p_left
      uint32_t p_left = 0xFFFFFFFFu;
#pragma unroll
      for (unsigned i = 0; i < GRB_MAX_PATTERNS; ++i) {
        if (i < h) {
          const uint32_t n_i = tl - (k + i) + 1; // valid positions of pattern i in this tile
          const uint32_t p = f < n_i ? f : n_i - 1; // stale tail keeps the last value
          const uint64_t at = (uint64_t)(p0 & 31) + p;
          if (p != p_left) {
            lh = grb_group_half(gL, ng, grb_lo64([&](uint64_t wi) { return sw[wi]; }, at));
            p_left = p;
          }
          const ulonglong2 rh =
            grb_group_half(gR, ng, grb_lo64([&](uint64_t wi) { return sw[wi]; }, at + half + i));
          const uint64_t hv = grb_combine(i, lh.x, lh.y, rh.x, rh.y);
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
    // arg-max (ties -> smallest id, goldrush_path.cpp:610-615) and the table itself to global
    unsigned long long best = 0;
    uint32_t* gk = b2.vk + (uint64_t)bt * prm.table_size;
    uint32_t* gc = b2.vc + (uint64_t)bt * prm.table_size;
    for (unsigned i = threadIdx.x; i < prm.table_size; i += BS) {
      const uint32_t c = counts[i];
      const uint32_t id = keys[i];
      gk[i] = id;
      gc[i] = c;
      if (c) {
        const unsigned long long key = ((unsigned long long)c << 32) | (0xFFFFFFFFu - id);
        best = key > best ? key : best;
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
    if (threadIdx.x == 0) {
      const unsigned long long bb = s_best;
      bd.best_count[bt] = (uint32_t)(bb >> 32);
      bd.best_id[bt] = bb ? 0xFFFFFFFFu - (uint32_t)(bb & 0xFFFFFFFFu) : 0u;
      bd.tile_hits[bt] = s_hits;
      bd.tile_miss[bt] = s_miss;
    }
  }
}

// Distinct arg-max ids of n tiles (best[0..n)) into uq_out / the id -> column map, by the whole
// CTA.  root[n], ukeys/uvals[us] are shared scratch.  Returns nu to all threads.
template<int BS>
__device__ __forceinline__ uint32_t
grb2_build_uq(uint32_t n, const uint32_t* best, uint32_t* root, uint32_t* ukeys, uint32_t* uvals,
              uint32_t us, uint32_t* uq_out)
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
  return s_nu;
}

// cmat[i * nu + u] = votes of uq[u] in tile i if > 2 (the reference's candidate list holds ids
// with count > 2, goldrush_path.cpp:616), else 0 -- from the tiles' vote tables, whole CTA
template<int BS>
__device__ __forceinline__ void
grb2_fill_cmat(const GrbB2& b2, uint32_t bt0, uint32_t n, uint32_t nu, const uint32_t* uq,
               uint32_t* cmat)
{
  const uint32_t ts = b2.table_size;
  for (uint32_t idx = threadIdx.x; idx < n * nu; idx += BS) {
    const uint32_t i = idx / nu, u = idx - i * nu;
    const uint32_t c = grb2_vote_get(b2.vk + (uint64_t)(bt0 + i) * ts, b2.vc + (uint64_t)(bt0 + i) * ts,
                                     ts - 1, uq[u]);
    cmat[idx] = c > 2 ? c : 0u;
  }
}

// After the speculative query: one CTA per read of the batch builds the read's count matrix and
// runs the smoothing passes + the plan on the speculative votes.  Dynamic shared memory:
//   uint32 best[n_cap] bcnt[n_cap] root[n_cap] tile_id[n_cap] snap[n_cap+2] uq[n_cap] ukeys[us]
//   uvals[us] | uint8 tile_as[n_cap] (padded to 16) | uint32 cmat[n_cap * n_cap] when cm_smem
template<int BS>
__global__ void __launch_bounds__(BS)
k2_cmat(GrbReadsDev reads, GrbSelParams prm, GrbBatchDev bd, GrbB2 b2,
        const GrbSelState* __restrict__ state, uint32_t n_cap, uint32_t us, uint32_t cm_smem)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_best = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* s_bcnt = s_best + n_cap;
  uint32_t* s_root = s_bcnt + n_cap;
  uint32_t* s_tile_id = s_root + n_cap;
  uint32_t* s_snap = s_tile_id + n_cap;
  uint32_t* s_uq = s_snap + n_cap + 2;
  uint32_t* s_ukeys = s_uq + n_cap;
  uint32_t* s_uvals = s_ukeys + us;
  uint8_t* s_tile_as = reinterpret_cast<uint8_t*>(s_uvals + us);
  uint32_t* s_cmat = reinterpret_cast<uint32_t*>(s_tile_as + ((n_cap + 15) / 16) * 16);
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
    const uint32_t nu = grb2_build_uq<BS>(n, s_best, s_root, s_ukeys, s_uvals, us, s_uq);
    uint32_t* cmat = cm_smem ? s_cmat : bd.cm + bd.cm_off[b];
    grb2_fill_cmat<BS>(b2, bt0, n, nu, s_uq, cmat);
    for (uint32_t u = threadIdx.x; u < nu; u += BS) {
      bd.uq[bt0 + u] = s_uq[u];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      bd.nu[b] = nu;
      const GrbMatrixVotes v{ s_best, s_bcnt, cmat, GrbUMap{ s_ukeys, s_uvals, us - 1 }, nu };
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
  const uint32_t per_tile = T * h;
  a.bt = idx / per_tile;
  const uint32_t rem = idx - a.bt * per_tile;
  a.f = rem / h;
  a.p = rem - a.f * h;
  a.b = bd.tile_read[a.bt];
  a.t = a.bt - bd.tile_first[a.b];
  a.tl = grb_tile_bases(reads.len[bd.read_idx[a.b]], a.t, T, k);
  a.valid = a.tl >= k + a.p && a.f < a.tl - (k + a.p) + 1;
  return a;
}

// Batch index: every valid probe registers its rank; the first one owns the entry, every later one
// turns the rank into a shared rank and joins the conflict list (together with the owner).
__global__ void __launch_bounds__(256)
k2_index(GrbReadsDev reads, GrbFilterDev filt, GrbSelParams prm, GrbBatchDev bd, GrbB2 b2,
         const GrbSelState* __restrict__ state)
{
  if (state->halt) {
    return;
  }
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h;
  const uint32_t per_tile = T * h;
  for (uint32_t bt = blockIdx.x; bt < bd.n_bt; bt += gridDim.x) {
    const uint32_t b = bd.tile_read[bt];
    const uint32_t t = bt - bd.tile_first[b];
    const uint32_t tl = grb_tile_bases(reads.len[bd.read_idx[b]], t, T, k);
    for (uint32_t rem = threadIdx.x; rem < per_tile; rem += blockDim.x) {
      const uint32_t f = rem / h, p = rem - f * h;
      if (tl < k + p || f >= tl - (k + p) + 1) {
        continue;
      }
      const uint32_t idx = bt * per_tile + rem;
      const uint64_t rank = bd.stash[idx] & ~GRB_STASH_NOFRAME;
      const uint64_t mine = grb_ix_pack(rank, idx);
      uint64_t slot = grb_mix64(rank) & b2.ix_mask;
      while (true) {
        const unsigned long long old = atomicCAS(&b2.ix_tab[slot], GRB_IX_EMPTY, mine);
        if (old == GRB_IX_EMPTY) {
          break;
        }
        if ((old >> 27) == rank) {
          const unsigned long long prev = atomicOr(&b2.ix_tab[slot], GRB_IX_FLAG);
          if (!(prev & GRB_IX_FLAG)) { // first duplicate: open the shared entry, list the owner
            const uint32_t sidx = atomicAdd(&b2.counters[1], 1u);
            b2.ix_sidx[slot] = sidx;
            const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(&filt.slots[rank]));
            GrbShared e;
            e.rank = rank;
            e.id0 = raw.x;
            e.id = raw.x;
            e.count = raw.y;
            e.head = GRB_NIL;
            e.pad[0] = e.pad[1] = 0;
            b2.shared[sidx] = e;
            const uint32_t ci = atomicAdd(&b2.counters[0], 1u);
            b2.c_slot[ci] = (uint32_t)slot;
            b2.c_probe[ci] = (uint32_t)((prev >> 1) & 0x3FFFFFFu);
          }
          const uint32_t ci = atomicAdd(&b2.counters[0], 1u);
          b2.c_slot[ci] = (uint32_t)slot;
          b2.c_probe[ci] = idx;
          break;
        }
        slot = (slot + 1) & b2.ix_mask;
      }
    }
  }
}

// Per conflict: chain it to its shared entry, mark its stash entries (stale-tail repeats
// included) and list the frames it votes in, once each.
__global__ void __launch_bounds__(256)
k2_conf(GrbReadsDev reads, GrbSelParams prm, GrbBatchDev bd, GrbB2 b2,
        const GrbSelState* __restrict__ state)
{
  if (state->halt) {
    return;
  }
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h;
  const uint32_t n_conf = b2.counters[0];
  for (uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x; ci < n_conf; ci += gridDim.x * blockDim.x) {
    const uint32_t sidx = b2.ix_sidx[b2.c_slot[ci]];
    const uint32_t probe = b2.c_probe[ci];
    b2.c_sidx[ci] = sidx;
    b2.c_next[ci] = atomicExch(&b2.shared[sidx].head, ci);
    const GrbProbeAt a = grb2_probe_at(reads, prm, bd, probe);
    const uint32_t frames = a.tl - k + 1;
    const uint32_t n_p = a.tl - (k + a.p) + 1;
    const uint32_t f_hi = (a.f == n_p - 1) ? frames - 1 : a.f;
    for (uint32_t ff = a.f; ff <= f_hi; ++ff) {
      uint64_t* e = bd.stash + ((uint64_t)a.bt * T + ff) * h + a.p;
      *e = (*e & GRB_STASH_NOFRAME) | GRB_ST_SHARED | sidx;
      const uint32_t g = a.bt * T + ff;
      const uint32_t bit = 1u << (g & 31);
      if (!(atomicOr(&b2.fbits[g >> 5], bit) & bit)) {
        const uint32_t at = atomicAdd(&b2.fl_n[a.b], 1u);
        b2.fl[(uint64_t)bd.tile_first[a.b] * T + at] = (a.t << 20) | ff;
      }
    }
  }
}

// (1) per conflict: is it the read's leader for its shared rank (smallest tile, then smallest
// conflict index, among the members of the same read)?  The leader alone enters the read's list.
// (2) per listed frame: its record {tile << 20 | frame, mask of shared patterns, per pattern the
// normalised id at batch start or the shared-table index}.
__global__ void __launch_bounds__(256)
k2_conf2(GrbReadsDev reads, GrbFilterDev filt, GrbSelParams prm, GrbBatchDev bd, GrbB2 b2,
         const GrbSelState* __restrict__ state)
{
  if (state->halt) {
    return;
  }
  const uint32_t T = prm.tile_len, h = prm.h;
  const uint32_t per_tile = T * h;
  const uint32_t n_conf = b2.counters[0];
  for (uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x; ci < n_conf; ci += gridDim.x * blockDim.x) {
    const uint32_t sidx = b2.c_sidx[ci];
    const uint32_t bt = b2.c_probe[ci] / per_tile;
    const uint32_t b = bd.tile_read[bt];
    bool leader = true, multi = false;
    for (uint32_t m = b2.shared[sidx].head; m != GRB_NIL; m = b2.c_next[m]) {
      if (m == ci) {
        continue;
      }
      const uint32_t bt_m = b2.c_probe[m] / per_tile;
      if (bd.tile_read[bt_m] != b) {
        continue;
      }
      multi = true;
      if (bt_m < bt || (bt_m == bt && m < ci)) {
        leader = false;
      }
    }
    if (leader) {
      const uint32_t at = atomicAdd(&b2.rl_n[b], 1u);
      b2.rl[(uint64_t)bd.tile_first[b] * per_tile + at] =
        make_uint2(sidx, (bt - bd.tile_first[b]) | (multi ? 0x80000000u : 0u));
    }
  }
  const uint32_t stride = 2 + h;
  for (uint32_t b = blockIdx.x; b < bd.nb; b += gridDim.x) {
    const uint32_t nfr = b2.fl_n[b];
    const uint32_t bt0 = bd.tile_first[b];
    const uint64_t off = (uint64_t)bt0 * T;
    for (uint32_t i = threadIdx.x; i < nfr; i += blockDim.x) {
      const uint32_t tf = b2.fl[off + i];
      const uint32_t t = tf >> 20, f = tf & 0xFFFFFu;
      const uint64_t* e = bd.stash + ((uint64_t)(bt0 + t) * T + f) * h;
      uint32_t* rec = b2.fr + (off + i) * stride;
      uint32_t smask = 0;
      const bool dead = (e[0] & GRB_STASH_NOFRAME) != 0;
      for (uint32_t p = 0; p < h && !dead; ++p) {
        const uint64_t v = e[p];
        if (v & GRB_ST_SHARED) {
          smask |= 1u << p;
          rec[2 + p] = (uint32_t)v;
        } else {
          rec[2 + p] = grb_norm_id(__ldcg(&filt.slots[v & ~GRB_STASH_NOFRAME].id));
        }
      }
      rec[0] = tf;
      rec[1] = dead ? GRB_FR_DEAD : smask;
    }
  }
}

// Shared-memory carve-up of k2_commit:
//   unsigned long long nbest[n_cap] | uint32 ukeys[us] uvals[us] best_id[n_cap] best_cnt[n_cap]
//   root[n_cap] uq[n_cap] tile_id[n_cap] snap[n_cap + 2] rescan[n_cap]
//   | uint8 tile_as[n_cap] (padded to 16) | uint32 cmat[n_cap * n_cap] when cm_smem
struct GrbCommit2Smem
{
  unsigned long long* nbest;
  uint32_t* ukeys;
  uint32_t* uvals;
  uint32_t* best_id;
  uint32_t* best_cnt;
  uint32_t* root;
  uint32_t* uq;
  uint32_t* tile_id;
  uint32_t* snap;
  uint32_t* rescan;
  uint8_t* tile_as;
  uint32_t* cmat;
};

// old / new normalised ids of one conflict frame; returns false when nothing changed
__device__ __forceinline__ bool
grb2_frame_ids(const uint32_t* __restrict__ rec, const GrbShared* __restrict__ shared, uint32_t h,
               uint32_t smask, uint32_t* oldv, uint32_t* newv)
{
  bool any = false;
#pragma unroll
  for (unsigned p = 0; p < GRB_MAX_PATTERNS; ++p) {
    if (p < h) {
      const uint32_t r = __ldcs(&rec[2 + p]);
      if ((smask >> p) & 1u) {
        const uint4 e = __ldcg(reinterpret_cast<const uint4*>(&shared[r]) + 0); // rank, id0, id
        oldv[p] = grb_norm_id(e.z);
        newv[p] = grb_norm_id(e.w);
        any = any || oldv[p] != newv[p];
      } else {
        oldv[p] = r;
        newv[p] = r;
      }
    }
  }
  return any;
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
// prof[] as in GrbSelState: 0 check, 2 re-smoothing, 3 decide, 4 insert (SM cycles), 6..9 counts.
template<int BS>
__global__ void __launch_bounds__(BS, 1)
k2_commit(GrbReadsDev reads, GrbSelParams prm, GrbBatchDev bd, GrbB2 b2,
          GrbSelState* __restrict__ state_g, grb_decision* __restrict__ decisions,
          const uint64_t* __restrict__ dec_idx, uint32_t us, uint32_t n_cap, uint32_t cm_smem)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GrbCommit2Smem sm;
  sm.nbest = reinterpret_cast<unsigned long long*>(smem_raw);
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
  __shared__ GrbSelState st;
  __shared__ GrbReadPlan s_plan;
  __shared__ uint32_t s_changed, s_uchg, s_overflow, s_any_rescan, s_n_as, s_adv, s_redone;
  __shared__ int s_dh;
  __shared__ unsigned long long s_scan_best;

  if (threadIdx.x == 0) {
    st = *state_g;
  }
  __syncthreads();
  if (st.halt) {
    return;
  }
  const uint32_t T = prm.tile_len, h = prm.h;
  const uint32_t B = (uint32_t)prm.block_size;
  const uint32_t vts = b2.table_size, vmask = vts - 1;
  const uint32_t stride = 2 + h;
  const uint32_t thr_hi = prm.threshold > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)prm.threshold;

  for (uint32_t b = 0; b < bd.nb; ++b) {
    const uint64_t read_idx = bd.read_idx[b];
    const uint32_t len = reads.len[read_idx];
    const uint32_t n = len / T;
    const uint32_t bt0 = bd.tile_first[b];
    long long t0 = clock64(), t1;
#define GRB_TICK(slot)                                                                             \
  t1 = clock64();                                                                                  \
  if (threadIdx.x == 0) {                                                                          \
    st.prof[slot] += (unsigned long long)(t1 - t0);                                                \
  }                                                                                                \
  t0 = t1;
    const uint32_t nfr = st.batch_inserts != 0 ? b2.fl_n[b] : 0u;
    if (threadIdx.x == 0) {
      s_changed = 0;
      s_uchg = 0;
      s_overflow = 0;
      s_any_rescan = 0;
      s_dh = 0;
      s_redone = 0;
    }
    if (nfr) {
      // ---- check: re-validate the read's conflict frames against the shared table ----
      const uint32_t nu = bd.nu[b];
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
      __syncthreads();
      for (uint32_t u = threadIdx.x; u < nu; u += BS) {
        const uint32_t id = bd.uq[bt0 + u];
        sm.uq[u] = id;
        grb_umap_insert_par(sm.ukeys, sm.uvals, us - 1, id, u);
      }
      __syncthreads();
      const GrbUMap um{ sm.ukeys, sm.uvals, us - 1 };
      const uint32_t* fr = b2.fr + (uint64_t)bt0 * T * stride;
      // phase A: vote deltas of the frames whose ids changed
      int dh = 0;
      for (uint32_t i = threadIdx.x; i < nfr; i += BS) {
        const uint32_t* rec = fr + (uint64_t)i * stride;
        const uint32_t smask = __ldcs(&rec[1]);
        if (smask == GRB_FR_DEAD) {
          continue;
        }
        uint32_t oldv[GRB_MAX_PATTERNS], newv[GRB_MAX_PATTERNS];
        if (!grb2_frame_ids(rec, b2.shared, h, smask, oldv, newv)) {
          continue;
        }
        const uint32_t t = __ldcs(&rec[0]) >> 20;
        uint32_t* vk = b2.vk + (uint64_t)(bt0 + t) * vts;
        uint32_t* vc = b2.vc + (uint64_t)(bt0 + t) * vts;
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
#pragma unroll
            for (int dir = 0; dir < 2; ++dir) {
              const bool go = dir == 0 ? (first && !stays) : (nfirst && !was);
              if (!go) {
                continue;
              }
              const uint32_t id = dir == 0 ? o : nw;
              const uint32_t d = dir == 0 ? 0xFFFFFFFFu : 1u;
              const uint32_t a = grb2_vote_add(vk, vc, vmask, id, d);
              if (a == 0xFFFFFFFFu) {
                s_overflow = 1;
                continue;
              }
              // the smoothing passes read a candidate count c only as c > 2 and c > threshold
              // (goldrush_path.cpp:616,628-682); any step across either bound may change them
              const uint32_t c = a + d;
              if (((a > 2) != (c > 2) || (a > thr_hi) != (c > thr_hi)) && um.lookup(id) != 0xFFFFFFFFu) {
                s_changed = 1;
              }
            }
          }
        }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        dh += __shfl_xor_sync(0xffffffffu, dh, d);
      }
      if ((threadIdx.x & 31) == 0 && dh) {
        atomicAdd(&s_dh, dh);
      }
      __syncthreads();
      // phase B: final counts of every id a changed frame touched -> the tiles' new arg-max
      for (uint32_t i = threadIdx.x; i < nfr && !s_overflow; i += BS) {
        const uint32_t* rec = fr + (uint64_t)i * stride;
        const uint32_t smask = __ldcs(&rec[1]);
        if (smask == GRB_FR_DEAD) {
          continue;
        }
        uint32_t oldv[GRB_MAX_PATTERNS], newv[GRB_MAX_PATTERNS];
        if (!grb2_frame_ids(rec, b2.shared, h, smask, oldv, newv)) {
          continue;
        }
        const uint32_t t = __ldcs(&rec[0]) >> 20;
        const uint32_t* vk = b2.vk + (uint64_t)(bt0 + t) * vts;
        const uint32_t* vc = b2.vc + (uint64_t)(bt0 + t) * vts;
#pragma unroll
        for (unsigned p = 0; p < GRB_MAX_PATTERNS; ++p) {
          if (p < h) {
#pragma unroll
            for (int dir = 0; dir < 2; ++dir) {
              const uint32_t id = dir == 0 ? oldv[p] : newv[p];
              if (id == 0 || oldv[p] == newv[p]) {
                continue;
              }
              const uint32_t c = grb2_vote_get(vk, vc, vmask, id);
              if (id == sm.best_id[t] && c < sm.best_cnt[t]) {
                sm.rescan[t] = 1; // the arg-max lost votes: any id of the tile may lead now
                s_any_rescan = 1;
              } else if (c) {
                atomicMax(&sm.nbest[t], grb2_pack_best(c, id));
              }
            }
          }
        }
      }
      __syncthreads();
      if (s_any_rescan && !s_overflow) {
        for (uint32_t t = 0; t < n; ++t) {
          if (!sm.rescan[t]) {
            continue;
          }
          if (threadIdx.x == 0) {
            s_scan_best = 0;
          }
          __syncthreads();
          const uint32_t* vk = b2.vk + (uint64_t)(bt0 + t) * vts;
          const uint32_t* vc = b2.vc + (uint64_t)(bt0 + t) * vts;
          unsigned long long best = 0;
          for (uint32_t i = threadIdx.x; i < vts; i += BS) {
            const uint32_t c = __ldcg(&vc[i]);
            if (c) {
              const unsigned long long key = grb2_pack_best(c, __ldcg(&vk[i]));
              best = key > best ? key : best;
            }
          }
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
            best = o > best ? o : best;
          }
          if ((threadIdx.x & 31) == 0 && best) {
            atomicMax(&s_scan_best, best);
          }
          __syncthreads();
          if (threadIdx.x == 0) {
            sm.nbest[t] = s_scan_best;
          }
          __syncthreads();
        }
      }
      // new arg-max per tile; did an input of the smoothing passes change?  They read the
      // arg-max count only as > max(2, threshold) (goldrush_path.cpp:616,630).
      for (uint32_t i = threadIdx.x; i < n && !s_overflow; i += BS) {
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
      if (threadIdx.x == 0) {
        st.prof[9] += 1;
      }
      GRB_TICK(0)
      if (s_overflow) {
        // a vote table filled up (ids handed out inside the batch exceed its slack): stop before
        // this read; the host starts a new batch at it
        if (threadIdx.x == 0) {
          st.halt = 2;
          st.halt_read = read_idx;
        }
        __syncthreads();
        break;
      }
      if (s_changed) {
        // ---- re-smoothing on the re-validated votes ----
        uint32_t nu2 = nu;
        if (s_uchg) {
          nu2 = grb2_build_uq<BS>(n, sm.best_id, sm.root, sm.ukeys, sm.uvals, us, sm.uq);
        }
        uint32_t* cmat = cm_smem ? sm.cmat : bd.cmat;
        grb2_fill_cmat<BS>(b2, bt0, n, nu2, sm.uq, cmat);
        __syncthreads();
        if (threadIdx.x == 0) {
          const GrbMatrixVotes v{ sm.best_id, sm.best_cnt, cmat, GrbUMap{ sm.ukeys, sm.uvals, us - 1 },
                                  nu2 };
          const uint32_t n_as = grb_smooth_tiles(n, v, prm.threshold, sm.tile_id, sm.tile_as, sm.snap);
          uint32_t rel = 0;
          GrbReadPlan plan;
          grb_plan_read(n, n_as, len, prm.tile_len, prm.block_size, prm.unassigned_min,
                        prm.assigned_max, sm.tile_id, sm.tile_as, &rel, &plan);
          s_plan = plan;
          s_n_as = n_as;
          s_adv = rel;
          s_redone = 1;
          st.prof[7] += 1;
        }
        __syncthreads();
        GRB_TICK(2)
      }
    }
    // ---- decide ----
    if (threadIdx.x == 0) {
      GrbSelState& s = st;
      GrbReadPlan plan;
      uint32_t n_as, adv;
      s.prof[5] += nfr;
      if (s_redone) {
        plan = s_plan;
        n_as = s_n_as;
        adv = s_adv;
        const GrbReadPlan sp = bd.sp_plan[b];
        if (sp.verdict != plan.verdict || sp.trim_start != plan.trim_start ||
            sp.trim_end != plan.trim_end || bd.sp_adv[b] != adv) {
          s.prof[1] += 1; // the re-validated plan differs from the speculative one
        }
      } else {
        plan = bd.sp_plan[b];
        n_as = bd.sp_n_as[b];
        adv = bd.sp_adv[b];
      }
      s.cur.queries += bd.rd_queries[b];
      s.cur.hits += (uint64_t)((int64_t)bd.rd_hits[b] + s_dh);
      s.cur.misses += (uint64_t)((int64_t)bd.rd_miss[b] - s_dh);
      s.cur.total_tiles += n;
      s.cur.assigned_tiles += n_as;
      s.cur.unassigned_tiles += n - n_as;
      const bool ins = plan.verdict == GRB_UNTRIMMED || plan.verdict == GRB_TRIMMED;
      if (ins) {
        plan.first_id += s.ids_inserted;
        s.ids_inserted += adv;
      }
      grb_decision d;
      d.verdict = plan.verdict;
      d.pad[0] = d.pad[1] = d.pad[2] = 0;
      d.path = (uint32_t)s.curr_path;
      d.trim_start = plan.trim_start;
      d.trim_end = plan.trim_end;
      d.num_tiles = n;
      d.num_assigned = n_as;
      decisions[dec_idx[b]] = d;
      if (ins) {
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
      s.prof[8] += (ins && plan.n_blocks != 0) ? 1 : 0;
      s_plan = plan;
      b2.plan_out[b] = plan;
    }
    __syncthreads();
    GRB_TICK(3)
    if (st.halt) {
      break; // path rollover: the host resets the ID slots and resumes after this read
    }
    // ---- insert into the read's shared ranks, one thread per rank, insert calls in order ----
    const GrbReadPlan plan = s_plan;
    if (plan.n_blocks != 0 && (plan.verdict == GRB_UNTRIMMED || plan.verdict == GRB_TRIMMED)) {
      const uint32_t nrl = b2.rl_n[b];
      const uint2* rl = b2.rl + (uint64_t)bt0 * T * h;
      const uint32_t per_tile = T * h;
      for (uint32_t i = threadIdx.x; i < nrl; i += BS) {
        const uint2 ent = __ldcs(&rl[i]);
        const uint32_t sidx = ent.x, t = ent.y & 0x7FFFFFFFu;
        const bool multi = (ent.y >> 31) != 0;
        const bool in = t >= plan.trim_start && t <= plan.trim_end;
        if (!multi && !in) {
          continue;
        }
        GrbShared* e = &b2.shared[sidx];
        const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(e));
        const uint64_t rank = ((uint64_t)raw.y << 32) | raw.x;
        uint32_t cur_id = raw.w;
        uint32_t cur_count = __ldcg(&e->count);
        if (!multi) {
          grb2_reservoir(rank, plan.first_id + (t - plan.trim_start) / B + plan.id_bump, cur_id,
                         cur_count);
        } else {
          // several probes of this read share the rank: one reservoir step per distinct insert
          // call, in call order (a rank counts once per call, MIBFConstructSupport.hpp:255-270)
          const uint32_t head = __ldcg(&e->head);
          uint32_t last_j = 0xFFFFFFFFu;
          while (true) {
            uint32_t next_j = 0xFFFFFFFFu;
            for (uint32_t m = head; m != GRB_NIL; m = b2.c_next[m]) {
              const uint32_t bt_m = b2.c_probe[m] / per_tile;
              if (bd.tile_read[bt_m] != b) {
                continue;
              }
              const uint32_t tm = bt_m - bt0;
              if (tm < plan.trim_start || tm > plan.trim_end) {
                continue;
              }
              const uint32_t j = (tm - plan.trim_start) / B;
              if ((last_j == 0xFFFFFFFFu || j > last_j) && j < next_j) {
                next_j = j;
              }
            }
            if (next_j == 0xFFFFFFFFu) {
              break;
            }
            grb2_reservoir(rank, plan.first_id + next_j + plan.id_bump, cur_id, cur_count);
            last_j = next_j;
          }
        }
        e->id = cur_id;
        e->count = cur_count;
      }
    }
    __syncthreads();
    GRB_TICK(4)
  }
#undef GRB_TICK
  if (threadIdx.x == 0) {
    *state_g = st;
  }
}

// After the commit: reservoir insert of the private ranks of every inserted read (any order: no
// other probe of the batch touches them), then the shared entries go back to their ID slots.
__global__ void __launch_bounds__(256)
k2_bulk(GrbReadsDev reads, GrbFilterDev filt, GrbSelParams prm, GrbBatchDev bd, GrbB2 b2,
        const GrbSelState* __restrict__ state)
{
  if (state->halt == 1 && !state->finished) {
    return; // path rollover inside this batch: every ID and count is wiped next
  }
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h;
  const uint32_t B = (uint32_t)prm.block_size;
  const uint32_t per_tile = T * h;
  for (uint32_t bt = blockIdx.x; bt < bd.n_bt; bt += gridDim.x) {
    const uint32_t b = bd.tile_read[bt];
    const GrbReadPlan plan = b2.plan_out[b];
    if (plan.n_blocks == 0 || (plan.verdict != GRB_UNTRIMMED && plan.verdict != GRB_TRIMMED)) {
      continue;
    }
    const uint32_t t = bt - bd.tile_first[b];
    if (t < plan.trim_start || t > plan.trim_end) {
      continue;
    }
    const uint32_t id = plan.first_id + (t - plan.trim_start) / B + plan.id_bump;
    const uint32_t tl = grb_tile_bases(reads.len[bd.read_idx[b]], t, T, k);
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
  const uint32_t n_shared = b2.counters[1];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_shared; i += gridDim.x * blockDim.x) {
    const GrbShared e = b2.shared[i];
    uint2* slot = reinterpret_cast<uint2*>(&filt.slots[e.rank]);
    *slot = make_uint2(e.id, e.count);
  }
}
