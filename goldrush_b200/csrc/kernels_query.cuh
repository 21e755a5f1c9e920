// K2 + K3 of the batch engine: the speculative query of a whole batch of reads.
//
//   k2_query    CTA per (read, tile): hash, probe, vote -> per-tile hash table of (id, count) in
//               global memory, arg-max, rank stash                              [whole GPU]
//   k2_cmat     CTA per read: distinct arg-max ids, their count matrix, smoothing + plan on the
//               speculative votes                                               [whole GPU]
// Replaces read_hashing.cpp:43-55 + goldrush_path.cpp:544-626 (query) and :628-889 (smoothing on
// the votes of the batch start; the ordered commit in kernels_commit.cuh re-validates them).
#pragma once
#include "batch_common.cuh"

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
    const uint32_t tl = grb_tile_bases(len, t, T, prm.kmer);
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
    uint64_t* stash = bd.stash + (uint64_t)bt * prm.tile_frames * h;
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
        my_q += grb_tile_bases(len, i, prm.tile_len, prm.kmer) - prm.k + 1;
      }
      if (my_q) {
        atomicAdd(&bd.rd_hits[b], my_h);
        atomicAdd(&bd.rd_miss[b], my_m);
        atomicAdd(&bd.rd_queries[b], my_q);
      }
    }
  }
}
