// K2 + K3 (tile hashing, miBF probe, per-tile ID vote), the per-read decision, and K4c (ordered
// ID insertion) — the pass-2 loop of GoldRush-Path, one read at a time in file order.
//
// Replaces read_hashing (goldrush_path/read_hashing.cpp:7-75), calc_num_assigned_tiles
// (goldrush_path/goldrush_path.cpp:529-890), process_read (:892-1094), silver_path_check
// (:156-187) and MIBFConstructSupport::insertMIBF (MIBFConstructSupport.hpp:247-283).
//
// Per visited read the stream carries   k_query -> k_decide -> k_insert_collect -> k_insert_apply.
// The loop-carried dependence of the reference (every query sees every earlier insert) is kept by
// stream order; nothing is decided on the host.
#pragma once
#include "common.cuh"
#include "decide.cuh"

#define GRB_SAT_MASK 0x80000000u // MIBloomFilter::s_mask (MIBloomFilter.hpp:38)
#define GRB_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

struct GrbSelParams
{
  uint32_t tile_len, k, h, cand_cap; // k = span of seed pattern 0 (k - 1 for an odd -k, spaced_seeds.cpp:28,58)
  uint32_t table_size; // shared-memory vote table entries (power of two)
  uint32_t sw_words;   // shared-memory words holding one tile's bases
  int32_t silver;
  uint32_t kmer;        // -k: a tile is cut as substr(i * T, T + kmer - 1) (read_hashing.cpp:43-46)
  uint32_t tile_frames; // frames of a full tile = per-tile stride of the stash: T + kmer - k
  uint32_t pad;
  uint64_t threshold, unassigned_min, assigned_max, block_size, max_paths, target_bases;
};

struct GrbSelState
{
  uint32_t pad0;
  uint32_t batch_inserts; // reads of the current batch that inserted so far
  uint32_t ids_inserted;
  uint32_t halt;     // set at a path rollover: later kernels are no-ops until the host resumes
  uint32_t finished; // the reference would have called exit(0) (goldrush_path.cpp:174-176)
  uint32_t n_snap;
  uint64_t halt_read;
  uint64_t curr_path;
  grb_path_stats cur;
  grb_path_stats snap; // counters at the last rollover (log_path_stat, :126-154)
  // k3_fix phase clocks of CTA 0 (SM cycles) and event counts, for grb_commit_profile: 0 walk,
  // 2 re-validation of its reads, 3 wait for the slowest CTA + order scan, 4 final pass; 1 scans that
  // found a plan contradicted, 5 conflict frames, 6 reads, 7 iterations, 8 inserting reads, 9 batches
  unsigned long long prof[10];
};

struct GrbSelScratch
{
  uint64_t* stash;      // [tiles * tile_len * h] rank of every probe of the current read
  uint32_t* best_id;    // [tiles]
  uint32_t* best_count; // [tiles]
  uint32_t* n_cand;     // [tiles]
  uint32_t* cand_id;    // [tiles * cand_cap]
  uint32_t* cand_cnt;   // [tiles * cand_cap]
  uint32_t* tile_id;    // [tiles] smoothed ids
  uint8_t* tile_as;     // [tiles]
  uint32_t* snap;       // [tiles]
  GrbReadPlan* plan;
  uint64_t* tab_key;    // insert de-dup table
  uint64_t* tab_mask;
};

__host__ __device__ __forceinline__ uint32_t
grb_mix32(uint32_t x)
{
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

__device__ __forceinline__ uint64_t
grb_mix64(uint64_t x)
{
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  return x;
}

// One CTA per tile (grid-strided).  Dynamic shared memory:
//   GrbSeedTables | uint64 sw[sw_words] | uint32 keys[table_size] | uint32 counts[table_size]
template<int BS>
__global__ void __launch_bounds__(BS)
k_query(GrbReadsDev reads, const GrbSeedTables* __restrict__ seeds_g, GrbFilterDev filt,
        GrbSelParams prm, GrbSelScratch sc, GrbSelState* __restrict__ state, uint64_t read_idx)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GrbSeedTables& st = *reinterpret_cast<GrbSeedTables*>(smem_raw);
  uint64_t* sw = reinterpret_cast<uint64_t*>(smem_raw + sizeof(GrbSeedTables));
  uint32_t* keys = reinterpret_cast<uint32_t*>(sw + prm.sw_words);
  uint32_t* counts = keys + prm.table_size;
  __shared__ uint32_t s_ncand;
  __shared__ unsigned long long s_best;
  __shared__ uint32_t s_hits, s_miss;

  if (state->halt) {
    return;
  }
  for (unsigned i = threadIdx.x; i < sizeof(GrbSeedTables) / 8; i += BS) {
    reinterpret_cast<uint64_t*>(&st)[i] = reinterpret_cast<const uint64_t*>(seeds_g)[i];
  }
  const uint32_t len = reads.len[read_idx];
  const uint64_t w_read = reads.word_off[read_idx];
  const uint32_t w_total = (len + 31) / 32;
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h;
  const uint32_t n_tiles = len / T;
  const uint32_t tmask = prm.table_size - 1;
  uint32_t my_hits = 0, my_miss = 0;
  uint64_t my_queries = 0;

  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
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
      s_ncand = 0;
      s_best = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      my_queries += frames;
    }
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
          sc.stash[((uint64_t)t * prm.tile_frames + f) * h + i] = rank[i];
        }
      }
      if (!all) { // MIBloomFilter::atRank fails on the first clear bit: the frame counts nothing
        continue;
      }
      uint32_t ids[GRB_MAX_PATTERNS];
#pragma unroll
      for (unsigned i = 0; i < GRB_MAX_PATTERNS; ++i) {
        if (i < h) {
          uint32_t v = __ldcg(&filt.slots[rank[i]].id);
          if (v > GRB_SAT_MASK) {
            v &= ~GRB_SAT_MASK;
          }
          ids[i] = v;
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
          if (dup) {
            continue;
          }
          uint32_t slot = grb_mix32(v) & tmask;
          while (true) {
            const uint32_t old = atomicCAS(&keys[slot], 0u, v);
            if (old == 0u || old == v) {
              atomicAdd(&counts[slot], 1u);
              break;
            }
            slot = (slot + 1) & tmask;
          }
        }
      }
    }
    __syncthreads();
    // arg-max (ties -> smallest id, goldrush_path.cpp:610-615) and the count > 2 list (:616-619)
    unsigned long long best = 0;
    for (unsigned i = threadIdx.x; i < prm.table_size; i += BS) {
      const uint32_t c = counts[i];
      if (c) {
        const uint32_t id = keys[i];
        const unsigned long long key = ((unsigned long long)c << 32) | (0xFFFFFFFFu - id);
        best = key > best ? key : best;
        if (c > 2) {
          const uint32_t at = atomicAdd(&s_ncand, 1u);
          if (at < prm.cand_cap) {
            sc.cand_id[(uint64_t)t * prm.cand_cap + at] = id;
            sc.cand_cnt[(uint64_t)t * prm.cand_cap + at] = c;
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
    if (threadIdx.x == 0) {
      const unsigned long long b = s_best;
      sc.best_count[t] = (uint32_t)(b >> 32);
      sc.best_id[t] = b ? 0xFFFFFFFFu - (uint32_t)(b & 0xFFFFFFFFu) : 0u;
      sc.n_cand[t] = s_ncand;
    }
  }
  // counters of log_info_struct (goldrush_path.cpp:46-48)
  if (threadIdx.x == 0) {
    s_hits = 0;
    s_miss = 0;
  }
  __syncthreads();
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    my_hits += __shfl_xor_sync(0xffffffffu, my_hits, d);
    my_miss += __shfl_xor_sync(0xffffffffu, my_miss, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_hits, my_hits);
    atomicAdd(&s_miss, my_miss);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (my_queries) {
      atomicAdd(reinterpret_cast<unsigned long long*>(&state->cur.queries),
                (unsigned long long)my_queries);
    }
    if (s_hits) {
      atomicAdd(reinterpret_cast<unsigned long long*>(&state->cur.hits), (unsigned long long)s_hits);
    }
    if (s_miss) {
      atomicAdd(reinterpret_cast<unsigned long long*>(&state->cur.misses),
                (unsigned long long)s_miss);
    }
  }
}

// One thread: smoothing, decision, ID bookkeeping, path rollover.
__global__ void
k_decide(GrbReadsDev reads, GrbSelParams prm, GrbSelScratch sc, GrbSelState* __restrict__ state,
         grb_decision* __restrict__ decisions, uint64_t read_idx, uint64_t dec_idx)
{
  if (threadIdx.x != 0 || state->halt) {
    return;
  }
  const uint32_t len = reads.len[read_idx];
  const uint32_t n_tiles = len / prm.tile_len;
  GrbTileVotes v{ sc.best_id, sc.best_count, sc.n_cand, sc.cand_id, sc.cand_cnt, prm.cand_cap };
  const uint32_t n_as =
    grb_smooth_tiles(n_tiles, v, prm.threshold, sc.tile_id, sc.tile_as, sc.snap);
  GrbSelState s = *state;
  s.cur.total_tiles += n_tiles;
  s.cur.assigned_tiles += n_as;
  s.cur.unassigned_tiles += n_tiles - n_as;
  GrbReadPlan plan;
  grb_plan_read(n_tiles, n_as, len, prm.tile_len, prm.block_size, prm.unassigned_min,
                prm.assigned_max, sc.tile_id, sc.tile_as, &s.ids_inserted, &plan);
  grb_decision d;
  d.verdict = plan.verdict;
  d.pad[0] = d.pad[1] = d.pad[2] = 0;
  d.path = (uint32_t)s.curr_path;
  d.trim_start = plan.trim_start;
  d.trim_end = plan.trim_end;
  d.num_tiles = n_tiles;
  d.num_assigned = n_as;
  decisions[dec_idx] = d;
  if (plan.verdict == GRB_UNTRIMMED || plan.verdict == GRB_TRIMMED) {
    s.cur.inserted_bases += plan.out_bases;
    s.cur.num_reads_in_path += 1;
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
  // queries/hits/misses were accumulated straight into state->cur by k_query (stream order)
  *state = s;
  *sc.plan = plan;
}

// Insert, step 1: every valid (tile, frame, pattern) of the planned tile range registers
// (rank -> bit of its insert block) in a global open-addressing table, which de-duplicates ranks
// per insert call exactly like the dense_hash_set of MIBFConstructSupport.hpp:255-270 while
// keeping calls that share a rank apart.  `round` selects insert blocks [64*round, 64*round+64).
__global__ void __launch_bounds__(256)
k_insert_collect(GrbReadsDev reads, GrbSelParams prm, GrbSelScratch sc,
                 const uint64_t* __restrict__ stash, const GrbSelState* __restrict__ state,
                 uint64_t read_idx, uint32_t round, uint32_t tab_size)
{
  const GrbReadPlan plan = *sc.plan;
  if (plan.n_blocks <= 64 * round || (plan.verdict != GRB_UNTRIMMED && plan.verdict != GRB_TRIMMED)) {
    return;
  }
  if (state->halt && state->halt_read != read_idx) {
    return;
  }
  const uint32_t len = reads.len[read_idx];
  const uint32_t T = prm.tile_len, k = prm.k, h = prm.h;
  const uint64_t B = prm.block_size;
  const uint64_t first_tile = plan.trim_start + 64ull * round * B;
  uint64_t last_tile = first_tile + 64ull * B - 1;
  if (last_tile > plan.trim_end) {
    last_tile = plan.trim_end;
  }
  const uint64_t per_tile = (uint64_t)prm.tile_frames * h;
  const uint64_t total = (last_tile - first_tile + 1) * per_tile;
  const uint64_t mask = tab_size - 1;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t trel = idx / per_tile;
    const uint32_t rem = (uint32_t)(idx - trel * per_tile);
    const uint32_t f = rem / h, p = rem - f * h;
    const uint32_t t = (uint32_t)(first_tile + trel);
    const uint32_t tl = grb_tile_bases(len, t, T, prm.kmer);
    if (tl < k + p || f >= tl - (k + p) + 1) {
      continue; // stale-tail repeat of the last valid position: same rank, already registered
    }
    const uint64_t key = stash[((uint64_t)t * prm.tile_frames + f) * h + p] & ~(1ull << 63);
    const uint32_t j = (uint32_t)((t - plan.trim_start) / B) - 64u * round;
    uint64_t slot = grb_mix64(key) & mask;
    while (true) {
      const unsigned long long old =
        atomicCAS(reinterpret_cast<unsigned long long*>(&sc.tab_key[slot]), GRB_EMPTY_KEY, key);
      if (old == GRB_EMPTY_KEY || old == key) {
        atomicOr(reinterpret_cast<unsigned long long*>(&sc.tab_mask[slot]), 1ull << j);
        break;
      }
      slot = (slot + 1) & mask;
    }
  }
}

// Insert, step 2: one thread per distinct rank applies its insert calls in block order:
//   count = ++m_counts[rank]; if (uint32(rank ^ id) % count == count - 1) setData(rank, id)
// (MIBFConstructSupport.hpp:271-282, MIBloomFilter.hpp:593-602), then clears the table entry.
__global__ void __launch_bounds__(256)
k_insert_apply(GrbFilterDev filt, GrbSelScratch sc, const GrbSelState* __restrict__ state,
               uint64_t read_idx, uint32_t round, uint32_t tab_size)
{
  const GrbReadPlan plan = *sc.plan;
  if (plan.n_blocks <= 64 * round || (plan.verdict != GRB_UNTRIMMED && plan.verdict != GRB_TRIMMED)) {
    return;
  }
  if (state->halt && state->halt_read != read_idx) {
    return;
  }
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tab_size;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t key = sc.tab_key[i];
    if (key == GRB_EMPTY_KEY) {
      continue;
    }
    uint64_t m = sc.tab_mask[i];
    GrbSlot s = filt.slots[key];
    while (m) {
      const uint32_t j = __ffsll((long long)m) - 1;
      m &= m - 1;
      const uint32_t id = plan.first_id + 64u * round + j + plan.id_bump;
      const uint32_t count = ++s.count;
      if ((uint32_t)(key ^ (uint64_t)id) % count == count - 1) {
        s.id = s.id > GRB_SAT_MASK ? (id | GRB_SAT_MASK) : id;
      }
    }
    filt.slots[key] = s;
    sc.tab_key[i] = GRB_EMPTY_KEY;
    sc.tab_mask[i] = 0;
  }
}
