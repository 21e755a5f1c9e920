// Per-read decision logic of GoldRush-Path, shared verbatim by the device kernel that runs it
// (one thread per read, select.cu) and by the host test hook grb_test_decide_host.
//
// Mirrors, without copying, goldrush_path/goldrush_path.cpp:
//   :628-889  threshold + the tile smoothing passes of calc_num_assigned_tiles
//   :195-233  find_longest_stretch
//   :341-527  eval_flanks
//   :968-1053 the insert / trim decision and the ID bookkeeping of process_read
// All "+-1" id comparisons are 32-bit wrap-around arithmetic except the end-tile rule (:827-838),
// which the reference evaluates in size_t.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define GRB_HD __host__ __device__ __forceinline__
#else
#define GRB_HD inline
#endif

// The smoothing code reads the per-tile votes through three accessors, so the same source serves
// the candidate-list layout below and the dense count matrix of the batch engine
// (batch_common.cuh: GrbMatrixVotes):
//   best_id(i), best_count(i)   arg-max id of tile i (ties -> smallest id) and its count
//   cand_count(i, id)           count of `id` in tile i if it is > 2 (a "candidate"), else 0
struct GrbTileVotes
{
  const uint32_t* best_id_;    // [n] arg-max id (ties -> smallest id), 0 if the tile saw no id
  const uint32_t* best_count_; // [n]
  const uint32_t* n_cand;      // [n] ids with count > 2
  const uint32_t* cand_id;     // [n * cand_cap]
  const uint32_t* cand_cnt;    // [n * cand_cap]
  uint32_t cand_cap;
  GRB_HD uint32_t best_id(uint32_t i) const { return best_id_[i]; }
  GRB_HD uint32_t best_count(uint32_t i) const { return best_count_[i]; }
  GRB_HD uint32_t cand_count(uint32_t i, uint32_t id) const
  {
    const uint32_t n = n_cand[i];
    const uint32_t* ids = cand_id + (uint64_t)i * cand_cap;
    for (uint32_t j = 0; j < n; ++j) {
      if (ids[j] == id) {
        return cand_cnt[(uint64_t)i * cand_cap + j];
      }
    }
    return 0;
  }
};

GRB_HD bool
grb_near(uint32_t a, uint32_t b)
{
  return a == b || a == (uint32_t)(b + 1u) || a == (uint32_t)(b - 1u);
}

// id[n], as[n] are outputs; snap[n + 2] is scratch.  Returns the number of assigned tiles.
template<class Votes>
GRB_HD uint32_t
grb_smooth_tiles(uint32_t n, const Votes& v, uint64_t threshold, uint32_t* id, uint8_t* as,
                 uint32_t* snap)
{
  for (uint32_t i = 0; i < n; ++i) {
    id[i] = v.best_id(i);
    // the candidate list is sorted by count, so its head is the arg-max count whenever it is > 2
    const uint32_t bc = v.best_count(i);
    as[i] = (bc > 2 && bc > threshold) ? 1 : 0;
  }
  if (n >= 3) {
    // adopt the neighbour's id when it is one of this tile's candidates: forward, then backward
    for (uint32_t i = 1; i < n; ++i) {
      const uint32_t nb = id[i - 1];
      if (id[i] != nb) {
        const uint32_t c = v.cand_count(i, nb);
        if (c) {
          id[i] = nb;
          as[i] = c > threshold ? 1 : 0;
        }
      }
    }
    for (uint32_t i = n - 1; i-- > 0;) {
      const uint32_t nb = id[i + 1];
      if (id[i] != nb) {
        const uint32_t c = v.cand_count(i, nb);
        if (c) {
          id[i] = nb;
          as[i] = c > threshold ? 1 : 0;
        }
      }
    }
    uint32_t any = 0;
    for (uint32_t i = 0; i < n; ++i) {
      any |= as[i];
    }
    if (any) {
      // unassigned tile next to an assigned tile with the same / adjacent id, or between two
      // assigned tiles that agree: forward, then backward
      for (int dir = 0; dir < 2; ++dir) {
        for (uint32_t s = 1; s + 1 < n; ++s) {
          const uint32_t i = dir ? (n - 1 - s) : s;
          if (as[i]) {
            continue;
          }
          const uint32_t c = id[i], p = id[i - 1], q = id[i + 1];
          const bool pa = as[i - 1], qa = as[i + 1];
          if ((pa && grb_near(c, p)) || (qa && grb_near(c, q))) {
            as[i] = 1;
          } else if (p == q && pa && qa) {
            as[i] = 1;
            id[i] = p;
          }
        }
      }
      // bridge interior unassigned runs whose flanking ids agree (runs found before any change)
      {
        uint32_t start = 0, nruns = 0;
        for (uint32_t i = 1; i + 1 < n; ++i) {
          if (!as[i] && as[i - 1]) {
            start = i;
          } else if (as[i] && !as[i - 1]) {
            snap[2 * nruns] = start;
            snap[2 * nruns + 1] = i - 1;
            ++nruns;
          }
        }
        for (uint32_t r = 0; r < nruns; ++r) {
          const uint32_t a = snap[2 * r], b = snap[2 * r + 1];
          if (a == 0 || b == n - 1) {
            continue;
          }
          const uint32_t left = id[a - 1], right = id[b + 1];
          if (grb_near(left, right)) {
            for (uint32_t i = a; i <= b; ++i) {
              as[i] = 1;
              id[i] = left;
            }
          }
        }
      }
      // isolated assigned tiles away from the ends: forward, then backward
      for (uint32_t i = 2; i + 2 < n; ++i) {
        if (as[i] && !as[i - 1] && !as[i + 1]) {
          as[i] = 0;
        }
      }
      for (uint32_t i = n - 3; i >= 2 && i < n; --i) {
        if (as[i] && !as[i - 1] && !as[i + 1]) {
          as[i] = 0;
        }
      }
      // per-id gap fill: ids in ascending order, occurrence lists taken before any change
      {
        for (uint32_t i = 0; i < n; ++i) {
          snap[i] = id[i];
        }
        bool have_prev = false;
        uint32_t prev_key = 0;
        while (true) {
          bool found = false;
          uint32_t key = 0;
          for (uint32_t i = 0; i < n; ++i) {
            if (as[i] && (!have_prev || snap[i] > prev_key) && (!found || snap[i] < key)) {
              key = snap[i];
              found = true;
            }
          }
          if (!found) {
            break;
          }
          bool first = true;
          uint32_t last_idx = 0;
          for (uint32_t i = 0; i < n; ++i) {
            if (as[i] && snap[i] == key) {
              if (!first && i > last_idx + 1) {
                const uint32_t fillv = id[last_idx];
                for (uint32_t t = last_idx + 1; t <= i; ++t) {
                  id[t] = fillv;
                }
              }
              first = false;
              last_idx = i;
            }
          }
          prev_key = key;
          have_prev = true;
        }
      }
    }
    // end tiles join a neighbour with the same / adjacent id (64-bit arithmetic here)
    {
      const uint64_t last = id[n - 1], last2 = id[n - 2], first = id[0], first2 = id[1];
      if (last == last2 || last == last2 + 1 || last == last2 - 1) {
        as[n - 1] = 1;
      }
      if (first == first2 || first == first2 + 1 || first == first2 - 1) {
        as[0] = 1;
      }
    }
    // drop interior tiles whose id is unrelated to both neighbours
    for (uint32_t i = 1; i + 1 < n; ++i) {
      if (as[i] && !grb_near(id[i], id[i + 1]) && !grb_near(id[i], id[i - 1])) {
        as[i] = 0;
      }
    }
    // drop assigned runs of at most 5 tiles (runs found before any change)
    {
      uint32_t start = 0, nruns = 0;
      for (uint32_t i = 1; i + 1 < n; ++i) {
        if (as[i] && !as[i - 1]) {
          start = i;
        } else if (!as[i] && as[i - 1]) {
          snap[2 * nruns] = start;
          snap[2 * nruns + 1] = i - 1;
          ++nruns;
        }
      }
      for (uint32_t r = 0; r < nruns; ++r) {
        const uint32_t a = snap[2 * r], b = snap[2 * r + 1];
        if (b - a + 1 <= 5) {
          for (uint32_t i = a; i <= b; ++i) {
            as[i] = 0;
          }
        }
      }
    }
  }
  uint32_t assigned = 0;
  for (uint32_t i = 0; i < n; ++i) {
    assigned += as[i] ? 1u : 0u;
  }
  return assigned;
}

GRB_HD void
grb_find_longest_stretch(const uint8_t* as, uint32_t n, int64_t* ls, int64_t* le)
{
  uint32_t start = 0, end = 0, cur = 0, best = 0;
  *ls = 0;
  *le = 0;
  for (uint32_t i = 1; i + 1 < n; ++i) {
    const bool a = as[i], p = as[i - 1];
    bool close = false;
    if (!a && p) {
      start = i;
      cur = 1;
    } else if (!a && !p && i + 1 != n - 1) {
      ++cur;
    } else if (a && !p) {
      end = i - 1;
      close = true;
    } else if (i + 1 == n - 1 && end < start) {
      end = i;
      ++cur;
      close = true;
    }
    if (close && best < cur) {
      best = cur;
      *ls = (int64_t)start;
      *le = (int64_t)end;
    }
  }
}

// highest multiplicity of any id among tiles [lo, hi]
GRB_HD uint32_t
grb_top_count(const uint32_t* id, int64_t lo, int64_t hi)
{
  uint32_t best = 0;
  for (int64_t i = lo; i <= hi; ++i) {
    uint32_t c = 0;
    for (int64_t j = lo; j <= hi; ++j) {
      c += id[j] == id[i] ? 1u : 0u;
    }
    best = c > best ? c : best;
  }
  return best;
}

GRB_HD bool
grb_eval_flanks(int64_t ls, int64_t le, const uint32_t* id, uint32_t n, uint32_t* trim_start,
                uint32_t* trim_end)
{
  uint64_t ts = ls != 0 ? (uint64_t)(ls - 1) : (uint64_t)ls;
  uint64_t te = (uint64_t)(le + 1);
  bool good = false;
  if (n < 15) {
    bool gl = false, gr = false;
    if (ls - 1 >= 0 && grb_top_count(id, 0, ls - 1) >= 2) {
      gl = true;
    }
    if (ts == 0) {
      gl = true;
    }
    if (le + 1 < (int64_t)n && grb_top_count(id, le + 1, (int64_t)n - 1) >= 2) {
      gr = true;
    }
    if (te == (uint64_t)n - 1) {
      gr = true;
    }
    good = gl && gr;
  } else {
    if (ls - 5 >= 1) {
      if (grb_top_count(id, ls - 5, ls - 1) >= 2) {
        good = true;
      }
    } else {
      good = true;
      ts = 0;
    }
    if (le + 5 < (int64_t)n - 1) {
      if (grb_top_count(id, le + 1, le + 5) >= 2) {
        good = true;
      }
    } else {
      good = true;
      te = (uint64_t)n - 1;
    }
  }
  *trim_start = (uint32_t)ts;
  *trim_end = (uint32_t)te;
  return good;
}

// What process_read decides for one visited read (goldrush_path.cpp:960-1080).
struct alignas(16) GrbReadPlan
{
  uint8_t verdict;      // grb_verdict
  uint32_t trim_start;  // first inserted tile
  uint32_t trim_end;    // last inserted tile (inclusive)
  uint32_t first_id;    // id of the first insert block
  uint32_t id_bump;     // 1 when the first trimmed block already gets first_id + 1 (block_size 1)
  uint32_t n_blocks;    // insert calls
  uint64_t out_bases;   // bases written to the path
};

// ids_inserted is updated in place exactly as :982-994 / :1040-1053 do.
GRB_HD void
grb_plan_read(uint32_t n_tiles, uint32_t n_assigned, uint64_t read_len, uint64_t tile_length,
              uint64_t block_size, uint64_t unassigned_min, uint64_t assigned_max,
              const uint32_t* id, const uint8_t* as, uint32_t* ids_inserted, GrbReadPlan* plan)
{
  const uint64_t n_un = (uint64_t)n_tiles - n_assigned;
  plan->trim_start = 0;
  plan->trim_end = 0;
  plan->first_id = 0;
  plan->id_bump = 0;
  plan->n_blocks = 0;
  plan->out_bases = 0;
  if (n_un >= unassigned_min && n_assigned <= assigned_max) {
    plan->verdict = 2; // GRB_UNTRIMMED
    *ids_inserted += 1;
    plan->first_id = *ids_inserted;
    plan->trim_end = n_tiles ? n_tiles - 1 : 0;
    plan->n_blocks = (uint32_t)((n_tiles + block_size - 1) / block_size);
    *ids_inserted += (uint32_t)(read_len / (tile_length * block_size));
    plan->out_bases = read_len;
    return;
  }
  plan->verdict = 4; // GRB_ASSIGNED
  if (n_assigned == n_tiles) {
    return;
  }
  int64_t ls, le;
  grb_find_longest_stretch(as, n_tiles, &ls, &le);
  uint32_t ts, te;
  if (!grb_eval_flanks(ls, le, id, n_tiles, &ts, &te)) {
    return;
  }
  plan->verdict = 3; // GRB_TRIMMED
  *ids_inserted += 1;
  plan->first_id = *ids_inserted;
  plan->id_bump = block_size == 1 ? 1u : 0u;
  plan->trim_start = ts;
  plan->trim_end = te;
  plan->n_blocks = te >= ts ? (uint32_t)((te - ts) / block_size + 1) : 0;
  *ids_inserted += (uint32_t)((te - ts) / block_size);
  plan->out_bases =
    (te == n_tiles - 1) ? read_len - (uint64_t)ts * tile_length : (uint64_t)(te - ts + 1) * tile_length;
}
