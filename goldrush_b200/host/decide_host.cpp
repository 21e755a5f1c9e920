// TEST HOOK — runs the per-read decision code of csrc/decide.cuh (the exact source the device
// kernel k_decide compiles) on the host, so the smoothing / trimming logic can be fuzzed against
// the oracle on a machine without a GPU.  Nothing in the product path calls this.
#include "../csrc/decide.cuh"
#include "goldrush_b200.h"

#include <vector>

extern "C" int
grb_test_decide_host(uint32_t n_tiles, const uint32_t* best_id, const uint32_t* best_count,
                     const uint32_t* n_cand, const uint32_t* cand_id, const uint32_t* cand_cnt,
                     uint32_t cand_cap, uint64_t threshold, uint64_t read_len, uint64_t tile_length,
                     uint64_t block_size, uint64_t unassigned_min, uint64_t assigned_max,
                     uint32_t* ids_inserted, uint32_t* out_ids, uint8_t* out_assigned,
                     uint32_t* out_plan /* verdict, trim_start, trim_end, first_id, id_bump,
                                           n_blocks, n_assigned, out_bases_lo, out_bases_hi */)
{
  GrbTileVotes v{ best_id, best_count, n_cand, cand_id, cand_cnt, cand_cap };
  std::vector<uint32_t> snap(n_tiles + 2);
  const uint32_t n_as =
    grb_smooth_tiles(n_tiles, v, threshold, out_ids, out_assigned, snap.data());
  GrbReadPlan plan;
  grb_plan_read(n_tiles, n_as, read_len, tile_length, block_size, unassigned_min, assigned_max,
                out_ids, out_assigned, ids_inserted, &plan);
  out_plan[0] = plan.verdict;
  out_plan[1] = plan.trim_start;
  out_plan[2] = plan.trim_end;
  out_plan[3] = plan.first_id;
  out_plan[4] = plan.id_bump;
  out_plan[5] = plan.n_blocks;
  out_plan[6] = n_as;
  out_plan[7] = (uint32_t)(plan.out_bases & 0xFFFFFFFFu);
  out_plan[8] = (uint32_t)(plan.out_bases >> 32);
  return 0;
}
