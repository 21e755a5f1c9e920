// miBF probe microbenchmark (BASELINE.json configs[4]): query / insert throughput
// of the filter's own probe sequence against filter footprint and the number of seed patterns h,
// on synthetic keys.  The probes are exactly the product path's (goldrush_path.cpp:544-626 query:
// block probe = bit test + rank from one 32-byte block, then the ID slot; MIBFConstructSupport.hpp:
// 247-283 insert: rank, then the reservoir read-modify-write of the slot) without hashing or voting,
// so the numbers bound what k2_query / k3_bulk can reach at a given footprint.
//
// Key c has the hashes splitmix64(seed + c * 8 + j), j < h.  The filter is filled from keys
// [0, n_fill), and the queried / inserted keys are drawn from the same range, so every probe meets
// a set bit, as on the product path (every queried k-mer was hashed into the bit vector in pass 1).
#pragma once
#include "common.cuh"
#include "kernels_filter.cuh"
#include "batch_common.cuh"

__host__ __device__ __forceinline__ uint64_t
grb_splitmix64(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256)
k_probe_fill(GrbFilterDev filt, uint64_t n_fill, uint32_t h, uint64_t seed)
{
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_fill;
       c += (uint64_t)gridDim.x * blockDim.x) {
    for (uint32_t j = 0; j < h; ++j) {
      grb_set_bit_pos(filt, grb_fastmod(grb_splitmix64(seed + c * 8 + j), filt.bits, filt.inv));
    }
  }
}

// half of the slots hold an ID ("IDs pre-filled at 50 % non-zero")
__global__ void __launch_bounds__(256)
k_probe_ids(GrbSlot* __restrict__ slots, uint64_t pop, uint64_t seed)
{
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < pop;
       r += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t m = grb_splitmix64(seed ^ r);
    const uint32_t id = (m & 1) ? (uint32_t)((m >> 8) & 0x00FFFFFFu) | 1u : 0u;
    slots[r] = GrbSlot{ id, id ? 1u : 0u };
  }
}

// thread per key: h block probes, then h slot reads (the order k2_query issues them in)
__global__ void __launch_bounds__(256)
k_probe_query(GrbFilterDev filt, uint64_t n_keys, uint64_t n_fill, uint32_t h, uint64_t seed,
              uint64_t pick, unsigned long long* __restrict__ checksum)
{
  unsigned long long sum = 0, missed = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_keys;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t c = grb_splitmix64(pick + i) % n_fill; // `pick` varies per launch, `seed` never
    uint64_t rank[GRB_MAX_PATTERNS];
    bool all = true;
#pragma unroll
    for (uint32_t j = 0; j < GRB_MAX_PATTERNS; ++j) {
      if (j < h) {
        bool bit;
        grb_probe_block(filt, grb_fastmod(grb_splitmix64(seed + c * 8 + j), filt.bits, filt.inv), bit,
                        rank[j]);
        all &= bit;
      }
    }
    if (!all) {
      ++missed; // never happens: counted so that it would show (checksum[1])
      continue;
    }
#pragma unroll
    for (uint32_t j = 0; j < GRB_MAX_PATTERNS; ++j) {
      if (j < h) {
        sum += grb_norm_id(__ldcg(&filt.slots[rank[j]].id));
      }
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, d);
    missed += __shfl_xor_sync(0xffffffffu, missed, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (sum) {
      atomicAdd(&checksum[0], sum);
    }
    if (missed) {
      atomicAdd(&checksum[1], missed);
    }
  }
}

// thread per key: h block probes, then the reservoir read-modify-write of each slot (k3_bulk's
// step; two keys meeting in one slot race here as they would not in the product, which is fine
// for a throughput figure)
__global__ void __launch_bounds__(256)
k_probe_insert(GrbFilterDev filt, uint64_t n_keys, uint64_t n_fill, uint32_t h, uint64_t seed,
               uint64_t pick, uint32_t id)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_keys;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t c = grb_splitmix64(pick + i) % n_fill; // `pick` varies per launch, `seed` never
    uint64_t rank[GRB_MAX_PATTERNS];
#pragma unroll
    for (uint32_t j = 0; j < GRB_MAX_PATTERNS; ++j) {
      if (j < h) {
        bool bit;
        grb_probe_block(filt, grb_fastmod(grb_splitmix64(seed + c * 8 + j), filt.bits, filt.inv), bit,
                        rank[j]);
      }
    }
#pragma unroll
    for (uint32_t j = 0; j < GRB_MAX_PATTERNS; ++j) {
      if (j < h) {
        uint2* slot = reinterpret_cast<uint2*>(&filt.slots[rank[j]]);
        uint2 s = __ldcg(slot);
        grb2_reservoir(rank[j], id + (uint32_t)(i & 1023), s.x, s.y);
        *slot = s;
      }
    }
  }
}

// What a line-local filter layout could reach: one random 128-byte line per probe,
// two dependent 16-byte reads inside it (header with bits + prefix, then the ID it selects).  `lines`
// is any buffer of n_lines * 128 bytes; its content only feeds the checksum and the dependent offset.
__global__ void __launch_bounds__(256)
k_probe_line(const uint4* __restrict__ lines, uint64_t n_lines, uint64_t n_keys, uint32_t h,
             uint64_t seed, uint64_t pick, unsigned long long* __restrict__ checksum)
{
  unsigned long long sum = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_keys;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t c = grb_splitmix64(pick + i);
    uint4 head[GRB_MAX_PATTERNS];
    uint64_t line[GRB_MAX_PATTERNS];
#pragma unroll
    for (uint32_t j = 0; j < GRB_MAX_PATTERNS; ++j) {
      if (j < h) {
        line[j] = __umul64hi(grb_splitmix64(seed + c * 8 + j), n_lines); // uniform in [0, n_lines)
        head[j] = __ldg(&lines[line[j] * 8]);
      }
    }
#pragma unroll
    for (uint32_t j = 0; j < GRB_MAX_PATTERNS; ++j) {
      if (j < h) {
        const uint4 id = __ldg(&lines[line[j] * 8 + 1 + (head[j].x + head[j].z) % 7]);
        sum += id.y & 0xFFFFu;
      }
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, d);
  }
  if ((threadIdx.x & 31) == 0 && sum) {
    atomicAdd(&checksum[0], sum);
  }
}

