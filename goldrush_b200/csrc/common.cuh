// Shared declarations of the engine's translation units: the context, the device-side views of
// the read store and of the multi-index Bloom filter, and small device utilities.
#pragma once
#include "goldrush_b200.h"
#include "nthash.cuh"

#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <string>
#include <vector>

// ---------------------------------------------------------------------------------------------
// HBM layout of the filter (replaces sdsl::bit_vector_il<512> + rank_support_il<1>,
// goldrush_path/MIBFConstructSupport.hpp:165-170): 32-byte blocks = one DRAM sector each,
//   word 0      number of set bits in all earlier blocks (filled by grb_finalize_bitvector)
//   words 1..3  192 filter bits, LSB first
// so one probe (bit test + rank) touches exactly one sector.  The ID / count pair of the slot
// with that rank (MIBloomFilter::m_data + MIBFConstructSupport::m_counts, 4 + 4 bytes per set bit,
// MIBloomFilter.hpp:538-546, MIBFConstructSupport.hpp:336-339) is one 8-byte GrbSlot {id, count}:
// a query reads its sector, an insert read-modify-writes it.
// ---------------------------------------------------------------------------------------------
#define GRB_BLK_BITS 192ull

struct __align__(8) GrbSlot
{
  uint32_t id;    // MIBloomFilter::m_data[rank] (bit 31 = saturation mask, MIBloomFilter.hpp:38)
  uint32_t count; // MIBFConstructSupport::m_counts[rank]
};

struct GrbFilterDev
{
  uint64_t* blocks; // 4 words per block
  uint64_t bits;    // m = filter size in bits
  uint64_t inv;     // floor(2^64 / m) for the exact fast modulo
  uint64_t n_blocks;
  GrbSlot* slots;   // one per set bit, indexed by rank
  uint64_t pop;
};

struct GrbReadsDev
{
  const uint64_t* bases;   // 2-bit packed, 32 bases per word; every read starts on a word boundary
  const uint32_t* nmask;   // 1 bit per base: byte was not ACGTacgt
  const uint64_t* word_off; // [n] first word of read i
  const uint32_t* len;      // [n] bases
  const uint8_t* flags;     // [n] GRB_READ_PASS1 | GRB_READ_PASS2 | 4 = has non-ACGT
};

__device__ __forceinline__ uint64_t
grb_fastmod(uint64_t x, uint64_t m, uint64_t inv)
{
  const uint64_t q = __umul64hi(x, inv);
  uint64_t r = x - q * m;
  if (r >= m) {
    r -= m;
  }
  return r;
}

__device__ __forceinline__ uint64_t
grb_div3(uint64_t x)
{
  return __umul64hi(x, 0xAAAAAAAAAAAAAAABull) >> 1;
}

// bit test + rank of filter position `pos` from its 32-byte block (MIBloomFilter.hpp:465-491)
__device__ __forceinline__ void
grb_probe_block(const GrbFilterDev& f, uint64_t pos, bool& bit, uint64_t& rank)
{
  const uint64_t blk = grb_div3(pos >> 6);
  const unsigned r = (unsigned)(pos - blk * GRB_BLK_BITS);
  // one 32-byte load (LDG.256, sm_100): the whole block in one request to L2 instead of two
  // (tools/sector_roofline.cu: random 32-byte sectors, 36.3 G/s with one load against 30-34 with two)
  ulonglong2 a, b; // cum, w0 | w1, w2
  asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
      : "=l"(a.x), "=l"(a.y), "=l"(b.x), "=l"(b.y)
      : "l"(f.blocks + blk * 4));
  const unsigned wi = r >> 6, bi = r & 63;
  const uint64_t w = wi == 0 ? a.y : (wi == 1 ? b.x : b.y);
  uint64_t rk = a.x;
  rk += wi > 0 ? __popcll(a.y) : 0;
  rk += wi > 1 ? __popcll(b.x) : 0;
  rk += __popcll(w & ((1ull << bi) - 1ull));
  bit = (w >> bi) & 1ull;
  rank = rk;
}

// frames and per-pattern valid positions of tile t of a read (read_hashing.cpp:43-46 and the
// stale-tail rule of multiLensfrHashIterator.hpp:49-68)
__host__ __device__ __forceinline__ uint32_t
grb_tile_bases(uint32_t read_len, uint32_t tile, uint32_t tile_len, uint32_t k)
{
  const uint64_t start = (uint64_t)tile * tile_len;
  const uint64_t want = (uint64_t)tile_len + k - 1;
  const uint64_t have = read_len - start;
  return (uint32_t)(want < have ? want : have);
}

struct grb_ctx;

#define GRB_CUDA(ctx, expr)                                                                        \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      return (ctx)->fail(GRB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));       \
    }                                                                                              \
  } while (0)

// Process-wide cache of device allocations (engine.cu).  cudaMalloc / cudaFree of the tens of GB behind
// a filter cost hundreds of milliseconds each, so a destroyed context hands its blocks back to the
// cache and the next context of the same shape (the next grb_run_path call of a service, the next
// bench step) takes them over.  grb_release_cached_memory() empties it; GRB_POOL=0 turns it off.
cudaError_t grb_pool_alloc(void** p, size_t bytes);
void grb_pool_free(void* p);

template<typename T>
struct DevBuf
{
  T* p = nullptr;
  size_t cap = 0; // elements
  ~DevBuf() { release(); }
  void release()
  {
    if (p) {
      grb_pool_free(p);
    }
    p = nullptr;
    cap = 0;
  }
  // grows (keeping contents up to `keep` elements) with 1.5x slack
  cudaError_t reserve(size_t n, size_t keep, cudaStream_t s)
  {
    if (n <= cap) {
      return cudaSuccess;
    }
    size_t ncap = n + n / 2 + 64;
    T* q = nullptr;
    cudaError_t e = grb_pool_alloc((void**)&q, ncap * sizeof(T));
    if (e != cudaSuccess) {
      ncap = n;
      e = grb_pool_alloc((void**)&q, ncap * sizeof(T));
      if (e != cudaSuccess) {
        return e;
      }
    }
    if (p && keep) {
      e = cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) {
        grb_pool_free(q);
        return e;
      }
    }
    if (p) {
      // the old block may be handed to another context: nothing of this stream may still use it
      cudaStreamSynchronize(s);
      grb_pool_free(p);
    }
    p = q;
    cap = ncap;
    return cudaSuccess;
  }
  // exact-size variant for the big one-shot buffers (no 1.5x slack)
  cudaError_t reserve_exact(size_t n, cudaStream_t s)
  {
    if (n <= cap) {
      return cudaSuccess;
    }
    if (p) {
      cudaStreamSynchronize(s);
      grb_pool_free(p);
      p = nullptr;
      cap = 0;
    }
    cudaError_t e = grb_pool_alloc((void**)&p, n * sizeof(T));
    if (e == cudaSuccess) {
      cap = n;
    }
    return e;
  }
};
