// K1 — batched FASTQ decode on the device: newline index, record table, 2-bit packing with the
// ACGTacgt check, and the sequential double-precision Phred sums.
//
// Replaces btllib::SeqReader record parsing (call sites goldrush_path/goldrush_path.cpp:87,246,
// read_hashing.cpp:89-90, ntcard.hpp:200) and the summation loop of calc_phred_average
// (goldrush_path/calc_phred_average.cpp:15-30).  The final log10 / integer casts stay on the
// host (grb_phred_finalize) so that glibc's log10 decides the integer boundaries exactly as in
// the reference; the per-base 10^(-q/10) values are a host-computed (glibc pow) 256-entry table.
#pragma once
#include "common.cuh"

__constant__ double c_delog[256];

__device__ __forceinline__ uint32_t
grb_warp_incl_scan(uint32_t v)
{
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= (unsigned)d) {
      v += t;
    }
  }
  return v;
}

// exclusive scan of one value per thread across the block; *total = block sum (all threads)
template<int BS>
__device__ __forceinline__ uint32_t
grb_block_excl_scan(uint32_t v, uint32_t* total)
{
  __shared__ uint32_t warp_sums[BS / 32];
  __shared__ uint32_t block_total;
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t incl = grb_warp_incl_scan(v);
  if (lane == 31) {
    warp_sums[wid] = incl;
  }
  __syncthreads();
  if (wid == 0) {
    uint32_t w = lane < BS / 32 ? warp_sums[lane] : 0;
    const uint32_t wi = grb_warp_incl_scan(w);
    if (lane < BS / 32) {
      warp_sums[lane] = wi - w;
    }
    if (lane == 31) {
      block_total = wi;
    }
  }
  __syncthreads();
  const uint32_t r = incl - v + warp_sums[wid];
  *total = block_total;
  __syncthreads();
  return r;
}

__device__ __forceinline__ uint32_t
grb_count_nl16(uint4 v)
{
  const uint32_t nl = 0x0A0A0A0Au;
  return __popc(__vcmpeq4(v.x, nl) & 0x01010101u) + __popc(__vcmpeq4(v.y, nl) & 0x01010101u) +
         __popc(__vcmpeq4(v.z, nl) & 0x01010101u) + __popc(__vcmpeq4(v.w, nl) & 0x01010101u);
}

// 256 threads x 16 bytes per CTA; buf is zero-padded to a multiple of 4096 bytes
__global__ void __launch_bounds__(256)
k_nl_count(const uint8_t* __restrict__ buf, uint32_t* __restrict__ blk_cnt)
{
  const uint64_t base = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 16;
  const uint4 v = *reinterpret_cast<const uint4*>(buf + base);
  uint32_t total;
  grb_block_excl_scan<256>(grb_count_nl16(v), &total);
  if (threadIdx.x == 0) {
    blk_cnt[blockIdx.x] = total;
  }
}

__global__ void __launch_bounds__(256)
k_nl_write(const uint8_t* __restrict__ buf, const uint64_t* __restrict__ blk_off,
           uint32_t* __restrict__ nl_pos)
{
  const uint64_t base = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 16;
  const uint4 v = *reinterpret_cast<const uint4*>(buf + base);
  uint32_t total;
  const uint32_t off = grb_block_excl_scan<256>(grb_count_nl16(v), &total);
  uint64_t o = blk_off[blockIdx.x] + off;
  const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (((w[i >> 2] >> ((i & 3) * 8)) & 0xFFu) == 0x0Au) {
      nl_pos[o++] = (uint32_t)(base + i);
    }
  }
}

// exclusive scan of n uint32 into uint64 by ONE block of 1024 threads; total -> out[n]
__global__ void __launch_bounds__(1024)
k_scan_u32(const uint32_t* __restrict__ in, uint64_t* __restrict__ out, uint64_t n)
{
  __shared__ uint64_t carry;
  if (threadIdx.x == 0) {
    carry = 0;
  }
  __syncthreads();
  for (uint64_t base = 0; base < n; base += 1024) {
    const uint64_t i = base + threadIdx.x;
    const uint32_t v = i < n ? in[i] : 0;
    uint32_t total;
    const uint32_t ex = grb_block_excl_scan<1024>(v, &total);
    if (i < n) {
      out[i] = carry + ex;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      carry += total;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[n] = carry;
  }
}

__device__ __forceinline__ bool
grb_is_ws(uint8_t c)
{
  return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f';
}

// one thread per record: line boundaries -> grb_read_meta (offsets relative to the whole input)
__global__ void
k_records(const uint8_t* __restrict__ buf, uint64_t n_bytes, const uint32_t* __restrict__ nl_pos,
          uint64_t n_nl, uint64_t n_rec, uint64_t chunk_base, grb_read_meta* __restrict__ meta,
          uint32_t* __restrict__ words_per_read, uint32_t* __restrict__ err)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) {
    return;
  }
  uint64_t b[4], e[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint64_t line = 4 * r + j;
    b[j] = line == 0 ? 0 : (uint64_t)nl_pos[line - 1] + 1;
    e[j] = line < n_nl ? (uint64_t)nl_pos[line] : n_bytes;
    while (e[j] > b[j] && grb_is_ws(buf[e[j] - 1])) {
      --e[j];
    }
  }
  if (e[0] == b[0] || buf[b[0]] != '@' || e[2] == b[2] || buf[b[2]] != '+') {
    atomicExch(err, 1u);
  }
  grb_read_meta m;
  m.hdr_off = chunk_base + b[0] + 1;
  m.hdr_len = (uint32_t)(e[0] > b[0] ? e[0] - b[0] - 1 : 0);
  m.seq_off = chunk_base + b[1];
  m.len = (uint32_t)(e[1] - b[1]);
  m.qual_off = chunk_base + b[3];
  m.phred_first_half_sum = 0.0;
  m.phred_total_sum = 0.0;
  m.non_acgt = 0;
  m.qual_len = (uint32_t)(e[3] - b[3]);
  // record id = header up to the first blank or tab (btllib Record::id): FNV-1a, the key under which
  // the host looks a read up among the names dropped in pass 1 / listed by -f
  uint64_t hsh = 1469598103934665603ull;
  for (uint64_t q = b[0] + 1; q < e[0] && buf[q] != ' ' && buf[q] != '\t'; ++q) {
    hsh = (hsh ^ buf[q]) * 1099511628211ull;
  }
  m.name_hash = hsh;
  meta[r] = m;
  words_per_read[r] = (m.len + 31) / 32;
}

// one CTA per record: bytes -> 2-bit codes (A0 C1 G2 T3, case-insensitive), 32 per word, plus the
// "not ACGTacgt" mask (goldrush_path.cpp:293) and the per-read flag
__global__ void __launch_bounds__(256)
k_pack(const uint8_t* __restrict__ buf, uint64_t chunk_base, grb_read_meta* __restrict__ meta,
       const uint64_t* __restrict__ word_off, uint64_t word_base, uint64_t* __restrict__ bases,
       uint32_t* __restrict__ nmask)
{
  __shared__ uint8_t stage[256 * 32];
  __shared__ uint32_t any_bad;
  const uint64_t r = blockIdx.x;
  const uint32_t len = meta[r].len;
  const uint8_t* seq = buf + (meta[r].seq_off - chunk_base);
  const uint64_t wo = word_base + word_off[r];
  const uint32_t nwords = (len + 31) / 32;
  if (threadIdx.x == 0) {
    any_bad = 0;
  }
  for (uint32_t w0 = 0; w0 < nwords; w0 += 256) {
    __syncthreads();
    const uint32_t byte0 = w0 * 32;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const uint32_t j = i * 256 + threadIdx.x;
      stage[j] = (byte0 + j < len) ? seq[byte0 + j] : (uint8_t)'A';
    }
    __syncthreads();
    const uint32_t w = w0 + threadIdx.x;
    if (w < nwords) {
      uint64_t packed = 0;
      uint32_t bad = 0;
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        const uint32_t c = stage[threadIdx.x * 32 + i] & 0xDFu;
        uint32_t code = (c >> 1) & 3u;
        code ^= code >> 1;
        const bool ok = c == 'A' || c == 'C' || c == 'G' || c == 'T';
        packed |= (uint64_t)(ok ? code : 0u) << (2 * i);
        bad |= (ok ? 0u : 1u) << i;
      }
      bases[wo + w] = packed;
      nmask[wo + w] = bad;
      if (bad) {
        any_bad = 1;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && any_bad) {
    meta[r].non_acgt = 1;
  }
}

// one thread per record: strictly left-to-right double sums (calc_phred_average.cpp:15-30).  The
// additions of one read are a dependent chain by definition (the sum must round as the
// reference's does); what can be saved is around them: the 10^(-q/10) table sits in shared memory
// (a data-dependent index into __constant__ memory serialises the 32 lanes of a warp) and the
// quality bytes arrive 16 at a time.
__global__ void __launch_bounds__(64)
k_phred(const uint8_t* __restrict__ buf, uint64_t chunk_base, grb_read_meta* __restrict__ meta,
        uint64_t n_rec)
{
  __shared__ double tab[256];
  for (unsigned i = threadIdx.x; i < 256; i += blockDim.x) {
    tab[i] = c_delog[i];
  }
  __syncthreads();
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) {
    return;
  }
  const uint32_t n = meta[r].qual_len;
  const uint8_t* q = buf + (meta[r].qual_off - chunk_base);
  const uint32_t mark = n / 2 - 1; // n/2 - 1 in size_t never equals i < n when n < 2
  const bool has_mark = n >= 2;
  double total = 0.0, first = 0.0;
  uint32_t i = 0;
  auto step = [&](uint32_t byte) {
    total += tab[byte];
    if (has_mark && i == mark) {
      first = total;
    }
    ++i;
  };
  const uint32_t head = (uint32_t)((16 - ((uintptr_t)q & 15)) & 15);
  while (i < n && i < head) {
    step(__ldg(q + i));
  }
  while (i + 16 <= n) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(q + i));
    const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      step((w[j >> 2] >> ((j & 3) * 8)) & 0xFFu);
    }
  }
  while (i < n) {
    step(__ldg(q + i));
  }
  meta[r].phred_first_half_sum = first;
  meta[r].phred_total_sum = total;
}

// Host -> device copy done by a few CTAs reading mapped pinned host memory over PCIe, used for the
// ingest read-ahead: it leaves the copy engine free, so the small descriptor uploads of the decode
// and pass-1 launches that run meanwhile are not queued behind a 1 GB transfer (measured: with
// cudaMemcpyAsync for the read-ahead, pass 1 under the ingest overlapped nothing).
// src and dst must have the same alignment modulo 16.
__global__ void __launch_bounds__(256)
k_copy_host(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint64_t n)
{
  const uint64_t mis = (16 - ((uintptr_t)src & 15)) & 15;
  const uint64_t head = mis < n ? mis : n;
  const uint64_t body = (n - head) / 16;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nth = (uint64_t)gridDim.x * blockDim.x;
  if (tid < head) {
    dst[tid] = src[tid];
  }
  const uint64_t tail0 = head + body * 16;
  if (tid < n - tail0) {
    dst[tail0 + tid] = src[tail0 + tid];
  }
  const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
  uint4* d4 = reinterpret_cast<uint4*>(dst + head);
  uint64_t i = tid;
  for (; i + 3 * nth < body; i += 4 * nth) { // four 16-byte PCIe reads in flight per thread
    const uint4 a = __ldcs(s4 + i), b = __ldcs(s4 + i + nth), c = __ldcs(s4 + i + 2 * nth),
                d = __ldcs(s4 + i + 3 * nth);
    d4[i] = a;
    d4[i + nth] = b;
    d4[i + 2 * nth] = c;
    d4[i + 3 * nth] = d;
  }
  for (; i < body; i += nth) {
    d4[i] = __ldcs(s4 + i);
  }
}

