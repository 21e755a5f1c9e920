// What bounds a random miBF probe on one B200?  Stand-alone measurement tool (not part of the
// product path): random-access rates of the access shapes the engine uses or could use, over
// footprints from L2-resident to most of HBM.
//
//   gather32x2  one random 32-byte sector per access, read as two 16-byte loads (LDG.128 x 2):
//               what grb_probe_block did in round 1
//   gather32    the same sector with ONE 32-byte load (LDG.256, new with sm_100)
//   gather16 / gather8 / gather4   narrower loads of a random sector (the ID slot read is 4-8 bytes)
//   line128     one random 128-byte line per access, read by 8 lanes x 16 bytes (coalesced)
//   chain       the product's probe: a random 32-byte block (LDG.256), then a dependent random
//               8-byte slot read whose address comes from the block (h = 3 chains per thread)
//   rmw8        random 8-byte read-modify-write (the reservoir insert of k3_bulk)
// U = independent accesses in flight per thread, T = resident threads per SM.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/sector-roofline tools/sector_roofline.cu
//   build/sector-roofline [footprint_GiB ...]        one JSON object per line on stdout
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x)                                                                                      \
  do {                                                                                             \
    cudaError_t e = (x);                                                                           \
    if (e != cudaSuccess) {                                                                        \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                                      \
      exit(1);                                                                                     \
    }                                                                                              \
  } while (0)

__device__ __forceinline__ uint64_t
splitmix(uint64_t x)
{
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

__device__ __forceinline__ uint64_t
pick(uint64_t x, uint64_t n)
{
  return __umul64hi(x, n);
}

__device__ __forceinline__ uint64_t
ld256(const void* p)
{
  uint64_t a, b, c, d;
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  return a ^ b ^ c ^ d;
}

enum Shape { G32X2 = 0, G32 = 1, G16 = 2, G8 = 3, G4 = 4 };

template<int SHAPE, int U>
__global__ void __launch_bounds__(256)
k_gather(const uint8_t* __restrict__ mem, uint64_t n_sectors, uint32_t per_thread, uint64_t seed,
         unsigned long long* sink)
{
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long acc = 0;
  uint64_t x = splitmix(seed ^ tid);
  for (uint32_t i = 0; i < per_thread; i += U) {
    uint64_t v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      x = splitmix(x);
      const uint8_t* p = mem + 32 * pick(x, n_sectors);
      if (SHAPE == G32X2) {
        const ulonglong2 a = __ldg((const ulonglong2*)p), b = __ldg((const ulonglong2*)p + 1);
        v[u] = a.x ^ a.y ^ b.x ^ b.y;
      } else if (SHAPE == G32) {
        v[u] = ld256(p);
      } else if (SHAPE == G16) {
        const ulonglong2 a = __ldg((const ulonglong2*)p);
        v[u] = a.x ^ a.y;
      } else if (SHAPE == G8) {
        v[u] = __ldg((const uint64_t*)p);
      } else {
        v[u] = __ldg((const uint32_t*)p);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      acc += v[u];
    }
  }
  if (acc == 0x1234567ull) {
    *sink = acc;
  }
}

// 8 lanes share one random 128-byte line
template<int U>
__global__ void __launch_bounds__(256)
k_line128(const uint8_t* __restrict__ mem, uint64_t n_lines, uint32_t per_group, uint64_t seed,
          unsigned long long* sink)
{
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t grp = tid >> 3;
  const unsigned sub = threadIdx.x & 7;
  unsigned long long acc = 0;
  uint64_t x = splitmix(seed ^ grp);
  for (uint32_t i = 0; i < per_group; i += U) {
    ulonglong2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      x = splitmix(x);
      v[u] = __ldg((const ulonglong2*)(mem + 128 * pick(x, n_lines)) + sub);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      acc += v[u].x ^ v[u].y;
    }
  }
  if (acc == 0x1234567ull) {
    *sink = acc;
  }
}

// the product's probe: random 32-byte block, then a dependent random 8-byte slot; H chains per thread
template<int H, bool WIDE>
__global__ void __launch_bounds__(256)
k_chain(const uint8_t* __restrict__ blocks, uint64_t n_blocks, const uint8_t* __restrict__ slots,
        uint64_t n_slots, uint32_t per_thread, uint64_t seed, unsigned long long* sink)
{
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long acc = 0;
  uint64_t x = splitmix(seed ^ tid);
  for (uint32_t i = 0; i < per_thread; ++i) {
    uint64_t b[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
      x = splitmix(x);
      const uint8_t* p = blocks + 32 * pick(x, n_blocks);
      if (WIDE) {
        b[j] = ld256(p);
      } else {
        const ulonglong2 a = __ldg((const ulonglong2*)p), c = __ldg((const ulonglong2*)p + 1);
        b[j] = a.x ^ a.y ^ c.x ^ c.y;
      }
    }
#pragma unroll
    for (int j = 0; j < H; ++j) {
      acc += __ldg((const uint64_t*)(slots + 8 * pick(splitmix(b[j] ^ x), n_slots)));
    }
  }
  if (acc == 0x1234567ull) {
    *sink = acc;
  }
}

template<int U>
__global__ void __launch_bounds__(256)
k_rmw8(uint64_t* __restrict__ mem, uint64_t n_slots, uint32_t per_thread, uint64_t seed)
{
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t x = splitmix(seed ^ tid);
  for (uint32_t i = 0; i < per_thread; i += U) {
    uint64_t s[U], v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      x = splitmix(x);
      s[u] = pick(x, n_slots);
      v[u] = __ldcg(mem + s[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      mem[s[u]] = v[u] + (x | 1);
    }
  }
}

struct Timer
{
  cudaEvent_t e0, e1;
  Timer()
  {
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
  }
  template<class F>
  float best_ms(F launch, int reps = 3)
  {
    float best = 1e30f;
    for (int r = 0; r <= reps; ++r) {
      CK(cudaEventRecord(e0));
      launch(r);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r > 0 && ms < best) {
        best = ms;
      }
    }
    CK(cudaGetLastError());
    return best;
  }
};

static void
row(const char* op, double gib, int threads_per_sm, int u, double accesses, float ms, double bytes_each)
{
  printf("{\"op\": \"%s\", \"footprint_gib\": %.3f, \"threads_per_sm\": %d, \"in_flight_per_thread\": %d, "
         "\"accesses\": %.0f, \"ms\": %.4f, \"gaccess_per_s\": %.3f, \"useful_gb_per_s\": %.1f}\n",
         op, gib, threads_per_sm, u, accesses, ms, accesses / (ms * 1e-3) / 1e9,
         accesses * bytes_each / (ms * 1e-3) / 1e9);
  fflush(stdout);
}

int
main(int argc, char** argv)
{
  std::vector<double> gib;
  for (int i = 1; i < argc; ++i) {
    gib.push_back(atof(argv[i]));
  }
  if (gib.empty()) {
    gib = { 0.0625, 1, 4, 22, 64 };
  }
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));
  Timer tm;
  for (double g : gib) {
    const uint64_t bytes = (uint64_t)(g * (1ull << 30)) / 128 * 128;
    uint8_t* mem = nullptr;
    if (cudaMalloc(&mem, bytes) != cudaSuccess) {
      cudaGetLastError();
      fprintf(stderr, "skip %.1f GiB: allocation failed\n", g);
      continue;
    }
    CK(cudaMemset(mem, 1, bytes));
    for (int tps : { 1024, 2048 }) {
      const unsigned grid = (unsigned)(sms * tps / 256);
      const double threads = (double)grid * 256;
      const uint32_t per = 256;
      const uint64_t nsec = bytes / 32;
#define GATHER(NAME, SHAPE, U, BYTES)                                                              \
  row(NAME, g, tps, U, threads * per,                                                              \
      tm.best_ms([&](int r) { k_gather<SHAPE, U><<<grid, 256>>>(mem, nsec, per, 42 + r, sink); }), BYTES)
      GATHER("gather32x2", G32X2, 1, 32);
      GATHER("gather32x2", G32X2, 4, 32);
      GATHER("gather32", G32, 1, 32);
      GATHER("gather32", G32, 2, 32);
      GATHER("gather32", G32, 4, 32);
      GATHER("gather32", G32, 8, 32);
      GATHER("gather16", G16, 4, 16);
      GATHER("gather8", G8, 4, 8);
      GATHER("gather4", G4, 4, 4);
      GATHER("gather4", G4, 8, 4);
#undef GATHER
      row("line128", g, tps, 4, threads / 8 * per,
          tm.best_ms([&](int r) { k_line128<4><<<grid, 256>>>(mem, bytes / 128, per, 42 + r, sink); }), 128);
      // chain: the first sixth of the memory plays the filter blocks, the rest the ID slots
      // (cfg2: 0.47 GB of blocks beside 10.7 GB of 8-byte slots)
      const uint64_t nblk = bytes / 6 / 32, nslot = (bytes - nblk * 32) / 8;
      const uint8_t* slots = mem + nblk * 32;
      row("chain_h3_2x128", g, tps, 3, threads * 64 * 3,
          tm.best_ms([&](int r) { k_chain<3, false><<<grid, 256>>>(mem, nblk, slots, nslot, 64, 42 + r, sink); }), 40);
      row("chain_h3_256", g, tps, 3, threads * 64 * 3,
          tm.best_ms([&](int r) { k_chain<3, true><<<grid, 256>>>(mem, nblk, slots, nslot, 64, 42 + r, sink); }), 40);
      row("chain_h1_256", g, tps, 1, threads * 128,
          tm.best_ms([&](int r) { k_chain<1, true><<<grid, 256>>>(mem, nblk, slots, nslot, 128, 42 + r, sink); }), 40);
      row("rmw8", g, tps, 1, threads * per,
          tm.best_ms([&](int r) { k_rmw8<1><<<grid, 256>>>((uint64_t*)mem, bytes / 8, per, 42 + r); }), 16);
      row("rmw8", g, tps, 4, threads * per,
          tm.best_ms([&](int r) { k_rmw8<4><<<grid, 256>>>((uint64_t*)mem, bytes / 8, per, 42 + r); }), 16);
    }
    CK(cudaFree(mem));
  }
  return 0;
}
