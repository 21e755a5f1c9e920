// Random 32-byte-sector gather / read-modify-write roofline of one B200 (the
// denominator the miBF probe can at best reach), as a function of the footprint and of
// cudaLimitMaxL2FetchGranularity.  Stand-alone measurement tool, not part of the product path.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/sector-roofline tools/sector_roofline.cu
//   build/sector-roofline [footprint_GiB ...]
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

// every thread gathers `per_thread` random 32-byte sectors (two 16-byte loads of one sector)
__global__ void
k_gather(const ulonglong2* __restrict__ mem, uint64_t n_sectors, uint64_t per_thread, uint64_t seed,
         unsigned long long* sink)
{
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long acc = 0;
  uint64_t x = splitmix(seed ^ tid);
  for (uint64_t i = 0; i < per_thread; i += 4) {
    ulonglong2 v[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x = splitmix(x);
      const uint64_t s = (uint64_t)(((unsigned __int128)x * n_sectors) >> 64);
      v[u][0] = __ldg(mem + 2 * s);
      v[u][1] = __ldg(mem + 2 * s + 1);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc += v[u][0].x ^ v[u][0].y ^ v[u][1].x ^ v[u][1].y;
    }
  }
  if (acc == 0x1234567ull) {
    *sink = acc;
  }
}

// random 16-byte read-modify-write (the ID-slot insert): one sector read + one sector write
__global__ void
k_rmw(ulonglong2* __restrict__ mem, uint64_t n_slots, uint64_t per_thread, uint64_t seed)
{
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t x = splitmix(seed ^ tid);
  for (uint64_t i = 0; i < per_thread; i += 4) {
    uint64_t s[4];
    ulonglong2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x = splitmix(x);
      s[u] = (uint64_t)(((unsigned __int128)x * n_slots) >> 64);
      v[u] = mem[s[u]];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u].x += 1;
      v[u].y ^= x;
      mem[s[u]] = v[u];
    }
  }
}

// random 64-bit atomicOr (the pass-1 bit fill) / atomicCAS (the batch rank index)
__global__ void
k_atom(unsigned long long* __restrict__ mem, uint64_t n_words, uint64_t per_thread, uint64_t seed,
       int cas)
{
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t x = splitmix(seed ^ tid);
  for (uint64_t i = 0; i < per_thread; ++i) {
    x = splitmix(x);
    const uint64_t s = (uint64_t)(((unsigned __int128)x * n_words) >> 64);
    if (cas) {
      atomicCAS(mem + s, 0x0101010101010101ull, x);
    } else {
      atomicOr(mem + s, 1ull << (x & 63));
    }
  }
}

int
main(int argc, char** argv)
{
  std::vector<double> gib;
  for (int i = 1; i < argc; ++i) {
    gib.push_back(atof(argv[i]));
  }
  if (gib.empty()) {
    gib = { 0.0625, 1, 4, 16, 64, 128 };
  }
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("{\"device_sms\": %d, \"rows\": [\n", sms);
  bool first = true;
  for (size_t gran : { (size_t)0, (size_t)32, (size_t)64, (size_t)128 }) {
    if (gran) {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
      if (e != cudaSuccess) {
        fprintf(stderr, "set granularity %zu: %s\n", gran, cudaGetErrorString(e));
        cudaGetLastError();
        continue;
      }
    }
    size_t got = 0;
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    for (double g : gib) {
      const uint64_t bytes = (uint64_t)(g * (1ull << 30)) / 32 * 32;
      void* mem = nullptr;
      if (cudaMalloc(&mem, bytes) != cudaSuccess) {
        cudaGetLastError();
        continue;
      }
      CK(cudaMemset(mem, 1, bytes));
      const uint64_t threads = (uint64_t)sms * 2048;
      const uint64_t per_thread = 256;
      const unsigned grid = (unsigned)(threads / 256);
      for (int mode = 0; mode < 4; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          CK(cudaEventRecord(e0));
          if (mode == 0) {
            k_gather<<<grid, 256>>>((const ulonglong2*)mem, bytes / 32, per_thread, 42 + rep, sink);
          } else if (mode == 1) {
            k_rmw<<<grid, 256>>>((ulonglong2*)mem, bytes / 16, per_thread, 42 + rep);
          } else {
            k_atom<<<grid, 256>>>((unsigned long long*)mem, bytes / 8, per_thread, 42 + rep, mode == 3);
          }
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (rep > 0 && ms < best) {
            best = ms;
          }
        }
        const double acc = (double)threads * per_thread;
        const double gbs = acc * (mode == 0 ? 32.0 : 64.0) / (best * 1e-3) / 1e9;
        printf("%s{\"granularity_limit\": %zu, \"footprint_gib\": %.4f, \"op\": \"%s\", "
               "\"accesses\": %.0f, \"ms\": %.4f, \"gacc_per_s\": %.3f, \"sector_gb_per_s\": %.1f}",
               first ? "" : ",\n", got, g, mode == 0 ? "gather32" : (mode == 1 ? "rmw16" : (mode == 2 ? "atomic_or64" : "atomic_cas64")), acc, best,
               acc / (best * 1e-3) / 1e9, gbs);
        first = false;
      }
      CK(cudaFree(mem));
    }
  }
  printf("\n]}\n");
  return 0;
}
