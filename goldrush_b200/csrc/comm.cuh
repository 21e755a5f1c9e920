// Multi-GPU exchange steps of the GoldRush-Path engine: one process per GPU, the
// filter replicated, NCCL over NVLink for the two places where the path really exchanges data:
//
//   pass 1  (goldrush_path.cpp:235-339, MIBFConstructSupport.hpp:134-147): reads are independent
//           and the bit OR commutes, so rank r fills its own zeroed bit vector from its share of the
//           reads and the vectors are OR-reduced.  NCCL has no bitwise-OR reduction: the vector is
//           all-gathered slice by slice and k_or_gathered folds the W-1 foreign slices in.
//   pass 2  (goldrush_path.cpp:544-626): the speculative query of one batch is read-only against
//           the filter as it stood at batch start, so rank r queries tiles [r*chunk, (r+1)*chunk) of
//           the batch and the per-tile results (rank stash, vote tables, arg-max, hit counters) are
//           all-gathered in one NCCL group.  The ordered commit that follows is replicated
//           (integer-only, deterministic), so every replica ends the batch with the same filter.
//
// libnccl is bound at run time (dlopen of the copy already loaded by torch, else the system one),
// so the library has no link-time dependency on it and single-GPU users never touch it.
#pragma once
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include <string>

struct GrbNccl
{
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct GrbComm
{
  std::mutex mu;
  GrbNccl api;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = -1;
  std::string err;
};

inline GrbComm&
grb_comm()
{
  static GrbComm* g = new GrbComm; // leaked on purpose: no NCCL / CUDA calls at static destruction
  return *g;
}

// binds the ten NCCL entry points; returns false and sets g.err if no libnccl can be found
inline bool
grb_nccl_load(GrbComm& g)
{
  if (g.api.lib) {
    return true;
  }
  const char* names[] = { getenv("GRB_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
  void* lib = nullptr;
  for (const char* n : names) {
    if (n && *n && (lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) {
      break;
    }
  }
  if (!lib) {
    g.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "dlopen failed");
    return false;
  }
  GrbNccl a;
  a.lib = lib;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(lib, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(lib, "ncclCommDestroy");
  a.AllGather = (decltype(a.AllGather))dlsym(lib, "ncclAllGather");
  a.Broadcast = (decltype(a.Broadcast))dlsym(lib, "ncclBroadcast");
  a.Send = (decltype(a.Send))dlsym(lib, "ncclSend");
  a.Recv = (decltype(a.Recv))dlsym(lib, "ncclRecv");
  a.GroupStart = (decltype(a.GroupStart))dlsym(lib, "ncclGroupStart");
  a.GroupEnd = (decltype(a.GroupEnd))dlsym(lib, "ncclGroupEnd");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.Broadcast || !a.Send ||
      !a.Recv ||
      !a.GroupStart ||
      !a.GroupEnd || !a.GetErrorString) {
    g.err = "libnccl lacks one of the entry points this library binds";
    dlclose(lib);
    return false;
  }
  g.api = a;
  return true;
}

// Share of `n` items of rank r: equal chunks of ceil(n / world) so that an all-gather is uniform;
// the last ranks may hold a short or empty share.
struct GrbShare
{
  uint64_t chunk, lo, hi;
};
__host__ __device__ inline GrbShare
grb_share(uint64_t n, int rank, int world)
{
  GrbShare s;
  s.chunk = (n + (uint64_t)world - 1) / (uint64_t)world;
  s.lo = (uint64_t)rank * s.chunk < n ? (uint64_t)rank * s.chunk : n;
  s.hi = s.lo + s.chunk < n ? s.lo + s.chunk : n;
  return s;
}

// dst[i] |= OR over r != rank of gathered[r * stride + i], i < n  (8-byte words)
__global__ void
k_or_gathered(uint64_t* __restrict__ dst, const uint64_t* __restrict__ gathered, uint64_t stride,
              uint64_t n, int rank, int world)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t v = dst[i];
    for (int r = 0; r < world; ++r) {
      if (r != rank) {
        v |= __ldcs(&gathered[(uint64_t)r * stride + i]);
      }
    }
    dst[i] = v;
  }
}
