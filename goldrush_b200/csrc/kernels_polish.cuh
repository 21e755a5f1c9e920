// (f4) GoldPolish targeted Bloom filters on the device: one thread per (batch, k) job of
// polish_core.h, thousands of jobs per launch.  See polish_core.h for what it replaces and why a
// job is sequential.
#pragma once
#include "polish_core.h"
#include <cuda_runtime.h>

struct GrbPolishWave
{
  const char* seqs;
  const uint64_t* off;
  const uint32_t* thr;
  const uint64_t* batch_first; // [n_batches + 1]
  const uint32_t* k_values;    // [n_k]
  uint32_t n_k, hash_num;
  uint32_t job0, n_jobs;       // jobs [job0, job0 + n_jobs) of the call: job = batch * n_k + k_index
  uint8_t* cbf;                // n_jobs counting filters of cbf_bytes each, zeroed
  uint8_t* bf;                 // n_jobs Bloom filters of bf_bytes each, zeroed
  uint64_t cbf_bytes, cbf_inv, bf_bytes, bf_inv;
  int* status;
};

__global__ void __launch_bounds__(32)
k_polish_fill(GrbPolishWave w)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= w.n_jobs) {
    return;
  }
  const uint32_t job = w.job0 + t;
  const uint32_t batch = job / w.n_k, ki = job - batch * w.n_k;
  GrbPolishJob j;
  j.seqs = w.seqs;
  j.off = w.off;
  j.thr = w.thr;
  j.first = w.batch_first[batch];
  j.last = w.batch_first[batch + 1];
  j.k = w.k_values[ki];
  j.k_index = ki;
  j.hash_num = w.hash_num;
  j.cbf = w.cbf + (uint64_t)t * w.cbf_bytes;
  j.cbf_bytes = w.cbf_bytes;
  j.cbf_inv = w.cbf_inv;
  j.bf = w.bf + (uint64_t)t * w.bf_bytes;
  j.bf_bits = w.bf_bytes * 8;
  j.bf_inv = w.bf_inv;
  if (grb_polish_run(j) != 0) {
    atomicExch(w.status, -1);
  }
}
