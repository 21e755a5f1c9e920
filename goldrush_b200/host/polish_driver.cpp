// (f4) GoldPolish targeted Bloom filters, host side: what serve_batch decides per target before any
// k-mer is hashed (subprojects/goldpolish/src/goldpolish_targeted_bfs.cpp:43-51,88-127), and a test
// hook that runs the device's job code (csrc/polish_core.h) on the host.
#include "goldrush_b200.h"

#include "../csrc/polish_core.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <tuple>
#include <vector>

extern "C" {

int
grb_polish_kmer_threshold(uint64_t mappings_bases)
{
  static const double a = 4.66943;
  static const double b = 2.11391e-07;
  static const int max_kmer_threshold = 13;
  const int kmer_threshold = int(std::round(a + double(mappings_bases) * b));
  return std::min(kmer_threshold, max_kmer_threshold);
}

int
grb_polish_plan_target(uint64_t target_len, double subsample_max_per_10kbp, uint32_t n_mappings,
                       const char* const* ids, const double* phred_avg, const uint64_t* lens,
                       uint32_t* order_out, uint32_t* n_used, int32_t* kmer_threshold)
{
  // :94-98: the cap is computed in double and cast to the (unsigned) size type
  const size_t num_max = size_t(double(target_len) * subsample_max_per_10kbp / 10000.0);
  const size_t adjusted = std::min<size_t>(n_mappings, num_max);
  struct Key
  {
    size_t phred; // :100-104: the tuple stores the Phred average as size_t
    const char* id;
    uint32_t index;
  };
  std::vector<Key> v(n_mappings);
  for (uint32_t i = 0; i < n_mappings; ++i) {
    v[i] = Key{ size_t(phred_avg[i]), ids[i], i };
  }
  std::sort(v.begin(), v.end(), [](const Key& x, const Key& y) { // :106-113
    return x.phred > y.phred || (x.phred == y.phred && strcmp(x.id, y.id) < 0);
  });
  uint64_t bases = 0;
  for (uint32_t i = 0; i < n_mappings; ++i) {
    order_out[i] = v[i].index;
    if (i < adjusted) {
      bases += lens[v[i].index]; // :117-123
    }
  }
  *n_used = (uint32_t)adjusted;
  *kmer_threshold = grb_polish_kmer_threshold(bases);
  return *kmer_threshold > 0 ? GRB_OK : GRB_ERR_ARG; // :126-127
}

int
grb_test_polish_fill_host(const grb_polish_params* p, uint32_t n_batches, const uint64_t* batch_first,
                          const char* seqs, const uint64_t* seq_off, const uint32_t* thresholds,
                          uint8_t* out_bfs)
{
  if (p->cbf_bytes < 2 || p->bf_bytes < 1) {
    return GRB_ERR_ARG;
  }
  std::vector<uint8_t> cbf(p->cbf_bytes);
  int rc = GRB_OK;
  for (uint32_t b = 0; b < n_batches; ++b) {
    for (uint32_t i = 0; i < p->n_k; ++i) {
      std::fill(cbf.begin(), cbf.end(), 0);
      uint8_t* bf = out_bfs + ((size_t)b * p->n_k + i) * p->bf_bytes;
      memset(bf, 0, p->bf_bytes);
      GrbPolishJob j;
      j.seqs = seqs;
      j.off = seq_off;
      j.thr = thresholds;
      j.first = batch_first[b];
      j.last = batch_first[b + 1];
      j.k = p->k_values[i];
      j.k_index = i;
      j.hash_num = p->hash_num;
      j.cbf = cbf.data();
      j.cbf_bytes = p->cbf_bytes;
      j.cbf_inv = (uint64_t)(((unsigned __int128)1 << 64) / p->cbf_bytes);
      j.bf = bf;
      j.bf_bits = p->bf_bytes * 8;
      j.bf_inv = (uint64_t)(((unsigned __int128)1 << 64) / j.bf_bits);
      if (grb_polish_run(j) != 0) {
        rc = GRB_ERR_ARG;
      }
    }
  }
  return rc;
}

} // extern "C"
