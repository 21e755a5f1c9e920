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

// The warp kernel's algorithm (csrc/kernels_polish.cuh) lane by lane on the host: segments packed
// with grb_p_pack32, k-mers hashed with grb_p_hash_packed through the per-k tables, groups of 32
// k-mers applied "at once" -- every lane's counts read BEFORE any lane writes -- when no counter is
// shared between two lanes, in lane order otherwise.  What the CPU suite checks against the port:
// the packing, the table hashing and the claim that conflict-free groups commute.
int
grb_test_polish_fill_host_grouped(const grb_polish_params* p, uint32_t n_batches, const uint64_t* batch_first,
                                  const char* seqs, const uint64_t* seq_off, const uint32_t* thresholds,
                                  uint8_t* out_bfs, uint64_t* groups_total, uint64_t* groups_in_lane_order)
{
  if (p->cbf_bytes < 2 || p->bf_bytes < 1 || p->hash_num == 0 || p->hash_num > 8) {
    return GRB_ERR_ARG;
  }
  const unsigned SEG = 1024, WORDS = (SEG + GRB_P_MAX_K) / 32 + 3;
  const unsigned h = p->hash_num;
  const uint64_t cbf_inv = (uint64_t)(((unsigned __int128)1 << 64) / p->cbf_bytes);
  const uint64_t bf_bits = p->bf_bytes * 8, bf_inv = (uint64_t)(((unsigned __int128)1 << 64) / bf_bits);
  std::vector<uint8_t> cbf(p->cbf_bytes);
  std::vector<GrbPolishPair> T((size_t)GRB_P_GROUPS * 256);
  std::vector<uint64_t> codes(WORDS);
  std::vector<uint32_t> bad(WORDS);
  uint64_t n_groups = 0, n_ordered = 0;
  for (uint32_t b = 0; b < n_batches; ++b) {
    for (uint32_t ki = 0; ki < p->n_k; ++ki) {
      const unsigned k = p->k_values[ki];
      if (k == 0 || k > GRB_P_MAX_K) {
        return GRB_ERR_ARG;
      }
      grb_p_build_table(k, T.data());
      std::fill(cbf.begin(), cbf.end(), 0);
      uint8_t* bf = out_bfs + ((size_t)b * p->n_k + ki) * p->bf_bytes;
      memset(bf, 0, p->bf_bytes);
      for (uint64_t r = batch_first[b]; r < batch_first[b + 1]; ++r) {
        if (thresholds[r] < 4) {
          return GRB_ERR_ARG;
        }
        const unsigned thr = thresholds[r] - 2 + ki, thr8 = thr > 255 ? 255 : thr;
        const char* seq = seqs + seq_off[r];
        const uint64_t len = seq_off[r + 1] - seq_off[r];
        if (len < k) {
          continue;
        }
        const uint64_t n_pos = len - k + 1;
        for (uint64_t seg0 = 0; seg0 < n_pos; seg0 += SEG) {
          const uint64_t seg_pos = std::min<uint64_t>(SEG, n_pos - seg0), seg_bases = seg_pos + k - 1;
          for (unsigned wd = 0; wd < WORDS; ++wd) {
            const uint64_t b0 = (uint64_t)wd * 32;
            codes[wd] = 0;
            bad[wd] = 0;
            if (b0 < seg_bases) {
              grb_p_pack32(seq + seg0 + b0, (unsigned)std::min<uint64_t>(32, seg_bases - b0), &codes[wd], &bad[wd]);
            }
          }
          for (uint64_t g0 = 0; g0 < seg_pos; g0 += 32) {
            bool valid[32];
            uint64_t idx[32][8], at[32][8];
            bool conflict = false;
            std::vector<std::pair<uint64_t, unsigned>> seen;
            for (unsigned lane = 0; lane < 32; ++lane) {
              const uint64_t q = g0 + lane;
              uint64_t base = 0;
              valid[lane] = q < seg_pos && grb_p_hash_packed(codes.data(), bad.data(), (unsigned)q, k, T.data(), &base);
              if (!valid[lane]) {
                continue;
              }
              idx[lane][0] = base;
              for (unsigned u = 1; u < h; ++u) {
                uint64_t x = base * (u ^ k * 0x90b45d39fb6da1faULL);
                x ^= x >> 27;
                idx[lane][u] = x;
              }
              for (unsigned u = 0; u < h; ++u) {
                at[lane][u] = grb_p_mod(idx[lane][u], p->cbf_bytes, cbf_inv);
                for (const auto& e : seen) {
                  conflict = conflict || (e.first == at[lane][u] && e.second != lane);
                }
                seen.emplace_back(at[lane][u], lane);
              }
            }
            ++n_groups;
            auto counts_of = [&](unsigned lane) {
              uint8_t c = 255;
              for (unsigned u = 0; u < h; ++u) {
                c = std::min(c, cbf[at[lane][u]]);
              }
              return c;
            };
            auto apply = [&](unsigned lane, uint8_t count) {
              unsigned after = count;
              if (count < thr8) {
                for (unsigned u = 0; u < h; ++u) {
                  if (cbf[at[lane][u]] == count) {
                    cbf[at[lane][u]] = (uint8_t)(count + 1);
                  }
                }
                after = count + 1u;
              }
              if (after >= thr) {
                for (unsigned u = 0; u < h; ++u) {
                  const uint64_t pos = grb_p_mod(idx[lane][u], bf_bits, bf_inv);
                  bf[pos >> 3] |= (uint8_t)(1u << (pos & 7));
                }
              }
            };
            if (!conflict) { // "at once": all reads, then all writes
              uint8_t cnt[32];
              for (unsigned lane = 0; lane < 32; ++lane) {
                cnt[lane] = valid[lane] ? counts_of(lane) : 0;
              }
              for (unsigned lane = 0; lane < 32; ++lane) {
                if (valid[lane]) {
                  apply(lane, cnt[lane]);
                }
              }
            } else {
              ++n_ordered;
              for (unsigned lane = 0; lane < 32; ++lane) {
                if (valid[lane]) {
                  apply(lane, counts_of(lane));
                }
              }
            }
          }
        }
      }
    }
  }
  if (groups_total) {
    *groups_total = n_groups;
  }
  if (groups_in_lane_order) {
    *groups_in_lane_order = n_ordered;
  }
  return GRB_OK;
}

} // extern "C"
