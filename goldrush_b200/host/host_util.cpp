// Host-side scalar pieces of GoldRush-Path that stay on the CPU because their exact value depends
// on glibc (rand, pow, log, log10) or on C++ arithmetic conversions of the reference.
#include "goldrush_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern "C" {

// goldrush_path/spaced_seeds.cpp:7-68.  Random design: srand(123); the left half has k/2 symbols,
// the first is 1, the others are redrawn as rand()%2 until exactly weight/2 are set; the right half
// mirrors it.  Preset: the two halves of the given string.  Pattern i puts i zeros between them.
int
grb_make_seed_pattern(const char* preset, unsigned k, unsigned weight, unsigned h, char** out)
{
  std::string left, right;
  if (preset == nullptr || preset[0] == '\0') {
    const unsigned half = k / 2;
    if (half == 0 || weight / 2 > half) {
      return GRB_ERR_ARG; // the reference would never leave its redraw loop
    }
    srand(123);
    std::vector<unsigned char> bits(half, 0);
    bits[0] = 1;
    unsigned ones = 0;
    while (ones != weight / 2) {
      ones = 1;
      for (unsigned i = 1; i < half; ++i) {
        bits[i] = (unsigned char)(rand() % 2);
        ones += bits[i];
      }
    }
    for (unsigned i = 0; i < half; ++i) {
      left.push_back(bits[i] ? '1' : '0');
    }
    right.assign(left.rbegin(), left.rend());
  } else {
    const std::string p(preset);
    left = p.substr(0, p.size() / 2);
    right = p.substr(p.size() / 2, p.size() / 2);
  }
  for (unsigned i = 0; i < h; ++i) {
    const std::string s = left + std::string(i, '0') + right;
    memcpy(out[i], s.c_str(), s.size() + 1);
  }
  return GRB_OK;
}

// goldrush_path/MIBloomFilter.hpp:94-101: always rounds UP to the next multiple of 64, even when
// the value already is one.
uint64_t
grb_calc_optimal_size(uint64_t entries, unsigned hash_num, double occupancy)
{
  const size_t approx = size_t(-double(entries) * double(hash_num) / log(1.0 - occupancy));
  return approx + (64 - approx % 64);
}

// goldrush_path/goldrush_path.cpp:1114-1121.  The 0.5 coefficient is a float in the reference, so
// the product is evaluated in single precision before the conversion to uint64.
uint64_t
grb_default_hash_universe(uint64_t weight, uint64_t genome_size, uint64_t hash_num)
{
  const size_t base = std::min((uint64_t)(pow((uint8_t)4, weight)), (uint64_t)2 * genome_size);
  const float coefficient = 0.5f;
  return (uint64_t)(base * coefficient * hash_num);
}

// goldrush_path/calc_phred_average.cpp:32-42 on the two running sums.
void
grb_phred_finalize(double first_half_sum, double total_sum, uint64_t n, uint32_t* avg,
                   uint32_t* delta)
{
  const size_t qual_size = (size_t)n;
  double second_avg = total_sum - first_half_sum;
  second_avg = second_avg / (qual_size * 0.5);
  const double first_avg = first_half_sum / (qual_size * 0.5);
  *avg = (uint32_t)(-10 * log10(total_sum / qual_size));
  *delta = (uint32_t)abs((int32_t)(-10 * log10(first_avg)) - (int32_t)(-10 * log10(second_avg)));
}

void
grb_abi_sizes(uint64_t* out9)
{
  out9[0] = sizeof(grb_params);
  out9[1] = sizeof(grb_read_meta);
  out9[2] = sizeof(grb_decision);
  out9[3] = sizeof(grb_path_stats);
  out9[4] = sizeof(grb_probe_bench_result);
  out9[5] = sizeof(grb_run_options);
  out9[6] = sizeof(grb_run_result);
  out9[7] = sizeof(grb_synth_params);
  out9[8] = sizeof(grb_host_msg);
}

// the same for n reads at once (host threads), from the metadata K1 returned
void
grb_phred_finalize_batch(const grb_read_meta* meta, uint64_t n, uint32_t* avg, uint32_t* delta)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    grb_phred_finalize(meta[i].phred_first_half_sum, meta[i].phred_total_sum, meta[i].qual_len,
                       &avg[i], &delta[i]);
  }
}

} // extern "C"
