// Deterministic synthetic long-read generator (SURVEY.md §8d shapes). Counter-based: the genome is a
// pure function of (seed, position) and every read has its own RNG stream keyed by (seed, index),
// so the output is identical for any thread count and no genome array is ever materialised.
//
//   genome      iid uniform ACGT, upper case
//   reads       start uniform, strand Bernoulli(0.5), file order = generation order
//   lengths     fixed `read_len`, or (read_len == 0) log-normal sigma=0.5 rescaled to N50 = `n50`,
//               clipped to [1000, 200000]
//   errors      per-base substitution / insertion / deletion (uniform replacement base)
//   qualities   per-read Q ~ UniformInt[qmin, qmax), per-base clamp(Q + round(N(0,2)), 2, 40), +33
//   header      "@read<i> pos=<start>"   (exercises the id / comment split)
#include "goldrush_b200.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#if _OPENMP
#include <omp.h>
#endif

namespace {

inline uint64_t
splitmix64(uint64_t x)
{
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}

struct Rng
{
  uint64_t s;
  explicit Rng(uint64_t seed)
    : s(seed)
  {}
  inline uint64_t next()
  {
    s += 0x9e3779b97f4a7c15ULL;
    uint64_t x = s;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
  }
  inline double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  inline uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next() * n) >> 64); }
};

inline unsigned
genome_base(uint64_t seed, uint64_t pos)
{
  const uint64_t w = splitmix64(seed ^ (0xa5a5a5a5ULL + (pos >> 5) * 0x2545f4914f6cdd1dULL));
  return (unsigned)(w >> ((pos & 31) * 2)) & 3u;
}

const char kBases[4] = { 'A', 'C', 'G', 'T' };

// 256-quantile table of round(N(0, 2))
struct NoiseTable
{
  int8_t v[256];
  NoiseTable()
  {
    for (int i = 0; i < 256; ++i) {
      // inverse normal CDF by bisection on erf
      const double p = (i + 0.5) / 256.0;
      double lo = -8, hi = 8;
      for (int it = 0; it < 60; ++it) {
        const double mid = 0.5 * (lo + hi);
        const double cdf = 0.5 * (1.0 + std::erf(mid / std::sqrt(2.0)));
        (cdf < p ? lo : hi) = mid;
      }
      v[i] = (int8_t)std::lround(2.0 * 0.5 * (lo + hi));
    }
  }
};

void
make_read(const grb_synth_params& p, uint64_t idx, std::string& out)
{
  static const NoiseTable noise;
  Rng rng(splitmix64(p.seed * 0x9e3779b97f4a7c15ULL + idx + 1));
  uint64_t len = p.read_len;
  if (len == 0) {
    // Box-Muller once per read
    const double u1 = std::max(rng.uniform(), 1e-300), u2 = rng.uniform();
    const double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    const double sigma = 0.5;
    const double mu = std::log((double)p.n50) - sigma * sigma; // length-weighted median = n50
    double l = std::exp(mu + sigma * z);
    l = std::min(200000.0, std::max(1000.0, l));
    len = (uint64_t)l;
  }
  if (len > p.genome_len) {
    len = p.genome_len;
  }
  // template long enough that deletions cannot exhaust it
  const uint64_t tmpl = std::min<uint64_t>(p.genome_len, len + len / 8 + 64);
  const uint64_t start = rng.below(p.genome_len - tmpl + 1);
  const bool rev = rng.next() >> 63;
  const unsigned qread = p.qmin + (unsigned)rng.below(std::max(1u, p.qmax - p.qmin));

  const uint64_t t_sub = (uint64_t)(p.sub_rate * 18446744073709551615.0);
  const uint64_t t_ins = (uint64_t)(p.ins_rate * 18446744073709551615.0);
  const uint64_t t_del = (uint64_t)(p.del_rate * 18446744073709551615.0);

  char hdr[96];
  const int hl = snprintf(hdr, sizeof hdr, "@read%llu pos=%llu\n", (unsigned long long)idx,
                          (unsigned long long)start);
  const size_t base0 = out.size();
  out.resize(base0 + hl + len + 3 + len + 1);
  char* o = &out[base0];
  memcpy(o, hdr, hl);
  char* seq = o + hl;
  char* qual = seq + len + 3;
  uint64_t n = 0;
  uint64_t t = 0;
  while (n < len) {
    unsigned b;
    if (t < tmpl) {
      const uint64_t gp = rev ? (start + tmpl - 1 - t) : (start + t);
      b = genome_base(p.seed, gp);
      if (rev) {
        b = 3 - b;
      }
    } else {
      b = (unsigned)rng.below(4);
    }
    ++t;
    if (t_del && rng.next() < t_del) {
      continue;
    }
    if (t_sub && rng.next() < t_sub) {
      b = (unsigned)rng.below(4);
    }
    seq[n++] = kBases[b];
    if (n < len && t_ins && rng.next() < t_ins) {
      seq[n++] = kBases[rng.below(4)];
    }
  }
  seq[len] = '\n';
  seq[len + 1] = '+';
  seq[len + 2] = '\n';
  for (uint64_t i = 0; i < len; i += 8) {
    uint64_t r = rng.next();
    const uint64_t m = std::min<uint64_t>(8, len - i);
    for (uint64_t j = 0; j < m; ++j, r >>= 8) {
      int q = (int)qread + noise.v[r & 255];
      q = std::min(40, std::max(2, q));
      qual[i + j] = (char)(33 + q);
    }
  }
  qual[len] = '\n';
}

} // namespace

extern "C" {

uint64_t
grb_synth_num_reads(const grb_synth_params* p)
{
  const double mean_len = p->read_len
                            ? (double)p->read_len
                            : std::exp(std::log((double)p->n50) - 0.25 + 0.125); // mu + sigma^2/2
  return (uint64_t)std::ceil(p->coverage * (double)p->genome_len / mean_len);
}

// Generates reads [first, first + count) of the set; returns a malloc'ed buffer the caller
// releases with grb_free_host. *out_len = bytes.
char*
grb_synth_fastq(const grb_synth_params* p, uint64_t first, uint64_t count, uint64_t* out_len)
{
  int nthreads = 1;
#if _OPENMP
  nthreads = omp_get_max_threads();
#endif
  const uint64_t chunk = 64;
  const uint64_t nchunks = (count + chunk - 1) / chunk;
  std::vector<std::string> parts(nchunks);
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
  for (int64_t c = 0; c < (int64_t)nchunks; ++c) {
    std::string& s = parts[c];
    const uint64_t lo = first + (uint64_t)c * chunk;
    const uint64_t hi = std::min(first + count, lo + chunk);
    for (uint64_t i = lo; i < hi; ++i) {
      make_read(*p, i, s);
    }
  }
  std::vector<uint64_t> offs(nchunks + 1, 0);
  for (uint64_t c = 0; c < nchunks; ++c) {
    offs[c + 1] = offs[c] + parts[c].size();
  }
  char* buf = (char*)malloc(std::max<uint64_t>(1, offs[nchunks]));
  if (!buf) {
    *out_len = 0;
    return nullptr;
  }
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
  for (int64_t c = 0; c < (int64_t)nchunks; ++c) {
    memcpy(buf + offs[c], parts[c].data(), parts[c].size());
    std::string().swap(parts[c]);
  }
  *out_len = offs[nchunks];
  return buf;
}

void
grb_free_host(void* p)
{
  free(p);
}

} // extern "C"
