// grb-synth: command-line front end of the deterministic synthetic read generator (synth.cpp).
//   grb-synth -G <genome bp> -c <coverage> -l <read len | 0> [-n N50] [-s seed] [-e rate] [-q lo,hi]
//             [-N max reads] -o out.fq
#include "goldrush_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

int
main(int argc, char** argv)
{
  grb_synth_params p{};
  p.genome_len = 1000000;
  p.seed = 1;
  p.coverage = 25;
  p.read_len = 20000;
  p.n50 = 20000;
  p.sub_rate = p.ins_rate = p.del_rate = 0.01;
  p.qmin = 12;
  p.qmax = 30;
  uint64_t max_reads = 0;
  std::string out;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string a = argv[i];
    const char* v = argv[i + 1];
    if (a == "-G") p.genome_len = (uint64_t)strtod(v, nullptr);
    else if (a == "-c") p.coverage = strtod(v, nullptr);
    else if (a == "-l") p.read_len = (uint32_t)strtoul(v, nullptr, 10);
    else if (a == "-n") p.n50 = (uint32_t)strtoul(v, nullptr, 10);
    else if (a == "-s") p.seed = strtoull(v, nullptr, 10);
    else if (a == "-e") p.sub_rate = p.ins_rate = p.del_rate = strtod(v, nullptr);
    else if (a == "-q") { sscanf(v, "%u,%u", &p.qmin, &p.qmax); }
    else if (a == "-N") max_reads = strtoull(v, nullptr, 10);
    else if (a == "-o") out = v;
    else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
  }
  if (out.empty()) { fprintf(stderr, "usage: grb-synth -G genome -c cov -l len -s seed -o out.fq\n"); return 1; }
  uint64_t n = grb_synth_num_reads(&p);
  if (max_reads && n > max_reads) n = max_reads;
  FILE* f = fopen(out.c_str(), "wb");
  if (!f) { perror("fopen"); return 1; }
  const uint64_t step = 4096;
  for (uint64_t first = 0; first < n; first += step) {
    uint64_t len = 0;
    const uint64_t cnt = (n - first < step) ? n - first : step;
    char* buf = grb_synth_fastq(&p, first, cnt, &len);
    if (!buf) { fprintf(stderr, "out of memory\n"); return 1; }
    fwrite(buf, 1, len, f);
    grb_free_host(buf);
  }
  fclose(f);
  fprintf(stderr, "wrote %llu reads to %s\n", (unsigned long long)n, out.c_str());
  return 0;
}
