#include "opt.hpp"

#include <cstdlib>
#include <getopt.h>
#include <iostream>

void
grb_print_usage(const std::string& progname)
{
  // same option set and bracketed defaults as the reference's help text (opt.cpp:34-88)
  std::cout
    << "Usage:  " << progname
    << "  -k K -w W -i INPUT -g G [-p prefix] [-P PHRED_AVG] [-o O] [-t T] [-f F] [-h H] [-u U] "
       "[-m M] [-H HASH_UNIVERSE] [-s S] [-x X] [-M MAX_PATHS][-a A] [-j J] [-b B] [-d D] "
       "[--silver_path] [--ntcard] [--help] \n\n"
    << "  -i INPUT                find golden paths from INPUT [required]\n"
    << "  -g G                    estimated genome size [required]\n"
    << "  -b B                    during insertion, B number of consecutive tiles to be inserted "
       "with the same ID [10]\n"
    << "  -d D                    remove reads with greater or equal then D phred average between "
       "first half and second half of the read [5]\n"
    << "  -f F                    don't use reads from F. Expects one read per line\n"
    << "  -o O                    use O as occupancy [0.1]\n"
    << "  -h H                    use h as number of spaced seed patterns [1]\n"
    << "  -H HASH_UNIVERSE        determine MiBF size based on HASH_UNIVERSE [Calculated based on "
       "W and h]\n"
    << "  -t T                    tile length [1000]\n"
    << "  -k K                    span of spaced seed [required]\n"
    << "  -w W                    weight of spaced seed [required]\n"
    << "  -m M                    use reads longer than M [20000]\n"
    << "  -u U                    U minimum unassigned tiles for read to be unassigned [5]\n"
    << "  -a A                    A maximum assigned tiles for read to be unassigned [1]\n"
    << "  -p prefix               write output to files with prefix [goldrush_out]\n"
    << "  -P PHRED_AVG            minimum average phred score for each read [0 (calculates phred "
       "score minimum automatically)]\n"
    << "  -j J                    number of threads [48]\n"
    << "  -s S                    use S seed preset. Must be consistent with k and w [n/a, generate "
       "one randomly based on k and w]\n"
    << "  -x X                    require X hits for a tile to be assigned [10]\n"
    << "  -M MAX_PATHS            output MAX_PATHS [5, used with --silver_path]\n"
    << "  --ntcard                use ntcard to estimate genome size [false, assume max entries]\n"
    << "  --silver_path           generate silver path(s) instead of golden path. Silver paths "
       "terminate when the number of bases recruited equals or exceeds T * r\n"
    << " --verbose                print verbose messages [false]\n"
    << "  --help                  display this help and exit\n";
}

int
grb_parse_cli(int argc, char** argv, GrbCli& cli)
{
  grb_params_default(&cli.params);
  cli.params.kmer_size = 0;
  cli.params.weight = 0;
  const struct option longopts[] = { { "debug", no_argument, &cli.debug, 1 },
                                     { "verbose", no_argument, &cli.verbose, 1 },
                                     { "silver_path", no_argument, &cli.silver_path, 1 },
                                     { "help", no_argument, &cli.help, 1 },
                                     { "ntcard", no_argument, &cli.ntcard, 1 },
                                     { nullptr, 0, nullptr, 0 } };
  grb_params& p = cli.params;
  int c, idx = 0;
  char* end = nullptr;
  while ((c = getopt_long(argc, argv, "a:b:d:f:g:h:i:j:k:m:M:o:r:s:t:u:w:x:p:P:H:", longopts,
                          &idx)) != -1) {
    switch (c) {
      case 0: break;
      case 'a': p.assigned_max = strtoul(optarg, &end, 10); break;
      case 'b': p.block_size = strtoul(optarg, &end, 10); break;
      case 'd': p.phred_delta = (uint32_t)strtoul(optarg, &end, 10); break;
      case 'f': cli.filter_file = optarg; break;
      case 'H': p.hash_universe = strtoull(optarg, &end, 10); break;
      case 'h': p.hash_num = strtoul(optarg, &end, 10); break;
      case 'i': cli.input = optarg; break;
      case 'j': cli.jobs = strtoul(optarg, &end, 10); break;
      case 'k': p.kmer_size = strtoul(optarg, &end, 10); break;
      case 'm': p.min_length = strtoul(optarg, &end, 10); break;
      case 'M': p.max_paths = strtoul(optarg, &end, 10); break;
      case 'o': p.occupancy = strtod(optarg, &end); break;
      case 'r': p.ratio = strtod(optarg, &end); break;
      case 'p': cli.prefix_file = optarg; break;
      case 'P': p.phred_min = (uint32_t)strtoul(optarg, &end, 10); break;
      case 's': cli.seed_preset = optarg; break;
      case 't': p.tile_length = strtoul(optarg, &end, 10); break;
      case 'g': p.genome_size = (uint64_t)strtod(optarg, &end); break;
      case 'u': p.unassigned_min = strtoul(optarg, &end, 10); break;
      case 'w': p.weight = strtoul(optarg, &end, 10); break;
      case 'x': p.threshold = strtoul(optarg, &end, 10); break;
      default: return EXIT_FAILURE;
    }
  }
  p.silver_path = cli.silver_path;
  if (cli.help) {
    grb_print_usage("goldrush_path");
    return 0;
  }
  auto reject = [](const char* msg) {
    std::cerr << msg << std::endl;
    grb_print_usage("goldrush_path");
    return 1;
  };
  if (!p.kmer_size) {
    return reject("span of spaced seed cannot be 0");
  }
  if (!p.weight) {
    return reject("weight of spaced seed cannot be 0");
  }
  if (p.genome_size == 0) {
    return reject("genome size cannot be 0");
  }
  if (!cli.seed_preset.empty()) {
    if (p.kmer_size != cli.seed_preset.size()) {
      return reject("seed preset must be the same size of k");
    }
    uint8_t ones = 0; // 8-bit counter, as in the reference
    for (char ch : cli.seed_preset) {
      ones += ch == '1';
    }
    if (p.weight != ones) {
      return reject("seed preset must have the same weight as w");
    }
  }
  return -1;
}
