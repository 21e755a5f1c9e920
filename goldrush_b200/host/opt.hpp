// Command line of goldrush-path: the same grammar, defaults, validation messages and exit codes as
// the reference's opt namespace (goldrush_path/opt.hpp:9-47, opt.cpp:5-32,90-217).
#ifndef GRB_HOST_OPT_HPP
#define GRB_HOST_OPT_HPP

#include "goldrush_b200.h"

#include <string>

struct GrbCli
{
  grb_params params;
  std::string prefix_file = "goldrush_out";
  std::string input;
  std::string seed_preset;
  std::string filter_file;
  unsigned long jobs = 48;
  int help = 0;
  int ntcard = 0;
  int silver_path = 0;
  int verbose = 0;
  int debug = 0;
};

// Parses argv into cli.  Returns -1 to continue, otherwise the process exit code (0 after --help,
// 1 on an invalid option set), having printed what the reference prints.
int grb_parse_cli(int argc, char** argv, GrbCli& cli);
void grb_print_usage(const std::string& progname);

#endif
