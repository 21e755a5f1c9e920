// goldrush-path — drop-in replacement of the reference executable
// (goldrush_path/goldrush_path.cpp main(), :1096-1275): same options, same <prefix>_N.fq /
// <prefix>.fa outputs, exit code 0 on success (also when the M-th silver path completes, where the
// reference calls exit(0)), 1 on option / format / no-reads errors.
#include "goldrush_b200.h"
#include "opt.hpp"

#include <cstdlib>
#include <iostream>

int
main(int argc, char** argv)
{
  GrbCli cli;
  const int early = grb_parse_cli(argc, argv, cli);
  if (early >= 0) {
    return early;
  }
  grb_run_options o{};
  o.params = cli.params;
  o.params.device = getenv("GRB_DEVICE") ? atoi(getenv("GRB_DEVICE")) : 0;
  o.seed_preset = cli.seed_preset.c_str();
  o.prefix = cli.prefix_file.c_str();
  o.filter_file = cli.filter_file.empty() ? nullptr : cli.filter_file.c_str();
  o.input_path = cli.input.c_str();
  o.ntcard = cli.ntcard;
  o.verbose = cli.verbose;
  o.debug = cli.debug;
  o.write_outputs = 1;
  o.quiet = 0;
  o.jobs = (int)cli.jobs;
  grb_run_result res{};
  char err[1024] = { 0 };
  const int rc = grb_run_path(&o, nullptr, 0, &res, err, sizeof err);
  if (rc != GRB_OK) {
    std::cerr << "goldrush-path: " << err << std::endl;
    return 1;
  }
  return 0;
}
