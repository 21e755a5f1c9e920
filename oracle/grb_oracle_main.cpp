// TEST INFRASTRUCTURE ONLY — command-line front end of the CPU oracle (same options as the
// reference's goldrush-path, goldrush_path/opt.cpp:90-217).
#include "grb_oracle.h"

int
main(int argc, char** argv)
{
  return grbo_main(argc, argv);
}
