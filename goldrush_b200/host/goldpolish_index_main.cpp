// Drop-in for the reference's `goldpolish-index` (subprojects/goldpolish/src/goldpolish_index.cpp):
// same two arguments, same "Wrong args." refusal; writes the sequence index goldpolish-targeted-bfs
// loads (grb_polish_index_build; lines in file order where the reference writes hash-table order).
#include "goldrush_b200.h"

#include <cstdio>
#include <cstdlib>
#include <iostream>

int
main(int argc, char** argv)
{
  if (argc != 3) {
    std::cerr << "Wrong args.\n";
    std::exit(EXIT_FAILURE);
  }
  char err[1024] = "";
  if (grb_polish_index_build(argv[1], argv[2], err, sizeof err) != GRB_OK) {
    std::cerr << "[ERROR] " << err << std::endl;
    return EXIT_FAILURE;
  }
  return 0;
}
