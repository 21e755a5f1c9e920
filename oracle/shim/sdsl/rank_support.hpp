// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for sdsl-lite's <sdsl/rank_support.hpp>;
// rank_support_il lives with bit_vector_il in this stand-in.
#ifndef GRB_SHIM_SDSL_RANK_SUPPORT_HPP
#define GRB_SHIM_SDSL_RANK_SUPPORT_HPP
#include "bit_vector_il.hpp"
#endif
