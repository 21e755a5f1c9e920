// TEST INFRASTRUCTURE ONLY (oracle/).  Stand-in for <btllib/util.hpp>.
#ifndef GRB_SHIM_POLISH_UTIL_HPP
#define GRB_SHIM_POLISH_UTIL_HPP
#include "status.hpp"
#endif
