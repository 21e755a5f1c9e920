// TEST INFRASTRUCTURE ONLY (oracle/). Empty stand-in for btllib's <btllib/util.hpp>: the
// reference includes it (goldrush_path/goldrush_path.cpp:10-11) but uses nothing from it.
#ifndef GRB_SHIM_BTLLIB_UTIL_HPP
#define GRB_SHIM_BTLLIB_UTIL_HPP
#endif
