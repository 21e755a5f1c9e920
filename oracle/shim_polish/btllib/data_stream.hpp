// TEST INFRASTRUCTURE ONLY (oracle/).  Stand-in for <btllib/data_stream.hpp> (nothing of it is used).
#ifndef GRB_SHIM_POLISH_DS_HPP
#define GRB_SHIM_POLISH_DS_HPP
#endif
