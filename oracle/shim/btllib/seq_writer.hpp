// TEST INFRASTRUCTURE ONLY (oracle/). Empty stand-in for btllib's <btllib/seq_writer.hpp>: the
// reference includes it (goldrush_path/goldrush_path.cpp:10-11) but uses nothing from it.
#ifndef GRB_SHIM_BTLLIB_SEQ_WRITER_HPP
#define GRB_SHIM_BTLLIB_SEQ_WRITER_HPP
#endif
