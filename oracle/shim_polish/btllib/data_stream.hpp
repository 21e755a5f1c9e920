// TEST INFRASTRUCTURE ONLY (oracle/).  Stand-in for <btllib/data_stream.hpp>: DataSource as the SAM
// and PAF loaders use it (mappings.cpp:140,196: an object that converts to FILE* for getline()).
// btllib pipes compressed and BAM inputs through external tools; here: plain files only.
#ifndef GRB_SHIM_POLISH_DS_HPP
#define GRB_SHIM_POLISH_DS_HPP
#include "status.hpp"
#include <cstdio>
#include <string>
namespace btllib {
class DataSource
{
public:
  explicit DataSource(const std::string& path)
    : f_(fopen(path.c_str(), "r"))
  {
    check_error(f_ == nullptr, "DataSource: cannot open " + path);
  }
  ~DataSource()
  {
    if (f_) {
      fclose(f_);
    }
  }
  DataSource(const DataSource&) = delete;
  DataSource& operator=(const DataSource&) = delete;
  operator FILE*() const { return f_; }

private:
  FILE* f_;
};
} // namespace btllib
#endif
