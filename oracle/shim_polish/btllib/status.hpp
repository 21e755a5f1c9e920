// TEST INFRASTRUCTURE ONLY (oracle/).  Stand-in for <btllib/status.hpp>.
#ifndef GRB_SHIM_POLISH_BTLLIB_STATUS_HPP
#define GRB_SHIM_POLISH_BTLLIB_STATUS_HPP
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
namespace btllib {
inline void
check_error(bool condition, const std::string& msg)
{
  if (condition) {
    std::cerr << "[ERROR] " << msg << std::endl;
    std::exit(EXIT_FAILURE);
  }
}
inline void log_info(const std::string& msg) { std::cerr << "[INFO] " << msg << std::endl; }
inline void log_error(const std::string& msg) { std::cerr << "[ERROR] " << msg << std::endl; }
inline std::string get_strerror() { return std::strerror(errno); }
template<class Stream>
inline void
check_stream(const Stream& stream, const std::string& name)
{
  check_error(!stream.good(), "'" + name + "' stream error: " + get_strerror());
}
} // namespace btllib
#endif
