// TEST SCAFFOLDING (oracle build only) -- not product code.
// Stand-ins for pcl::StopWatch / pcl::ScopeTime and the PCL_* console macros.
#pragma once
#include <chrono>
#include <cstdio>
#include <string>
#include <iostream>

#define PCL_INFO(...)  do { std::fprintf(stdout, __VA_ARGS__); } while (0)
#define PCL_WARN(...)  do { std::fprintf(stderr, __VA_ARGS__); } while (0)
#define PCL_ERROR(...) do { std::fprintf(stderr, __VA_ARGS__); } while (0)

namespace pcl {
class StopWatch {
 public:
  StopWatch() : t0_(std::chrono::steady_clock::now()) {}
  double getTime() const {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0_).count();
  }
 protected:
  std::chrono::steady_clock::time_point t0_;
};
class ScopeTime : public StopWatch {
 public:
  explicit ScopeTime(const char* title) : title_(title) {}
  ~ScopeTime() { std::cerr << title_ << " took " << getTime() << "ms.\n"; }
 private:
  std::string title_;
};
}  // namespace pcl
