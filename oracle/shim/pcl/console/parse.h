// TEST SCAFFOLDING (oracle build only) -- not product code.
// pcl::console::{parse_argument,find_switch}: "--key value" lookup returning the
// index of the value (or -1), as the reference mains expect (> 0 means found).
#pragma once
#include <cstring>
#include <cstdlib>
#include <string>

namespace pcl { namespace console {
inline int find_argument(int argc, char** argv, const char* name) {
  for (int i = 1; i < argc; ++i)
    if (std::strcmp(argv[i], name) == 0) return i;
  return -1;
}
inline bool find_switch(int argc, char** argv, const char* name) {
  return find_argument(argc, argv, name) != -1;
}
inline int parse_argument(int argc, char** argv, const char* name, std::string& v) {
  int i = find_argument(argc, argv, name) + 1;
  if (i > 0 && i < argc) { v = argv[i]; return i - 1; }
  return -1;
}
inline int parse_argument(int argc, char** argv, const char* name, int& v) {
  int i = find_argument(argc, argv, name) + 1;
  if (i > 0 && i < argc) { v = std::atoi(argv[i]); return i - 1; }
  return -1;
}
inline int parse_argument(int argc, char** argv, const char* name, double& v) {
  int i = find_argument(argc, argv, name) + 1;
  if (i > 0 && i < argc) { v = std::atof(argv[i]); return i - 1; }
  return -1;
}
}}  // namespace pcl::console
