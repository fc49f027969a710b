// TEST SCAFFOLDING (oracle build only) -- not product code.
// pcl::console verbosity + the PCL_* print macros (FragmentOptimizer/FragmentOptimizer.cpp:44-45).
#pragma once
#include <pcl/common/time.h>   // PCL_INFO / PCL_WARN / PCL_ERROR stand-ins
namespace pcl { namespace console {
enum VERBOSITY_LEVEL { L_ALWAYS, L_ERROR, L_WARN, L_INFO, L_DEBUG, L_VERBOSE };
inline void setVerbosityLevel(VERBOSITY_LEVEL) {}
}}  // namespace pcl::console
