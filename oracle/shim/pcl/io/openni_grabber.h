// TEST SCAFFOLDING (oracle build only): there is no OpenNI device here, so the
// live grabber always fails to open, which the reference main reports and exits on.
#pragma once
#include <pcl/io/grabber.h>
namespace pcl {
class OpenNIGrabber : public Grabber {
 public:
  explicit OpenNIGrabber(const std::string& = "") { throw PCLException("no OpenNI device in the oracle shim"); }
};
}  // namespace pcl
