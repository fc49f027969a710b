// TEST SCAFFOLDING (oracle build only) -- not product code.
// Stand-ins for pcl::Grabber, pcl::PCLException and the openni_wrapper image
// types, just wide enough for the reference IntegrateApp to compile unmodified.
#pragma once
#include <cstdint>
#include <cstring>
#include <exception>
#include <functional>
#include <memory>
#include <string>
#include <vector>
#include <boost/filesystem.hpp>

namespace openni_wrapper {
class Image {
 public:
  Image(unsigned w, unsigned h) : w_(w), h_(h) {}
  unsigned getWidth() const { return w_; }
  unsigned getHeight() const { return h_; }
 private:
  unsigned w_, h_;
};
struct DepthMetaData {
  int frame_id;
  int FrameID() const { return frame_id; }
};
class DepthImage {
 public:
  DepthImage(unsigned w, unsigned h, int frame_id, const uint16_t* data)
      : w_(w), h_(h), data_(data, data + (size_t)w * h) { md_.frame_id = frame_id; }
  unsigned getWidth() const { return w_; }
  unsigned getHeight() const { return h_; }
  void fillDepthImageRaw(unsigned w, unsigned h, unsigned short* dst) const {
    std::memcpy(dst, data_.data(), sizeof(uint16_t) * (size_t)w * h);
  }
  const DepthMetaData& getDepthMetaData() const { return md_; }
 private:
  unsigned w_, h_;
  std::vector<uint16_t> data_;
  DepthMetaData md_;
};
}  // namespace openni_wrapper

namespace pcl {
class PCLException : public std::exception {
 public:
  explicit PCLException(const std::string& m) : m_(m) {}
  const char* what() const noexcept override { return m_.c_str(); }
 private:
  std::string m_;
};

class Grabber {
 public:
  typedef void(sig_cb_openni_image_depth_image)(const boost::shared_ptr<openni_wrapper::Image>&,
                                                const boost::shared_ptr<openni_wrapper::DepthImage>&,
                                                float);
  virtual ~Grabber() {}
  template <typename T> bool providesCallback() const { return true; }
  template <typename F> boost::signals2::connection registerCallback(const F& f) {
    cb_ = f;
    return boost::signals2::connection();
  }
  virtual void start() {}
  virtual void stop() {}
 protected:
  std::function<sig_cb_openni_image_depth_image> cb_;
};
}  // namespace pcl
