// TEST SCAFFOLDING (oracle build only) -- not product code.
// Fake pcl::ONIGrabber: "-oni <file>" names a raw depth container
//   int32 n, int32 w, int32 h (little endian) followed by n*w*h uint16 millimetres,
// frames are handed out one per trigger() with FrameID = 1, 2, ... .  The callback
// is fired from a worker thread because the reference main loop holds its mutex
// while it calls trigger() (Integrate/IntegrateApp.cpp:97-107).
// With ER_SHIM_TIMING set, prints the wall time from the first trigger() to the
// end-of-stream trigger() (= the time spent in Execute for all frames, without
// the reference's 10 x 300 ms end-of-stream timeout).  With ER_SHIM_TIMING_FILE=<path>
// it also writes the time of every trigger() call (seconds since the first, one per
// line): consecutive differences are the per-frame processing times.
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <pcl/io/grabber.h>

namespace pcl {
class ONIGrabber : public Grabber {
 public:
  typedef Grabber::sig_cb_openni_image_depth_image sig_cb_openni_image_depth_image;

  ONIGrabber(const std::string& file, bool /*repeat*/, bool /*stream*/)
      : f_(std::fopen(file.c_str(), "rb")), n_(0), w_(0), h_(0), next_(0), pending_(0), quit_(false),
        started_(false), reported_(false) {
    if (!f_) throw PCLException("cannot open depth container " + file);
    int32_t hdr[3];
    if (std::fread(hdr, sizeof(int32_t), 3, f_) != 3) throw PCLException("bad depth container header");
    n_ = hdr[0]; w_ = hdr[1]; h_ = hdr[2];
    buf_.resize((size_t)w_ * h_);
    worker_ = std::thread([this] { this->run(); });
  }
  ~ONIGrabber() override {
    { std::lock_guard<std::mutex> g(m_); quit_ = true; }
    cv_.notify_one();
    worker_.join();
    if (f_) std::fclose(f_);
  }
  void trigger() {
    auto now = std::chrono::steady_clock::now();
    std::lock_guard<std::mutex> g(m_);
    if (!started_) { started_ = true; t0_ = now; }
    if (!reported_) stamps_.push_back(std::chrono::duration<double>(now - t0_).count());
    if (next_ >= n_) {
      if (!reported_ && std::getenv("ER_SHIM_TIMING_FILE")) {
        if (std::FILE* tf = std::fopen(std::getenv("ER_SHIM_TIMING_FILE"), "w")) {
          for (double t : stamps_) std::fprintf(tf, "%.9f\n", t);
          std::fclose(tf);
        }
      }
      if (!reported_ && std::getenv("ER_SHIM_TIMING")) {
        double s = std::chrono::duration<double>(now - t0_).count();
        std::fprintf(stderr, "[shim] frames=%d elapsed_s=%.6f ms_per_frame=%.4f\n", n_, s, n_ ? 1e3 * s / n_ : 0.0);
      }
      reported_ = true;
      return;
    }
    ++pending_;
    cv_.notify_one();
  }
 private:
  void run() {
    for (;;) {
      int id;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [this] { return quit_ || pending_ > 0; });
        if (quit_) return;
        --pending_;
        if (std::fread(buf_.data(), sizeof(uint16_t), buf_.size(), f_) != buf_.size()) { next_ = n_; continue; }
        id = ++next_;
      }
      boost::shared_ptr<openni_wrapper::Image> img(new openni_wrapper::Image(w_, h_));
      boost::shared_ptr<openni_wrapper::DepthImage> dep(new openni_wrapper::DepthImage(w_, h_, id, buf_.data()));
      if (cb_) cb_(img, dep, 1.0f);
    }
  }
  std::FILE* f_;
  int n_, w_, h_, next_, pending_;
  bool quit_, started_, reported_;
  std::vector<uint16_t> buf_;
  std::vector<double> stamps_;
  std::mutex m_;
  std::condition_variable cv_;
  std::thread worker_;
  std::chrono::steady_clock::time_point t0_;
};
}  // namespace pcl
