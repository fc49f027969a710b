// TEST SCAFFOLDING (oracle build only) -- not product code.
// Minimal stand-ins for the Boost pieces the reference Integrate sources use
// (shared_ptr, function/bind, filesystem::exists, mutex/condition_variable,
// posix_time::millisec, signals2::connection), mapped onto the C++11 library.
#pragma once
#include <memory>
#include <functional>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <string>
#include <sys/stat.h>

namespace boost {

using std::shared_ptr;
using std::function;
using std::bind;

namespace filesystem {
inline bool exists(const std::string& p) {
  struct stat st;
  return !p.empty() && ::stat(p.c_str(), &st) == 0;
}
}  // namespace filesystem

namespace posix_time {
struct millisec {
  long ms;
  explicit millisec(long v) : ms(v) {}
};
}  // namespace posix_time

class mutex {
 public:
  std::mutex m_;
  void lock() { m_.lock(); }
  void unlock() { m_.unlock(); }
  bool try_lock() { return m_.try_lock(); }
  class scoped_lock {
   public:
    explicit scoped_lock(mutex& m) : m_(m) { m_.lock(); }
    ~scoped_lock() { m_.unlock(); }
   private:
    mutex& m_;
  };
  class scoped_try_lock {
   public:
    explicit scoped_try_lock(mutex& m) : m_(m), owns_(m.try_lock()) {}
    ~scoped_try_lock() { if (owns_) m_.unlock(); }
    bool operator!() const { return !owns_; }
   private:
    mutex& m_;
    bool owns_;
  };
};

template <class M>
class unique_lock {
 public:
  explicit unique_lock(M& m) : m_(m), owns_(true) { m_.lock(); }
  ~unique_lock() { if (owns_) m_.unlock(); }
  void lock() { m_.lock(); owns_ = true; }
  void unlock() { m_.unlock(); owns_ = false; }
 private:
  M& m_;
  bool owns_;
};

// notify/wait go through an internal mutex so that a notify issued after the
// producer dropped the user mutex can never be lost and a spurious wake-up is
// never reported as data (either would change which frames get integrated).
class condition_variable {
 public:
  condition_variable() : flag_(false) {}
  void notify_one() {
    { std::lock_guard<std::mutex> g(im_); flag_ = true; }
    icv_.notify_one();
  }
  template <class L>
  bool timed_wait(L& user_lock, const posix_time::millisec& d) {
    std::unique_lock<std::mutex> il(im_);
    user_lock.unlock();
    bool r = icv_.wait_for(il, std::chrono::milliseconds(d.ms), [this] { return flag_; });
    flag_ = false;
    il.unlock();
    user_lock.lock();
    return r;
  }
 private:
  std::mutex im_;
  std::condition_variable icv_;
  bool flag_;
};

namespace signals2 {
struct connection {
  void disconnect() {}
};
}  // namespace signals2

}  // namespace boost

using std::placeholders::_1;
using std::placeholders::_2;
using std::placeholders::_3;
