// hostpool.h -- a handful of persistent host threads for the short parallel sections of the frame writer and the
// encoder (recording the rows of a token partition, coding the partitions, writing a frame while the device searches
// for its loop-filter level).  These sections last a millisecond or less, so the threads are created once per
// process: starting a thread per section costs as much as the section on some hosts.
//
//   vp8::HostPool::Group g;
//   g.run([&] { ... });      // queued for a pool thread
//   ...                      // the caller's own share of the work
//   g.wait();                // returns when every task of the group has finished; while it waits the caller
//                            // executes queued tasks itself, so nested groups cannot starve each other
#pragma once
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace vp8 {

class HostPool {
 public:
  class Group {
   public:
    Group() = default;
    Group(const Group&) = delete;
    Group& operator=(const Group&) = delete;
    ~Group() { wait(); }
    void run(std::function<void()> fn) {
      {
        std::lock_guard<std::mutex> lk(m_);
        pending_++;
      }
      HostPool::instance().submit([this, fn = std::move(fn)] {
        fn();
        // counted down and announced under the group's mutex: once wait() has seen 0 under that mutex this task no
        // longer touches the group, which may then be destroyed
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) cv_.notify_all();
      });
    }
    void wait() {
      HostPool& p = HostPool::instance();
      for (;;) {
        {
          std::lock_guard<std::mutex> lk(m_);
          if (pending_ == 0) return;
        }
        if (p.run_one()) continue;  // help: a task of this or of another group
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait_for(lk, std::chrono::microseconds(200), [this] { return pending_ == 0; });
      }
    }

   private:
    int pending_ = 0;  // guarded by m_
    std::mutex m_;
    std::condition_variable cv_;
  };

  static HostPool& instance() {
    static HostPool* p = new HostPool();  // never destroyed: its threads may outlive static destructors
    return *p;
  }

 private:
  HostPool() {
    unsigned n = std::thread::hardware_concurrency();
    n = n < 2 ? 2 : (n > 8 ? 8 : n);
    for (unsigned i = 0; i < n; i++) std::thread([this] { loop(); }).detach();
  }
  void submit(std::function<void()> task) {
    {
      std::lock_guard<std::mutex> lk(m_);
      q_.push_back(std::move(task));
    }
    cv_.notify_one();
  }
  bool run_one() {
    std::function<void()> task;
    {
      std::lock_guard<std::mutex> lk(m_);
      if (q_.empty()) return false;
      task = std::move(q_.front());
      q_.pop_front();
    }
    task();
    return true;
  }
  void loop() {
    for (;;) {
      std::function<void()> task;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [this] { return !q_.empty(); });
        task = std::move(q_.front());
        q_.pop_front();
      }
      task();
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> q_;
};

}  // namespace vp8
