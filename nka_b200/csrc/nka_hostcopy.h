// nka_hostcopy.h -- a few host threads that move bytes between a caller's pageable array and the
// library's pinned staging buffers (nka_accel_update with a host pointer that is not page-locked:
// what the reference's own callers pass, src-C/nka_example.c:139).
//
// cudaMemcpyAsync from pageable memory is staged by the driver through its own bounce buffers on
// one thread: ~14 GB/s over H2D + D2H on the bench box (e2e.pageable, 295-336 ms per 2 GiB update
// against 78 ms from pinned memory).  Copying into pinned chunks with several threads while the
// previous chunk is on the bus gets most of that back without touching the caller's memory
// (page-locking a caller's buffer behind its back is not safe: it may be freed while registered).
// Host data movement only: no arithmetic happens on the CPU.

#pragma once

#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

class NkaHostCopier {
 public:
  // nthreads helper threads (the calling thread copies a share too); 0 = plain memcpy
  explicit NkaHostCopier(int nthreads)
  {
    for (int t = 0; t < nthreads; ++t) workers_.emplace_back([this, t] { run(t); });
  }
  ~NkaHostCopier()
  {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      ++generation_;
    }
    cv_.notify_all();
    for (std::thread& w : workers_) w.join();
  }
  int threads() const { return (int)workers_.size() + 1; }

  // memcpy(dst, src, bytes) split over the helpers and the caller; returns when all of it is done
  void copy(void* dst, const void* src, size_t bytes)
  {
    const size_t parts = workers_.size() + 1;
    if (parts == 1 || bytes < (size_t)(4u << 20)) { memcpy(dst, src, bytes); return; }
    std::lock_guard<std::mutex> one_job(job_);          // the pool holds one job: callers on different threads take turns
    const size_t per = (((bytes + parts - 1) / parts) + 4095) & ~(size_t)4095;   // parts * per >= bytes
    {
      std::lock_guard<std::mutex> lk(m_);
      dst_ = (char*)dst; src_ = (const char*)src; bytes_ = bytes; per_ = per;
      pending_ = (int)workers_.size();
      ++generation_;
    }
    cv_.notify_all();
    piece(parts - 1);                                  // the caller takes the last share
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  void piece(size_t t)
  {
    const size_t lo = t * per_;
    if (lo >= bytes_) return;
    const size_t len = bytes_ - lo < per_ ? bytes_ - lo : per_;
    memcpy(dst_ + lo, src_ + lo, len);
  }
  void run(int t)
  {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
      }
      piece((size_t)t);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }

  std::vector<std::thread> workers_;
  std::mutex m_, job_;
  std::condition_variable cv_, done_;
  unsigned long long generation_ = 0;
  bool stop_ = false;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t bytes_ = 0, per_ = 0;
  int pending_ = 0;
};
