// Staging buffers of the batch decoders (pinned upload slices, packed-input and scratch device buffers) are expensive to make
// (cudaHostAlloc ~0.2 ms per MB) and the scene driver decodes on short-lived threads, so they are pooled per device for the
// life of the process: a call leases one, the lease returns it.  T needs `int device`, a default constructor and release().
#pragma once
#include <memory>
#include <mutex>
#include <vector>

extern "C" int scn_current_device_(void);      // cudaGetDevice of the calling thread (scn_common.cu); keeps this header free of CUDA includes

namespace scn {

template <class T>
class StagePool {
 public:
  class Lease {
   public:
    Lease(StagePool& p) : pool_(p) {
      const int dev = scn_current_device_();
      {
        std::lock_guard<std::mutex> l(p.m_);
        for (size_t i = 0; i < p.idle_.size(); ++i)
          if (p.idle_[i]->device == dev) { s_ = std::move(p.idle_[i]); p.idle_.erase(p.idle_.begin() + (long)i); break; }
      }
      if (!s_) { s_.reset(new T()); s_->device = dev; }
    }
    ~Lease() { std::lock_guard<std::mutex> l(pool_.m_); pool_.idle_.push_back(std::move(s_)); }
    T& operator*() { return *s_; }
    T* operator->() { return s_.get(); }
   private:
    StagePool& pool_; std::unique_ptr<T> s_;
  };
  // frees every idle stage of the calling thread's current device (scn_release_cached_staging)
  void trim() {
    const int dev = scn_current_device_();
    std::vector<std::unique_ptr<T>> mine;
    {
      std::lock_guard<std::mutex> l(m_);
      for (size_t i = 0; i < idle_.size();) if (idle_[i]->device == dev) { mine.push_back(std::move(idle_[i])); idle_.erase(idle_.begin() + (long)i); } else ++i;
    }
    for (auto& s : mine) s->release();
  }
 private:
  std::mutex m_; std::vector<std::unique_ptr<T>> idle_;     // never destroyed at exit: the CUDA runtime may be gone by then
};

}  // namespace scn
