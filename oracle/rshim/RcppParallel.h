// oracle/rshim/RcppParallel.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// From-scratch stand-in for the pieces of the R package "RcppParallel" used by
// /root/reference/src/rcpp_parallel_jaccard_coeff.cpp:10-73 so that file can be
// compiled unmodified into oracle/_ref/ (RcppParallel/TBB are not installed and
// cannot be).  Provided semantics:
//
//   * RMatrix<T>: a non-owning view (pointer, nrow, ncol) of a column-major
//     matrix; (i,j) at p[j*nrow+i]; row(i) is a strided range of length ncol
//     (used at :13,:16,:28,:30,:34,:49-51).
//   * Worker: abstract functor over [begin,end) (:10,:24).
//   * parallelFor(begin, end, worker, grain=1): runs disjoint sub-ranges of
//     [begin,end) on a pool of threads (:73).  RcppParallel does this with TBB
//     work stealing; here std::thread workers pull fixed-size chunks from an
//     atomic counter (dynamic scheduling).  The thread count comes from the
//     environment variable RCPP_PARALLEL_NUM_THREADS -- the variable
//     RcppParallel::setThreadOptions(numThreads=) itself sets
//     (R/clustCells.R:64) -- defaulting to all hardware threads.
#ifndef GFICF_ORACLE_RSHIM_RCPPPARALLEL_H
#define GFICF_ORACLE_RSHIM_RCPPPARALLEL_H

#include <atomic>
#include <cstddef>
#include <cstdlib>
#include <iterator>
#include <thread>
#include <vector>

namespace RcppParallel {

template <typename T>
class RMatrix {
 public:
  class Row {
   public:
    class iterator {
     public:
      typedef std::random_access_iterator_tag iterator_category;
      typedef T value_type;
      typedef std::ptrdiff_t difference_type;
      typedef T* pointer;
      typedef T& reference;
      iterator(T* p, std::size_t stride) : p_(p), stride_((std::ptrdiff_t)stride) {}
      reference operator*() const { return *p_; }
      iterator& operator++() { p_ += stride_; return *this; }
      iterator operator++(int) { iterator t = *this; p_ += stride_; return t; }
      iterator& operator--() { p_ -= stride_; return *this; }
      iterator operator+(difference_type n) const { return iterator(p_ + n * stride_, stride_); }
      iterator operator-(difference_type n) const { return iterator(p_ - n * stride_, stride_); }
      iterator& operator+=(difference_type n) { p_ += n * stride_; return *this; }
      difference_type operator-(const iterator& o) const { return (p_ - o.p_) / stride_; }
      reference operator[](difference_type n) const { return p_[n * stride_]; }
      bool operator==(const iterator& o) const { return p_ == o.p_; }
      bool operator!=(const iterator& o) const { return p_ != o.p_; }
      bool operator<(const iterator& o) const { return p_ < o.p_; }

     private:
      T* p_;
      std::ptrdiff_t stride_;
    };
    Row(T* first, std::size_t stride, std::size_t len) : p_(first), stride_(stride), len_(len) {}
    iterator begin() const { return iterator(p_, stride_); }
    iterator end() const { return iterator(p_ + len_ * stride_, stride_); }
    std::size_t length() const { return len_; }
    std::size_t size() const { return len_; }
    T& operator[](std::size_t j) const { return p_[j * stride_]; }

   private:
    T* p_;
    std::size_t stride_, len_;
  };

  template <typename Source>
  RMatrix(const Source& src)
      : p_(const_cast<T*>(src.begin())), nrow_((std::size_t)src.nrow()), ncol_((std::size_t)src.ncol()) {}
  RMatrix(T* p, std::size_t nrow, std::size_t ncol) : p_(p), nrow_(nrow), ncol_(ncol) {}

  std::size_t nrow() const { return nrow_; }
  std::size_t ncol() const { return ncol_; }
  std::size_t length() const { return nrow_ * ncol_; }
  T& operator()(std::size_t i, std::size_t j) const { return p_[j * nrow_ + i]; }
  Row row(std::size_t i) const { return Row(p_ + i, nrow_, ncol_); }

 private:
  T* p_;
  std::size_t nrow_, ncol_;
};

struct Worker {
  virtual ~Worker() {}
  virtual void operator()(std::size_t begin, std::size_t end) = 0;
};

inline int resolved_threads() {
  const char* e = std::getenv("RCPP_PARALLEL_NUM_THREADS");
  int nt = e ? std::atoi(e) : 0;
  if (nt <= 0) nt = (int)std::thread::hardware_concurrency();
  if (nt <= 0) nt = 1;
  return nt;
}

inline void parallelFor(std::size_t begin, std::size_t end, Worker& worker,
                        std::size_t grainSize = 1) {
  if (end <= begin) return;
  const std::size_t total = end - begin;
  int nt = resolved_threads();
  if ((std::size_t)nt > total) nt = (int)total;
  // chunk: small enough for load balance, large enough to amortise the atomic
  std::size_t chunk = total / ((std::size_t)nt * 64);
  if (chunk < grainSize) chunk = grainSize;
  if (chunk < 1) chunk = 1;
  if (nt == 1) {
    worker(begin, end);
    return;
  }
  std::atomic<std::size_t> next(begin);
  std::vector<std::thread> pool;
  pool.reserve((std::size_t)nt);
  for (int t = 0; t < nt; ++t) {
    pool.emplace_back([&]() {
      for (;;) {
        std::size_t lo = next.fetch_add(chunk);
        if (lo >= end) break;
        std::size_t hi = lo + chunk < end ? lo + chunk : end;
        worker(lo, hi);
      }
    });
  }
  for (std::size_t t = 0; t < pool.size(); ++t) pool[t].join();
}

}  // namespace RcppParallel

#endif  // GFICF_ORACLE_RSHIM_RCPPPARALLEL_H
