// oracle/rshim/Rcpp.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A from-scratch, header-only stand-in for the few pieces of the R package
// "Rcpp" that the reference's Jaccard sources touch, so that
//   /root/reference/src/rcpp_parallel_jaccard_coeff.cpp   and
//   /root/reference/src/jaccard_coeff.cpp
// can be compiled UNMODIFIED, from where they lie, into oracle/_ref/ (R, Rcpp,
// RcppParallel and RcppProgress are not installed in this image and cannot be:
// there is no network).  Nothing here is copied from Rcpp; only the observable
// semantics the two reference files rely on are provided:
//
//   * NumericMatrix: column-major double storage, element (i,j) at p[j*nrow+i],
//     zero-filled on (rows, cols) construction, reference (shared) copy
//     semantics like an R SEXP handle
//     (used at rcpp_parallel_jaccard_coeff.cpp:59,67 and jaccard_coeff.cpp:19-21).
//   * idx(i, _) row proxy convertible to NumericVector (jaccard_coeff.cpp:31-32).
//   * intersect(a, b): the DISTINCT values of a that also occur in b, i.e.
//     unique-set semantics (jaccard_coeff.cpp:33) -- that is what Rcpp sugar's
//     intersect documents (it mirrors R's base::intersect).
//   * Rprintf (rcpp_parallel_jaccard_coeff.cpp:63,77; jaccard_coeff.cpp:25).
#ifndef GFICF_ORACLE_RSHIM_RCPP_H
#define GFICF_ORACLE_RSHIM_RCPP_H

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <functional>
#include <numeric>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <iterator>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

// Banner sink: when the harness sets this, Rprintf appends here instead of
// stdout (so tests can assert the reference's banner strings).
namespace rshim {
inline std::vector<char>*& printf_sink() {
  static std::vector<char>* sink = nullptr;
  return sink;
}
}  // namespace rshim

inline void Rprintf(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  int m = vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (m < 0) return;
  if (m > (int)sizeof buf - 1) m = (int)sizeof buf - 1;
  if (rshim::printf_sink())
    rshim::printf_sink()->insert(rshim::printf_sink()->end(), buf, buf + m);
  else
    fwrite(buf, 1, (size_t)m, stdout);
}

typedef std::ptrdiff_t R_xlen_t;

// SEXP stand-in (used by the drop-in bodies of gficf_b200/rpkg/src that take the matrix as it
// comes from R, integer or double; not used by the reference sources): a tagged matrix handle.
struct SEXPREC {
  int type;         // INTSXP or REALSXP
  int nrow, ncol;
  void* data;       // column-major int / double storage, owned by the harness
};
typedef SEXPREC* SEXP;
#define INTSXP 13
#define REALSXP 14
inline int TYPEOF(SEXP x) { return x->type; }

namespace Rcpp {

// Rcpp::stop(fmt, ...): raises an R error; here a C++ exception the harness catches.
// (Not used by the reference sources; used by the drop-in bodies in gficf_b200/rpkg/src
// when tests compile them against this stand-in.)
struct exception : public std::runtime_error {
  explicit exception(const std::string& m) : std::runtime_error(m) {}
};
inline void stop(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw exception(buf);
}

struct Placeholder {};
static const Placeholder _ = Placeholder();

class NumericVector {
 public:
  typedef std::vector<double>::iterator iterator;
  typedef std::vector<double>::const_iterator const_iterator;
  NumericVector() {}
  explicit NumericVector(std::size_t len) : v_(len, 0.0) {}
  template <typename It>
  NumericVector(It first, It last) : v_(first, last) {}
  std::size_t size() const { return v_.size(); }
  std::size_t length() const { return v_.size(); }
  double& operator[](std::size_t i) { return v_[i]; }
  const double& operator[](std::size_t i) const { return v_[i]; }
  double& operator()(std::size_t i) { return v_[i]; }  // mann_whitney.cpp:117 res(k)
  const double& operator()(std::size_t i) const { return v_[i]; }
  iterator begin() { return v_.begin(); }
  iterator end() { return v_.end(); }
  const_iterator begin() const { return v_.begin(); }
  const_iterator end() const { return v_.end(); }
  void push_back(double x) { v_.push_back(x); }

 private:
  std::vector<double> v_;
};

class NumericMatrix {
 public:
  // A strided view of one matrix row; converts to an owning NumericVector.
  class RowProxy {
   public:
    RowProxy(const double* p, std::size_t stride, std::size_t len)
        : p_(p), stride_(stride), len_(len) {}
    operator NumericVector() const {
      NumericVector out(len_);
      for (std::size_t j = 0; j < len_; ++j) out[j] = p_[j * stride_];
      return out;
    }

   private:
    const double* p_;
    std::size_t stride_, len_;
  };

  NumericMatrix() : nrow_(0), ncol_(0), p_(nullptr) {}
  // From an R object: a double matrix is wrapped, an integer one is coerced into a fresh double
  // copy -- what Rcpp's input_parameter<NumericMatrix> does (reference src/RcppExports.cpp:65).
  NumericMatrix(SEXP x) : nrow_(x->nrow), ncol_(x->ncol), p_(nullptr) {  // NOLINT: implicit like Rcpp's
    if (x->type == REALSXP) {
      p_ = static_cast<double*>(x->data);
    } else {
      own_.reset(new std::vector<double>((std::size_t)nrow_ * (std::size_t)ncol_));
      const int* src = static_cast<const int*>(x->data);
      for (std::size_t i = 0; i < own_->size(); ++i) (*own_)[i] = (double)src[i];
      p_ = own_->data();
    }
  }
  // Fresh zero-filled matrix (what R's allocMatrix + Rcpp's fill does).
  NumericMatrix(int nrow, int ncol)
      : nrow_(nrow), ncol_(ncol),
        own_(new std::vector<double>((std::size_t)nrow * (std::size_t)ncol, 0.0)),
        p_(own_->data()) {}
  // Harness-only: wrap caller-owned column-major memory without copying.
  static NumericMatrix wrap_external(double* p, int nrow, int ncol) {
    NumericMatrix m;
    m.nrow_ = nrow;
    m.ncol_ = ncol;
    m.p_ = p;
    return m;
  }
  int nrow() const { return nrow_; }
  int ncol() const { return ncol_; }
  int rows() const { return nrow_; }
  int cols() const { return ncol_; }
  double* begin() { return p_; }
  const double* begin() const { return p_; }
  double& operator()(std::size_t i, std::size_t j) { return p_[j * (std::size_t)nrow_ + i]; }
  const double& operator()(std::size_t i, std::size_t j) const {
    return p_[j * (std::size_t)nrow_ + i];
  }
  RowProxy operator()(std::size_t i, Placeholder) const {
    return RowProxy(p_ + i, (std::size_t)nrow_, (std::size_t)ncol_);
  }

 private:
  int nrow_, ncol_;
  std::shared_ptr<std::vector<double> > own_;  // shared like an R object handle
  double* p_;
};

// An integer matrix as R holds it (INTSXP): column-major int storage, wrapped without a copy.
class IntegerMatrix {
 public:
  IntegerMatrix(SEXP x) : nrow_(x->nrow), ncol_(x->ncol), p_(static_cast<int*>(x->data)) {  // NOLINT
    if (x->type != INTSXP) throw exception("not an integer matrix");
  }
  int nrow() const { return nrow_; }
  int ncol() const { return ncol_; }
  int* begin() { return p_; }
  const int* begin() const { return p_; }

 private:
  int nrow_, ncol_;
  int* p_;
};

// Distinct values of lhs that occur in rhs (set semantics; order unspecified,
// the reference only takes .size()).
inline NumericVector intersect(const NumericVector& lhs, const NumericVector& rhs) {
  std::unordered_set<double> a(lhs.begin(), lhs.end());
  std::unordered_set<double> b(rhs.begin(), rhs.end());
  NumericVector out;
  for (std::unordered_set<double>::const_iterator it = a.begin(); it != a.end(); ++it)
    if (b.count(*it)) out.push_back(*it);
  return out;
}

}  // namespace Rcpp

#endif  // GFICF_ORACLE_RSHIM_RCPP_H
