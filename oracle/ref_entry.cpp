// oracle/ref_entry.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry points around the UNMODIFIED reference sources, compiled from where
// they lie under /root/reference/src (see oracle/Makefile; outputs go to
// oracle/_ref/ only).  The reference files are pulled in by #include so that
// the worker type `JCoefficient` (rcpp_parallel_jaccard_coeff.cpp:10-56) is
// visible and can be driven over a row sub-range for bounded CPU-baseline
// samples; the R runtime they expect is provided by oracle/rshim/*.h.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
// --impl reference legs may load the resulting library.
#include <Rcpp.h>
#include <RcppParallel.h>

#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

// The reference translation unit itself (path resolved by -I$(REF)/src).
#include "rcpp_parallel_jaccard_coeff.cpp"

// Defined in the reference's jaccard_coeff.cpp (compiled as its own TU).
Rcpp::NumericMatrix jaccard_coeff(Rcpp::NumericMatrix idx, bool printOutput);

namespace {
void set_threads(int nthreads) {
  if (nthreads > 0)
    setenv("RCPP_PARALLEL_NUM_THREADS", std::to_string(nthreads).c_str(), 1);
  else
    unsetenv("RCPP_PARALLEL_NUM_THREADS");
}
}  // namespace

extern "C" {

// Whole-matrix call of the reference export
// rcpp_parallel_jaccard_coef(mat, printOutput)  (rcpp_parallel_jaccard_coeff.cpp:59-80).
// idx: n x k column-major doubles, 1-based.  out: (n*k) x 3 column-major.
// Returns elapsed seconds of the reference call (allocation included, as in R).
double gficf_ref_parallel_jaccard(const double* idx, int64_t n, int32_t k, double* out,
                                  int32_t print_output, int32_t nthreads) {
  set_threads(nthreads);
  Rcpp::NumericMatrix mat =
      Rcpp::NumericMatrix::wrap_external(const_cast<double*>(idx), (int)n, (int)k);
  auto t0 = std::chrono::steady_clock::now();
  Rcpp::NumericMatrix res = rcpp_parallel_jaccard_coef(mat, print_output != 0);
  auto t1 = std::chrono::steady_clock::now();
  if (out) std::memcpy(out, res.begin(), sizeof(double) * 3 * (size_t)n * (size_t)k);
  return std::chrono::duration<double>(t1 - t0).count();
}

// Bounded sample: the reference worker JCoefficient::operator() (:24-55) run by
// the parallelFor stand-in over rows [row_lo,row_hi) only, gathering from the
// full matrix.  out_slab: ((row_hi-row_lo)*k) x 3 column-major, caller-zeroed.
double gficf_ref_parallel_jaccard_rows(const double* idx, int64_t n, int32_t k, int64_t row_lo,
                                       int64_t row_hi, double* out_slab, int32_t nthreads) {
  set_threads(nthreads);
  Rcpp::NumericMatrix mat =
      Rcpp::NumericMatrix::wrap_external(const_cast<double*>(idx), (int)n, (int)k);
  const int64_t slab_e = (row_hi - row_lo) * (int64_t)k;
  // The worker addresses output row i*k+j; bias the base so that row_lo*k maps to 0.
  Rcpp::NumericMatrix rmat =
      Rcpp::NumericMatrix::wrap_external(out_slab - row_lo * (int64_t)k, (int)slab_e, 3);
  JCoefficient worker(mat, rmat);
  auto t0 = std::chrono::steady_clock::now();
  RcppParallel::parallelFor((size_t)row_lo, (size_t)row_hi, worker);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// Whole-matrix call of the reference's serial export jaccard_coeff(idx, printOutput)
// (jaccard_coeff.cpp:19-45): compacted rows, unique-set semantics.
double gficf_ref_serial_jaccard(const double* idx, int64_t n, int32_t k, double* out,
                                int32_t print_output) {
  Rcpp::NumericMatrix mat =
      Rcpp::NumericMatrix::wrap_external(const_cast<double*>(idx), (int)n, (int)k);
  auto t0 = std::chrono::steady_clock::now();
  Rcpp::NumericMatrix res = jaccard_coeff(mat, print_output != 0);
  auto t1 = std::chrono::steady_clock::now();
  if (out) std::memcpy(out, res.begin(), sizeof(double) * 3 * (size_t)n * (size_t)k);
  return std::chrono::duration<double>(t1 - t0).count();
}

// Capture what the reference printed through Rprintf during the next calls.
static std::vector<char> g_sink;
void gficf_ref_capture_begin() {
  g_sink.clear();
  rshim::printf_sink() = &g_sink;
}
int64_t gficf_ref_capture_end(char* buf, int64_t buflen) {
  rshim::printf_sink() = nullptr;
  int64_t m = (int64_t)g_sink.size();
  if (buf && buflen > 0) {
    int64_t c = m < buflen - 1 ? m : buflen - 1;
    std::memcpy(buf, g_sink.data(), (size_t)c);
    buf[c] = 0;
  }
  return m;
}

int32_t gficf_ref_hw_threads() { return RcppParallel::resolved_threads(); }

}  // extern "C"
