// oracle/ref_wmu_entry.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry point around the UNMODIFIED reference Mann-Whitney sources, compiled from where they lie
// under /root/reference/src (oracle/Makefile -> oracle/_ref/libgficf_ref_wmu.so):
//   mann_whitney.cpp (helpers :17-131) and rcpp_parallel_mann_whitney.cpp (worker :12-103,
//   export :106-127), pulled into ONE translation unit because the header only declares the
//   template sort_indexes (mann_whitney.h:9-10) whose definition lives in mann_whitney.cpp.
// The R runtime is oracle/rshim/; GSL's normal cdf (third-party, absent) is oracle/gauss_cdf.c --
// PARITY UNPINNED ON THE CDF, everything else is the reference's own arithmetic.
#include <Rcpp.h>
#include <RcppParallel.h>

#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

#include "mann_whitney.cpp"
#include "rcpp_parallel_mann_whitney.cpp"

extern "C" {

// rcpp_parallel_WMU_test(matX, matY, printOutput): matX n_genes x n1, matY n_genes x n2 column-major
// doubles; out n_genes x 2 column-major (p-value, log2FC).  Returns elapsed seconds.
double gficf_ref_wmu(const double* mx, const double* my, int64_t n_genes, int64_t n1, int64_t n2, double* out,
                     int32_t print_output, int32_t nthreads) {
  if (nthreads > 0) setenv("RCPP_PARALLEL_NUM_THREADS", std::to_string(nthreads).c_str(), 1);
  else unsetenv("RCPP_PARALLEL_NUM_THREADS");
  Rcpp::NumericMatrix X = Rcpp::NumericMatrix::wrap_external(const_cast<double*>(mx), (int)n_genes, (int)n1);
  Rcpp::NumericMatrix Y = Rcpp::NumericMatrix::wrap_external(const_cast<double*>(my), (int)n_genes, (int)n2);
  auto t0 = std::chrono::steady_clock::now();
  Rcpp::NumericMatrix res = rcpp_parallel_WMU_test(X, Y, print_output != 0);
  auto t1 = std::chrono::steady_clock::now();
  if (out) std::memcpy(out, res.begin(), sizeof(double) * 2 * (size_t)n_genes);
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
