// Drop-in replacement body for gficf's src/rcpp_parallel_jaccard_coeff.cpp.
//
// Same exported symbol, same signature, same (n*k) x 3 result as the reference
// (reference src/rcpp_parallel_jaccard_coeff.cpp:58-80); the RcppParallel worker
// (:10-56) is gone: the edges are computed by libgficf_cuda (include/gficf_cuda.h)
// on the GPU(s).  The generated shim src/RcppExports.cpp:61-70 and R/RcppExports.R:16-18
// stay byte-for-byte what Rcpp::compileAttributes() produced.
//
// Reviewed, not executed here: this build environment has no R toolchain.
#include <Rcpp.h>

#include "gficf_cuda.h"

// [[Rcpp::export]]
Rcpp::NumericMatrix rcpp_parallel_jaccard_coef(Rcpp::NumericMatrix mat, bool printOutput) {
  if (printOutput) Rprintf("Running Parallell Jaccard Coefficient Estimation...\n");

  const R_xlen_t n = mat.nrow(), k = mat.ncol();
  // R-owned, zero-filled result; the library overwrites every slot (zeros where u == 0)
  Rcpp::NumericMatrix edges(n * k, 3);

  char msg[512] = {0};
  const int rc = gficf_cuda_jaccard(mat.begin(), (int64_t)n, (int32_t)k, edges.begin(),
                                    /*n_devices=*/0 /* gficf_cuda_set_devices() / GFICF_CUDA_DEVICES */,
                                    GFICF_MODE_PARALLEL, /*n_written=*/NULL, msg, sizeof msg);
  if (rc != GFICF_OK) Rcpp::stop("gficf CUDA Jaccard failed (%d): %s", rc, msg);

  if (printOutput) Rprintf("Done!!\n");
  return edges;
}
