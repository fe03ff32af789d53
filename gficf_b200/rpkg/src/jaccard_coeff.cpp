// Drop-in replacement body for gficf's src/jaccard_coeff.cpp (the serial export,
// reference src/jaccard_coeff.cpp:19-45): compacted rows (only edges with a non-empty
// intersection are emitted, in (i,j) order, the tail of the matrix stays zero) and
// unique-set intersection counts, computed by libgficf_cuda in GFICF_MODE_SERIAL.
// The RcppProgress bar of the reference (:27,41) has nothing left to report and is dropped.
//
// Reviewed, not executed here: this build environment has no R toolchain.
#include <Rcpp.h>

#include "gficf_cuda.h"

// [[Rcpp::export]]
Rcpp::NumericMatrix jaccard_coeff(Rcpp::NumericMatrix idx, bool printOutput) {
  if (printOutput) Rprintf("Running Jaccard Coefficient Estimation...\n");
  const R_xlen_t n = idx.nrow(), k = idx.ncol();
  Rcpp::NumericMatrix weights(n * k, 3);
  char msg[512] = {0};
  int64_t emitted = 0;
  const int rc = gficf_cuda_jaccard(idx.begin(), (int64_t)n, (int32_t)k, weights.begin(), 1,
                                    GFICF_MODE_SERIAL, &emitted, msg, sizeof msg);
  if (rc != GFICF_OK) Rcpp::stop("gficf CUDA Jaccard failed (%d): %s", rc, msg);
  return weights;
}
