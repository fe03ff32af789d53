// Drop-in replacement body for gficf's src/rcpp_parallel_mann_whitney.cpp.
//
// Same exported R function, same three arguments, same n_genes x 2 result (p-value, log2 fold
// change) as the reference (src/rcpp_parallel_mann_whitney.cpp:106-127); the RcppParallel worker
// (:12-103) is gone: sorting, ranks, tie groups, U and sigma are computed per gene by libgficf_cuda
// (include/gficf_cuda.h, gficf_cuda_wmu_test) on the GPU.  src/mann_whitney.cpp keeps serving the
// serial rcpp_WMU_test and is not touched; GSL is no longer needed by THIS file.
//
// Compiled and executed against a stand-in R runtime by tests/test_rpkg_sources.py; R itself is not
// installed in this build environment.
#include <Rcpp.h>

#include "gficf_cuda.h"

// [[Rcpp::export]]
Rcpp::NumericMatrix rcpp_parallel_WMU_test(Rcpp::NumericMatrix matX, Rcpp::NumericMatrix matY, bool printOutput) {
  if (printOutput) Rprintf("Running Parallell WM-U test...\n");
  if (matX.nrow() != matY.nrow()) Rcpp::stop("matX and matY must have the same number of rows (genes)");
  Rcpp::NumericMatrix rmat(matX.nrow(), 2);
  char msg[512] = {0};
  const int rc = gficf_cuda_wmu_test(matX.begin(), matY.begin(), (int64_t)matX.nrow(), (int64_t)matX.ncol(),
                                     (int64_t)matY.ncol(), rmat.begin(), msg, sizeof msg);
  if (rc != GFICF_OK) Rcpp::stop("gficf CUDA Mann-Whitney failed (%d): %s", rc, msg);
  if (printOutput) Rprintf("Done!!\n");
  return rmat;
}
