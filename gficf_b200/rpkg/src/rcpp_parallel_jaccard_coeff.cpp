// Drop-in replacement body for gficf's src/rcpp_parallel_jaccard_coeff.cpp.
//
// Same exported R function, same two arguments, same (n*k) x 3 result as the reference
// (reference src/rcpp_parallel_jaccard_coeff.cpp:58-80); the RcppParallel worker (:10-56) is gone:
// the edges are computed by libgficf_cuda (include/gficf_cuda.h) on the GPU(s).
//
// The matrix is taken as the SEXP R passes.  uwot hands clustcells() an INTEGER matrix
// (R/clustCells.R:57-63); the reference's generated shim coerces it to double before the native code
// sees it (src/RcppExports.cpp:65, a full copy).  Here an integer matrix goes to the device as it is
// (gficf_cuda_jaccard_i32: no coercion copy, 4 bytes per id over PCIe); a double matrix takes
// gficf_cuda_jaccard (which narrows it to int32 on the way to the device); anything else is coerced to
// double like before.  Rcpp::compileAttributes() regenerates src/RcppExports.cpp with
// `input_parameter<SEXP>`; the registered symbol _gficf_rcpp_parallel_jaccard_coef and its arity (2)
// stay what they are (src/RcppExports.cpp:61,89).
//
// Compiled and executed against a stand-in R runtime by tests/test_rpkg_sources.py; R itself is not
// installed in this build environment.
#include <Rcpp.h>

#include "gficf_cuda.h"

// [[Rcpp::export]]
Rcpp::NumericMatrix rcpp_parallel_jaccard_coef(SEXP mat, bool printOutput) {
  if (printOutput) Rprintf("Running Parallell Jaccard Coefficient Estimation...\n");

  char msg[512] = {0};
  int rc;
  Rcpp::NumericMatrix edges;
  if (TYPEOF(mat) == INTSXP) {
    Rcpp::IntegerMatrix im(mat);
    const R_xlen_t n = im.nrow(), k = im.ncol();
    edges = Rcpp::NumericMatrix(n * k, 3);  // R-owned; the library writes every slot (zeros where u == 0)
    rc = gficf_cuda_jaccard_i32(im.begin(), (int64_t)n, (int32_t)k, edges.begin(),
                                /*n_devices=*/0 /* gficf_cuda_set_devices() / GFICF_CUDA_DEVICES */,
                                GFICF_MODE_PARALLEL, /*n_written=*/NULL, msg, sizeof msg);
  } else {
    Rcpp::NumericMatrix nm(mat);
    const R_xlen_t n = nm.nrow(), k = nm.ncol();
    edges = Rcpp::NumericMatrix(n * k, 3);
    rc = gficf_cuda_jaccard(nm.begin(), (int64_t)n, (int32_t)k, edges.begin(), 0, GFICF_MODE_PARALLEL, NULL, msg,
                            sizeof msg);
  }
  if (rc != GFICF_OK) Rcpp::stop("gficf CUDA Jaccard failed (%d): %s", rc, msg);

  if (printOutput) Rprintf("Done!!\n");
  return edges;
}
