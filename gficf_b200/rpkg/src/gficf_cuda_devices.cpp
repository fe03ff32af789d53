// NEW routine (added to the registration table, nothing existing changes arity):
// lets R/clustCells.R's `n.gpu` argument choose how many GPUs of the node the Jaccard
// step shards its cell rows over.  After adding this file run Rcpp::compileAttributes()
// so that src/RcppExports.cpp / R/RcppExports.R gain `_gficf_gficf_cuda_devices`.
#include <Rcpp.h>

#include "gficf_cuda.h"

// [[Rcpp::export]]
int gficf_cuda_devices(int n) {
  if (n > 0 && gficf_cuda_set_devices(n) != GFICF_OK) Rcpp::stop("invalid GPU count %d", n);
  return gficf_cuda_get_devices();
}

// [[Rcpp::export]]
int gficf_cuda_visible_devices() { return gficf_cuda_device_count(); }
