// jaccard_weight.cuh -- the one floating-point operation of the Jaccard path, shared by the count / expand
// kernels (jaccard_kernels.cuh) and the graph build (snn_kernels.cuh).  PTX-free: with GFICF_CUDA_EMU
// it compiles as plain C++ (see scan_kernels.cuh).
#pragma once
#include "scan_kernels.cuh"

namespace gficf {

__device__ __forceinline__ double jaccard_weight(int u, int k) {
  // rcpp_parallel_jaccard_coeff.cpp:51  u/(2.0*mat.ncol() - u) : exact integer
  // operands, ONE IEEE-754 double division (correctly rounded on the device too).
  return __ddiv_rn((double)u, 2.0 * (double)k - (double)u);
}

}  // namespace gficf
