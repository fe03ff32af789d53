// host_stats.h -- the two transcendental steps of the Mann-Whitney path that stay on the host (one
// evaluation per gene, after the device has produced z and the mean ratio bit-exactly):
// the cumulative unit Gaussian the reference takes from GSL (src/mann_whitney.cpp:101-110) and log2.
#pragma once

namespace gficf_host {

// gsl_cdf_ugaussian_P / _Q: W. J. Cody's rational Chebyshev approximations (Math. Comp. 23, 1969),
// the algorithm GSL's cdf/gauss.c implements; libm exp() like GSL.
double ugaussian_P(double x);
double ugaussian_Q(double x);

// getPvalue (mann_whitney.cpp:101-110): two-sided p from the continuity-corrected z
double wmu_pvalue(double z);

}  // namespace gficf_host
