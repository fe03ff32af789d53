// oracle/rshim/gsl/gsl_cdf.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Stand-in for the one GSL header the reference's Mann-Whitney code includes
// (/root/reference/src/mann_whitney.cpp:9, used at :105-107).  GSL is a third-party dependency that
// is absent from this image and from /root/reference (DESCRIPTION: LinkingTo RcppGSL, version
// unpinned); the two functions are restated in oracle/gauss_cdf.c from the published algorithm GSL
// implements (W. J. Cody's rational Chebyshev approximations, cdf/gauss.c).  PARITY UNPINNED ON THE
// CDF: no GSL build exists here to compare bits with; the restatement is checked against
// scipy.special.ndtr to a few ulp (tests/test_wmu_oracle.py).
#ifndef GFICF_ORACLE_RSHIM_GSL_CDF_H
#define GFICF_ORACLE_RSHIM_GSL_CDF_H
#ifdef __cplusplus
extern "C" {
#endif
double gsl_cdf_gaussian_P(double x, double sigma);
double gsl_cdf_gaussian_Q(double x, double sigma);
double gsl_cdf_ugaussian_P(double x);
double gsl_cdf_ugaussian_Q(double x);
#ifdef __cplusplus
}
#endif
#endif
