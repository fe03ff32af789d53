/* oracle/gauss_cdf.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restatement of GSL's cumulative unit Gaussian (gsl_cdf_ugaussian_P / _Q, GSL cdf/gauss.c), the
 * third-party function the reference's Mann-Whitney p-value goes through
 * (/root/reference/src/mann_whitney.cpp:101-110, getPvalue: gsl_cdf_gaussian_P(z,1)*2 for z<0,
 * gsl_cdf_gaussian_Q(z,1)*2 otherwise).  GSL is absent here (no source, no library, no network), so
 * this follows the published algorithm it implements: W. J. Cody, "Rational Chebyshev approximations
 * for the error function", Math. Comp. 23 (1969) 631-637 -- three ranges (|x| < 0.66291,
 * < sqrt(32), beyond) with the rational coefficients of that paper, and the exp(-x^2/2) factor
 * split as exp(-xsq^2/2) * exp(-del) with xsq = x truncated to 1/16 to avoid cancellation.
 * PARITY UNPINNED ON THE CDF (see oracle/rshim/gsl/gsl_cdf.h).
 */
#include <math.h>

#define GAUSS_EPSILON (2.2204460492503131e-16 / 2)
#define GAUSS_XUPPER (8.572)
#define GAUSS_XLOWER (-37.519)
#define GAUSS_SCALE (16.0)
#define SQRT32 (4.0 * 1.41421356237309504880)
#define M_1_SQRT2PI 0.39894228040143267794

static double get_del(double x, double rational) {
  double xsq = floor(x * GAUSS_SCALE) / GAUSS_SCALE;
  double del = (x - xsq) * (x + xsq);
  del *= 0.5;
  return exp(-0.5 * xsq * xsq) * exp(-1.0 * del) * rational;
}

/* |x| < 0.66291 */
static double gauss_small(const double x) {
  static const double a[5] = {2.2352520354606839287, 161.02823106855587881, 1067.6894854603709582,
                              18154.981253343561249, 0.065682337918207449113};
  static const double b[4] = {47.20258190468824187, 976.09855173777669322, 10260.932208618978205,
                              45507.789335026729956};
  unsigned int i;
  double xsq = x * x, xnum = a[4] * xsq, xden = xsq;
  for (i = 0; i < 3; i++) {
    xnum = (xnum + a[i]) * xsq;
    xden = (xden + b[i]) * xsq;
  }
  return x * (xnum + a[3]) / (xden + b[3]);
}

/* 0.66291 <= |x| < sqrt(32) */
static double gauss_medium(const double x) {
  static const double c[9] = {0.39894151208813466764, 8.8831497943883759412, 93.506656132177855979,
                              597.27027639480026226, 2494.5375852903726711, 6848.1904505362823326,
                              11602.651437647350124, 9842.7148383839780218, 1.0765576773720192317e-8};
  static const double d[8] = {22.266688044328115691, 235.38790178262499861, 1519.377599407554805,
                              6485.558298266760755, 18615.571640885098091, 34900.952721145977266,
                              38912.003286093271411, 19685.429676859990727};
  unsigned int i;
  double absx = fabs(x), xnum = c[8] * absx, xden = absx, temp;
  for (i = 0; i < 7; i++) {
    xnum = (xnum + c[i]) * absx;
    xden = (xden + d[i]) * absx;
  }
  temp = (xnum + c[7]) / (xden + d[7]);
  return get_del(x, temp);
}

/* sqrt(32) <= |x| */
static double gauss_large(const double x) {
  static const double p[6] = {0.21589853405795699, 0.1274011611602473639, 0.022235277870649807,
                              0.001421619193227893466, 2.9112874951168792e-5, 0.02307344176494017303};
  static const double q[5] = {1.28426009614491121, 0.468238212480865118, 0.0659881378689285515,
                              0.00378239633202758244, 7.29751555083966205e-5};
  int i;
  double absx = fabs(x), xsq = 1.0 / (x * x), xnum = p[5] * xsq, xden = xsq, temp;
  for (i = 0; i < 4; i++) {
    xnum = (xnum + p[i]) * xsq;
    xden = (xden + q[i]) * xsq;
  }
  temp = xsq * (xnum + p[4]) / (xden + q[4]);
  temp = (M_1_SQRT2PI - temp) / absx;
  return get_del(x, temp);
}

double gsl_cdf_ugaussian_P(const double x) {
  double result, absx = fabs(x);
  if (absx < GAUSS_EPSILON) return 0.5;
  if (absx < 0.66291) return 0.5 + gauss_small(x);
  if (absx < SQRT32) {
    result = gauss_medium(x);
    if (x > 0.0) result = 1.0 - result;
    return result;
  }
  if (x > GAUSS_XUPPER) return 1.0;
  if (x < GAUSS_XLOWER) return 0.0;
  result = gauss_large(x);
  if (x > 0.0) result = 1.0 - result;
  return result;
}

double gsl_cdf_ugaussian_Q(const double x) {
  double result, absx = fabs(x);
  if (absx < GAUSS_EPSILON) return 0.5;
  if (absx < 0.66291) return 0.5 - gauss_small(x);
  if (absx < SQRT32) {
    result = gauss_medium(x);
    if (x < 0.0) result = 1.0 - result;
    return result;
  }
  if (x > -(GAUSS_XLOWER)) return 0.0;
  if (x < -(GAUSS_XUPPER)) return 1.0;
  result = gauss_large(x);
  if (x < 0.0) result = 1.0 - result;
  return result;
}

double gsl_cdf_gaussian_P(const double x, const double sigma) { return gsl_cdf_ugaussian_P(x / sigma); }
double gsl_cdf_gaussian_Q(const double x, const double sigma) { return gsl_cdf_ugaussian_Q(x / sigma); }
