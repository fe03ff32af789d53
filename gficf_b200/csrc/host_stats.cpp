// host_stats.cpp -- see host_stats.h.  Plain host C++.
#include "host_stats.h"

#include <math.h>

namespace gficf_host {
namespace {

const double kEps = 2.2204460492503131e-16 / 2;
const double kXUpper = 8.572, kXLower = -37.519, kScale = 16.0;
const double kSqrt32 = 4.0 * 1.41421356237309504880;
const double kInvSqrt2Pi = 0.39894228040143267794;

// exp(-x^2/2) * rational, with x split at a multiple of 1/16 so that the large part of the exponent
// is exact and only a small correction is exponentiated
double scaled_tail(double x, double rational) {
  const double xsq = floor(x * kScale) / kScale;
  double del = (x - xsq) * (x + xsq);
  del *= 0.5;
  return exp(-0.5 * xsq * xsq) * exp(-1.0 * del) * rational;
}

double centre(double x) {  // |x| < 0.66291
  static const double a[5] = {2.2352520354606839287, 161.02823106855587881, 1067.6894854603709582,
                              18154.981253343561249, 0.065682337918207449113};
  static const double b[4] = {47.20258190468824187, 976.09855173777669322, 10260.932208618978205,
                              45507.789335026729956};
  const double xsq = x * x;
  double num = a[4] * xsq, den = xsq;
  for (int i = 0; i < 3; ++i) {
    num = (num + a[i]) * xsq;
    den = (den + b[i]) * xsq;
  }
  return x * (num + a[3]) / (den + b[3]);
}

double middle(double x) {  // 0.66291 <= |x| < sqrt(32)
  static const double c[9] = {0.39894151208813466764, 8.8831497943883759412, 93.506656132177855979,
                              597.27027639480026226, 2494.5375852903726711, 6848.1904505362823326,
                              11602.651437647350124, 9842.7148383839780218, 1.0765576773720192317e-8};
  static const double d[8] = {22.266688044328115691, 235.38790178262499861, 1519.377599407554805,
                              6485.558298266760755, 18615.571640885098091, 34900.952721145977266,
                              38912.003286093271411, 19685.429676859990727};
  const double ax = fabs(x);
  double num = c[8] * ax, den = ax;
  for (int i = 0; i < 7; ++i) {
    num = (num + c[i]) * ax;
    den = (den + d[i]) * ax;
  }
  return scaled_tail(x, (num + c[7]) / (den + d[7]));
}

double far(double x) {  // sqrt(32) <= |x|
  static const double p[6] = {0.21589853405795699, 0.1274011611602473639, 0.022235277870649807,
                              0.001421619193227893466, 2.9112874951168792e-5, 0.02307344176494017303};
  static const double q[5] = {1.28426009614491121, 0.468238212480865118, 0.0659881378689285515,
                              0.00378239633202758244, 7.29751555083966205e-5};
  const double ax = fabs(x), xsq = 1.0 / (x * x);
  double num = p[5] * xsq, den = xsq;
  for (int i = 0; i < 4; ++i) {
    num = (num + p[i]) * xsq;
    den = (den + q[i]) * xsq;
  }
  double t = xsq * (num + p[4]) / (den + q[4]);
  t = (kInvSqrt2Pi - t) / ax;
  return scaled_tail(x, t);
}

}  // namespace

double ugaussian_P(double x) {
  const double ax = fabs(x);
  if (ax < kEps) return 0.5;
  if (ax < 0.66291) return 0.5 + centre(x);
  if (ax < kSqrt32) {
    const double r = middle(x);
    return x > 0.0 ? 1.0 - r : r;
  }
  if (x > kXUpper) return 1.0;
  if (x < kXLower) return 0.0;
  const double r = far(x);
  return x > 0.0 ? 1.0 - r : r;
}

double ugaussian_Q(double x) {
  const double ax = fabs(x);
  if (ax < kEps) return 0.5;
  if (ax < 0.66291) return 0.5 - centre(x);
  if (ax < kSqrt32) {
    const double r = middle(x);
    return x < 0.0 ? 1.0 - r : r;
  }
  if (x > -kXLower) return 0.0;
  if (x < -kXUpper) return 1.0;
  const double r = far(x);
  return x < 0.0 ? 1.0 - r : r;
}

double wmu_pvalue(double z) {
  // gsl_cdf_gaussian_P(z, 1) = ugaussian_P(z / 1)
  return z < 0 ? ugaussian_P(z / 1.0) * 2 : ugaussian_Q(z / 1.0) * 2;
}

}  // namespace gficf_host
