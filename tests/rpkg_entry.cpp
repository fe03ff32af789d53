// tests/rpkg_entry.cpp -- TEST HARNESS.  C entry points around the drop-in R-package sources
// (gficf_b200/rpkg/src/*.cpp), compiled against the stand-in R runtime of oracle/rshim/ and
// linked with the product library, so the exact files a maintainer would put into gficf's src/
// are exercised end to end (R itself is not installed here).
#include <Rcpp.h>

#include <cstring>
#include <string>

Rcpp::NumericMatrix rcpp_parallel_jaccard_coef(SEXP mat, bool printOutput);
Rcpp::NumericMatrix jaccard_coeff(Rcpp::NumericMatrix idx, bool printOutput);
Rcpp::NumericMatrix rcpp_parallel_WMU_test(Rcpp::NumericMatrix matX, Rcpp::NumericMatrix matY, bool printOutput);
int gficf_cuda_devices(int n);
int gficf_cuda_visible_devices();

static std::vector<char> g_sink;

extern "C" {

// which: 0 = rcpp_parallel_jaccard_coef, 1 = jaccard_coeff, 2 = rcpp_parallel_jaccard_coef on an INTEGER
// matrix (idx then points at int32 data).  Returns 0, or 1 with the R error text.
int rpkg_call(int which, const void* idx, int n, int k, double* out, int print_output, char* err,
              int errlen, char* printed, int printedlen) {
  g_sink.clear();
  rshim::printf_sink() = &g_sink;
  int rc = 0;
  try {
    SEXPREC obj = {which == 2 ? INTSXP : REALSXP, n, k, const_cast<void*>(idx)};
    Rcpp::NumericMatrix res =
        which == 1 ? jaccard_coeff(Rcpp::NumericMatrix::wrap_external((double*)const_cast<void*>(idx), n, k),
                                   print_output != 0)
                   : rcpp_parallel_jaccard_coef(&obj, print_output != 0);
    std::memcpy(out, res.begin(), sizeof(double) * 3 * (size_t)n * (size_t)k);
  } catch (const std::exception& e) {
    snprintf(err, errlen, "%s", e.what());
    rc = 1;
  }
  rshim::printf_sink() = nullptr;
  int m = (int)g_sink.size() < printedlen - 1 ? (int)g_sink.size() : printedlen - 1;
  if (printed && printedlen > 0) {
    std::memcpy(printed, g_sink.data(), m);
    printed[m] = 0;
  }
  return rc;
}

// rcpp_parallel_WMU_test(matX, matY, printOutput): out is n_genes x 2 column-major
int rpkg_wmu(const double* x, const double* y, int genes, int n1, int n2, double* out, int print_output, char* err,
             int errlen, char* printed, int printedlen) {
  g_sink.clear();
  rshim::printf_sink() = &g_sink;
  int rc = 0;
  try {
    Rcpp::NumericMatrix mx = Rcpp::NumericMatrix::wrap_external(const_cast<double*>(x), genes, n1);
    Rcpp::NumericMatrix my = Rcpp::NumericMatrix::wrap_external(const_cast<double*>(y), genes, n2);
    Rcpp::NumericMatrix res = rcpp_parallel_WMU_test(mx, my, print_output != 0);
    std::memcpy(out, res.begin(), sizeof(double) * 2 * (size_t)genes);
  } catch (const std::exception& e) {
    snprintf(err, errlen, "%s", e.what());
    rc = 1;
  }
  rshim::printf_sink() = nullptr;
  int m = (int)g_sink.size() < printedlen - 1 ? (int)g_sink.size() : printedlen - 1;
  if (printed && printedlen > 0) {
    std::memcpy(printed, g_sink.data(), m);
    printed[m] = 0;
  }
  return rc;
}

int rpkg_devices(int n) {
  try {
    return gficf_cuda_devices(n);
  } catch (const std::exception&) {
    return -1;
  }
}
int rpkg_visible_devices() { return gficf_cuda_visible_devices(); }
}
