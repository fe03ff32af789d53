// oracle/rshim/RcppGSL.h -- TEST INFRASTRUCTURE.  The reference includes RcppGSL.h only to reach
// GSL's headers (mann_whitney.cpp:8); nothing of RcppGSL itself is used on the path.
#ifndef GFICF_ORACLE_RSHIM_RCPPGSL_H
#define GFICF_ORACLE_RSHIM_RCPPGSL_H
#include <Rcpp.h>
#endif
