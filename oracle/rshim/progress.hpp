// oracle/rshim/progress.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Stand-in for RcppProgress's `Progress` (used at
// /root/reference/src/jaccard_coeff.cpp:27,41): a progress bar has no effect on
// results, so this one only counts.
#ifndef GFICF_ORACLE_RSHIM_PROGRESS_HPP
#define GFICF_ORACLE_RSHIM_PROGRESS_HPP

class Progress {
 public:
  Progress(unsigned long max, bool display) : max_(max), cur_(0), display_(display) {}
  bool increment(unsigned long by = 1) { cur_ += by; return true; }
  static bool check_abort() { return false; }

 private:
  unsigned long max_, cur_;
  bool display_;
};

#endif
