# testthat cases for the two Jaccard exports (the reference ships none: tests/testthat.R:4 has
# test_check() commented out).  Known answers derived by hand from
# src/rcpp_parallel_jaccard_coeff.cpp:24-55 and src/jaccard_coeff.cpp:28-42; the same vectors are
# checked against the reference's sources in this repository's tests/golden/.
# Not executed in the build environment of this repository (no R toolchain there).

test_that("every pair of lists sharing one id gives u = 1, w = 1/3", {
  idx <- matrix(c(2L, 1L, 1L, 1L,  3L, 3L, 2L, 2L), nrow = 4)   # N(1)={2,3} N(2)={1,3} N(3)={1,2} N(4)={1,2}
  r <- gficf:::rcpp_parallel_jaccard_coef(idx, FALSE)
  expect_equal(dim(r), c(8L, 3L))
  expect_identical(r[, 1], as.numeric(rep(1:4, each = 2)))
  expect_identical(r[, 2], as.numeric(c(2, 3, 1, 3, 1, 2, 1, 2)))
  expect_identical(r[, 3], rep(1 / (2 * 2 - 1), 8))
  expect_identical(gficf:::jaccard_coeff(idx, FALSE), r)        # nothing to compact here
})

test_that("disjoint lists leave a zero row (parallel) / are skipped (serial); identical lists give w = 1", {
  idx <- matrix(c(2, 1, 5, 5, 1, 1, 1,
                  3, 3, 6, 6, 2, 2, 2,
                  4, 4, 7, 7, 3, 3, 3), nrow = 7)
  p <- gficf:::rcpp_parallel_jaccard_coef(idx, FALSE)
  expect_identical(p[2 * 3 + 1, ], c(0, 0, 0))                  # cell 3 -> cell 5: {5,6,7} vs {1,2,3}
  expect_identical(p[4 * 3 + 1, ], c(5, 1, 0.5))                # cell 5 -> cell 1: u = 2 -> 2/(6-2)
  s <- gficf:::jaccard_coeff(idx, FALSE)
  m <- sum(s[, 3] > 0)
  expect_true(all(s[seq_len(m), 3] > 0) && all(s[-seq_len(m), ] == 0))
  expect_identical(s[seq_len(m), ], p[p[, 3] > 0, ])
  same <- matrix(rep(c(2, 3), each = 3), nrow = 3)              # every list is {2,3}
  expect_true(all(gficf:::rcpp_parallel_jaccard_coef(same, FALSE)[, 3] == 1))
})

test_that("repeated ids: multiset counts in the parallel export, distinct ids in the serial one", {
  idx <- matrix(c(2, 2, 1,  2, 3, 1), nrow = 3)                 # N(1)={2,2} N(2)={2,3} N(3)={1,1}
  p <- gficf:::rcpp_parallel_jaccard_coef(idx, FALSE)
  expect_identical(p[1, 3], 1 / 3)
  expect_identical(p[5, ], c(0, 0, 0))                          # {1,1} vs {2,2}
})

test_that("ids outside 1..nrow are an error instead of an out-of-bounds read", {
  idx <- matrix(c(2L, 1L, 9L, 1L), nrow = 2)
  expect_error(gficf:::rcpp_parallel_jaccard_coef(idx, FALSE), "gficf CUDA Jaccard failed")
})

test_that("the result does not depend on the number of GPUs", {
  skip_if(gficf:::gficf_cuda_visible_devices() < 2)
  set.seed(1)
  n <- 5000; k <- 15
  idx <- t(sapply(seq_len(n), function(i) sample(setdiff(seq_len(n), i), k)))
  gficf:::gficf_cuda_devices(1); a <- gficf:::rcpp_parallel_jaccard_coef(idx, FALSE)
  gficf:::gficf_cuda_devices(2); b <- gficf:::rcpp_parallel_jaccard_coef(idx, FALSE)
  gficf:::gficf_cuda_devices(1)
  expect_identical(a, b)
})
