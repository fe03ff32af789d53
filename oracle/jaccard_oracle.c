/* oracle/jaccard_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's Phenograph Jaccard edge weighting,
 * used ONLY as the checker by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  The product path (gficf_b200/csrc) never links, loads or
 * calls anything in this directory.
 *
 * Pinning: the reference's own tests pin nothing for this path
 * (tests/testthat.R:4 is commented out).  This restatement is instead pinned
 * against the UNMODIFIED reference sources compiled into oracle/_ref/ (R runtime
 * stubbed by oracle/rshim/) -- see tests/test_oracle.py and tests/golden/.
 *
 * Conventions (all from the reference):
 *   idx : n x k doubles, column-major, 1-based neighbour ids   (RcppExports.cpp:65)
 *   out : (n*k) x 3 doubles, column-major: from[], to[], w[]  (rcpp_parallel_jaccard_coeff.cpp:67)
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int cmp_double(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

/* One edge, multiset semantics.
 * Follows rcpp_parallel_jaccard_coeff.cpp:30-46: copy row i and row t out of the
 * column-major matrix (stride n), sort both ascending, count the merge
 * intersection (std::set_intersection counts min(multiplicity) per value). */
static int edge_count_multiset(const double* idx, int64_t n, int32_t k, int64_t i, int64_t t,
                               double* v1, double* v2) {
  for (int32_t c = 0; c < k; ++c) v1[c] = idx[(int64_t)c * n + i];
  for (int32_t c = 0; c < k; ++c) v2[c] = idx[(int64_t)c * n + t];
  qsort(v1, (size_t)k, sizeof(double), cmp_double);
  qsort(v2, (size_t)k, sizeof(double), cmp_double);
  int32_t a = 0, b = 0, u = 0;
  while (a < k && b < k) {
    if (v1[a] < v2[b]) ++a;
    else if (v2[b] < v1[a]) ++b;
    else { ++u; ++a; ++b; }
  }
  return u;
}

/* One edge, unique-set semantics.
 * Follows jaccard_coeff.cpp:31-33: Rcpp::intersect(nodei, nodej).size() is the
 * number of DISTINCT values common to both rows. */
static int edge_count_set(const double* idx, int64_t n, int32_t k, int64_t i, int64_t t,
                          double* v1, double* v2) {
  for (int32_t c = 0; c < k; ++c) v1[c] = idx[(int64_t)c * n + i];
  for (int32_t c = 0; c < k; ++c) v2[c] = idx[(int64_t)c * n + t];
  qsort(v1, (size_t)k, sizeof(double), cmp_double);
  qsort(v2, (size_t)k, sizeof(double), cmp_double);
  int32_t a = 0, b = 0, u = 0;
  while (a < k && b < k) {
    if (v1[a] < v2[b]) ++a;
    else if (v2[b] < v1[a]) ++b;
    else {
      double x = v1[a];
      ++u;
      while (a < k && v1[a] == x) ++a;
      while (b < k && v2[b] == x) ++b;
    }
  }
  return u;
}

/* rows [lo,hi) of the parallel export, fixed output slots.
 * Follows JCoefficient::operator() rcpp_parallel_jaccard_coeff.cpp:24-55.
 * out_slab is ((hi_total-lo_total)*k) x 3 column-major with slab stride slab_e;
 * slab row 0 corresponds to output row slab_lo*k. */
typedef struct {
  const double* idx;
  int64_t n;
  int32_t k;
  double* out;
  int64_t slab_lo, slab_e;
  int64_t lo, hi;
  int64_t* next;
  int64_t chunk;
  pthread_mutex_t* mu;
} job_t;

static void rows_fixed(const job_t* jb, int64_t lo, int64_t hi, double* v1, double* v2) {
  const int32_t k = jb->k;
  for (int64_t i = lo; i < hi; ++i) {
    for (int32_t j = 0; j < k; ++j) {
      int t = (int)(jb->idx[(int64_t)j * jb->n + i] - 1); /* :28  int k = mat(i,j)-1 */
      int u = edge_count_multiset(jb->idx, jb->n, k, i, (int64_t)t, v1, v2);
      if (u > 0) { /* :48-52 */
        int64_t r = (i - jb->slab_lo) * (int64_t)k + j;
        jb->out[r] = (double)(i + 1);
        jb->out[jb->slab_e + r] = (double)(t + 1);
        jb->out[2 * jb->slab_e + r] = u / (2.0 * k - u);
      }
    }
  }
}

static void* worker_main(void* arg) {
  job_t* jb = (job_t*)arg;
  double* v1 = (double*)malloc(sizeof(double) * (size_t)jb->k);
  double* v2 = (double*)malloc(sizeof(double) * (size_t)jb->k);
  for (;;) {
    pthread_mutex_lock(jb->mu);
    int64_t lo = *jb->next;
    *jb->next = lo + jb->chunk;
    pthread_mutex_unlock(jb->mu);
    if (lo >= jb->hi) break;
    int64_t hi = lo + jb->chunk < jb->hi ? lo + jb->chunk : jb->hi;
    rows_fixed(jb, lo, hi, v1, v2);
  }
  free(v1);
  free(v2);
  return NULL;
}

/* Parallel export, rows [row_lo,row_hi), gathering from the full matrix.
 * out_slab must be zero-filled by the caller (the reference allocates a
 * zero-filled matrix, :67, and leaves u==0 rows untouched). */
void gficf_oracle_parallel_jaccard_rows(const double* idx, int64_t n, int32_t k, int64_t row_lo,
                                        int64_t row_hi, double* out_slab, int32_t nthreads) {
  if (row_hi <= row_lo || k <= 0) return;
  if (nthreads < 1) nthreads = 1;
  int64_t total = row_hi - row_lo;
  if ((int64_t)nthreads > total) nthreads = (int32_t)total;
  int64_t next = row_lo;
  pthread_mutex_t mu;
  pthread_mutex_init(&mu, NULL);
  job_t jb;
  jb.idx = idx; jb.n = n; jb.k = k; jb.out = out_slab;
  jb.slab_lo = row_lo; jb.slab_e = total * (int64_t)k;
  jb.lo = row_lo; jb.hi = row_hi; jb.next = &next; jb.mu = &mu;
  jb.chunk = total / ((int64_t)nthreads * 64);
  if (jb.chunk < 1) jb.chunk = 1;
  if (nthreads == 1) {
    worker_main(&jb);
  } else {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int32_t t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, worker_main, &jb);
    for (int32_t t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
  }
  pthread_mutex_destroy(&mu);
}

/* Whole matrix, parallel export (rcpp_parallel_jaccard_coeff.cpp:59-80). `out` is
 * fully written here (zero-filled first, as the reference's fresh matrix is). */
void gficf_oracle_parallel_jaccard(const double* idx, int64_t n, int32_t k, double* out,
                                   int32_t nthreads) {
  memset(out, 0, sizeof(double) * 3 * (size_t)n * (size_t)k);
  gficf_oracle_parallel_jaccard_rows(idx, n, k, 0, n, out, nthreads);
}

/* Whole matrix, serial export (jaccard_coeff.cpp:19-45): rows compacted (r++ only
 * when u>0, :34-39), unique-set intersection (:33), trailing rows stay zero.
 * Returns the number of rows written. */
int64_t gficf_oracle_serial_jaccard(const double* idx, int64_t n, int32_t k, double* out) {
  const int64_t e = n * (int64_t)k;
  memset(out, 0, sizeof(double) * 3 * (size_t)e);
  double* v1 = (double*)malloc(sizeof(double) * (size_t)(k > 0 ? k : 1));
  double* v2 = (double*)malloc(sizeof(double) * (size_t)(k > 0 ? k : 1));
  int64_t r = 0;
  for (int64_t i = 0; i < n; ++i) {
    for (int32_t j = 0; j < k; ++j) {
      int t = (int)(idx[(int64_t)j * n + i] - 1); /* :30 */
      int u = edge_count_set(idx, n, k, i, (int64_t)t, v1, v2);
      if (u > 0) { /* :34-39 */
        out[r] = (double)(i + 1);
        out[e + r] = (double)(t + 1);
        out[2 * e + r] = u / (2.0 * k - u);
        ++r;
      }
    }
  }
  free(v1);
  free(v2);
  return r;
}
