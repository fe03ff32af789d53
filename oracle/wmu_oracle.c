/* oracle/wmu_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's per-gene Mann-Whitney U worker, line by line:
 *   /root/reference/src/rcpp_parallel_mann_whitney.cpp:27-100  (WMU_test::operator())
 *   /root/reference/src/mann_whitney.cpp:17-29  sort_indexes   (sort indices by value)
 *                                        :31-54  getRanks      (ranks averaged over ties)
 *                                        :65-85  getCounts     (sizes of the tie groups, in order)
 *                                        :87-99  getSigma      (sigma with the tie correction)
 *                                        :101-110 getPvalue    (2 * normal cdf; GSL -> oracle/gauss_cdf.c)
 *                                        :128-131 avg
 * Matrices are R's: n_genes x n column-major doubles, element (g, c) at p[c * n_genes + g].
 * out: n_genes x 2 column-major -- column 0 the p-value, column 1 log2(mean(x+1) / mean(y+1)).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use this file.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

double gsl_cdf_gaussian_P(double x, double sigma);
double gsl_cdf_gaussian_Q(double x, double sigma);

typedef struct {
  double v;
  size_t i;
} Item;

static int cmp_item(const void* a, const void* b) {
  const double x = ((const Item*)a)->v, y = ((const Item*)b)->v;
  return (x < y) ? -1 : (y < x) ? 1 : 0; /* :26 comparator v[i1] < v[i2]; order inside ties is irrelevant */
}

static void one_gene(const double* mx, const double* my, int64_t G, int64_t n1, int64_t n2, int64_t k,
                     double* out) {
  const size_t N = (size_t)(n1 + n2);
  Item* it = (Item*)malloc(N * sizeof(Item));
  double* ord = (double*)malloc(N * sizeof(double));
  double* rnk = (double*)malloc(N * sizeof(double));
  double* cnt = (double*)malloc(N * sizeof(double));
  size_t i, j, q, ncnt = 0;
  double pval = 1;
  for (i = 0; i < (size_t)n1; i++) { it[i].v = mx[i * (size_t)G + (size_t)k]; it[i].i = i; }        /* :30-36 */
  for (i = 0; i < (size_t)n2; i++) { it[n1 + i].v = my[i * (size_t)G + (size_t)k]; it[n1 + i].i = (size_t)n1 + i; }
  qsort(it, N, sizeof(Item), cmp_item);                                                       /* :48 */
  for (i = 0; i < N; i++) ord[i] = it[i].v;                                                   /* :49-53 */
  i = 0;                                                                                      /* getRanks */
  while (i < N) {
    j = i + 1;
    while (j < N) {
      if (ord[i] != ord[j]) break;
      j++;
    }
    for (q = i; q <= j - 1; q++) rnk[q] = 1 + (double)(i + j - 1) / (double)2;
    i = j;
  }
  {                                                                                           /* getCounts */
    double prev = ord[0];
    size_t kk = 0;
    cnt[0] = 0;
    for (i = 0; i < N; i++) {
      if (prev == ord[i]) {
        cnt[kk]++;
      } else {
        kk++;
        cnt[kk] = 0;
        cnt[kk]++;
        prev = ord[i];
      }
    }
    ncnt = kk + 1;
  }
  if (ncnt > 1) {                                                                             /* :58 */
    long double U1 = ((size_t)n1 * ((size_t)n1 + 1)) * -0.5;
    long double U2 = ((size_t)n2 * ((size_t)n2 + 1)) * -0.5;
    double mu, sig, z, nties = 0, d1 = (double)n1, d2 = (double)n2;
    for (i = 0; i < N; i++) {
      if (it[i].i < (size_t)n1) U1 += rnk[i];
      else U2 += rnk[i];
    }
    mu = (double)(((size_t)n1 * (size_t)n2) / 2);                                             /* :86 size_t division */
    if (ncnt < d1 + d2)
      for (i = 0; i < ncnt; i++) nties += ((cnt[i] * cnt[i] * cnt[i]) - cnt[i]);
    sig = sqrt((d1 * d2 / 12) * ((d1 + d2 + 1) - nties / ((d1 + d2) * (d1 + d2 - 1))));
    z = U1 < U2 ? U1 - mu : U2 - mu;
    z = z < 0 ? z + 0.5 : z - 0.5;
    z = z / sig;
    pval = z < 0 ? gsl_cdf_gaussian_P(z, 1) * 2 : gsl_cdf_gaussian_Q(z, 1) * 2;
  }
  out[k] = pval;
  {                                                                                           /* :97-99 */
    double s1 = 0.0, s2 = 0.0;
    for (i = 0; i < (size_t)n1; i++) s1 = s1 + (mx[i * (size_t)G + (size_t)k] + 1.0);
    for (i = 0; i < (size_t)n2; i++) s2 = s2 + (my[i * (size_t)G + (size_t)k] + 1.0);
    out[(size_t)G + (size_t)k] = log2((double)((s1 / (size_t)n1) / (s2 / (size_t)n2)));
  }
  free(it); free(ord); free(rnk); free(cnt);
}

typedef struct {
  const double *mx, *my;
  int64_t G, n1, n2, lo, hi;
  double* out;
} Job;

static void* worker(void* p) {
  Job* j = (Job*)p;
  int64_t k;
  for (k = j->lo; k < j->hi; k++) one_gene(j->mx, j->my, j->G, j->n1, j->n2, k, j->out);
  return 0;
}

void gficf_oracle_wmu(const double* mx, const double* my, int64_t G, int64_t n1, int64_t n2, double* out,
                      int32_t nthreads) {
  int t;
  pthread_t th[256];
  Job jobs[256];
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (nthreads > G) nthreads = (int32_t)(G > 0 ? G : 1);
  for (t = 0; t < nthreads; t++) {
    Job j = {mx, my, G, n1, n2, G * t / nthreads, G * (t + 1) / nthreads, out};
    jobs[t] = j;
    pthread_create(&th[t], 0, worker, &jobs[t]);
  }
  for (t = 0; t < nthreads; t++) pthread_join(th[t], 0);
}
