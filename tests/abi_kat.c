/* tests/abi_kat.c -- plain C99 client of include/gficf_cuda.h (what a cgo / .Call / FFI stub sees).
 * Runs the hand-derived known-answer test of SURVEY appendix B through gficf_cuda_jaccard.
 * exit 0: answers match; exit 3: no CUDA device (GFICF_E_CUDA, expected on CPU-only hosts);
 * anything else: failure. */
#include <stdio.h>
#include <string.h>

#include "gficf_cuda.h"

int main(void) {
  /* n=4, k=2, 1-based, column-major: N(1)={2,3} N(2)={1,3} N(3)={1,2} N(4)={1,2} */
  const double idx[8] = {2, 1, 1, 1, /* column 2 */ 3, 3, 2, 2};
  double out[24];
  char err[256];
  int64_t written = 123;
  int rc, r;
  const double third = 1.0 / (2.0 * 2 - 1);
  memset(out, 0xff, sizeof out);
  rc = gficf_cuda_jaccard(idx, 4, 2, out, 1, GFICF_MODE_PARALLEL, &written, err, sizeof err);
  if (rc == GFICF_E_CUDA) {
    printf("no device: %s\n", err);
    return 3;
  }
  if (rc != GFICF_OK) {
    printf("error %d: %s\n", rc, err);
    return 1;
  }
  /* every pair of lists shares exactly one id -> u=1 -> w=1/3 on all 8 edges */
  for (r = 0; r < 8; ++r) {
    const double from = (double)(r / 2 + 1);
    const double to = idx[(r % 2) * 4 + r / 2];
    if (out[r] != from || out[8 + r] != to || out[16 + r] != third) {
      printf("edge %d: got (%g,%g,%.17g) want (%g,%g,%.17g)\n", r, out[r], out[8 + r], out[16 + r], from, to, third);
      return 2;
    }
  }
  if (written != -1) return 4;
  /* the serial export on the same input: all rows emitted, r = 8 */
  rc = gficf_cuda_jaccard(idx, 4, 2, out, 1, GFICF_MODE_SERIAL, &written, err, sizeof err);
  if (rc != GFICF_OK || written != 8 || out[16 + 7] != third) return 5;
  /* bad id -> GFICF_E_RANGE with a message */
  {
    double bad[8];
    memcpy(bad, idx, sizeof bad);
    bad[3] = 9;
    rc = gficf_cuda_jaccard(bad, 4, 2, out, 1, GFICF_MODE_PARALLEL, NULL, err, sizeof err);
    if (rc != GFICF_E_RANGE || !strlen(err)) return 6;
  }
  printf("abi_kat ok (%s)\n", gficf_cuda_version());
  return 0;
}
