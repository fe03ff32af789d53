// host_expand.h -- host half of the "counts over PCIe" output mode of the host-buffer entry points.
//
// The reference's output row r = i*k+j is (i+1, idx(i,j), u/(2.0*k-u)) when u>0 and (0,0,0) otherwise
// (src/rcpp_parallel_jaccard_coeff.cpp:48-52).  Everything except u is already on the host: i is
// the row number, idx(i,j) is the caller's own matrix, and the weight takes one of k+1 values.  So
// instead of copying 24 bytes per edge back over PCIe, the GPU path may send the 1-byte count u
// computed by the CUDA kernels and these routines write the three output columns straight into the
// caller's (R-owned) matrix with non-temporal stores.  The division table is computed on the host
// with the reference's own expression, `u / (2.0 * k - u)` in IEEE double (:51), so the bytes are
// identical to the device's __ddiv_rn table and to the reference.
//
// This is output formatting of device-computed counts, not a CPU implementation of the path: the
// intersection counts never come from here.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace gficf_host {

struct ExpandJob {
  const void* idx;    // caller's n x k matrix, COLUMN-major, 1-based ids (validated by the device pre-pass)
  int elem;           // 8 = double (REALSXP), 4 = int32 (INTSXP)
  long long n;
  int k;
  const uint8_t* counts;  // u of edge (i,j) at counts[(i - counts_row0)*k + j]
  long long counts_row0;
  double* out;        // (n*k) x 3 column-major: from = out, to = out + E, weight = out + 2E
  long long E;
  const double* lut;  // k+1 weights, lut[u] = u / (2.0*k - u), lut[0] = 0
};

// k+1 doubles; the reference's expression with its operand types (int u, size_t ncol -> double)
void fill_weight_table(int k, double* lut);

// Fixed-slot rows [row_lo,row_hi) of ONE output column (0 = from, 1 = to, 2 = weight) of the
// parallel export.  Thread-safe for disjoint (column, row range) pieces.
void expand_column(const ExpandJob& job, int col, long long row_lo, long long row_hi);
// all three columns of the rows
void expand_rows(const ExpandJob& job, long long row_lo, long long row_hi);

// memcpy whose destination is written with non-temporal stores (large, write-once destinations)
void stream_copy(void* dst, const void* src, size_t bytes);

// ids held as doubles -> int32 (exact for valid ids; NaN / fractional / out-of-int32-range -> 0,
// which the device pre-pass rejects): the host half of the compressing H2D
void f64_to_i32(const double* src, int32_t* dst, size_t count);

// "avx512" / "avx2" / "scalar": which body expand_rows dispatches to on this CPU
const char* isa();

}  // namespace gficf_host
