// snn_kernels.cuh -- the step right after the Jaccard path ("next" row 1):
// from the per-edge counts to the graph the community detection reads.
//
// What it replaces in the reference (all single-threaded R / igraph / C++ on the host):
//   R/clustCells.R:66      relations <- relations[relations[,3] > 0, ]
//   R/clustCells.R:67-69   igraph::graph.data.frame(relations, directed = FALSE)
//   R/clustCells.R:81      as_adjacency_matrix(g, attr = "weight")   (parallel edges i->j, j->i SUMMED)
//   src/RModularityOptimizer.cpp:67-83   strictly-lower-triangle scan in column order
// Result: CSC of the strictly lower triangle (colptr[n+1], row[], weight[]), rows ascending
// within a column -- node1 = column, node2 = row, exactly the order matrixToNetwork
// (src/ModularityOptimizer.cpp:761-806) is fed with.
//
// Input: the count byte of every edge slot with bit 7 = "mutual" (i is in N(t) too), produced by
// the count kernels in OUT==2 mode.  For rows without repeated ids u(i,t) == u(t,i), so an
// undirected pair {i,t} carries w (one direction) or w + w (both), and the direction that emits it
// is: the one from the smaller id if both exist, else the only one.
// Supported domain (flags otherwise): k <= 127, no repeated ids, every cell has at least one edge
// with u > 0 (then igraph's first-appearance vertex numbering is the cell numbering).
#pragma once
#include "jaccard_kernels.cuh"

namespace gficf {

constexpr unsigned kFlagIsolated = 16u;  // some cell has no edge with u>0: vertex numbering differs

// one staged entry of pass 2: a single 16-byte store per entry (the scattered ones land at random
// positions, three separate 4/4/8-byte stores tripled the partial-sector writes)
struct __align__(16) SnnEntry {
  int col, row;
  double w;
};

// pass 1 (count) / pass 2 (scatter): one warp per row
template <bool SCATTER>
__global__ void __launch_bounds__(256)
snn_edges_kernel(const int* __restrict__ idx, const uint8_t* __restrict__ um, long long n, int k, int kp,
                 int* __restrict__ cnt_or_cursor, const long long* __restrict__ colptr,
                 SnnEntry* __restrict__ tmp, unsigned* __restrict__ flags) {
  __shared__ double lut[128];
  if ((int)threadIdx.x <= k && threadIdx.x < 128) lut[threadIdx.x] = jaccard_weight((int)threadIdx.x, k);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  bool isolated = false;
  for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nwarps) {
    bool any_nz = false;
    for (int j0 = 0; j0 < k; j0 += 32) {
      const int j = j0 + lane;
      int u = 0, t = 0;
      bool mut = false;
      if (j < k) {
        const unsigned b = um[i * (long long)k + j];
        u = (int)(b & 0x7Fu);
        mut = (b & 0x80u) != 0;
        t = __ldg(idx + i * (long long)kp + j);
      }
      const bool nz = u > 0;
      any_nz |= nz;
      const bool off_diag = nz && (long long)t != i;
      const bool own = off_diag && (long long)t > i;            // emitted in this row's own column i
      const bool other = off_diag && (long long)t < i && !mut;  // one-directional edge into column t
      const unsigned own_mask = __ballot_sync(kFull, own);
      if (!SCATTER) {
        if (lane == 0 && own_mask) atomicAdd(cnt_or_cursor + i, __popc(own_mask));
        if (other) atomicAdd(cnt_or_cursor + t, 1);
      } else {
        int base = 0;
        if (lane == 0 && own_mask) base = atomicAdd(cnt_or_cursor + i, __popc(own_mask));
        base = __shfl_sync(kFull, base, 0);
        long long pos = -1;
        int col = 0, row = 0;
        if (own) {
          pos = colptr[i] + base + __popc(own_mask & ((1u << lane) - 1u));
          col = (int)i;
          row = t;
        } else if (other) {
          pos = colptr[t] + atomicAdd(cnt_or_cursor + t, 1);
          col = t;
          row = (int)i;
        }
        if (pos >= 0) {
          const double w = lut[u];
          SnnEntry en;
          en.col = col;
          en.row = row;
          en.w = mut ? w + w : w;  // both directions present: the two equal weights summed
          tmp[pos] = en;
        }
      }
    }
    if (!__any_sync(kFull, any_nz)) isolated = true;
  }
  if (!SCATTER && isolated && lane == 0) atomicOr(flags, kFlagIsolated);
}

// exclusive scan of int32 counts into int64 offsets: per-block sums, scan of the sums
// (compact_scan_kernel), block-local scan + offset
constexpr int kScanBlock = 1024;

__global__ void __launch_bounds__(kScanBlock)
scan_block_sums_kernel(const int* __restrict__ cnt, long long n, long long* __restrict__ block_sums) {
  __shared__ long long ws[32];
  const long long x = (long long)blockIdx.x * kScanBlock + threadIdx.x;
  long long v = x < n ? cnt[x] : 0;
  for (int m = 16; m; m >>= 1) v += __shfl_xor_sync(kFull, v, m);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    long long s = ws[threadIdx.x];
    for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(kFull, s, m);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(kScanBlock)
scan_finish_kernel(const int* __restrict__ cnt, long long n, const long long* __restrict__ block_off,
                   const long long* __restrict__ total, long long* __restrict__ colptr) {
  __shared__ long long ws[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long x = (long long)blockIdx.x * kScanBlock + threadIdx.x;
  const long long v = x < n ? cnt[x] : 0;
  long long inc = v;
  for (int m = 1; m < 32; m <<= 1) {
    const long long o = __shfl_up_sync(kFull, inc, m);
    if (lane >= m) inc += o;
  }
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    long long s = ws[lane];
    for (int m = 1; m < 32; m <<= 1) {
      const long long o = __shfl_up_sync(kFull, s, m);
      if (lane >= m) s += o;
    }
    ws[lane] = s;
  }
  __syncthreads();
  if (x < n) colptr[x] = block_off[blockIdx.x] + (warp ? ws[warp - 1] : 0) + inc - v;
  if (x == 0) colptr[n] = *total;
}

// rank sort inside every column: one thread per entry counts the entries of its column with a
// smaller row (rows are distinct inside a column) and writes itself at that rank.  Threads of a
// warp mostly share a column, so the scans are broadcast reads.
__global__ void __launch_bounds__(256)
snn_rank_sort_kernel(const long long* __restrict__ colptr, const long long* __restrict__ nnz,
                     const SnnEntry* __restrict__ tmp, int* __restrict__ row_out,
                     double* __restrict__ w_out) {
  const long long total = *nnz;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    const SnnEntry en = tmp[p];
    const long long lo = colptr[en.col], hi = colptr[en.col + 1];
    int rank = 0;
    for (long long q = lo; q < hi; ++q) rank += tmp[q].row < en.row;
    row_out[lo + rank] = en.row;
    w_out[lo + rank] = en.w;
  }
}

}  // namespace gficf
