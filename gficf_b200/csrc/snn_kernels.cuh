// snn_kernels.cuh -- the step right after the Jaccard path ("next" row 1):
// from the per-edge counts to the graph the community detection reads.
//
// What it replaces in the reference (all single-threaded R / igraph / C++ on the host):
//   R/clustCells.R:66      relations <- relations[relations[,3] > 0, ]
//   R/clustCells.R:67-69   igraph::graph.data.frame(relations, directed = FALSE)
//   R/clustCells.R:81      as_adjacency_matrix(g, attr = "weight")   (parallel edges i->j, j->i SUMMED)
//   src/RModularityOptimizer.cpp:67-83   strictly-lower-triangle scan in column order
// Result: CSC of the strictly lower triangle (colptr[nv+1], row[], weight[]), rows ascending
// within a column -- node1 = column, node2 = row, exactly the order matrixToNetwork
// (src/ModularityOptimizer.cpp:761-806) is fed with -- plus the vertex -> cell map.
//
// Vertex numbering is igraph's: graph.data.frame numbers vertices by first appearance in
// c(relations$from, relations$to).  The kept rows are in (i, j) order, so
//   * cells with at least one edge u>0 of their own ("active") come first, in cell order;
//   * then cells that only ever appear as a target, ordered by the first kept edge that names them;
//   * cells that appear in no kept edge are not vertices at all.
// For a kNN graph where every cell has a neighbour with a shared neighbour this is the identity.
//
// Input: the count byte of every edge slot with bit 7 = "mutual" (i is in N(t) too), produced by
// the count kernels in OUT==2 mode.  For rows without repeated ids u(i,t) == u(t,i), so an
// undirected pair {i,t} carries w (one direction) or w + w (both), and the direction that emits it
// is: the one from the smaller vertex id if both exist, else the only one.
// Supported domain (flags otherwise): k <= 127, no repeated ids.
// PTX-free: with GFICF_CUDA_EMU the header compiles as plain C++ (tests/test_snn_emu.py runs the kernels that way).
#pragma once
#include "jaccard_weight.cuh"
#include "scan_kernels.cuh"

namespace gficf {

constexpr unsigned kFlagIsolated = 16u;  // informational: some cell has no kept edge of its own

// one staged entry of pass 2: a single 16-byte store per entry (the scattered ones land at random
// positions, three separate 4/4/8-byte stores tripled the partial-sector writes)
struct __align__(16) SnnEntry {
  int col, row;
  double w;
};

// act[i] = 1 when row i keeps at least one edge (u > 0).  One warp per row.
__global__ void __launch_bounds__(256)
snn_active_kernel(const uint8_t* __restrict__ um, long long n, int k, int* __restrict__ act,
                  unsigned* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  bool isolated = false;
  for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nwarps) {
    bool nz = false;
    for (int j = lane; j < k; j += 32) nz |= (um[i * (long long)k + j] & 0x7Fu) != 0;
    const bool any = __any_sync(kFull, nz);
    if (lane == 0) act[i] = any ? 1 : 0;
    isolated |= !any;
  }
  if (isolated && lane == 0) atomicOr(flags, kFlagIsolated);
}

// Target-only vertices.  pass 0: first[t] = smallest kept edge number that names inactive cell t;
// pass 1: cntB[i] = edges of row i that are such a first appearance; pass 2: number them.
// All three leave at once when every cell is active (*n_active == n).
template <int PASS>
__global__ void __launch_bounds__(256)
snn_targets_kernel(const int* __restrict__ idx, const uint8_t* __restrict__ um, long long n, int k, int kp,
                   const int* __restrict__ act, const long long* __restrict__ off_a,
                   unsigned* __restrict__ first, int* __restrict__ cnt_b,
                   const long long* __restrict__ off_b, int* __restrict__ vid) {
  if (off_a[n] == n) return;
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nwarps) {
    int row_cnt = 0;
    for (int j0 = 0; j0 < k; j0 += 32) {
      const int j = j0 + lane;
      bool hit = false;
      int t = 0;
      const long long e = i * (long long)k + j;
      if (j < k && (um[e] & 0x7Fu) != 0) {
        t = __ldg(idx + i * (long long)kp + j);
        if (!act[t]) {
          if (PASS == 0) atomicMin(first + t, (unsigned)e);
          else hit = first[t] == (unsigned)e;
        }
      }
      if (PASS >= 1) {
        const unsigned m = __ballot_sync(kFull, hit);
        if (PASS == 2 && hit)
          vid[t] = (int)(off_a[n] + off_b[i] + row_cnt + __popc(m & ((1u << lane) - 1u)));
        row_cnt += __popc(m);
      }
    }
    if (PASS == 1 && lane == 0) cnt_b[i] = row_cnt;
  }
}

// vid[i] for active cells, vertex -> cell map (1-based cell ids, as R sees them), vertex count
__global__ void __launch_bounds__(256)
snn_vertex_ids_kernel(long long n, const int* __restrict__ act, const long long* __restrict__ off_a,
                      const long long* __restrict__ off_b, int* __restrict__ vid, int pass,
                      int* __restrict__ vertex_cell, long long* __restrict__ n_vertices) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    if (pass == 0) {
      vid[i] = act[i] ? (int)off_a[i] : -1;  // target-only cells are numbered by snn_targets_kernel<2>
    } else {
      const int v = vid[i];
      if (v >= 0 && vertex_cell) vertex_cell[v] = (int)(i + 1);
    }
  }
  if (pass == 1 && blockIdx.x == 0 && threadIdx.x == 0)
    *n_vertices = off_a[n] + (off_a[n] == n ? 0 : off_b[n]);
}

// pass 1 (count) / pass 2 (scatter) over the kept edges: one warp per row
template <bool SCATTER>
__global__ void __launch_bounds__(256)
snn_edges_kernel(const int* __restrict__ idx, const uint8_t* __restrict__ um, long long n, int k, int kp,
                 const int* __restrict__ vid, int* __restrict__ cnt_or_cursor,
                 const long long* __restrict__ colptr, SnnEntry* __restrict__ tmp) {
  __shared__ double lut[128];
  if ((int)threadIdx.x <= k && threadIdx.x < 128) lut[threadIdx.x] = jaccard_weight((int)threadIdx.x, k);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nwarps) {
    const int vi = vid[i];
    for (int j0 = 0; j0 < k; j0 += 32) {
      const int j = j0 + lane;
      int u = 0, vt = 0;
      bool mut = false;
      if (j < k) {
        const unsigned b = um[i * (long long)k + j];
        u = (int)(b & 0x7Fu);
        mut = (b & 0x80u) != 0;
        if (u > 0) vt = vid[__ldg(idx + i * (long long)kp + j)];
      }
      const bool off_diag = u > 0 && vt != vi;
      const bool own = off_diag && vt > vi;            // emitted in this row's own column vi
      const bool other = off_diag && vt < vi && !mut;  // one-directional edge into column vt
      const unsigned own_mask = __ballot_sync(kFull, own);
      if (!SCATTER) {
        if (lane == 0 && own_mask) atomicAdd(cnt_or_cursor + vi, __popc(own_mask));
        if (other) atomicAdd(cnt_or_cursor + vt, 1);
      } else {
        int base = 0;
        if (lane == 0 && own_mask) base = atomicAdd(cnt_or_cursor + vi, __popc(own_mask));
        base = __shfl_sync(kFull, base, 0);
        long long pos = -1;
        int col = 0, row = 0;
        if (own) {
          pos = colptr[vi] + base + __popc(own_mask & ((1u << lane) - 1u));
          col = vi;
          row = vt;
        } else if (other) {
          pos = colptr[vt] + atomicAdd(cnt_or_cursor + vt, 1);
          col = vt;
          row = vi;
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
  }
}

// ---------------------------------------------------------------------------
// Rows ascending inside every column (the staged entries of a column are in atomic arrival order).
//   deg <= 32    one warp, bitonic network over the lanes (shuffles)
//   deg <= 512   one warp, rank sort (each lane counts the smaller rows of its entries)
//   deg >  512   queued for snn_sort_big_kernel: one CTA per column, bitonic in shared memory up
//                to 4096 entries, CTA-wide rank sort beyond (hub vertices)
// Rows are distinct inside a column, so ranks are a permutation.
// ---------------------------------------------------------------------------
constexpr int kSnnWarpRankMax = 512;
constexpr int kSnnSmemSortMax = 4096;

__global__ void __launch_bounds__(256)
snn_sort_columns_kernel(const long long* __restrict__ colptr, const long long* __restrict__ n_vertices,
                        const SnnEntry* __restrict__ tmp, int* __restrict__ row_out,
                        double* __restrict__ w_out, int* __restrict__ big_cols, int* __restrict__ n_big) {
  const int lane = threadIdx.x & 31;
  const long long nv = *n_vertices;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nv; c += nwarps) {
    const long long lo = colptr[c];
    const int deg = (int)(colptr[c + 1] - lo);
    if (deg <= 0) continue;
    if (deg <= 32) {
      int r = 0x7fffffff;
      double w = 0.0;
      if (lane < deg) {
        const SnnEntry en = tmp[lo + lane];
        r = en.row;
        w = en.w;
      }
#pragma unroll
      for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          const int r2 = __shfl_xor_sync(kFull, r, stride);
          const double w2 = __shfl_xor_sync(kFull, w, stride);
          const bool up = (lane & size) == 0;          // ascending half of the bitonic merge
          const bool low = (lane & stride) == 0;       // this lane keeps the smaller of the pair
          const bool take = (r2 < r) == (up == low);   // rows are distinct: no ties among real entries
          if (r2 != r && take) {
            r = r2;
            w = w2;
          }
        }
      }
      if (lane < deg) {
        row_out[lo + lane] = r;
        w_out[lo + lane] = w;
      }
    } else if (deg <= kSnnWarpRankMax) {
      for (int e = lane; e < deg; e += 32) {
        const SnnEntry en = tmp[lo + e];
        int rank = 0;
        for (int q = 0; q < deg; ++q) rank += tmp[lo + q].row < en.row;
        row_out[lo + rank] = en.row;
        w_out[lo + rank] = en.w;
      }
    } else if (lane == 0) {
      big_cols[atomicAdd(n_big, 1)] = (int)c;
    }
  }
}

__global__ void __launch_bounds__(256)
snn_sort_big_kernel(const long long* __restrict__ colptr, const SnnEntry* __restrict__ tmp,
                    int* __restrict__ row_out, double* __restrict__ w_out, const int* __restrict__ big_cols,
                    const int* __restrict__ n_big) {
  GFICF_DYNAMIC_SMEM(snn_smem);
  int* s_row = reinterpret_cast<int*>(snn_smem);                            // [kSnnSmemSortMax]
  double* s_w = reinterpret_cast<double*>(snn_smem + kSnnSmemSortMax * 4);  // [kSnnSmemSortMax]
  const int nb = *n_big;
  for (int b = blockIdx.x; b < nb; b += gridDim.x) {
    const int c = big_cols[b];
    const long long lo = colptr[c];
    const int deg = (int)(colptr[c + 1] - lo);
    if (deg <= kSnnSmemSortMax) {
      int p2 = 1;
      while (p2 < deg) p2 <<= 1;
      for (int x = threadIdx.x; x < p2; x += blockDim.x) {
        if (x < deg) {
          const SnnEntry en = tmp[lo + x];
          s_row[x] = en.row;
          s_w[x] = en.w;
        } else {
          s_row[x] = 0x7fffffff;
          s_w[x] = 0.0;
        }
      }
      __syncthreads();
      for (int size = 2; size <= p2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int x = threadIdx.x; x < p2 / 2; x += blockDim.x) {
            const int a = 2 * x - (x & (stride - 1));  // lower index of the pair
            const int bb = a + stride;
            const bool up = (a & size) == 0;
            const int ra = s_row[a], rb = s_row[bb];
            if ((ra > rb) == up) {
              s_row[a] = rb;
              s_row[bb] = ra;
              const double wa = s_w[a];
              s_w[a] = s_w[bb];
              s_w[bb] = wa;
            }
          }
          __syncthreads();
        }
      }
      for (int x = threadIdx.x; x < deg; x += blockDim.x) {
        row_out[lo + x] = s_row[x];
        w_out[lo + x] = s_w[x];
      }
      __syncthreads();
    } else {
      for (int e = threadIdx.x; e < deg; e += blockDim.x) {
        const SnnEntry en = tmp[lo + e];
        int rank = 0;
        for (int q = 0; q < deg; ++q) rank += tmp[lo + q].row < en.row;
        row_out[lo + rank] = en.row;
        w_out[lo + rank] = en.w;
      }
    }
  }
}

}  // namespace gficf
