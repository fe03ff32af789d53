// wmu_kernels.cuh -- Mann-Whitney U per gene on the device ("next" row 4).
//
// What it replaces (reference = dibbelab/gficf):
//   src/rcpp_parallel_mann_whitney.cpp:27-100   WMU_test::operator(): per gene concatenate the two
//       groups, sort (sort_indexes, mann_whitney.cpp:17-29), average ranks over ties (getRanks
//       :31-54), tie-group sizes (getCounts :65-85), rank sums U1/U2, sigma with the tie
//       correction (getSigma :87-99), continuity-corrected z, and the group means for log2FC
//   called per cluster from R/deGenes.R:44-54 (findClusterMarkers)
//
// Everything here is integer or correctly-rounded IEEE arithmetic, so it is bit-identical to the
// reference's doubles:
//   * rank sums are sums of half-integers below 2^53 -- exact in any order; here 2*R1 is
//     accumulated in 64-bit integers as sum over tie groups of cx * (i + j + 1) (cx = members of
//     group 1 in the group that occupies sorted positions [i, j));
//   * the tie term sum(t^3 - t) is accumulated by the reference sequentially in sorted order in
//     double; below N = 208064 every partial sum is an exact integer < 2^53 and the order is
//     irrelevant (integer accumulation here); above, the non-trivial groups are compacted in order
//     and summed sequentially by one thread, with the reference's rounding ((t*t)*t) - t;
//   * sigma and z use __dmul_rn / __ddiv_rn / __dsqrt_rn in the reference's operation order;
//   * the group means are sequential sums over the cells in their original order (one thread per
//     gene, coalesced because the R matrix is column-major: consecutive genes are adjacent).
// The two transcendental steps -- the normal cdf of z (GSL in the reference) and log2 of the mean
// ratio -- are left to the host wrapper (libm), one evaluation per gene.
//
// Sort: one CTA per gene, least-significant-digit radix sort (8-bit digits, 8 passes, passes whose
// digit is constant are skipped) on order-preserving 64-bit keys with the element's original
// position as payload, ping-pong buffers in global memory (L2 resident for typical N).
// With GFICF_CUDA_EMU defined the header compiles as plain C++ against tests/cuda_emu/cuda_emu.h (the CPU test
// suite runs both kernels that way); the product build never defines it.
#pragma once
#ifdef GFICF_CUDA_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace gficf {

constexpr int kWmuThreads = 512;
constexpr int kWmuWarps = kWmuThreads / 32;
constexpr long long kWmuExactTieN = 208064;  // N^3 < 2^53 below this

__device__ __forceinline__ unsigned long long wmu_key(double v) {
  v = v + 0.0;  // -0.0 -> +0.0: they compare equal in the reference (one tie group)
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

struct WmuGeneOut {
  unsigned long long two_r1;  // 2 * (sum of the averaged ranks of group 1)
  double tie_sum;             // sum over tie groups of t^3 - t, the reference's double
  unsigned distinct;          // number of distinct values (nties.size() in the reference)
  unsigned pad;
};

// z (continuity corrected, divided by sigma) from the exact pieces; *single = 1 when all values are
// equal (the reference then leaves pval = 1, rcpp_parallel_mann_whitney.cpp:56)
__device__ __forceinline__ double wmu_z(const WmuGeneOut& o, long long n1, long long n2, int* single) {
  *single = o.distinct <= 1;
  if (*single) return 0.0;
  const long long n = n1 + n2;
  // U1 = R1 - n1(n1+1)/2, U2 = R2 - n2(n2+1)/2 with R1 + R2 = n(n+1)/2: all multiples of 1/2, exact
  const long long two_u1 = (long long)o.two_r1 - n1 * (n1 + 1);
  const long long two_u2 = (n * (n + 1) - (long long)o.two_r1) - n2 * (n2 + 1);
  const double u1 = 0.5 * (double)two_u1, u2 = 0.5 * (double)two_u2;
  const double mu = (double)((unsigned long long)(n1 * n2) / 2ull);  // :86 size_t division, then double
  // getSigma (mann_whitney.cpp:87-99) with n1, n2 as doubles, in its operation order
  const double d1 = (double)n1, d2 = (double)n2;
  const double a = __ddiv_rn(__dmul_rn(d1, d2), 12.0);
  const double np1 = __dadd_rn(__dadd_rn(d1, d2), 1.0);
  const double den = __dmul_rn(__dadd_rn(d1, d2), __dadd_rn(__dadd_rn(d1, d2), -1.0));
  const double sig = __dsqrt_rn(__dmul_rn(a, __dadd_rn(np1, -__ddiv_rn(o.tie_sum, den))));
  double z = u1 < u2 ? u1 - mu : u2 - mu;  // exact
  z = z < 0 ? z + 0.5 : z - 0.5;           // :90 continuity correction, exact
  return __ddiv_rn(z, sig);
}

// ---------------------------------------------------------------------------
// One CTA per gene (persistent over genes).  scratch per CTA: keys[2][N] | pay[2][N] | px[N+1].
//
// Expression matrices are sparse: most of a gene's values are exactly 0, i.e. ONE tie group that
// needs no sorting.  Only the M non-zero values are compacted and sorted; the zeros enter the rank
// sums, the tie term and the distinct-value count analytically as the group that occupies sorted
// positions [n_neg, n_neg + nz) (n_neg = number of negative values, nz = N - M).  Dense input
// simply has M = N.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kWmuThreads, 2)
wmu_rank_kernel(const double* __restrict__ mat_x, const double* __restrict__ mat_y, long long n_genes,
                long long n1, long long n2, unsigned long long* __restrict__ scratch_keys,
                unsigned* __restrict__ scratch_pay, double* __restrict__ out_z,
                int* __restrict__ out_single) {
  __shared__ unsigned hist[256];
  __shared__ unsigned base[256];
  __shared__ unsigned short wcnt[kWmuWarps][256];
  __shared__ unsigned tile_total[256];
  __shared__ unsigned long long s_red[kWmuWarps];
  __shared__ unsigned s_scan[kWmuWarps];
  __shared__ unsigned s_carry, s_flag, s_m, s_z1, s_neg, s_negties;
  __shared__ WmuGeneOut s_out;

  const long long N = n1 + n2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned long long* const k0 = scratch_keys + (size_t)blockIdx.x * 2 * N;
  unsigned long long* const k1 = k0 + N;
  unsigned* const p0 = scratch_pay + (size_t)blockIdx.x * (3 * N + 2);
  unsigned* const p1 = p0 + N;
  unsigned* const px = p0 + 2 * N;  // [M+1]
  const unsigned long long kZeroKey = 0x8000000000000000ull;  // wmu_key(0.0)

  for (long long g = blockIdx.x; g < n_genes; g += gridDim.x) {
    if (tid == 0) {
      s_m = 0;
      s_z1 = 0;
      s_neg = 0;
      s_negties = 0;
    }
    __syncthreads();
    // ---- the gene's row: element (g, c) of a column-major matrix sits at g + c * n_genes.
    //      Non-zero values are appended (any order) as (key, original position); zeros are counted.
    for (long long t0 = 0; t0 < N; t0 += kWmuThreads) {
      const long long i = t0 + tid;
      double v = 0.0;
      const bool in = i < N;
      if (in) v = i < n1 ? __ldg(mat_x + g + i * n_genes) : __ldg(mat_y + g + (i - n1) * n_genes);
      const bool nzv = in && v != 0.0;  // -0.0 == 0.0: same tie group as +0.0, like in the reference
      const unsigned m_nz = __ballot_sync(0xffffffffu, nzv);
      const unsigned m_z1 = __ballot_sync(0xffffffffu, in && !nzv && i < n1);
      const unsigned m_neg = __ballot_sync(0xffffffffu, nzv && v < 0.0);
      unsigned wbase = 0;
      if (lane == 0) {
        if (m_nz) wbase = atomicAdd(&s_m, __popc(m_nz));
        if (m_z1) atomicAdd(&s_z1, __popc(m_z1));
        if (m_neg) atomicAdd(&s_neg, __popc(m_neg));
      }
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (nzv) {
        const unsigned pos = wbase + __popc(m_nz & lt_mask);
        k0[pos] = wmu_key(v);
        p0[pos] = (unsigned)i;
      }
    }
    __syncthreads();
    const long long M = s_m;
    const long long nz = N - M, z1 = s_z1, n_neg = s_neg;
    int cur = 0;
    // ---- LSD radix sort of the M non-zero keys, 8 bits per pass
    for (int pass = 0; pass < 8 && M > 1; ++pass) {
      const int shift = pass * 8;
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      const unsigned long long* kin = cur ? k1 : k0;
      const unsigned* pin = cur ? p1 : p0;
      unsigned long long* kout = cur ? k0 : k1;
      unsigned* pout = cur ? p0 : p1;
      for (long long t0 = 0; t0 < M; t0 += kWmuThreads) {  // warp-aggregated histogram
        const long long i = t0 + tid;
        const unsigned d = i < M ? ((unsigned)(kin[i] >> shift) & 255u) : 256u + (unsigned)lane;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (i < M && (peers & lt_mask) == 0) atomicAdd(&hist[d], (unsigned)__popc(peers));
      }
      __syncthreads();
      if (tid == 0) s_flag = 0;
      __syncthreads();
      if (tid < 256 && hist[tid] == (unsigned)M) s_flag = 1;  // every key has the same digit: nothing moves
      __syncthreads();
      if (s_flag) continue;
      if (warp == 0) {  // exclusive scan of the 256 counts: 8 per lane
        unsigned c[8], sum = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          c[q] = hist[lane * 8 + q];
          sum += c[q];
        }
        unsigned inc = sum;
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) {
          const unsigned o = __shfl_up_sync(0xffffffffu, inc, m);
          if (lane >= m) inc += o;
        }
        unsigned run = inc - sum;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          base[lane * 8 + q] = run;
          run += c[q];
        }
      }
      __syncthreads();
      for (long long t0 = 0; t0 < M; t0 += kWmuThreads) {
        for (int x = tid; x < kWmuWarps * 256; x += kWmuThreads) (&wcnt[0][0])[x] = 0;
        __syncthreads();
        const long long i = t0 + tid;
        const bool valid = i < M;
        unsigned long long key = 0;
        unsigned pay = 0, d = 256u + (unsigned)lane;  // invalid lanes: a digit nobody shares
        if (valid) {
          key = kin[i];
          pay = pin[i];
          d = (unsigned)(key >> shift) & 255u;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned rank_in_warp = __popc(peers & lt_mask);
        if (valid && rank_in_warp == 0) wcnt[warp][d] = (unsigned short)__popc(peers);
        __syncthreads();
        if (tid < 256) {  // exclusive prefix over the warps, per digit
          unsigned acc = 0;
#pragma unroll
          for (int w = 0; w < kWmuWarps; ++w) {
            const unsigned c = wcnt[w][tid];
            wcnt[w][tid] = (unsigned short)acc;
            acc += c;
          }
          tile_total[tid] = acc;
        }
        __syncthreads();
        if (valid) {
          const unsigned pos = base[d] + wcnt[warp][d] + rank_in_warp;
          kout[pos] = key;
          pout[pos] = pay;
        }
        __syncthreads();
        if (tid < 256) base[tid] += tile_total[tid];
      }
      cur ^= 1;
      __syncthreads();
    }
    const unsigned long long* keys = cur ? k1 : k0;
    const unsigned* pays = cur ? p1 : p0;
    // ---- px[p] = number of group-1 elements among sorted non-zero positions [0, p)
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (long long t0 = 0; t0 < M; t0 += kWmuThreads) {
      const long long i = t0 + tid;
      const unsigned f = (i < M && (long long)pays[i] < n1) ? 1u : 0u;
      unsigned inc = f;
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, inc, m);
        if (lane >= m) inc += o;
      }
      if (lane == 31) s_scan[warp] = inc;
      __syncthreads();
      unsigned before = s_carry;
      for (int w = 0; w < warp; ++w) before += s_scan[w];
      if (i < M) px[i] = before + inc - f;
      __syncthreads();
      if (tid == kWmuThreads - 1) s_carry = before + inc;
      __syncthreads();
    }
    if (tid == 0) px[M] = s_carry;
    __syncthreads();
    // ---- tie groups of the non-zero values: a thread that sits on the first element of a group
    //      finds the group's end by binary search (the keys are sorted) and adds the group's
    //      contributions.  Sorted position q of the non-zero array is global position q (negative
    //      values) or q + nz (positive values): the zeros sit in between.
    unsigned long long two_r1 = 0, tie_int = 0;
    unsigned distinct = 0;
    const bool ordered_ties = N >= kWmuExactTieN;
    double* tie_list = reinterpret_cast<double*>(cur ? k0 : k1);  // the other key buffer is free now
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (long long t0 = 0; t0 < M; t0 += kWmuThreads) {
      const long long i = t0 + tid;
      bool start = false;
      long long j = 0;
      if (i < M) {
        const unsigned long long key = keys[i];
        start = i == 0 || keys[i - 1] != key;
        if (start) {
          long long lo = i + 1, hi = M;  // first position > i whose key differs
          while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (keys[mid] == key) lo = mid + 1;
            else hi = mid;
          }
          j = lo;
          const long long off = key > kZeroKey ? nz : 0;
          const unsigned long long cx = px[j] - px[i];
          two_r1 += cx * (unsigned long long)(i + j + 2 * off + 1);  // 2 * averaged rank = gi + gj + 1 (0-based)
          ++distinct;
          const unsigned long long t = (unsigned long long)(j - i);
          if (!ordered_ties) tie_int += t * t * t - t;
        }
      }
      if (ordered_ties) {  // non-trivial groups, compacted in sorted order
        const unsigned f = (start && j - i >= 2) ? 1u : 0u;
        unsigned inc = f;
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) {
          const unsigned o = __shfl_up_sync(0xffffffffu, inc, m);
          if (lane >= m) inc += o;
        }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        unsigned before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_scan[w];
        if (f) {
          const double t = (double)(j - i);
          tie_list[before + inc - 1] = __dadd_rn(__dmul_rn(__dmul_rn(t, t), t), -t);  // ((t*t)*t) - t, :94
          if (i < n_neg) atomicAdd(&s_negties, 1u);  // groups that precede the zeros
        }
        __syncthreads();
        if (tid == kWmuThreads - 1) s_carry = before + inc;
        __syncthreads();
      }
    }
    // ---- block reduction of the integer pieces
#pragma unroll
    for (int m = 16; m; m >>= 1) {
      two_r1 += __shfl_xor_sync(0xffffffffu, two_r1, m);
      tie_int += __shfl_xor_sync(0xffffffffu, tie_int, m);
      distinct += __shfl_xor_sync(0xffffffffu, distinct, m);
    }
    if (lane == 0) s_red[warp] = two_r1;
    __syncthreads();
    if (tid == 0) {
      unsigned long long s = 0;
      for (int w = 0; w < kWmuWarps; ++w) s += s_red[w];
      s_out.two_r1 = s;
    }
    __syncthreads();
    if (lane == 0) s_red[warp] = tie_int;
    __syncthreads();
    if (tid == 0) {
      unsigned long long s = 0;
      for (int w = 0; w < kWmuWarps; ++w) s += s_red[w];
      s_out.tie_sum = (double)s;  // exact: < 2^53 in the integer regime (completed below)
      s_red[0] = s;
    }
    __syncthreads();
    const unsigned long long tie_total_int = s_red[0];
    __syncthreads();
    if (lane == 0) s_red[warp] = distinct;
    __syncthreads();
    if (tid == 0) {
      unsigned long long s = 0;
      for (int w = 0; w < kWmuWarps; ++w) s += s_red[w];
      // the zeros: one group at sorted positions [n_neg, n_neg + nz)
      unsigned long long tr = s_out.two_r1, ti = tie_total_int;
      if (nz > 0) {
        tr += (unsigned long long)z1 * (unsigned long long)(2 * n_neg + nz + 1);
        s += 1;
        ti += (unsigned long long)nz * (unsigned long long)nz * (unsigned long long)nz - (unsigned long long)nz;
      }
      s_out.two_r1 = tr;
      s_out.distinct = (unsigned)s;
      s_out.tie_sum = (double)ti;
      if (ordered_ties) {  // the reference's sequential double accumulation, in sorted order
        double acc = 0.0;
        const unsigned cnt = s_carry, negc = s_negties;
        const double tz = (double)nz;
        for (unsigned q = 0; q < negc; ++q) acc = __dadd_rn(acc, tie_list[q]);
        if (nz >= 2) acc = __dadd_rn(acc, __dadd_rn(__dmul_rn(__dmul_rn(tz, tz), tz), -tz));
        for (unsigned q = negc; q < cnt; ++q) acc = __dadd_rn(acc, tie_list[q]);
        s_out.tie_sum = acc;
      }
      int single;
      const double z = wmu_z(s_out, n1, n2, &single);
      out_z[g] = z;
      out_single[g] = single;
    }
    __syncthreads();
  }
}

// Group means for the fold change: avg(v + 1) with the reference's sequential accumulation
// (std::accumulate from 0.0 over the cells in order, rcpp_parallel_mann_whitney.cpp:97-99,
// mann_whitney.cpp:123-126).  The sum of a gene is inherently sequential, the loading is not: a CTA
// owns 32 adjacent genes; its 8 warps stream tiles of 64 cells x 32 genes (coalesced: the R matrix is
// column-major, 32 adjacent genes are 256 contiguous bytes) through a double-buffered shared-memory
// tile while warp 0 -- one lane per gene -- adds its gene's values in cell order.
// out_ratio[g] = avg1 / avg2.
constexpr int kMeanGenes = 32, kMeanCells = 64, kMeanThreads = 256;

__global__ void __launch_bounds__(kMeanThreads)
wmu_means_kernel(const double* __restrict__ mat_x, const double* __restrict__ mat_y, long long n_genes,
                 long long n1, long long n2, double* __restrict__ out_ratio) {
  __shared__ double tile[2][kMeanCells][kMeanGenes + 1];
  const int tid = threadIdx.x, lane = tid & 31;
  const long long g0 = (long long)blockIdx.x * kMeanGenes;
  const long long N = n1 + n2;
  const long long ntiles = (N + kMeanCells - 1) / kMeanCells;
  auto load_tile = [&](long long t, int buf) {
    for (int x = tid; x < kMeanCells * kMeanGenes; x += kMeanThreads) {
      const int c = x / kMeanGenes, gg = x % kMeanGenes;
      const long long cell = t * kMeanCells + c, gene = g0 + gg;
      double v = 0.0;
      if (cell < N && gene < n_genes)
        v = cell < n1 ? __ldg(mat_x + gene + cell * n_genes) : __ldg(mat_y + gene + (cell - n1) * n_genes);
      tile[buf][c][gg] = v;
    }
  };
  double s1 = 0.0, s2 = 0.0;
  load_tile(0, 0);
  __syncthreads();
  for (long long t = 0; t < ntiles; ++t) {
    const int buf = (int)(t & 1);
    if (tid >= 32) {
      if (t + 1 < ntiles) {  // warps 1..7 fetch the next tile while warp 0 sums this one
        for (int x = tid - 32; x < kMeanCells * kMeanGenes; x += kMeanThreads - 32) {
          const int c = x / kMeanGenes, gg = x % kMeanGenes;
          const long long cell = (t + 1) * kMeanCells + c, gene = g0 + gg;
          double v = 0.0;
          if (cell < N && gene < n_genes)
            v = cell < n1 ? __ldg(mat_x + gene + cell * n_genes) : __ldg(mat_y + gene + (cell - n1) * n_genes);
          tile[buf ^ 1][c][gg] = v;
        }
      }
    } else {
      const long long c_lo = t * kMeanCells;
#pragma unroll 8
      for (int c = 0; c < kMeanCells; ++c) {
        const long long cell = c_lo + c;
        if (cell >= N) break;
        const double v1 = __dadd_rn(tile[buf][c][lane], 1.0);
        if (cell < n1) s1 = __dadd_rn(s1, v1);
        else s2 = __dadd_rn(s2, v1);
      }
    }
    __syncthreads();
  }
  if (tid < 32 && g0 + lane < n_genes)
    out_ratio[g0 + lane] = __ddiv_rn(__ddiv_rn(s1, (double)n1), __ddiv_rn(s2, (double)n2));
}

}  // namespace gficf
