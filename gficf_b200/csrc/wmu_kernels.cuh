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
#pragma once
#include <cuda_runtime.h>
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
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kWmuThreads)
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
  __shared__ unsigned s_carry, s_flag;
  __shared__ WmuGeneOut s_out;

  const long long N = n1 + n2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long* const k0 = scratch_keys + (size_t)blockIdx.x * 2 * N;
  unsigned long long* const k1 = k0 + N;
  unsigned* const p0 = scratch_pay + (size_t)blockIdx.x * (3 * N + 2);
  unsigned* const p1 = p0 + N;
  unsigned* const px = p0 + 2 * N;  // [N+1]

  for (long long g = blockIdx.x; g < n_genes; g += gridDim.x) {
    // ---- keys of the gene's row: element (g, c) of a column-major matrix sits at g + c * n_genes
    for (long long i = tid; i < N; i += kWmuThreads) {
      const double v = i < n1 ? __ldg(mat_x + g + i * n_genes) : __ldg(mat_y + g + (i - n1) * n_genes);
      k0[i] = wmu_key(v);
      p0[i] = (unsigned)i;
    }
    int cur = 0;
    __syncthreads();
    // ---- LSD radix sort, 8 bits per pass
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = pass * 8;
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      const unsigned long long* kin = cur ? k1 : k0;
      const unsigned* pin = cur ? p1 : p0;
      unsigned long long* kout = cur ? k0 : k1;
      unsigned* pout = cur ? p0 : p1;
      for (long long i = tid; i < N; i += kWmuThreads)
        atomicAdd(&hist[(unsigned)(kin[i] >> shift) & 255u], 1u);
      __syncthreads();
      if (tid == 0) s_flag = 0;
      __syncthreads();
      if (tid < 256 && hist[tid] == (unsigned)N) s_flag = 1;  // every key has the same digit: nothing moves
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
      for (long long t0 = 0; t0 < N; t0 += kWmuThreads) {
        for (int x = tid; x < kWmuWarps * 256; x += kWmuThreads) (&wcnt[0][0])[x] = 0;
        __syncthreads();
        const long long i = t0 + tid;
        const bool valid = i < N;
        unsigned long long key = 0;
        unsigned pay = 0, d = 256u + (unsigned)lane;  // invalid lanes: a digit nobody shares
        if (valid) {
          key = kin[i];
          pay = pin[i];
          d = (unsigned)(key >> shift) & 255u;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned rank_in_warp = __popc(peers & ((1u << lane) - 1u));
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
    // ---- px[p] = number of group-1 elements among sorted positions [0, p)
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (long long t0 = 0; t0 < N; t0 += kWmuThreads) {
      const long long i = t0 + tid;
      const unsigned f = (i < N && (long long)pays[i] < n1) ? 1u : 0u;
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
      if (i < N) px[i] = before + inc - f;
      __syncthreads();
      if (tid == kWmuThreads - 1) s_carry = before + inc;
      __syncthreads();
    }
    if (tid == 0) px[N] = s_carry;
    __syncthreads();
    // ---- tie groups: a thread that sits on the first element of a group finds the group's end
    //      by binary search (the keys are sorted) and adds the group's contributions
    unsigned long long two_r1 = 0, tie_int = 0;
    unsigned distinct = 0;
    const bool ordered_ties = N >= kWmuExactTieN;
    double* tie_list = reinterpret_cast<double*>(cur ? k0 : k1);  // the other key buffer is free now
    if (ordered_ties) {
      if (tid == 0) s_carry = 0;
      __syncthreads();
    }
    for (long long t0 = 0; t0 < N; t0 += kWmuThreads) {
      const long long i = t0 + tid;
      bool start = false;
      long long j = 0;
      if (i < N) {
        const unsigned long long key = keys[i];
        start = i == 0 || keys[i - 1] != key;
        if (start) {
          long long lo = i + 1, hi = N;  // first position > i whose key differs
          while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (keys[mid] == key) lo = mid + 1;
            else hi = mid;
          }
          j = lo;
          const unsigned long long cx = px[j] - px[i];
          two_r1 += cx * (unsigned long long)(i + j + 1);  // 2 * averaged rank = i + j + 1 (0-based i, j)
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
      s_out.tie_sum = (double)s;  // exact: < 2^53 in the integer regime
    }
    __syncthreads();
    if (lane == 0) s_red[warp] = distinct;
    __syncthreads();
    if (tid == 0) {
      unsigned long long s = 0;
      for (int w = 0; w < kWmuWarps; ++w) s += s_red[w];
      s_out.distinct = (unsigned)s;
      if (ordered_ties) {  // the reference's sequential double accumulation, in sorted order
        double acc = 0.0;
        const unsigned cnt = s_carry;
        for (unsigned q = 0; q < cnt; ++q) acc = __dadd_rn(acc, tie_list[q]);
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
// mann_whitney.cpp:123-126).  One thread per gene; out_ratio[g] = avg1 / avg2.
__global__ void __launch_bounds__(128)
wmu_means_kernel(const double* __restrict__ mat_x, const double* __restrict__ mat_y, long long n_genes,
                 long long n1, long long n2, double* __restrict__ out_ratio) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genes) return;
  double s1 = 0.0, s2 = 0.0;
#pragma unroll 8
  for (long long c = 0; c < n1; ++c) s1 = __dadd_rn(s1, __dadd_rn(__ldg(mat_x + g + c * n_genes), 1.0));
#pragma unroll 8
  for (long long c = 0; c < n2; ++c) s2 = __dadd_rn(s2, __dadd_rn(__ldg(mat_y + g + c * n_genes), 1.0));
  out_ratio[g] = __ddiv_rn(__ddiv_rn(s1, (double)n1), __ddiv_rn(s2, (double)n2));
}

}  // namespace gficf
