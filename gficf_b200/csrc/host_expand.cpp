// host_expand.cpp -- see host_expand.h.  Plain host C++ (compiled by the host compiler, no CUDA).
#include "host_expand.h"

#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace gficf_host {

void fill_weight_table(int k, double* lut) {
  // src/rcpp_parallel_jaccard_coeff.cpp:51   rmat(i*ncol+j,2) = u/(2.0*mat.ncol() - u)
  // with `int u` and `size_t ncol`: 2.0*ncol is a double, u converts to double, one IEEE division
  const size_t ncol = (size_t)k;
  for (int u = 0; u <= k; ++u) lut[u] = u / (2.0 * ncol - u);
}

// ------------------------------------------------------------------ streaming copy
void stream_copy(void* dst_v, const void* src_v, size_t bytes) {
  char* dst = (char*)dst_v;
  const char* src = (const char*)src_v;
#if defined(__x86_64__) && defined(__SSE2__)
  if (bytes >= 256) {
    // head: up to the next 16-byte boundary of the destination
    const size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    if (head) {
      memcpy(dst, src, head);
      dst += head;
      src += head;
      bytes -= head;
    }
    const size_t blocks = bytes / 64;
    const __m128i* s = (const __m128i*)src;
    __m128i* d = (__m128i*)dst;
    for (size_t b = 0; b < blocks; ++b) {
      const __m128i x0 = _mm_loadu_si128(s + 0), x1 = _mm_loadu_si128(s + 1);
      const __m128i x2 = _mm_loadu_si128(s + 2), x3 = _mm_loadu_si128(s + 3);
      _mm_stream_si128(d + 0, x0);
      _mm_stream_si128(d + 1, x1);
      _mm_stream_si128(d + 2, x2);
      _mm_stream_si128(d + 3, x3);
      s += 4;
      d += 4;
    }
    _mm_sfence();
    dst += blocks * 64;
    src += blocks * 64;
    bytes -= blocks * 64;
  }
#endif
  if (bytes) memcpy(dst, src, bytes);
}

// ------------------------------------------------------------------ one tile of one column
// A tile of an output column is built in a small L1-resident array (plain stores; for the `to`
// column the transposed read of the caller's column-major matrix lands here) and then leaves with
// stream_copy: one contiguous run of the column, written once, never read back -> no
// read-for-ownership traffic on the destination.
constexpr int kTileEdges = 2048;

// generic bodies (the compiler vectorises what it can for the CPU it runs on)
template <typename T>
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx2", "default")))
#endif
static void build_tile(int col, const T* __restrict__ mat, long long n, int k, const uint8_t* __restrict__ u,
                       long long r0, int rows, const double* __restrict__ lut, double* __restrict__ t) {
  if (col == 2) {  // weight, :51 through the table
    const int ne = rows * k;
    for (int x = 0; x < ne; ++x) t[x] = lut[u[x]];
  } else if (col == 0) {  // from, :49 rmat(.,0) = i+1
    for (int i = 0; i < rows; ++i) {
      const double fi = (double)(r0 + i + 1);
      const uint8_t* ui = u + (size_t)i * k;
      double* f = t + (size_t)i * k;
      for (int j = 0; j < k; ++j) f[j] = ui[j] ? fi : 0.0;
    }
  } else {  // to, :50 k+1 == mat(i,j); column j of the caller's matrix is contiguous in i
    for (int j = 0; j < k; ++j) {
      const T* c = mat + (size_t)j * n + r0;
      for (int i = 0; i < rows; ++i) t[(size_t)i * k + j] = u[(size_t)i * k + j] ? (double)c[i] : 0.0;
    }
  }
}

#if defined(__x86_64__) && defined(__GNUC__)
#define GFICF_HOST_AVX512 1
// AVX-512 bodies, 8 edges per step: weight = gather from the table; from = masked broadcast of
// i+1; to = gather of the row's 8 ids out of 8 columns of the caller's matrix (those lines stay in
// L1 for the 8 rows that share them), loaded only where u>0.
template <typename T>
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) static void build_tile_avx512(
    int col, const T* __restrict__ mat, long long n, int k, const uint8_t* __restrict__ u, long long r0,
    int rows, const double* __restrict__ lut, double* __restrict__ t) {
  const __m256i zero = _mm256_setzero_si256();
  if (col == 2) {
    const int ne = rows * k;
    for (int x = 0; x < ne; x += 8) {
      const int rem = ne - x < 8 ? ne - x : 8;
      const __mmask8 m = (__mmask8)((1u << rem) - 1u);
      const __m256i u32 = _mm256_cvtepu8_epi32(_mm_maskz_loadu_epi8((__mmask16)m, u + x));
      _mm512_mask_storeu_pd(t + x, m, _mm512_i32gather_pd(u32, lut, 8));
    }
    return;
  }
  const __m512i vidx = _mm512_mullo_epi64(_mm512_set_epi64(7, 6, 5, 4, 3, 2, 1, 0), _mm512_set1_epi64(n));
  for (int i = 0; i < rows; ++i) {
    const uint8_t* ui = u + (size_t)i * k;
    const __m512d vfi = _mm512_set1_pd((double)(r0 + i + 1));
    double* o = t + (size_t)i * k;
    for (int j0 = 0; j0 < k; j0 += 8) {
      const int rem = k - j0 < 8 ? k - j0 : 8;
      const __mmask8 m = (__mmask8)((1u << rem) - 1u);
      const __m256i u32 = _mm256_cvtepu8_epi32(_mm_maskz_loadu_epi8((__mmask16)m, ui + j0));
      const __mmask8 nz = _mm256_cmpneq_epi32_mask(u32, zero);
      if (col == 0) {
        _mm512_mask_storeu_pd(o + j0, m, _mm512_maskz_mov_pd(nz, vfi));
      } else {
        const T* base = mat + (size_t)j0 * n + r0 + i;
        __m512d tv;
        if (sizeof(T) == 8)
          tv = _mm512_mask_i64gather_pd(_mm512_setzero_pd(), nz, vidx, base, 8);
        else
          tv = _mm512_cvtepi32_pd(_mm512_mask_i64gather_epi32(zero, nz, vidx, base, 4));
        _mm512_mask_storeu_pd(o + j0, m, tv);
      }
    }
  }
}

// Direct form for k >= 8: no tile, every aligned group of 8 edges is built in a register and leaves
// with one 64-byte non-temporal store.  A group covers at most two rows (k >= 8): the lanes at and
// beyond `k - rem` belong to row r0+1.
template <typename T>
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) static void expand_column_avx512(
    const ExpandJob& job, int col, long long row_lo, long long row_hi) {
  const int k = job.k;
  const long long n = job.n;
  const uint8_t* u = job.counts - (size_t)job.counts_row0 * k;  // indexed by absolute edge number
  const T* mat = (const T*)job.idx;
  const double* lut = job.lut;
  double* dst = job.out + (size_t)col * job.E;
  long long e = row_lo * k;
  const long long e_end = row_hi * k;
  auto scalar = [&](long long x) {
    const long long i = x / k;
    const int j = (int)(x - i * k);
    const int uu = u[x];
    dst[x] = col == 2 ? lut[uu] : (uu == 0 ? 0.0 : (col == 0 ? (double)(i + 1) : (double)mat[(size_t)j * n + i]));
  };
  while (e < e_end && ((uintptr_t)(dst + e) & 63)) scalar(e++);
  long long r0 = e / k;
  int rem = (int)(e - r0 * k);
  const __m256i zero = _mm256_setzero_si256();
  const __m512i iota = _mm512_set_epi64(7, 6, 5, 4, 3, 2, 1, 0);
  const __m512i vn = _mm512_set1_epi64(n), vk = _mm512_set1_epi64(k), one = _mm512_set1_epi64(1);
  for (; e + 8 <= e_end; e += 8) {
    const __m256i u32 = _mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i*)(u + e)));
    __m512d v;
    if (col == 2) {
      v = _mm512_i32gather_pd(u32, lut, 8);
    } else {
      const __mmask8 nz = _mm256_cmpneq_epi32_mask(u32, zero);
      const int first = k - rem;  // lanes >= first are in the next row
      const __mmask8 over = first >= 8 ? (__mmask8)0 : (__mmask8)(0xFFu << first);
      if (col == 0) {
        v = _mm512_mask_mov_pd(_mm512_set1_pd((double)(r0 + 1)), over, _mm512_set1_pd((double)(r0 + 2)));
        v = _mm512_maskz_mov_pd(nz, v);
      } else {
        __m512i jl = _mm512_add_epi64(iota, _mm512_set1_epi64(rem));
        jl = _mm512_mask_sub_epi64(jl, over, jl, vk);
        const __m512i r0v = _mm512_set1_epi64(r0);
        const __m512i il = _mm512_mask_add_epi64(r0v, over, r0v, one);
        const __m512i off = _mm512_add_epi64(_mm512_mullo_epi64(jl, vn), il);
        if (sizeof(T) == 8)
          v = _mm512_mask_i64gather_pd(_mm512_setzero_pd(), nz, off, mat, 8);
        else
          v = _mm512_cvtepi32_pd(_mm512_mask_i64gather_epi32(zero, nz, off, mat, 4));
      }
    }
    _mm512_stream_pd(dst + e, v);
    rem += 8;
    if (rem >= k) {
      rem -= k;
      ++r0;
    }
  }
  _mm_sfence();
  while (e < e_end) scalar(e++);
}

// `to` column: an 8 x 8 register transpose turns 8 column segments of the caller's matrix (8
// consecutive rows each, one 64-byte load) into 8 row segments, so the column-major matrix is read
// with full-line loads instead of one gathered element per load.
__attribute__((target("avx512f"))) static inline void transpose8x8(__m512d (&r)[8]) {
  const __m512d t0 = _mm512_unpacklo_pd(r[0], r[1]), t1 = _mm512_unpackhi_pd(r[0], r[1]);
  const __m512d t2 = _mm512_unpacklo_pd(r[2], r[3]), t3 = _mm512_unpackhi_pd(r[2], r[3]);
  const __m512d t4 = _mm512_unpacklo_pd(r[4], r[5]), t5 = _mm512_unpackhi_pd(r[4], r[5]);
  const __m512d t6 = _mm512_unpacklo_pd(r[6], r[7]), t7 = _mm512_unpackhi_pd(r[6], r[7]);
  const __m512d v0 = _mm512_shuffle_f64x2(t0, t2, 0x44), v1 = _mm512_shuffle_f64x2(t0, t2, 0xEE);
  const __m512d v2 = _mm512_shuffle_f64x2(t1, t3, 0x44), v3 = _mm512_shuffle_f64x2(t1, t3, 0xEE);
  const __m512d v4 = _mm512_shuffle_f64x2(t4, t6, 0x44), v5 = _mm512_shuffle_f64x2(t4, t6, 0xEE);
  const __m512d v6 = _mm512_shuffle_f64x2(t5, t7, 0x44), v7 = _mm512_shuffle_f64x2(t5, t7, 0xEE);
  r[0] = _mm512_shuffle_f64x2(v0, v4, 0x88);
  r[1] = _mm512_shuffle_f64x2(v2, v6, 0x88);
  r[2] = _mm512_shuffle_f64x2(v0, v4, 0xDD);
  r[3] = _mm512_shuffle_f64x2(v2, v6, 0xDD);
  r[4] = _mm512_shuffle_f64x2(v1, v5, 0x88);
  r[5] = _mm512_shuffle_f64x2(v3, v7, 0x88);
  r[6] = _mm512_shuffle_f64x2(v1, v5, 0xDD);
  r[7] = _mm512_shuffle_f64x2(v3, v7, 0xDD);
}

constexpr int kToTileRows = 512;  // rows per tile: 4 KB of every column per tile (one page touch per column)

template <typename T>
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) static void expand_to_avx512(
    const ExpandJob& job, long long row_lo, long long row_hi) {
  const int k = job.k;
  const long long n = job.n;
  const uint8_t* u = job.counts - (size_t)job.counts_row0 * k;  // indexed by absolute edge number
  const T* mat = (const T*)job.idx;
  double* dst = job.out + (size_t)job.E;
  static thread_local double* tile = nullptr;  // kToTileRows x 255 doubles (1 MB, L2 resident), per thread
  if (!tile) tile = (double*)aligned_alloc(64, (size_t)kToTileRows * 256 * sizeof(double));
  const int tr = kToTileRows;
  const __m256i zero = _mm256_setzero_si256();
  for (long long t0 = row_lo; t0 < row_hi; t0 += tr) {
    const int rows = (int)((row_hi - t0) < tr ? (row_hi - t0) : tr);
    // 1. transposed read of rows [t0, t0+rows) into the tile (row-major, k ids per row); 8 columns at
    //    a time over all the tile's rows, so only 8 pages of the caller's matrix are live at once
    for (int j0 = 0; j0 < k; j0 += 8) {
      const int nc = k - j0 < 8 ? k - j0 : 8;
      const __mmask8 mc = (__mmask8)((1u << nc) - 1u);
      for (int b = 0; b < rows; b += 8) {
        const int nr = rows - b < 8 ? rows - b : 8;
        const __mmask8 mr = (__mmask8)((1u << nr) - 1u);
        __m512d c[8];
        for (int l = 0; l < 8; ++l) {
          if (l < nc) {
            const T* p = mat + (size_t)(j0 + l) * n + t0 + b;
            if (sizeof(T) == 8) c[l] = _mm512_maskz_loadu_pd(mr, p);
            else c[l] = _mm512_cvtepi32_pd(_mm256_maskz_loadu_epi32(mr, p));
          } else {
            c[l] = _mm512_setzero_pd();
          }
        }
        transpose8x8(c);
        for (int m = 0; m < nr; ++m) _mm512_mask_storeu_pd(tile + (size_t)(b + m) * k + j0, mc, c[m]);
      }
    }
    // 2. the tile is the edge range [t0*k, (t0+rows)*k): zero where u == 0, stream out
    long long e = t0 * k;
    const long long e_end = (t0 + rows) * k;
    const double* src = tile - e;
    while (e < e_end && ((uintptr_t)(dst + e) & 63)) {
      dst[e] = u[e] ? src[e] : 0.0;
      ++e;
    }
    for (; e + 8 <= e_end; e += 8) {
      const __m256i u32 = _mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i*)(u + e)));
      const __mmask8 nz = _mm256_cmpneq_epi32_mask(u32, zero);
      _mm512_stream_pd(dst + e, _mm512_maskz_loadu_pd(nz, src + e));
    }
    for (; e < e_end; ++e) dst[e] = u[e] ? src[e] : 0.0;
  }
  _mm_sfence();
}

static bool cpu_has_avx512() {
  static const bool ok = [] {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
           __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq");
  }();
  return ok;
}
#endif

void expand_column(const ExpandJob& job, int col, long long row_lo, long long row_hi) {
  const int k = job.k;
  if (k < 1 || k > 255 || col < 0 || col > 2 || row_hi <= row_lo) return;
  const int tr = kTileEdges / k;  // >= 8 rows: k <= 255 (counts are one byte)
  alignas(64) double tile[kTileEdges];
  double* dst = job.out + (size_t)col * job.E;
#ifdef GFICF_HOST_AVX512
  const bool wide = cpu_has_avx512();
  if (wide && col == 1) {
    if (job.elem == 8) expand_to_avx512<double>(job, row_lo, row_hi);
    else expand_to_avx512<int32_t>(job, row_lo, row_hi);
    return;
  }
  if (wide && k >= 8) {
    if (job.elem == 8) expand_column_avx512<double>(job, col, row_lo, row_hi);
    else expand_column_avx512<int32_t>(job, col, row_lo, row_hi);
    return;
  }
#endif
  for (long long r0 = row_lo; r0 < row_hi; r0 += tr) {
    const int rows = (int)((row_hi - r0) < tr ? (row_hi - r0) : tr);
    const uint8_t* u = job.counts + (size_t)(r0 - job.counts_row0) * k;
#ifdef GFICF_HOST_AVX512
    if (wide) {
      if (job.elem == 8)
        build_tile_avx512<double>(col, (const double*)job.idx, job.n, k, u, r0, rows, job.lut, tile);
      else
        build_tile_avx512<int32_t>(col, (const int32_t*)job.idx, job.n, k, u, r0, rows, job.lut, tile);
    } else
#endif
    if (job.elem == 8)
      build_tile<double>(col, (const double*)job.idx, job.n, k, u, r0, rows, job.lut, tile);
    else
      build_tile<int32_t>(col, (const int32_t*)job.idx, job.n, k, u, r0, rows, job.lut, tile);
    stream_copy(dst + (size_t)r0 * k, tile, (size_t)rows * k * sizeof(double));
  }
}

void expand_rows(const ExpandJob& job, long long row_lo, long long row_hi) {
  for (int col = 0; col < 3; ++col) expand_column(job, col, row_lo, row_hi);
}

// ------------------------------------------------------------------ f64 ids -> int32 ids
// The R matrix holds integers as doubles (src/RcppExports.cpp:65 coerces uwot's INTSXP); half of
// its bytes are zeros of the exponent/mantissa.  Converting on the host while the copy engine
// streams the converted pieces halves the H2D bytes.  Exact for every valid id; a value that is
// not an integer in int32 range (NaN, fractional, huge) becomes 0, which the device pre-pass
// rejects like any id outside [1, n] -- the error behaviour is unchanged.
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx512f,avx512vl"))) static void f64_to_i32_avx512(const double* src, int32_t* dst,
                                                                          size_t count) {
  size_t x = 0;
  for (; x + 8 <= count; x += 8) {
    const __m512d d = _mm512_loadu_pd(src + x);
    const __m256i v = _mm512_cvttpd_epi32(d);
    const __mmask8 ok = _mm512_cmp_pd_mask(_mm512_cvtepi32_pd(v), d, _CMP_EQ_OQ);
    _mm256_storeu_si256((__m256i*)(dst + x), _mm256_maskz_mov_epi32(ok, v));
  }
  for (; x < count; ++x) {
    const double d = src[x];
    const int32_t v = (d >= -2147483648.0 && d <= 2147483647.0) ? (int32_t)d : 0;
    dst[x] = ((double)v == d) ? v : 0;
  }
}
#endif

void f64_to_i32(const double* src, int32_t* dst, size_t count) {
#ifdef GFICF_HOST_AVX512
  if (cpu_has_avx512()) {
    f64_to_i32_avx512(src, dst, count);
    return;
  }
#endif
  for (size_t x = 0; x < count; ++x) {
    const double d = src[x];
    const int32_t v = (d >= -2147483648.0 && d <= 2147483647.0) ? (int32_t)d : 0;  // NaN fails both
    dst[x] = ((double)v == d) ? v : 0;
  }
}

const char* isa() {
#if defined(__x86_64__) && defined(__GNUC__)
  if (cpu_has_avx512()) return "avx512";
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx2")) return "avx2";
#endif
  return "scalar";
}

}  // namespace gficf_host
