// host_expand.cpp -- see host_expand.h.  Plain host C++ (compiled by the host compiler, no CUDA).
#include "host_expand.h"

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

const char* isa() {
#if defined(__x86_64__) && defined(__GNUC__)
  if (cpu_has_avx512()) return "avx512";
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx2")) return "avx2";
#endif
  return "scalar";
}

}  // namespace gficf_host
