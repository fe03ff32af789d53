// tools/host_expand_bench.cpp -- host-side microbenchmark of gficf_host::expand_rows (the CPU half of
// the "counts over PCIe" output mode): GB/s of output written by T threads, next to a plain
// non-temporal fill of the same bytes (the host's write ceiling).
//   g++ -O3 -std=c++17 -pthread -I gficf_b200/csrc tools/host_expand_bench.cpp gficf_b200/csrc/host_expand.cpp -o /tmp/hxb
//   /tmp/hxb [n=2000000] [k=30] [threads...]
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "host_expand.h"

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 2000000;
  const int k = argc > 2 ? atoi(argv[2]) : 30;
  const long long E = n * k;
  std::vector<double> mat((size_t)E);
  std::vector<uint8_t> cnt((size_t)E);
  double* out = (double*)aligned_alloc(64, (size_t)E * 24);
  unsigned long long s = 88172645463325252ull;
  for (long long x = 0; x < E; ++x) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    mat[x] = (double)(s % n + 1);
    cnt[x] = (uint8_t)((s >> 32) % (k + 1));
  }
  memset(out, 0, (size_t)E * 24);
  double lut[256];
  gficf_host::fill_weight_table(k, lut);
  const bool i32 = getenv("HXB_I32") != nullptr;
  std::vector<int32_t> mat32;
  if (i32) { mat32.resize((size_t)E); for (long long x = 0; x < E; ++x) mat32[x] = (int32_t)mat[x]; }
  gficf_host::ExpandJob job{i32 ? (const void*)mat32.data() : (const void*)mat.data(), i32 ? 4 : 8, n, k, cnt.data(), 0, out, E, lut};
  printf("isa %s, n=%lld k=%d, output %.2f GB\n", gficf_host::isa(), n, k, E * 24 / 1e9);
  std::vector<int> ts;
  for (int a = 3; a < argc; ++a) ts.push_back(atoi(argv[a]));
  if (ts.empty()) ts = {1, 2, 4, 8};
  const long long chunk = 16384;
  const int only_col = getenv("HXB_COL") ? atoi(getenv("HXB_COL")) : -1;  // time one column only (GB/s then counts 3x)
  for (int T : ts) {
    for (int what = 1; what >= 0; --what) {
      double best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        std::atomic<long long> next(0);
        const double t0 = now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
          th.emplace_back([&] {
            static thread_local std::vector<double> zero(chunk * 30, 1.0);
            for (;;) {
              const long long lo = next.fetch_add(chunk);
              if (lo >= n) return;
              const long long hi = lo + chunk < n ? lo + chunk : n;
              if (what == 0) {
                if (only_col >= 0) gficf_host::expand_column(job, only_col, lo, hi);
                else gficf_host::expand_rows(job, lo, hi);
              } else {  // ceiling: stream the same number of bytes from an L2-resident source
                for (int c = 0; c < 3; ++c)
                  for (long long r = lo; r < hi; r += 1024) {
                    const long long m = (r + 1024 < hi ? 1024 : hi - r) * k;
                    gficf_host::stream_copy(out + c * E + r * k, zero.data(), (size_t)m * 8);
                  }
              }
            }
          });
        for (auto& x : th) x.join();
        const double dt = now() - t0;
        if (dt < best) best = dt;
      }
      printf("  T=%2d %-12s %7.1f ms  %6.1f GB/s written  %.2f ns/edge/thread\n", T, what ? "nt-fill" : "expand_rows",
             best * 1e3, E * 24 / best / 1e9, best * 1e9 * T / E);
    }
  }
  // spot check
  long long bad = 0;
  for (long long x = 0; x < E; x += 997) {
    const long long i = x / k; const int j = (int)(x % k);
    const int u = cnt[x];
    const double ef = u ? (double)(i + 1) : 0.0, et = u ? mat[(size_t)j * n + i] : 0.0, ew = u / (2.0 * k - u);
    bad += out[x] != ef || out[E + x] != et || out[2 * E + x] != ew;
  }
  printf("spot check mismatches: %lld\n", bad);
  return bad != 0;
}
