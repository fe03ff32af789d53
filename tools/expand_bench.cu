// Micro-benchmark: what limits the counts -> (from,to,w) expand kernel?  Variants of the store path.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

template <int MODE>  // 0: stcs 8B; 1: plain 8B; 2: stcs, no idx read; 3: 16B stcs (2 edges/thread); 4: stwt
__global__ void __launch_bounds__(256) expand(const int* __restrict__ idx, int k, int kp, long long rows,
    const uint8_t* d_u, double* __restrict__ f, double* __restrict__ t_, double* __restrict__ w) {
  __shared__ double lut[256];
  if ((int)threadIdx.x <= k) lut[threadIdx.x] = (double)threadIdx.x / (2.0 * k - threadIdx.x);
  __syncthreads();
  const long long total = rows * k;
  if (MODE == 3 || MODE == 6) {
    const long long stride = (long long)gridDim.x * 256 * 2;
    long long g0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 2;
    for (long long e = g0; e < total; e += stride) {
      const int u0 = __ldcg(d_u + e), u1 = __ldcg(d_u + e + 1);
      const long long r0 = e / k, r1 = (e + 1) / k;
      const int j0 = (int)(e - r0 * k), j1 = (int)(e + 1 - r1 * k);
      const int t0 = (MODE == 6 || u0) ? __ldg(idx + r0 * kp + j0) : 0, t1 = (MODE == 6 || u1) ? __ldg(idx + r1 * kp + j1) : 0;
      if (MODE == 6) {  // default-policy stores
        *(double2*)(f + e) = make_double2(u0 ? r0 + 1.0 : 0.0, u1 ? r1 + 1.0 : 0.0);
        *(double2*)(t_ + e) = make_double2(u0 ? t0 + 1.0 : 0.0, u1 ? t1 + 1.0 : 0.0);
        *(double2*)(w + e) = make_double2(lut[u0], lut[u1]);
      } else {
      __stcs((double2*)(f + e), make_double2(u0 ? r0 + 1.0 : 0.0, u1 ? r1 + 1.0 : 0.0));
      __stcs((double2*)(t_ + e), make_double2(u0 ? t0 + 1.0 : 0.0, u1 ? t1 + 1.0 : 0.0));
      __stcs((double2*)(w + e), make_double2(lut[u0], lut[u1]));
      }
    }
    return;
  }
  const long long stride = (long long)gridDim.x * 256;
  const long long g0 = (long long)blockIdx.x * 256 + threadIdx.x;
  long long row = g0 / k; int j = (int)(g0 % k);
  const long long d_row = stride / k; const int d_j = (int)(stride % k);
#pragma unroll 4
  for (long long e = g0; e < total; e += stride) {
    const int u = (int)__ldcg(d_u + e);
    const bool nz = u > 0;
    const int t = (MODE == 2) ? 7 : (MODE == 5 ? __ldg(idx + row * (long long)kp + j) : (nz ? __ldg(idx + row * (long long)kp + j) : 0));
    const double a = nz ? (double)(row + 1) : 0.0, b = nz ? (double)(t + 1) : 0.0, c = lut[u];
    if (MODE == 1) { f[e] = a; t_[e] = b; w[e] = c; }
    else if (MODE == 4) { __stwt(f + e, a); __stwt(t_ + e, b); __stwt(w + e, c); }
    else { __stcs(f + e, a); __stcs(t_ + e, b); __stcs(w + e, c); }
    row += d_row; j += d_j; if (j >= k) { j -= k; ++row; }
  }
}

// r02 variants: what separates the 3-array expand store pattern from a plain fill?
//  7: fill only, three arrays, 8-byte st.cs (no loads at all)
//  8: fill only, three arrays, 16-byte default stores
//  9: 4 consecutive edges per thread: one 32-bit count load, 4 id loads, 2 x 16-byte st.cs per array
// 10: like the shipped kernel (mode 5) but default-policy stores
// 11: three passes, one output array per pass (single write stream each)
template <int MODE>
__global__ void __launch_bounds__(256) expand2(const int* __restrict__ idx, int k, int kp, long long rows,
    const uint8_t* d_u, double* __restrict__ f, double* __restrict__ t_, double* __restrict__ w, int pass) {
  __shared__ double lut[256];
  if ((int)threadIdx.x <= k) lut[threadIdx.x] = (double)threadIdx.x / (2.0 * k - threadIdx.x);
  __syncthreads();
  const long long total = rows * k;
  if (MODE == 7) {
    const long long stride = (long long)gridDim.x * 256;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += stride) {
      __stcs(f + e, 1.0); __stcs(t_ + e, 2.0); __stcs(w + e, 3.0);
    }
  } else if (MODE == 8) {
    const long long stride = (long long)gridDim.x * 256 * 2;
    for (long long e = ((long long)blockIdx.x * 256 + threadIdx.x) * 2; e < total; e += stride) {
      *(double2*)(f + e) = make_double2(1.0, 1.0); *(double2*)(t_ + e) = make_double2(2.0, 2.0);
      *(double2*)(w + e) = make_double2(3.0, 3.0);
    }
  } else if (MODE == 9) {
    const long long stride = (long long)gridDim.x * 256 * 4;
    for (long long e = ((long long)blockIdx.x * 256 + threadIdx.x) * 4; e < total; e += stride) {
      const unsigned u4 = __ldcg((const unsigned*)(d_u + e));
      long long r = e / k; int j = (int)(e - r * k);
      double a[4], b[4], c[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int u = (u4 >> (8 * q)) & 0xFF;
        const int t = __ldg(idx + r * (long long)kp + j);
        a[q] = u ? (double)(r + 1) : 0.0; b[q] = u ? (double)(t + 1) : 0.0; c[q] = lut[u];
        if (++j == k) { j = 0; ++r; }
      }
      __stcs((double2*)(f + e), make_double2(a[0], a[1])); __stcs((double2*)(f + e + 2), make_double2(a[2], a[3]));
      __stcs((double2*)(t_ + e), make_double2(b[0], b[1])); __stcs((double2*)(t_ + e + 2), make_double2(b[2], b[3]));
      __stcs((double2*)(w + e), make_double2(c[0], c[1])); __stcs((double2*)(w + e + 2), make_double2(c[2], c[3]));
    }
  } else {
    const long long stride = (long long)gridDim.x * 256;
    const long long g0 = (long long)blockIdx.x * 256 + threadIdx.x;
    long long row = g0 / k; int j = (int)(g0 % k);
    const long long d_row = stride / k; const int d_j = (int)(stride % k);
#pragma unroll 4
    for (long long e = g0; e < total; e += stride) {
      const int u = (int)__ldcg(d_u + e);
      const bool nz = u > 0;
      if (MODE == 10) {
        const int t = __ldg(idx + row * (long long)kp + j);
        f[e] = nz ? (double)(row + 1) : 0.0; t_[e] = nz ? (double)(t + 1) : 0.0; w[e] = lut[u];
      } else {  // 11: one array per pass
        if (pass == 0) __stcs(f + e, nz ? (double)(row + 1) : 0.0);
        else if (pass == 1) { const int t = __ldg(idx + row * (long long)kp + j); __stcs(t_ + e, nz ? (double)(t + 1) : 0.0); }
        else __stcs(w + e, lut[u]);
      }
      row += d_row; j += d_j; if (j >= k) { j -= k; ++row; }
    }
  }
}

int main() {
  const long long n = 4000000; const int k = 30, kp = 32; const long long E = n * k;
  int* idx; uint8_t* u; double* out; uint8_t* flush;
  CK(cudaMalloc(&idx, n * kp * 4)); CK(cudaMalloc(&u, E)); CK(cudaMalloc(&out, 3 * E * 8)); CK(cudaMalloc(&flush, 256 << 20));
  CK(cudaMemset(idx, 1, n * kp * 4)); CK(cudaMemset(u, 3, E));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // r02: does the distance between the three output arrays matter (DRAM channel / bank mapping)?
  for (long long padb : {0ll, 256ll, 4096ll + 256, 65536ll + 512, (1ll << 20) + 4096 + 256, (32ll << 20) + 768}) {
    double* o2; CK(cudaMalloc(&o2, 3 * E * 8 + 4 * padb + 1024));
    const long long pd = padb / 8;
    for (int mode : {7, 8, 6}) {
      float tot = 0;
      for (int it = 0; it < 6; ++it) {
        CK(cudaMemset(flush, it, 256 << 20));
        cudaEventRecord(e0);
        const int g = 148 * 6;
        if (mode == 7) expand2<7><<<g, 256>>>(idx, k, kp, n, u, o2, o2 + E + pd, o2 + 2 * E + 2 * pd, 0);
        else if (mode == 8) expand2<8><<<g, 256>>>(idx, k, kp, n, u, o2, o2 + E + pd, o2 + 2 * E + 2 * pd, 0);
        else expand<6><<<g, 256>>>(idx, k, kp, n, u, o2, o2 + E + pd, o2 + 2 * E + 2 * pd);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it) tot += ms;
      }
      printf("array gap +%lld B, mode %d (6 CTA/SM): %.3f ms\n", padb, mode, tot / 5);
    }
    cudaFree(o2);
  }
  for (int grid_mul : {4, 8}) {
    for (int mode : {7, 8, 9, 10, 11}) {
      float tot = 0;
      for (int it = 0; it < 6; ++it) {
        CK(cudaMemset(flush, it, 256 << 20));
        cudaEventRecord(e0);
        const int g = 148 * grid_mul;
        switch (mode) {
          case 7: expand2<7><<<g, 256>>>(idx, k, kp, n, u, out, out + E, out + 2 * E, 0); break;
          case 8: expand2<8><<<g, 256>>>(idx, k, kp, n, u, out, out + E, out + 2 * E, 0); break;
          case 9: expand2<9><<<g, 256>>>(idx, k, kp, n, u, out, out + E, out + 2 * E, 0); break;
          case 10: expand2<10><<<g, 256>>>(idx, k, kp, n, u, out, out + E, out + 2 * E, 0); break;
          case 11: for (int p = 0; p < 3; ++p) expand2<11><<<g, 256>>>(idx, k, kp, n, u, out, out + E, out + 2 * E, p); break;
        }
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it) tot += ms;
      }
      printf("grid %2dx148 mode %2d: %.3f ms  %.0f GB/s\n", grid_mul, mode, tot / 5, E * 29.0 / (tot / 5) / 1e6);
    }
  }
  for (int grid_mul : {5, 6, 7}) {
    for (int mode : {3, 5, 6}) {
      float tot = 0;
      for (int it = 0; it < 6; ++it) {
        CK(cudaMemset(flush, it, 256 << 20));
        cudaEventRecord(e0);
        const int g = 148 * grid_mul;
        switch (mode) {
          case 0: expand<0><<<g, 256>>>(idx, k, kp, n, u, out, out + E, out + 2 * E); break;
          case 1: expand<1><<<g, 256>>>(idx, k, kp, n, u, out, out + E, out + 2 * E); break;
          case 2: expand<2><<<g, 256>>>(idx, k, kp, n, u, out, out + E, out + 2 * E); break;
          case 3: expand<3><<<g, 256>>>(idx, k, kp, n, u, out, out + E, out + 2 * E); break;
          case 4: expand<4><<<g, 256>>>(idx, k, kp, n, u, out, out + E, out + 2 * E); break;
          case 5: expand<5><<<g, 256>>>(idx, k, kp, n, u, out, out + E, out + 2 * E); break;
          case 6: expand<6><<<g, 256>>>(idx, k, kp, n, u, out, out + E, out + 2 * E); break;
        }
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it) tot += ms;
      }
      printf("grid %2dx148 mode %d: %.3f ms  %.0f GB/s\n", grid_mul, mode, tot / 5, E * 29.0 / (tot / 5) / 1e6);
    }
  }
  return 0;
}
