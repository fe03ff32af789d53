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
      __stcs((double2*)(f + e), make_double2(u0 ? r0 + 1.0 : 0.0, u1 ? r1 + 1.0 : 0.0));
      __stcs((double2*)(t_ + e), make_double2(u0 ? t0 + 1.0 : 0.0, u1 ? t1 + 1.0 : 0.0));
      __stcs((double2*)(w + e), make_double2(lut[u0], lut[u1]));
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

int main() {
  const long long n = 4000000; const int k = 30, kp = 32; const long long E = n * k;
  int* idx; uint8_t* u; double* out; uint8_t* flush;
  CK(cudaMalloc(&idx, n * kp * 4)); CK(cudaMalloc(&u, E)); CK(cudaMalloc(&out, 3 * E * 8)); CK(cudaMalloc(&flush, 256 << 20));
  CK(cudaMemset(idx, 1, n * kp * 4)); CK(cudaMemset(u, 3, E));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int grid_mul : {6, 8}) {
    for (int mode : {0, 3, 5, 6}) {
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
