// Micro-benchmark: ceiling of a random 128-byte-row gather (+ 24 B/edge streaming write) on B200.
// Same access pattern as jaccard_small_k_kernel at k=30 (8 lanes per row, 8 LDG.128 per lane in
// flight, index read first), but no hashing / probing at all.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <random>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

__device__ __forceinline__ int4 ld_na(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ld_ef(const int4* p) {
  int4 r;
  asm volatile("ld.global.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

template <int WARPS, bool WRITE, int LD = 0>
__global__ void __launch_bounds__(WARPS * 32) gather(const int* __restrict__ idx, long long n,
    double* __restrict__ f, double* __restrict__ t_, double* __restrict__ w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nwarps = (long long)gridDim.x * WARPS;
  const int sub4 = (lane & 7) * 4, grp = lane >> 3;
  long long row = (long long)blockIdx.x * WARPS + warp;
  int a_next = row < n ? __ldg(idx + row * 32 + lane) : 0;
  for (; row < n; row += nwarps) {
    const int a = a_next;
    const long long nrow = row + nwarps;
    a_next = nrow < n ? __ldg(idx + nrow * 32 + lane) : 0;
    int4 v[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int t = __shfl_sync(0xffffffffu, a, grp * 8 + s);
      const int4* p = reinterpret_cast<const int4*>(idx + (long long)t * 32 + sub4);
      v[s] = LD == 1 ? ld_na(p) : (LD == 2 ? ld_ef(p) : __ldg(p));
    }
    int acc = 0;
#pragma unroll
    for (int s = 0; s < 8; ++s) acc += v[s].x ^ v[s].y ^ v[s].z ^ v[s].w;
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (WRITE) {
      if (lane < 30) {
        const long long r = row * 30 + lane;
        __stcs(f + r, (double)(row + 1)); __stcs(t_ + r, (double)(a + 1)); __stcs(w + r, (double)acc);
      }
    } else if (acc == 0x7fffffff) f[0] = 1.0;
  }
}

// ---- TMA variant: every neighbour row is one 128-byte cp.async.bulk (UBLKCP) into shared memory,
// completion through a per-warp mbarrier, double-buffered across rows; the data is then read back
// from shared memory.  Same gathered bytes, same stores.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) gather_tma(const int* __restrict__ idx, long long n,
    double* __restrict__ f, double* __restrict__ t_, double* __restrict__ w) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // per warp: 2 x (32 rows x 128 B) + 2 mbarriers
  unsigned char* base = smem + (size_t)warp * (2 * 4096 + 16);
  int* buf[2] = {(int*)base, (int*)(base + 4096)};
  unsigned long long* mbar = (unsigned long long*)(base + 8192);
  const unsigned mb[2] = {smem_u32(mbar), smem_u32(mbar + 1)};
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb[0]));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb[1]));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const long long nwarps = (long long)gridDim.x * WARPS;
  const int sub = (lane & 7), grp = lane >> 3;
  long long row = (long long)blockIdx.x * WARPS + warp;
  auto issue = [&](int b, int a) {
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb[b]), "r"(30 * 128) : "memory");
    __syncwarp();
    if (lane < 30) {
      const int* src = idx + (long long)a * 32;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];"
                   ::"r"(smem_u32(buf[b] + lane * 32)), "l"(src), "r"(mb[b]) : "memory");
    }
  };
  int a = row < n ? __ldg(idx + row * 32 + lane) : 0;
  if (row < n) issue(0, a);
  int cur = 0;
  unsigned phase[2] = {0, 0};
  for (; row < n; row += nwarps) {
    const long long nrow = row + nwarps;
    const int a_next = nrow < n ? __ldg(idx + nrow * 32 + lane) : 0;
    if (nrow < n) issue(cur ^ 1, a_next);
    // wait for this row's 30 copies
    unsigned done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(mb[cur]), "r"(phase[cur]) : "memory");
    }
    phase[cur] ^= 1;
    int acc = 0;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int e = grp * 8 + s;
      if (e < 30) {
        const int4 v = *reinterpret_cast<const int4*>(buf[cur] + e * 32 + sub * 4);
        acc += v.x ^ v.y ^ v.z ^ v.w;
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    __syncwarp();  // all lanes have read the buffer before it is refilled two iterations later
    if (lane < 30) {
      const long long r = row * 30 + lane;
      __stcs(f + r, (double)(row + 1)); __stcs(t_ + r, (double)(a + 1)); __stcs(w + r, (double)acc);
    }
    a = a_next;
    cur ^= 1;
  }
}

int main() {
  const long long n = 4000000; const long long E = n * 30;
  std::vector<int> h(n * 32);
  std::mt19937_64 rng(1);
  for (long long i = 0; i < n * 32; ++i) h[i] = (int)(rng() % n);
  int* idx; double* out; uint8_t* flush;
  CK(cudaMalloc(&idx, n * 32 * 4)); CK(cudaMalloc(&out, 3 * E * 8)); CK(cudaMalloc(&flush, 256 << 20));
  CK(cudaMemcpy(idx, h.data(), n * 32 * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  {
    const size_t smem = 8 * (2 * 4096 + 16);
    CK(cudaFuncSetAttribute(gather_tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int ctas : {2, 3}) {
      float tot = 0;
      for (int it = 0; it < 6; ++it) {
        CK(cudaMemset(flush, it, 256 << 20));
        cudaEventRecord(e0);
        gather_tma<8><<<148 * ctas, 256, smem>>>(idx, n, out, out + E, out + 2 * E);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it) tot += ms;
      }
      printf("TMA bulk rows, %d CTA/SM (2 row-sets in flight per warp)  %.3f ms  %.2f Gedges/s\n", ctas, tot / 5, E / (tot / 5) / 1e6);
    }
  }
  for (int cfg = 0; cfg < 8; ++cfg) {
    float tot = 0;
    for (int it = 0; it < 6; ++it) {
      CK(cudaMemset(flush, it, 256 << 20));
      cudaEventRecord(e0);
      switch (cfg) {
        case 0: gather<8, true><<<148 * 4, 256>>>(idx, n, out, out + E, out + 2 * E); break;
        case 1: gather<8, true><<<148 * 6, 256>>>(idx, n, out, out + E, out + 2 * E); break;
        case 2: gather<8, true><<<148 * 8, 256>>>(idx, n, out, out + E, out + 2 * E); break;
        case 3: gather<8, false><<<148 * 8, 256>>>(idx, n, out, out + E, out + 2 * E); break;
        case 4: gather<8, true><<<148 * 3, 256>>>(idx, n, out, out + E, out + 2 * E); break;
        case 5: gather<8, true, 1><<<148 * 4, 256>>>(idx, n, out, out + E, out + 2 * E); break;
        case 6: gather<8, true, 1><<<148 * 8, 256>>>(idx, n, out, out + E, out + 2 * E); break;
        case 7: gather<8, true, 2><<<148 * 4, 256>>>(idx, n, out, out + E, out + 2 * E); break;
      }
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it) tot += ms;
    }
    const char* nm[] = {"4 CTA/SM write", "6 CTA/SM write", "8 CTA/SM write", "8 CTA/SM read-only", "3 CTA/SM write", "4 CTA/SM no_allocate", "8 CTA/SM no_allocate", "4 CTA/SM ld.global na (coherent)"};
    printf("%-20s %.3f ms  %.2f Gedges/s  %.0f GB/s algorithmic\n", nm[cfg], tot / 5, E / (tot / 5) / 1e6, E * 148.0 / (tot / 5) / 1e6);
  }
  return 0;
}
