// scan_kernels.cuh -- exclusive prefix sums shared by the compaction of the serial export
// (jaccard_kernels.cuh), the graph build (snn_kernels.cuh) and the network kernels
// (network_kernels.cuh).  Integer arithmetic only.
//
// With GFICF_CUDA_EMU defined the header compiles as plain C++ against tests/cuda_emu/cuda_emu.h
// (a fibre-per-thread emulation of the CUDA execution model used by the CPU test suite to run the
// kernels' logic without a GPU); the product build never defines it.
#pragma once
#ifdef GFICF_CUDA_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

// dynamic shared memory of a kernel: `extern __shared__` on the device, a per-launch buffer in the emulation
#ifdef GFICF_CUDA_EMU
#define GFICF_DYNAMIC_SMEM_T(type, name) type* name = reinterpret_cast<type*>(cuda_emu::dynamic_smem())
#else
#define GFICF_DYNAMIC_SMEM_T(type, name) extern __shared__ type name[]
#endif
#define GFICF_DYNAMIC_SMEM(name) GFICF_DYNAMIC_SMEM_T(unsigned char, name)

namespace gficf {

constexpr unsigned kFull = 0xFFFFFFFFu;

// exclusive scan of chunk counts in place, single CTA; total -> *n_written
__global__ void __launch_bounds__(1024)
compact_scan_kernel(long long* __restrict__ chunk_cnt, long long nchunks,
                    long long* __restrict__ n_written) {
  __shared__ long long wtot[32];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long b = 0; b < nchunks; b += 1024) {
    const long long x = b + threadIdx.x;
    const long long v = x < nchunks ? chunk_cnt[x] : 0;
    long long inc = v;
    for (int m = 1; m < 32; m <<= 1) {
      const long long o = __shfl_up_sync(kFull, inc, m);
      if (lane >= m) inc += o;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      long long w = wtot[lane];
      for (int m = 1; m < 32; m <<= 1) {
        const long long o = __shfl_up_sync(kFull, w, m);
        if (lane >= m) w += o;
      }
      wtot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const long long carry = carry_s;
    const long long before = carry + (warp ? wtot[warp - 1] : 0) + inc - v;
    if (x < nchunks) chunk_cnt[x] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wtot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_written = carry_s;
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

}  // namespace gficf
