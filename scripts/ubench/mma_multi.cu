// Micro-benchmark: is the ~48-cycle floor per tcgen05.mma (scripts/ubench/mma_rate.cu) a limit of the ISSUING THREAD
// or of the SM's tensor core?  `nw` warps each issue n MMAs (M=128, K=16, given N) into their own accumulators at the
// same time, with descriptors either computed in the loop or hoisted; reported: cycles from the common start to the
// last completion, per MMA (all warps together).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../painter_b200/csrc mma_multi.cu -o mma_multi
#include "common.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace pk;

__global__ void __launch_bounds__(160) multi_kernel(int N, int n_mma, int nw, int hoist, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 32768, bar = base + 98304, holder = base + 98304 + 64;
  volatile uint32_t* holder_gen =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (base - smem_u32(smem_raw)) + 98304 + 64);
  __shared__ long long t_end[4];
  __shared__ long long t_beg[4];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int w = 0; w < 4; ++w) mbar_init(bar + 8 * w, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder_gen;
  if (warp < nw && (threadIdx.x & 31) == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint64_t a0 = make_sdesc(sA + warp * 8192, 16, 1024);
    const uint64_t b0 = make_sdesc(sB + warp * 8192, 16, 1024);
    const uint32_t acc = tmem + warp * 128;
    uint64_t da[4], db[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      da[k] = sdesc_add(a0, k * 32);
      db[k] = sdesc_add(b0, k * 32);
    }
    const long long t0 = clock64();
    if (hoist) {
      for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(acc, da[k], db[k], idesc, 1u);
      }
    } else {
      for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(acc, sdesc_add(a0, k * 32), sdesc_add(b0, k * 32), idesc, 1u);
      }
    }
    umma_commit(bar + 8 * warp);
    while (!mbar_try_wait(bar + 8 * warp, 0)) {
    }
    t_beg[warp] = t0;
    t_end[warp] = clock64();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    long long b = t_beg[0], e = t_end[0];
    for (int w = 1; w < nw; ++w) {
      b = t_beg[w] < b ? t_beg[w] : b;
      e = t_end[w] > e ? t_end[w] : e;
    }
    out[0] = e - b;
  }
  if (warp == 4) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  for (int N : {32, 64, 112})
    for (int hoist = 0; hoist < 2; ++hoist)
      for (int nw = 1; nw <= 4; nw *= 2) {
        const int n = 64;
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
          multi_kernel<<<1, 160, 100 * 1024 + 1024>>>(N, n, nw, hoist, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("error %s\n", cudaGetErrorString(e));
            return 1;
          }
        }
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("M=128 N=%3d hoisted_desc=%d issuing_warps=%d  n=%d each  total=%6lld cyc  per-MMA (all warps) = %.1f  floor=%d\n",
               N, hoist, nw, n, h, (double)h / (n * nw), 128 * N / 256);
      }
  return 0;
}
