// Micro-benchmark: dispatch rate of tcgen05.mma (kind::f16, cta_group::1, SS mode) for small N.
// One CTA per launch; an elected thread issues `n` MMAs (K = 16 each) back to back into one accumulator,
// commits, waits; cycles / n is printed for several (N, a_major, b_major) combinations.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../painter_b200/csrc mma_rate.cu -o mma_rate
#include "common.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace pk;

// time from the first issue to the completion seen by the issuing lane
__global__ void __launch_bounds__(128) rate2_kernel(int N, int a_mn, int b_mn, int n_mma, int distinct_acc,
                                                    long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 32768, bar = base + 98304, holder = base + 98304 + 16;
  volatile uint32_t* holder_gen =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (base - smem_u32(smem_raw)) + 98304 + 16);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder_gen;
  if (threadIdx.x == 32) {
    const uint32_t idesc = make_idesc_bf16(128, N, a_mn != 0, b_mn != 0);
    const uint64_t a0 = a_mn ? make_sdesc(sA, 8192, 1024) : make_sdesc(sA, 16, 1024);
    const uint64_t b0 = b_mn ? make_sdesc(sB, 8192, 1024) : make_sdesc(sB, 16, 1024);
    const uint32_t a_step = a_mn ? 2048u : 32u, b_step = b_mn ? 2048u : 32u;
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_ss(tmem + (distinct_acc ? ((i >> 2) & 1) * 256 : 0), sdesc_add(a0, k * a_step), sdesc_add(b0, k * b_step),
                idesc, 1u);
    }
    const long long t1 = clock64();
    umma_commit(bar);
    while (!mbar_try_wait(bar, 0)) {
    }
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  const int Ns[] = {32, 64, 112, 128, 256};
  for (int amn = 0; amn < 2; ++amn)
    for (int bmn = 0; bmn < 2; ++bmn)
      for (int N : Ns)
        for (int n : {8, 64})
          for (int da = 0; da < 2; ++da) {
            long long h[2] = {0, 0};
            for (int rep = 0; rep < 2; ++rep) {
              rate2_kernel<<<1, 128, 100 * 1024 + 1024>>>(N, amn, bmn, n, da, d);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) {
                printf("error %s\n", cudaGetErrorString(e));
                return 1;
              }
            }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("M=128 N=%3d a_mn=%d b_mn=%d n=%2d alt_acc=%d  issue=%6lld cyc (%.1f/mma)  total=%6lld cyc (%.1f/mma)  floor=%d\n",
                   N, amn, bmn, n, da, h[0], (double)h[0] / n, h[1], (double)h[1] / n, 128 * N / 256);
          }
  return 0;
}
