// Micro-benchmark + layout check: tcgen05.mma with the A operand in TMEM (TS form), kind::f16, cta_group::1, M = 128.
//   1. correctness: A[128 x 64] bf16 is written to TMEM by its row owners (tcgen05.st.32x32b, two K elements per
//      32-bit column: element 2c in the low half, 2c+1 in the high half) and multiplied with B from shared memory
//      (K-major and MN-major); the result is compared with the SS form on the same data.
//   2. rate: cycles per MMA for a back-to-back burst (N = 64, 112, 128; K = 16 per MMA), next to the SS form.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../painter_b200/csrc mma_ts.cu -o mma_ts
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
using namespace pk;

// smem: sA (K-major, 128 rows x 64 k, SW128) | sB K-major [256 rows][64 k] | sBt MN-major [64 k][64.. N] groups
constexpr uint32_t OFF_A = 0, OFF_B = 16384, OFF_BT = 16384 + 32768, OFF_BAR = OFF_BT + 32768;

__device__ __forceinline__ float a_val(int r, int k) { return static_cast<float>(((r * 7 + k * 3) % 17) - 8) * 0.125f; }
__device__ __forceinline__ float b_val(int n, int k) { return static_cast<float>(((n * 5 + k * 11) % 13) - 6) * 0.25f; }

__global__ void __launch_bounds__(128) ts_kernel(int N, int b_mn, int n_mma, float* out_ss, float* out_ts,
                                                 long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base + OFF_A, sB = base + OFF_B, sBt = base + OFF_BT, bar = base + OFF_BAR, holder = bar + 16;
  volatile uint32_t* holder_gen = reinterpret_cast<volatile uint32_t*>(gen + OFF_BAR + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
  // A: K-major SW128: row r at r * 128 B, 16-byte chunk c at ((c ^ (r & 7)) << 4)
  for (int k = 0; k < 64; ++k) {
    __nv_bfloat16 v = __float2bfloat16(a_val(row, k));
    const uint32_t off = row * 128 + ((((k >> 3) ^ (row & 7)) << 4)) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(gen + OFF_A + off) = v;
  }
  // B K-major: [n][k], rows n < 256 (two per thread)
  for (int n = row; n < 256; n += 128)
    for (int k = 0; k < 64; ++k) {
      const uint32_t off = n * 128 + ((((k >> 3) ^ (n & 7)) << 4)) + (k & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(gen + OFF_B + off) = __float2bfloat16(b_val(n, k));
    }
  // B MN-major: [k][n] in 64-wide n groups: group g at g * 8192 B (64 k rows x 128 B), k row at k * 128 B
  for (int n = row; n < 256; n += 128)
    for (int k = 0; k < 64; ++k) {
      const int g = n >> 6, nn = n & 63;
      const uint32_t off = g * 8192 + k * 128 + ((((nn >> 3) ^ (k & 7)) << 4)) + (nn & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(gen + OFF_BT + off) = __float2bfloat16(b_val(n, k));
    }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(holder, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder_gen;
  const uint32_t tSS = tmem, tTS = tmem + 256, tA = tmem + 448;   // A: 32 columns (64 bf16 per row)
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  {
    uint32_t v[16];
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 16) {
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = pack_bf16x2(a_val(row, 2 * (c0 + c)), a_val(row, 2 * (c0 + c) + 1));
      tmem_st_x16(tA + lane_addr + c0, v);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t idesc = make_idesc_bf16(128, N, false, b_mn != 0);
  const uint64_t a0 = make_sdesc(sA, 16, 1024);
  const uint64_t b0 = b_mn ? make_sdesc(sBt, 8192, 1024) : make_sdesc(sB, 16, 1024);
  const uint32_t b_step = b_mn ? 2048u : 32u;
  if (threadIdx.x == 32) {
    // correctness pass: 4 K-slices each
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ss(tSS, sdesc_add(a0, k * 32), sdesc_add(b0, k * b_step), idesc, k != 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ts(tTS, tA + k * 8, sdesc_add(b0, k * b_step), idesc, k != 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16], w[16];
    tmem_ld_x16(tSS + lane_addr + c0, v);
    tmem_ld_x16(tTS + lane_addr + c0, w);
    tmem_wait_ld();
    for (int c = 0; c < 16; ++c) {
      out_ss[row * 256 + c0 + c] = __uint_as_float(v[c]);
      out_ts[row * 256 + c0 + c] = __uint_as_float(w[c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // rate passes (results garbage: accumulate into the same tiles)
  if (threadIdx.x == 32) {
    for (int form = 0; form < 2; ++form) {
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (form == 0) umma_ss(tSS, sdesc_add(a0, k * 32), sdesc_add(b0, k * b_step), idesc, 1u);
          else umma_ts(tTS, tA + k * 8, sdesc_add(b0, k * b_step), idesc, 1u);
        }
      }
      const long long t1 = clock64();
      umma_commit(bar);
      while (!mbar_try_wait(bar, (form + 1) & 1)) {
      }
      const long long t2 = clock64();
      cyc[form * 2] = t1 - t0;
      cyc[form * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  float *d_ss, *d_ts;
  long long* d_c;
  cudaMalloc(&d_ss, 128 * 256 * 4);
  cudaMalloc(&d_ts, 128 * 256 * 4);
  cudaMalloc(&d_c, 64);
  cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  float* h_ss = (float*)malloc(128 * 256 * 4);
  float* h_ts = (float*)malloc(128 * 256 * 4);
  for (int bmn = 0; bmn < 2; ++bmn)
    for (int N : {64, 112, 128}) {
      if (bmn && N == 112) continue;   // MN-major B is laid out in 64-wide groups here
      for (int n : {8, 64}) {
        cudaMemset(d_ss, 0, 128 * 256 * 4);
        cudaMemset(d_ts, 0, 128 * 256 * 4);
        ts_kernel<<<1, 128, OFF_BAR + 1024 + 64>>>(N, bmn, n, d_ss, d_ts, d_c);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("error %s\n", cudaGetErrorString(e));
          return 1;
        }
        long long c[4];
        cudaMemcpy(c, d_c, 32, cudaMemcpyDeviceToHost);
        cudaMemcpy(h_ss, d_ss, 128 * 256 * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(h_ts, d_ts, 128 * 256 * 4, cudaMemcpyDeviceToHost);
        double max_ref = 0, err_ss = 0, err_ts = 0;
        for (int r = 0; r < 128; ++r)
          for (int j = 0; j < N; ++j) {
            double ref = 0;
            for (int k = 0; k < 64; ++k)
              ref += (double)((((r * 7 + k * 3) % 17) - 8) * 0.125f) * (double)((((j * 5 + k * 11) % 13) - 6) * 0.25f);
            max_ref = fmax(max_ref, fabs(ref));
            err_ss = fmax(err_ss, fabs(h_ss[r * 256 + j] - ref));
            err_ts = fmax(err_ts, fabs(h_ts[r * 256 + j] - ref));
          }
        printf("M=128 N=%3d b_mn=%d n=%2d  |ref|max=%.2f err_ss=%.3g err_ts=%.3g   SS issue %.1f total %.1f /mma   "
               "TS issue %.1f total %.1f /mma   math floor %d\n",
               N, bmn, n, max_ref, err_ss, err_ts, (double)c[0] / n, (double)c[1] / n, (double)c[2] / n,
               (double)c[3] / n, 128 * N / 256);
      }
    }
  return 0;
}
