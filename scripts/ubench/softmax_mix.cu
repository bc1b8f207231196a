// Micro-benchmark: how the per-score instruction mix of the attention softmax warps shares one SM sub-partition.
// Every thread runs 8 independent "pairs"; per pair and iteration: NF packed-fp32 ops (FFMA2), NA ALU-pipe ops
// (FMNMX + F2FP) and NM MUFU.EX2.  Reported: cycles per pair-iteration per sub-partition (all resident warps
// together) for 1, 2, 4 warps per sub-partition - the floor of each pipe and how well they overlap.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../painter_b200/csrc softmax_mix.cu -o softmax_mix
#include "common.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace pk;

template <int NF, int NA, int NM>
__global__ void __launch_bounds__(512) mix_kernel(float* out, int iters, long long* cyc, float a, float b) {
  f32x2 x[8];
  float mx = -1e30f;
  uint32_t acc = 0;
#pragma unroll
  for (int p = 0; p < 8; ++p) x[p] = pack_f2(0.001f * threadIdx.x + p, 0.002f * threadIdx.x - p);
  const f32x2 a2 = pack_f2(a, a), b2 = pack_f2(b, b);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int p = 0; p < 8; ++p) {
#pragma unroll
      for (int k = 0; k < NF; ++k) x[p] = fma_f2(x[p], a2, b2);
      float lo, hi;
      unpack_f2(x[p], lo, hi);
      if (NA >= 1) mx = fmaxf(mx, fmaxf(lo, hi));
      if (NM >= 1) lo = fast_exp2(lo);
      if (NM >= 2) hi = fast_exp2(hi);
      if (NA >= 2) acc ^= pack_bf16x2(lo, hi);
      x[p] = pack_f2(lo, hi);
    }
  }
  const long long t1 = clock64();
  float s = mx + __uint_as_float(acc);
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    float lo, hi;
    unpack_f2(x[p], lo, hi);
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + (threadIdx.x >> 5)] = t1 - t0;
}

template <int NF, int NA, int NM>
void run(const char* label, float* out, long long* cyc) {
  const int iters = 2000;
  printf("%-44s", label);
  for (int wps = 1; wps <= 4; wps *= 2) {
    const int threads = 128 * wps;
    mix_kernel<NF, NA, NM><<<148, threads>>>(out, iters, cyc, 0.999f, 0.0001f);
    mix_kernel<NF, NA, NM><<<148, threads>>>(out, iters, cyc, 0.999f, 0.0001f);
    cudaDeviceSynchronize();
    long long h[148 * 16];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long m = 0;
    for (int b = 0; b < 148; ++b)
      for (int w = 0; w < threads / 32; ++w) m = h[b * 16 + w] > m ? h[b * 16 + w] : m;
    // cycles per pair per sub-partition: wps warps each did iters * 8 pairs
    printf("  wps=%d %6.2f", wps, static_cast<double>(m) / (static_cast<double>(iters) * 8 * wps));
  }
  printf("   (cycles per pair-iteration per sub-partition)\n");
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 16 * 8);
  run<0, 0, 2>("2 MUFU.EX2", out, cyc);
  run<0, 0, 1>("1 MUFU.EX2", out, cyc);
  run<4, 0, 0>("4 FFMA2", out, cyc);
  run<8, 0, 0>("8 FFMA2", out, cyc);
  run<0, 2, 0>("FMNMX3 + F2FP", out, cyc);
  run<3, 2, 2>("fwd mix: 3 FFMA2 + FMNMX3 + F2FP + 2 EX2", out, cyc);
  run<3, 0, 2>("3 FFMA2 + 2 EX2", out, cyc);
  run<6, 2, 2>("dq mix: 6 FFMA2 + 2 ALU + 2 EX2", out, cyc);
  run<6, 0, 2>("6 FFMA2 + 2 EX2", out, cyc);
  run<10, 2, 2>("10 FFMA2 + 2 ALU + 2 EX2", out, cyc);
  run<10, 2, 1>("10 FFMA2 + 2 ALU + 1 EX2", out, cyc);
  run<14, 2, 0>("14 FFMA2 + 2 ALU", out, cyc);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
