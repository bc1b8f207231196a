// painter_b200 — the per-sample work of the training data path that does not need file I/O, on the device
// (SURVEY §8 f.4): BEiT-style block masks (Painter/util/masking_generator.py:15-93, drawn per sample in
// Painter/data/pairdataset.py:183-188, incl. the half-mask alternative) and the per-task `valid` weighting maps
// (pairdataset.py:154-181).  At 8 x 120 images/s the host loader is the next wall; these two run on the GPU
// next to the step they feed and cost microseconds.
#include "common.cuh"
#include "host_common.h"
#include "../../include/painter_b200.h"

namespace pk {

// counter-based uniform in [0, 1): splitmix64 finaliser over (seed, sample, counter)
__device__ __forceinline__ float uni01(unsigned long long seed, unsigned sample, unsigned& counter) {
  unsigned long long x = seed * 0x9E3779B97F4A7C15ull + (static_cast<unsigned long long>(sample) << 32) + counter++;
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return static_cast<float>(x >> 40) * (1.0f / 16777216.0f);
}

// One warp per sample.  The algorithm is masking_generator.py's, decision for decision: blocks of random area
// in [min_patches, min(remaining, max_patches)] and log-uniform aspect ratio are tried (<= 10 attempts per block)
// until `num_masking` cells are covered or no block fits; the count is then made exact by flipping uniformly chosen
// cells (selection sampling = np.random.choice(..., replace=False)).  With probability half_mask_ratio the sample
// gets the bottom-half mask instead (pairdataset.py:149,183-186).  Only the random number source differs from the
// reference (counter-based hash instead of Python's Mersenne twister), so parity is distributional.
__global__ void __launch_bounds__(32)
block_masks_kernel(int* __restrict__ out, int B, int H, int W, int num_masking, int min_patches, int max_patches,
                   float log_ar_lo, float log_ar_hi, float half_mask_ratio, unsigned long long seed) {
  extern __shared__ unsigned char m[];   // H * W cells
  const unsigned s = blockIdx.x;
  const int lane = threadIdx.x;
  const int n = H * W;
  unsigned ctr = 0;
  for (int i = lane; i < n; i += 32) m[i] = 0;
  __syncwarp();
  const bool half = uni01(seed, s, ctr) < half_mask_ratio;       // every lane evaluates the same stream
  if (half) {
    for (int i = lane; i < n; i += 32) m[i] = (i / W) >= H / 2 ? 1 : 0;
    __syncwarp();
  } else {
    int count = 0;
    while (count < num_masking) {
      int max_mask = num_masking - count;
      if (max_mask > max_patches) max_mask = max_patches;
      int delta = 0;
      for (int attempt = 0; attempt < 10 && delta == 0; ++attempt) {
        const float area = min_patches + (max_mask - min_patches) * uni01(seed, s, ctr);
        const float ar = expf(log_ar_lo + (log_ar_hi - log_ar_lo) * uni01(seed, s, ctr));
        const int hh = static_cast<int>(rintf(sqrtf(area * ar)));
        const int ww = static_cast<int>(rintf(sqrtf(area / ar)));
        if (ww < W && hh < H) {
          int top = static_cast<int>(uni01(seed, s, ctr) * (H - hh + 1));
          int left = static_cast<int>(uni01(seed, s, ctr) * (W - ww + 1));
          if (top > H - hh) top = H - hh;
          if (left > W - ww) left = W - ww;
          int covered = 0;
          for (int i = lane; i < hh * ww; i += 32) covered += m[(top + i / ww) * W + left + i % ww];
          covered = __reduce_add_sync(0xffffffffu, covered);
          const int fresh = hh * ww - covered;
          if (fresh > 0 && fresh <= max_mask) {
            for (int i = lane; i < hh * ww; i += 32) m[(top + i / ww) * W + left + i % ww] = 1;
            __syncwarp();
            delta = fresh;
          }
        }
      }
      if (delta == 0) break;
      count += delta;
    }
    // exact count: flip `need` uniformly chosen cells of the right kind (selection sampling, lane 0)
    if (count != num_masking && lane == 0) {
      const unsigned char from = count > num_masking ? 1 : 0;
      int need = count > num_masking ? count - num_masking : num_masking - count;
      int remaining = from ? count : n - count;
      for (int i = 0; i < n && need > 0; ++i) {
        if (m[i] != from) continue;
        if (uni01(seed, s, ctr) * remaining < need) {
          m[i] = 1 - from;
          --need;
        }
        --remaining;
      }
    }
    __syncwarp();
  }
  for (int i = lane; i < n; i += 32) out[static_cast<size_t>(s) * n + i] = m[i];
}

// rule per sample (pairdataset.py:154-181): 0 ones | 1 valid[t < thr] = 0 | 2 valid[t > thr] = 10, all 0 if the
// foreground count is < 300 | 3 all 0 if the foreground count is < 300.  thr[rule-specific 3 channels] in `thr`.
__global__ void __launch_bounds__(256)
valid_count_kernel(const float* __restrict__ tgt, const int* __restrict__ rule, const float* __restrict__ thr,
                   int* __restrict__ fg, int plane3, int plane) {
  const int b = blockIdx.y;
  const int r = rule[b];
  if (r < 2) return;
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane3; i += gridDim.x * blockDim.x)
    c += tgt[static_cast<size_t>(b) * plane3 + i] > thr[b * 3 + i / plane] ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(fg + b, c);
}
__global__ void __launch_bounds__(256)
valid_apply_kernel(const float* __restrict__ tgt, const int* __restrict__ rule, const float* __restrict__ thr,
                   const int* __restrict__ fg, float* __restrict__ valid, int plane3, int plane) {
  const int b = blockIdx.y;
  const int r = rule[b];
  const bool dead = r >= 2 && fg[b] < 300;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane3; i += gridDim.x * blockDim.x) {
    const float t = tgt[static_cast<size_t>(b) * plane3 + i];
    const float th = thr[b * 3 + i / plane];
    float v = 1.f;
    if (r == 1) v = t < th ? 0.f : 1.f;
    else if (r == 2) v = t > th ? 10.f : 1.f;
    if (dead) v = 0.f;
    valid[static_cast<size_t>(b) * plane3 + i] = v;
  }
}

}  // namespace pk

using namespace pk;

extern "C" int pk_block_masks(int* out, int B, int H, int W, int num_masking, int min_patches, int max_patches,
                              float log_ar_lo, float log_ar_hi, float half_mask_ratio, unsigned long long seed,
                              void* stream) {
  PK_CHECK(out && B > 0 && H > 0 && W > 0 && num_masking >= 0 && num_masking <= H * W && min_patches >= 1 &&
               max_patches >= min_patches && H * W <= 48 * 1024,
           "pk_block_masks: bad arguments");
  block_masks_kernel<<<B, 32, static_cast<size_t>(H) * W, static_cast<cudaStream_t>(stream)>>>(
      out, B, H, W, num_masking, min_patches, max_patches, log_ar_lo, log_ar_hi, half_mask_ratio, seed);
  PK_LAUNCH_CHECK("pk_block_masks");
  return 0;
}

extern "C" int pk_valid_maps(const float* targets, const int* rule_dev, const float* thr_dev, int* fg_zeroed,
                             float* valid, int B, int H, int W, void* stream) {
  PK_CHECK(targets && rule_dev && thr_dev && fg_zeroed && valid && B > 0, "pk_valid_maps: bad arguments");
  const int plane = H * W, plane3 = 3 * plane;
  int gx = (plane3 + 255) / 256;
  if (gx > sm_count() * 4) gx = sm_count() * 4;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  valid_count_kernel<<<dim3(gx, B), 256, 0, st>>>(targets, rule_dev, thr_dev, fg_zeroed, plane3, plane);
  PK_LAUNCH_CHECK("pk_valid_maps(count)");
  valid_apply_kernel<<<dim3(gx, B), 256, 0, st>>>(targets, rule_dev, thr_dev, fg_zeroed, valid, plane3, plane);
  PK_LAUNCH_CHECK("pk_valid_maps(apply)");
  return 0;
}
