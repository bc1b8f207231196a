// painter_b200 — device-side pre/post-processing of the inference wrappers: everything the reference
// does in numpy / PIL-free torch around the forward of seggpt_engine.inference_image / inference_video
// (SegGPT/SegGPT_inference/seggpt_engine.py:56-181) and painter_inference_segm.run_one_image
// (Painter/eval/ade20k_semantic/painter_inference_segm.py:67-93), so that a frame costs one small H2D copy, the
// forward, and one D2H copy of the finished uint8 image.
//
// The reference does this arithmetic in float64 (numpy broadcasting of the float64 mean/std arrays); the kernels
// below keep float64 wherever the reference result is consumed as float64, so results are bit-comparable.
// All kernels are HBM-bound streaming passes over <= a few MB.
#include "common.cuh"
#include "host_common.h"
#include "../../include/painter_b200.h"

namespace pk {

__constant__ double c_mean[3] = {0.485, 0.456, 0.406};
__constant__ double c_std[3] = {0.229, 0.224, 0.225};

// Source pixel in [0, 1] as a double: uint8 / 255. (np.array(PIL) / 255.), float32 or float64 as stored.
__device__ __forceinline__ double load_unit(const void* src, int dtype, size_t idx) {
  if (dtype == 0) return static_cast<double>(static_cast<const uint8_t*>(src)[idx]) / 255.0;
  if (dtype == 1) return static_cast<double>(static_cast<const float*>(src)[idx]);
  return static_cast<const double*>(src)[idx];
}

// canvas[p, c, y, x] (fp32 NCHW, [P, 3, 2S, S]) = ((y < S ? top[p] : bottom[p])[y % S, x, c] - mean[c]) / std[c]
// top / bottom: per-member pointers to HWC images [S, S, 3] in [0,1] (or uint8); seggpt_engine.py:75-91 (stitch,
// normalise), :29-34 (nhwc -> nchw, .float()).
struct CanvasSrc {
  const void* top[16];
  const void* bottom[16];
  int top_dtype[16];
  int bottom_dtype[16];
};
__global__ void __launch_bounds__(256)
stitch_normalize_kernel(const CanvasSrc src, float* __restrict__ out, int P, int S) {
  const size_t total = static_cast<size_t>(P) * 3 * 2 * S * S;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % S);
    size_t r = i / S;
    const int y = static_cast<int>(r % (2 * S));
    r /= (2 * S);
    const int c = static_cast<int>(r % 3), p = static_cast<int>(r / 3);
    const bool is_top = y < S;
    const int yy = is_top ? y : y - S;
    const size_t sidx = (static_cast<size_t>(yy) * S + x) * 3 + c;
    const double v = is_top ? load_unit(src.top[p], src.top_dtype[p], sidx)
                            : load_unit(src.bottom[p], src.bottom_dtype[p], sidx);
    out[i] = static_cast<float>(__dsub_rn(v, c_mean[c]) / c_std[c]);
  }
}

// nhwc (fp64 or fp32, already normalised) -> nchw fp32: torch.einsum('nhwc->nchw', x).float() (seggpt_engine.py:29-34)
__global__ void __launch_bounds__(256)
nhwc_to_nchw_f32_kernel(const void* __restrict__ in, int in_is_f64, float* __restrict__ out, int n, int H, int W) {
  const size_t total = static_cast<size_t>(n) * 3 * H * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    size_t r = i / W;
    const int y = static_cast<int>(r % H);
    r /= H;
    const int c = static_cast<int>(r % 3), b = static_cast<int>(r / 3);
    const size_t s = ((static_cast<size_t>(b) * H + y) * W + x) * 3 + c;
    out[i] = in_is_f64 ? static_cast<float>(static_cast<const double*>(in)[s]) : static_cast<const float*>(in)[s];
  }
}

// out[y, x, c] (fp64 [H/2, W, 3]) = clip((pred_pixel(sample 0, y + H/2, x, c) * std[c] + mean[c]) * 255, 0, 255)
// patch: patchified prediction [B, h*w, p*p*3] fp32 (inner order (pr, pc, c), models_painter.py:355-368);
// = unpatchify -> 'nchw->nhwc' -> bottom half -> de-normalise (seggpt_engine.py:48-52).  Optionally also writes
// bin[y, x] = (mean_c(out) > 128) (the video path's next-frame target, seggpt_engine.py:164-169).
__global__ void __launch_bounds__(256)
seg_postprocess_kernel(const float* __restrict__ patch, double* __restrict__ out, float* __restrict__ bin, int h,
                       int w, int p) {
  const int H = h * p, W = w * p, Hh = H / 2;
  const size_t total = static_cast<size_t>(Hh) * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W), y = static_cast<int>(i / W) + Hh;
    const int ti = y / p, tj = x / p, pr = y % p, pc = x % p;
    const float* src = patch + (static_cast<size_t>(ti) * w + tj) * (p * p * 3) + (pr * p + pc) * 3;
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // explicit IEEE mul / add: numpy rounds the product before the addition (no fused multiply-add)
      double v = __dmul_rn(__dadd_rn(__dmul_rn(static_cast<double>(src[c]), c_std[c]), c_mean[c]), 255.0);
      v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
      out[i * 3 + c] = v;
      s += v;
    }
    if (bin) bin[i] = (s / 3.0 > 128.0) ? 1.f : 0.f;
  }
}

// ATen nearest (legacy) source index: upsample_nearest's nearest_idx with scale = float(in) / out
__device__ __forceinline__ int nearest_idx(int dst, int in_size, int out_size) {
  if (out_size == in_size) return dst;
  if (out_size == 2 * in_size) return dst >> 1;
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  const int s = static_cast<int>(floorf(dst * scale));
  return s < in_size - 1 ? s : in_size - 1;
}
// dst[y, x, c] (uint8 [OH, OW, 3]) = uint8(image[y, x, c] * (0.6 * nearest(seg)[y, x, c] / 255 + 0.4))
// (seggpt_engine.py:97-102: F.interpolate(mode='nearest') to the original size, alpha blend, astype(uint8))
__global__ void __launch_bounds__(256)
nearest_blend_kernel(const double* __restrict__ seg, int SH, int SW, const uint8_t* __restrict__ image,
                     uint8_t* __restrict__ dst, int OH, int OW) {
  const size_t total = static_cast<size_t>(OH) * OW * 3;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % 3);
    const size_t r = i / 3;
    const int x = static_cast<int>(r % OW), y = static_cast<int>(r / OW);
    const int sy = nearest_idx(y, SH, OH), sx = nearest_idx(x, SW, OW);
    const double o = seg[(static_cast<size_t>(sy) * SW + sx) * 3 + c];
    const double v = __dmul_rn(static_cast<double>(image[i]), __dadd_rn(__dmul_rn(0.6, o) / 255.0, 0.4));
    dst[i] = static_cast<uint8_t>(v);   // numpy astype(uint8): truncation (values are within [0, 255])
  }
}

// dst[y, x, c] (uint8 [OH, OW, 3]) = uint8(int(bilinear(seg)[y, x, c]))   (painter_inference_segm.py:89-92:
// F.interpolate(mode='bilinear', align_corners=False) on the float64 tensor, .int(), astype(uint8))
__global__ void __launch_bounds__(256)
bilinear_u8_kernel(const double* __restrict__ seg, int SH, int SW, uint8_t* __restrict__ dst, int OH, int OW) {
  const size_t total = static_cast<size_t>(OH) * OW * 3;
  const double sch = static_cast<double>(SH) / OH, scw = static_cast<double>(SW) / OW;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % 3);
    const size_t r = i / 3;
    const int x = static_cast<int>(r % OW), y = static_cast<int>(r / OW);
    double fy = sch * (y + 0.5) - 0.5, fx = scw * (x + 0.5) - 0.5;
    if (fy < 0.0) fy = 0.0;
    if (fx < 0.0) fx = 0.0;
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const int y1 = y0 + (y0 < SH - 1 ? 1 : 0), x1 = x0 + (x0 < SW - 1 ? 1 : 0);
    const double ly = fy - y0, lx = fx - x0, hy = 1.0 - ly, hx = 1.0 - lx;
    auto at = [&](int yy, int xx) { return seg[(static_cast<size_t>(yy) * SW + xx) * 3 + c]; };
    const double v = hy * (hx * at(y0, x0) + lx * at(y0, x1)) + ly * (hx * at(y1, x0) + lx * at(y1, x1));
    dst[i] = static_cast<uint8_t>(static_cast<int>(v));
  }
}

static inline int stream_grid(size_t total) {
  size_t g = (total + 255) / 256;
  const size_t cap = static_cast<size_t>(sm_count()) * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace pk

using namespace pk;

extern "C" int pk_stitch_normalize(const void* const* top, const int* top_dtype, const void* const* bottom,
                                   const int* bottom_dtype, float* out, int P, int S, void* stream) {
  PK_CHECK(top && bottom && top_dtype && bottom_dtype && out && P >= 1 && P <= 16 && S > 0,
           "pk_stitch_normalize: bad arguments (1 <= P <= 16)");
  CanvasSrc src;
  for (int p = 0; p < P; ++p) {
    PK_CHECK(top[p] && bottom[p] && top_dtype[p] >= 0 && top_dtype[p] <= 2 && bottom_dtype[p] >= 0 &&
                 bottom_dtype[p] <= 2, "pk_stitch_normalize: bad member %d", p);
    src.top[p] = top[p];
    src.bottom[p] = bottom[p];
    src.top_dtype[p] = top_dtype[p];
    src.bottom_dtype[p] = bottom_dtype[p];
  }
  const size_t total = static_cast<size_t>(P) * 3 * 2 * S * S;
  stitch_normalize_kernel<<<stream_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, out, P, S);
  PK_LAUNCH_CHECK("pk_stitch_normalize");
  return 0;
}

extern "C" int pk_nhwc_to_nchw_f32(const void* in, int in_is_f64, float* out, int n, int H, int W, void* stream) {
  PK_CHECK(in && out && n > 0 && H > 0 && W > 0, "pk_nhwc_to_nchw_f32: bad arguments");
  const size_t total = static_cast<size_t>(n) * 3 * H * W;
  nhwc_to_nchw_f32_kernel<<<stream_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, in_is_f64, out, n, H,
                                                                                             W);
  PK_LAUNCH_CHECK("pk_nhwc_to_nchw_f32");
  return 0;
}

extern "C" int pk_seg_postprocess(const float* patch, double* out, float* bin_or_null, int h, int w, int p,
                                  void* stream) {
  PK_CHECK(patch && out && h > 0 && w > 0 && p > 0 && h % 2 == 0, "pk_seg_postprocess: bad arguments");
  const size_t total = static_cast<size_t>(h) * p / 2 * w * p;
  seg_postprocess_kernel<<<stream_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(patch, out, bin_or_null, h,
                                                                                            w, p);
  PK_LAUNCH_CHECK("pk_seg_postprocess");
  return 0;
}

extern "C" int pk_nearest_blend(const double* seg, int SH, int SW, const uint8_t* image, uint8_t* dst, int OH, int OW,
                                void* stream) {
  PK_CHECK(seg && image && dst && SH > 0 && SW > 0 && OH > 0 && OW > 0, "pk_nearest_blend: bad arguments");
  nearest_blend_kernel<<<stream_grid(static_cast<size_t>(OH) * OW * 3), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      seg, SH, SW, image, dst, OH, OW);
  PK_LAUNCH_CHECK("pk_nearest_blend");
  return 0;
}

extern "C" int pk_bilinear_u8(const double* seg, int SH, int SW, uint8_t* dst, int OH, int OW, void* stream) {
  PK_CHECK(seg && dst && SH > 0 && SW > 0 && OH > 0 && OW > 0, "pk_bilinear_u8: bad arguments");
  bilinear_u8_kernel<<<stream_grid(static_cast<size_t>(OH) * OW * 3), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      seg, SH, SW, dst, OH, OW);
  PK_LAUNCH_CHECK("pk_bilinear_u8");
  return 0;
}
