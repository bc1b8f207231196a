// painter_b200 — the step either side of backward: fused multi-tensor AdamW and the global gradient
// norm that `misc.NativeScalerWithGradNormCount.__call__` (Painter/util/misc.py:252-278) computes for clipping.
//
// Reference semantics: torch.optim.AdamW as built by main_train.py:344-348 over lr_decay.param_groups_lrd groups
// (per-group lr (= lr * lr_scale) and weight_decay), decoupled weight decay:
//     p <- p (1 - lr wd);  m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2;
//     p <- p - (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The gradient may carry a device-side multiplier (1 / loss scale, and/or the clip coefficient), so unscale + clip +
// step are ONE pass over p, g, m, v: 7 fp32 streams, HBM-bound (10.4 GB per step for the 370.7 M parameters).
#include <cuda_fp16.h>
#include "common.cuh"
#include "host_common.h"
#include "../../include/painter_b200.h"

namespace pk {

constexpr int OPT_CHUNK = 16384;   // elements per block: 256 threads x 16 float4

struct OptTensor {       // mirrors PkOptTensor (include/painter_b200.h)
  float* p;
  float* g;
  float* m;
  float* v;
  __nv_bfloat16* w16;    // optional bf16 copy of the updated parameter (the GEMM / attention operand), or null
  long long n;
  float lr, wd;
};
static_assert(sizeof(OptTensor) == sizeof(PkOptTensor), "PkOptTensor layout");

__global__ void __launch_bounds__(256)
adamw_kernel(const OptTensor* __restrict__ tensors, const int2* __restrict__ chunks, float b1, float omb1, float b2,
             float omb2, float eps, float inv_bc1, float inv_sqrt_bc2, const float* __restrict__ gscale,
             float gscale_cap, int zero_grad, const float* __restrict__ found_inf,
             const float* __restrict__ bc_dev) {
  if (found_inf && found_inf[0] != 0.f) return;   // GradScaler semantics: an overflowed step is skipped entirely
  if (bc_dev) {   // CUDA-graph replay: the step-dependent bias corrections come from device memory
    inv_bc1 = bc_dev[0];
    inv_sqrt_bc2 = bc_dev[1];
  }
  const int2 ck = chunks[blockIdx.x];
  const OptTensor t = tensors[ck.x];
  const long long base = static_cast<long long>(ck.y) * OPT_CHUNK;
  const long long left = t.n - base;
  const int cnt = left < OPT_CHUNK ? static_cast<int>(left) : OPT_CHUNK;
  float gs = gscale ? gscale[0] : 1.f;
  if (gscale_cap > 0.f && gs > gscale_cap) gs = gscale_cap;   // clip coefficient: min(1, max_norm / (norm + 1e-6))
  const float decay = 1.f - t.lr * t.wd, step = t.lr * inv_bc1;
  float* p = t.p + base;
  float* g = t.g + base;
  float* m = t.m + base;
  float* v = t.v + base;
  __nv_bfloat16* w16 = t.w16 ? t.w16 + base : nullptr;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                     reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(w16) & 7) == 0;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= gs;
    mm = b1 * mm + omb1 * gg;
    vv = b2 * vv + omb2 * gg * gg;
    pp = pp * decay - step * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
  };
  if (vec) {
    const int n4 = cnt >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i],
             vv = reinterpret_cast<float4*>(v)[i];
      const float4 gg = reinterpret_cast<const float4*>(g)[i];
      upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y);
      upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
      if (w16) {
        uint2 u;
        u.x = pack_bf16x2(pp.x, pp.y);
        u.y = pack_bf16x2(pp.z, pp.w);
        reinterpret_cast<uint2*>(w16)[i] = u;
      }
      if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < cnt; i += 256) {
      upd(p[i], g[i], m[i], v[i]);
      if (w16) w16[i] = __float2bfloat16_rn(p[i]);
      if (zero_grad) g[i] = 0.f;
    }
  } else {
    for (int i = threadIdx.x; i < cnt; i += 256) {
      upd(p[i], g[i], m[i], v[i]);
      if (w16) w16[i] = __float2bfloat16_rn(p[i]);
      if (zero_grad) g[i] = 0.f;
    }
  }
}

// out[0] += sum over all tensors of g^2 (fp32 partials per block, one atomic per block)
__global__ void __launch_bounds__(256)
sumsq_kernel(const OptTensor* __restrict__ tensors, const int2* __restrict__ chunks, float* __restrict__ out) {
  const int2 ck = chunks[blockIdx.x];
  const OptTensor t = tensors[ck.x];
  const long long base = static_cast<long long>(ck.y) * OPT_CHUNK;
  const long long left = t.n - base;
  const int cnt = left < OPT_CHUNK ? static_cast<int>(left) : OPT_CHUNK;
  const float* g = t.g + base;
  float s = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int n4 = cnt >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) {
      const float4 gg = reinterpret_cast<const float4*>(g)[i];
      s += (gg.x * gg.x + gg.y * gg.y) + (gg.z * gg.z + gg.w * gg.w);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < cnt; i += 256) s += g[i] * g[i];
  } else {
    for (int i = threadIdx.x; i < cnt; i += 256) s += g[i] * g[i];
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[w];
    atomicAdd(out, a);
  }
}

}  // namespace pk

extern "C" int pk_opt_chunk_elems(void) { return pk::OPT_CHUNK; }

extern "C" int pk_adamw_step(const PkOptTensor* tensors_dev, const int* chunks_dev, int nchunks, double beta1,
                             double beta2, double eps, int step, const float* gscale, float gscale_cap,
                             int zero_grad, const float* found_inf, void* stream) {
  using namespace pk;
  PK_CHECK(tensors_dev && chunks_dev && nchunks > 0 && step >= 1, "pk_adamw_step: bad arguments");
  // hyper-parameters arrive as doubles and 1 - beta / the bias corrections are formed in double, as torch does with
  // its Python floats: 1.f - 0.999f is off by 1.3e-5 relative, which would show in exp_avg_sq
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  adamw_kernel<<<nchunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const OptTensor*>(tensors_dev), reinterpret_cast<const int2*>(chunks_dev),
      static_cast<float>(beta1), static_cast<float>(1.0 - beta1), static_cast<float>(beta2),
      static_cast<float>(1.0 - beta2), static_cast<float>(eps), static_cast<float>(1.0 / bc1),
      static_cast<float>(1.0 / sqrt(bc2)), gscale, gscale_cap, zero_grad, found_inf, nullptr);
  PK_LAUNCH_CHECK("pk_adamw_step");
  return 0;
}

// The same update for a captured training step: the two step-dependent scalars 1 / (1 - beta1^t) and
// 1 / sqrt(1 - beta2^t) are read from bc_dev[0..1], which the host rewrites before every replay; learning rates and
// weight decays already live in the device table.
extern "C" int pk_adamw_step_graph(const PkOptTensor* tensors_dev, const int* chunks_dev, int nchunks, double beta1,
                                   double beta2, double eps, const float* bc_dev, const float* gscale,
                                   float gscale_cap, int zero_grad, const float* found_inf, void* stream) {
  using namespace pk;
  PK_CHECK(tensors_dev && chunks_dev && nchunks > 0 && bc_dev, "pk_adamw_step_graph: bad arguments");
  adamw_kernel<<<nchunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const OptTensor*>(tensors_dev), reinterpret_cast<const int2*>(chunks_dev),
      static_cast<float>(beta1), static_cast<float>(1.0 - beta1), static_cast<float>(beta2),
      static_cast<float>(1.0 - beta2), static_cast<float>(eps), 1.0f, 1.0f, gscale, gscale_cap, zero_grad, found_inf,
      bc_dev);
  PK_LAUNCH_CHECK("pk_adamw_step_graph");
  return 0;
}

extern "C" int pk_grad_sumsq(const PkOptTensor* tensors_dev, const int* chunks_dev, int nchunks, float* out_zeroed,
                             void* stream) {
  using namespace pk;
  PK_CHECK(tensors_dev && chunks_dev && nchunks > 0 && out_zeroed, "pk_grad_sumsq: bad arguments");
  sumsq_kernel<<<nchunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const OptTensor*>(tensors_dev), reinterpret_cast<const int2*>(chunks_dev), out_zeroed);
  PK_LAUNCH_CHECK("pk_grad_sumsq");
  return 0;
}

namespace pk {
__global__ void droppath_scales_kernel(const void* __restrict__ r, int dtype_code, const float* __restrict__ keep,
                                       float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float k = keep[i];
  float v;
  if (dtype_code == 1) {
    v = k + __bfloat162float(static_cast<const __nv_bfloat16*>(r)[i]);
    v = __bfloat162float(__float2bfloat16_rn(v));
  } else if (dtype_code == 2) {
    v = k + __half2float(static_cast<const __half*>(r)[i]);
    v = __half2float(__float2half_rn(v));
  } else {
    v = k + static_cast<const float*>(r)[i];
  }
  out[i] = floorf(v) / k;
}
}  // namespace pk

extern "C" int pk_droppath_scales(const void* r, int dtype_code, const float* keep, float* out, int n,
                                  void* stream) {
  PK_CHECK(r && keep && out && n > 0 && dtype_code >= 0 && dtype_code <= 2, "pk_droppath_scales: bad arguments");
  pk::droppath_scales_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(r, dtype_code, keep,
                                                                                              out, n);
  PK_LAUNCH_CHECK("pk_droppath_scales");
  return 0;
}
