// painter_b200 — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = A[M,K] . B[N,K]^T      bf16 operands, fp32 accumulation in TMEM
//
// One CTA per SM, 256 threads:
//   warp 0      TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      MMA issuer     (one thread: tcgen05.mma 128 x BN x 16, commit -> mbarriers)
//   warp 2      TMEM allocator (2 accumulator stages of BN fp32 columns)
//   warps 4..7  epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile
// i+1.  Operands may be K-major (row-major [rows,K]) or MN-major (row-major [K,rows]); the latter is
// what the backward GEMMs (dgrad: B = W[out,in]; wgrad: A = dY[tokens,out], B = X[tokens,in]) need, so no
// transposed copies of weights or activations are ever materialised.
//
// Reference call sites replaced: every nn.Linear on the path (models_painter.py:60-61,76,87; timm Mlp
// fc1/fc2 via :201; decoder_embed :327,423) and autograd's mm backward for them.
#include "common.cuh"
#include "host_common.h"
#include "../../include/painter_b200.h"

namespace pk {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KiB
constexpr int GEMM_THREADS = 256;

struct GemmArgs {
  int M, N, K;
  int BN;
  int stages;
  int transA, transB;
  int num_m_tiles, num_n_tiles;
  PkEpilogue epi;
};

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&x)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_bf16x2(x[q * 8 + 0], x[q * 8 + 1]);
    u.y = pack_bf16x2(x[q * 8 + 2], x[q * 8 + 3]);
    u.z = pack_bf16x2(x[q * 8 + 4], x[q * 8 + 5]);
    u.w = pack_bf16x2(x[q * 8 + 6], x[q * 8 + 7]);
    d[q] = u;
  }
}

// One 32-column chunk of one output row.
__device__ __forceinline__ void gemm_epilogue_chunk(const PkEpilogue& e, int row, int col,
                                                    const uint32_t (&v)[32]) {
  float x[32];
  const float alpha = e.alpha;
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) * alpha;
  if (e.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(e.bias + col);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = __ldg(b4 + q);
      x[q * 4 + 0] += b.x;
      x[q * 4 + 1] += b.y;
      x[q * 4 + 2] += b.z;
      x[q * 4 + 3] += b.w;
    }
  }
  const size_t off = static_cast<size_t>(row) * e.ldc + col;
  switch (e.kind) {
    case PK_EPI_BF16: {
      store_bf16x32(reinterpret_cast<__nv_bfloat16*>(e.out) + off, x);
    } break;
    case PK_EPI_F32: {
      float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + off);
      if (e.accumulate) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 o = d[q];
          o.x += x[q * 4 + 0];
          o.y += x[q * 4 + 1];
          o.z += x[q * 4 + 2];
          o.w += x[q * 4 + 3];
          d[q] = o;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          d[q] = make_float4(x[q * 4 + 0], x[q * 4 + 1], x[q * 4 + 2], x[q * 4 + 3]);
      }
    } break;
    case PK_EPI_GELU: {
      store_bf16x32(reinterpret_cast<__nv_bfloat16*>(e.out) + off, x);
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = gelu_erf(bf16_round(x[j]));
      store_bf16x32(reinterpret_cast<__nv_bfloat16*>(e.out2) + off, x);
    } break;
    case PK_EPI_RESID: {
      const float sc = e.rowscale ? __ldg(e.rowscale + row / e.rows_per_group) : 1.0f;
      const float4* r4 = reinterpret_cast<const float4*>(
          reinterpret_cast<const float*>(e.aux) + static_cast<size_t>(row) * e.ld_aux + col);
      float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + off);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 r = r4[q];
        d[q] = make_float4(fmaf(sc, x[q * 4 + 0], r.x), fmaf(sc, x[q * 4 + 1], r.y),
                           fmaf(sc, x[q * 4 + 2], r.z), fmaf(sc, x[q * 4 + 3], r.w));
      }
    } break;
    case PK_EPI_DGELU: {
      const uint4* z4 = reinterpret_cast<const uint4*>(
          reinterpret_cast<const __nv_bfloat16*>(e.aux) + static_cast<size_t>(row) * e.ld_aux + col);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 z = z4[q];
        const uint32_t w[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float z0 = __uint_as_float(w[t] << 16);
          const float z1 = __uint_as_float(w[t] & 0xFFFF0000u);
          x[q * 8 + t * 2 + 0] *= gelu_erf_grad(z0);
          x[q * 8 + t * 2 + 1] *= gelu_erf_grad(z1);
        }
      }
      store_bf16x32(reinterpret_cast<__nv_bfloat16*>(e.out) + off, x);
    } break;
    case PK_EPI_PIXSHUF: {
      const int hw = e.ps_h * e.ps_w;
      const int b = row / hw, t = row - b * hw;
      const int i = t / e.ps_w, j = t - i * e.ps_w;
      const int pc = e.ps_p * e.ps_c;
      const int r = col / pc, rem = col - r * pc;
      const int s = rem / e.ps_c, c = rem - s * e.ps_c;
      const size_t o = ((static_cast<size_t>(b) * (e.ps_h * e.ps_p) + i * e.ps_p + r) *
                            (static_cast<size_t>(e.ps_w) * e.ps_p) +
                        j * e.ps_p + s) *
                           e.ps_c +
                       c;
      store_bf16x32(reinterpret_cast<__nv_bfloat16*>(e.out) + o, x);
    } break;
    default:
      break;
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_base = smem_u32(smem_raw);
  const uint32_t base = (raw_base + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (base - raw_base);

  const int BN = g.BN;
  const int stages = g.stages;
  const uint32_t B_BYTES = static_cast<uint32_t>(BN) * 128u;
  const uint32_t sA = base;
  const uint32_t sB = base + stages * GEMM_A_BYTES;
  const uint32_t bar_base = sB + stages * B_BYTES;
  // barrier layout: full[stages], empty[stages], tfull[2], tempty[2], holder
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * stages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * stages + 2 + s); };
  const uint32_t holder = bar_base + 8u * (2 * stages + 4);
  volatile uint32_t* holder_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (holder - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (g.K + GEMM_BK - 1) / GEMM_BK;
  const int total_tiles = g.num_m_tiles * g.num_n_tiles;
  const uint32_t tmem_cols = 2u * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(holder, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder_gen;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------ TMA producer ------------------------------
      uint32_t s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile % g.num_m_tiles, n_blk = tile / g.num_m_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), GEMM_A_BYTES + B_BYTES);
          const uint32_t a_dst = sA + s * GEMM_A_BYTES;
          const uint32_t b_dst = sB + s * B_BYTES;
          if (!g.transA) {
            tma_load_2d(a_dst, &tmA, full_bar(s), kb * GEMM_BK, m_blk * GEMM_BM);
          } else {
            tma_load_2d(a_dst, &tmA, full_bar(s), m_blk * GEMM_BM, kb * GEMM_BK);
            tma_load_2d(a_dst + 8192, &tmA, full_bar(s), m_blk * GEMM_BM + 64, kb * GEMM_BK);
          }
          if (!g.transB) {
            tma_load_2d(b_dst, &tmB, full_bar(s), kb * GEMM_BK, n_blk * BN);
          } else {
            for (int gi = 0; gi < BN / 64; ++gi)
              tma_load_2d(b_dst + gi * 8192, &tmB, full_bar(s), n_blk * BN + gi * 64, kb * GEMM_BK);
          }
          if (++s == static_cast<uint32_t>(stages)) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------- MMA issuer -------------------------------
      const uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, g.transA != 0, g.transB != 0);
      uint32_t s = 0, ph = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = sA + s * GEMM_A_BYTES;
          const uint32_t b_addr = sB + s * B_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t adesc = g.transA ? make_sdesc(a_addr + k * 2048, 8192, 1024)
                                            : make_sdesc(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = g.transB ? make_sdesc(b_addr + k * 2048, 8192, 1024)
                                            : make_sdesc(b_addr + k * 32, 16, 1024);
            umma_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
          if (++s == static_cast<uint32_t>(stages)) {
            s = 0;
            ph ^= 1u;
          }
        }
        umma_commit(tfull_bar(as));
      }
    }
  } else if (warp >= 4) {
    // --------------------------------- epilogue ---------------------------------
    const int ew = warp & 3;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile % g.num_m_tiles, n_blk = tile / g.num_m_tiles;
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const int row = m_blk * GEMM_BM + ew * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(taddr + c0, v);
        tmem_wait_ld();
        if (row < g.M) gemm_epilogue_chunk(g.epi, row, n_blk * BN + c0, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

int g_force_bn = 0;
}  // namespace pk

// test hook: force the N tile (64/128/256); 0 restores the heuristic
extern "C" void pk_gemm_force_bn(int bn) { pk::g_force_bn = bn; }

extern "C" int pk_gemm_bf16(const void* A, const void* B, int M, int N, int K, int lda, int ldb,
                            int transA, int transB, const PkEpilogue* epi, void* stream) {
  using namespace pk;
  PK_CHECK(A && B && epi && epi->out, "pk_gemm_bf16: null pointer");
  PK_CHECK(M > 0 && N > 0 && K > 0, "pk_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
  PK_CHECK(N % 64 == 0, "pk_gemm_bf16: N=%d must be a multiple of 64", N);
  PK_CHECK(lda % 8 == 0 && ldb % 8 == 0, "pk_gemm_bf16: lda/ldb must be multiples of 8 (16 B)");
  PK_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
           "pk_gemm_bf16: operands must be 16-byte aligned");
  PK_CHECK(epi->kind >= PK_EPI_BF16 && epi->kind <= PK_EPI_PIXSHUF, "pk_gemm_bf16: bad epilogue %d",
           epi->kind);
  if (epi->kind != PK_EPI_PIXSHUF)
    PK_CHECK(epi->ldc % 8 == 0 && epi->ldc >= N, "pk_gemm_bf16: bad ldc %d", epi->ldc);
  if (epi->kind == PK_EPI_RESID || epi->kind == PK_EPI_DGELU)
    PK_CHECK(epi->aux && epi->ld_aux % 8 == 0, "pk_gemm_bf16: epilogue %d needs aux", epi->kind);
  if (epi->kind == PK_EPI_GELU) PK_CHECK(epi->out2, "pk_gemm_bf16: GELU epilogue needs out2");
  if (epi->kind == PK_EPI_RESID && epi->rowscale)
    PK_CHECK(epi->rows_per_group > 0, "pk_gemm_bf16: rows_per_group must be > 0");
  if (epi->kind == PK_EPI_PIXSHUF)
    PK_CHECK(epi->ps_c % 32 == 0 && epi->ps_h * epi->ps_w > 0 && M % (epi->ps_h * epi->ps_w) == 0 &&
                 N == epi->ps_p * epi->ps_p * epi->ps_c,
             "pk_gemm_bf16: bad pixel-shuffle geometry");

  GemmArgs g;
  g.M = M;
  g.N = N;
  g.K = K;
  g.transA = transA ? 1 : 0;
  g.transB = transB ? 1 : 0;
  g.epi = *epi;
  if (g.epi.alpha == 0.0f) g.epi.alpha = 1.0f;
  const int sms = sm_count();
  const int mt = (M + GEMM_BM - 1) / GEMM_BM;
  int BN = 64;
  if (N % 256 == 0 && static_cast<long long>(mt) * (N / 256) >= sms) BN = 256;
  else if (N % 128 == 0 && static_cast<long long>(mt) * (N / 128) >= sms / 2) BN = 128;
  if (g_force_bn > 0 && N % g_force_bn == 0) BN = g_force_bn;
  g.BN = BN;
  g.stages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  g.num_m_tiles = mt;
  g.num_n_tiles = N / BN;

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (!g.transA) {
      dims[0] = K; dims[1] = M; box[0] = 64; box[1] = 128;
    } else {
      dims[0] = M; dims[1] = K; box[0] = 64; box[1] = 64;
    }
    strides[0] = static_cast<uint64_t>(lda) * 2;
    if (!make_tmap_bf16(&tmA, A, 2, dims, strides, box)) return 3;
    if (!g.transB) {
      dims[0] = K; dims[1] = N; box[0] = 64; box[1] = static_cast<uint32_t>(BN);
    } else {
      dims[0] = N; dims[1] = K; box[0] = 64; box[1] = 64;
    }
    strides[0] = static_cast<uint64_t>(ldb) * 2;
    if (!make_tmap_bf16(&tmB, B, 2, dims, strides, box)) return 3;
  }
  const size_t smem = 1024 + static_cast<size_t>(g.stages) * (GEMM_A_BYTES + BN * 128) + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    PK_CHECK(e == cudaSuccess, "pk_gemm_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int total = g.num_m_tiles * g.num_n_tiles;
  const int grid = total < sms ? total : sms;
  gemm_bf16_kernel<<<grid, GEMM_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, g);
  PK_LAUNCH_CHECK("pk_gemm_bf16");
  return 0;
}
