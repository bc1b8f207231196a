// painter_b200 — 2-CTA (cta_group::2) variant of the persistent tcgen05 GEMM.
//
// A CTA pair (cluster of 2 on one TPC) owns a 256 x BN output tile: CTA r stages A rows [128 r, 128 r + 128) and
// B rows [BN/2 r, BN/2 r + BN/2) of the tile in its own shared memory; the leader CTA issues one
// tcgen05.mma.cta_group::2 (M = 256) per K = 16 slice that reads both CTAs' operands and writes 128 accumulator
// rows into each CTA's TMEM.  Versus the 1-CTA kernel (gemm.cu) every SM stages half of B, which takes the
// kernel off the shared-memory bandwidth limit (the 128 x 256 1-CTA tile needs 192 B/clk of smem traffic against
// 128 B/clk available) — same epilogues, same operand-major options, same split-K.
//
// Synchronisation: per-stage "full" mbarriers live in the leader (both CTAs' TMA loads complete_tx on it through
// the .cta_group::2 TMA form); "empty" / "accumulator full" barriers exist in both CTAs and are signalled by
// multicast tcgen05.commit; the peer's epilogue warps release an accumulator stage with a remote mbarrier arrive.
#include "gemm_common.cuh"

namespace pk {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0,
                                             int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t holder_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread retired) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

constexpr int G2_STAGES_256 = 6;  // 6 x (16 KiB A + 16 KiB B-half) per CTA

template <int KIND>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_base = smem_u32(smem_raw);
  const uint32_t base = (raw_base + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (base - raw_base);

  const int BN = g.BN;          // tile width of the PAIR
  const int BNh = BN / 2;       // B rows staged by one CTA
  const int stages = g.stages;
  const uint32_t B_BYTES = static_cast<uint32_t>(BNh) * 128u;
  const uint32_t sA = base;
  const uint32_t sB = base + stages * GEMM_A_BYTES;
  const uint32_t bar_base = sB + stages * B_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * stages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * stages + 2 + s); };
  const uint32_t holder = bar_base + 8u * (2 * stages + 4);
  volatile uint32_t* holder_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (holder - base));
  float* stg_gen = reinterpret_cast<float*>(smem_gen + (bar_base - base) + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_kb = (g.K + GEMM_BK - 1) / GEMM_BK;
  const int tiles_mn = g.num_m_tiles * g.num_n_tiles;   // m tiles are 256 rows here
  const uint32_t tmem_cols = 2u * BN;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 9 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * GEMM_EPI_WARPS);  // epilogue warps of both CTAs (only the leader's copy is waited on)
    }
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc2(holder, tmem_cols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *holder_gen;
  // the prologue above overlaps the tail of the previous kernel (programmatic dependent launch, host_common.h)
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 8) {
    if (lane == 0) {
      // ------------------------------ TMA producer (both CTAs) ------------------------------
      uint32_t s = 0, ph = 0;
      GemmSched sch;
      sch.init(g, tiles_mn, num_kb, cid, ncl);
      int mn, kb0, kb1;
      while (sch.next(mn, kb0, kb1)) {
        int m_blk, n_blk;
        gemm_tile_coords(g, mn, m_blk, n_blk);
        const int m0 = m_blk * 256 + static_cast<int>(rank) * 128;
        const int n0 = n_blk * BN + static_cast<int>(rank) * BNh;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (leader) mbar_expect_tx(full_bar(s), 2u * (GEMM_A_BYTES + B_BYTES));
          const uint32_t fb = mapa_shared(full_bar(s), 0);
          const uint32_t a_dst = sA + s * GEMM_A_BYTES;
          const uint32_t b_dst = sB + s * B_BYTES;
          if (!g.transA) {
            tma2_load_2d(a_dst, &tmA, fb, kb * GEMM_BK, m0);
          } else {
            tma2_load_2d(a_dst, &tmA, fb, m0, kb * GEMM_BK);
            tma2_load_2d(a_dst + 8192, &tmA, fb, m0 + 64, kb * GEMM_BK);
          }
          if (!g.transB) {
            tma2_load_2d(b_dst, &tmB, fb, kb * GEMM_BK, n0);
          } else {
            for (int gi = 0; gi < BNh / 64; ++gi)
              tma2_load_2d(b_dst + gi * 8192, &tmB, fb, n0 + gi * 64, kb * GEMM_BK);
          }
          if (++s == static_cast<uint32_t>(stages)) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 9) {
    if (leader) {
      // ------------------------------- MMA issuer (leader CTA) -------------------------------
      const uint32_t idesc = make_idesc_bf16(256, BN, g.transA != 0, g.transB != 0);
      uint32_t s = 0, ph = 0, it = 0;
      GemmSched sch;
      sch.init(g, tiles_mn, num_kb, cid, ncl);
      int mn, kb0, kb1;
      for (; sch.next(mn, kb0, kb1); ++it) {
        const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = sA + s * GEMM_A_BYTES;
          const uint32_t b_addr = sB + s * B_BYTES;
          const uint64_t a0 = g.transA ? make_sdesc(a_addr, 8192, 1024) : make_sdesc(a_addr, 16, 1024);
          const uint64_t b0 = g.transB ? make_sdesc(b_addr, 8192, 1024) : make_sdesc(b_addr, 16, 1024);
          const uint32_t a_step = g.transA ? 2048u : 32u, b_step = g.transB ? 2048u : 32u;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              umma2_ss(d_tmem, sdesc_add(a0, k * a_step), sdesc_add(b0, k * b_step), idesc,
                       (kb != kb0 || k != 0) ? 1u : 0u);
            umma2_commit_mc(empty_bar(s));
          }
          __syncwarp();
          if (++s == static_cast<uint32_t>(stages)) {
            s = 0;
            ph ^= 1u;
          }
        }
        if (elect_one()) umma2_commit_mc(tfull_bar(as));
        __syncwarp();
      }
    }
  } else if (warp < GEMM_EPI_WARPS) {
    // --------------------------------- epilogue (both CTAs) ---------------------------------
    const int ew = warp & 3;   // TMEM lane quarter
    const int eg = warp >> 2;  // column group: this warp owns the 64-column blocks with (c0 / 64) % 2 == eg
    uint32_t it = 0;
    float* stg = stg_gen + warp * (32 * STG_LD);
    // per-warp copy of the bias values of the (up to four) 32-column chunks this warp handles in a tile: fetched with
    // one coalesced load per lane BEFORE the accumulator is waited for, read back as shared-memory broadcasts
    float* bias_s = stg_gen + GEMM_EPI_WARPS * (32 * STG_LD) + warp * 128;
    // The accumulator chunk of step i+1 is loaded from TMEM while chunk i is processed where the register budget
    // allows it (the fp32-aux epilogues hold two prefetched aux chunks instead; pipelining those too spills and
    // measured 3 % slower, scripts/time_gemms.py).
    constexpr bool PIPE = (KIND == PK_EPI_BF16 || KIND == PK_EPI_GELU || KIND == PK_EPI_PIXSHUF);
    GemmSched sch;
    sch.init(g, tiles_mn, num_kb, cid, ncl);
    int mn, kb0, kb1;
    for (; sch.next(mn, kb0, kb1); ++it) {
      int m_blk, n_blk;
      gemm_tile_coords(g, mn, m_blk, n_blk);
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      const bool has_bias = g.epi.bias != nullptr;
      float4 bias_v = make_float4(0.f, 0.f, 0.f, 0.f);
      {
        // lane l: chunk (l >> 3) of this warp's chunk list {eg*64, eg*64+32, eg*64+128, eg*64+160}, 4 floats
        const int ch = lane >> 3;
        const int ccol = eg * 64 + (ch & 1) * 32 + (ch >> 1) * 128;
        if (has_bias && ccol < BN)
          bias_v = __ldg(reinterpret_cast<const float4*>(g.epi.bias + n_blk * BN + ccol + (lane & 7) * 4));
      }
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      if (has_bias) {
        *reinterpret_cast<float4*>(bias_s + lane * 4) = bias_v;
        __syncwarp();
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
      const int row0 = m_blk * 256 + static_cast<int>(rank) * 128 + ew * 32;
      EpiAux auxA, auxB;   // explicit ping-pong (BN is a multiple of 64): keeps both in registers
      if (eg * 64 < BN) gemm_epilogue_prefetch<KIND>(g.epi, row0, g.M, n_blk * BN + eg * 64, auxA);
      if constexpr (PIPE) {
        uint32_t vA[32], vB[32];
        if (eg * 64 < BN) tmem_ld_x32(taddr + eg * 64, vA);
        int k = 0;
        for (int c0 = eg * 64; c0 < BN; c0 += 128, ++k) {
          tmem_wait_ld();
          tmem_ld_x32(taddr + c0 + 32, vB);
          gemm_epilogue_prefetch<KIND>(g.epi, row0, g.M, n_blk * BN + c0 + 32, auxB);
          gemm_epilogue_chunk<KIND>(g.epi, stg, row0, g.M, n_blk * BN + c0, vA, auxA,
                                    has_bias ? bias_s + (2 * k) * 32 : nullptr);
          tmem_wait_ld();
          if (c0 + 128 < BN) {
            tmem_ld_x32(taddr + c0 + 128, vA);
            gemm_epilogue_prefetch<KIND>(g.epi, row0, g.M, n_blk * BN + c0 + 128, auxA);
          }
          gemm_epilogue_chunk<KIND>(g.epi, stg, row0, g.M, n_blk * BN + c0 + 32, vB, auxB,
                                    has_bias ? bias_s + (2 * k + 1) * 32 : nullptr);
        }
      } else {
        int k = 0;
        for (int c0 = eg * 64; c0 < BN; c0 += 128, ++k) {
          uint32_t v[32];
          tmem_ld_x32(taddr + c0, v);
          gemm_epilogue_prefetch<KIND>(g.epi, row0, g.M, n_blk * BN + c0 + 32, auxB);
          tmem_wait_ld();
          gemm_epilogue_chunk<KIND>(g.epi, stg, row0, g.M, n_blk * BN + c0, v, auxA,
                                    has_bias ? bias_s + (2 * k) * 32 : nullptr);
          tmem_ld_x32(taddr + c0 + 32, v);
          if (c0 + 128 < BN) gemm_epilogue_prefetch<KIND>(g.epi, row0, g.M, n_blk * BN + c0 + 128, auxA);
          tmem_wait_ld();
          gemm_epilogue_chunk<KIND>(g.epi, stg, row0, g.M, n_blk * BN + c0 + 32, v, auxB,
                                    has_bias ? bias_s + (2 * k + 1) * 32 : nullptr);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(as), 0));
    }
  }

  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 10) tmem_dealloc2(tmem_base, tmem_cols);
}

// Launcher used by pk_gemm_bf16 (gemm.cu).  g.BN / g.stages / g.num_*_tiles are already in pair units.
int launch_gemm2(const void* A, const void* B, int lda, int ldb, GemmArgs& g, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  uint64_t dims[2], strides[1];
  uint32_t box[2];
  if (!g.transA) {
    dims[0] = g.K; dims[1] = g.M; box[0] = 64; box[1] = 128;
  } else {
    dims[0] = g.M; dims[1] = g.K; box[0] = 64; box[1] = 64;
  }
  strides[0] = static_cast<uint64_t>(lda) * 2;
  if (!make_tmap_bf16(&tmA, A, 2, dims, strides, box)) return 3;
  if (!g.transB) {
    dims[0] = g.K; dims[1] = g.N; box[0] = 64; box[1] = static_cast<uint32_t>(g.BN / 2);
  } else {
    dims[0] = g.N; dims[1] = g.K; box[0] = 64; box[1] = 64;
  }
  strides[0] = static_cast<uint64_t>(ldb) * 2;
  if (!make_tmap_bf16(&tmB, B, 2, dims, strides, box)) return 3;

  const size_t smem = 1024 + static_cast<size_t>(g.stages) * (GEMM_A_BYTES + (g.BN / 2) * 128) + 256 +
                      GEMM_EPI_WARPS * 32 * STG_LD * sizeof(float) + GEMM_EPI_WARPS * 128 * sizeof(float);
  const int total = g.num_m_tiles * g.num_n_tiles * g.splits;
  int clusters = sm_count() / 2;
  if (g.streamk_units > 0) {
    const int units = g.num_m_tiles * g.num_n_tiles * ((g.K + GEMM_BK - 1) / GEMM_BK);
    clusters = (units + g.streamk_units - 1) / g.streamk_units;
  } else if (clusters > total) {
    clusters = total;
  }
  PdlLaunch L(dim3(2 * clusters), dim3(GEMM_THREADS), smem, st, 2);
  cudaLaunchConfig_t& cfg = L.cfg;
#define PK_GEMM2_CASE(KK)                                                                                       \
  case KK: {                                                                                                    \
    static bool attr_set = false;                                                                               \
    if (!attr_set) {                                                                                            \
      cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           227 * 1024);                                                         \
      PK_CHECK(e == cudaSuccess, "pk_gemm_bf16(2cta): cudaFuncSetAttribute: %s", cudaGetErrorString(e));        \
      attr_set = true;                                                                                          \
    }                                                                                                           \
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm2_bf16_kernel<KK>, tmA, tmB, g);                               \
    PK_CHECK(e == cudaSuccess, "pk_gemm_bf16(2cta): launch: %s", cudaGetErrorString(e));                        \
  } break;
  switch (g.epi.kind) {
    PK_GEMM2_CASE(PK_EPI_BF16)
    PK_GEMM2_CASE(PK_EPI_F32)
    PK_GEMM2_CASE(PK_EPI_GELU)
    PK_GEMM2_CASE(PK_EPI_RESID)
    PK_GEMM2_CASE(PK_EPI_DGELU)
    PK_GEMM2_CASE(PK_EPI_PIXSHUF)
    default:
      PK_CHECK(false, "pk_gemm_bf16(2cta): bad epilogue kind %d", g.epi.kind);
  }
#undef PK_GEMM2_CASE
  PK_LAUNCH_CHECK("pk_gemm_bf16(2cta)");
  return 0;
}

}  // namespace pk
