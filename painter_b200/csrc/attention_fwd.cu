// painter_b200 — fused global attention forward with decomposed relative-position bias (sm_100a).
//
// Replaces models_painter.py:73-86 + util/vitdet_utils.py:63-125 (q*scale @ k^T, add_decomposed_rel_pos,
// softmax, @ v) without ever materialising the [B*heads, N, N] score tensor.
//
// One CTA = 128 query tokens of one (batch, head); two CTAs co-reside per SM so one CTA's softmax
// overlaps the other's tensor-core work.  6 warps:
//   warp 4       TMA producer: Q tile, bf16 rel-pos tables, then K / V tiles (single-stage each)
//   warp 5       MMA issuer (tcgen05, accumulators in TMEM): S = Q.K^T, O += P.V
//                (highest warp ids: the SM's issue arbiter prefers higher warp ids, and these two warps sit on
//                 the critical path while the softmax warps saturate the issue slots)
//   warps 0..3   softmax: one thread per query row (tcgen05.ld 32x32b), online softmax in the log2 domain
//                with lazy rescaling; P stays in TENSOR MEMORY: the row owners pack it as bf16 pairs into 56 TMEM
//                columns (tcgen05.st) and O += P.V is issued in the TS form (A operand from TMEM, only V is read
//                from shared memory) - an SS-form MMA re-reads 4 KB of A + 2 KB of B per K = 16 slice through the
//                128 B/clk shared-memory port, which is what bounds the d = 64 MMAs (scripts/ubench/mma_ts.cu);
//                PK_ATTN_FWD_TS=0 builds the round-1 variant (P through a 128B-swizzled smem tile)
// Key tiles are KT = 112 keys = R whole image rows (R = 112 / W) so that, inside a tile, the key's image
// row / column are compile-time: rel_w lives in registers, rel_h needs R values per tile.
// The bias itself comes from the tensor cores too: G_h = Q . T_h^T and G_w = Q . T_w^T (T = bf16 rel-pos
// tables) are two extra MMAs at CTA start; rel_h[r, i] = G_h[r, i_r - i + h - 1] (Toeplitz gather).
#include "common.cuh"
#include "host_common.h"
#include "../../include/painter_b200.h"

namespace pk {

constexpr int ATT_BM = 128;
constexpr int ATT_KT = 112;
constexpr int ATT_THREADS = 192;
constexpr float LOG2E = 1.4426950408889634f;
// PK_ATTN_FWD_TS: 0 = P through shared memory (SS form); 1 (default) = P in TMEM, S(j+1) issued after P(j).V(j);
// 2 = P in TMEM and S(j+1) issued BEFORE P(j).V(j), so that the next tile's scores do not wait behind the seven P.V
// MMAs (~390 cycles of a CTA's ~2650-cycle tile chain, profiles/r02_attn_cta_timelines.txt): P(j+1) then starts being
// written while P(j).V(j) may still read P(j), so the 16-key P chunks rotate through ten 8-column TMEM slots (tile j's
// chunk c lives in slot (7 j + c) mod 10: the first three chunks of a tile land in free slots, the fourth waits for
// the previous tile's P.V to retire).  Correct (same tests green) but SLOWER on the B200 at the B = 8 geometry:
// 0.171 ms against 0.161 ms - the two co-resident CTAs' softmax phases then overlap and a tile's softmax pass
// stretches from ~1370 to ~1650 cycles: the SM's softmax throughput, not the per-CTA chain, is the bound.
#ifndef PK_ATTN_FWD_TS
#define PK_ATTN_FWD_TS 1
#endif

// shared-memory map (offsets from the 1024-aligned base)
constexpr uint32_t ATT_SQ = 0;                 // 128 x 128 B
constexpr uint32_t ATT_SK = 16384;             // 112 x 128 B
constexpr uint32_t ATT_SV = ATT_SK + 14336;    // 112 x 128 B   (T_w staged here before the main loop)
constexpr uint32_t ATT_SP = ATT_SV + 14336;    // 2 x (128 x 128 B) (T_h staged here before the main loop)
constexpr uint32_t ATT_SRELH = ATT_SP + 32768; // 128 x (h+1) fp32  (>= 4 x 32 x 17 fp32 scratch)

struct AttnFwdArgs {
  int h, N, heads;
  int th_pad, tw_pad;     // padded row counts of the bf16 tables (multiples of 16)
  int relh_bytes;         // size of the sRelh region
  float scale_log2;       // head_dim^-0.5 * log2(e)
  __nv_bfloat16* out;     // [B*N, ldo]
  int ldo;
  float* lse;             // [B*heads, N]  (log2 domain: m + log2(l))
  // optional (training): the log2e-scaled bias rows rel_h / rel_w of every query, kept for the backward kernels, which
  // then start from two coalesced loads instead of recomputing G_h / G_w on the tensor cores and re-doing the Toeplitz
  // gathers (that prologue was 10 k of a dQ CTA's 48 k cycles).  Layout: per 128-query tile, query row innermost
  // (attention_bwd.cu AttnBwdArgs): relh [B*heads, q tiles, h, 128], relw [B*heads, q tiles, W/4, 128] float4.
  float* relh_out;
  float* relw_out;
  long long* trace;       // optional debug timeline (CTA 0 only): [role][tile][event] clock64 stamps
};

// debug timeline: role 0 = producer, 1 = MMA issuer, 2 = softmax row 0; 8 events per tile
#define ATT_TRACE(role, tile, ev)                                                                     \
  do {                                                                                                \
    if (a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tile) < 16 &&   \
        ((role) == 2 || (threadIdx.x & 31) == 0))                                                       \
      a.trace[((role) * 16 + (tile)) * 8 + (ev)] = clock64();                                          \
  } while (0)

template <int W>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                const __grid_constant__ CUtensorMap tmTh, const __grid_constant__ CUtensorMap tmTw,
                const AttnFwdArgs a) {
  constexpr int R = ATT_KT / W;
  static_assert(R * W == ATT_KT, "W must divide 112");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_base = smem_u32(smem_raw);
  const uint32_t base = (raw_base + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw_base);

  const uint32_t sQ = base + ATT_SQ, sK = base + ATT_SK, sV = base + ATT_SV, sP = base + ATT_SP;
  float* relh_gen = reinterpret_cast<float*>(gen + ATT_SRELH);
  const uint32_t bar0 = base + ATT_SRELH + a.relh_bytes;
  const uint32_t bar_q = bar0, bar_kf = bar0 + 8, bar_ke = bar0 + 16, bar_vf = bar0 + 24,
                 bar_ve = bar0 + 32, bar_s = bar0 + 40, bar_p = bar0 + 48, bar_o = bar0 + 56,
                 bar_g = bar0 + 64, bar_gr = bar0 + 72,
                 bar_gw = bar0 + 88;  // G_w retired (single completion: the producer must not alias bar_g's phases)
  const uint32_t holder = bar0 + 80;
  volatile uint32_t* holder_gen = reinterpret_cast<volatile uint32_t*>(gen + ATT_SRELH + a.relh_bytes + 80);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM;
  const int head = blockIdx.y, b = blockIdx.z;
  const int C = a.heads * 64;
  const int num_tiles = (a.h + R - 1) / R;
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmTh);
    tma_prefetch_desc(&tmTw);
    mbar_init(bar_q, 1);
    mbar_init(bar_kf, 1);
    mbar_init(bar_ke, 1);
    mbar_init(bar_vf, 1);
    mbar_init(bar_ve, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 1);
    mbar_init(bar_g, 1);
    mbar_init(bar_gr, 128);
    mbar_init(bar_gw, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(holder, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder_gen;
  // TMEM columns (256 per CTA, two CTAs per SM): S [0,112) | O [112,176) | P as packed bf16 pairs: ten 8-column
  // slots [176,256) (PK_ATTN_FWD_TS=1 uses the first seven in place)
  const uint32_t tS = tmem, tO = tmem + 112;
#if PK_ATTN_FWD_TS
  const uint32_t tP = tmem + 176;
#endif
  pdl_launch_dependents();   // programmatic dependent launch (host_common.h): the prologue above overlaps the
  pdl_wait();                // previous kernel's tail; nothing before this line touches global memory

  if (warp == 4) {
    if (lane == 0) {
      // ------------------------------------ TMA producer ------------------------------------
      mbar_expect_tx(bar_q, 16384u + static_cast<uint32_t>(a.th_pad + a.tw_pad) * 128u);
      tma_load_3d(sQ, &tmQ, bar_q, head * 64, q0, b);
      tma_load_2d(sP, &tmTh, bar_q, 0, 0);
      tma_load_2d(sV, &tmTw, bar_q, 0, 0);
      mbar_expect_tx(bar_kf, ATT_KT * 128);
      tma_load_3d(sK, &tmKV, bar_kf, C + head * 64, 0, b);
      mbar_wait(bar_gw, 0);  // G_w MMA finished reading T_w out of the V buffer
      mbar_expect_tx(bar_vf, ATT_KT * 128);
      tma_load_3d(sV, &tmKV, bar_vf, 2 * C + head * 64, 0, b);
      for (int j = 1; j < num_tiles; ++j) {
        mbar_wait(bar_ke, (j - 1) & 1);
        ATT_TRACE(0, j, 0);
        mbar_expect_tx(bar_kf, ATT_KT * 128);
        tma_load_3d(sK, &tmKV, bar_kf, C + head * 64, j * ATT_KT, b);
        mbar_wait(bar_ve, (j - 1) & 1);
        ATT_TRACE(0, j, 1);
        mbar_expect_tx(bar_vf, ATT_KT * 128);
        tma_load_3d(sV, &tmKV, bar_vf, 2 * C + head * 64, j * ATT_KT, b);
      }
    }
  } else if (warp == 5) {
    {
      // -------------------------------------- MMA issuer --------------------------------------
      // The whole warp runs the (uniform) control flow; only the tcgen05 instructions are issued by one elected lane.
      mbar_wait(bar_q, 0);
      tc_fence_after();
      {  // G_w = Q . T_w^T
        const uint32_t idesc = make_idesc_bf16(128, a.tw_pad, false, false);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tS, make_sdesc(sQ + k * 32, 16, 1024), make_sdesc(sV + k * 32, 16, 1024), idesc, k != 0);
          umma_commit(bar_g);
          umma_commit(bar_gw);
        }
        __syncwarp();
      }
      mbar_wait(bar_gr, 0);
      tc_fence_after();
      {  // G_h = Q . T_h^T
        const uint32_t idesc = make_idesc_bf16(128, a.th_pad, false, false);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tS, make_sdesc(sQ + k * 32, 16, 1024), make_sdesc(sP + k * 32, 16, 1024), idesc, k != 0);
          umma_commit(bar_g);
        }
        __syncwarp();
      }
      mbar_wait(bar_gr, 1);
      tc_fence_after();
      const uint32_t idesc_qk = make_idesc_bf16(128, ATT_KT, false, false);
      const uint32_t idesc_pv = make_idesc_bf16(128, 64, false, true);
      // single-stage operand buffers: every descriptor is loop-invariant
      uint64_t dq[4], dk[4], dp[ATT_KT / 16], dv[ATT_KT / 16];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dq[k] = make_sdesc(sQ + k * 32, 16, 1024);
        dk[k] = make_sdesc(sK + k * 32, 16, 1024);
      }
#pragma unroll
      for (int kk = 0; kk < ATT_KT / 16; ++kk) {
        dp[kk] = make_sdesc(sP + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
        dv[kk] = make_sdesc(sV + kk * 2048, 16, 1024);
      }
#if PK_ATTN_FWD_TS == 2
      // S(0), then per tile: [P(j) ready] -> S(j+1) -> P(j).V(j)
      mbar_wait(bar_kf, 0);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tS, dq[k], dk[k], idesc_qk, k != 0);
        umma_commit(bar_ke);
        umma_commit(bar_s);
      }
      __syncwarp();
      for (int j = 0; j < num_tiles; ++j) {
        ATT_TRACE(1, j, 0);
        mbar_wait(bar_p, j & 1);      // softmax(j) done: S is free, P(j) sits in its TMEM slots
        ATT_TRACE(1, j, 4);
        if (j + 1 < num_tiles) {
          mbar_wait(bar_kf, (j + 1) & 1);
          ATT_TRACE(1, j, 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tS, dq[k], dk[k], idesc_qk, k != 0);
            umma_commit(bar_ke);
            umma_commit(bar_s);
          }
          __syncwarp();
          ATT_TRACE(1, j, 2);
        }
        mbar_wait(bar_vf, j & 1);
        ATT_TRACE(1, j, 3);
        tc_fence_after();
        const int slot0 = (7 * j) % 10;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < ATT_KT / 16; ++kk) {
            int sl = slot0 + kk;
            if (sl >= 10) sl -= 10;
            umma_ts(tO, tP + sl * 8, dv[kk], idesc_pv, (j | kk) != 0);
          }
          umma_commit(bar_ve);
          if (j == num_tiles - 1) umma_commit(bar_o);
        }
        __syncwarp();
      }
#else
      for (int j = 0; j < num_tiles; ++j) {
        ATT_TRACE(1, j, 0);
        mbar_wait(bar_kf, j & 1);
        ATT_TRACE(1, j, 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tS, dq[k], dk[k], idesc_qk, k != 0);
          umma_commit(bar_ke);
          umma_commit(bar_s);
        }
        __syncwarp();
        ATT_TRACE(1, j, 2);
        mbar_wait(bar_vf, j & 1);   // completes long before P is ready
        ATT_TRACE(1, j, 3);
        mbar_wait(bar_p, j & 1);
        ATT_TRACE(1, j, 4);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < ATT_KT / 16; ++kk) {
#if PK_ATTN_FWD_TS
            umma_ts(tO, tP + kk * 8, dv[kk], idesc_pv, (j | kk) != 0);
#else
            umma_ss(tO, dp[kk], dv[kk], idesc_pv, (j | kk) != 0);
#endif
          }
          umma_commit(bar_ve);
          if (j == num_tiles - 1) umma_commit(bar_o);
        }
        __syncwarp();
      }
#endif
    }
  } else {
    // ------------------------------------ softmax warps ------------------------------------
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    int t = q0 + row;
    const bool valid = t < a.N;
    if (!valid) t = a.N - 1;
    const int i_r = t / W, j_r = t - i_r * W;
    const int h = a.h;
    const int ldr = h + 1;
    float* my_relh = relh_gen + static_cast<size_t>(row) * ldr;

    // ---- rel_w -> registers (through a private smem scratch row for the dynamic shift) ----
    float relw[W];
    {
      float* scratch = relh_gen + (static_cast<size_t>(quarter) * 32 + lane) * 17;
      mbar_wait(bar_g, 0);
      tc_fence_after();
      for (int c0 = 0; c0 < a.tw_pad; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(tS + lane_addr + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 16; ++c) scratch[c] = __uint_as_float(v[c]) * LOG2E;
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const int tt = j_r + (W - 1) - j - c0;
          if (tt >= 0 && tt < 16) relw[j] = scratch[tt];
        }
      }
      tc_fence_before();
      __syncwarp();
      mbar_arrive(bar_gr);
      if (a.relw_out != nullptr) {
        float* dst = a.relw_out + ((static_cast<size_t>(b) * a.heads + head) * gridDim.x + blockIdx.x) *
                                      static_cast<size_t>(W) * ATT_BM;
        if constexpr (W % 4 == 0) {
#pragma unroll
          for (int j = 0; j < W / 4; ++j)
            reinterpret_cast<float4*>(dst)[j * ATT_BM + row] =
                make_float4(relw[4 * j], relw[4 * j + 1], relw[4 * j + 2], relw[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < W; ++j) dst[j * ATT_BM + row] = relw[j];
        }
      }
    }
    // ---- rel_h -> smem [row][h]  (Toeplitz gather: rel_h[i] = G_h[i_r - i + h - 1]) ----
    {
      mbar_wait(bar_g, 1);  // also implies every thread is done with the scratch rows (bar_gr phase 0)
      tc_fence_after();
      for (int c0 = 0; c0 < a.th_pad; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(tS + lane_addr + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int i = i_r + (h - 1) - (c0 + c);
          if (i >= 0 && i < h) my_relh[i] = __uint_as_float(v[c]) * LOG2E;
        }
      }
      tc_fence_before();
      __syncwarp();
      mbar_arrive(bar_gr);
      if (a.relh_out != nullptr) {
        // rows past N carry the (finite) values of the clamped token: the backward masks them but reads them
        float* dst = a.relh_out + ((static_cast<size_t>(b) * a.heads + head) * gridDim.x + blockIdx.x) *
                                      static_cast<size_t>(h) * ATT_BM + row;
        for (int i = 0; i < h; ++i) dst[static_cast<size_t>(i) * ATT_BM] = my_relh[i];
      }
    }

    float m_ref = -INFINITY, l_sum = 0.f;
    const float sc = a.scale_log2;
    for (int j = 0; j < num_tiles; ++j) {
      float hb[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int i = j * R + r;
        hb[r] = i < h ? my_relh[i] : -INFINITY;   // image rows past the end: bias -inf -> p = 0, no per-score select
      }
#if PK_ATTN_FWD_TS == 2
      const int pslot0 = (7 * j) % 10;
#endif
      if (row == 0) ATT_TRACE(2, j, 0);
      mbar_wait(bar_s, j & 1);
      if (row == 0) ATT_TRACE(2, j, 1);
      tc_fence_after();
      if (j == 0) {
        // first tile: an explicit max pass fixes the reference point m_ref
        float mx = -INFINITY;
#pragma unroll
        for (int c0 = 0; c0 < ATT_KT; c0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(tS + lane_addr + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int kc = c0 + c;
            const float tv = fmaf(__uint_as_float(v[c]), sc, hb[kc / W] + relw[kc % W]);
            mx = fmaxf(mx, tv);
          }
        }
        m_ref = mx;
      }
      // Optimistic single pass: p = exp2(t - m_ref) with the reference point of the previous tiles while the
      // tile max is tracked; only if some row's max outgrew m_ref by more than 2^8 is O rescaled and the pass
      // repeated (rare after the first tiles).  The result is exact for any threshold.
      // The per-score arithmetic runs on packed fp32 pairs (FFMA2 / FADD2) when W is even: the softmax warps are
      // issue-bound, a pair of adjacent keys shares its image row, so bias add, scale-fma and the row-sum
      // accumulation cost one issue slot per two scores.
      constexpr bool PK2 = (W % 2 == 0);
      for (int attempt = 0; attempt < 2; ++attempt) {
        float hbm[R];
        f32x2 hbm2[R], relw2[PK2 ? W / 2 : 1];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          hbm[r] = hb[r] - m_ref;
          hbm2[r] = pack_f2(hbm[r], hbm[r]);
        }
        if constexpr (PK2) {
#pragma unroll
          for (int jj = 0; jj < W / 2; ++jj) relw2[jj] = pack_f2(relw[2 * jj], relw[2 * jj + 1]);
        }
        const f32x2 sc2 = pack_f2(sc, sc);
        f32x2 l2 = pack_f2(0.f, 0.f);
        float mx = -INFINITY, l_tile = 0.f;
        uint32_t v[2][16];
        tmem_ld_x16(tS + lane_addr, v[0]);
#pragma unroll
        for (int ci = 0; ci < ATT_KT / 16; ++ci) {
          const int c0 = ci * 16;
          tmem_wait_ld();
          if (ci + 1 < ATT_KT / 16) tmem_ld_x16(tS + lane_addr + c0 + 16, v[(ci + 1) & 1]);
          float p[16];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const int kc = c0 + c;
            float t0, t1;
            if constexpr (PK2) {
              unpack_f2(fma_f2(pack_u2(v[ci & 1][c], v[ci & 1][c + 1]), sc2,
                               add_f2(hbm2[kc / W], relw2[(kc % W) / 2])), t0, t1);
            } else {
              const int k1 = kc + 1;
              t0 = fmaf(__uint_as_float(v[ci & 1][c]), sc, hbm[kc / W] + relw[kc % W]);
              t1 = fmaf(__uint_as_float(v[ci & 1][c + 1]), sc, hbm[k1 / W] + relw[k1 % W]);
            }
            mx = fmaxf(mx, fmaxf(t0, t1));
            if constexpr (PK2) {
              exp2_pair<PK_EXP_POLY_MASK_FWD>(c >> 1, t0, t1, p[c], p[c + 1]);
            } else {
              p[c] = fast_exp2(t0);
              p[c + 1] = fast_exp2(t1);
            }
            if constexpr (PK2) l2 = add_f2(l2, pack_f2(p[c], p[c + 1]));
            else l_tile += p[c] + p[c + 1];
          }
#if PK_ATTN_FWD_TS
          {
            // 16 consecutive keys = 8 packed columns of this row's A operand in TMEM (key 2c low half, 2c+1 high)
            uint32_t pk8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) pk8[q] = pack_bf16x2(p[2 * q], p[2 * q + 1]);
#if PK_ATTN_FWD_TS == 2
            // chunk ci of tile j -> slot (7 j + ci) mod 10; from the fourth chunk on the slot was last read by the
            // previous tile's P.V MMAs, which must have retired (the first three slots are free by construction)
            if (ci == 3 && j >= 1) {
              mbar_wait(bar_ve, (j - 1) & 1);
              tc_fence_after();
            }
            int sl = pslot0 + ci;
            if (sl >= 10) sl -= 10;
            tmem_st_x8(tP + lane_addr + sl * 8, pk8);
#else
            tmem_st_x8(tP + lane_addr + (c0 >> 1), pk8);
#endif
          }
#else
          // 16 consecutive keys = two 16-byte chunks of this row inside K-block (c0 / 64)
          const uint32_t rowbase = sP + (c0 >> 6) * 16384 + row * 128;
          const int ch = (c0 & 63) >> 3;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint32_t addr = rowbase + (((ch + q) ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                         "r"(pack_bf16x2(p[q * 8 + 0], p[q * 8 + 1])), "r"(pack_bf16x2(p[q * 8 + 2], p[q * 8 + 3])),
                         "r"(pack_bf16x2(p[q * 8 + 4], p[q * 8 + 5])), "r"(pack_bf16x2(p[q * 8 + 6], p[q * 8 + 7]))
                         : "memory");
          }
#endif
        }
        if constexpr (PK2) {
          float l0, l1;
          unpack_f2(l2, l0, l1);
          l_tile = l0 + l1;
        }
        const bool grow = mx > 8.0f;  // relative to m_ref
        if (attempt == 1 || !__any_sync(0xffffffffu, grow)) {
          l_sum += l_tile;
          break;
        }
        // rescale the running state to the new reference point and redo the tile
#if PK_ATTN_FWD_TS == 2
        if (j >= 1) {   // S(j) was issued ahead of P(j-1).V(j-1): O is only stable once that has retired
          mbar_wait(bar_ve, (j - 1) & 1);
          tc_fence_after();
        }
#endif
        float alpha = 1.0f;
        if (grow) {
          alpha = fast_exp2(-mx);
          m_ref += mx;
          l_sum *= alpha;
        }
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t o[16];
          tmem_ld_x16(tO + lane_addr + c0, o);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
          tmem_st_x16(tO + lane_addr + c0, o);
        }
        tmem_wait_st();
      }
      if (row == 0) ATT_TRACE(2, j, 2);
#if PK_ATTN_FWD_TS
      tmem_wait_st();
#else
      fence_proxy_async_smem();
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      if (row == 0) ATT_TRACE(2, j, 3);
    }
    // ---------------- epilogue: O / l -> bf16, LSE ----------------
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.0f / l_sum;
    __nv_bfloat16* orow = a.out + (static_cast<size_t>(b) * a.N + t) * a.ldo + head * 64;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t o[16];
      tmem_ld_x16(tO + lane_addr + c0, o);
      tmem_wait_ld();
      if (valid) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c0 + q * 8) = u;
        }
      }
    }
    if (valid && a.lse)
      a.lse[(static_cast<size_t>(b) * a.heads + head) * a.N + t] = m_ref + log2f(l_sum);
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 5) tmem_dealloc(tmem, 256);
}

}  // namespace pk

using namespace pk;

static long long* g_attn_trace = nullptr;

// debug hook: device buffer of 3*16*8 int64 receiving a clock64 timeline of CTA (0,0,0); nullptr disables
extern "C" void pk_attn_set_trace(void* buf) { g_attn_trace = static_cast<long long*>(buf); }

// qkv: bf16 [B*N, 3C] (columns (3, head, 64) as produced by the qkv Linear, models_painter.py:76-78)
// th / tw: bf16 rel-pos tables, zero-padded to th_pad / tw_pad rows (multiples of 16), [rows, 64]
static int attn_fwd_impl(const void* qkv, const void* th, const void* tw, void* out, float* lse, float* relh_out,
                         float* relw_out, int B, int heads, int h, int w, int th_pad, int tw_pad, void* stream);

extern "C" int pk_attn_fwd(const void* qkv, const void* th, const void* tw, void* out, float* lse, int B,
                           int heads, int h, int w, int th_pad, int tw_pad, void* stream) {
  return attn_fwd_impl(qkv, th, tw, out, lse, nullptr, nullptr, B, heads, h, w, th_pad, tw_pad, stream);
}
// training forward: additionally stores every query's bias rows (relh_g / relw_g: B*heads*Np*h resp. *w floats,
// Np = N rounded up to 128) for pk_attn_bwd_saved
extern "C" int pk_attn_fwd_save(const void* qkv, const void* th, const void* tw, void* out, float* lse, float* relh_g,
                                float* relw_g, int B, int heads, int h, int w, int th_pad, int tw_pad, void* stream) {
  PK_CHECK(lse && relh_g && relw_g, "pk_attn_fwd_save: null pointer");
  return attn_fwd_impl(qkv, th, tw, out, lse, relh_g, relw_g, B, heads, h, w, th_pad, tw_pad, stream);
}

static int attn_fwd_impl(const void* qkv, const void* th, const void* tw, void* out, float* lse, float* relh_out,
                         float* relw_out, int B, int heads, int h, int w, int th_pad, int tw_pad, void* stream) {
  PK_CHECK(qkv && th && tw && out, "pk_attn_fwd: null pointer");
  PK_CHECK(B > 0 && heads > 0 && h > 0 && w > 0, "pk_attn_fwd: bad shape");
  PK_CHECK(th_pad % 16 == 0 && tw_pad % 16 == 0 && th_pad >= 2 * h - 1 && tw_pad >= 2 * w - 1 &&
               th_pad <= 256 && tw_pad <= 112,
           "pk_attn_fwd: bad table padding th_pad=%d tw_pad=%d (h=%d w=%d)", th_pad, tw_pad, h, w);
  const int N = h * w, C = heads * 64;
  AttnFwdArgs a;
  a.h = h;
  a.N = N;
  a.heads = heads;
  a.th_pad = th_pad;
  a.tw_pad = tw_pad;
  int relh_bytes = 128 * (h + 1) * 4;
  if (relh_bytes < 128 * 17 * 4) relh_bytes = 128 * 17 * 4;
  relh_bytes = (relh_bytes + 15) & ~15;
  a.relh_bytes = relh_bytes;
  a.scale_log2 = 0.125f * LOG2E;
  a.out = static_cast<__nv_bfloat16*>(out);
  a.ldo = C;
  a.lse = lse;
  a.relh_out = relh_out;
  a.relw_out = relw_out;
  a.trace = g_attn_trace;

  CUtensorMap tmQ, tmKV, tmTh, tmTw;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(3 * C), static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
    uint64_t strides[2] = {static_cast<uint64_t>(3 * C) * 2, static_cast<uint64_t>(N) * 3 * C * 2};
    uint32_t boxq[3] = {64, ATT_BM, 1};
    uint32_t boxk[3] = {64, ATT_KT, 1};
    if (!make_tmap_bf16(&tmQ, qkv, 3, dims, strides, boxq)) return 3;
    if (!make_tmap_bf16(&tmKV, qkv, 3, dims, strides, boxk)) return 3;
    uint64_t d2[2] = {64, static_cast<uint64_t>(th_pad)};
    uint64_t s2[1] = {128};
    uint32_t b2[2] = {64, static_cast<uint32_t>(th_pad)};
    if (!make_tmap_bf16(&tmTh, th, 2, d2, s2, b2)) return 3;
    d2[1] = tw_pad;
    b2[1] = tw_pad;
    if (!make_tmap_bf16(&tmTw, tw, 2, d2, s2, b2)) return 3;
  }
  const size_t smem = 1024 + ATT_SRELH + relh_bytes + 128;
  PK_CHECK(smem <= 227 * 1024, "pk_attn_fwd: h=%d needs %zu B of shared memory", h, smem);
  dim3 grid((N + ATT_BM - 1) / ATT_BM, heads, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define PK_ATT_LAUNCH(WW)                                                                                   \
  case WW: {                                                                                                \
    static bool attr = false;                                                                               \
    if (!attr) {                                                                                            \
      cudaFuncSetAttribute(attn_fwd_kernel<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);   \
      attr = true;                                                                                          \
    }                                                                                                       \
    launch_pdl(attn_fwd_kernel<WW>, grid, dim3(ATT_THREADS), smem, st, tmQ, tmKV, tmTh, tmTw, a);           \
  } break;
  switch (w) {
    PK_ATT_LAUNCH(2)
    PK_ATT_LAUNCH(4)
    PK_ATT_LAUNCH(7)
    PK_ATT_LAUNCH(8)
    PK_ATT_LAUNCH(14)
    PK_ATT_LAUNCH(28)
    PK_ATT_LAUNCH(56)
    default:
      PK_CHECK(false, "pk_attn_fwd: token-grid width %d unsupported (must divide 112: 2,4,7,8,14,28,56)", w);
  }
#undef PK_ATT_LAUNCH
  PK_LAUNCH_CHECK("pk_attn_fwd");
  return 0;
}

// fp32 table [L, 64] -> bf16 [Lpad, 64], rows >= L zero  (rel_pos_h / rel_pos_w, models_painter.py:66-67)
__global__ void pad_cast_table_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int L,
                                      int Lpad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Lpad * 64) return;
  out[idx] = __float2bfloat16_rn(idx < L * 64 ? in[idx] : 0.f);
}
extern "C" int pk_relpos_table_bf16(const float* table, void* out_bf16, int L, int Lpad, void* stream) {
  PK_CHECK(table && out_bf16 && Lpad >= L, "pk_relpos_table_bf16: bad args");
  pad_cast_table_kernel<<<(Lpad * 64 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      table, static_cast<__nv_bfloat16*>(out_bf16), L, Lpad);
  PK_LAUNCH_CHECK("pk_relpos_table_bf16");
  return 0;
}
