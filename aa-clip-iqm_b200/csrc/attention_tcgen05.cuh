// Fused flash-style multi-head attention forward for sm_100a (tcgen05 + TMEM + TMA), head_dim 64.
//
// Replaces nn.MultiheadAttention's slow path on the reference (model/transformer.py:200,226-237 ->
// F.multi_head_attention_forward): q*(1/8), bmm, softmax, bmm, plus the head-averaged weights that
// AdaptedCLIP.forward discards (model/adapter.py:162). Nothing N x N ever reaches HBM here.
//
// Input : packed QKV fp16|bf16 [B, N, 3*D] (the QKV GEMM output, q | k | v along the last axis, head h at h*64)
// Output: O same dtype [B, N, D] (heads concatenated) -- directly the A operand of the out-projection GEMM.
//
// One CTA = (image b, head h, 256 query rows) = two 128-row Q tiles that share every K/V tile:
//   warp 8      : TMA producer (Q once; K_j and V_j through separate 3-deep rings)
//   warp 9 / 10 : tcgen05.mma issuer (one lane each) for Q tile 0 / 1; warp 9 also owns the TMEM allocation.
//                 One issuer per Q tile keeps the two softmax pipelines independent: with a single in-order
//                 issuer a late s_free of one tile delayed the other tile's P V product (ncu: 18 % of the softmax
//                 warps' samples sat on o_full). Warp 11 is idle: it completes the third warpgroup so that
//                 setmaxnreg can hand its registers to the softmax warpgroups.
//   warps 0..3  : softmax for Q tile 0 (thread = one query row; TMEM lane quarter = warp % 4)
//   warps 4..7  : same for Q tile 1
// TMEM (512 cols): S_t [128x128] fp32 at t*128 | O_t [128x64] fp32 at 256+t*64 | P_t [128x128] 16-bit at 384+t*64.
//
// Per K/V tile j and Q tile t:
//   MMA   : S_t = Q_t K_j^T (SS MMA, M128 N128 K64)                                     -> s_full[t]
//   softmax: whole S row (128 fp32) TMEM -> registers in one go                         -> s_free[t]  (S_t(j+1) may
//            now be issued: the next QK^T runs under this tile's exponentials)
//            row max; LAZY rescale: the running max used for the exponentials only moves when the true max grew by
//            more than 2^8 (then O_t in TMEM and the row sum are rescaled once); P = exp2(S*scale - m) <= 256
//            packed to 16 bit and written back to TENSOR MEMORY (tcgen05.st)            -> p_full[t]
//   MMA   : O_t += P_t V_j  (TS MMA: A = P_t from TMEM, B = V_j MN-major straight from the TMA tile) -> o_full[t]
// P never touches shared memory: smem bandwidth (128 B/clk) is what bounds SS MMAs at this tile shape, and a
// P round trip through smem would double the traffic. O stays in TMEM for the whole K/V loop and is read once at
// the end (scaled by 1/rowsum). While one Q tile is in its exponential phase the tensor core works for the other.
//
// What bounds it (measured, tools/ubench/{pipe_rates2,exp_loop,tmem_rates2}.cu and the clock64 timeline behind
// -DAACLIP_ATTN_DBG_TIMELINE): each SM sub-partition runs one softmax warp of each Q tile. MUFU.EX2 issues once per
// 8 clk per sub-partition; FFMA2 / FADD2 / F2FP / FMNMX3 are half-rate and hold the issue port for 2 clk each, so the
// plain loop (scale-subtract, ex2, row sum, pack) needs 8 XU clk and ~4 issue clk per element, and every MIO
// instruction of the warp that is NOT exponentiating (mbarrier try_wait / arrive, tcgen05.ld / st) queues behind the
// other warp's MUFU stream (~140 clk per mbarrier op). Hence: (a) packed fp32 pairs for the scale-subtract and the row
// sum (same issue cost, half the instructions and registers); (b) every POLY_EVERY-th pair of exponentials goes to a
// degree-3 Cody-Waite polynomial on the FMA pipe, which balances the XU (6 clk/element) against the issue port
// (6.25 clk/element) -- 0.412 -> 0.385 ms at B=32; (c) o_full(j-1) is awaited after the first half of tile j's
// exponentials instead of before them. A steady-state tile takes ~2 550 clk per Q-tile pair against the 1 536 clk of
// MUFU work left in it; prologue + epilogue are ~11 % of a CTA's life.
// Build knobs kept for the record (all measured slower or equal on B200, see tools/experiments/README.md):
// AACLIP_ATTN_QT=1 (two one-tile CTAs per SM), AACLIP_ATTN_SPLIT=2 (four softmax warps per sub-partition, half a row
// each), AACLIP_ATTN_LAZYMAX=1 (no row maximum after the first tile, row-sum overflow check), AACLIP_ATTN_WIDE_TMEM=1
// (one .x128 load / .x64 store per tile). AACLIP_ATTN_DBG_* are timing-only what-if switches (wrong results).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace aaclip {

namespace attn_cfg {
// QT = Q tiles per CTA. 2: one CTA per SM, the two tiles share every K/V tile (default). 1: two independent CTAs
// per SM (256 TMEM columns, 32 Ki registers, ~113 KiB smem each) -- one CTA's prologue / epilogue runs under the
// other's steady state.
#ifndef AACLIP_ATTN_QT
#define AACLIP_ATTN_QT 2
#endif
// SPLIT = softmax warps per 32 query rows. 1: a thread owns a whole 128-column S row. 2: two warps of the same TMEM
// lane quarter own 64 columns each (four softmax warps per SM sub-partition instead of two: more independent
// instruction streams to cover each other's TMEM-load / maximum / barrier phases); the row maximum and the final
// row sum are exchanged through shared memory under a 64-thread named barrier.
#ifndef AACLIP_ATTN_SPLIT
#define AACLIP_ATTN_SPLIT 1
#endif
constexpr int BQ = 128, QT = AACLIP_ATTN_QT, SPLIT = AACLIP_ATTN_SPLIT, BKV = 128, HD = 64, KV_STAGES = 3;
static_assert(QT == 1 || QT == 2, "one or two Q tiles per CTA");
static_assert(SPLIT == 1 || SPLIT == 2, "one or two softmax warps per 32 rows");
static_assert(!(QT == 1 && SPLIT == 2), "the split softmax is built for the one-CTA-per-SM layout");
constexpr int CTAS_PER_SM = QT == 1 ? 2 : 1;
constexpr int TILE_BYTES = 128 * HD * 2;          // one Q / K / V tile: 128 rows x 128 B = 16 KiB
constexpr int N_SOFTMAX_WARPS = 4 * QT * SPLIT;
constexpr int THREADS = 32 * N_SOFTMAX_WARPS + 128;   // softmax warpgroups + {TMA, QT MMA issuers, idle}
constexpr int W_TMA = N_SOFTMAX_WARPS, W_MMA0 = N_SOFTMAX_WARPS + 1;
// setmaxnreg only re-distributes what the CTA was launched with (THREADS x the launch register count, which ptxas
// derives from __launch_bounds__: 384 x 168 | 640 x 96): 256 x 216 + 128 x 40 <= 64 512 | 512 x R + 128 x D <= 61 440.
#ifndef AACLIP_ATTN_SPLIT_REGS
#define AACLIP_ATTN_SPLIT_REGS 112
#endif
constexpr int SOFTMAX_REGS = SPLIT == 1 ? 216 : AACLIP_ATTN_SPLIT_REGS;
constexpr int OTHER_REGS = SPLIT == 1 ? 40 : (AACLIP_ATTN_SPLIT_REGS > 104 ? 24 : 40);
static_assert(SPLIT == 1 || 512 * SOFTMAX_REGS + 128 * OTHER_REGS <= 640 * 96, "register pool of the CTA");
constexpr int COLS = BKV / SPLIT;                 // S columns per softmax thread
constexpr int OCOLS = HD / SPLIT;                 // O columns per softmax thread (rescale / write-back)
constexpr int TMEM_COLS = 256 * QT;
constexpr int TM_S = 0, TM_O = QT * BKV, TM_P = TM_O + QT * HD;   // TMEM column map
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + QT * TILE_BYTES;
constexpr int OFF_V = OFF_K + KV_STAGES * TILE_BYTES;
constexpr int OFF_BAR = OFF_V + KV_STAGES * TILE_BYTES;
constexpr int OFF_X = OFF_BAR + 256;              // SPLIT == 2: float xmax[2][QT][2][128], xsum[QT][2][128]
constexpr int X_BYTES = SPLIT == 2 ? 3 * QT * 2 * 128 * 4 : 0;
// QT == 1: two CTAs must fit one SM's 228 KiB, so there is no room for a 1 KiB alignment slack; the dynamic smem
// window is declared 1024-aligned and the kernel traps if the base is not.
constexpr int SMEM_BYTES = OFF_X + X_BYTES + (QT == 1 ? 0 : 1024);
constexpr float RESCALE_THRESHOLD = 8.0f;         // log2 units: P never exceeds 2^8 on the exact-max path
// LAZYMAX: K/V tiles after the first skip the row maximum. The exponentials are taken against the maximum in use and
// the tile's row sum (needed anyway) is the overflow check: every P <= row sum < LAZY_LIMIT < 65504 (fp16 max). A
// warp with a row over the limit redoes the tile on the exact-max path (S is still in registers).
#ifndef AACLIP_ATTN_LAZYMAX
#define AACLIP_ATTN_LAZYMAX 0
#endif
constexpr bool LAZYMAX = AACLIP_ATTN_LAZYMAX != 0 && SPLIT == 1;
constexpr float LAZY_LIMIT = 6.0e4f;
// POLY_EVERY = n > 0: every n-th pair of exponentials is evaluated by a polynomial on the FMA pipe instead of the
// MUFU (the MUFU's 4 ex2/clk per sub-partition bound the loop; tools/ubench/exp_loop.cu measures the trade).
#ifndef AACLIP_ATTN_POLY_EVERY
#define AACLIP_ATTN_POLY_EVERY 4
#endif
constexpr int POLY_EVERY = AACLIP_ATTN_POLY_EVERY;
// one .x128 TMEM load / .x64 store per tile instead of 4 / 2 (measured: no gain, 0.395 vs 0.385 ms)
#ifndef AACLIP_ATTN_WIDE_TMEM
#define AACLIP_ATTN_WIDE_TMEM 0
#endif
constexpr bool WIDE_TMEM = AACLIP_ATTN_WIDE_TMEM != 0;
#ifndef AACLIP_ATTN_STAGGER
#define AACLIP_ATTN_STAGGER 1
#endif
constexpr bool STAGGER = AACLIP_ATTN_STAGGER != 0 && QT == 2;
}  // namespace attn_cfg

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));   // FMNMX3 on sm_100
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
#ifdef AACLIP_ATTN_DBG_NOEXP   // what-if (timing only, wrong results): no MUFU work
  return fmaf(x, 1e-3f, 1.0f);
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}
// packed fp32 pairs (sm_100: FFMA2 / FADD2 do two fp32 lanes per issue slot)
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// exp2 of a packed pair on the FMA pipe (no MUFU): Cody-Waite split x = n + f, f in [-0.5, 0.5], degree-3 minimax
// polynomial for 2^f (max relative error 7.5e-5 -- P is rounded to 16 bits, half-ulp 4.9e-4, right afterwards), n added
// into the exponent field. Inputs below -126 (masked columns hold -inf) clamp to 2^-126.
__device__ __forceinline__ void poly_exp2_pair(uint64_t x2, float& p0, float& p1) {
  constexpr float kMagic = 12582912.f;   // 1.5 * 2^23: x + kMagic holds round(x) in its low mantissa bits
  float x0, x1;
  unpack_f32x2(x2, x0, x1);
  const uint64_t x = pack_f32x2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const uint64_t xm = add_f32x2(x, pack_f32x2(kMagic, kMagic));
  const uint64_t nf = add_f32x2(xm, pack_f32x2(-kMagic, -kMagic));
  const uint64_t f = fma_f32x2(nf, pack_f32x2(-1.f, -1.f), x);
  uint64_t r = fma_f32x2(f, pack_f32x2(0.05517132207751274f, 0.05517132207751274f),
                         pack_f32x2(0.24261054396629333f, 0.24261054396629333f));
  r = fma_f32x2(r, f, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
  r = fma_f32x2(r, f, pack_f32x2(0.9999281167984009f, 0.9999281167984009f));
  float r0, r1, n0, n1;
  unpack_f32x2(r, r0, r1);
  unpack_f32x2(xm, n0, n1);
  p0 = __uint_as_float(__float_as_uint(r0) + (__float_as_uint(n0) << 23));
  p1 = __uint_as_float(__float_as_uint(r1) + (__float_as_uint(n1) << 23));
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

#ifdef AACLIP_ATTN_DBG_TIMELINE
// developer instrumentation: clock64() stamps of one CTA's softmax warps 0 (Q tile 0) and 4*SPLIT (Q tile 1)
__device__ long long g_attn_timeline[2][16][8];
#define ATTN_STAMP(k)                                                                                 \
  do {                                                                                                \
    if (blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 9 && lane == 0 && q == 0 && half == 0 && j < 16) \
      g_attn_timeline[t][j][k] = clock64();                                                           \
  } while (0)
#else
#define ATTN_STAMP(k)
#endif

template <bool BF16>
__global__ void __launch_bounds__(attn_cfg::THREADS, attn_cfg::CTAS_PER_SM)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, uint16_t* __restrict__ out, int n_tok,
                         int D, int causal, float scale_log2) {
  using namespace attn_cfg;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (QT == 1 && smem != smem_raw) __trap();   // no alignment slack in the two-CTAs-per-SM layout
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* k_full = q_full + 1;                  // [KV_STAGES] TMA -> MMA
  uint64_t* k_empty = k_full + KV_STAGES;         // [KV_STAGES] MMA -> TMA : every S MMA reading K_j has retired
  uint64_t* v_full = k_empty + KV_STAGES;         // [KV_STAGES]
  uint64_t* v_empty = v_full + KV_STAGES;         // [KV_STAGES] MMA -> TMA : every P V MMA reading V_j has retired
  uint64_t* s_full = v_empty + KV_STAGES;         // [QT] MMA -> softmax : S_t(j) complete
  uint64_t* s_free = s_full + QT;                 // [QT] softmax -> MMA : S_t(j) is in registers
  uint64_t* p_full = s_free + QT;                 // [QT] softmax -> MMA : P_t(j) in TMEM, O_t rescaled if needed
  uint64_t* o_full = p_full + QT;                 // [QT] MMA -> softmax : O_t += P_t(j) V_j complete
  uint64_t* stagger = o_full + QT;                // softmax tile 0 -> tile 1, once: start half an exp phase later
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(stagger + 1);
  [[maybe_unused]] float* xmax = reinterpret_cast<float*>(smem + OFF_X);   // [2 (tile parity)][QT][2 (half)][128]
  [[maybe_unused]] float* xsum = xmax + 2 * QT * 2 * 128;                  // [QT][2][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_base = blockIdx.x * (BQ * QT);
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_kv_tiles = (n_tok + BKV - 1) / BKV;
  const int n_qt = (QT == 2 && q_base + BQ < n_tok) ? 2 : 1;  // second Q tile entirely out of range -> skipped

  if (warp == W_TMA && lane == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&k_empty[s], QT);  // one tcgen05.commit per MMA issuer (twice from issuer 0 if n_qt < QT)
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&v_empty[s], QT);
    }
    for (int t = 0; t < QT; ++t) {
      ptx::mbar_init(&s_full[t], 1);
      ptx::mbar_init(&s_free[t], 4 * SPLIT);
      ptx::mbar_init(&p_full[t], 4 * SPLIT);
      ptx::mbar_init(&o_full[t], 1);
    }
    ptx::mbar_init(stagger, 4 * SPLIT);
    ptx::fence_mbar_init();
  }
  if (warp == W_MMA0) {
    ptx::tmem_alloc<TMEM_COLS>(tmem_base_smem);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  ptx::pdl_trigger();      // programmatic dependent launch (ptx.cuh): no global access above this line
  ptx::pdl_wait();

  // The softmax warpgroups need a whole S row in registers: warpgroup 2 (TMA, MMA, 2 idle warps) hands its
  // registers over with setmaxnreg (issued inside each role's branch so ptxas allocates per role).
  if (warp >= W_MMA0 + QT) {
    ptx::reg_dec<OTHER_REGS>();
  } else if (warp == W_TMA) {
    ptx::reg_dec<OTHER_REGS>();
    // ===================== TMA producer =====================
    // (warp-uniform loop, one elected lane issues: keeps the UTMALDG / UTCHMMA operands in uniform registers
    // without the per-instruction ELECT + R2UR.BROADCAST + BRA.U.ANY loop ptxas emits under `if (lane == 0)`)
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(q_full, QT * TILE_BYTES);
      for (int t = 0; t < QT; ++t)
        ptx::tma_load_3d(smem + OFF_Q + t * TILE_BYTES, &tmap_qkv, q_full, h * HD, q_base + t * BQ, b);
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_kv_tiles; ++j) {
      ptx::mbar_wait(&k_empty[st], ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        ptx::tma_load_3d(smem + OFF_K + st * TILE_BYTES, &tmap_qkv, &k_full[st], D + h * HD, j * BKV, b);
      }
      __syncwarp();
      ptx::mbar_wait(&v_empty[st], ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
        ptx::tma_load_3d(smem + OFF_V + st * TILE_BYTES, &tmap_qkv, &v_full[st], 2 * D + h * HD, j * BKV, b);
      }
      __syncwarp();
      if (++st == KV_STAGES) { st = 0; ph ^= 1; }
    }
  } else if (warp >= W_MMA0) {
    ptx::reg_dec<OTHER_REGS>();
    // ===================== MMA issuers (one warp per Q tile; warp-uniform loop, one elected lane issues) ===========
    const int t = warp - W_MMA0;
    if (t < n_qt) {
      constexpr uint32_t idesc_s = ptx::make_idesc_16(BQ, BKV, 0, BF16);   // S = Q K^T : both operands K-major
      constexpr uint32_t idesc_pv = ptx::make_idesc_16(BQ, HD, 1, BF16);   // O += P V  : V is MN-major
      const uint32_t q_addr = ptx::smem_u32(smem + OFF_Q) + t * TILE_BYTES;
      const uint32_t k_addr = ptx::smem_u32(smem + OFF_K);
      const uint32_t v_addr = ptx::smem_u32(smem + OFF_V);
      const bool alone = (n_qt < QT);   // then this thread supplies both arrivals of k_empty / v_empty

      auto issue_s = [&](int stage) {
        const uint64_t a = ptx::make_sw128_desc(q_addr);
        const uint64_t bd = ptx::make_sw128_desc(k_addr + stage * TILE_BYTES);
        if (ptx::elect_one()) {
#pragma unroll
#ifdef AACLIP_ATTN_DBG_NOS       // what-if (timing only, wrong results): one quarter of the Q K^T product
          for (int k = 0; k < 1; ++k)
#else
          for (int k = 0; k < HD / 16; ++k)
#endif
            ptx::tc_mma_f16(tmem_base + TM_S + t * BKV, a + 2 * k, bd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          ptx::tc_commit(&s_full[t]);
          ptx::tc_commit(&k_empty[stage]);
          if (alone) ptx::tc_commit(&k_empty[stage]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int stage, uint32_t accumulate) {
        const uint32_t va = v_addr + stage * TILE_BYTES;
        if (ptx::elect_one()) {
#pragma unroll
#ifdef AACLIP_ATTN_DBG_NOPV      // what-if (timing only, wrong results): one eighth of the P V product
          for (int k = 0; k < 1; ++k) {
#else
          for (int k = 0; k < BKV / 16; ++k) {
#endif
            // A: P_t in TMEM, 8 columns (16 packed K elements) per k-step.  B (MN-major): 16 kv rows of 128 B.
            const uint64_t bd = ptx::make_sw128_desc(va + k * 16 * 128);
            ptx::tc_mma_f16_ts(tmem_base + TM_O + t * HD, tmem_base + TM_P + t * (BKV / 2) + k * 8, bd, idesc_pv,
                               (accumulate | k) != 0 ? 1u : 0u);
          }
          ptx::tc_commit(&o_full[t]);
          ptx::tc_commit(&v_empty[stage]);
          if (alone) ptx::tc_commit(&v_empty[stage]);
        }
        __syncwarp();
      };

      ptx::mbar_wait(q_full, 0);
      ptx::mbar_wait(&k_full[0], 0);
      ptx::tc_fence_after();
      issue_s(0);

      int st = 0;             // ring slot of K_j / V_j
      uint32_t ph = 0;
      for (int j = 0; j < n_kv_tiles; ++j) {
        int nst = st + 1;
        uint32_t nph = ph;
        if (nst == KV_STAGES) { nst = 0; nph ^= 1; }
        if (j + 1 < n_kv_tiles) {
          ptx::mbar_wait(&k_full[nst], nph);
          ptx::mbar_wait(&s_free[t], j & 1);     // S_t(j) has been copied to registers
          ptx::tc_fence_after();
          issue_s(nst);                          // runs under this tile's exponentials
        }
        ptx::mbar_wait(&v_full[st], ph);
        ptx::mbar_wait(&p_full[t], j & 1);       // P_t(j) in TMEM; O_t rescaled if the running max moved
        ptx::tc_fence_after();
        issue_pv(st, j > 0 ? 1u : 0u);
        st = nst;
        ph = nph;
      }
    }
  } else {
    ptx::reg_inc<SOFTMAX_REGS>();
    // ===================== softmax warps =====================
    const int q = warp & 3;                     // TMEM lane quarter (= SM sub-partition)
    const int half = (warp >> 2) % SPLIT;       // which COLS-wide slice of the S row
    const int t = warp / (4 * SPLIT);
    if (t < n_qt) {
      const int r = q * 32 + lane;              // row inside the Q tile == TMEM lane
      const int qrow = q_base + t * BQ + r;     // token index of this query
      const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
      const uint32_t s_tmem = tmem_base + lane_off + TM_S + t * BKV + half * COLS;
      const uint32_t o_tmem = tmem_base + lane_off + TM_O + t * HD + half * OCOLS;
      const uint32_t p_tmem = tmem_base + lane_off + TM_P + t * (BKV / 2) + half * (COLS / 2);
      [[maybe_unused]] const int pair_bar = 1 + t * 4 + q;   // named barrier of the two warps sharing these 32 rows

      float m_used = -1e30f;   // the maximum the exponentials are taken against (log2 domain, scaled)
      float l = 0.f;           // running row sum of P (this thread's columns)

      // Two Q tiles share each SM sub-partition's MUFU (4 ex2/clk): if both run in lock-step, the MUFU idles while
      // both are in their load/max/wait phases and is oversubscribed while both exponentiate. Q tile 1 starts half
      // an exponential phase late, so one tile's exponentials cover the other's non-MUFU phases.
      if (STAGGER && t == 1) ptx::mbar_wait(stagger, 0);

      for (int j = 0; j < n_kv_tiles; ++j) {
        ATTN_STAMP(0);
        ptx::mbar_wait(&s_full[t], j & 1);
        ptx::tc_fence_after();
        ATTN_STAMP(1);
        uint32_t s[COLS];
#ifdef AACLIP_ATTN_DBG_NOLD      // what-if (timing only, wrong results): no TMEM read of S
#pragma unroll
        for (int i = 0; i < COLS; ++i) s[i] = __float_as_uint(1e-3f * static_cast<float>((lane + i + j) & 63));
#else
        // every MIO instruction of this warp (TMEM access, mbarrier op) queues behind the other Q tile's MUFU
        // stream: one wide load instead of four
        if constexpr (WIDE_TMEM && COLS == 128) ptx::tmem_ld_32x32b_x128(s_tmem, *reinterpret_cast<uint32_t(*)[128]>(&s[0]));
        else if constexpr (WIDE_TMEM && COLS == 64) ptx::tmem_ld_32x32b_x64(s_tmem, *reinterpret_cast<uint32_t(*)[64]>(&s[0]));
        else {
#pragma unroll
          for (int c = 0; c < COLS / 32; ++c)
            ptx::tmem_ld_32x32b_x32(s_tmem + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
        }
        ptx::tmem_ld_wait();
#endif
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&s_free[t]);
        ATTN_STAMP(2);

        const int col0 = j * BKV + half * COLS;
        int lim = n_tok - col0;                       // number of valid columns in this thread's slice
        if (causal) lim = min(lim, qrow - col0 + 1);
        const bool full_tile = __all_sync(0xffffffffu, lim >= COLS);
        if (!full_tile) {
#pragma unroll
          for (int i = 0; i < COLS; ++i)
            if (i >= lim) s[i] = 0xff800000u;   // -inf: exp2 gives 0, contributes nothing
        }

        // P_t (TMEM) and O_t are still being read / accumulated by the P V product of the previous tile: o_full(j-1)
        // is awaited as late as possible -- right before the first write to P_t or O_t.
        bool o_done = (j == 0);
        auto wait_o = [&]() {
          if (!o_done) {
            ptx::mbar_wait(&o_full[t], (j - 1) & 1);
            ptx::tc_fence_after();
            o_done = true;
          }
        };
        // exact: row maximum first (first tile, masked tiles, and the redo of a tile whose lazy pass overflowed)
#ifdef AACLIP_ATTN_DBG_NOMAX      // what-if (timing only, wrong results): no row maximum after the first tile
        bool exact = j == 0;
#else
        bool exact = !LAZYMAX || j == 0 || !full_tile;
#endif
        float rs_tot;
#pragma unroll 1
        for (;;) {
          if (exact) {
            // ---- row maximum: independent FMNMX3 chains (one 128-long dependent chain costs ~500 cycles)
            float mxp[COLS / 16];
#pragma unroll
            for (int c = 0; c < COLS / 16; ++c) {
              mxp[c] = max3(__uint_as_float(s[c * 16]), __uint_as_float(s[c * 16 + 1]), __uint_as_float(s[c * 16 + 2]));
#pragma unroll
              for (int i = 3; i < 15; i += 2)
                mxp[c] = max3(mxp[c], __uint_as_float(s[c * 16 + i]), __uint_as_float(s[c * 16 + i + 1]));
              mxp[c] = fmaxf(mxp[c], __uint_as_float(s[c * 16 + 15]));
            }
            float mx = max3(mxp[0], mxp[1], mxp[2]);
            mx = fmaxf(mx, mxp[3]);
            if constexpr (COLS == 128) mx = max3(mx, max3(mxp[4], mxp[5], mxp[6]), mxp[7]);
            if constexpr (SPLIT == 2) {
              // the other 64 columns of this row live in the partner warp (same lane): exchange through smem
              float* slot = xmax + ((j & 1) * QT + t) * 2 * 128;
              slot[half * 128 + r] = mx;
              named_bar_sync(pair_bar, 64);
              mx = fmaxf(mx, slot[(half ^ 1) * 128 + r]);
            }
            const float m_tile = mx * scale_log2;
            ATTN_STAMP(3);

            // ---- lazy rescale: only when the true maximum ran away from the one in use by more than 2^8
            // (SPLIT == 2: both halves see the same m_tile and m_used, so they take the same decision)
            const bool need = m_tile > m_used + RESCALE_THRESHOLD;
            if (__any_sync(0xffffffffu, need)) {
              const float m_new = need ? m_tile : m_used;
              const float alpha = ex2_approx(m_used - m_new);   // 1 for rows that keep their maximum
              if (j > 0) {
                wait_o();
#pragma unroll
                for (int c = 0; c < OCOLS / 16; ++c) {
                  uint32_t v[16];
                  ptx::tmem_ld_32x32b_x16(o_tmem + c * 16, v);
                  ptx::tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                  ptx::tmem_st_32x32b_x16(o_tmem + c * 16, v);
                }
              }
              l *= alpha;
              m_used = m_new;
            }
          }

          // ---- P = exp2(S*scale - m_used) -> packed 16-bit pairs -> tensor memory; row sum in fp32
          // The loop is bound by issue slots around the MUFU (one ex2 per 8 clk per warp): the scale-and-subtract and
          // the row sum use packed fp32 pairs (FFMA2 / FADD2), 2.5 instead of 3.5 instructions per element.
          const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
          const uint64_t negm2 = pack_f32x2(-m_used, -m_used);
          uint64_t rs2[2] = {0ull, 0ull};        // independent partial sums (no long FADD chain)
          uint32_t w[COLS / 2];                  // the whole packed P slice (one wide store) or two halves
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int i = 0; i < COLS / 4; ++i) {
              const int k = c * (COLS / 2) + 2 * i;
              const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(s[k]), __uint_as_float(s[k + 1])), scale2, negm2);
              float p0, p1;
              if (POLY_EVERY > 0 && (i % (POLY_EVERY > 0 ? POLY_EVERY : 1)) == POLY_EVERY - 1) {
                poly_exp2_pair(x2, p0, p1);    // this pair's exponentials on the FMA pipe
              } else {
                float x0, x1;
                unpack_f32x2(x2, x0, x1);
                p0 = ex2_approx(x0);
                p1 = ex2_approx(x1);
              }
#ifdef AACLIP_ATTN_DBG_NOSUM      // what-if (timing only, wrong results): no row sum in the loop
              if (i == 0) rs2[0] = add_f32x2(rs2[0], pack_f32x2(p0, p1));
#else
              rs2[i & 1] = add_f32x2(rs2[i & 1], pack_f32x2(p0, p1));
#endif
              w[c * (COLS / 4) + i] = ptx::pack16<BF16>(p0, p1);
            }
            if (c == 0) ATTN_STAMP(4);
#ifndef AACLIP_ATTN_DBG_NOST     // what-if (timing only, wrong results): no TMEM write of P
            if constexpr (!WIDE_TMEM) {
              if (c == 0) wait_o();              // the first exponentials ran under the previous P V product
              if constexpr (SPLIT == 1)
                ptx::tmem_st_32x32b_x32(p_tmem + c * (COLS / 4), *reinterpret_cast<const uint32_t(*)[32]>(&w[c * 32]));
              else
                ptx::tmem_st_32x32b_x16(p_tmem + c * (COLS / 4), *reinterpret_cast<const uint32_t(*)[16]>(&w[c * 16]));
            }
#else
            if (w[c] == 0x12345678u) l += 1.f;
#endif
            if (c == 0) ATTN_STAMP(5);
            if (STAGGER && t == 0 && j == 0 && c == 0) {
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(stagger);
            }
          }
#ifndef AACLIP_ATTN_DBG_NOST
          if constexpr (WIDE_TMEM) {
            wait_o();                            // every exponential of this tile ran under the previous P V product
            if constexpr (SPLIT == 1) ptx::tmem_st_32x32b_x64(p_tmem, *reinterpret_cast<const uint32_t(*)[64]>(&w[0]));
            else ptx::tmem_st_32x32b_x32(p_tmem, *reinterpret_cast<const uint32_t(*)[32]>(&w[0]));
          }
#endif
          {
            float r0, r1, r2, r3;
            unpack_f32x2(rs2[0], r0, r1);
            unpack_f32x2(rs2[1], r2, r3);
            rs_tot = (r0 + r1) + (r2 + r3);
          }
#ifdef AACLIP_ATTN_DBG_NOMAX
          break;
#endif
          if (exact) break;
          // lazy pass: every P <= row sum, so a row sum under the limit proves that nothing saturated
          if (!__any_sync(0xffffffffu, !(rs_tot < LAZY_LIMIT))) break;
          ptx::tmem_st_wait();                   // rare: redo this tile with the exact maximum
          exact = true;
        }
        l += rs_tot;
        ATTN_STAMP(6);

        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[t]);
        ATTN_STAMP(7);
      }

      // ---- all P V accumulated: normalise and store
      if constexpr (SPLIT == 2) {
        xsum[(t * 2 + half) * 128 + r] = l;
        named_bar_sync(pair_bar, 64);
        l += xsum[(t * 2 + (half ^ 1)) * 128 + r];
      }
      ptx::mbar_wait(&o_full[t], (n_kv_tiles - 1) & 1);
      ptx::tc_fence_after();
      const float inv_l = 1.0f / l;
      uint16_t* dst_row = out + (static_cast<size_t>(b) * n_tok + qrow) * D + h * HD + half * OCOLS;
#pragma unroll
      for (int c = 0; c < OCOLS / 32; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(o_tmem + c * 32, v);
        ptx::tmem_ld_wait();
        if (qrow < n_tok) {
          uint4* dst = reinterpret_cast<uint4*>(dst_row + c * 32);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 w;
            w.x = ptx::pack16<BF16>(__uint_as_float(v[8 * u + 0]) * inv_l, __uint_as_float(v[8 * u + 1]) * inv_l);
            w.y = ptx::pack16<BF16>(__uint_as_float(v[8 * u + 2]) * inv_l, __uint_as_float(v[8 * u + 3]) * inv_l);
            w.z = ptx::pack16<BF16>(__uint_as_float(v[8 * u + 4]) * inv_l, __uint_as_float(v[8 * u + 5]) * inv_l);
            w.w = ptx::pack16<BF16>(__uint_as_float(v[8 * u + 6]) * inv_l, __uint_as_float(v[8 * u + 7]) * inv_l);
            dst[u] = w;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace aaclip
