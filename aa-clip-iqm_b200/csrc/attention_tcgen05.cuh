// Fused flash-style multi-head attention forward for sm_100a (tcgen05 + TMEM + TMA), head_dim 64.
//
// Replaces nn.MultiheadAttention's slow path on the reference (model/transformer.py:200,226-237 ->
// F.multi_head_attention_forward): q*(1/8), bmm, softmax, bmm, plus the head-averaged weights that
// AdaptedCLIP.forward discards (model/adapter.py:162). Nothing N x N ever reaches HBM here.
//
// Input : packed QKV fp16|bf16 [B, N, 3*D] (the QKV GEMM output, q | k | v along the last axis, head h at h*64)
// Output: O same dtype [B, N, D] (heads concatenated) -- directly the A operand of the out-projection GEMM.
//
// One work item = (image b, head h, 256 query rows) = two 128-row Q tiles that share every K/V tile. The kernel is
// PERSISTENT with a DYNAMIC schedule: one CTA per SM takes items from an atomic counter (first item = blockIdx.x),
// two-tile items first and the half-empty last query block of every head at the end (longest first). All pipelines
// run on across item boundaries:
//   * the producer loads Q / K_0 / V_0 ... of item i+1 while item i computes (double-buffered Q, continuous K/V rings);
//   * the MMA issuers issue S(0) of item i+1 BEFORE P V of the last tile of item i (one flat issue order S(g+1), PV(g)
//     over all tiles of all items), so the next item's first scores are ready when its softmax starts;
//   * the two Q tiles are re-staggered by half an exponential phase at the start of EVERY item;
//   * the finished O tile leaves through shared memory and ONE bulk tensor store per softmax warp.
// A one-item-per-CTA launch (rounds 1-2a, git 0f65755) spent ~5 500 of every CTA's ~33 000 clk outside the steady state
// (1 900 clk until the first S, 1 400-1 900 waiting for the last P V, 1 200 in per-thread global stores, launch / TMEM
// allocation / teardown); this form spends ~2 500 per item: 0.380 -> 0.338 ms at B = 32 x 16 heads x 1370 tokens
// (729 TFLOP/s), 414 -> 388 mJ per launch under the 1 kW cap, attention 12.7 -> 11.4-12.2 ms per step.
// Two things that did NOT work on the way (tools/experiments/README.md): deferring the write-back of item i into the
// first tile of item i+1 (the extra code inside the tile body costs every tile more than the wait it hides), and
// leaving the stagger to chance after the first item (tiles that drift into phase run 10-25 % slower).
//   warp 8      : scheduler + TMA producer (item ring of 2 slots in shared memory -> issuers and softmax warps)
//   warp 9 / 10 : tcgen05.mma issuer (one lane each) for Q tile 0 / 1; warp 9 also owns the TMEM allocation.
//                 One issuer per Q tile keeps the two softmax pipelines independent: with a single in-order
//                 issuer a late s_free of one tile delayed the other tile's P V product (ncu: 18 % of the softmax
//                 warps' samples sat on o_full). Warp 11 is idle: it completes the third warpgroup so that
//                 setmaxnreg can hand its registers to the softmax warpgroups.
//   warps 0..3  : softmax for Q tile 0 (thread = one query row; TMEM lane quarter = warp % 4)
//   warps 4..7  : same for Q tile 1
// TMEM (512 cols): S_t [128x128] fp32 at t*128 | O_t [128x64] fp32 at 256+t*64 | P_t [128x128] 16-bit at 384+t*64.
//
// Per K/V tile and Q tile t (g = running tile count of this Q tile = the parity of the four hand-over barriers):
//   MMA   : S_t = Q_t K_j^T (SS MMA, M128 N128 K64)                                     -> s_full[t]
//   softmax: whole S row (128 fp32) TMEM -> registers in one go                         -> s_free[t]  (the next S_t may
//            now be issued: the next QK^T runs under this tile's exponentials)
//            row max; LAZY rescale: the running max used for the exponentials only moves when the true max grew by
//            more than 2^8 (then O_t in TMEM and the row sum are rescaled once); P = exp2(S*scale - m) <= 256
//            packed to 16 bit and written back to TENSOR MEMORY (tcgen05.st)            -> p_full[t]
//   MMA   : O_t += P_t V_j  (TS MMA: A = P_t from TMEM, B = V_j MN-major straight from the TMA tile) -> o_full[t]
// P never touches shared memory: smem bandwidth (128 B/clk) is what bounds SS MMAs at this tile shape, and a
// P round trip through smem would double the traffic. O stays in TMEM for the whole K/V loop and is read once at
// the end (scaled by 1/rowsum). While one Q tile is in its exponential phase the tensor core works for the other.
//
// What bounds the steady state (measured, tools/ubench/{pipe_rates2,exp_loop,tmem_rates2}.cu and the clock64 timeline
// behind -DAACLIP_ATTN_DBG_TIMELINE, tools/attn_timeline.py): each SM sub-partition runs one softmax warp of each Q
// tile. MUFU.EX2 issues once per 8 clk per sub-partition; FFMA2 / FADD2 / F2FP / FMNMX3 are half-rate and hold the issue
// port for 2 clk each, and every MIO instruction of the warp that is NOT exponentiating (mbarrier try_wait / arrive,
// tcgen05.ld / st) queues behind the other warp's MUFU stream (a try_wait on a barrier that completed long ago still
// takes ~250 clk; probing the barrier a phase early only moves that wait to the next MIO instruction). Hence: (a) packed
// fp32 pairs for the scale-subtract and the row sum; (b) every POLY_EVERY-th pair of exponentials goes to a degree-3
// Cody-Waite polynomial on the FMA pipe. A steady-state tile takes ~2 600 clk per Q-tile pair against ~1 850 issue
// clocks and 1 536 clk of MUFU work in it. The tile body has to stay straight-line code inside the 32 KB L1.5
// instruction cache: run-time column counts (a trimmed last K/V tile) or a second instantiation made EVERY tile slower.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace aaclip {

namespace attn_cfg {
#ifndef AACLIP_ATTN_KV_STAGES
#define AACLIP_ATTN_KV_STAGES 3
#endif
constexpr int BQ = 128, QT = 2, BKV = 128, HD = 64, KV_STAGES = AACLIP_ATTN_KV_STAGES, Q_STAGES = 2, ITEM_SLOTS = 2;
constexpr int TILE_BYTES = 128 * HD * 2;          // one Q / K / V tile: 128 rows x 128 B = 16 KiB
constexpr int N_SOFTMAX_WARPS = 4 * QT;
constexpr int THREADS = 32 * N_SOFTMAX_WARPS + 128;   // softmax warpgroups + {TMA, QT MMA issuers, idle}
constexpr int W_TMA = N_SOFTMAX_WARPS, W_MMA0 = N_SOFTMAX_WARPS + 1;
// setmaxnreg only re-distributes what the CTA was launched with (THREADS x the launch register count, which ptxas
// derives from __launch_bounds__: 384 x 168): 256 x 216 + 128 x 56 <= 64 512.
constexpr int SOFTMAX_REGS = 216;
#ifndef AACLIP_ATTN_OTHER_REGS
#define AACLIP_ATTN_OTHER_REGS 56
#endif
constexpr int OTHER_REGS = AACLIP_ATTN_OTHER_REGS;   // 256 x 216 + 128 x 56 = 62 464 <= 64 512: the issuers keep their cursors in registers
constexpr int TMEM_COLS = 256 * QT;
constexpr int TM_S = 0, TM_O = QT * BKV, TM_P = TM_O + QT * HD;   // TMEM column map
constexpr int OFF_Q = 0;                          // [Q_STAGES][QT] tiles
constexpr int OFF_K = OFF_Q + Q_STAGES * QT * TILE_BYTES;
constexpr int OFF_V = OFF_K + KV_STAGES * TILE_BYTES;
constexpr int OFF_O = OFF_V + KV_STAGES * TILE_BYTES;   // [QT] O tiles staged for the TMA store (4 KiB per softmax warp)
constexpr int OFF_BAR = OFF_O + QT * TILE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;  // barriers + item ring, + 1 KiB alignment slack
constexpr int ITEM_CONSUMERS = QT + N_SOFTMAX_WARPS;   // warps that read every item slot
constexpr float RESCALE_THRESHOLD = 8.0f;         // log2 units: P never exceeds 2^8
// POLY_EVERY = n > 0: every n-th pair of exponentials is evaluated by a polynomial on the FMA pipe instead of the
// MUFU (the MUFU's 4 ex2/clk per sub-partition bound the loop; tools/ubench/exp_loop.cu measures the trade).
#ifndef AACLIP_ATTN_POLY_EVERY
#define AACLIP_ATTN_POLY_EVERY 4
#endif
constexpr int POLY_EVERY = AACLIP_ATTN_POLY_EVERY;
#ifndef AACLIP_ATTN_STAGGER
#define AACLIP_ATTN_STAGGER 1
#endif
constexpr bool STAGGER = AACLIP_ATTN_STAGGER != 0;
constexpr int SCHED_SLOTS = 256;
}  // namespace attn_cfg

// Work counters of the dynamic schedule: one slot per launch in flight (the launcher hands slots out round-robin).
// `next` = items handed out beyond the first gridDim.x; `exited` = CTAs that have drawn their terminating index; the
// last of them resets the slot, so every launch finds it zeroed (device-global: zero-initialised at module load).
struct AttnSched {
  unsigned int next;
  unsigned int exited;
};
__device__ AttnSched g_attn_sched[attn_cfg::SCHED_SLOTS];

struct AttnItem {
  int b, h, q_base, n_qt;
};
// Item order. SINGLES_LAST = 1: items 0 .. n_pair_items-1 are the query blocks whose two Q tiles both hold rows, the
// one-tile last blocks of every (image, head) come at the end (longest first: the best tail balance). SINGLES_LAST = 0:
// natural order, the blocks of one (image, head) are consecutive INCLUDING its one-tile block -- every K/V tile is read
// from HBM once (the one-tile blocks at the end re-read the K/V of all 512 (image, head) pairs after they have left L2:
// ncu: 449 instead of 269 MB of DRAM reads per launch at B = 32 @518) at the price of a lumpier tail.
#ifndef AACLIP_ATTN_SINGLES_LAST
#define AACLIP_ATTN_SINGLES_LAST 0
#endif
struct AttnShape {
  int n_heads, n_qblocks, n_pair_blocks, n_pair_items, n_items, n_tok;
  __device__ __forceinline__ AttnShape(int n_tok_, int D, int n_images) {
    n_tok = n_tok_;
    n_heads = D / attn_cfg::HD;
    n_qblocks = (n_tok + 2 * attn_cfg::BQ - 1) / (2 * attn_cfg::BQ);
    n_pair_blocks = n_qblocks - (((n_qblocks - 1) * 2 * attn_cfg::BQ + attn_cfg::BQ >= n_tok) ? 1 : 0);
    n_pair_items = n_images * n_heads * n_pair_blocks;
    n_items = n_images * n_heads * n_qblocks;
  }
  __device__ __forceinline__ AttnItem decode(int idx) const {
    int bh, qb, n_qt;
    if (AACLIP_ATTN_SINGLES_LAST) {
      if (idx < n_pair_items) {
        bh = idx / n_pair_blocks;
        qb = idx - bh * n_pair_blocks;
        n_qt = 2;
      } else {
        bh = idx - n_pair_items;
        qb = n_qblocks - 1;
        n_qt = 1;
      }
    } else {
      bh = idx / n_qblocks;
      qb = idx - bh * n_qblocks;
      n_qt = qb < n_pair_blocks ? 2 : 1;
    }
    const int b = bh / n_heads;
    return AttnItem{b, bh - b * n_heads, qb * 2 * attn_cfg::BQ, n_qt};
  }
};

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));   // FMNMX3 on sm_100
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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

#ifdef AACLIP_ATTN_DBG_TIMELINE
// developer instrumentation: clock64() stamps of CTA 5's softmax warps 0 (Q tile 0) and 4 (Q tile 1) and of its two
// MMA issuers, over the first 16 tiles that CTA processes (tools/attn_timeline.py)
__device__ long long g_attn_timeline[2][16][8];
__device__ long long g_attn_timeline_mma[2][16][8];
#define ATTN_STAMP(k)                                                                                 \
  do {                                                                                                \
    if (blockIdx.x == 5 && lane == 0 && q == 0 && g < 16) g_attn_timeline[t][g][k] = clock64();       \
  } while (0)
#define ATTN_MMA_STAMP(gg, k)                                                                         \
  do {                                                                                                \
    if (blockIdx.x == 5 && lane == 0 && (gg) < 16) g_attn_timeline_mma[t][gg][k] = clock64();         \
  } while (0)
#else
#define ATTN_STAMP(k)
#define ATTN_MMA_STAMP(gg, k)
#endif

template <bool BF16>
__global__ void __launch_bounds__(attn_cfg::THREADS, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out,
                            int n_tok, int D, int causal, float scale_log2, int n_images,
                            AttnSched* __restrict__ sched, int descending) {
  using namespace attn_cfg;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);   // [Q_STAGES] TMA -> MMA
  uint64_t* q_empty = q_full + Q_STAGES;          // [Q_STAGES] MMA -> TMA : the item's last S MMAs have retired
  uint64_t* k_full = q_empty + Q_STAGES;          // [KV_STAGES] TMA -> MMA
  uint64_t* k_empty = k_full + KV_STAGES;         // [KV_STAGES] MMA -> TMA : every S MMA reading K_j has retired
  uint64_t* v_full = k_empty + KV_STAGES;         // [KV_STAGES]
  uint64_t* v_empty = v_full + KV_STAGES;         // [KV_STAGES] MMA -> TMA : every P V MMA reading V_j has retired
  uint64_t* s_full = v_empty + KV_STAGES;         // [QT] MMA -> softmax : S_t(g) complete
  uint64_t* s_free = s_full + QT;                 // [QT] softmax -> MMA : S_t(g) is in registers
  uint64_t* p_full = s_free + QT;                 // [QT] softmax -> MMA : P_t(g) in TMEM, O_t rescaled if needed
  uint64_t* o_full = p_full + QT;                 // [QT] MMA -> softmax : O_t += P_t(g) V complete
  uint64_t* item_full = o_full + QT;              // [ITEM_SLOTS] scheduler -> everyone
  uint64_t* item_empty = item_full + ITEM_SLOTS;  // [ITEM_SLOTS] everyone -> scheduler
  uint64_t* stagger = item_empty + ITEM_SLOTS;    // softmax tile 0 -> tile 1, once: start half an exp phase later
  int4* item_slot = reinterpret_cast<int4*>(stagger + 2);            // [ITEM_SLOTS] (b, h, q_base, n_qt); n_qt = 0: no more work
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(item_slot + ITEM_SLOTS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv_tiles = (n_tok + BKV - 1) / BKV;
  const AttnShape shape(n_tok, D, n_images);

  if (warp == W_TMA && lane == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::prefetch_tmap(&tmap_out);
    for (int s = 0; s < Q_STAGES; ++s) {
      ptx::mbar_init(&q_full[s], 1);
      ptx::mbar_init(&q_empty[s], QT);  // one tcgen05.commit per MMA issuer (twice from issuer 0 on a one-tile item)
    }
    for (int s = 0; s < KV_STAGES; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&k_empty[s], QT);
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&v_empty[s], QT);
    }
    for (int t = 0; t < QT; ++t) {
      ptx::mbar_init(&s_full[t], 1);
      ptx::mbar_init(&s_free[t], 4);
      ptx::mbar_init(&p_full[t], 4);
      ptx::mbar_init(&o_full[t], 1);
    }
    for (int s = 0; s < ITEM_SLOTS; ++s) {
      ptx::mbar_init(&item_full[s], 1);
      ptx::mbar_init(&item_empty[s], ITEM_CONSUMERS);
    }
    ptx::mbar_init(stagger, 4);
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

  // every consumer warp: wait for the next published item, read it, hand the slot back
  auto next_item = [&](int& slot, uint32_t& slot_ph) {
    ptx::mbar_wait(&item_full[slot], slot_ph);
    const int4 v = item_slot[slot];
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&item_empty[slot]);
    if (++slot == ITEM_SLOTS) { slot = 0; slot_ph ^= 1; }
    return AttnItem{v.x, v.y, v.z, v.w};
  };

  // The softmax warpgroups need a whole S row in registers: warpgroup 2 (TMA, MMA, idle warp) hands its
  // registers over with setmaxnreg (issued inside each role's branch so ptxas allocates per role).
  if (warp >= W_MMA0 + QT) {
    ptx::reg_dec<OTHER_REGS>();
  } else if (warp == W_TMA) {
    ptx::reg_dec<OTHER_REGS>();
    // ===================== scheduler + TMA producer =====================
    // (warp-uniform loops, one elected lane issues: keeps the UTMALDG operands in uniform registers)
    int slot = 0, qs = 0, st = 0;
    uint32_t slot_ph = 0, q_ph = 0, ph = 0;
    int idx = blockIdx.x;
    for (;;) {
      const bool more = idx < shape.n_items;
      AttnItem it{0, 0, 0, 0};
      // the only place that divides: everyone else reads the decoded item. descending: last image first -- the rows the
      // QKV GEMM wrote last (when it ran ascending) are still in L2 (gemm_tcgen05.cuh: GemmEpi::descending)
      if (more) it = shape.decode(descending ? shape.n_items - 1 - idx : idx);
      ptx::mbar_wait(&item_empty[slot], slot_ph ^ 1);
      if (lane == 0) {
        item_slot[slot] = make_int4(it.b, it.h, it.q_base, it.n_qt);
        ptx::mbar_arrive(&item_full[slot]);     // release: the item is visible to whoever sees the phase complete
      }
      __syncwarp();
      if (++slot == ITEM_SLOTS) { slot = 0; slot_ph ^= 1; }
      if (!more) break;
      // draw the next item now: the atomic's round trip runs under this item's loads
      int nxt = 0;
      if (lane == 0) nxt = static_cast<int>(gridDim.x + atomicAdd(&sched->next, 1u));
      ptx::mbar_wait(&q_empty[qs], q_ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&q_full[qs], QT * TILE_BYTES);
        for (int t = 0; t < QT; ++t)
          ptx::tma_load_3d(smem + OFF_Q + (qs * QT + t) * TILE_BYTES, &tmap_qkv, &q_full[qs], it.h * HD,
                           it.q_base + t * BQ, it.b);
      }
      __syncwarp();
      if (++qs == Q_STAGES) { qs = 0; q_ph ^= 1; }
      for (int j = 0; j < n_kv_tiles; ++j) {
        ptx::mbar_wait(&k_empty[st], ph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
          ptx::tma_load_3d(smem + OFF_K + st * TILE_BYTES, &tmap_qkv, &k_full[st], D + it.h * HD, j * BKV, it.b);
        }
        __syncwarp();
        ptx::mbar_wait(&v_empty[st], ph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
          ptx::tma_load_3d(smem + OFF_V + st * TILE_BYTES, &tmap_qkv, &v_full[st], 2 * D + it.h * HD, j * BKV, it.b);
        }
        __syncwarp();
        if (++st == KV_STAGES) { st = 0; ph ^= 1; }
      }
      idx = __shfl_sync(0xffffffffu, nxt, 0);
    }
    // this CTA has drawn its terminating index; the last CTA to get here zeroes the counters for the next launch
    if (lane == 0 && atomicAdd(&sched->exited, 1u) == gridDim.x - 1) {
      sched->next = 0;
      sched->exited = 0;
    }
  } else if (warp >= W_MMA0) {
    ptx::reg_dec<OTHER_REGS>();
    // ===================== MMA issuers (one warp per Q tile; warp-uniform loops, one elected lane issues) ===========
    // Issue order over ALL tiles of ALL items of this Q tile: S(0); then for every further tile S(g+1), PV(g) -- so
    // the first S of the next item is issued before the last P V of the current one.
    const int t = warp - W_MMA0;
    constexpr uint32_t idesc_s = ptx::make_idesc_16(BQ, BKV, 0, BF16);   // S = Q K^T : both operands K-major
    constexpr uint32_t idesc_pv = ptx::make_idesc_16(BQ, HD, 1, BF16);   // O += P V  : V is MN-major
    const uint32_t q_addr0 = ptx::smem_u32(smem + OFF_Q) + t * TILE_BYTES;
    const uint32_t k_addr = ptx::smem_u32(smem + OFF_K);
    const uint32_t v_addr = ptx::smem_u32(smem + OFF_V);
    int slot = 0, qs = 0, ks = 0, vs = 0;    // item ring, Q stage, K ring (S cursor), V ring (P V cursor)
    uint32_t slot_ph = 0, q_ph = 0, kph = 0, vph = 0;
    uint32_t gs = 0, gp = 0;                 // S / P V products issued for this Q tile
    bool pv_pending = false, pv_alone = false;
    uint32_t pv_acc = 0;

    auto flush_pv = [&]() {                  // P V of the tile whose S was issued one step ago
      ATTN_MMA_STAMP(gp, 4);
      ptx::mbar_wait(&v_full[vs], vph);
      ptx::mbar_wait(&p_full[t], gp & 1);    // P_t in TMEM; O_t rescaled if the running max moved
      ptx::tc_fence_after();
      ATTN_MMA_STAMP(gp, 5);
      const uint32_t va = v_addr + vs * TILE_BYTES;
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          // A: P_t in TMEM, 8 columns (16 packed K elements) per k-step.  B (MN-major): 16 kv rows of 128 B.
          const uint64_t bd = ptx::make_sw128_desc(va + k * 16 * 128);
          ptx::tc_mma_f16_ts(tmem_base + TM_O + t * HD, tmem_base + TM_P + t * (BKV / 2) + k * 8, bd, idesc_pv,
                             (pv_acc | k) != 0 ? 1u : 0u);
        }
        ptx::tc_commit(&o_full[t]);
        ptx::tc_commit(&v_empty[vs]);
        if (pv_alone) ptx::tc_commit(&v_empty[vs]);
      }
      __syncwarp();
      ATTN_MMA_STAMP(gp, 6);
      ++gp;
      if (++vs == KV_STAGES) { vs = 0; vph ^= 1; }
      pv_pending = false;
    };

    // One loop body for every case, so that the (unrolled) S and P V issue sequences exist once in the code: a live item
    // takes n_kv_tiles steps (S(j), then the pending P V of the previous tile); an item this issuer does not work on
    // (one-tile item, issuer 1) and the end of the work take one step that only drains the pending P V.
    for (;;) {
      const AttnItem it = next_item(slot, slot_ph);
      const bool end = it.n_qt == 0;
      const bool live = !end && t < it.n_qt;
      const bool alone = live && it.n_qt < QT;   // issuer 0 then supplies both arrivals of the shared barriers
      const uint32_t q_addr = q_addr0 + qs * QT * TILE_BYTES;
      if (live) ptx::mbar_wait(&q_full[qs], q_ph);
      const int steps = live ? n_kv_tiles : 1;
      for (int j = 0; j < steps; ++j) {
        if (live) {
          ATTN_MMA_STAMP(gs, 0);
          ptx::mbar_wait(&k_full[ks], kph);
          ATTN_MMA_STAMP(gs, 1);
          if (gs > 0) ptx::mbar_wait(&s_free[t], (gs - 1) & 1);   // the previous S_t has been copied to registers
          ptx::tc_fence_after();
          ATTN_MMA_STAMP(gs, 2);
          const uint64_t a = ptx::make_sw128_desc(q_addr);
          const uint64_t bd = ptx::make_sw128_desc(k_addr + ks * TILE_BYTES);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
              ptx::tc_mma_f16(tmem_base + TM_S + t * BKV, a + 2 * k, bd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
            ptx::tc_commit(&s_full[t]);
            ptx::tc_commit(&k_empty[ks]);
            if (alone) ptx::tc_commit(&k_empty[ks]);
            if (j + 1 == n_kv_tiles) {          // the Q tiles of this item are no longer needed
              ptx::tc_commit(&q_empty[qs]);
              if (alone) ptx::tc_commit(&q_empty[qs]);
            }
          }
          __syncwarp();
          ATTN_MMA_STAMP(gs, 3);
          ++gs;
          if (++ks == KV_STAGES) { ks = 0; kph ^= 1; }
        }
        if (pv_pending) flush_pv();             // runs after the NEXT tile's S has been issued
        pv_pending = live;
        pv_acc = j > 0 ? 1u : 0u;
        pv_alone = alone;
      }
      if (end) break;
      if (!live) {                              // the rings move on without this issuer
        for (int j = 0; j < n_kv_tiles; ++j) {
          if (++ks == KV_STAGES) { ks = 0; kph ^= 1; }
          if (++vs == KV_STAGES) { vs = 0; vph ^= 1; }
        }
      }
      if (++qs == Q_STAGES) { qs = 0; q_ph ^= 1; }
    }
  } else {
    ptx::reg_inc<SOFTMAX_REGS>();
    // ===================== softmax warps =====================
    const int q = warp & 3;                     // TMEM lane quarter (= SM sub-partition)
    const int t = warp >> 2;
    const int r = q * 32 + lane;                // row inside the Q tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_off + TM_S + t * BKV;
    const uint32_t o_tmem = tmem_base + lane_off + TM_O + t * HD;
    const uint32_t p_tmem = tmem_base + lane_off + TM_P + t * (BKV / 2);
    int slot = 0;
    uint32_t slot_ph = 0;
    uint32_t g = 0;            // tiles processed for this Q tile: parity of s_full / s_free / p_full / o_full
    uint32_t n_seen = 0;       // items seen (live or not): Q tile 0 completes one phase of `stagger` per item

    // O_t complete in TMEM -> out[b, row0 .. row0+31, h*64 .. +64) = O / rowsum for this warp's 32 rows: normalised,
    // packed, staged in shared memory (row = 128 B, 16-byte chunks XOR-swizzled as SWIZZLE_128B expects) and written
    // with ONE bulk tensor store per warp; rows beyond the sequence are clipped by the tensor map. (Per-thread global
    // stores put 32 different rows into every store instruction: ~1 200 clk per item.)
    uint8_t* const o_stage = smem + OFF_O + (t * 4 + q) * 4096;
    const uint32_t o_stage_row = ptx::smem_u32(o_stage) + lane * 128;
    auto write_back = [&](float lsum, int b, int h, int row0) {
      const float inv_l = 1.0f / lsum;
      if (lane == 0) ptx::bulk_wait_group_read0();   // the previous item's store has read the staging tile
      __syncwarp();
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(o_tmem + c * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t w0 = ptx::pack16<BF16>(__uint_as_float(v[8 * u + 0]) * inv_l, __uint_as_float(v[8 * u + 1]) * inv_l);
          const uint32_t w1 = ptx::pack16<BF16>(__uint_as_float(v[8 * u + 2]) * inv_l, __uint_as_float(v[8 * u + 3]) * inv_l);
          const uint32_t w2 = ptx::pack16<BF16>(__uint_as_float(v[8 * u + 4]) * inv_l, __uint_as_float(v[8 * u + 5]) * inv_l);
          const uint32_t w3 = ptx::pack16<BF16>(__uint_as_float(v[8 * u + 6]) * inv_l, __uint_as_float(v[8 * u + 7]) * inv_l);
          ptx::sts128(o_stage_row + (((c * 4 + u) ^ (lane & 7)) << 4), w0, w1, w2, w3);
        }
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_3d(&tmap_out, o_stage, h * HD, row0, b);
        ptx::bulk_commit_group();
      }
    };

    for (;;) {
      const AttnItem it = next_item(slot, slot_ph);
      if (it.n_qt == 0) break;
      ++n_seen;
      if (t >= it.n_qt) continue;               // one-tile item: this warpgroup has nothing to do
      const int qrow = it.q_base + t * BQ + r;  // token index of this query

      float m_used = -1e30f;   // the maximum the exponentials are taken against (log2 domain, scaled)
      float l = 0.f;           // running row sum of P

      // Two Q tiles share each SM sub-partition's MUFU (4 ex2/clk): if both run in lock-step, the MUFU idles while
      // both are in their load/max/wait phases and is oversubscribed while both exponentiate. Q tile 1 starts half
      // an exponential phase late, so one tile's exponentials cover the other's non-MUFU phases.
      // (re-done at the start of EVERY item: the write-back and the item hand-over do not cost the two tiles the same,
      // and two tiles that drift into phase slow every tile of the item down by 10-25 %)
      if (STAGGER && t == 1) ptx::mbar_wait(stagger, (n_seen - 1) & 1);

      for (int j = 0; j < n_kv_tiles; ++j, ++g) {
        ATTN_STAMP(0);
        ptx::mbar_wait(&s_full[t], g & 1);
        ptx::tc_fence_after();
        ATTN_STAMP(1);
        uint32_t s[BKV];
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c)
          ptx::tmem_ld_32x32b_x32(s_tmem + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&s_free[t]);
        ATTN_STAMP(2);

        const int col0 = j * BKV;
        int lim = n_tok - col0;                       // number of valid columns of this tile for this row
        if (causal) lim = min(lim, qrow - col0 + 1);
        const bool full_tile = __all_sync(0xffffffffu, lim >= BKV);
        if (!full_tile) {
#pragma unroll
          for (int i = 0; i < BKV; ++i)
            if (i >= lim) s[i] = 0xff800000u;   // -inf: exp2 gives 0, contributes nothing
        }

        // P_t (TMEM) and O_t are still being read / accumulated by the P V product of the previous tile: o_full(g-1)
        // is awaited as late as possible -- right before the first write to P_t or O_t. (First tile of an item: the
        // previous item's last product was awaited before its write-back.)
        bool o_done = (j == 0);
        auto wait_o = [&]() {
          if (!o_done) {
            ptx::mbar_wait(&o_full[t], (g - 1) & 1);
            ptx::tc_fence_after();
            o_done = true;
          }
        };

        // ---- row maximum: independent FMNMX3 chains (one 128-long dependent chain costs ~500 cycles)
        float mxp[BKV / 16];
#pragma unroll
        for (int c = 0; c < BKV / 16; ++c) {
          mxp[c] = max3(__uint_as_float(s[c * 16]), __uint_as_float(s[c * 16 + 1]), __uint_as_float(s[c * 16 + 2]));
#pragma unroll
          for (int i = 3; i < 15; i += 2)
            mxp[c] = max3(mxp[c], __uint_as_float(s[c * 16 + i]), __uint_as_float(s[c * 16 + i + 1]));
          mxp[c] = fmaxf(mxp[c], __uint_as_float(s[c * 16 + 15]));
        }
        float mx = max3(mxp[0], mxp[1], mxp[2]);
        mx = fmaxf(mx, mxp[3]);
        mx = max3(mx, max3(mxp[4], mxp[5], mxp[6]), mxp[7]);
        const float m_tile = mx * scale_log2;
        ATTN_STAMP(3);

        // ---- lazy rescale: only when the true maximum ran away from the one in use by more than 2^8
        const bool need = m_tile > m_used + RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? m_tile : m_used;
          const float alpha = ex2_approx(m_used - m_new);   // 1 for rows that keep their maximum
          if (j > 0) {
            wait_o();
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
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
        // ---- P = exp2(S*scale - m_used) -> packed 16-bit pairs -> tensor memory; row sum in fp32
        // The loop is bound by issue slots around the MUFU (one ex2 per 8 clk per warp): the scale-and-subtract and
        // the row sum use packed fp32 pairs (FFMA2 / FADD2), 2.5 instead of 3.5 instructions per element.
        const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
        const uint64_t negm2 = pack_f32x2(-m_used, -m_used);
        uint64_t rs2[2] = {0ull, 0ull};        // independent partial sums (no long FADD chain)
        uint32_t w[BKV / 2];                   // the packed P row, stored in two halves
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < BKV / 4; ++i) {
            const int k = c * (BKV / 2) + 2 * i;
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
            rs2[i & 1] = add_f32x2(rs2[i & 1], pack_f32x2(p0, p1));
            w[c * (BKV / 4) + i] = ptx::pack16<BF16>(p0, p1);
          }
          if (c == 0) {
            ATTN_STAMP(4);
            wait_o();                        // the first exponentials ran under the previous P V product
          }
          ptx::tmem_st_32x32b_x32(p_tmem + c * (BKV / 4), *reinterpret_cast<const uint32_t(*)[32]>(&w[c * (BKV / 4)]));
          if (c == 0) {
            ATTN_STAMP(5);
            if (STAGGER && t == 0 && j == 0) {
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(stagger);
            }
          }
        }
        {
          float r0, r1, r2, r3;
          unpack_f32x2(rs2[0], r0, r1);
          unpack_f32x2(rs2[1], r2, r3);
          l += (r0 + r1) + (r2 + r3);
        }
        ATTN_STAMP(6);

        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[t]);
        ATTN_STAMP(7);
      }

      // ---- write the item back once its last P V product has landed. (Deferring this into the first tile of the next
      // item, under its row maximum and first exponentials, was measured SLOWER: 0.366 vs 0.338 ms -- the write-back
      // code inside the tile body costs every tile more than the ~1 500-clk wait it hides once per item.) The next
      // item's Q / K / V tiles are already in flight and its first S product is already issued at this point.
      ptx::mbar_wait(&o_full[t], (g - 1) & 1);
      ptx::tc_fence_after();
      write_back(l, it.b, it.h, it.q_base + t * BQ + q * 32);
    }
    if (lane == 0) ptx::bulk_wait_group0();     // the last store has left shared memory (and is complete) before exit
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace aaclip
