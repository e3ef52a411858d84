// Fused flash-style multi-head attention forward for sm_100a (tcgen05 + TMEM + TMA), head_dim 64.
//
// Replaces nn.MultiheadAttention's slow path on the reference (model/transformer.py:200,226-237 ->
// F.multi_head_attention_forward): q*(1/8), bmm, softmax, bmm, plus the head-averaged weights that
// AdaptedCLIP.forward discards (model/adapter.py:162). Nothing N x N ever reaches HBM here.
//
// Input : packed QKV fp16|bf16 [B, N, 3*D] (the QKV GEMM output, q | k | v along the last axis, head h at h*64)
// Output: O same dtype [B, N, D] (heads concatenated) -- directly the A operand of the out-projection GEMM.
//
// One CTA = (image b, head h, 256 query rows) = two 128-row Q tiles that share every K/V tile (64 kv rows each):
//   warps 0..3  : softmax for Q tile 0 (thread = one query row; TMEM lane quarter = warp % 4)
//   warps 4..7  : same for Q tile 1
//   warp 8      : TMA producer (Q once; K_j and V_j through separate 6-deep rings)
//   warp 9 / 10 : tcgen05.mma issuer of S_t = Q_t K_j^T for Q tile 0 / 1          (one elected lane each)
//   warp 11 / 12: tcgen05.mma issuer of O_t += P_t V_j   for Q tile 0 / 1; warp 9 also owns the TMEM allocation
//   warps 13..15: idle (they complete the fourth warpgroup so that setmaxnreg can move its registers)
// One issuer per (Q tile, product) keeps all four MMA streams independent: every issuer only waits for its own
// operands, so a late softmax warp of one tile never delays the other tile or the other product.
//
// TMEM (512 columns): S_t double-buffered [128 x 64] fp32 at t*128 + b*64 | O_t [128 x 64] fp32 at 256 + t*64 |
//                     P_t double-buffered [128 x 64] 16-bit at 384 + t*64 + b*32.
//
// The MUFU (ex2: 4 per clock per SM sub-partition, shared by one warp of each Q tile) is the binding unit of this
// kernel: 64 * 128 ex2 per (Q tile, K/V tile) = 512 MUFU cycles against 320 tensor-pipe cycles. The softmax loop is
// therefore software-pipelined inside each thread so that the MUFU never waits for anything else:
//   iteration j holds S(j) in 64 registers (its row maximum is already known) and
//     - starts the TMEM load of S(j+1) into the other 64 registers,
//     - exponentiates the first half of S(j),
//     - (one sync point) waits for the load and for the tcgen05.st of P(j-1); publishes s_free(j+1), p_full(j-1),
//     - exponentiates the second half of S(j) INTERLEAVED with the row maximum of S(j+1) (FMNMX3, ALU pipe),
//     - writes P(j) to its TMEM buffer (tcgen05.st, completed at the next sync point),
//     - LAZY rescale: the maximum the exponentials are taken against only moves when the true maximum grew by more
//       than 2^8 (then O_t in TMEM and the row sum are rescaled once, after P V of this tile has drained).
// P never touches shared memory (TS MMA: A = P_t from TMEM, B = V_j MN-major straight from the TMA tile), O stays in
// TMEM for the whole K/V loop and is read once at the end (scaled by 1/rowsum).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include "ptx.cuh"

namespace aaclip {

namespace attn_cfg {
constexpr int BQ = 128, QT = 2, BKV = 64, HD = 64, KV_STAGES = 6;
constexpr int Q_TILE_BYTES = BQ * HD * 2;          // 16 KiB
constexpr int KV_TILE_BYTES = BKV * HD * 2;        // 8 KiB
constexpr int THREADS = 512;                       // 4 warpgroups: softmax 0, softmax 1, {TMA, issuers}, {issuer, idle}
constexpr int TMEM_COLS = 512;
constexpr int TM_S = 0, TM_O = 256, TM_P = 384;    // TMEM column map (see above)
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + QT * Q_TILE_BYTES;
constexpr int OFF_V = OFF_K + KV_STAGES * KV_TILE_BYTES;
constexpr int OFF_BAR = OFF_V + KV_STAGES * KV_TILE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
constexpr float RESCALE_THRESHOLD = 8.0f;          // log2 units: P never exceeds 2^8
#ifndef AACLIP_ATTN_STAGGER
#define AACLIP_ATTN_STAGGER 1
#endif
constexpr bool STAGGER = AACLIP_ATTN_STAGGER != 0;
}  // namespace attn_cfg

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

template <bool BF16>
__global__ void __launch_bounds__(attn_cfg::THREADS, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                         uint16_t* __restrict__ out, int n_tok, int D, int causal, float scale_log2) {
  using namespace attn_cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* k_full = q_full + 1;                  // [KV_STAGES] TMA -> S issuers
  uint64_t* k_empty = k_full + KV_STAGES;         // [KV_STAGES] S issuers -> TMA (2 arrivals)
  uint64_t* v_full = k_empty + KV_STAGES;         // [KV_STAGES] TMA -> PV issuers
  uint64_t* v_empty = v_full + KV_STAGES;         // [KV_STAGES] PV issuers -> TMA (2 arrivals)
  uint64_t* s_full = v_empty + KV_STAGES;         // [QT][2] S issuer -> softmax : S_t(j) complete in buffer j & 1
  uint64_t* s_free = s_full + 2 * QT;             // [QT][2] softmax -> S issuer : that buffer is in registers
  uint64_t* p_full = s_free + 2 * QT;             // [QT][2] softmax -> PV issuer: P_t(j) in TMEM buffer j & 1
  uint64_t* p_free = p_full + 2 * QT;             // [QT][2] PV issuer -> softmax: O_t += P_t(j) V_j complete
  uint64_t* stagger = p_free + 2 * QT;            // softmax WG0 -> WG1, once: start half an exponential phase later
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(stagger + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_base = blockIdx.x * (BQ * QT);
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (n_tok + BKV - 1) / BKV;
  const int n_qt = (q_base + BQ < n_tok) ? 2 : 1;  // second Q tile entirely out of range -> skipped
  const bool alone = (n_qt == 1);                  // then issuer 0 supplies both arrivals of k_empty / v_empty

  if (warp == 8 && lane == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_kv);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&k_empty[s], 2);
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&v_empty[s], 2);
    }
    for (int i = 0; i < 2 * QT; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&s_free[i], 4);
      ptx::mbar_init(&p_full[i], 4);
      ptx::mbar_init(&p_free[i], 1);
    }
    ptx::mbar_init(stagger, 4);
    ptx::fence_mbar_init();
  }
  if (warp == 9) {
    ptx::tmem_alloc<TMEM_COLS>(tmem_base_smem);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  // The softmax warpgroups keep two S tiles in registers: warpgroups 2 and 3 hand their registers over with
  // setmaxnreg (issued inside each role's branch so ptxas allocates per role).
  if (warp >= 13) {
    ptx::reg_dec<40>();
  } else if (warp == 8) {
    ptx::reg_dec<40>();
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(q_full, QT * Q_TILE_BYTES);
      for (int t = 0; t < QT; ++t)
        ptx::tma_load_3d(smem + OFF_Q + t * Q_TILE_BYTES, &tmap_q, q_full, h * HD, q_base + t * BQ, b);
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_kv; ++j) {
      ptx::mbar_wait(&k_empty[st], ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&k_full[st], KV_TILE_BYTES);
        ptx::tma_load_3d(smem + OFF_K + st * KV_TILE_BYTES, &tmap_kv, &k_full[st], D + h * HD, j * BKV, b);
      }
      __syncwarp();
      ptx::mbar_wait(&v_empty[st], ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&v_full[st], KV_TILE_BYTES);
        ptx::tma_load_3d(smem + OFF_V + st * KV_TILE_BYTES, &tmap_kv, &v_full[st], 2 * D + h * HD, j * BKV, b);
      }
      __syncwarp();
      if (++st == KV_STAGES) { st = 0; ph ^= 1; }
    }
  } else if (warp == 9 || warp == 10) {
    ptx::reg_dec<40>();
    // ===================== S = Q_t K_j^T issuer =====================
    const int t = warp - 9;
    if (t < n_qt) {
      constexpr uint32_t idesc_s = ptx::make_idesc_16(BQ, BKV, 0, BF16);   // both operands K-major
      const uint64_t q_desc = ptx::make_sw128_desc(ptx::smem_u32(smem + OFF_Q) + t * Q_TILE_BYTES);
      const uint32_t k_addr = ptx::smem_u32(smem + OFF_K);
      ptx::mbar_wait(q_full, 0);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        const int sb = j & 1;
        ptx::mbar_wait(&k_full[st], ph);
        if (j >= 2) ptx::mbar_wait(&s_free[2 * t + sb], ((j - 2) >> 1) & 1);   // S_t(j-2) has been copied to registers
        ptx::tc_fence_after();
        const uint64_t k_desc = ptx::make_sw128_desc(k_addr + st * KV_TILE_BYTES);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            ptx::tc_mma_f16(tmem_base + TM_S + t * 128 + sb * BKV, q_desc + 2 * k, k_desc + 2 * k, idesc_s,
                            k != 0 ? 1u : 0u);
          ptx::tc_commit(&s_full[2 * t + sb]);
          ptx::tc_commit(&k_empty[st]);
          if (alone) ptx::tc_commit(&k_empty[st]);
        }
        __syncwarp();
        if (++st == KV_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 11 || warp == 12) {
    ptx::reg_dec<40>();
    // ===================== O_t += P_t V_j issuer =====================
    const int t = warp - 11;
    if (t < n_qt) {
      constexpr uint32_t idesc_pv = ptx::make_idesc_16(BQ, HD, 1, BF16);   // V is the MN-major B operand
      const uint32_t v_addr = ptx::smem_u32(smem + OFF_V);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        const int pb = j & 1;
        ptx::mbar_wait(&v_full[st], ph);
        ptx::mbar_wait(&p_full[2 * t + pb], (j >> 1) & 1);   // P_t(j) in TMEM; O_t rescaled if the running max moved
        ptx::tc_fence_after();
        const uint32_t va = v_addr + st * KV_TILE_BYTES;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            // A: P_t in TMEM, 8 columns (16 packed K elements) per k-step.  B (MN-major): 16 kv rows of 128 B.
            const uint64_t v_desc = ptx::make_sw128_desc(va + k * 16 * 128);
            ptx::tc_mma_f16_ts(tmem_base + TM_O + t * HD, tmem_base + TM_P + t * 64 + pb * 32 + k * 8, v_desc,
                               idesc_pv, (j | k) != 0 ? 1u : 0u);
          }
          ptx::tc_commit(&p_free[2 * t + pb]);
          ptx::tc_commit(&v_empty[st]);
          if (alone) ptx::tc_commit(&v_empty[st]);
        }
        __syncwarp();
        if (++st == KV_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else {
    ptx::reg_inc<216>();
    // ===================== softmax warps =====================
    const int t = warp >> 2;
    const int q = warp & 3;
    if (t < n_qt) {
      const int r = q * 32 + lane;              // row inside the Q tile == TMEM lane
      const int qrow = q_base + t * BQ + r;     // token index of this query
      const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
      const uint32_t s_tmem = tmem_base + lane_off + TM_S + t * 128;
      const uint32_t o_tmem = tmem_base + lane_off + TM_O + t * HD;
      const uint32_t p_tmem = tmem_base + lane_off + TM_P + t * 64;

      float m_used = -1e30f;   // the maximum the exponentials are taken against (log2 domain, scaled)
      float l = 0.f;           // running row sum of P
      bool pub_prev = true;    // P(j-1) already published (warp-uniform); nothing to publish before tile 0

      // -inf for the columns of K/V tile j this row must not see (tail of the sequence, causal mask)
      auto mask_tile = [&](uint32_t (&s)[BKV], int j) {
        const int col0 = j * BKV;
        int lim = n_tok - col0;
        if (causal) lim = min(lim, qrow - col0 + 1);
        if (!__all_sync(0xffffffffu, lim >= BKV)) {
#pragma unroll
          for (int i = 0; i < BKV; ++i)
            if (i >= lim) s[i] = 0xff800000u;
        }
      };
      auto load_s = [&](uint32_t (&s)[BKV], int j) {        // asynchronous: complete after tmem_ld_wait()
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        ptx::tmem_ld_32x32b_x32(s_tmem + (j & 1) * BKV, s0);
        ptx::tmem_ld_32x32b_x32(s_tmem + (j & 1) * BKV + 32, s1);
      };

      // Two warps share each SM sub-partition's MUFU: Q tile 1 starts half an exponential phase late so that the
      // (few) non-MUFU stretches of one tile fall into the other's exponentials.
      if (STAGGER && t == 1) ptx::mbar_wait(stagger, 0);

      // One K/V tile. cur = S(j) (already masked, row maximum folded into m_used); nxt receives S(j+1).
      auto iteration = [&](uint32_t (&cur)[BKV], uint32_t (&nxt)[BKV], int j) {
        const bool have_next = j + 1 < n_kv;
        if (have_next) {
          ptx::mbar_wait(&s_full[2 * t + ((j + 1) & 1)], ((j + 1) >> 1) & 1);
          ptx::tc_fence_after();
          load_s(nxt, j + 1);
        }
        float rs[4] = {0.f, 0.f, 0.f, 0.f};   // independent partial row sums
        uint32_t w[BKV / 2];                  // P(j) packed to 16 bit
        // ---- first half of the exponentials (runs under the TMEM load)
#pragma unroll
        for (int i = 0; i < BKV / 4; ++i) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(cur[2 * i]), scale_log2, -m_used));
          const float p1 = ex2_approx(fmaf(__uint_as_float(cur[2 * i + 1]), scale_log2, -m_used));
          rs[(2 * i) & 3] += p0;
          rs[(2 * i + 1) & 3] += p1;
          w[i] = ptx::pack16<BF16>(p0, p1);
        }
        // ---- the one synchronisation point of the iteration
        if (have_next) ptx::tmem_ld_wait();
        if (!pub_prev) ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (have_next) ptx::mbar_arrive(&s_free[2 * t + ((j + 1) & 1)]);
          if (!pub_prev) ptx::mbar_arrive(&p_full[2 * t + ((j - 1) & 1)]);
          if (STAGGER && t == 0 && j == 0) ptx::mbar_arrive(stagger);
        }
        pub_prev = false;
        // ---- second half of the exponentials, interleaved with the row maximum of S(j+1)
        float mxp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (have_next) mask_tile(nxt, j + 1);
#pragma unroll
        for (int i = BKV / 4; i < BKV / 2; ++i) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(cur[2 * i]), scale_log2, -m_used));
          const float p1 = ex2_approx(fmaf(__uint_as_float(cur[2 * i + 1]), scale_log2, -m_used));
          rs[(2 * i) & 3] += p0;
          rs[(2 * i + 1) & 3] += p1;
          w[i] = ptx::pack16<BF16>(p0, p1);
          if (have_next) {
            const int c = (i - BKV / 4) * 4;   // 4 columns of S(j+1) per step: two FMNMX3 into alternating chains
            mxp[i & 1] = max3(mxp[i & 1], __uint_as_float(nxt[c]), __uint_as_float(nxt[c + 1]));
            mxp[2 + (i & 1)] = max3(mxp[2 + (i & 1)], __uint_as_float(nxt[c + 2]), __uint_as_float(nxt[c + 3]));
          }
        }
        l += (rs[0] + rs[1]) + (rs[2] + rs[3]);
        // ---- P(j) -> its TMEM buffer (still being read by P V of tile j-2, issued more than an iteration ago)
        if (j >= 2) {
          ptx::mbar_wait(&p_free[2 * t + (j & 1)], ((j - 2) >> 1) & 1);
          ptx::tc_fence_after();
        }
        ptx::tmem_st_32x32b_x32(p_tmem + (j & 1) * 32, w);
        // ---- lazy rescale for the next tile: only when its maximum ran away from the one in use by more than 2^8
        if (have_next) {
          const float m_tile = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3])) * scale_log2;
          const bool need = m_tile > m_used + RESCALE_THRESHOLD;
          if (__any_sync(0xffffffffu, need)) {
            // O_t must be quiescent: publish P(j) now and wait until its P V product has drained
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p_full[2 * t + (j & 1)]);
            pub_prev = true;
            ptx::mbar_wait(&p_free[2 * t + (j & 1)], (j >> 1) & 1);
            ptx::tc_fence_after();
            const float m_new = need ? m_tile : m_used;
            const float alpha = ex2_approx(m_used - m_new);   // 1 for rows that keep their maximum
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
              uint32_t v[16];
              ptx::tmem_ld_32x32b_x16(o_tmem + c * 16, v);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              ptx::tmem_st_32x32b_x16(o_tmem + c * 16, v);
            }
            ptx::tmem_st_wait();
            l *= alpha;
            m_used = m_new;
          }
        }
      };

      // ---- prologue: S(0) into registers, its maximum becomes the first reference
      uint32_t sa[BKV], sb[BKV];
      ptx::mbar_wait(&s_full[2 * t + 0], 0);
      ptx::tc_fence_after();
      load_s(sa, 0);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s_free[2 * t + 0]);
      mask_tile(sa, 0);
      {
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < BKV; i += 2) mx = max3(mx, __uint_as_float(sa[i]), __uint_as_float(sa[i + 1]));
        m_used = mx * scale_log2;        // column 0 is always visible, so this is finite
      }
      for (int j = 0; j < n_kv; j += 2) {
        iteration(sa, sb, j);
        if (j + 1 < n_kv) iteration(sb, sa, j + 1);
      }

      // ---- publish the last P, wait for all P V products, normalise and store
      if (!pub_prev) {
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[2 * t + ((n_kv - 1) & 1)]);
      }
      ptx::mbar_wait(&p_free[2 * t + ((n_kv - 1) & 1)], ((n_kv - 1) >> 1) & 1);
      ptx::tc_fence_after();
      const float inv_l = 1.0f / l;
      uint16_t* dst_row = out + (static_cast<size_t>(b) * n_tok + qrow) * D + h * HD;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(o_tmem + c * 32, v);
        ptx::tmem_ld_wait();
        if (qrow < n_tok) {
          uint4* dst = reinterpret_cast<uint4*>(dst_row + c * 32);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 wv;
            wv.x = ptx::pack16<BF16>(__uint_as_float(v[8 * u + 0]) * inv_l, __uint_as_float(v[8 * u + 1]) * inv_l);
            wv.y = ptx::pack16<BF16>(__uint_as_float(v[8 * u + 2]) * inv_l, __uint_as_float(v[8 * u + 3]) * inv_l);
            wv.z = ptx::pack16<BF16>(__uint_as_float(v[8 * u + 4]) * inv_l, __uint_as_float(v[8 * u + 5]) * inv_l);
            wv.w = ptx::pack16<BF16>(__uint_as_float(v[8 * u + 6]) * inv_l, __uint_as_float(v[8 * u + 7]) * inv_l);
            dst[u] = wv;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace aaclip
