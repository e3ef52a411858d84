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
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace aaclip {

namespace attn_cfg {
constexpr int BQ = 128, QT = 2, BKV = 128, HD = 64, KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * HD * 2;          // one Q / K / V tile: 128 rows x 128 B = 16 KiB
constexpr int THREADS = 384;                      // 3 warpgroups: softmax tile 0, softmax tile 1, {TMA, MMA, 2 idle}
constexpr int TMEM_COLS = 512;
constexpr int TM_S = 0, TM_O = 2 * BKV, TM_P = TM_O + 2 * HD;   // TMEM column map
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + QT * TILE_BYTES;
constexpr int OFF_V = OFF_K + KV_STAGES * TILE_BYTES;
constexpr int OFF_BAR = OFF_V + KV_STAGES * TILE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr float RESCALE_THRESHOLD = 8.0f;         // log2 units: P never exceeds 2^8
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
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, uint16_t* __restrict__ out, int n_tok,
                         int D, int causal, float scale_log2) {
  using namespace attn_cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* k_full = q_full + 1;                  // [KV_STAGES] TMA -> MMA
  uint64_t* k_empty = k_full + KV_STAGES;         // [KV_STAGES] MMA -> TMA : every S MMA reading K_j has retired
  uint64_t* v_full = k_empty + KV_STAGES;         // [KV_STAGES]
  uint64_t* v_empty = v_full + KV_STAGES;         // [KV_STAGES] MMA -> TMA : every P V MMA reading V_j has retired
  uint64_t* s_full = v_empty + KV_STAGES;         // [QT] MMA -> softmax : S_t(j) complete
  uint64_t* s_free = s_full + QT;                 // [QT] softmax -> MMA : S_t(j) is in registers
  uint64_t* p_full = s_free + QT;                 // [QT] softmax -> MMA : P_t(j) in TMEM, O_t rescaled if needed
  uint64_t* o_full = p_full + QT;                 // [QT] MMA -> softmax : O_t += P_t(j) V_j complete
  uint64_t* stagger = o_full + QT;                // softmax WG0 -> WG1, once: start half an exponential phase later
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(stagger + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_base = blockIdx.x * (BQ * QT);
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_kv_tiles = (n_tok + BKV - 1) / BKV;
  const int n_qt = (q_base + BQ < n_tok) ? 2 : 1;  // second Q tile entirely out of range -> skipped

  if (warp == 8 && lane == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&k_empty[s], 2);   // one tcgen05.commit per MMA issuer (twice from issuer 0 if n_qt == 1)
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&v_empty[s], 2);
    }
    for (int t = 0; t < QT; ++t) {
      ptx::mbar_init(&s_full[t], 1);
      ptx::mbar_init(&s_free[t], 4);
      ptx::mbar_init(&p_full[t], 4);
      ptx::mbar_init(&o_full[t], 1);
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

  // The softmax warpgroups need a whole S row in registers: warpgroup 2 (TMA, MMA, 2 idle warps) hands its
  // registers over with setmaxnreg (issued inside each role's branch so ptxas allocates per role).
  if (warp == 11) {
    ptx::reg_dec<40>();
  } else if (warp == 8) {
    ptx::reg_dec<40>();
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
  } else if (warp == 9 || warp == 10) {
    ptx::reg_dec<40>();
    // ===================== MMA issuers (one warp per Q tile; warp-uniform loop, one elected lane issues) ===========
    const int t = warp - 9;
    if (t < n_qt) {
      constexpr uint32_t idesc_s = ptx::make_idesc_16(BQ, BKV, 0, BF16);   // S = Q K^T : both operands K-major
      constexpr uint32_t idesc_pv = ptx::make_idesc_16(BQ, HD, 1, BF16);   // O += P V  : V is MN-major
      const uint32_t q_addr = ptx::smem_u32(smem + OFF_Q) + t * TILE_BYTES;
      const uint32_t k_addr = ptx::smem_u32(smem + OFF_K);
      const uint32_t v_addr = ptx::smem_u32(smem + OFF_V);
      const bool alone = (n_qt == 1);   // then this thread supplies both arrivals of k_empty / v_empty

      auto issue_s = [&](int stage) {
        const uint64_t a = ptx::make_sw128_desc(q_addr);
        const uint64_t bd = ptx::make_sw128_desc(k_addr + stage * TILE_BYTES);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
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
          for (int k = 0; k < BKV / 16; ++k) {
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
    ptx::reg_inc<216>();
    // ===================== softmax warps =====================
    const int t = warp >> 2;
    const int q = warp & 3;
    if (t < n_qt) {
      const int r = q * 32 + lane;              // row inside the Q tile == TMEM lane
      const int qrow = q_base + t * BQ + r;     // token index of this query
      const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
      const uint32_t s_tmem = tmem_base + lane_off + TM_S + t * BKV;
      const uint32_t o_tmem = tmem_base + lane_off + TM_O + t * HD;
      const uint32_t p_tmem = tmem_base + lane_off + TM_P + t * (BKV / 2);

      float m_used = -1e30f;   // the maximum the exponentials are taken against (log2 domain, scaled)
      float l = 0.f;           // running row sum of P

      // Two warps share each SM sub-partition's MUFU (4 ex2/clk): if both Q tiles run in lock-step, the MUFU idles
      // while both are in their load/max/wait phases and is oversubscribed while both exponentiate. Q tile 1 starts
      // half an exponential phase late, so one tile's exponentials cover the other's non-MUFU phases.
      if (STAGGER && t == 1) ptx::mbar_wait(stagger, 0);

      for (int j = 0; j < n_kv_tiles; ++j) {
        ptx::mbar_wait(&s_full[t], j & 1);
        ptx::tc_fence_after();
        uint32_t s[BKV];
        {
          uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
          uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
          uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
          uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
          ptx::tmem_ld_32x32b_x32(s_tmem + 0, s0);
          ptx::tmem_ld_32x32b_x32(s_tmem + 32, s1);
          ptx::tmem_ld_32x32b_x32(s_tmem + 64, s2);
          ptx::tmem_ld_32x32b_x32(s_tmem + 96, s3);
          ptx::tmem_ld_wait();
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&s_free[t]);

        const int col0 = j * BKV;
        int lim = n_tok - col0;                       // number of valid columns in this tile
        if (causal) lim = min(lim, qrow - col0 + 1);
        const bool full_tile = __all_sync(0xffffffffu, lim >= BKV);
        if (!full_tile) {
#pragma unroll
          for (int i = 0; i < BKV; ++i)
            if (i >= lim) s[i] = 0xff800000u;   // -inf: exp2 gives 0, contributes nothing
        }

        // ---- row maximum: 8 independent FMNMX3 chains (one 128-long dependent chain costs ~500 cycles)
        float mxp[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          mxp[c] = max3(__uint_as_float(s[c * 16]), __uint_as_float(s[c * 16 + 1]), __uint_as_float(s[c * 16 + 2]));
#pragma unroll
          for (int i = 3; i < 15; i += 2)
            mxp[c] = max3(mxp[c], __uint_as_float(s[c * 16 + i]), __uint_as_float(s[c * 16 + i + 1]));
          mxp[c] = fmaxf(mxp[c], __uint_as_float(s[c * 16 + 15]));
        }
        const float mx = max3(max3(mxp[0], mxp[1], mxp[2]), max3(mxp[3], mxp[4], mxp[5]), fmaxf(mxp[6], mxp[7]));
        const float m_tile = mx * scale_log2;

        // ---- P_t (TMEM) and O_t are still being read / accumulated by the P V product of the previous tile
        if (j > 0) {
          ptx::mbar_wait(&o_full[t], (j - 1) & 1);
          ptx::tc_fence_after();
        }

        // ---- lazy rescale: only when the true maximum ran away from the one in use by more than 2^8
        const bool need = m_tile > m_used + RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? m_tile : m_used;
          const float alpha = ex2_approx(m_used - m_new);   // 1 for rows that keep their maximum
          if (j > 0) {
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
        float rs[4] = {0.f, 0.f, 0.f, 0.f};   // independent partial sums (no 128-long FADD chain)
#pragma unroll
        for (int c = 0; c < 2; ++c) {          // 64 P values -> 32 packed columns per store
          uint32_t w[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(s[c * 64 + 2 * i]), scale_log2, -m_used));
            const float p1 = ex2_approx(fmaf(__uint_as_float(s[c * 64 + 2 * i + 1]), scale_log2, -m_used));
            rs[(2 * i) & 3] += p0;
            rs[(2 * i + 1) & 3] += p1;
            w[i] = ptx::pack16<BF16>(p0, p1);
          }
          ptx::tmem_st_32x32b_x32(p_tmem + c * 32, w);
          if (STAGGER && t == 0 && j == 0 && c == 0) {
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(stagger);
          }
        }
        l += (rs[0] + rs[1]) + (rs[2] + rs[3]);

        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[t]);
      }

      // ---- all P V accumulated: normalise and store
      ptx::mbar_wait(&o_full[t], (n_kv_tiles - 1) & 1);
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
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace aaclip
