// Fused flash-style multi-head attention forward for sm_100a (tcgen05 + TMEM + TMA), head_dim 64.
//
// Replaces nn.MultiheadAttention's slow path on the reference (model/transformer.py:200,226-237 ->
// F.multi_head_attention_forward): q*(1/8), bmm, softmax, bmm, plus the head-averaged weights that
// AdaptedCLIP.forward discards (model/adapter.py:162). Nothing N x N ever reaches HBM here.
//
// Input : packed QKV bf16 [B, N, 3*D] (the QKV GEMM output, q | k | v along the last axis, head h at h*64)
// Output: O bf16 [B, N, D] (heads concatenated) -- directly the A operand of the out-projection GEMM.
//
// One CTA = (image b, head h, 256 query rows) = two 128-row Q tiles that share every K/V tile:
//   warp 8      : TMA producer (Q once; K_j/V_j through a 3-stage ring)
//   warp 9      : tcgen05.mma issuer (one lane) + TMEM owner
//   warps 0..3  : softmax + O accumulation for Q tile 0 (thread = one query row; lane quarter = warp % 4)
//   warps 4..7  : same for Q tile 1
// TMEM (512 cols): S_t [128 x 128] fp32 at t*128, PV_t [128 x 64] fp32 at 256 + t*64.
// Per K/V tile j and Q tile t:  S_t = Q_t K_j^T (SS MMA) -> softmax warps read S_t twice from TMEM (row max,
// then exp2) and write P_t (bf16) into a 128B-swizzled K-major smem tile -> PV_t = P_t V_j (V is the MN-major
// B operand, straight from the TMA tile) -> the owning threads fold PV_t into their fp32 O registers with the
// usual online-softmax rescale. While one Q tile is in its softmax phase the tensor core runs the other tile's
// S / PV MMAs (ping-pong), so MUFU and tensor pipes overlap.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace aaclip {

namespace attn_cfg {
constexpr int BQ = 128, QT = 2, BKV = 128, HD = 64, KV_STAGES = 3;
constexpr int Q_BYTES = BQ * HD * 2;         // 16 KiB per Q tile
constexpr int KV_STAGE_BYTES = 2 * BKV * HD * 2;  // K tile + V tile = 32 KiB
constexpr int P_BYTES = BQ * BKV * 2;        // 32 KiB per Q tile
constexpr int THREADS = 320;
constexpr int TMEM_COLS = 512;
constexpr int OFF_Q = 0;
constexpr int OFF_KV = QT * Q_BYTES;
constexpr int OFF_P = OFF_KV + KV_STAGES * KV_STAGE_BYTES;
constexpr int OFF_BAR = OFF_P + QT * P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
}  // namespace attn_cfg

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
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
  uint64_t* kv_full = q_full + 1;             // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES;   // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;    // [QT]
  uint64_t* p_full = s_full + QT;             // [QT]
  uint64_t* o_full = p_full + QT;             // [QT]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(o_full + QT);

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
      ptx::mbar_init(&kv_full[s], 1);
      ptx::mbar_init(&kv_empty[s], 1);
    }
    for (int t = 0; t < QT; ++t) {
      ptx::mbar_init(&s_full[t], 1);
      ptx::mbar_init(&p_full[t], 4);
      ptx::mbar_init(&o_full[t], 1);
    }
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

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, QT * Q_BYTES);
      for (int t = 0; t < QT; ++t)
        ptx::tma_load_3d(smem + OFF_Q + t * Q_BYTES, &tmap_qkv, q_full, h * HD, q_base + t * BQ, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv_tiles; ++j) {
        ptx::mbar_wait(&kv_empty[stage], phase ^ 1);
        ptx::mbar_arrive_expect_tx(&kv_full[stage], KV_STAGE_BYTES);
        uint8_t* sk = smem + OFF_KV + stage * KV_STAGE_BYTES;
        ptx::tma_load_3d(sk, &tmap_qkv, &kv_full[stage], D + h * HD, j * BKV, b);
        ptx::tma_load_3d(sk + BKV * HD * 2, &tmap_qkv, &kv_full[stage], 2 * D + h * HD, j * BKV, b);
        if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_16(BQ, BKV, 0, BF16);   // S = Q K^T : both operands K-major
      constexpr uint32_t idesc_pv = ptx::make_idesc_16(BQ, HD, 1, BF16);   // PV = P V  : V is MN-major
      const uint32_t q_addr = ptx::smem_u32(smem + OFF_Q);
      const uint32_t kv_addr = ptx::smem_u32(smem + OFF_KV);
      const uint32_t p_addr = ptx::smem_u32(smem + OFF_P);

      auto issue_s = [&](int t, int stage) {
        const uint64_t a = ptx::make_sw128_desc(q_addr + t * Q_BYTES);
        const uint64_t bd = ptx::make_sw128_desc(kv_addr + stage * KV_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ptx::tc_mma_f16(tmem_base + t * BKV, a + 2 * k, bd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        ptx::tc_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int stage) {
        const uint32_t pa = p_addr + t * P_BYTES;
        const uint32_t va = kv_addr + stage * KV_STAGE_BYTES + BKV * HD * 2;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          // A: two 64-wide K chunks (16 KiB apart), 32 B per k-step inside a chunk.
          const uint64_t a = ptx::make_sw128_desc(pa + (k >> 2) * (BQ * 128) + (k & 3) * 32);
          // B (MN-major): 16 kv rows of 128 B per k-step.
          const uint64_t bd = ptx::make_sw128_desc(va + k * 16 * 128);
          ptx::tc_mma_f16(tmem_base + 2 * BKV + t * HD, a, bd, idesc_pv, k != 0 ? 1u : 0u);
        }
        ptx::tc_commit(&o_full[t]);
      };

      ptx::mbar_wait(q_full, 0);
      ptx::mbar_wait(&kv_full[0], 0);
      ptx::tc_fence_after();
      for (int t = 0; t < n_qt; ++t) issue_s(t, 0);

      int stage = 0;          // stage of K_j / V_j
      uint32_t phase = 0;
      for (int j = 0; j < n_kv_tiles; ++j) {
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == KV_STAGES) { nstage = 0; nphase ^= 1; }
        for (int t = 0; t < n_qt; ++t) {
          ptx::mbar_wait(&p_full[t], j & 1);   // P_t(j) in smem, S_t and PV_t drained by the softmax warps
          ptx::tc_fence_after();
          if (j + 1 < n_kv_tiles) {
            if (t == 0) {
              ptx::mbar_wait(&kv_full[nstage], nphase);
              ptx::tc_fence_after();
            }
            issue_s(t, nstage);
          }
          issue_pv(t, stage);
          if (t == n_qt - 1) ptx::tc_commit(&kv_empty[stage]);
        }
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // ===================== softmax / accumulate warps =====================
    const int t = warp >> 2;
    const int q = warp & 3;
    if (t < n_qt) {
      const int r = q * 32 + lane;              // row inside the Q tile == TMEM lane
      const int qrow = q_base + t * BQ + r;     // token index of this query
      const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
      const uint32_t s_tmem = tmem_base + lane_off + t * BKV;
      const uint32_t o_tmem = tmem_base + lane_off + 2 * BKV + t * HD;
      uint8_t* p_row = smem + OFF_P + t * P_BYTES + r * 128;
      const int rx = r & 7;

      float m = -1e30f, l = 0.f;
      float o[HD];
#pragma unroll
      for (int i = 0; i < HD; ++i) o[i] = 0.f;

      for (int j = 0; j < n_kv_tiles; ++j) {
        ptx::mbar_wait(&s_full[t], j & 1);
        ptx::tc_fence_after();
        const int col0 = j * BKV;
        int lim_all = n_tok - col0;                       // valid columns in this tile
        if (causal) lim_all = min(lim_all, qrow - col0 + 1);

        // ---- pass 1: row maximum
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < BKV / 32; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(s_tmem + c * 32, v);
          ptx::tmem_ld_wait();
          const int lim = lim_all - c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (i < lim) ? __uint_as_float(v[i]) : -INFINITY);
        }
        const float m_new = fmaxf(m, mx * scale_log2);
        const float alpha = ex2_approx(m - m_new);

        // ---- fold the previous tile's PV into O, then rescale to the new maximum
        if (j > 0) {
          ptx::mbar_wait(&o_full[t], (j - 1) & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32b_x32(o_tmem + c * 32, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[c * 32 + i] = (o[c * 32 + i] + __uint_as_float(v[i])) * alpha;
          }
        }
        l *= alpha;
        m = m_new;

        // ---- pass 2: P = exp2(S*scale - m) -> bf16 -> swizzled K-major smem tile
        float rowsum = 0.f;
#pragma unroll 1
        for (int c = 0; c < BKV / 32; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(s_tmem + c * 32, v);
          ptx::tmem_ld_wait();
          const int lim = lim_all - c * 32;
          float p[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = ex2_approx(fmaf(__uint_as_float(v[i]), scale_log2, -m_new));
            p[i] = (i < lim) ? e : 0.f;
            rowsum += p[i];
          }
          uint8_t* chunk = p_row + (c >> 1) * (BQ * 128);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 w;
            w.x = ptx::pack16<BF16>(p[8 * u + 0], p[8 * u + 1]);
            w.y = ptx::pack16<BF16>(p[8 * u + 2], p[8 * u + 3]);
            w.z = ptx::pack16<BF16>(p[8 * u + 4], p[8 * u + 5]);
            w.w = ptx::pack16<BF16>(p[8 * u + 6], p[8 * u + 7]);
            const int unit = ((c & 1) * 4 + u) ^ rx;
            *reinterpret_cast<uint4*>(chunk + unit * 16) = w;
          }
        }
        l += rowsum;

        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[t]);
      }

      // ---- last PV, normalise, store
      ptx::mbar_wait(&o_full[t], (n_kv_tiles - 1) & 1);
      ptx::tc_fence_after();
      const float inv_l = 1.0f / l;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(o_tmem + c * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = (o[c * 32 + i] + __uint_as_float(v[i])) * inv_l;
      }
      if (qrow < n_tok) {
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * n_tok + qrow) * D + h * HD);
#pragma unroll
        for (int u = 0; u < HD / 8; ++u) {
          uint4 w;
          w.x = ptx::pack16<BF16>(o[8 * u + 0], o[8 * u + 1]);
          w.y = ptx::pack16<BF16>(o[8 * u + 2], o[8 * u + 3]);
          w.z = ptx::pack16<BF16>(o[8 * u + 4], o[8 * u + 5]);
          w.w = ptx::pack16<BF16>(o[8 * u + 6], o[8 * u + 7]);
          dst[u] = w;
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
