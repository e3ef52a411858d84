// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   A : fp16|bf16 row-major [M,K] (activations, token-major)         -- TMA, SWIZZLE_128B, box 64(K) x 128(M)
//   W : fp16|bf16 row-major [N,K] (nn.Linear weight layout, K-major) -- TMA, SWIZZLE_128B, box 64(K) x 256(N)
//   D : fp32 accumulator in TMEM, 128 lanes x 256 columns, double-buffered (512 columns)
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane) + TMEM owner,
// warps 2..9 = epilogue (TMEM -> registers -> fused math -> global). The epilogue of tile i overlaps the
// main loop of tile i+1 through the second TMEM accumulator stage.
//
// Replaces, on the reference's path: F.linear / addmm calls inside nn.MultiheadAttention
// (model/transformer.py:200,237), the MLP (model/transformer.py:211-219), SimpleAdapter / SimpleProj
// (model/adapter_modules.py:6-26) and the patch-embedding conv (model/adapter.py:139).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace aaclip {

enum EpiMode : int {
  EPI_BIAS_BF16 = 0,       // out_16 = acc + bias   ("BF16" in these names = the 16-bit operand dtype)
  EPI_BIAS_GELU_BF16 = 1,  // out_bf16 = gelu_erf(acc + bias)          (nn.GELU, model/model.py:84)
  EPI_BIAS_QGELU_BF16 = 2, // out_bf16 = x*sigmoid(1.702x), x=acc+bias (QuickGELU, model/transformer.py:46-49)
  EPI_BIAS_RESID_F32 = 3,  // out_f32 += acc + bias ; optional bf16 copy of the new value
  EPI_F32 = 4,             // out_f32 = acc
  EPI_LEAKY_F32 = 5,       // out_f32 = leaky_relu(acc, 0.01)          (nn.LeakyReLU default slope)
  EPI_PATCH_F32 = 6,       // out_f32[b*(P+1)+1+p, :] = acc + pos[p+1, :]   (m = b*P + p)
  EPI_BIAS_F32 = 7,        // out_f32 = acc + bias
};

struct GemmEpi {
  const float* bias = nullptr;        // [N]
  float* out_f32 = nullptr;           // [*, ldo]
  uint16_t* out_16 = nullptr;         // [*, ldo] fp16 or bf16 (the operand dtype)
  const float* pos = nullptr;         // EPI_PATCH_F32: positional embedding [(P+1), N]
  int ldo = 0;                        // leading dimension of the outputs, in elements
  int patches = 0;                    // EPI_PATCH_F32: P
};

namespace gemm_cfg {
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;             // 16 KiB
constexpr int B_BYTES = BN * BK * 2;             // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 48 KiB
constexpr int THREADS = 320;                     // 10 warps
constexpr int EPI_WARPS = 8;
constexpr int TMEM_COLS = 512;                   // 2 accumulator stages x 256 columns
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
}  // namespace gemm_cfg

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float quick_gelu(float x) { return x / (1.0f + __expf(-1.702f * x)); }

template <int EPI, bool BF16>
__device__ __forceinline__ void epilogue_store(const uint32_t (&r)[32], int row, int n0, int M, const GemmEpi& e) {
  if (row >= M) return;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);

  if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_QGELU_BF16 ||
                EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_F32) {
    const float4* b4 = reinterpret_cast<const float4*>(e.bias + n0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = __ldg(b4 + i);
      v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
  }
  if constexpr (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  }
  if constexpr (EPI == EPI_BIAS_QGELU_BF16) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = quick_gelu(v[i]);
  }
  if constexpr (EPI == EPI_LEAKY_F32) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : 0.01f * v[i];
  }

  if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_QGELU_BF16) {
    uint4* o = reinterpret_cast<uint4*>(e.out_16 + static_cast<size_t>(row) * e.ldo + n0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = ptx::pack16<BF16>(v[8 * i + 0], v[8 * i + 1]);
      u.y = ptx::pack16<BF16>(v[8 * i + 2], v[8 * i + 3]);
      u.z = ptx::pack16<BF16>(v[8 * i + 4], v[8 * i + 5]);
      u.w = ptx::pack16<BF16>(v[8 * i + 6], v[8 * i + 7]);
      o[i] = u;
    }
  } else if constexpr (EPI == EPI_BIAS_RESID_F32) {
    float4* o = reinterpret_cast<float4*>(e.out_f32 + static_cast<size_t>(row) * e.ldo + n0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 x = o[i];
      x.x += v[4 * i + 0]; x.y += v[4 * i + 1]; x.z += v[4 * i + 2]; x.w += v[4 * i + 3];
      o[i] = x;
      v[4 * i + 0] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
    }
    if (e.out_16 != nullptr) {
      uint4* ob = reinterpret_cast<uint4*>(e.out_16 + static_cast<size_t>(row) * e.ldo + n0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = ptx::pack16<BF16>(v[8 * i + 0], v[8 * i + 1]);
        u.y = ptx::pack16<BF16>(v[8 * i + 2], v[8 * i + 3]);
        u.z = ptx::pack16<BF16>(v[8 * i + 4], v[8 * i + 5]);
        u.w = ptx::pack16<BF16>(v[8 * i + 6], v[8 * i + 7]);
        ob[i] = u;
      }
    }
  } else {
    size_t orow = static_cast<size_t>(row);
    if constexpr (EPI == EPI_PATCH_F32) {
      const int b = row / e.patches, p = row - b * e.patches;
      orow = static_cast<size_t>(b) * (e.patches + 1) + 1 + p;
      const float4* pe = reinterpret_cast<const float4*>(e.pos + static_cast<size_t>(p + 1) * e.ldo + n0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 q = __ldg(pe + i);
        v[4 * i + 0] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
      }
    }
    float4* o = reinterpret_cast<float4*>(e.out_f32 + orow * e.ldo + n0);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i + 0], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
}

template <int EPI, bool BF16>
__global__ void __launch_bounds__(gemm_cfg::THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M,
                    int N, int K, GemmEpi epi) {
  using namespace gemm_cfg;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = N / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = K / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full_bar[s], 1);
      ptx::mbar_init(&tmem_empty_bar[s], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<TMEM_COLS>(tmem_base_smem);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
          ptx::tma_load_2d(sa + A_BYTES, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_16(BM, BN, 0, BF16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_desc = ptx::make_sw128_desc(sa);
          const uint64_t b_desc = ptx::make_sw128_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128-B swizzle row: +2 in the (addr>>4) field
            ptx::tc_mma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::tc_commit(&empty_bar[stage]);                       // frees the smem slot when the MMAs retire
          if (kb == num_kb - 1) ptx::tc_commit(&tmem_full_bar[acc]);  // accumulator complete
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int e = warp - 2;
    const int q = warp & 3;        // TMEM lane quarter this warp may access
    const int half = e >> 2;       // which 128 of the 256 accumulator columns
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + half * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(t_base + c * 32, r);
        ptx::tmem_ld_wait();
        if (c == 3) {
          // all of this warp's TMEM reads for the tile are done: hand the accumulator stage back
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
        }
        epilogue_store<EPI, BF16>(r, row, n_blk * BN + half * 128 + c * 32, M, epi);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace aaclip
