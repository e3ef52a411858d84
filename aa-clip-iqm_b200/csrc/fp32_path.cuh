// fp32-accuracy mode (BASELINE config 4: "ViT-L-14-336 @336 batch 128 fp32-accuracy mode, tolerance check").
//
// The reference computes everything in fp32; the tensor-core path rounds GEMM / attention operands to
// 16 bit and lands within 1e-2 of it. This mode exists to show that the REST of the pipeline (data layout, fused
// epilogues, LayerNorm, adapters, scoring) is exact: the same launch sequence with plain fp32 FFMA kernels in place
// of the tcgen05 ones reproduces the reference maps to <= 1e-4. It is a verification mode, not a fast path:
// fp32 has no tensor-core format on sm_100a (kind::tf32 keeps 10 mantissa bits), so these are classic SIMT kernels --
// a 128x128x16 register-tiled SGEMM and a thread-per-query-row flash attention.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "gemm_tcgen05.cuh"   // EpiMode, GemmEpi

namespace aaclip {

namespace sgemm_cfg {
constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256, PAD = 4;
}

__device__ __forceinline__ float gelu_erf_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// C[M,N] = A[M,K] W[N,K]^T (+ epilogue), all fp32, A and W K-major. N % 128 == 0, K % 16 == 0, any M.
// 256 threads, each an 8x8 block of the 128x128 tile; operands staged transposed ([k][m]) in shared memory so the
// inner product reads two float4 per operand per k. Sequential-k fp32 FMA accumulation (what the CPU reference does,
// up to summation order).
template <int EPI>
__global__ void __launch_bounds__(sgemm_cfg::THREADS)
sgemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ W, int M, int N, int K, GemmEpi epi) {
  using namespace sgemm_cfg;
  __shared__ float As[2][BK][BM + PAD];
  __shared__ float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int ty = tid >> 4, tx = tid & 15;           // 16 x 16 threads, 8 x 8 outputs each
  // global -> smem: 128 rows x 16 k per operand = 512 float4; thread t loads rows (t >> 2) and (t >> 2) + 64, k-quad (t & 3)
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto load_tile = [&](int kb, float4 (&ra)[2], float4 (&rb)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int am = m0 + lr + h * 64;
      ra[h] = (am < M) ? *reinterpret_cast<const float4*>(A + static_cast<size_t>(am) * K + kb * BK + lk)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[h] = *reinterpret_cast<const float4*>(W + static_cast<size_t>(n0 + lr + h * 64) * K + kb * BK + lk);
    }
  };
  auto store_tile = [&](int buf, const float4 (&ra)[2], const float4 (&rb)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + h * 64;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y; Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
    }
  };

  const int nkb = K / BK;
  float4 ra[2], rb[2];
  load_tile(0, ra, rb);
  store_tile(0, ra, rb);
  __syncthreads();
  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nkb) load_tile(kb + 1, ra, rb);     // global loads of the next tile fly under this tile's FMAs
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nkb) {
      store_tile(buf ^ 1, ra, rb);
      __syncthreads();
    }
  }

  // ---- epilogue (same semantics as the tensor-core epilogues, exact erff GELU)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + ty * 8 + i;
    if (row >= M) continue;
#pragma unroll
    for (int jv = 0; jv < 2; ++jv) {
      const int n = n0 + tx * 8 + jv * 4;
      float4 v = make_float4(acc[i][jv * 4], acc[i][jv * 4 + 1], acc[i][jv * 4 + 2], acc[i][jv * 4 + 3]);
      if constexpr (EPI == EPI_BIAS_F32 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_QGELU_BF16 || EPI == EPI_BIAS_RESID_F32) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(epi.bias + n));
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
      }
      if constexpr (EPI == EPI_BIAS_GELU_BF16) {
        v.x = gelu_erf_exact(v.x); v.y = gelu_erf_exact(v.y); v.z = gelu_erf_exact(v.z); v.w = gelu_erf_exact(v.w);
      }
      if constexpr (EPI == EPI_BIAS_QGELU_BF16) {
        v.x = v.x / (1.0f + expf(-1.702f * v.x)); v.y = v.y / (1.0f + expf(-1.702f * v.y));
        v.z = v.z / (1.0f + expf(-1.702f * v.z)); v.w = v.w / (1.0f + expf(-1.702f * v.w));
      }
      if constexpr (EPI == EPI_LEAKY_F32) {
        v.x = v.x > 0.f ? v.x : 0.01f * v.x; v.y = v.y > 0.f ? v.y : 0.01f * v.y;
        v.z = v.z > 0.f ? v.z : 0.01f * v.z; v.w = v.w > 0.f ? v.w : 0.01f * v.w;
      }
      size_t orow = static_cast<size_t>(row);
      if constexpr (EPI == EPI_PATCH_F32) {
        const int bi = row / epi.patches, p = row - bi * epi.patches;
        orow = static_cast<size_t>(bi) * (epi.patches + 1) + 1 + p;
        const float4 q = __ldg(reinterpret_cast<const float4*>(epi.pos + static_cast<size_t>(p + 1) * epi.ldo + n));
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      float4* o = reinterpret_cast<float4*>(epi.out_f32 + orow * epi.ldo + n);
      if constexpr (EPI == EPI_BIAS_RESID_F32) {
        const float4 x = *o;
        v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
      }
      *o = v;
    }
  }
}

// fp32 flash attention, one thread per query row: qkv fp32 [B, N, 3D] -> out fp32 [B, N, D]; head dim 64.
// grid = (ceil(N / 128), H, B), block = 128. K / V stream through shared memory in 32-key tiles (broadcast reads);
// scores are taken 8 keys at a time so the running maximum moves (and O is rescaled) at most once per 8 keys.
// q is pre-scaled by 1/8 exactly as nn.MultiheadAttention does (q * head_dim^-0.5 before q k^T).
__global__ void __launch_bounds__(128)
attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int n_tok, int D, int causal) {
  constexpr int HD = 64, KT = 32;
  __shared__ float4 ks[KT][HD / 4];
  __shared__ float4 vs[KT][HD / 4];
  const int b = blockIdx.z, h = blockIdx.y;
  const int row = blockIdx.x * 128 + threadIdx.x;
  const bool valid = row < n_tok;
  const float* base = qkv + static_cast<size_t>(b) * n_tok * 3 * D;
  float q[HD], o[HD];
  {
    const float4* qp = reinterpret_cast<const float4*>(base + static_cast<size_t>(valid ? row : 0) * 3 * D + h * HD);
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) {
      const float4 t = qp[i];
      q[4 * i] = t.x * 0.125f; q[4 * i + 1] = t.y * 0.125f; q[4 * i + 2] = t.z * 0.125f; q[4 * i + 3] = t.w * 0.125f;
    }
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const int n_tiles = (n_tok + KT - 1) / KT;
  for (int tile = 0; tile < n_tiles; ++tile) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * (HD / 4); i += 128) {
      const int kr = i / (HD / 4), c = i - kr * (HD / 4);
      const int key = tile * KT + kr;
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (key < n_tok) {
        const float* kp = base + static_cast<size_t>(key) * 3 * D + D + h * HD;
        kk = reinterpret_cast<const float4*>(kp)[c];
        vv = reinterpret_cast<const float4*>(kp + D)[c];
      }
      ks[kr][c] = kk;
      vs[kr][c] = vv;
    }
    __syncthreads();
#pragma unroll 1
    for (int g = 0; g < KT / 8; ++g) {
      float s[8];
      float gm = -INFINITY;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int kr = g * 8 + u, key = tile * KT + kr;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) {
          const float4 kk = ks[kr][c];
          acc = fmaf(q[4 * c], kk.x, acc); acc = fmaf(q[4 * c + 1], kk.y, acc);
          acc = fmaf(q[4 * c + 2], kk.z, acc); acc = fmaf(q[4 * c + 3], kk.w, acc);
        }
        const bool vis = key < n_tok && (!causal || key <= row);
        s[u] = vis ? acc : -INFINITY;
        gm = fmaxf(gm, s[u]);
      }
      if (gm == -INFINITY) continue;             // nothing visible in this group for this row
      const float m_new = fmaxf(m, gm);
      if (m_new > m) {
        const float alpha = expf(m - m_new);     // exp(-inf) = 0 on the first visible group
        l *= alpha;
#pragma unroll
        for (int i = 0; i < HD; ++i) o[i] *= alpha;
        m = m_new;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float p = expf(s[u] - m);          // masked keys: exp(-inf) = 0
        l += p;
        const int kr = g * 8 + u;
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) {
          const float4 vv = vs[kr][c];
          o[4 * c] = fmaf(p, vv.x, o[4 * c]); o[4 * c + 1] = fmaf(p, vv.y, o[4 * c + 1]);
          o[4 * c + 2] = fmaf(p, vv.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(p, vv.w, o[4 * c + 3]);
        }
      }
    }
  }
  if (valid) {
    const float inv = 1.0f / l;
    float4* op = reinterpret_cast<float4*>(out + (static_cast<size_t>(b) * n_tok + row) * D + h * HD);
#pragma unroll
    for (int i = 0; i < HD / 4; ++i)
      op[i] = make_float4(o[4 * i] * inv, o[4 * i + 1] * inv, o[4 * i + 2] * inv, o[4 * i + 3] * inv);
  }
}

}  // namespace aaclip
