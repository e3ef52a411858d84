// Memory-bound kernels of the AA-CLIP hot path (everything that is not a tensor-core contraction).
// All are coalesced / float4-vectorised, one warp per token row where a row reduction is needed, and have a
// fixed reduction order (batch-invariant: a row's result never depends on which other rows are in the batch).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <limits.h>
#include "ptx.cuh"

namespace aaclip {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------------------
// Patch gather ("im2col") for the 14x14/stride-14 patch-embedding conv (model/adapter.py:139,
// model/transformer.py:359-365):  images fp32 NCHW [B,3,S,S]  ->  A bf16 [B*G*G, KP]  with
// k = c*ps*ps + ky*ps + kx (the flattening order of conv1.weight[1024,3,14,14]); columns >= 3*ps*ps stay 0.
// grid = (G, B), block = 256: one CTA per row of patches. The 3*ps image rows of the patch row are converted into
// shared memory with coalesced reads, then written out as 32-bit pairs in the order of A: consecutive threads store
// consecutive words (the direct scatter wrote 28-byte runs 2 bytes at a time: 22 % of the HBM rate).
// Needs an even patch size and image size (pairs never straddle a run of kx; 8-byte aligned image rows) and
// 3*ps*W*2 bytes of shared memory.
template <bool BF16>
__global__ void __launch_bounds__(256)
patch_gather_kernel(const float* __restrict__ img, uint16_t* __restrict__ A, int S, int G, int ps, int KP) {
  extern __shared__ __align__(16) uint32_t rows32[];   // [3*ps][W/2] packed 16-bit pairs
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int py = blockIdx.x, b = blockIdx.y;
  const int W = G * ps;  // pixels actually covered by patches (== S when S % ps == 0)
  const int wp = W / 2, n_rows = 3 * ps;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // phase 1: a warp per image row (no integer division in the inner loop), 8-byte loads, one packed store per pair
  for (int cr = warp; cr < n_rows; cr += 8) {
    const int c = cr / ps, ky = cr - c * ps;
    const float2* src = reinterpret_cast<const float2*>(img + ((static_cast<size_t>(b) * 3 + c) * S + (py * ps + ky)) * S);
    uint32_t* dst = rows32 + cr * wp;
    for (int i = lane; i < wp; i += 32) {
      const float2 v = src[i];
      dst[i] = ptx::pack16<BF16>(v.x, v.y);
    }
  }
  __syncthreads();
  // phase 2: thread = one 32-bit word position of an A row (its source offset is fixed: computed once), loop over patches
  const int k_raw = 3 * ps * ps, wpr = KP / 2;         // 32-bit words per row of A
  uint32_t* out = reinterpret_cast<uint32_t*>(A + (static_cast<size_t>(b) * G + py) * G * KP);
  for (int w = threadIdx.x; w < wpr; w += blockDim.x) {
    const int k = 2 * w;
    if (k < k_raw) {
      const int cr = k / ps, kx = k - cr * ps;
      const uint32_t* src = rows32 + cr * wp + (kx >> 1);
      const int step = ps >> 1;
      for (int px = 0; px < G; ++px) out[px * wpr + w] = src[px * step];
    } else {
      for (int px = 0; px < G; ++px) out[px * wpr + w] = 0;   // K padding (the workspace may hold anything)
    }
  }
}

// general form of the same gather (odd patch size / width, or rows too large for shared memory)
template <bool BF16>
__global__ void patch_gather_simple_kernel(const float* __restrict__ img, uint16_t* __restrict__ A, int S, int G,
                                           int ps, int KP) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int py = blockIdx.x, b = blockIdx.y;
  const int W = G * ps;
  for (int cr = 0; cr < 3 * ps; ++cr) {
    const int c = cr / ps, ky = cr - c * ps;
    const float* src = img + ((static_cast<size_t>(b) * 3 + c) * S + (py * ps + ky)) * S;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
      const int px = x / ps, kx = x - px * ps;
      const size_t row = (static_cast<size_t>(b) * G + py) * G + px;
      A[row * KP + (c * ps + ky) * ps + kx] = ptx::cvt16<BF16>(src[x]);
    }
  }
  const int k_raw = 3 * ps * ps, pad = KP - k_raw;
  for (int i = threadIdx.x; i < G * pad; i += blockDim.x) {
    const int px = i / pad, k = k_raw + (i - px * pad);
    A[((static_cast<size_t>(b) * G + py) * G + px) * KP + k] = 0;
  }
}

// fp32-accuracy mode: same gather, fp32 out (columns >= 3*ps*ps zeroed)
__global__ void patch_gather_f32_kernel(const float* __restrict__ img, float* __restrict__ A, int S, int G, int ps,
                                        int KP) {
  const int py = blockIdx.x, b = blockIdx.y;
  const int W = G * ps;
  for (int cr = 0; cr < 3 * ps; ++cr) {
    const int c = cr / ps, ky = cr - c * ps;
    const float* src = img + ((static_cast<size_t>(b) * 3 + c) * S + (py * ps + ky)) * S;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
      const int px = x / ps, kx = x - px * ps;
      A[((static_cast<size_t>(b) * G + py) * G + px) * KP + (c * ps + ky) * ps + kx] = src[x];
    }
  }
  const int k_raw = 3 * ps * ps, pad = KP - k_raw;
  for (int i = threadIdx.x; i < G * pad; i += blockDim.x) {
    const int px = i / pad, k = k_raw + (i - px * pad);
    A[((static_cast<size_t>(b) * G + py) * G + px) * KP + k] = 0.f;
  }
}

// x[b*(P+1), :] = class_embedding + positional_embedding[0, :]      (model/adapter.py:143-153)
__global__ void cls_row_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                               int tokens, int D) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < D; i += blockDim.x) x[static_cast<size_t>(b) * tokens * D + i] = cls[i] + pos[i];
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm over D (eps 1e-5), fp32 statistics, one warp per row (model/transformer.py LayerNorm; ln_pre,
// ln_1, ln_2, ln_post, ln_final). Two-pass in registers: mean, then sum of squared deviations.
// skipP > 0: output row r = b*skipP + p reads input row b*(skipP+1) + 1 + p (drops the CLS row: x[1:], adapter.py:172).
// NORMALIZE=false degrades to a plain fp32 -> bf16 cast (operand staging for the adapter GEMM).
// stats_out != nullptr additionally emits (sum, sum of squares) of the OUTPUT row as partial 0 of stats_out[row][np][2]
// (partials 1..np-1 zeroed): the LayerNorm statistics the next GEMM's folded-LN epilogue needs when this kernel
// produces the residual stream (ln_pre).
// x and out_f32 may be the same buffer (ln_pre runs in place: each warp loads its whole row before it stores), so
// neither is declared __restrict__.
template <int D, bool NORMALIZE, bool BF16>
__global__ void layernorm_kernel(const float* x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, uint16_t* __restrict__ out_bf16,
                                 float* out_f32, int rows, int skipP, float eps,
                                 float* __restrict__ stats_out = nullptr, int np = 0,
                                 uint16_t* __restrict__ out_lo = nullptr) {
  // out_lo != nullptr: the result enters the residual stream in its (hi, lo) form: out_bf16 receives the hi plane (the
  // rounded 16-bit value = the GEMM operand) and out_lo the rounded remainder (ptx::pack16_hilo)
  ptx::pdl_trigger();
  ptx::pdl_wait();
  constexpr int V = D / 128;  // float4 per lane
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  size_t in_row = warp;
  if (skipP > 0) {
    const int b = warp / skipP;
    in_row = static_cast<size_t>(b) * (skipP + 1) + 1 + (warp - b * skipP);
  }
  const float4* xr = reinterpret_cast<const float4*>(x + in_row * D);
  float4 v[V];
#pragma unroll
  for (int j = 0; j < V; ++j) v[j] = xr[lane + 32 * j];
  if constexpr (NORMALIZE) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float a = v[j].x - mean, b2 = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      ss += (a * a + b2 * b2) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(ss) * (1.0f / D) + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 g = __ldg(g4 + lane + 32 * j), bb = __ldg(b4 + lane + 32 * j);
      v[j].x = (v[j].x - mean) * rstd * g.x + bb.x;
      v[j].y = (v[j].y - mean) * rstd * g.y + bb.y;
      v[j].z = (v[j].z - mean) * rstd * g.z + bb.z;
      v[j].w = (v[j].w - mean) * rstd * g.w + bb.w;
    }
  }
  if (out_bf16 != nullptr && out_lo != nullptr) {
    uint2* oh = reinterpret_cast<uint2*>(out_bf16 + static_cast<size_t>(warp) * D);
    uint2* ol = reinterpret_cast<uint2*>(out_lo + static_cast<size_t>(warp) * D);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      uint2 hi, lo;
      ptx::pack16_hilo<BF16>(v[j].x, v[j].y, hi.x, lo.x);
      ptx::pack16_hilo<BF16>(v[j].z, v[j].w, hi.y, lo.y);
      oh[lane + 32 * j] = hi;
      ol[lane + 32 * j] = lo;
    }
  } else if (out_bf16 != nullptr) {
    uint2* o = reinterpret_cast<uint2*>(out_bf16 + static_cast<size_t>(warp) * D);
#pragma unroll
    for (int j = 0; j < V; ++j)
      o[lane + 32 * j] = make_uint2(ptx::pack16<BF16>(v[j].x, v[j].y), ptx::pack16<BF16>(v[j].z, v[j].w));
  }
  if (out_f32 != nullptr) {
    float4* o = reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(warp) * D);
#pragma unroll
    for (int j = 0; j < V; ++j) o[lane + 32 * j] = v[j];
  }
  if (stats_out != nullptr) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
    s = warp_sum(s);
    q = warp_sum(q);
    float2* o = reinterpret_cast<float2*>(stats_out) + static_cast<size_t>(warp) * np;
    if (lane < np) o[lane] = (lane == 0) ? make_float2(s, q) : make_float2(0.f, 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Residual adapter blend (model/adapter.py:163-170 visual, :286-295 text):
//   a <- a * ||x||_2 / ||a||_2  (per token, no eps) ;  x <- w*a + (1-w)*x.   a = LeakyReLU(x W^T) from the GEMM.
// Optionally also emits bf16(x_new) (operand for the next GEMM). One warp per row.
// x16 != nullptr: also the 16-bit copy of the new x (A operand of the next GEMM); stats_out: its LayerNorm partials
// (same format as layernorm_kernel's stats_out).
template <int D, bool BF16>
__global__ void adapter_blend_kernel(float* __restrict__ x, const float* __restrict__ a, int rows, float w,
                                     float one_minus_w, uint16_t* __restrict__ x16 = nullptr,
                                     float* __restrict__ stats_out = nullptr, int np = 0) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  constexpr int V = D / 128;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float4* xr = reinterpret_cast<float4*>(x + static_cast<size_t>(warp) * D);
  const float4* ar = reinterpret_cast<const float4*>(a + static_cast<size_t>(warp) * D);
  float4 xv[V], av[V];
  float sx = 0.f, sa = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    xv[j] = xr[lane + 32 * j];
    av[j] = ar[lane + 32 * j];
    sx += (xv[j].x * xv[j].x + xv[j].y * xv[j].y) + (xv[j].z * xv[j].z + xv[j].w * xv[j].w);
    sa += (av[j].x * av[j].x + av[j].y * av[j].y) + (av[j].z * av[j].z + av[j].w * av[j].w);
  }
  const float nx = sqrtf(warp_sum(sx)), na = sqrtf(warp_sum(sa));
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    float4 o;
    o.x = w * (av[j].x * nx / na) + one_minus_w * xv[j].x;
    o.y = w * (av[j].y * nx / na) + one_minus_w * xv[j].y;
    o.z = w * (av[j].z * nx / na) + one_minus_w * xv[j].z;
    o.w = w * (av[j].w * nx / na) + one_minus_w * xv[j].w;
    xr[lane + 32 * j] = o;
    if (x16 != nullptr)
      reinterpret_cast<uint2*>(x16 + static_cast<size_t>(warp) * D)[lane + 32 * j] =
          make_uint2(ptx::pack16<BF16>(o.x, o.y), ptx::pack16<BF16>(o.z, o.w));
    s += (o.x + o.y) + (o.z + o.w);
    q += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
  }
  if (stats_out != nullptr) {
    s = warp_sum(s);
    q = warp_sum(q);
    float2* o = reinterpret_cast<float2*>(stats_out) + static_cast<size_t>(warp) * np;
    if (lane < np) o[lane] = (lane == 0) ? make_float2(s, q) : make_float2(0.f, 0.f);
  }
}

// The blend of the 16-bit modes. a arrives from the adapter GEMM's EPI_LEAKY_16_SQ epilogue in the operand dtype with
// its squared norm as `np` per-row partials (sum, sumsq) of the unrounded values (fixed order: batch-invariant); the
// residual stream is read and rewritten in its (hi, lo) form (ptx::pack16_hilo), hi being the next GEMM's operand:
// 10 bytes per element (2 + 2 + 2 read, 2 + 2 written). One scale per row: a is already rounded to 16 bits, so the
// reference's per-element order a*|x|/|a| buys nothing here (the fp32-accuracy mode keeps it).
template <int D, bool BF16>
__global__ void adapter_blend_hl_kernel(uint16_t* __restrict__ x_hi, uint16_t* __restrict__ x_lo,
                                        const uint16_t* __restrict__ a16, const float* __restrict__ a_stats, int rows,
                                        float w, float one_minus_w, float* __restrict__ stats_out, int np) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  constexpr int V = D / 128;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  uint2* hr = reinterpret_cast<uint2*>(x_hi + static_cast<size_t>(warp) * D);
  uint2* lr = reinterpret_cast<uint2*>(x_lo + static_cast<size_t>(warp) * D);
  const uint2* ar = reinterpret_cast<const uint2*>(a16 + static_cast<size_t>(warp) * D);
  float4 xv[V];
  uint2 aw[V];
  float sx = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    xv[j] = ptx::unpack_hilo4<BF16>(hr[lane + 32 * j], lr[lane + 32 * j]);
    aw[j] = ar[lane + 32 * j];
    sx += (xv[j].x * xv[j].x + xv[j].y * xv[j].y) + (xv[j].z * xv[j].z + xv[j].w * xv[j].w);
  }
  float sa = 0.f;
  {
    const float2* ap = reinterpret_cast<const float2*>(a_stats) + static_cast<size_t>(warp) * np;
    for (int i = 0; i < np; ++i) sa += __ldg(ap + i).y;   // every lane: same order, same value
  }
  const float r = w * (sqrtf(warp_sum(sx)) / sqrtf(sa));
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    float a0, a1, a2, a3;
    ptx::unpack16<BF16>(aw[j].x, a0, a1);
    ptx::unpack16<BF16>(aw[j].y, a2, a3);
    float4 o;
    o.x = fmaf(r, a0, one_minus_w * xv[j].x);
    o.y = fmaf(r, a1, one_minus_w * xv[j].y);
    o.z = fmaf(r, a2, one_minus_w * xv[j].z);
    o.w = fmaf(r, a3, one_minus_w * xv[j].w);
    uint2 hi, lo;
    ptx::pack16_hilo<BF16>(o.x, o.y, hi.x, lo.x);
    ptx::pack16_hilo<BF16>(o.z, o.w, hi.y, lo.y);
    hr[lane + 32 * j] = hi;
    lr[lane + 32 * j] = lo;
    s += (o.x + o.y) + (o.z + o.w);
    q += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
  }
  s = warp_sum(s);
  q = warp_sum(q);
  float2* o = reinterpret_cast<float2*>(stats_out) + static_cast<size_t>(warp) * np;
  if (lane < np) o[lane] = (lane == 0) ? make_float2(s, q) : make_float2(0.f, 0.f);
}

// fp32 -> (hi, lo) planes (ptx::pack16_hilo); exported as aaclip_split_hilo for the building-block tests
template <bool BF16>
__global__ void f32_to_hilo_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                   size_t n4) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 h, l;
    ptx::pack16_hilo<BF16>(v.x, v.y, h.x, l.x);
    ptx::pack16_hilo<BF16>(v.z, v.w, h.y, l.y);
    reinterpret_cast<uint2*>(hi)[i] = h;
    reinterpret_cast<uint2*>(lo)[i] = l;
  }
}

// (hi, lo) planes -> fp32 (exports of the residual stream: the CLS rows, the rows the fp32 token export normalises)
template <bool BF16>
__global__ void hilo_to_f32_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo,
                                   float* __restrict__ out, size_t n4) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    reinterpret_cast<float4*>(out)[i] =
        ptx::unpack_hilo4<BF16>(reinterpret_cast<const uint2*>(hi)[i], reinterpret_cast<const uint2*>(lo)[i]);
}

// Validation pass for the 16-bit operand mode (aaclip_saturation_check): counts the elements of a 16-bit activation
// buffer that hit the format's ceiling -- fp16: +-65504 (what cvt.rn.satfinite stores for anything larger) or a
// non-finite pattern; bf16: inf / nan (bf16 has fp32's range, so only non-finite inputs get there). n is a multiple of 8
// (every activation row is), p is 16-byte aligned. Not launched unless the check is enabled.
template <bool BF16>
__global__ void count_saturated16_kernel(const uint16_t* __restrict__ p, size_t n8, unsigned long long* counter) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const uint4* v = reinterpret_cast<const uint4*>(p);
  unsigned int local = 0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 q = v[i];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t lo = w[k] & 0x7FFFu, hi = (w[k] >> 16) & 0x7FFFu;
      constexpr uint32_t kCeil = BF16 ? 0x7F80u : 0x7BFFu;
      local += (lo >= kCeil) + (hi >= kCeil);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local != 0) atomicAdd(counter, static_cast<unsigned long long>(local));
}

// LayerNorm folding at weight-load time (model/transformer.py:245-254: LN followed by a Linear):
//   W'[n,k] = W[n,k] * gamma[k] (rounded to the 16-bit operand type),  colsum[n] = sum_k W'_16[n,k],
//   bias'[n] = b[n] + sum_k W[n,k] * beta[k].
// The GEMM then consumes the RAW residual stream: LN(x) W^T + b = rstd * (x W'^T - mean * colsum) + bias'.
// colsum is taken over the ROUNDED W' so that the mean term cancels exactly what the tensor core accumulates.
// One warp per output feature n; fixed summation order.
template <bool BF16>
__global__ void ln_fold_kernel(const float* __restrict__ W, const float* __restrict__ b,
                               const float* __restrict__ gamma, const float* __restrict__ beta,
                               uint16_t* __restrict__ W16, float* __restrict__ colsum, float* __restrict__ bias_f,
                               int N, int K) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* wr = W + static_cast<size_t>(n) * K;
  uint16_t* o = W16 + static_cast<size_t>(n) * K;
  float cs = 0.f, bs = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = wr[k];
    const uint16_t h = ptx::cvt16<BF16>(w * gamma[k]);
    o[k] = h;
    float hv;
    if constexpr (BF16) hv = __uint_as_float(static_cast<uint32_t>(h) << 16);
    else hv = __half2float(__ushort_as_half(h));
    cs += hv;
    bs = fmaf(w, beta[k], bs);
  }
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  if (lane == 0) {
    colsum[n] = cs;
    bias_f[n] = b[n] + bs;
  }
}

// F.normalize(t, dim=-1) in place: t / max(||t||_2, 1e-12)   (model/adapter.py:182-184). One warp per row.
template <int D>
__global__ void l2norm_rows_kernel(float* __restrict__ t, int rows) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  constexpr int V = D / 128;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float4* r = reinterpret_cast<float4*>(t + static_cast<size_t>(warp) * D);
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    v[j] = r[lane + 32 * j];
    s += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(s)), 1e-12f);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    v[j].x *= inv; v[j].y *= inv; v[j].z *= inv; v[j].w *= inv;
    r[lane + 32 * j] = v[j];
  }
}

// det_token = normalized_rows.mean(dim=1): [B,P,E] -> [B,E]   (model/adapter.py:184).
// block = (MEAN_COLS columns, MEAN_SEGS patch segments): every thread sums one contiguous segment of patches with eight
// independent accumulators, the segments are combined through shared memory in a fixed order. The summation order
// depends on (P, MEAN_SEGS) alone: bit-reproducible and independent of the batch.
constexpr int MEAN_SEGS = 16, MEAN_COLS = 64;
__global__ void __launch_bounds__(MEAN_COLS * MEAN_SEGS)
mean_over_patches_kernel(const float* __restrict__ t, float* __restrict__ out, int P, int E) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  __shared__ float part[MEAN_SEGS][MEAN_COLS];
  const int b = blockIdx.y;
  const int e = blockIdx.x * MEAN_COLS + threadIdx.x;
  const int seg = threadIdx.y;
  const int per = (P + MEAN_SEGS - 1) / MEAN_SEGS;
  const int p0 = seg * per, p1 = min(P, p0 + per);
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (e < E) {
    const float* src = t + static_cast<size_t>(b) * P * E + e;
    int p = p0;
    for (; p + 8 <= p1; p += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += src[static_cast<size_t>(p + u) * E];
    }
    for (; p < p1; ++p) a[0] += src[static_cast<size_t>(p) * E];
  }
  part[seg][threadIdx.x] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  if (seg == 0 && e < E) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < MEAN_SEGS; ++k) s += part[k][threadIdx.x];
    out[static_cast<size_t>(b) * E + e] = s / static_cast<float>(P);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Patch-text similarity, forward_utils.py:199: s[b,l,c,p] = 100 * <f_l[b,p,:], text[(b),:,c]>, c in {0,1}.
// One warp per patch row, both anchors in one pass over the row. text_bstride = 0 for a shared [E,2] anchor,
// E*2 for per-image anchors [B,E,2].
struct SegPtrs { const float* p[4]; };

// NCH 16-byte feature loads per lane (issued before the first FMA) against the anchor pair [E,2], read as two float4
// per feature float4 (6 KB, L1-resident)
// (q accumulates the squared norm of the row for the RAW form of the kernel; dead code otherwise)
template <int NCH>
__device__ __forceinline__ void similarity_row(const float* f, const float* t, int lane, int i0, float& s0, float& s1,
                                               float& q) {
  const float4* f4 = reinterpret_cast<const float4*>(f);
  const float4* t4 = reinterpret_cast<const float4*>(t);
  float4 fv[NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) fv[j] = f4[i0 + lane + 32 * j];
  float u0 = 0.f, u1 = 0.f;
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int i = i0 + lane + 32 * j;
    const float4 ta = __ldg(t4 + 2 * i), tb = __ldg(t4 + 2 * i + 1);   // (x0 y0 x1 y1) (x2 y2 x3 y3)
    s0 = fmaf(fv[j].x, ta.x, s0); s1 = fmaf(fv[j].x, ta.y, s1);
    u0 = fmaf(fv[j].y, ta.z, u0); u1 = fmaf(fv[j].y, ta.w, u1);
    s0 = fmaf(fv[j].z, tb.x, s0); s1 = fmaf(fv[j].z, tb.y, s1);
    u0 = fmaf(fv[j].w, tb.z, u0); u1 = fmaf(fv[j].w, tb.w, u1);
    q += (fv[j].x * fv[j].x + fv[j].y * fv[j].y) + (fv[j].z * fv[j].z + fv[j].w * fv[j].w);
  }
  s0 += u0;
  s1 += u1;
}

// RAW = true: the feature rows are the un-normalised projections (aaclip_seg_normalize(h, 0)); the kernel applies
// F.normalize(dim=-1, eps=1e-12) (model/adapter.py:182) itself -- <f/max(|f|,eps), t> -- instead of a separate pass
// that reads and re-writes all four feature tensors.
template <bool RAW>
__global__ void similarity_kernel(SegPtrs seg, const float* __restrict__ text, long long text_bstride,
                                  float* __restrict__ s, int B, int L, int P, int E) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(B) * L * P;
  if (warp >= total) return;
  const int p = static_cast<int>(warp % P);
  const int l = static_cast<int>((warp / P) % L);
  const int b = static_cast<int>(warp / (static_cast<long long>(P) * L));
  const float* f = seg.p[l] + (static_cast<size_t>(b) * P + p) * E;
  const float* t = text + static_cast<size_t>(b) * text_bstride;
  float s0 = 0.f, s1 = 0.f, q = 0.f;
  if (E == 768) {
    similarity_row<6>(f, t, lane, 0, s0, s1, q);   // all six 16-byte loads of the lane in flight at once
  } else if ((E & 127) == 0) {
    for (int i0 = 0; i0 < E / 4; i0 += 32) similarity_row<1>(f, t, lane, i0, s0, s1, q);
  } else {
    for (int i = lane; i < E; i += 32) {
      const float fv = f[i];
      const float2 tv = __ldg(reinterpret_cast<const float2*>(t) + i);
      s0 = fmaf(fv, tv.x, s0);
      s1 = fmaf(fv, tv.y, s1);
      q = fmaf(fv, fv, q);
    }
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  float scale = 100.0f;
  if constexpr (RAW) scale = 100.0f / fmaxf(sqrtf(warp_sum(q)), 1e-12f);
  if (lane == 0) {
    float* o = s + ((static_cast<size_t>(b) * L + l) * 2) * P + p;
    o[0] = scale * s0;
    o[P] = scale * s1;
  }
}

// image score, test_last.py:90-92 at the well-posed level: ((det[b,:] @ text[(b),:,1]) + 1) / 2
__global__ void image_score_kernel(const float* __restrict__ det, const float* __restrict__ text,
                                   long long text_bstride, float* __restrict__ score, int B, int E) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* d = det + static_cast<size_t>(warp) * E;
  const float* t = text + static_cast<size_t>(warp) * text_bstride;
  float s = 0.f;
  for (int i = lane; i < E; i += 32) s = fmaf(d[i], __ldg(t + 2 * i + 1), s);
  s = warp_sum(s);
  if (lane == 0) score[warp] = (s + 1.0f) * 0.5f;
}

// ---------------------------------------------------------------------------------------------------------
// Anomaly-map tail, forward_utils.py:200-215 (+ the level sum of test_last.py:143,149).
//  test mode : g_l = (s1 + 1 - s0)/2 on the GxG grid; Gaussian blur ksize x ksize (reflect border, kornia
//              gaussian_blur2d); bilinear upsample align_corners=True to SxS; summed over the L levels given.
//              out [B, S, S]  (L = 1 reproduces a single calculate_similarity_map(test=True) call).
//  train mode: both channels upsampled, then softmax over the channel pair. out [B, 2, S, S], L must be 1.
// grid = (ceil(S / ROWS), B); every CTA rebuilds the (tiny) blurred grid in smem and writes ROWS output rows
// with fully coalesced stores.
constexpr int MAP_ROWS = 14;
constexpr int MAP_MAXK = 15;

struct MapTaps { float w[MAP_MAXK]; };

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__global__ void anomaly_map_kernel(const float* __restrict__ s, float* __restrict__ out, int L, int G, int S,
                                   int test_mode, int ksize, MapTaps taps) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  extern __shared__ float sm[];
  const int P = G * G;
  float* g0 = sm;            // test: summed blurred grid ; train: channel 0
  float* g1 = sm + P;        // scratch / channel 1
  float* g2 = sm + 2 * P;    // scratch
  const int b = blockIdx.y;
  const float* sb = s + static_cast<size_t>(b) * L * 2 * P;

  // bilinear, align_corners=True (ATen upsample_bilinear2d index/lambda arithmetic)
  const float scale = (S > 1) ? static_cast<float>(G - 1) / static_cast<float>(S - 1) : 0.f;
  const int y_begin = blockIdx.x * MAP_ROWS;
  const int y_end = min(S, y_begin + MAP_ROWS);

  if (test_mode) {
    // Only the grid rows this CTA's MAP_ROWS output rows interpolate between are blurred (3 of 37 at 518 px), from
    // the horizontally blurred rows within the kernel radius of them: every CTA used to rebuild the whole grid.
    // Same operations in the same order per element: bit-identical maps.
    const int r = ksize / 2;
    const int gy_lo = static_cast<int>(scale * y_begin);
    const int gy_hi = min(G - 1, static_cast<int>(scale * (y_end - 1)) + 1);
    const int h_lo = max(0, gy_lo - r), h_hi = min(G - 1, gy_hi + r);   // reflected indices stay inside [h_lo, h_hi]
    const int nh = (h_hi - h_lo + 1) * G, nv = (gy_hi - gy_lo + 1) * G;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) g0[gy_lo * G + i] = 0.f;
    __syncthreads();
    for (int l = 0; l < L; ++l) {
      const float* s0 = sb + static_cast<size_t>(l) * 2 * P;
      const float* s1 = s0 + P;
      for (int i = h_lo * G + threadIdx.x; i < h_lo * G + nh; i += blockDim.x) g1[i] = (s1[i] + 1.0f - s0[i]) / 2.0f;
      __syncthreads();
      // horizontal pass
      for (int i = h_lo * G + threadIdx.x; i < h_lo * G + nh; i += blockDim.x) {
        const int y = i / G, x = i - y * G;
        float acc = 0.f;
        for (int k = 0; k < ksize; ++k) acc = fmaf(taps.w[k], g1[y * G + reflect_idx(x + k - r, G)], acc);
        g2[i] = acc;
      }
      __syncthreads();
      // vertical pass, accumulated over levels
      for (int i = gy_lo * G + threadIdx.x; i < gy_lo * G + nv; i += blockDim.x) {
        const int y = i / G, x = i - y * G;
        float acc = 0.f;
        for (int k = 0; k < ksize; ++k) acc = fmaf(taps.w[k], g2[reflect_idx(y + k - r, G) * G + x], acc);
        g0[i] += acc;
      }
      __syncthreads();
    }
  } else {
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
      g0[i] = sb[i];
      g1[i] = sb[P + i];
    }
    __syncthreads();
  }

  for (int y = y_begin; y < y_end; ++y) {
    const float fy = scale * y;
    const int y0 = static_cast<int>(fy);
    const int yp = (y0 < G - 1) ? 1 : 0;
    const float ly1 = fy - y0, ly0 = 1.0f - ly1;
    for (int x = threadIdx.x; x < S; x += blockDim.x) {
      const float fx = scale * x;
      const int x0 = static_cast<int>(fx);
      const int xp = (x0 < G - 1) ? 1 : 0;
      const float lx1 = fx - x0, lx0 = 1.0f - lx1;
      const int i00 = y0 * G + x0, i01 = i00 + xp, i10 = i00 + yp * G, i11 = i10 + xp;
      const float v0 = ly0 * (lx0 * g0[i00] + lx1 * g0[i01]) + ly1 * (lx0 * g0[i10] + lx1 * g0[i11]);
      if (test_mode) {
        out[(static_cast<size_t>(b) * S + y) * S + x] = v0;
      } else {
        const float v1 = ly0 * (lx0 * g1[i00] + lx1 * g1[i01]) + ly1 * (lx0 * g1[i10] + lx1 * g1[i11]);
        const float m = fmaxf(v0, v1);
        const float e0 = expf(v0 - m), e1 = expf(v1 - m);
        const float inv = 1.0f / (e0 + e1);
        out[((static_cast<size_t>(b) * 2 + 0) * S + y) * S + x] = e0 * inv;
        out[((static_cast<size_t>(b) * 2 + 1) * S + y) * S + x] = e1 * inv;
      }
    }
  }
}

}  // namespace aaclip

// =========================================================================================================
// Text-anchor path (AdaptedCLIP.encode_text, model/adapter.py:273-304; forward_utils.py:138-162)
namespace aaclip {

// x[s*ctx + t, :] = token_embedding[ids[s,t], :] + positional_embedding[t, :]   (model/adapter.py:277-281)
// grid = n_sent * ctx rows, block = D/4 threads (float4 each). Out-of-vocabulary ids are clamped.
__global__ void text_embed_kernel(const int32_t* __restrict__ ids, const float* __restrict__ tok_emb,
                                  const float* __restrict__ pos, float* __restrict__ x, int ctx, int D, int vocab) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int row = blockIdx.x;
  const int t = row % ctx;
  const int id = min(max(ids[row], 0), vocab - 1);
  const float4* e = reinterpret_cast<const float4*>(tok_emb + static_cast<size_t>(id) * D);
  const float4* p = reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * D);
  float4* o = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * D);
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    const float4 a = __ldg(e + i), b = __ldg(p + i);
    o[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// out[s, :] = x[s*ctx + argmax_t ids[s,t], :]  -- the EOT row (highest token id; torch.argmax returns the FIRST
// maximum), model/adapter.py:298. ln_final is row-wise, so it is applied after this gather on n_sent rows only.
__global__ void text_gather_eot_kernel(const int32_t* __restrict__ ids, const float* __restrict__ x,
                                       float* __restrict__ out, int ctx, int D) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int s = blockIdx.x;
  __shared__ int s_best;
  if (threadIdx.x < 32) {
    int best_v = INT_MIN, best_i = 0;
    for (int t = threadIdx.x; t < ctx; t += 32) {
      const int v = ids[s * ctx + t];
      if (v > best_v) { best_v = v; best_i = t; }     // strided scan keeps the first maximum per lane
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int ov = __shfl_xor_sync(0xffffffffu, best_v, o), oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
    }
    if (threadIdx.x == 0) s_best = best_i;
  }
  __syncthreads();
  const float4* src = reinterpret_cast<const float4*>(x + (static_cast<size_t>(s) * ctx + s_best) * D);
  float4* dst = reinterpret_cast<float4*>(out + static_cast<size_t>(s) * D);
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) dst[i] = src[i];
}

// the same gather from the (hi, lo) form of the residual stream (16-bit modes)
template <bool BF16>
__global__ void text_gather_eot_hl_kernel(const int32_t* __restrict__ ids, const uint16_t* __restrict__ x_hi,
                                          const uint16_t* __restrict__ x_lo, float* __restrict__ out, int ctx, int D) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int s = blockIdx.x;
  __shared__ int s_best;
  if (threadIdx.x < 32) {
    int best_v = INT_MIN, best_i = 0;
    for (int t = threadIdx.x; t < ctx; t += 32) {
      const int v = ids[s * ctx + t];
      if (v > best_v) { best_v = v; best_i = t; }     // strided scan keeps the first maximum per lane
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int ov = __shfl_xor_sync(0xffffffffu, best_v, o), oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
    }
    if (threadIdx.x == 0) s_best = best_i;
  }
  __syncthreads();
  const size_t row = (static_cast<size_t>(s) * ctx + s_best) * D;
  const uint2* h = reinterpret_cast<const uint2*>(x_hi + row);
  const uint2* l = reinterpret_cast<const uint2*>(x_lo + row);
  float4* dst = reinterpret_cast<float4*>(out + static_cast<size_t>(s) * D);
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) dst[i] = ptx::unpack_hilo4<BF16>(h[i], l[i]);
}

// Prompt-ensemble anchors, forward_utils.py:155-161: for group g (rows offsets[g] .. offsets[g+1]-1 of emb):
//   e_r = emb_r / ||emb_r|| ; m = mean_r e_r ; a = m / ||m|| ; anchors[g / 2, :, g % 2] = a
// (two consecutive groups = normal / abnormal state of one class -> the [E,2] layout of torch.stack(dim=1)).
// One CTA per group, fixed summation order.
__global__ void text_ensemble_kernel(const float* __restrict__ emb, const int32_t* __restrict__ offsets,
                                     float* __restrict__ anchors, int E) {
  extern __shared__ float sm[];      // [E] accumulator + [32] reduction scratch
  float* acc = sm;
  float* red = sm + E;
  const int g = blockIdx.x;
  const int r0 = offsets[g], r1 = offsets[g + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  auto block_sum = [&](float v) -> float {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += red[w];
    return t;
  };
  for (int i = threadIdx.x; i < E; i += blockDim.x) acc[i] = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float* row = emb + static_cast<size_t>(r) * E;
    float ss = 0.f;
    for (int i = threadIdx.x; i < E; i += blockDim.x) ss += row[i] * row[i];
    const float inv = 1.0f / sqrtf(block_sum(ss));
    for (int i = threadIdx.x; i < E; i += blockDim.x) acc[i] += row[i] * inv;
  }
  const float cnt = static_cast<float>(r1 - r0);
  float ss = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    acc[i] /= cnt;
    ss += acc[i] * acc[i];
  }
  const float inv = 1.0f / sqrtf(block_sum(ss));
  float* out = anchors + static_cast<size_t>(g >> 1) * E * 2 + (g & 1);
  for (int i = threadIdx.x; i < E; i += blockDim.x) out[2 * i] = acc[i] * inv;
}

}  // namespace aaclip
