// Image pre-processing on the device: the reference's per-image transform
//   transforms.Resize((S, S), Image.BICUBIC) -> ToTensor() -> Normalize(CLIP mean, std)      dataset/__init__.py:150-159
// for a batch of decoded RGB images, bit-exact with PIL + torchvision:
//   * Pillow's ImagingResample (src/libImaging/Resample.c) in its 8-bit fixed-point form -- antialiased bicubic
//     (a = -0.5) taps normalised per output pixel and rounded to 22 fractional bits, accumulator seeded with 1 << 21,
//     clip8 = clamp(acc >> 22, 0, 255) -- horizontal pass first, its uint8 result feeds the vertical pass;
//   * ToTensor: float(u8) / 255 (IEEE fp32 division); Normalize: (x - mean) / std (IEEE fp32).
// The tap tables (first tap, tap count, fixed-point weights per output coordinate) are computed once per (in, out) size
// on the HOST in double precision exactly as precompute_coeffs / normalize_coeffs_8bpc do (aaclip_resample_coeffs_host).
// Byte work, HBM bound: 3 B/pixel in, 12 B/pixel out; the uint8 intermediate [B, H, S, 3] stays in L2 for typical sizes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace aaclip {

constexpr int RESAMPLE_PRECISION_BITS = 32 - 8 - 2;

// ---- host: Resample.c precompute_coeffs + normalize_coeffs_8bpc for the box (0, in_size), bicubic filter ----------
inline double bicubic_filter_host(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
inline int resample_ksize_host(int in_size, int out_size) {
  double filterscale = static_cast<double>(in_size) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  return static_cast<int>(ceil(2.0 * filterscale)) * 2 + 1;
}
// bounds [out_size][2] = (first tap, tap count), kk [out_size][ksize] (unused tail zero)
inline void resample_coeffs_host(int in_size, int out_size, int ksize, int32_t* bounds, int32_t* kk) {
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const double ss = 1.0 / filterscale;
  double* k = new double[ksize];
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter_host((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = 0; x < ksize; ++x) {
      int32_t v = 0;
      if (x < xmax) {
        const double f = k[x] * (1 << RESAMPLE_PRECISION_BITS);
        v = k[x] < 0 ? static_cast<int32_t>(-0.5 + f) : static_cast<int32_t>(0.5 + f);
      }
      kk[static_cast<size_t>(xx) * ksize + x] = v;
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  delete[] k;
}

__device__ __forceinline__ uint8_t clip8_fixed(int32_t acc) {
  const int32_t v = acc >> RESAMPLE_PRECISION_BITS;   // arithmetic shift, as in Resample.c's clip8 lookup
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// Horizontal pass: src [rows, W, 3] uint8 -> dst [rows, S, 3] uint8. One CTA per HROWS consecutive image rows: their
// bytes (one contiguous span) are staged in shared memory with 16-byte loads (rows start at arbitrary byte offsets:
// W * 3 is rarely a multiple of 16), then the CTA's threads share the HROWS * S output pixels (all three channels each).
// Horizontal pass, unrolled form (KS >= ksize taps, KS in {5, 9, 13}): a thread owns up to two output columns for all of
// the CTA's HROWS_X rows -- their first tap and fixed-point weights live in registers (no table loads, no integer
// division in the row loop) -- and reads a pixel window as aligned 32-bit shared-memory words realigned with funnel
// shifts (8 word loads instead of 27 byte loads per pixel; ncu showed the byte version at 94 % of the L1/LSU rate).
// Taps beyond a column's count carry weight 0, so the bytes they cover (the next pixels / the zeroed slack) are inert.
constexpr int HROWS = 4;       // generic kernel
constexpr int HROWS_X = 8;     // unrolled kernel
template <int KS>
__global__ void __launch_bounds__(1024)
resample_horizontal_x_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int W, int S,
                             const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize,
                             long long n_rows, long long total_bytes) {
  constexpr int NW = (3 * KS + 3 + 3) / 4;           // words covering 3*KS bytes at any byte offset
  extern __shared__ __align__(16) uint8_t row_smem[];
  const long long row0 = static_cast<long long>(blockIdx.x) * HROWS_X;
  const int rows = static_cast<int>(n_rows - row0 < HROWS_X ? n_rows - row0 : HROWS_X);
  const long long r0 = row0 * W * 3;                 // first byte of this CTA's span
  const long long r1 = r0 + static_cast<long long>(rows) * W * 3;
  const long long a0 = r0 & ~15ll;
  const int skew = static_cast<int>(r0 - a0);
  const int n_chunks = static_cast<int>((r1 - a0 + 15) >> 4) + 3;   // + 48 bytes of zeroed slack behind the last row
  for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) {
    const long long g = a0 + (static_cast<long long>(c) << 4);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (g + 16 <= total_bytes) {
      v = *reinterpret_cast<const uint4*>(src + g);
    } else if (g < total_bytes) {                    // last chunk of the buffer: never read past its end
      uint8_t tail[16];
      for (int b = 0; b < 16; ++b) tail[b] = (g + b < total_bytes) ? src[g + b] : 0;
      v = *reinterpret_cast<const uint4*>(tail);
    }
    *reinterpret_cast<uint4*>(row_smem + (c << 4)) = v;
  }
  const int half = (S + 1) >> 1;                     // thread t owns columns t and t + half
  int xmin[2] = {0, 0};
  int32_t w[2][KS];
  bool live[2];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int xo = threadIdx.x + it * half;
    live[it] = threadIdx.x < half && xo < S;
#pragma unroll
    for (int x = 0; x < KS; ++x) w[it][x] = 0;
    if (live[it]) {
      xmin[it] = bounds[2 * xo];
      const int n = bounds[2 * xo + 1];
      const int32_t* k = kk + static_cast<size_t>(xo) * ksize;
#pragma unroll
      for (int x = 0; x < KS; ++x)
        if (x < n) w[it][x] = k[x];
    }
  }
  __syncthreads();
  const uint32_t* words = reinterpret_cast<const uint32_t*>(row_smem);
  uint8_t* out = dst + row0 * S * 3;
  for (int r = 0; r < rows; ++r) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      if (!live[it]) continue;
      const int o = skew + (r * W + xmin[it]) * 3;   // first byte of the window
      const uint32_t* wp = words + (o >> 2);
      const uint32_t sh = (o & 3) * 8;
      uint32_t raw[NW];
#pragma unroll
      for (int i = 0; i < NW; ++i) raw[i] = wp[i];
      int32_t s0 = 1 << (RESAMPLE_PRECISION_BITS - 1), s1 = s0, s2 = s0;
#pragma unroll
      for (int x = 0; x < KS; ++x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int k = 3 * x + c;                   // byte k of the realigned window (compile-time position)
          const uint32_t al = __funnelshift_r(raw[k >> 2], raw[(k >> 2) + 1 < NW ? (k >> 2) + 1 : NW - 1], sh);
          const int32_t px = static_cast<int32_t>(__byte_perm(al, 0u, 0x4440u | (k & 3)));   // one PRMT per byte
          if (c == 0) s0 += px * w[it][x];
          if (c == 1) s1 += px * w[it][x];
          if (c == 2) s2 += px * w[it][x];
        }
      }
      const int xo = threadIdx.x + it * half;
      uint8_t* q = out + (static_cast<size_t>(r) * S + xo) * 3;
      q[0] = clip8_fixed(s0);
      q[1] = clip8_fixed(s1);
      q[2] = clip8_fixed(s2);
    }
  }
}

// Horizontal pass, generic form (any ksize): src [rows, W, 3] uint8 -> dst [rows, S, 3] uint8. One CTA per HROWS image
// rows, staged in shared memory with 16-byte loads (rows start at arbitrary byte offsets: W * 3 is rarely a multiple of
// 16); the CTA's threads share the HROWS * S output pixels and loop over the taps.
__global__ void __launch_bounds__(256)
resample_horizontal_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int W, int S,
                           const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize,
                           long long n_rows, long long total_bytes) {
  extern __shared__ __align__(16) uint8_t row_smem[];
  const long long row0 = static_cast<long long>(blockIdx.x) * HROWS;
  const int rows = static_cast<int>(n_rows - row0 < HROWS ? n_rows - row0 : HROWS);
  const long long r0 = row0 * W * 3;                 // first byte of this CTA's span
  const long long r1 = r0 + static_cast<long long>(rows) * W * 3;
  const long long a0 = r0 & ~15ll;
  const int skew = static_cast<int>(r0 - a0);
  const int n_chunks = static_cast<int>((r1 - a0 + 15) >> 4);
  for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) {
    const long long g = a0 + (static_cast<long long>(c) << 4);
    if (g + 16 <= total_bytes) {
      *reinterpret_cast<uint4*>(row_smem + (c << 4)) = *reinterpret_cast<const uint4*>(src + g);
    } else {                                         // last chunk of the buffer: never read past its end
      for (int b = 0; b < 16; ++b) row_smem[(c << 4) + b] = (g + b < total_bytes) ? src[g + b] : 0;
    }
  }
  __syncthreads();
  uint8_t* out = dst + row0 * S * 3;
  for (int o = threadIdx.x; o < rows * S; o += blockDim.x) {
    const int r = o / S, xo = o - r * S;
    const int xmin = bounds[2 * xo], n = bounds[2 * xo + 1];
    const int32_t* k = kk + static_cast<size_t>(xo) * ksize;
    int32_t s0 = 1 << (RESAMPLE_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    const uint8_t* p = row_smem + skew + (static_cast<size_t>(r) * W + xmin) * 3;
    for (int x = 0; x < n; ++x) {
      const int32_t w = k[x];
      s0 += p[3 * x + 0] * w;
      s1 += p[3 * x + 1] * w;
      s2 += p[3 * x + 2] * w;
    }
    out[3 * o + 0] = clip8_fixed(s0);
    out[3 * o + 1] = clip8_fixed(s1);
    out[3 * o + 2] = clip8_fixed(s2);
  }
}

// ToTensor + Normalize of every possible byte: lut[c][v] = ((float(v) / 255) - mean_c) / std_c with IEEE fp32 division
// and subtraction (bit parity with torch; never compiled with -use_fast_math). 768 entries, rebuilt per call (~2 us):
// the vertical pass then needs no division at all.
__global__ void resample_lut_kernel(float* __restrict__ lut, float m0, float m1, float m2, float d0, float d1, float d2) {
  const int v = threadIdx.x, c = blockIdx.x;
  const float m = c == 0 ? m0 : (c == 1 ? m1 : m2), d = c == 0 ? d0 : (c == 1 ? d1 : d2);
  lut[c * 256 + v] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), m), d);
}

// Vertical pass fused with ToTensor + Normalize: tmp [B, H, S, 3] uint8 -> out [B, 3, S, S] fp32.
// A warp reads 96 contiguous bytes per tap row and writes three 128-byte runs (one per channel plane).
template <int KS>
__global__ void __launch_bounds__(128)
resample_vertical_normalize_kernel(const uint8_t* __restrict__ tmp, float* __restrict__ out, int H, int S,
                                   const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize,
                                   const float* __restrict__ lut) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (yo, xo) flattened: no idle lanes at the row ends
  const int b = blockIdx.y;
  if (idx >= S * S) return;
  const int yo = idx / S, xo = idx - yo * S;
  const int ymin = bounds[2 * yo], n = bounds[2 * yo + 1];
  const int32_t* k = kk + static_cast<size_t>(yo) * ksize;
  const uint8_t* p = tmp + ((static_cast<size_t>(b) * H + ymin) * S + xo) * 3;
  const size_t stride = static_cast<size_t>(S) * 3;
  int32_t s0 = 1 << (RESAMPLE_PRECISION_BITS - 1), s1 = s0, s2 = s0;
  if constexpr (KS > 0) {
    int32_t w[KS];
#pragma unroll
    for (int y = 0; y < KS; ++y) w[y] = (y < n) ? k[y] : 0;
#pragma unroll
    for (int y = 0; y < KS; ++y) {
      const uint8_t* py = p + static_cast<size_t>(min(y, n - 1)) * stride;   // unused taps re-read the last row (x 0)
      s0 += py[0] * w[y];
      s1 += py[1] * w[y];
      s2 += py[2] * w[y];
    }
  } else {
    for (int y = 0; y < n; ++y) {
      const int32_t w = k[y];
      s0 += p[0] * w;
      s1 += p[1] * w;
      s2 += p[2] * w;
      p += stride;
    }
  }
  // ToTensor + Normalize through the 3 x 256 table of resample_lut_kernel (L1-resident)
  const size_t plane = static_cast<size_t>(S) * S;
  float* o = out + static_cast<size_t>(b) * 3 * plane + static_cast<size_t>(yo) * S + xo;
  o[0] = __ldg(lut + clip8_fixed(s0));
  o[plane] = __ldg(lut + 256 + clip8_fixed(s1));
  o[2 * plane] = __ldg(lut + 512 + clip8_fixed(s2));
}

}  // namespace aaclip
