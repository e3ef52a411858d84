// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   A : fp16|bf16 row-major [M,K] (activations, token-major)         -- TMA, SWIZZLE_128B, box 64(K) x 128(M)
//   W : fp16|bf16 row-major [N,K] (nn.Linear weight layout, K-major) -- TMA, SWIZZLE_128B, box 64(K) x 128(N half)
//   D : fp32 accumulator in TMEM, 128 lanes x 256 columns per CTA, double-buffered (512 columns)
//
// One cluster of 2 CTAs (tcgen05 cta_group::2) computes a 256 x 256 output tile; see gemm_pair_kernel below.
// Roles per CTA (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only) + TMEM owner,
// warps 2..9 = epilogue (TMEM -> registers -> swizzled smem transpose -> fused math -> row-coalesced global
// accesses). The epilogue of tile i overlaps the main loop of tile i+1 through the second TMEM accumulator stage.
//
// Replaces, on the reference's path: F.linear / addmm calls inside nn.MultiheadAttention
// (model/transformer.py:200,237), the MLP (model/transformer.py:211-219), SimpleAdapter / SimpleProj
// (model/adapter_modules.py:6-26), the patch-embedding conv (model/adapter.py:139) and -- folded into the epilogues
// of the GEMMs that consume / produce the residual stream -- the LayerNorms ln_1 / ln_2 (model/transformer.py:245,252).
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
  // LayerNorm folded into the consuming GEMM: A is the RAW 16-bit copy of the residual stream, W already carries
  // gamma (W' = W diag(gamma)), and  LN(x) W^T + b  =  rstd * (x W'^T - mean * colsum(W')) + (b + W beta):
  EPI_LN_BIAS_BF16 = 8,       // out_16 = rstd*(acc - mean*colsum) + bias'
  EPI_LN_BIAS_GELU_BF16 = 9,  // out_16 = gelu_erf(...)
  EPI_LN_BIAS_QGELU_BF16 = 10,
  // residual adapter (model/adapter_modules.py:6-13): out_16 = LeakyReLU(acc) in the operand dtype, plus per-row partial
  // (sum, sum of squares) of the UNROUNDED values over each 128-column slice (stats_out) -- the blend kernel needs
  // ||a||_2 of the whole row and then reads a as 2 bytes per element instead of 4
  EPI_LEAKY_16_SQ = 11,
  // token taps (model/adapter.py:171-184): ln_post folded into the seg / det / query projections. A = the raw 16-bit
  // residual stream INCLUDING the CLS rows (row m = b*(P+1) + t); rows with t = 0 are dropped and row m lands at
  // out_f32[b*P + t - 1] -- x[1:] of every image, without a LayerNorm pass or a token copy
  EPI_LN_F32_TOK = 12,        // out_f32 = rstd*(acc - mean*colsum) + bias'
  EPI_LN_LEAKY_F32_TOK = 13,  // ... followed by LeakyReLU(0.01) (--relu)
  // residual update on the (hi, lo) form of the stream: x = hi + lo (two 16-bit planes, ptx::pack16_hilo);
  // x += acc + bias; hi (== the 16-bit operand copy the next GEMM reads) and lo are rewritten in place, plus the
  // LayerNorm partials of the new rows. 8 bytes per element per update instead of the 10 of EPI_BIAS_RESID_F32.
  EPI_BIAS_RESID_HL = 14,
};

struct GemmEpi {
  const float* bias = nullptr;        // [N]
  float* out_f32 = nullptr;           // [*, ldo]
  uint16_t* out_16 = nullptr;         // [*, ldo] fp16 or bf16 (the operand dtype); EPI_BIAS_RESID_HL: the hi plane (in / out)
  uint16_t* out_lo = nullptr;         // EPI_BIAS_RESID_HL: the lo plane (in / out)
  const float* pos = nullptr;         // EPI_PATCH_F32: positional embedding [(P+1), N]
  int ldo = 0;                        // leading dimension of the outputs, in elements
  int patches = 0;                    // EPI_PATCH_F32: P
  // LayerNorm statistics of the residual stream, as per-row partial sums [rows][n_partials][2] = (sum, sum of squares)
  // over disjoint column ranges; a row's mean / rstd are finished by whoever consumes it (fixed order: batch-invariant)
  const float* stats = nullptr;       // EPI_LN_*: statistics of the A rows
  const float* colsum = nullptr;      // EPI_LN_*: [N] column sums of the 16-bit W' actually multiplied
  float* stats_out = nullptr;         // EPI_BIAS_RESID_F32 (optional): statistics of the new residual rows, N/128 partials
  int n_partials = 0;                 // EPI_LN_*: partials per row in `stats`
  float inv_width = 0.f;              // EPI_LN_*: 1 / (LayerNorm width)
  float eps = 1e-5f;
  // Row-block order of the persistent schedule: 0 = ascending, 1 = descending. The activation tensors of the step
  // (90-360 MB) are larger than the 126 MB L2, so a consumer that walks the rows in the SAME direction as its producer
  // finds nothing of it in L2; walking them the other way round starts with the rows the producer wrote last. The
  // launch sequence alternates (aaclip_api.cu: BlockRun); the result is bit-identical either way.
  int descending = 0;
};


// erf-GELU, x * Phi(x) = max(x, 0) - |x| Phi(-|x|)  (x >= 0: x - x q;  x < 0: x q -- no compare / select, and no
// cancellation on the negative side). This epilogue runs 128 times per thread per tile and bounds the c_fc GEMM; its
// MUFUs share the MIO queue with the epilogue's TMEM loads and shared-memory transpose. Measured on B200 (c_fc in the
// step, tools/variant_check.sh; map error against the reference's golden vectors unchanged unless noted):
//   form 0  Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7): rcp + ex2, 11 FMA-pipe + 2 MUFU            7.9 ms
//   form 5  Phi(-a) = 2^P(a), P = degree-5 fit of log2 Phi(-a) weighted by the GELU error
//           a |2^P - Phi(-a)| (max 4.6e-6 abs, far below the 16-bit store): ONE MUFU on a short chain,
//           6 FMA-pipe + 1 FMNMX + 1 MUFU; leading coefficient negative, so P -> -inf for large |x|       7.5 ms  (default)
//   not kept: tanh.approx of a fitted odd polynomial (7.0 ms, but MUFU.TANH's 2^-11 error doubles the map error on the
//   negative side), the same fit as x * sigmoid on ex2 + rcp (no gain), A&S 7.1.28 with one rcp and a 16th power
//   (slower: 10 dependent FMA-pipe operations before the MUFU).
#ifndef AACLIP_GELU_FORM
#define AACLIP_GELU_FORM 5
#endif
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
#if AACLIP_GELU_FORM == 5
  float p = fmaf(-0.000417569448472932f, ax, 0.006652015261352062f);
  p = fmaf(p, ax, -0.05059274658560753f);
  p = fmaf(p, ax, -0.46160048246383667f);
  p = fmaf(p, ax, -1.1498371362686157f);
  p = fmaf(p, ax, -1.0002421140670776f);
  float q;                                                         // = Phi(-|x|)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(p));
#else
  //   u = |x| sqrt(log2(e)/2)  so that  exp(-z^2) = ex2(-u^2)  and  1 + p z = 1 + (p / sqrt(log2 e)) u;
  //   q = Phi(-|x|) = (poly t) e / 2 (the 1/2 folded into the coefficients)
  const float u = ax * 0.84932180028801904f;                       // sqrt(log2(e) / 2)
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.27273748f, u, 1.0f)));   // p / sqrt(log2 e), p = 0.3275911
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-u * u));
  const float q = (poly * t) * e;                                  // = (1 - erf(z)) / 2 = Phi(-|x|)
#endif
  return fmaf(-ax, q, fmaxf(x, 0.0f));
}
__device__ __forceinline__ float quick_gelu(float x) { return x / (1.0f + __expf(-1.702f * x)); }

// Fused epilogue on one float4 = accumulator values of (row, columns n..n+3), executed in the "coalesced
// domain": after the smem transpose a quarter-warp owns 128 contiguous bytes of one output row.
template <int EPI>
__host__ __device__ constexpr bool epi_has_bias() {
  return EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_QGELU_BF16 ||
         EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_F32 || EPI == EPI_LN_BIAS_BF16 ||
         EPI == EPI_LN_BIAS_GELU_BF16 || EPI == EPI_LN_BIAS_QGELU_BF16 || EPI == EPI_LN_F32_TOK ||
         EPI == EPI_LN_LEAKY_F32_TOK || EPI == EPI_BIAS_RESID_HL;
}

template <int EPI>
__host__ __device__ constexpr bool epi_is_resid() {    // adds the residual row the caller loads
  return EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_RESID_HL;
}

template <int EPI>
__host__ __device__ constexpr bool epi_is_tok() {      // token-tap projections: CLS rows dropped, output rows re-based
  return EPI == EPI_LN_F32_TOK || EPI == EPI_LN_LEAKY_F32_TOK;
}

template <int EPI>
__host__ __device__ constexpr bool epi_has_stats() {   // emits per-row (sum, sumsq) partials of what it wrote
  return EPI == EPI_BIAS_RESID_F32 || EPI == EPI_LEAKY_16_SQ || EPI == EPI_BIAS_RESID_HL;
}

template <int EPI>
__host__ __device__ constexpr bool epi_is_ln() {
  return EPI == EPI_LN_BIAS_BF16 || EPI == EPI_LN_BIAS_GELU_BF16 || EPI == EPI_LN_BIAS_QGELU_BF16 ||
         EPI == EPI_LN_F32_TOK || EPI == EPI_LN_LEAKY_F32_TOK;
}

// `b` = bias[n..n+3] (and `cs` = colsum[n..n+3]) are loaded once per 32-column chunk by the caller; `mean` / `rstd`
// are the LayerNorm statistics of this row (EPI_LN_* only). Returns the fp32 value written (EPI_BIAS_RESID_F32).
// `ok` predicates the STORES only (row < M): the math runs unconditionally so that the 8 rows a thread handles per
// chunk stay in one basic block and their dependency chains interleave (a branch per row serialised them: ~95
// dependent instructions per 4 elements with 2 warps per scheduler made the c_fc epilogue slower than its MMAs).
template <int EPI, bool BF16>
__device__ __forceinline__ float4 epilogue_apply(bool ok, float4 v, int row, int n, const GemmEpi& e, float4 b,
                                                 float4 resid = make_float4(0.f, 0.f, 0.f, 0.f),
                                                 float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), float mean = 0.f,
                                                 float rstd = 1.f) {
  if constexpr (epi_is_ln<EPI>()) {
    v.x = fmaf(rstd, fmaf(-mean, cs.x, v.x), b.x); v.y = fmaf(rstd, fmaf(-mean, cs.y, v.y), b.y);
    v.z = fmaf(rstd, fmaf(-mean, cs.z, v.z), b.z); v.w = fmaf(rstd, fmaf(-mean, cs.w, v.w), b.w);
  } else if constexpr (epi_has_bias<EPI>()) {
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  if constexpr (EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_LN_BIAS_GELU_BF16) {
    v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
  }
  if constexpr (EPI == EPI_BIAS_QGELU_BF16 || EPI == EPI_LN_BIAS_QGELU_BF16) {
    v.x = quick_gelu(v.x); v.y = quick_gelu(v.y); v.z = quick_gelu(v.z); v.w = quick_gelu(v.w);
  }
  if constexpr (EPI == EPI_LEAKY_F32 || EPI == EPI_LEAKY_16_SQ || EPI == EPI_LN_LEAKY_F32_TOK) {
    v.x = v.x > 0.f ? v.x : 0.01f * v.x; v.y = v.y > 0.f ? v.y : 0.01f * v.y;
    v.z = v.z > 0.f ? v.z : 0.01f * v.z; v.w = v.w > 0.f ? v.w : 0.01f * v.w;
  }
  if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_QGELU_BF16 ||
                (epi_is_ln<EPI>() && !epi_is_tok<EPI>()) || EPI == EPI_LEAKY_16_SQ) {
    const uint2 w = make_uint2(ptx::pack16<BF16>(v.x, v.y), ptx::pack16<BF16>(v.z, v.w));
    if (ok) *reinterpret_cast<uint2*>(e.out_16 + static_cast<size_t>(row) * e.ldo + n) = w;
  } else if constexpr (EPI == EPI_BIAS_RESID_HL) {
    // `resid` carries the raw words of the residual row: (hi.x, hi.y, lo.x, lo.y) of these four elements. Added with the
    // mixed-precision FHADD (16-bit operand + fp32 -> fp32): two instructions per element, no conversions.
    {
      const uint32_t h0 = __float_as_uint(resid.x), h1 = __float_as_uint(resid.y);
      const uint32_t l0 = __float_as_uint(resid.z), l1 = __float_as_uint(resid.w);
      v.x = ptx::add16<BF16, 0>(l0, ptx::add16<BF16, 0>(h0, v.x));
      v.y = ptx::add16<BF16, 1>(l0, ptx::add16<BF16, 1>(h0, v.y));
      v.z = ptx::add16<BF16, 0>(l1, ptx::add16<BF16, 0>(h1, v.z));
      v.w = ptx::add16<BF16, 1>(l1, ptx::add16<BF16, 1>(h1, v.w));
    }
    uint2 hi, lo;
    ptx::pack16_hilo<BF16>(v.x, v.y, hi.x, lo.x);
    ptx::pack16_hilo<BF16>(v.z, v.w, hi.y, lo.y);
    if (ok) {
      *reinterpret_cast<uint2*>(e.out_16 + static_cast<size_t>(row) * e.ldo + n) = hi;
      *reinterpret_cast<uint2*>(e.out_lo + static_cast<size_t>(row) * e.ldo + n) = lo;
    }
  } else if constexpr (EPI == EPI_BIAS_RESID_F32) {
    v.x += resid.x; v.y += resid.y; v.z += resid.z; v.w += resid.w;   // residual loaded by the caller
    const uint2 w = make_uint2(ptx::pack16<BF16>(v.x, v.y), ptx::pack16<BF16>(v.z, v.w));
    if (ok) {
#ifndef AACLIP_GEMM_DBG_NOXSTORE  // what-if (timing only, wrong results): no fp32 residual write
      *reinterpret_cast<float4*>(e.out_f32 + static_cast<size_t>(row) * e.ldo + n) = v;
#endif
#ifndef AACLIP_GEMM_DBG_NOX16     // what-if (timing only, wrong results): no 16-bit copy
      if (e.out_16 != nullptr) *reinterpret_cast<uint2*>(e.out_16 + static_cast<size_t>(row) * e.ldo + n) = w;
#endif
    }
  } else {
    size_t orow = static_cast<size_t>(row);
    if constexpr (EPI == EPI_PATCH_F32) {
      const int rowc = ok ? row : 0;                                  // keep the positional-embedding load in range
      const int bi = rowc / e.patches, p = rowc - bi * e.patches;
      orow = static_cast<size_t>(bi) * (e.patches + 1) + 1 + p;
      const float4 q = __ldg(reinterpret_cast<const float4*>(e.pos + static_cast<size_t>(p + 1) * e.ldo + n));
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    if (ok) *reinterpret_cast<float4*>(e.out_f32 + orow * e.ldo + n) = v;
  }
  return v;
}

// =========================================================================================================
// CTA-pair variant (tcgen05 cta_group::2): one cluster of 2 CTAs computes a 256 x BN output tile.
//   CTA r of the pair owns rows [256*m_pair + 128*r, +128) of C (its 128 TMEM lanes) and stages, per K step,
//   its own 128 x 64 slice of A and HALF of the W tile (BN/2 rows) -- the tensor cores of both SMs read both
//   halves, so every W byte is fetched from L2 and written to shared memory once per PAIR (2/3 of the single-CTA
//   kernel's operand traffic at BN = 256) and the per-SM shared-memory read rate drops below the MMA rate.
//   Only the leader (cluster rank 0) issues MMAs; TMA completions of both CTAs are counted on the leader's `full`
//   barriers; tcgen05.commit multicasts to the `empty` / `tmem_full` barriers of both CTAs; the epilogue warps of
//   both CTAs arrive on the leader's `tmem_empty`.
template <int BN_>
struct GemmPairCfg {
  static constexpr int BM = 128, BN = BN_, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;             // 16 KiB
  static constexpr int B_BYTES = (BN / 2) * BK * 2;       // 16 KiB (BN 256) / 8 KiB (BN 128): this CTA's half
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
#ifndef AACLIP_GEMM_RESID_PREFETCH
#define AACLIP_GEMM_RESID_PREFETCH 1
#endif
#ifndef AACLIP_GEMM_RESID_PREFETCH_MAXKB
#define AACLIP_GEMM_RESID_PREFETCH_MAXKB 32   // K / 64 up to which the residual rows are L2-prefetched at tile start
#endif
#ifndef AACLIP_GEMM_EPI_WARPS
#define AACLIP_GEMM_EPI_WARPS 8
#endif
  // Epilogue warps: EPI_WARPS/4 per TMEM lane quarter, splitting the accumulator columns. 16 warps (5-stage ring) run
  // the c_fc epilogue 12 % faster on L2-resident inputs but lose 11 % inside the real step (HBM-fed, power-capped:
  // 767 vs 809 images/s), so 8 warps with the deeper 6-stage ring is the default.
  static constexpr int EPI_WARPS = AACLIP_GEMM_EPI_WARPS;
  static constexpr int COL_GROUPS = EPI_WARPS / 4;        // warps sharing a lane quarter split the columns
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int STAGES = (EPI_WARPS == 16) ? 5 : 6;   // 160 / 192 KiB of operand ring (+ 4 KiB staging per epilogue warp)
  static constexpr int TMEM_COLS = 2 * BN;                // double-buffered accumulator
  static constexpr int EPI_STAGE_BYTES = 32 * 128;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 512 + EPI_WARPS * EPI_STAGE_BYTES + 1024;
};

template <int EPI, bool BF16, int BN>
// (18 warps are allocated like 20: 96 registers per thread at most -- __maxnreg__(112) is refused at launch)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GemmPairCfg<BN>::THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M, int N,
                 int K, GemmEpi epi) {
  using C = GemmPairCfg<BN>;
  constexpr int BM = C::BM, BK = C::BK, STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, A_BYTES = C::A_BYTES;
  constexpr int EPI_WARPS = C::EPI_WARPS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();           // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int m_pairs = (M + 2 * BM - 1) / (2 * BM);
  const int n_tiles = N / BN;
  const int num_tiles = m_pairs * n_tiles;
  const int num_kb = K / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);                    // leader's producer: arrive.expect_tx (bytes of both CTAs)
      ptx::mbar_init(&empty_bar[s], 1);                   // multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full_bar[s], 1);               // multicast tcgen05.commit
      ptx::mbar_init(&tmem_empty_bar[s], 2 * EPI_WARPS);  // epilogue warps of both CTAs (used on the leader only)
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair<C::TMEM_COLS>(tmem_base_smem);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();                                // peer barriers initialised, both TMEM allocations done
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  // programmatic dependent launch: everything above (barriers, TMEM, tensor maps) ran while the previous kernel of the
  // stream was still draining; nothing below may touch global memory before that kernel has completed
  ptx::pdl_trigger();
  ptx::pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (one per CTA) =====================
    // The whole warp runs the loop (warp-uniform control flow and addresses -> plain uniform-register moves for the
    // TMA / MMA operands); one elected lane issues. Under `if (lane == 0)` ptxas cannot prove uniformity and wraps
    // every UTMALDG / UTCHMMA in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop, which made the single issuing thread
    // the bottleneck of the main loop (ncu: tensor pipe 76 % active with no barrier stall anywhere).
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t full0_leader = ptx::mapa_u32(ptx::smem_u32(&full_bar[0]), 0);
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m_lin = tile / n_tiles, n_blk = tile - m_lin * n_tiles;
      const int m_pair = epi.descending ? m_pairs - 1 - m_lin : m_lin;
      const int m_row = (2 * m_pair + static_cast<int>(rank)) * BM;
      const int n_row = n_blk * BN + static_cast<int>(rank) * (BN / 2);
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (ptx::elect_one()) {
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          ptx::tma_load_2d_pair(sa, &tmap_a, full0_leader + stage * 8, kb * BK, m_row);
          ptx::tma_load_2d_pair(sa + A_BYTES, &tmap_b, full0_leader + stage * 8, kb * BK, n_row);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_16(2 * BM, BN, 0, BF16);
      const uint32_t smem_base = ptx::smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t a_desc = ptx::make_sw128_desc(sa);
          const uint64_t b_desc = ptx::make_sw128_desc(sa + A_BYTES);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              ptx::tc_mma_f16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            ptx::tc_commit_pair(&empty_bar[stage]);                          // frees the slot in both CTAs
            if (kb == num_kb - 1) ptx::tc_commit_pair(&tmem_full_bar[acc]);  // accumulators complete in both CTAs
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    constexpr int COLS_PER_WARP = BN / C::COL_GROUPS;   // warps sharing a lane quarter split the accumulator columns
    constexpr int CHUNKS = COLS_PER_WARP / 32;
    const int e = warp - 2;
    const int q = warp & 3;
    const int cg = e >> 2;          // column group of this warp
    const uint32_t stage_u32 = ptx::smem_u32(smem + STAGES * STAGE_BYTES + 512 + e * C::EPI_STAGE_BYTES);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m_lin = tile / n_tiles, n_blk = tile - m_lin * n_tiles;
      const int m_pair = epi.descending ? m_pairs - 1 - m_lin : m_lin;
      const int row0 = (2 * m_pair + static_cast<int>(rank)) * BM + q * 32;
      if constexpr (EPI == EPI_BIAS_RESID_HL && AACLIP_GEMM_RESID_PREFETCH) {
        // (hi, lo) form: COLS_PER_WARP 16-bit elements per plane and row = COLS_PER_WARP / 64 lines of 128 B each
        if (num_kb <= AACLIP_GEMM_RESID_PREFETCH_MAXKB && row0 + lane < M) {
          const size_t off = static_cast<size_t>(row0 + lane) * epi.ldo + n_blk * BN + cg * COLS_PER_WARP;
#pragma unroll
          for (int i = 0; i < COLS_PER_WARP / 64; ++i) {
            ptx::prefetch_l2(epi.out_16 + off + i * 64);
            ptx::prefetch_l2(epi.out_lo + off + i * 64);
          }
        }
      }
      if constexpr (EPI == EPI_BIAS_RESID_F32 && AACLIP_GEMM_RESID_PREFETCH)
        // The residual rows this warp will add come from HBM (the stream is 180 MB, far beyond L2): pull them into L2
        // now, while the MMAs of this tile are still running, so that the loads in the chunk loop below are L2 hits
        // (out-proj 3.39 -> 3.18 ms/step; skipped for long-K tiles, whose 20 us main loop outlives the prefetch).
        // lane l: row row0 + l, COLS_PER_WARP fp32 = COLS_PER_WARP / 32 lines of 128 B.
        // (Round 2: requesting the lines one TILE ahead instead, and fetching bias + residual one chunk ahead into
        // registers, both measured slower -- tools/experiments/README.md. ncu: the out-projection moves 4.5 TB/s of
        // DRAM traffic in isolation, 70 % of the measured copy bandwidth; it is bandwidth-, not latency-bound.)
        if (num_kb <= AACLIP_GEMM_RESID_PREFETCH_MAXKB && row0 + lane < M) {
          const float* rp = epi.out_f32 + static_cast<size_t>(row0 + lane) * epi.ldo + n_blk * BN + cg * COLS_PER_WARP;
#pragma unroll
          for (int i = 0; i < COLS_PER_WARP / 32; ++i) ptx::prefetch_l2(rp + i * 32);
        }
      // EPI_LN_*: lane l finishes the LayerNorm statistics of row row0 + l from its partials (contiguous, fixed order)
      // BEFORE waiting for the accumulator -- the loads' latency (the statistics were written by the previous kernel:
      // an L2 or HBM round trip) runs under this tile's MMAs instead of on the epilogue's critical path.
      float mean = 0.f, rstd = 0.f;
      if constexpr (epi_is_ln<EPI>()) {
        if (row0 + lane < M) {
          const float4* sp = reinterpret_cast<const float4*>(epi.stats + static_cast<size_t>(row0 + lane) * epi.n_partials * 2);
          float sx = 0.f, sq = 0.f;
          for (int i = 0; i < epi.n_partials / 2; ++i) {      // two (sum, sumsq) partials per 16-byte load
            const float4 p = __ldg(sp + i);
            sx += p.x + p.z;
            sq += p.y + p.w;
          }
          if (epi.n_partials & 1) {
            const float2 p = __ldg(reinterpret_cast<const float2*>(sp) + epi.n_partials - 1);
            sx += p.x; sq += p.y;
          }
          mean = sx * epi.inv_width;
          rstd = rsqrtf(fmaxf(sq * epi.inv_width - mean * mean, 0.f) + epi.eps);
        }
      }
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + cg * COLS_PER_WARP;
      const int sub = lane >> 3, unit = lane & 7;
      // coalesced domain: this thread handles rows row0 + j*4 + sub (j < 8) in every 32-column chunk
      float mean_j[8], rstd_j[8];   // EPI_LN_*: LayerNorm statistics of its 8 rows
      float st_s[8], st_q[8];       // epi_has_stats + stats_out: running sum / sum of squares of the rows written
      if constexpr (epi_is_ln<EPI>()) {
        // the coalesced-domain owners of a row pick its statistics up with one shuffle each
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          mean_j[j] = __shfl_sync(0xffffffffu, mean, j * 4 + sub);
          rstd_j[j] = __shfl_sync(0xffffffffu, rstd, j * 4 + sub);
        }
      }
      if constexpr (epi_has_stats<EPI>()) {
#pragma unroll
        for (int j = 0; j < 8; ++j) st_s[j] = st_q[j] = 0.f;
      }
      int tok_row[8];               // epi_is_tok: output row of each of the 8 rows, -1 = dropped (CLS row or beyond M)
      if constexpr (epi_is_tok<EPI>()) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rr = row0 + j * 4 + sub;
          const int bi = rr / (epi.patches + 1), t = rr - bi * (epi.patches + 1);
          tok_row[j] = (rr < M && t != 0) ? bi * epi.patches + t - 1 : -1;
        }
      }
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        const int n = n_blk * BN + cg * COLS_PER_WARP + c * 32 + unit * 4;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), cs4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (epi_has_bias<EPI>()) bias4 = __ldg(reinterpret_cast<const float4*>(epi.bias + n));
        if constexpr (epi_is_ln<EPI>()) cs4 = __ldg(reinterpret_cast<const float4*>(epi.colsum + n));
        // residual rows of this chunk: rows j = 0..3 are fetched before the TMEM load, rows 4..7 right after the
        // staging stores (when the 32 accumulator registers are dead) -- 4 + 4 coalesced loads in flight per thread
        float4 resid_a[4], resid_b[4];
        auto load_resid = [&](float4 (&dst)[4], int j0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rr = row0 + (j0 + j) * 4 + sub;
#ifdef AACLIP_GEMM_DBG_NORESID   // what-if (timing only, wrong results): no residual read
            dst[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#else
            if constexpr (EPI == EPI_BIAS_RESID_HL) {
              uint2 hi = make_uint2(0u, 0u), lo = make_uint2(0u, 0u);
              if (rr < M) {
                hi = *reinterpret_cast<const uint2*>(epi.out_16 + static_cast<size_t>(rr) * epi.ldo + n);
                lo = *reinterpret_cast<const uint2*>(epi.out_lo + static_cast<size_t>(rr) * epi.ldo + n);
              }
              dst[j] = make_float4(__uint_as_float(hi.x), __uint_as_float(hi.y), __uint_as_float(lo.x), __uint_as_float(lo.y));
            } else {
              dst[j] = (rr < M) ? *reinterpret_cast<const float4*>(epi.out_f32 + static_cast<size_t>(rr) * epi.ldo + n)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#endif
          }
        };
        if constexpr (epi_is_resid<EPI>()) load_resid(resid_a, 0);
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(t_base + c * 32, r);
        ptx::tmem_ld_wait();
        if (c == CHUNKS - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&tmem_empty_bar[acc]), 0));   // leader's barrier
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          ptx::sts128(stage_u32 + lane * 128 + ((u ^ (lane & 7)) << 4), r[4 * u], r[4 * u + 1], r[4 * u + 2],
                      r[4 * u + 3]);
        if constexpr (epi_is_resid<EPI>()) load_resid(resid_b, 4);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rr = j * 4 + sub;
          const float4 v = ptx::lds128(stage_u32 + rr * 128 + ((unit ^ (rr & 7)) << 4));
          const bool ok = row0 + rr < M;
          if constexpr (epi_is_resid<EPI>()) {
            const float4 x = epilogue_apply<EPI, BF16>(ok, v, row0 + rr, n, epi, bias4, j < 4 ? resid_a[j & 3] : resid_b[j & 3]);
            st_s[j] += (x.x + x.y) + (x.z + x.w);      // rows >= M accumulate garbage that is never stored
            st_q[j] += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
          } else if constexpr (EPI == EPI_LEAKY_16_SQ) {
            const float4 x = epilogue_apply<EPI, BF16>(ok, v, row0 + rr, n, epi, bias4);
            st_s[j] += (x.x + x.y) + (x.z + x.w);
            st_q[j] += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
          } else if constexpr (epi_is_tok<EPI>()) {
            epilogue_apply<EPI, BF16>(tok_row[j] >= 0, v, tok_row[j], n, epi, bias4, make_float4(0.f, 0.f, 0.f, 0.f), cs4,
                                      mean_j[j], rstd_j[j]);
          } else if constexpr (epi_is_ln<EPI>()) {
            epilogue_apply<EPI, BF16>(ok, v, row0 + rr, n, epi, bias4, make_float4(0.f, 0.f, 0.f, 0.f), cs4, mean_j[j],
                                      rstd_j[j]);
          } else {
            epilogue_apply<EPI, BF16>(ok, v, row0 + rr, n, epi, bias4);
          }
        }
        __syncwarp();
      }
      if constexpr (epi_has_stats<EPI>()) {
        if (epi.stats_out != nullptr) {
          // partial (sum, sumsq) of this warp's COLS_PER_WARP columns for each of its 32 rows. The 8 lanes of a row
          // group (unit = 0..7) hold 8 rows x 2 values each; a halving butterfly (each xor step exchanges half of the
          // still-live values) leaves lane `unit` with the complete sums of row j = unit: 14 shuffles instead of 48,
          // every lane stores one row. Fixed order: bit-reproducible and batch-invariant.
          const int np = N / COLS_PER_WARP, pi = n_blk * C::COL_GROUPS + cg;
          float v[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) { v[j] = st_s[j]; v[8 + j] = st_q[j]; }
          // step xor 4: lanes with (unit & 4) keep rows 4..7, the others rows 0..3
          {
            const bool hi = (unit & 4) != 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float send_s = hi ? v[j] : v[4 + j], send_q = hi ? v[8 + j] : v[12 + j];
              const float got_s = __shfl_xor_sync(0xffffffffu, send_s, 4), got_q = __shfl_xor_sync(0xffffffffu, send_q, 4);
              v[j] = (hi ? v[4 + j] : v[j]) + got_s;
              v[8 + j] = (hi ? v[12 + j] : v[8 + j]) + got_q;
            }
          }
          // step xor 2: of the 4 live rows, lanes with (unit & 2) keep the upper two
          {
            const bool hi = (unit & 2) != 0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float send_s = hi ? v[j] : v[2 + j], send_q = hi ? v[8 + j] : v[10 + j];
              const float got_s = __shfl_xor_sync(0xffffffffu, send_s, 2), got_q = __shfl_xor_sync(0xffffffffu, send_q, 2);
              v[j] = (hi ? v[2 + j] : v[j]) + got_s;
              v[8 + j] = (hi ? v[10 + j] : v[8 + j]) + got_q;
            }
          }
          // step xor 1
          {
            const bool hi = (unit & 1) != 0;
            const float send_s = hi ? v[0] : v[1], send_q = hi ? v[8] : v[9];
            const float got_s = __shfl_xor_sync(0xffffffffu, send_s, 1), got_q = __shfl_xor_sync(0xffffffffu, send_q, 1);
            v[0] = (hi ? v[1] : v[0]) + got_s;
            v[8] = (hi ? v[9] : v[8]) + got_q;
          }
          const int rr = row0 + unit * 4 + sub;      // lane `unit` now owns row j = unit of its sub-group
          if (rr < M) reinterpret_cast<float2*>(epi.stats_out)[static_cast<size_t>(rr) * np + pi] = make_float2(v[0], v[8]);
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();         // no CTA may free its TMEM / exit while its peer still multicasts into it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair<C::TMEM_COLS>(tmem_base);
  }
}

}  // namespace aaclip
