// C-ABI + host orchestration of the B200-native AA-CLIP hot path. See include/aaclip.h for the contract.
// Compiled with: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/aaclip.h"
#include "attention_tcgen05.cuh"
#include "preprocess.cuh"
#include "elementwise.cuh"
#include "fp32_path.cuh"
#include "gemm_tcgen05.cuh"
#include "metrics.cuh"

namespace {

using namespace aaclip;

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) return fail(AACLIP_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                       __FILE__, __LINE__);                                              \
  } while (0)

#define LAUNCH_CHECK(what)                                                                               \
  do {                                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                                 \
    if (_e != cudaSuccess) return fail(AACLIP_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(_e)); \
  } while (0)

// ------------------------------------------------------------------------------------------ TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

// fp16/bf16 tensor, innermost dim contiguous; rank 2 or 3; 128-byte swizzle; OOB reads return zero.
int make_tmap_16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, bool bf16) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(AACLIP_ENOSYS, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gdim[3];
  cuuint64_t gstr[2];
  cuuint32_t bx[3], estr[3];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AACLIP_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return AACLIP_OK;
}

// Per-device state. Function attributes (the > 48 KiB dynamic shared-memory opt-in) and the SM count belong to the
// CURRENT device's context: a process that drives several GPUs (AaclipEngine(device="cuda:1")) must see them set on
// each. Everything is keyed by device ordinal; the flags are atomics, so first calls from several threads may both set
// an attribute (idempotent) but never skip it.
constexpr int kMaxDevices = 64;
std::atomic<int> g_sms[kMaxDevices];

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

int ensure_device(int* num_sms = nullptr) {
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return fail(AACLIP_EINVAL, "device ordinal %d outside 0..%d", dev, kMaxDevices - 1);
  int sms = g_sms[dev].load(std::memory_order_relaxed);
  if (sms == 0) {
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(AACLIP_ENOSYS, "device '%s' is sm_%d%d; this library is sm_100a only", prop.name,
                                      prop.major, prop.minor);
    sms = prop.multiProcessorCount;
    g_sms[dev].store(sms, std::memory_order_relaxed);
  }
  if (num_sms) *num_sms = sms;
  return AACLIP_OK;
}

// One flag word per kernel instantiation: bit d set = cudaFuncSetAttribute has run on device d.
struct PerDeviceOnce {
  std::atomic<uint64_t> mask{0};
  bool needed() const { return (mask.load(std::memory_order_acquire) >> (current_device() & 63) & 1ull) == 0; }
  void done() { mask.fetch_or(1ull << (current_device() & 63), std::memory_order_release); }
};

template <typename K>
int opt_in_smem(PerDeviceOnce& once, K kernel, size_t bytes) {
  if (bytes <= 48 * 1024 || !once.needed()) return AACLIP_OK;
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  once.done();
  return AACLIP_OK;
}

// Programmatic dependent launch (ptx.cuh: pdl_wait / pdl_trigger). AACLIP_PDL=0 turns the attribute off (plain stream
// order), e.g. to A/B the overlap.
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("AACLIP_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------ kernel launchers
template <int EPI, bool BF16>
int launch_gemm_t(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const GemmEpi& epi,
                  cudaStream_t st) {
  using C = GemmPairCfg<256>;
  static PerDeviceOnce once;
  if (int rc = opt_in_smem(once, gemm_pair_kernel<EPI, BF16, 256>, C::SMEM_BYTES)) return rc;
  int sms = 0;
  if (int rc = ensure_device(&sms)) return rc;
  // persistent: one CTA pair (cluster of 2) per two SMs, each looping over 256 x 256 output tiles
  const int tiles = ((M + 2 * C::BM - 1) / (2 * C::BM)) * (N / C::BN);
  const int pairs = tiles < sms / 2 ? tiles : sms / 2;
  CUDA_TRY(launch_pdl(gemm_pair_kernel<EPI, BF16, 256>, dim3(2 * pairs), dim3(C::THREADS), C::SMEM_BYTES, st, ta, tb, M, N,
                      K, epi));
  return AACLIP_OK;
}

int launch_gemm(int epi_mode, bool bf16, const void* A, const void* W, int M, int N, int K, const GemmEpi& epi,
                cudaStream_t st) {
  using C = GemmPairCfg<256>;
  if (int rc = ensure_device()) return rc;
  if (M <= 0 || N <= 0 || K <= 0 || N % C::BN != 0 || K % C::BK != 0)
    return fail(AACLIP_EINVAL, "gemm: unsupported shape M=%d N=%d K=%d (need N%%256==0, K%%64==0)", M, N, K);
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return fail(AACLIP_EINVAL, "gemm: operands must be 16-byte aligned");
  CUtensorMap ta, tb;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {C::BK, C::BM};                 // this CTA's 128 rows of A
    if (int rc = make_tmap_16(&ta, A, 2, dims, str, box, bf16)) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t str[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {C::BK, C::BN / 2};             // this CTA's half of the 256-row W tile
    if (int rc = make_tmap_16(&tb, W, 2, dims, str, box, bf16)) return rc;
  }
#define AACLIP_GEMM_CASE(E)                                                              \
  case E:                                                                                \
    return bf16 ? launch_gemm_t<E, true>(ta, tb, M, N, K, epi, st) : launch_gemm_t<E, false>(ta, tb, M, N, K, epi, st);
  switch (epi_mode) {
    AACLIP_GEMM_CASE(EPI_BIAS_BF16)
    AACLIP_GEMM_CASE(EPI_BIAS_GELU_BF16)
    AACLIP_GEMM_CASE(EPI_BIAS_QGELU_BF16)
    AACLIP_GEMM_CASE(EPI_BIAS_RESID_F32)
    AACLIP_GEMM_CASE(EPI_F32)
    AACLIP_GEMM_CASE(EPI_LEAKY_F32)
    AACLIP_GEMM_CASE(EPI_PATCH_F32)
    AACLIP_GEMM_CASE(EPI_BIAS_F32)
    AACLIP_GEMM_CASE(EPI_LN_BIAS_BF16)
    AACLIP_GEMM_CASE(EPI_LN_BIAS_GELU_BF16)
    AACLIP_GEMM_CASE(EPI_LN_BIAS_QGELU_BF16)
    AACLIP_GEMM_CASE(EPI_LEAKY_16_SQ)
    AACLIP_GEMM_CASE(EPI_LN_F32_TOK)
    AACLIP_GEMM_CASE(EPI_LN_LEAKY_F32_TOK)
    AACLIP_GEMM_CASE(EPI_BIAS_RESID_HL)
#undef AACLIP_GEMM_CASE
    default: return fail(AACLIP_EINVAL, "gemm: unknown epilogue %d", epi_mode);
  }
}

// one round-robin slot counter for every attention launch of the process (both operand dtypes, every device)
std::atomic<unsigned> g_attn_next_slot{0};

template <bool BF16>
int launch_attention_t(const CUtensorMap& tq, void* out, int B, int N, int D, int causal, cudaStream_t st,
                       int descending) {
  static PerDeviceOnce once;
  if (int rc = opt_in_smem(once, attention_tcgen05_kernel<BF16>, attn_cfg::SMEM_BYTES)) return rc;
  int sms = 0;
  if (int rc = ensure_device(&sms)) return rc;
  // persistent CTAs draw (image, head, 256 query rows) items from a counter slot that the kernel leaves zeroed; slots
  // are handed out round-robin, so launches in flight on different streams (or captured in different graphs) do not
  // share one unless more than SCHED_SLOTS of them are pending at once
  AttnSched* sched = nullptr;
  CUDA_TRY(cudaGetSymbolAddress(reinterpret_cast<void**>(&sched), g_attn_sched));
  sched += g_attn_next_slot.fetch_add(1, std::memory_order_relaxed) % attn_cfg::SCHED_SLOTS;
  const int items = B * (D / attn_cfg::HD) * ((N + attn_cfg::BQ * attn_cfg::QT - 1) / (attn_cfg::BQ * attn_cfg::QT));
  // the output as a 3-D tensor [B][N][D] for the bulk stores: one 32-row x 64-column box per softmax warp and item
  CUtensorMap to;
  {
    const uint64_t dims[3] = {(uint64_t)D, (uint64_t)N, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)D * 2, (uint64_t)N * D * 2};
    const uint32_t box[3] = {attn_cfg::HD, 32, 1};
    if (int rc = make_tmap_16(&to, out, 3, dims, str, box, BF16)) return rc;
  }
  const float scale_log2 = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
  CUDA_TRY(launch_pdl(attention_tcgen05_kernel<BF16>, dim3(items < sms ? items : sms), dim3(attn_cfg::THREADS),
                      attn_cfg::SMEM_BYTES, st, tq, to, N, D, causal, scale_log2, B, sched, descending));
  return AACLIP_OK;
}

int launch_attention(bool bf16, const void* qkv, void* out, int B, int N, int D, int causal, cudaStream_t st,
                     int descending = 0) {
  if (int rc = ensure_device()) return rc;
  if (B <= 0 || N <= 0 || D <= 0 || D % 64 != 0) return fail(AACLIP_EINVAL, "attention: bad shape B=%d N=%d D=%d", B, N, D);
  CUtensorMap tq;
  const uint64_t dims[3] = {(uint64_t)3 * D, (uint64_t)N, (uint64_t)B};
  const uint64_t str[2] = {(uint64_t)3 * D * 2, (uint64_t)N * 3 * D * 2};
  const uint32_t box[3] = {attn_cfg::HD, attn_cfg::BKV, 1};
  if (int rc = make_tmap_16(&tq, qkv, 3, dims, str, box, bf16)) return rc;
  return bf16 ? launch_attention_t<true>(tq, out, B, N, D, causal, st, descending)
              : launch_attention_t<false>(tq, out, B, N, D, causal, st, descending);
}

template <int D, bool BF16>
void launch_layernorm_t(const float* x, const float* g, const float* b, void* out16, float* out_f32, int rows,
                        int skipP, bool normalize, cudaStream_t st, float* stats_out, int np, uint16_t* out_lo) {
  const int threads = 256, grid = (rows + 7) / 8;
  uint16_t* o = reinterpret_cast<uint16_t*>(out16);
  if (normalize) launch_pdl(layernorm_kernel<D, true, BF16>, dim3(grid), dim3(threads), 0, st, x, g, b, o, out_f32, rows, skipP, 1e-5f, stats_out, np, out_lo);
  else launch_pdl(layernorm_kernel<D, false, BF16>, dim3(grid), dim3(threads), 0, st, x, g, b, o, out_f32, rows, skipP, 1e-5f, stats_out, np, out_lo);
}

int launch_layernorm(bool bf16, const float* x, const float* g, const float* b, void* out16, float* out_f32, int rows,
                     int D, int skipP, bool normalize, cudaStream_t st, float* stats_out = nullptr, int np = 0,
                     uint16_t* out_lo = nullptr) {
  if (rows <= 0) return AACLIP_OK;
  if (D == 1024) {
    if (bf16) launch_layernorm_t<1024, true>(x, g, b, out16, out_f32, rows, skipP, normalize, st, stats_out, np, out_lo);
    else launch_layernorm_t<1024, false>(x, g, b, out16, out_f32, rows, skipP, normalize, st, stats_out, np, out_lo);
  } else if (D == 768) {
    if (bf16) launch_layernorm_t<768, true>(x, g, b, out16, out_f32, rows, skipP, normalize, st, stats_out, np, out_lo);
    else launch_layernorm_t<768, false>(x, g, b, out16, out_f32, rows, skipP, normalize, st, stats_out, np, out_lo);
  } else {
    return fail(AACLIP_EINVAL, "layernorm: D=%d unsupported (768 or 1024)", D);
  }
  LAUNCH_CHECK("layernorm_kernel");
  return AACLIP_OK;
}

template <bool BF16>
__global__ void pack16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long R, long long C,
                              long long Cp) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= R * Cp) return;
  const long long r = i / Cp, c = i - r * Cp;
  dst[i] = ptx::cvt16<BF16>(c < C ? src[r * C + c] : 0.f);
}

__global__ void gather_cls_kernel(const float* __restrict__ x, float* __restrict__ out, int tokens, int D) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < D; i += blockDim.x) out[static_cast<size_t>(b) * D + i] = x[static_cast<size_t>(b) * tokens * D + i];
}

// one row per group of `tokens` rows of the (hi, lo) residual stream -> fp32 [groups, D]: row `pick[g]` of group g
// (pick == nullptr: row 0 = the CLS token, model/adapter.py:190)
template <bool BF16>
__global__ void gather_row_hl_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo,
                                     float* __restrict__ out, int tokens, int D, const int32_t* __restrict__ pick) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int g = blockIdx.x;
  const size_t row = static_cast<size_t>(g) * tokens + (pick != nullptr ? pick[g] : 0);
  const uint2* h = reinterpret_cast<const uint2*>(hi + row * D);
  const uint2* l = reinterpret_cast<const uint2*>(lo + row * D);
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x)
    reinterpret_cast<float4*>(out + static_cast<size_t>(g) * D)[i] = ptx::unpack_hilo4<BF16>(h[i], l[i]);
}

int launch_pack(bool bf16, const float* src, void* dst, long long R, long long C, long long Cp, cudaStream_t st) {
  const long long n = R * Cp;
  if (n <= 0) return AACLIP_OK;
  const unsigned grid = static_cast<unsigned>((n + 255) / 256);
  if (bf16) pack16_kernel<true><<<grid, 256, 0, st>>>(src, reinterpret_cast<uint16_t*>(dst), R, C, Cp);
  else pack16_kernel<false><<<grid, 256, 0, st>>>(src, reinterpret_cast<uint16_t*>(dst), R, C, Cp);
  LAUNCH_CHECK("pack16_kernel");
  return AACLIP_OK;
}

// fp32-accuracy mode: SIMT SGEMM / attention (fp32_path.cuh)
int launch_sgemm(int epi_mode, const float* A, const float* W, int M, int N, int K, const GemmEpi& epi, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0 || N % sgemm_cfg::BN != 0 || K % sgemm_cfg::BK != 0)
    return fail(AACLIP_EINVAL, "sgemm: unsupported shape M=%d N=%d K=%d (need N%%128==0, K%%16==0)", M, N, K);
  dim3 grid(N / sgemm_cfg::BN, (M + sgemm_cfg::BM - 1) / sgemm_cfg::BM);
#define AACLIP_SGEMM_CASE(E) \
  case E: sgemm_f32_kernel<E><<<grid, sgemm_cfg::THREADS, 0, st>>>(A, W, M, N, K, epi); break;
  switch (epi_mode) {
    AACLIP_SGEMM_CASE(EPI_BIAS_GELU_BF16)     // (names shared with the 16-bit epilogues; outputs are fp32 here)
    AACLIP_SGEMM_CASE(EPI_BIAS_QGELU_BF16)
    AACLIP_SGEMM_CASE(EPI_BIAS_RESID_F32)
    AACLIP_SGEMM_CASE(EPI_F32)
    AACLIP_SGEMM_CASE(EPI_LEAKY_F32)
    AACLIP_SGEMM_CASE(EPI_PATCH_F32)
    AACLIP_SGEMM_CASE(EPI_BIAS_F32)
#undef AACLIP_SGEMM_CASE
    default: return fail(AACLIP_EINVAL, "sgemm: unknown epilogue %d", epi_mode);
  }
  LAUNCH_CHECK("sgemm_f32_kernel");
  return AACLIP_OK;
}

int launch_attention_f32(const float* qkv, float* out, int B, int N, int D, int causal, cudaStream_t st) {
  if (B <= 0 || N <= 0 || D <= 0 || D % 64 != 0) return fail(AACLIP_EINVAL, "attention: bad shape B=%d N=%d D=%d", B, N, D);
  attention_f32_kernel<<<dim3((N + 127) / 128, D / 64, B), 128, 0, st>>>(qkv, out, N, D, causal);
  LAUNCH_CHECK("attention_f32_kernel");
  return AACLIP_OK;
}

// fp32 [R,C] -> fp32 [R,Cp] (destination pre-zeroed): the fp32-mode counterpart of launch_pack
int copy_pad_f32(const float* src, void* dst, long long R, long long C, long long Cp, cudaStream_t st) {
  CUDA_TRY(cudaMemcpy2DAsync(dst, Cp * sizeof(float), src, C * sizeof(float), C * sizeof(float), R,
                             cudaMemcpyDeviceToDevice, st));
  return AACLIP_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// =========================================================================================== transformer core
// Shared by the visual tower (AaclipHandle) and the text tower (AaclipTextHandle): per-block weights, LayerNorm
// folding, and the launch sequence of one pre-LN residual attention block (model/transformer.py:239-258).
struct BlockW {
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *b_qkv, *b_out, *b_fc, *b_proj;   // raw fp32 parameters (arena)
  uint16_t *w_qkv, *w_out, *w_fc, *w_proj;                               // 16-bit GEMM operands (arena)
  // ln_1 / ln_2 folded into the QKV / c_fc GEMMs (ln_fold_kernel): w_qkv = in_proj * gamma_1, w_fc = c_fc * gamma_2
  float *cs_qkv, *bf_qkv, *cs_fc, *bf_fc;                                // column sums of W', bias' (arena)
  float *stage_qkv = nullptr, *stage_fc = nullptr;                       // fp32 weights kept until the fold (cudaMalloc)
};

namespace {

struct ArenaPlan {
  size_t off = 0;
  template <typename T>
  size_t take(size_t n) {
    size_t o = off;
    off = align_up(off + n * sizeof(T), 256);
    return o;
  }
};

int64_t numel(const int64_t* shape, int ndim) {
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  return n;
}

void erase_key(std::vector<std::string>& v, const std::string& key) {
  for (size_t i = 0; i < v.size(); ++i)
    if (v[i] == key) { v.erase(v.begin() + i); return; }
}

struct BlockOff { size_t ln1g, ln1b, ln2g, ln2b, bqkv, bout, bfc, bproj, wqkv, wout, wfc, wproj, csq, bfq, csf, bff; };

// wb = bytes per GEMM-weight element: 2 (fp16 / bf16 operands) or 4 (fp32-accuracy mode: the uint16_t* weight
// pointers of BlockW then address fp32 data)
BlockOff plan_block(ArenaPlan& plan, size_t W, size_t F, size_t wb) {
  BlockOff o;
  o.ln1g = plan.take<float>(W); o.ln1b = plan.take<float>(W);
  o.ln2g = plan.take<float>(W); o.ln2b = plan.take<float>(W);
  o.bqkv = plan.take<float>(3 * W); o.bout = plan.take<float>(W);
  o.bfc = plan.take<float>(F); o.bproj = plan.take<float>(W);
  o.wqkv = plan.take<uint8_t>(3 * W * W * wb); o.wout = plan.take<uint8_t>(W * W * wb);
  o.wfc = plan.take<uint8_t>(F * W * wb); o.wproj = plan.take<uint8_t>(W * F * wb);
  o.csq = plan.take<float>(3 * W); o.bfq = plan.take<float>(3 * W);
  o.csf = plan.take<float>(F); o.bff = plan.take<float>(F);
  return o;
}

void bind_block(BlockW& b, uint8_t* a, const BlockOff& o) {
  b.ln1_g = (float*)(a + o.ln1g); b.ln1_b = (float*)(a + o.ln1b);
  b.ln2_g = (float*)(a + o.ln2g); b.ln2_b = (float*)(a + o.ln2b);
  b.b_qkv = (float*)(a + o.bqkv); b.b_out = (float*)(a + o.bout);
  b.b_fc = (float*)(a + o.bfc); b.b_proj = (float*)(a + o.bproj);
  b.w_qkv = (uint16_t*)(a + o.wqkv); b.w_out = (uint16_t*)(a + o.wout);
  b.w_fc = (uint16_t*)(a + o.wfc); b.w_proj = (uint16_t*)(a + o.wproj);
  b.cs_qkv = (float*)(a + o.csq); b.bf_qkv = (float*)(a + o.bfq);
  b.cs_fc = (float*)(a + o.csf); b.bf_fc = (float*)(a + o.bff);
}

const char* const kBlockTensors[] = {"ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias", "attn.in_proj_weight",
                                     "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias",
                                     "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias"};

// One tensor of a residual attention block ("ln_1.weight", "attn.in_proj_weight", ...). The weights that LayerNorm is
// folded into (in_proj, c_fc) are kept in fp32 until fold_blocks() runs at the first forward; after that the block's
// LayerNorm / in_proj / c_fc tensors are frozen (create a new handle to load different ones).
int set_block_tensor(BlockW& b, const std::string& t, const std::string& key, std::vector<std::string>& missing,
                     bool folded, bool bf16, bool f32, int64_t W, int64_t F, const char* name, const float* data,
                     int64_t n, cudaStream_t st) {
  auto copy_f32 = [&](float* dst, int64_t expect) -> int {
    if (n != expect) return fail(AACLIP_EINVAL, "%s: expected %lld elements, got %lld", name, (long long)expect, (long long)n);
    CUDA_TRY(cudaMemcpyAsync(dst, data, expect * sizeof(float), cudaMemcpyDeviceToDevice, st));
    erase_key(missing, key);
    return AACLIP_OK;
  };
  auto pack = [&](uint16_t* dst, int64_t R, int64_t C) -> int {
    if (n != R * C) return fail(AACLIP_EINVAL, "%s: expected %lld x %lld elements, got %lld", name, (long long)R, (long long)C, (long long)n);
    if (f32) {
      CUDA_TRY(cudaMemcpyAsync(dst, data, R * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else if (int rc = launch_pack(bf16, data, dst, R, C, C, st)) {
      return rc;
    }
    erase_key(missing, key);
    return AACLIP_OK;
  };
  auto stage = [&](float** slot, uint16_t* direct, int64_t R, int64_t C) -> int {
    if (f32) return pack(direct, R, C);          // fp32 mode: no LayerNorm folding, the weight is used as is
    if (n != R * C) return fail(AACLIP_EINVAL, "%s: expected %lld x %lld elements, got %lld", name, (long long)R, (long long)C, (long long)n);
    if (*slot == nullptr) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(slot), R * C * sizeof(float)));
    CUDA_TRY(cudaMemcpyAsync(*slot, data, R * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    erase_key(missing, key);
    return AACLIP_OK;
  };
  const bool fold_input = t == "ln_1.weight" || t == "ln_1.bias" || t == "ln_2.weight" || t == "ln_2.bias" ||
                          t == "attn.in_proj_weight" || t == "attn.in_proj_bias" || t == "mlp.c_fc.weight" ||
                          t == "mlp.c_fc.bias";
  if (fold_input && folded)
    return fail(AACLIP_ESTATE, "%s: LayerNorm has already been folded into this block's GEMM weights (first forward "
                               "ran); create a new handle to load different backbone weights", name);
  if (t == "ln_1.weight") return copy_f32(b.ln1_g, W);
  if (t == "ln_1.bias") return copy_f32(b.ln1_b, W);
  if (t == "ln_2.weight") return copy_f32(b.ln2_g, W);
  if (t == "ln_2.bias") return copy_f32(b.ln2_b, W);
  if (t == "attn.in_proj_bias") return copy_f32(b.b_qkv, 3 * W);
  if (t == "attn.out_proj.bias") return copy_f32(b.b_out, W);
  if (t == "mlp.c_fc.bias") return copy_f32(b.b_fc, F);
  if (t == "mlp.c_proj.bias") return copy_f32(b.b_proj, W);
  if (t == "attn.in_proj_weight") return stage(&b.stage_qkv, b.w_qkv, 3 * W, W);
  if (t == "attn.out_proj.weight") return pack(b.w_out, W, W);
  if (t == "mlp.c_fc.weight") return stage(&b.stage_fc, b.w_fc, F, W);
  if (t == "mlp.c_proj.weight") return pack(b.w_proj, W, F);
  return fail(AACLIP_EINVAL, "%s: unknown block tensor", name);
}

int fold_blocks(std::vector<BlockW>& blk, bool bf16, int W, int F, cudaStream_t st) {
  for (BlockW& b : blk) {
    if (b.stage_qkv == nullptr || b.stage_fc == nullptr) return fail(AACLIP_ESTATE, "fold_blocks: staged weights missing");
    if (bf16) {
      ln_fold_kernel<true><<<(3 * W + 7) / 8, 256, 0, st>>>(b.stage_qkv, b.b_qkv, b.ln1_g, b.ln1_b, b.w_qkv, b.cs_qkv, b.bf_qkv, 3 * W, W);
      ln_fold_kernel<true><<<(F + 7) / 8, 256, 0, st>>>(b.stage_fc, b.b_fc, b.ln2_g, b.ln2_b, b.w_fc, b.cs_fc, b.bf_fc, F, W);
    } else {
      ln_fold_kernel<false><<<(3 * W + 7) / 8, 256, 0, st>>>(b.stage_qkv, b.b_qkv, b.ln1_g, b.ln1_b, b.w_qkv, b.cs_qkv, b.bf_qkv, 3 * W, W);
      ln_fold_kernel<false><<<(F + 7) / 8, 256, 0, st>>>(b.stage_fc, b.b_fc, b.ln2_g, b.ln2_b, b.w_fc, b.cs_fc, b.bf_fc, F, W);
    }
    LAUNCH_CHECK("ln_fold_kernel");
  }
  CUDA_TRY(cudaStreamSynchronize(st));      // load-time only: the staging copies are released right away
  for (BlockW& b : blk) {
    cudaFree(b.stage_qkv); cudaFree(b.stage_fc);
    b.stage_qkv = b.stage_fc = nullptr;
  }
  return AACLIP_OK;
}

void free_block_staging(std::vector<BlockW>& blk) {
  for (BlockW& b : blk) {
    if (b.stage_qkv) cudaFree(b.stage_qkv);
    if (b.stage_fc) cudaFree(b.stage_fc);
    b.stage_qkv = b.stage_fc = nullptr;
  }
}

// workspace of the block loop, shared layout for both towers (rows M = items * tokens)
// partials per row of width W written by the residual-GEMM epilogue (one per epilogue warp column slice)
constexpr int kStatCols = 256 / GemmPairCfg<256>::COL_GROUPS;
inline int kStatPartials(size_t W) { return static_cast<int>(W) / kStatCols; }

struct CoreWs { size_t x, x16, xlo, qkv, att, hid, abuf, stats, end; };
// ab = bytes per activation-operand element: 2, or 4 in fp32-accuracy mode (x16 then holds the fp32 LayerNorm output)
CoreWs plan_core_ws(ArenaPlan& p, size_t M, size_t W, size_t F, size_t ab) {
  CoreWs w;
  w.x = p.take<float>(M * W);            // fp32 staging of the stream before ln_pre / the text embedding; fp32 mode: the stream
  w.x16 = p.take<uint8_t>(M * W * ab);   // 16-bit modes: the hi plane of the residual stream = A operand of QKV / c_fc / adapter GEMMs
  w.xlo = p.take<uint8_t>(M * W * 2);    // 16-bit modes: the lo plane (x = hi + lo, ptx::pack16_hilo)
  w.qkv = p.take<uint8_t>(M * 3 * W * ab);
  w.att = p.take<uint8_t>(M * W * ab);
  w.hid = p.take<uint8_t>(M * F * ab);
  w.abuf = p.take<float>(M * W);
  w.stats = p.take<float>(M * kStatPartials(W) * 2);   // LayerNorm partials (sum, sumsq) per epilogue-warp column slice
  w.end = p.off;
  return w;
}

}  // namespace

struct AaclipHandle {
  AaclipCfg cfg;
  bool bf16 = false;      // 16-bit operand dtype: fp16 (default) or bf16
  bool f32 = false;       // fp32-accuracy mode (AACLIP_DT_F32): SIMT kernels, fp32 weights behind the uint16_t* pointers
  int G, P, N, KP;        // grid side, patches, tokens, padded patch K
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  // fp32 params
  float *cls, *pos, *lnpre_g, *lnpre_b, *lnpost_g, *lnpost_b;
  uint16_t* w_conv;  // [W, KP]
  std::vector<BlockW> blk;
  bool folded = false;
  std::vector<uint16_t*> w_adapter;  // [W,W] each
  std::vector<uint16_t*> w_seg;      // [E,W] each
  uint16_t* w_det;                   // [E,W]
  std::vector<uint16_t*> w_query;    // [query_dim,W] each (IQM query_adapters; optional)
  std::vector<bool> have_query;
  // ln_post folded into the token-tap projections (16-bit modes): W' = W diag(gamma_post) in w_seg / w_det / w_query,
  // plus colsum(W') and bias' = W beta_post per projection. Index: 0..n_levels-1 seg, n_levels det, then query.
  struct ProjFold { float* stage = nullptr; uint16_t* w16 = nullptr; float* colsum = nullptr; float* bias_f = nullptr; int rows = 0; };
  std::vector<ProjFold> proj;
  float* zero_bias = nullptr;        // max(E, query_dim) zeros: the projections have no bias of their own
  std::vector<std::string> missing;
  int last_launches = 0;
  bool seg_normalize = true;   // aaclip_seg_normalize: false = seg_out stays un-normalised (the scoring call normalises)
  unsigned long long* sat_counter = nullptr;   // device counter of aaclip_saturation_check (cudaMalloc'ed at create)
  bool sat_on = false;
  // optional per-kernel-class device timing (aaclip_profile_*)
  struct ProfRec { int cls; double flops; cudaEvent_t a, b; };
  bool prof_on = false;
  std::vector<ProfRec> prof;
};

// Records a CUDA-event pair around one launch on the launch stream when profiling is enabled (h may be null).
struct ProfScope {
  AaclipHandle* h;
  cudaStream_t st;
  size_t idx = 0;
  bool on;
  ProfScope(AaclipHandle* h_, int cls, double flops, cudaStream_t st_) : h(h_), st(st_), on(h_ && h_->prof_on) {
    if (!on) return;
    AaclipHandle::ProfRec r;
    r.cls = cls;
    r.flops = flops;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    h->prof.push_back(r);
    idx = h->prof.size() - 1;
  }
  ~ProfScope() {
    if (on) cudaEventRecord(h->prof[idx].b, st);
  }
};

namespace {

// One residual attention block (+ optional residual adapter), model/transformer.py:239-258 + model/adapter.py:163-170.
//   x, x16, stats hold the residual stream on entry (fp32, raw 16-bit copy, LayerNorm partials) and on exit.
// 5 launches (7 with an adapter): LayerNorms live in the GEMM epilogues.
// aaclip_saturation_check: after every kernel that stores 16-bit activations, count the stored elements that sit at the
// format's ceiling (elementwise.cuh: count_saturated16_kernel). Off by default: zero cost on the hot path.
struct SatCheck {
  unsigned long long* counter = nullptr;   // device counter, nullptr = disabled
  bool bf16 = false;
  int launches = 0;
  void scan(const void* p, size_t n, cudaStream_t st) {
    if (counter == nullptr || n == 0) return;
    const size_t n8 = n / 8;
    const unsigned grid = static_cast<unsigned>(n8 / 256 + 1 < 148 * 8 ? n8 / 256 + 1 : 148 * 8);
    if (bf16) launch_pdl(count_saturated16_kernel<true>, dim3(grid), dim3(256), 0, st, static_cast<const uint16_t*>(p), n8, counter);
    else launch_pdl(count_saturated16_kernel<false>, dim3(grid), dim3(256), 0, st, static_cast<const uint16_t*>(p), n8, counter);
    ++launches;
  }
};

struct BlockRun {
  AaclipHandle* prof;     // profiling sink (visual handle) or nullptr
  SatCheck sat;
  bool bf16;
  int items, tokens, W, F, causal, act;
  // the residual stream x = x16 + xlo: two 16-bit planes (ptx::pack16_hilo); x16, the rounded value, is the A operand
  // of the QKV / c_fc / adapter / tap GEMMs. stats: its LayerNorm partials.
  uint16_t* x16; uint16_t* xlo; uint16_t* qkv; uint16_t* att; uint16_t* hid; float* abuf; float* stats;
  cudaStream_t st;
  int launches = 0;
  // Serpentine row order (GemmEpi::descending): every GEMM / attention launch walks the row blocks in the direction
  // opposite to the kernel that wrote its input, so it starts with the rows that are still in L2. `last_desc` = the
  // direction of the last kernel that wrote an activation of the stream; element-wise kernels run ascending.
  // AACLIP_SERPENTINE=0 in the environment keeps every launch ascending (A/B runs).
  bool last_desc = false;
  static bool serpentine() {
    static const bool on = [] {
      const char* e = getenv("AACLIP_SERPENTINE");
      return !(e != nullptr && e[0] == '0');
    }();
    return on;
  }
  int next_dir() {
    if (!serpentine()) return 0;
    last_desc = !last_desc;
    return last_desc ? 1 : 0;
  }

  int gemm(int epi, const void* A, const void* Wt, int m, int n, int k, const GemmEpi& e, int cls = AACLIP_PROF_GEMM) {
    ProfScope ps(prof, cls, 2.0 * m * n * k, st);
    ++launches;
    GemmEpi ed = e;
    ed.descending = next_dir();
    return launch_gemm(epi, bf16, A, Wt, m, n, k, ed, st);
  }
  // a GEMM that only READS the stream (token-tap projections): opposite to the last writer, and the next kernel still
  // sees that writer's direction
  int gemm_tap(int epi, const void* A, const void* Wt, int m, int n, int k, const GemmEpi& e) {
    ProfScope ps(prof, AACLIP_PROF_GEMM, 2.0 * m * n * k, st);
    ++launches;
    GemmEpi ed = e;
    ed.descending = (serpentine() && !last_desc) ? 1 : 0;
    return launch_gemm(epi, bf16, A, Wt, m, n, k, ed, st);
  }

  int run(const BlockW& b, const uint16_t* w_adapter, float adapt_weight) {
    const int M = items * tokens, np = kStatPartials(W);
    int rc;
    { GemmEpi e; e.bias = b.bf_qkv; e.colsum = b.cs_qkv; e.stats = stats; e.n_partials = np; e.inv_width = 1.0f / W;
      e.out_16 = qkv; e.ldo = 3 * W;
      if ((rc = gemm(EPI_LN_BIAS_BF16, x16, b.w_qkv, M, 3 * W, W, e, AACLIP_PROF_GEMM_QKV))) return rc; }
    sat.scan(qkv, static_cast<size_t>(M) * 3 * W, st);
    {
      ProfScope ps(prof, AACLIP_PROF_ATTENTION, 4.0 * items * (W / 64) * (double)tokens * tokens * 64, st);
      ++launches;
      if ((rc = launch_attention(bf16, qkv, att, items, tokens, W, causal, st, next_dir()))) return rc;
    }
    sat.scan(att, static_cast<size_t>(M) * W, st);
    { GemmEpi e; e.bias = b.b_out; e.out_16 = x16; e.out_lo = xlo; e.stats_out = stats; e.ldo = W;
      if ((rc = gemm(EPI_BIAS_RESID_HL, att, b.w_out, M, W, W, e, AACLIP_PROF_GEMM_OUT))) return rc; }
    sat.scan(x16, static_cast<size_t>(M) * W, st);
    { GemmEpi e; e.bias = b.bf_fc; e.colsum = b.cs_fc; e.stats = stats; e.n_partials = np; e.inv_width = 1.0f / W;
      e.out_16 = hid; e.ldo = F;
      if ((rc = gemm(act == AACLIP_ACT_QUICK_GELU ? EPI_LN_BIAS_QGELU_BF16 : EPI_LN_BIAS_GELU_BF16, x16, b.w_fc, M, F, W, e, AACLIP_PROF_GEMM_FC))) return rc; }
    sat.scan(hid, static_cast<size_t>(M) * F, st);
    { GemmEpi e; e.bias = b.b_proj; e.out_16 = x16; e.out_lo = xlo; e.stats_out = stats; e.ldo = W;
      if ((rc = gemm(EPI_BIAS_RESID_HL, hid, b.w_proj, M, W, F, e, AACLIP_PROF_GEMM_PROJ))) return rc; }
    sat.scan(x16, static_cast<size_t>(M) * W, st);
    if (w_adapter != nullptr) return adapter(w_adapter, M, adapt_weight);
    return AACLIP_OK;
  }

  // residual adapter + blend, model/adapter.py:163-170: a = LeakyReLU(x W^T); a <- a |x| / |a|; x <- w a + (1-w) x.
  // a leaves the GEMM in the 16-bit operand dtype together with the per-row partials of sum a^2 (EPI_LEAKY_16_SQ); both
  // live in the fp32-sized `abuf` slot of the workspace.
  int adapter(const uint16_t* w_adapter, int M, float adapt_weight) {
    const int np = kStatPartials(W);
    int rc;
    uint16_t* a16 = reinterpret_cast<uint16_t*>(abuf);
    float* a_stats = reinterpret_cast<float*>(a16 + static_cast<size_t>(M) * W);
    { GemmEpi e; e.out_16 = a16; e.stats_out = a_stats; e.ldo = W;
      if ((rc = gemm(EPI_LEAKY_16_SQ, x16, w_adapter, M, W, W, e))) return rc; }
    sat.scan(a16, static_cast<size_t>(M) * W, st);
    ProfScope ps(prof, AACLIP_PROF_OTHER, 0.0, st);
    const int grid = (M + 7) / 8;
    const float w = adapt_weight, w1 = 1.0f - adapt_weight;
    if (W == 1024) {
      if (bf16) launch_pdl(adapter_blend_hl_kernel<1024, true>, dim3(grid), dim3(256), 0, st, x16, xlo, static_cast<const uint16_t*>(a16), static_cast<const float*>(a_stats), M, w, w1, stats, np);
      else launch_pdl(adapter_blend_hl_kernel<1024, false>, dim3(grid), dim3(256), 0, st, x16, xlo, static_cast<const uint16_t*>(a16), static_cast<const float*>(a_stats), M, w, w1, stats, np);
    } else {
      if (bf16) launch_pdl(adapter_blend_hl_kernel<768, true>, dim3(grid), dim3(256), 0, st, x16, xlo, static_cast<const uint16_t*>(a16), static_cast<const float*>(a_stats), M, w, w1, stats, np);
      else launch_pdl(adapter_blend_hl_kernel<768, false>, dim3(grid), dim3(256), 0, st, x16, xlo, static_cast<const uint16_t*>(a16), static_cast<const float*>(a_stats), M, w, w1, stats, np);
    }
    LAUNCH_CHECK("adapter_blend_hl_kernel");
    last_desc = false;                     // an element-wise kernel: rows ascending
    ++launches;
    sat.scan(x16, static_cast<size_t>(M) * W, st);
    return AACLIP_OK;
  }
};

// The same block in fp32-accuracy mode: un-folded LayerNorms, SIMT fp32 GEMMs and attention (fp32_path.cuh).
// The uint16_t* weight pointers of BlockW address fp32 data in this mode.
struct BlockRunF32 {
  AaclipHandle* prof;
  int items, tokens, W, F, causal, act;
  float* x; float* xn; float* qkv; float* att; float* hid; float* abuf;
  cudaStream_t st;
  int launches = 0;

  int gemm(int epi, const float* A, const void* Wt, int m, int n, int k, const GemmEpi& e, int cls = AACLIP_PROF_GEMM) {
    ProfScope ps(prof, cls, 2.0 * m * n * k, st);
    ++launches;
    return launch_sgemm(epi, A, reinterpret_cast<const float*>(Wt), m, n, k, e, st);
  }
  int lnorm(const float* in, const float* g, const float* b, float* o, int rows, int skipP) {
    ProfScope ps(prof, AACLIP_PROF_LAYERNORM, 0.0, st);
    ++launches;
    return launch_layernorm(false, in, g, b, nullptr, o, rows, W, skipP, true, st);
  }
  int run(const BlockW& b, const uint16_t* w_adapter, float adapt_weight) {
    const int M = items * tokens;
    int rc;
    if ((rc = lnorm(x, b.ln1_g, b.ln1_b, xn, M, 0))) return rc;
    { GemmEpi e; e.bias = b.b_qkv; e.out_f32 = qkv; e.ldo = 3 * W;
      if ((rc = gemm(EPI_BIAS_F32, xn, b.w_qkv, M, 3 * W, W, e, AACLIP_PROF_GEMM_QKV))) return rc; }
    {
      ProfScope ps(prof, AACLIP_PROF_ATTENTION, 4.0 * items * (W / 64) * (double)tokens * tokens * 64, st);
      ++launches;
      if ((rc = launch_attention_f32(qkv, att, items, tokens, W, causal, st))) return rc;
    }
    { GemmEpi e; e.bias = b.b_out; e.out_f32 = x; e.ldo = W;
      if ((rc = gemm(EPI_BIAS_RESID_F32, att, b.w_out, M, W, W, e, AACLIP_PROF_GEMM_OUT))) return rc; }
    if ((rc = lnorm(x, b.ln2_g, b.ln2_b, xn, M, 0))) return rc;
    { GemmEpi e; e.bias = b.b_fc; e.out_f32 = hid; e.ldo = F;
      if ((rc = gemm(act == AACLIP_ACT_QUICK_GELU ? EPI_BIAS_QGELU_BF16 : EPI_BIAS_GELU_BF16, xn, b.w_fc, M, F, W, e, AACLIP_PROF_GEMM_FC))) return rc; }
    { GemmEpi e; e.bias = b.b_proj; e.out_f32 = x; e.ldo = W;
      if ((rc = gemm(EPI_BIAS_RESID_F32, hid, b.w_proj, M, W, F, e, AACLIP_PROF_GEMM_PROJ))) return rc; }
    if (w_adapter != nullptr) {
      { GemmEpi e; e.out_f32 = abuf; e.ldo = W;
        if ((rc = gemm(EPI_LEAKY_F32, x, w_adapter, M, W, W, e))) return rc; }
      ProfScope ps(prof, AACLIP_PROF_OTHER, 0.0, st);
      const int grid = (M + 7) / 8;
      const float w = adapt_weight, w1 = 1.0f - adapt_weight;
      if (W == 1024) adapter_blend_kernel<1024, false><<<grid, 256, 0, st>>>(x, abuf, M, w, w1);
      else adapter_blend_kernel<768, false><<<grid, 256, 0, st>>>(x, abuf, M, w, w1);
      LAUNCH_CHECK("adapter_blend_kernel");
      ++launches;
    }
    return AACLIP_OK;
  }
};

}  // namespace

#ifdef AACLIP_ATTN_DBG_TIMELINE
extern "C" int aaclip_debug_attn_timeline(long long* dst, long long* dst_mma) {
  if (cudaMemcpyFromSymbol(dst, aaclip::g_attn_timeline, sizeof(long long) * 2 * 16 * 8) != cudaSuccess) return -1;
  return cudaMemcpyFromSymbol(dst_mma, aaclip::g_attn_timeline_mma, sizeof(long long) * 2 * 16 * 8) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" {

const char* aaclip_version(void) { return "aaclip_b200 0.2 sm_100a"; }
const char* aaclip_last_error(void) { return g_err; }

int aaclip_create(const AaclipCfg* cfg, AaclipHandle** out) {
  if (cfg == nullptr || out == nullptr) return fail(AACLIP_EINVAL, "aaclip_create: null argument");
  *out = nullptr;
  if (int rc = ensure_device()) return rc;
  const AaclipCfg& c = *cfg;
  if (c.patch <= 0 || c.img_size < c.patch) return fail(AACLIP_EINVAL, "bad img_size/patch %d/%d", c.img_size, c.patch);
  if (c.width != 1024 && c.width != 768) return fail(AACLIP_EINVAL, "width %d unsupported (768 or 1024)", c.width);
  if (c.heads * 64 != c.width) return fail(AACLIP_EINVAL, "head dim must be 64 (width %d, heads %d)", c.width, c.heads);
  if (c.mlp_dim % 256 != 0 || c.embed_dim % 256 != 0 || (c.embed_dim != 768 && c.embed_dim != 1024))
    return fail(AACLIP_EINVAL, "mlp_dim/embed_dim unsupported (%d/%d)", c.mlp_dim, c.embed_dim);
  if (c.layers <= 0 || c.adapt_until < 0 || c.adapt_until > c.layers) return fail(AACLIP_EINVAL, "bad layers/adapt_until");
  if (c.n_levels <= 0 || c.n_levels > AACLIP_MAX_LEVELS) return fail(AACLIP_EINVAL, "n_levels must be 1..4");
  for (int l = 0; l < c.n_levels; ++l) {
    if (c.levels[l] < 1 || c.levels[l] > c.layers || (l > 0 && c.levels[l] <= c.levels[l - 1]))
      return fail(AACLIP_EINVAL, "levels must be ascending in 1..layers");
  }
  if (c.max_batch <= 0) return fail(AACLIP_EINVAL, "max_batch must be > 0");
  if (c.query_dim < 0 || c.query_dim % 256 != 0) return fail(AACLIP_EINVAL, "query_dim must be 0 or a multiple of 256 (got %d)", c.query_dim);
  if (c.operand_dtype != AACLIP_DT_F16 && c.operand_dtype != AACLIP_DT_BF16 && c.operand_dtype != AACLIP_DT_F32)
    return fail(AACLIP_EINVAL, "operand_dtype must be AACLIP_DT_F16, AACLIP_DT_BF16 or AACLIP_DT_F32");

  AaclipHandle* h = new AaclipHandle();
  h->cfg = c;
  h->bf16 = (c.operand_dtype == AACLIP_DT_BF16);
  h->f32 = (c.operand_dtype == AACLIP_DT_F32);
  const size_t wb = h->f32 ? 4 : 2;       // bytes per GEMM-weight element
  h->G = c.img_size / c.patch;
  h->P = h->G * h->G;
  h->N = h->P + 1;
  const int k_raw = 3 * c.patch * c.patch;
  h->KP = (k_raw + 63) / 64 * 64;
  const size_t W = c.width, F = c.mlp_dim, E = c.embed_dim;

  ArenaPlan plan;
  const size_t o_cls = plan.take<float>(W), o_pos = plan.take<float>((size_t)h->N * W);
  const size_t o_lpg = plan.take<float>(W), o_lpb = plan.take<float>(W), o_log = plan.take<float>(W), o_lob = plan.take<float>(W);
  const size_t o_conv = plan.take<uint8_t>(W * h->KP * wb);
  std::vector<BlockOff> bo(c.layers);
  for (int i = 0; i < c.layers; ++i) bo[i] = plan_block(plan, W, F, wb);
  std::vector<size_t> o_ad(c.adapt_until), o_seg(c.n_levels);
  for (int i = 0; i < c.adapt_until; ++i) o_ad[i] = plan.take<uint8_t>(W * W * wb);
  for (int l = 0; l < c.n_levels; ++l) o_seg[l] = plan.take<uint8_t>(E * W * wb);
  const size_t o_det = plan.take<uint8_t>(E * W * wb);
  std::vector<size_t> o_qry(c.query_dim > 0 ? c.n_levels : 0);
  for (size_t l = 0; l < o_qry.size(); ++l) o_qry[l] = plan.take<uint8_t>((size_t)c.query_dim * W * wb);
  const size_t o_sat = plan.take<unsigned long long>(1);
  const size_t n_proj = (size_t)c.n_levels + 1 + o_qry.size();
  const size_t proj_rows_max = E > (size_t)c.query_dim ? E : (size_t)c.query_dim;
  std::vector<size_t> o_pcs(n_proj), o_pbf(n_proj);
  for (size_t i = 0; i < n_proj; ++i) { o_pcs[i] = plan.take<float>(proj_rows_max); o_pbf[i] = plan.take<float>(proj_rows_max); }
  const size_t o_zero = plan.take<float>(proj_rows_max);

  h->arena_bytes = plan.off;
  cudaError_t e = cudaMalloc(&h->arena, h->arena_bytes);
  if (e != cudaSuccess) {
    delete h;
    return fail(AACLIP_ECUDA, "cudaMalloc(%zu) for the weight arena failed: %s", plan.off, cudaGetErrorString(e));
  }
  e = cudaMemset(h->arena, 0, h->arena_bytes);
  if (e != cudaSuccess) {
    cudaFree(h->arena);
    delete h;
    return fail(AACLIP_ECUDA, "cudaMemset of the weight arena failed: %s", cudaGetErrorString(e));
  }
  uint8_t* a = h->arena;
  h->cls = (float*)(a + o_cls); h->pos = (float*)(a + o_pos);
  h->lnpre_g = (float*)(a + o_lpg); h->lnpre_b = (float*)(a + o_lpb);
  h->lnpost_g = (float*)(a + o_log); h->lnpost_b = (float*)(a + o_lob);
  h->w_conv = (uint16_t*)(a + o_conv);
  h->blk.resize(c.layers);
  for (int i = 0; i < c.layers; ++i) bind_block(h->blk[i], a, bo[i]);
  for (int i = 0; i < c.adapt_until; ++i) h->w_adapter.push_back((uint16_t*)(a + o_ad[i]));
  for (int l = 0; l < c.n_levels; ++l) h->w_seg.push_back((uint16_t*)(a + o_seg[l]));
  h->w_det = (uint16_t*)(a + o_det);
  for (size_t l = 0; l < o_qry.size(); ++l) h->w_query.push_back((uint16_t*)(a + o_qry[l]));
  h->have_query.assign(o_qry.size(), false);
  h->sat_counter = (unsigned long long*)(a + o_sat);      // zeroed with the arena
  h->proj.resize(n_proj);
  for (size_t i = 0; i < n_proj; ++i) {
    AaclipHandle::ProjFold& pf = h->proj[i];
    pf.colsum = (float*)(a + o_pcs[i]); pf.bias_f = (float*)(a + o_pbf[i]);
    if (i < (size_t)c.n_levels) { pf.w16 = h->w_seg[i]; pf.rows = (int)E; }
    else if (i == (size_t)c.n_levels) { pf.w16 = h->w_det; pf.rows = (int)E; }
    else { pf.w16 = h->w_query[i - c.n_levels - 1]; pf.rows = c.query_dim; }
  }
  h->zero_bias = (float*)(a + o_zero);

  // the set of tensors that must arrive through aaclip_set_weight before forward may run
  auto need = [&](const std::string& k) { h->missing.push_back(k); };
  need("conv1.weight"); need("class_embedding"); need("positional_embedding");
  need("ln_pre.weight"); need("ln_pre.bias"); need("ln_post.weight"); need("ln_post.bias");
  for (int i = 0; i < c.layers; ++i)
    for (const char* t : kBlockTensors) need("blk." + std::to_string(i) + "." + t);
  for (int i = 0; i < c.adapt_until; ++i) need("adapter." + std::to_string(i));
  for (int l = 0; l < c.n_levels; ++l) need("seg." + std::to_string(l));
  need("det");
  *out = h;
  return AACLIP_OK;
}

void aaclip_destroy(AaclipHandle* h) {
  if (h == nullptr) return;
  free_block_staging(h->blk);
  for (auto& pf : h->proj) if (pf.stage) cudaFree(pf.stage);
  if (h->arena) cudaFree(h->arena);
  for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  delete h;
}

int aaclip_missing_weights(AaclipHandle* h) { return h ? static_cast<int>(h->missing.size()) : -1; }

int aaclip_set_weight(AaclipHandle* h, const char* name, const float* data, const int64_t* shape, int32_t ndim,
                      void* stream) {
  if (!h || !name || !data || !shape || ndim <= 0) return fail(AACLIP_EINVAL, "aaclip_set_weight: null/empty argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const AaclipCfg& c = h->cfg;
  const int64_t W = c.width, F = c.mlp_dim, E = c.embed_dim;
  const int64_t n = numel(shape, ndim);
  std::string k(name);
  for (const char* pre : {"image_encoder.", "clipmodel.visual.", "visual."}) {
    if (k.rfind(pre, 0) == 0) { k = k.substr(strlen(pre)); break; }
  }
  auto copy_f32 = [&](float* dst, int64_t expect, const std::string& key) -> int {
    if (n != expect) return fail(AACLIP_EINVAL, "%s: expected %lld elements, got %lld", name, (long long)expect, (long long)n);
    CUDA_TRY(cudaMemcpyAsync(dst, data, expect * sizeof(float), cudaMemcpyDeviceToDevice, st));
    erase_key(h->missing, key);
    return AACLIP_OK;
  };
  auto pack = [&](uint16_t* dst, int64_t R, int64_t C, int64_t Cp, const std::string& key) -> int {
    if (n != R * C) return fail(AACLIP_EINVAL, "%s: expected %lld x %lld elements, got %lld", name, (long long)R, (long long)C, (long long)n);
    if (int rc = h->f32 ? copy_pad_f32(data, dst, R, C, Cp, st) : launch_pack(h->bf16, data, dst, R, C, Cp, st)) return rc;
    erase_key(h->missing, key);
    return AACLIP_OK;
  };

  if (k == "conv1.weight") return pack(h->w_conv, W, 3 * c.patch * c.patch, h->KP, k);
  if (k == "class_embedding") return copy_f32(h->cls, W, k);
  if (k == "positional_embedding") return copy_f32(h->pos, (int64_t)h->N * W, k);
  if (k == "ln_pre.weight") return copy_f32(h->lnpre_g, W, k);
  if (k == "ln_pre.bias") return copy_f32(h->lnpre_b, W, k);
  if ((k == "ln_post.weight" || k == "ln_post.bias") && h->folded)
    return fail(AACLIP_ESTATE, "%s: ln_post has already been folded into the token-tap projections (first forward ran); "
                               "create a new handle to load different backbone weights", name);
  if (k == "ln_post.weight") return copy_f32(h->lnpost_g, W, k);
  if (k == "ln_post.bias") return copy_f32(h->lnpost_b, W, k);
  // seg / det / query projections: ln_post is folded into them (16-bit modes). Before the fold the fp32 weight is
  // staged; a projection that arrives after it (trained heads re-fed, late query adapters) is folded on the spot.
  auto proj = [&](size_t pi, uint16_t* dst, int64_t R, const std::string& key) -> int {
    if (h->f32) return pack(dst, R, W, W, key);
    if (n != R * W) return fail(AACLIP_EINVAL, "%s: expected %lld x %lld elements, got %lld", name, (long long)R, (long long)W, (long long)n);
    AaclipHandle::ProjFold& pf = h->proj[pi];
    if (h->folded) {
      if (h->bf16) ln_fold_kernel<true><<<(R + 7) / 8, 256, 0, st>>>(data, h->zero_bias, h->lnpost_g, h->lnpost_b, pf.w16, pf.colsum, pf.bias_f, (int)R, (int)W);
      else ln_fold_kernel<false><<<(R + 7) / 8, 256, 0, st>>>(data, h->zero_bias, h->lnpost_g, h->lnpost_b, pf.w16, pf.colsum, pf.bias_f, (int)R, (int)W);
      LAUNCH_CHECK("ln_fold_kernel");
    } else {
      if (pf.stage == nullptr) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&pf.stage), R * W * sizeof(float)));
      CUDA_TRY(cudaMemcpyAsync(pf.stage, data, R * W * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (!key.empty()) erase_key(h->missing, key);
    return AACLIP_OK;
  };

  int idx = -1;
  char tail[96];
  if (sscanf(k.c_str(), "transformer.resblocks.%d.%95s", &idx, tail) == 2) {
    if (idx < 0 || idx >= c.layers) return fail(AACLIP_EINVAL, "%s: block index out of range", name);
    return set_block_tensor(h->blk[idx], tail, "blk." + std::to_string(idx) + "." + tail, h->missing, h->folded, h->bf16,
                            h->f32, W, F, name, data, n, st);
  }
  if (sscanf(k.c_str(), "image_adapter.layer_adapters.%d.%95s", &idx, tail) == 2) {
    if (strcmp(tail, "fc.0.weight") != 0) return fail(AACLIP_EINVAL, "%s: unknown tensor", name);
    if (idx < 0 || idx >= c.adapt_until) return fail(AACLIP_EINVAL, "%s: adapter index out of range", name);
    return pack(h->w_adapter[idx], W, W, W, "adapter." + std::to_string(idx));
  }
  if (sscanf(k.c_str(), "image_adapter.seg_proj.%d.fc.%95s", &idx, tail) == 2) {
    if (idx < 0 || idx >= c.n_levels) return fail(AACLIP_EINVAL, "%s: seg_proj index out of range", name);
    if (strcmp(tail, "weight") != 0 && strcmp(tail, "0.weight") != 0) return fail(AACLIP_EINVAL, "%s: unknown tensor", name);
    return proj((size_t)idx, h->w_seg[idx], E, "seg." + std::to_string(idx));
  }
  if (k == "image_adapter.det_proj.fc.weight" || k == "image_adapter.det_proj.fc.0.weight") return proj((size_t)c.n_levels, h->w_det, E, "det");
  if (sscanf(k.c_str(), "query_adapters.%d.fc.%95s", &idx, tail) == 2) {
    if (c.query_dim <= 0) return fail(AACLIP_EINVAL, "%s: the handle was created with query_dim = 0", name);
    if (idx < 0 || idx >= c.n_levels) return fail(AACLIP_EINVAL, "%s: query_adapters index out of range", name);
    if (strcmp(tail, "weight") != 0 && strcmp(tail, "0.weight") != 0) return fail(AACLIP_EINVAL, "%s: unknown tensor", name);
    if (int rc = proj((size_t)c.n_levels + 1 + idx, h->w_query[idx], c.query_dim, "")) return rc;
    h->have_query[idx] = true;
    return AACLIP_OK;
  }
  return fail(AACLIP_EINVAL, "%s: unknown tensor name", name);
}

// workspace carve-up shared by aaclip_workspace_bytes and aaclip_forward
struct WsPlan {
  CoreWs core;
  size_t patchA, tok, det, total;
};
static WsPlan plan_ws(const AaclipHandle* h, int B) {
  const size_t W = h->cfg.width, F = h->cfg.mlp_dim, E = h->cfg.embed_dim;
  const size_t M = (size_t)B * h->N, MP = (size_t)B * h->P;
  ArenaPlan p;
  WsPlan w;
  const size_t ab = h->f32 ? 4 : 2;
  w.core = plan_core_ws(p, M, W, F, ab);
  w.patchA = p.take<uint8_t>(MP * h->KP * ab);
  w.tok = p.take<uint8_t>(MP * W * ab);
  w.det = p.take<float>(MP * E);
  w.total = p.off;
  return w;
}

}  // extern "C"

// aaclip_forward in fp32-accuracy mode: same data layout and launch order as the tensor-core path, un-folded
// LayerNorms, SIMT fp32 GEMMs / attention, fp32 patch operand. Verification mode (BASELINE config 4).
static int forward_f32(AaclipHandle* h, const float* images, int B, uint8_t* wsb, const WsPlan& ws, float* const* seg_out,
                       float* det_out, float* const* tokens_out, float* cls_out, float* const* query_out, cudaStream_t st) {
  const AaclipCfg& c = h->cfg;
  const int W = c.width, F = c.mlp_dim, E = c.embed_dim, N = h->N, P = h->P;
  const int M = B * N, MP = B * P;
  int rc;
  float* patchA = (float*)(wsb + ws.patchA);
  float* tok = (float*)(wsb + ws.tok);
  float* detbuf = (float*)(wsb + ws.det);
  BlockRunF32 run;
  run.prof = h; run.items = B; run.tokens = N; run.W = W; run.F = F; run.causal = 0; run.act = c.act;
  run.x = (float*)(wsb + ws.core.x); run.xn = (float*)(wsb + ws.core.x16); run.qkv = (float*)(wsb + ws.core.qkv);
  run.att = (float*)(wsb + ws.core.att); run.hid = (float*)(wsb + ws.core.hid); run.abuf = (float*)(wsb + ws.core.abuf);
  run.st = st;
  float* x = run.x;
  int& launches = run.launches;
  auto l2norm = [&](float* t, int rows) -> int {
    ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
    ++launches;
    if (E == 768) l2norm_rows_kernel<768><<<(rows + 7) / 8, 256, 0, st>>>(t, rows);
    else l2norm_rows_kernel<1024><<<(rows + 7) / 8, 256, 0, st>>>(t, rows);
    LAUNCH_CHECK("l2norm_rows_kernel");
    return AACLIP_OK;
  };
  {
    ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
    patch_gather_f32_kernel<<<dim3(h->G, B), 256, 0, st>>>(images, patchA, c.img_size, h->G, c.patch, h->KP);
    LAUNCH_CHECK("patch_gather_f32_kernel"); ++launches;
  }
  { GemmEpi e; e.out_f32 = x; e.ldo = W; e.pos = h->pos; e.patches = P;
    if ((rc = run.gemm(EPI_PATCH_F32, patchA, h->w_conv, MP, W, h->KP, e))) return rc; }
  {
    ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
    cls_row_kernel<<<B, 256, 0, st>>>(x, h->cls, h->pos, N, W);
    LAUNCH_CHECK("cls_row_kernel"); ++launches;
  }
  if ((rc = run.lnorm(x, h->lnpre_g, h->lnpre_b, x, M, 0))) return rc;
  int level_idx = 0;
  for (int i = 0; i < c.layers; ++i) {
    if ((rc = run.run(h->blk[i], i < c.adapt_until ? h->w_adapter[i] : nullptr, c.adapt_weight))) return rc;
    if (level_idx < c.n_levels && c.levels[level_idx] == i + 1) {
      float* tok_out = (tokens_out != nullptr && tokens_out[level_idx] != nullptr) ? tokens_out[level_idx] : tok;
      if ((rc = run.lnorm(x, h->lnpost_g, h->lnpost_b, tok_out, MP, P))) return rc;
      const int proj_epi = c.proj_relu ? EPI_LEAKY_F32 : EPI_F32;
      { GemmEpi e; e.out_f32 = seg_out[level_idx]; e.ldo = E;
        if ((rc = run.gemm(proj_epi, tok_out, h->w_seg[level_idx], MP, E, W, e))) return rc; }
      if (query_out != nullptr && query_out[level_idx] != nullptr) {
        GemmEpi e; e.out_f32 = query_out[level_idx]; e.ldo = c.query_dim;
        if ((rc = run.gemm(proj_epi, tok_out, h->w_query[level_idx], MP, c.query_dim, W, e))) return rc;
      }
      if ((rc = l2norm(seg_out[level_idx], MP))) return rc;
      if (level_idx == c.n_levels - 1) {
        { GemmEpi e; e.out_f32 = detbuf; e.ldo = E;
          if ((rc = run.gemm(proj_epi, tok_out, h->w_det, MP, E, W, e))) return rc; }
        if ((rc = l2norm(detbuf, MP))) return rc;
        ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
        mean_over_patches_kernel<<<dim3((E + MEAN_COLS - 1) / MEAN_COLS, B), dim3(MEAN_COLS, MEAN_SEGS), 0, st>>>(detbuf, det_out, P, E);
        LAUNCH_CHECK("mean_over_patches_kernel");
        ++launches;
      }
      ++level_idx;
    }
  }
  if (cls_out != nullptr) {
    ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
    gather_cls_kernel<<<B, 256, 0, st>>>(x, cls_out, N, W);
    LAUNCH_CHECK("gather_cls_kernel"); ++launches;
  }
  h->last_launches = launches;
  return AACLIP_OK;
}

extern "C" {

size_t aaclip_workspace_bytes(AaclipHandle* h, int32_t B) {
  if (!h || B <= 0) return 0;
  return plan_ws(h, B).total;
}

int aaclip_forward(AaclipHandle* h, const float* images, int32_t B, void* workspace, size_t workspace_bytes,
                   float* const* seg_out, float* det_out, float* const* tokens_out, float* cls_out,
                   float* const* query_out, void* stream) {
  if (!h || !images || !workspace || !seg_out || !det_out) return fail(AACLIP_EINVAL, "aaclip_forward: null argument");
  if (B <= 0 || B > h->cfg.max_batch) return fail(AACLIP_EINVAL, "aaclip_forward: batch %d outside 1..%d", B, h->cfg.max_batch);
  if (!h->missing.empty())
    return fail(AACLIP_ESTATE, "aaclip_forward: %zu weight tensors not set (first: %s)", h->missing.size(), h->missing[0].c_str());
  const WsPlan ws = plan_ws(h, B);
  if (workspace_bytes < ws.total) return fail(AACLIP_EINVAL, "aaclip_forward: workspace %zu < required %zu", workspace_bytes, ws.total);
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(AACLIP_EINVAL, "aaclip_forward: workspace must be 256-byte aligned");
  for (int l = 0; l < h->cfg.n_levels; ++l)
    if (!seg_out[l]) return fail(AACLIP_EINVAL, "aaclip_forward: seg_out[%d] is null", l);
  if (query_out != nullptr) {
    for (int l = 0; l < h->cfg.n_levels; ++l) {
      if (query_out[l] == nullptr) continue;
      if (h->cfg.query_dim <= 0) return fail(AACLIP_EINVAL, "aaclip_forward: query_out given but the handle has query_dim = 0");
      if (!h->have_query[l]) return fail(AACLIP_ESTATE, "aaclip_forward: query_out[%d] requested but query_adapters.%d was never set", l, l);
    }
  }

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const AaclipCfg& c = h->cfg;
  const int W = c.width, F = c.mlp_dim, E = c.embed_dim, N = h->N, P = h->P;
  const int M = B * N, MP = B * P;
  int rc;
  if (h->f32) return forward_f32(h, images, B, static_cast<uint8_t*>(workspace), ws, seg_out, det_out, tokens_out, cls_out, query_out, st);
  // LayerNorm folding is load-time work (kernels + a synchronisation + cudaFree of the staging copies): callers that
  // want a capture-safe, non-blocking first forward call aaclip_finalize_weights() after the last aaclip_set_weight();
  // otherwise it happens here, once.
  if ((rc = aaclip_finalize_weights(h, stream))) return rc;
  uint8_t* wsb = static_cast<uint8_t*>(workspace);
  uint16_t* patchA = (uint16_t*)(wsb + ws.patchA);
  float* detbuf = (float*)(wsb + ws.det);
  const bool bf = h->bf16;
  BlockRun run;
  run.prof = h; run.bf16 = bf; run.items = B; run.tokens = N; run.W = W; run.F = F; run.causal = 0; run.act = c.act;
  run.x16 = (uint16_t*)(wsb + ws.core.x16); run.xlo = (uint16_t*)(wsb + ws.core.xlo); run.qkv = (uint16_t*)(wsb + ws.core.qkv);
  run.att = (uint16_t*)(wsb + ws.core.att); run.hid = (uint16_t*)(wsb + ws.core.hid);
  run.abuf = (float*)(wsb + ws.core.abuf); run.stats = (float*)(wsb + ws.core.stats); run.st = st;
  if (h->sat_on) { run.sat.counter = h->sat_counter; run.sat.bf16 = bf; }
  float* x = (float*)(wsb + ws.core.x);   // fp32 staging: patch embedding + CLS + positional embedding, until ln_pre
  int& launches = run.launches;

  auto lnorm = [&](const float* in, const float* g, const float* bta, void* o16, float* o32, int rows, int skipP,
                   float* stats_out, uint16_t* o_lo = nullptr) -> int {
    ProfScope ps(h, AACLIP_PROF_LAYERNORM, 0.0, st);
    ++launches;
    return launch_layernorm(bf, in, g, bta, o16, o32, rows, W, skipP, true, st, stats_out, kStatPartials(W), o_lo);
  };
  // the (hi, lo) residual stream back as fp32 rows in the staging buffer (exports only: fp32 ln_post tokens)
  auto stream_to_f32 = [&]() -> int {
    ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
    ++launches;
    const size_t n4 = static_cast<size_t>(M) * W / 4;
    if (bf) launch_pdl(hilo_to_f32_kernel<true>, dim3(148 * 8), dim3(256), 0, st, static_cast<const uint16_t*>(run.x16), static_cast<const uint16_t*>(run.xlo), x, n4);
    else launch_pdl(hilo_to_f32_kernel<false>, dim3(148 * 8), dim3(256), 0, st, static_cast<const uint16_t*>(run.x16), static_cast<const uint16_t*>(run.xlo), x, n4);
    LAUNCH_CHECK("hilo_to_f32_kernel");
    return AACLIP_OK;
  };
  auto l2norm = [&](float* t, int rows) -> int {
    ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
    ++launches;
    const int grid = (rows + 7) / 8;
    if (E == 768) launch_pdl(l2norm_rows_kernel<768>, dim3(grid), dim3(256), 0, st, t, rows);
    else launch_pdl(l2norm_rows_kernel<1024>, dim3(grid), dim3(256), 0, st, t, rows);
    LAUNCH_CHECK("l2norm_rows_kernel");
    return AACLIP_OK;
  };

  // ---- patch embedding + CLS + positional embedding + ln_pre (model/adapter.py:139-156)
  {
    ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
    const size_t gsmem = static_cast<size_t>(3) * c.patch * h->G * c.patch * 2;
    // AACLIP_PATCH_GATHER_SIMPLE=1 forces the general kernel (tests compare the two paths bit for bit)
    static const bool force_simple = getenv("AACLIP_PATCH_GATHER_SIMPLE") != nullptr;
    if (!force_simple && c.patch % 2 == 0 && c.img_size % 2 == 0 && h->KP % 2 == 0 && gsmem <= 160 * 1024 &&
        (reinterpret_cast<uintptr_t>(images) & 7) == 0) {
      // (the opt-in is per image geometry, so it is simply repeated: one host call per forward)
      if (gsmem > 48 * 1024) {
        if (bf) CUDA_TRY(cudaFuncSetAttribute(patch_gather_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(gsmem)));
        else CUDA_TRY(cudaFuncSetAttribute(patch_gather_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(gsmem)));
      }
      if (bf) launch_pdl(patch_gather_kernel<true>, dim3(h->G, B), dim3(256), gsmem, st, images, patchA, c.img_size, h->G, c.patch, h->KP);
      else launch_pdl(patch_gather_kernel<false>, dim3(h->G, B), dim3(256), gsmem, st, images, patchA, c.img_size, h->G, c.patch, h->KP);
    } else {
      if (bf) launch_pdl(patch_gather_simple_kernel<true>, dim3(h->G, B), dim3(256), 0, st, images, patchA, c.img_size, h->G, c.patch, h->KP);
      else launch_pdl(patch_gather_simple_kernel<false>, dim3(h->G, B), dim3(256), 0, st, images, patchA, c.img_size, h->G, c.patch, h->KP);
    }
    LAUNCH_CHECK("patch_gather_kernel"); ++launches;
    run.sat.scan(patchA, static_cast<size_t>(MP) * h->KP, st);
  }
  {
    GemmEpi e; e.out_f32 = x; e.ldo = W; e.pos = h->pos; e.patches = P;
    if ((rc = run.gemm(EPI_PATCH_F32, patchA, h->w_conv, MP, W, h->KP, e))) return rc;
  }
  {
    ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
    launch_pdl(cls_row_kernel, dim3(B), dim3(256), 0, st, x, static_cast<const float*>(h->cls), static_cast<const float*>(h->pos), N, W);
    LAUNCH_CHECK("cls_row_kernel"); ++launches;
  }
  // ln_pre: fp32 staging -> the residual stream in its (hi, lo) form + the LayerNorm partials block 0's folded ln_1 needs
  if ((rc = lnorm(x, h->lnpre_g, h->lnpre_b, run.x16, nullptr, M, 0, run.stats, run.xlo))) return rc;
  run.last_desc = false;                   // an element-wise kernel wrote the stream: rows ascending
  run.sat.scan(run.x16, static_cast<size_t>(M) * W, st);

  // ---- residual attention blocks (model/transformer.py:239-258) + adapters (model/adapter.py:161-172)
  int level_idx = 0;
  for (int i = 0; i < c.layers; ++i) {
    if ((rc = run.run(h->blk[i], i < c.adapt_until ? h->w_adapter[i] : nullptr, c.adapt_weight))) return rc;
    if (level_idx < c.n_levels && c.levels[level_idx] == i + 1) {
      // token tap: x[1:] of every image -> ln_post -> projections (model/adapter.py:171-184). ln_post is folded into
      // the projections: they read the raw 16-bit residual stream (all N rows per image) and the epilogue drops the CLS
      // rows; the fp32 ln_post tokens are only materialised when the caller asks for them (host IQM branch).
      float* tok_f32 = (tokens_out != nullptr) ? tokens_out[level_idx] : nullptr;
      if (tok_f32 != nullptr) {
        if ((rc = stream_to_f32())) return rc;
        if ((rc = lnorm(x, h->lnpost_g, h->lnpost_b, nullptr, tok_f32, MP, P, nullptr))) return rc;
      }
      const int proj_epi = c.proj_relu ? EPI_LN_LEAKY_F32_TOK : EPI_LN_F32_TOK;
      auto tap = [&](size_t pi, float* out, int cols) -> int {
        const AaclipHandle::ProjFold& pf = h->proj[pi];
        GemmEpi e; e.out_f32 = out; e.ldo = cols; e.patches = P;
        e.bias = pf.bias_f; e.colsum = pf.colsum; e.stats = run.stats; e.n_partials = kStatPartials(W); e.inv_width = 1.0f / W;
        return run.gemm_tap(proj_epi, run.x16, pf.w16, M, cols, W, e);
      };
      if ((rc = tap((size_t)level_idx, seg_out[level_idx], E))) return rc;
      if (query_out != nullptr && query_out[level_idx] != nullptr) {
        // IQM branch: query_adapters[l] (model/adapter.py:202-205) on the same tokens, same epilogue
        if ((rc = tap((size_t)c.n_levels + 1 + level_idx, query_out[level_idx], c.query_dim))) return rc;
      }
      if (h->seg_normalize && (rc = l2norm(seg_out[level_idx], MP))) return rc;
      if (level_idx == c.n_levels - 1) {
        if ((rc = tap((size_t)c.n_levels, detbuf, E))) return rc;
        if ((rc = l2norm(detbuf, MP))) return rc;
        ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
        launch_pdl(mean_over_patches_kernel, dim3((E + MEAN_COLS - 1) / MEAN_COLS, B), dim3(MEAN_COLS, MEAN_SEGS), 0, st, static_cast<const float*>(detbuf), det_out, P, E);
        LAUNCH_CHECK("mean_over_patches_kernel");
        ++launches;
      }
      ++level_idx;
    }
  }
  if (cls_out != nullptr) {
    ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
    if (bf) launch_pdl(gather_row_hl_kernel<true>, dim3(B), dim3(256), 0, st, static_cast<const uint16_t*>(run.x16), static_cast<const uint16_t*>(run.xlo), cls_out, N, W, static_cast<const int32_t*>(nullptr));
    else launch_pdl(gather_row_hl_kernel<false>, dim3(B), dim3(256), 0, st, static_cast<const uint16_t*>(run.x16), static_cast<const uint16_t*>(run.xlo), cls_out, N, W, static_cast<const int32_t*>(nullptr));
    LAUNCH_CHECK("gather_row_hl_kernel"); ++launches;
  }
  h->last_launches = launches + run.sat.launches;
  return AACLIP_OK;
}

int aaclip_last_launch_count(AaclipHandle* h) { return h ? h->last_launches : -1; }

int aaclip_finalize_weights(AaclipHandle* h, void* stream) {
  if (!h) return fail(AACLIP_EINVAL, "aaclip_finalize_weights: null handle");
  if (h->f32 || h->folded) return AACLIP_OK;
  if (!h->missing.empty())
    return fail(AACLIP_ESTATE, "aaclip_finalize_weights: %zu weight tensors not set (first: %s)", h->missing.size(), h->missing[0].c_str());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int W = h->cfg.width;
  for (AaclipHandle::ProjFold& pf : h->proj) {
    if (pf.stage == nullptr) continue;      // an optional query adapter that was never set
    if (h->bf16) ln_fold_kernel<true><<<(pf.rows + 7) / 8, 256, 0, st>>>(pf.stage, h->zero_bias, h->lnpost_g, h->lnpost_b, pf.w16, pf.colsum, pf.bias_f, pf.rows, W);
    else ln_fold_kernel<false><<<(pf.rows + 7) / 8, 256, 0, st>>>(pf.stage, h->zero_bias, h->lnpost_g, h->lnpost_b, pf.w16, pf.colsum, pf.bias_f, pf.rows, W);
    LAUNCH_CHECK("ln_fold_kernel");
  }
  if (int rc = fold_blocks(h->blk, h->bf16, W, h->cfg.mlp_dim, st)) return rc;   // synchronises the stream
  for (AaclipHandle::ProjFold& pf : h->proj) {
    if (pf.stage) cudaFree(pf.stage);
    pf.stage = nullptr;
  }
  h->folded = true;
  return AACLIP_OK;
}

int aaclip_saturation_check(AaclipHandle* h, int32_t enable) {
  if (!h) return fail(AACLIP_EINVAL, "aaclip_saturation_check: null handle");
  if (h->f32 && enable) return fail(AACLIP_EINVAL, "aaclip_saturation_check: fp32-accuracy mode stores no 16-bit activations");
  h->sat_on = enable != 0;
  return AACLIP_OK;
}

int aaclip_saturation_count(AaclipHandle* h, uint64_t* count_host, int32_t reset, void* stream) {
  if (!h || !count_host) return fail(AACLIP_EINVAL, "aaclip_saturation_count: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long v = 0;
  CUDA_TRY(cudaMemcpyAsync(&v, h->sat_counter, sizeof(v), cudaMemcpyDeviceToHost, st));
  if (reset) CUDA_TRY(cudaMemsetAsync(h->sat_counter, 0, sizeof(v), st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *count_host = v;
  return AACLIP_OK;
}

int aaclip_stat_columns(void) { return kStatCols; }

int aaclip_profile_enable(AaclipHandle* h, int32_t enable) {
  if (!h) return fail(AACLIP_EINVAL, "aaclip_profile_enable: null handle");
  for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  h->prof.clear();
  h->prof_on = enable != 0;
  return AACLIP_OK;
}

int aaclip_profile_read(AaclipHandle* h, float* ms, int32_t* launches, double* flops) {
  if (!h || !ms || !launches || !flops) return fail(AACLIP_EINVAL, "aaclip_profile_read: null argument");
  for (int i = 0; i < AACLIP_PROF_CLASSES; ++i) { ms[i] = 0.f; launches[i] = 0; flops[i] = 0.0; }
  for (auto& r : h->prof) {
    CUDA_TRY(cudaEventSynchronize(r.b));
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.cls] += t;
    launches[r.cls] += 1;
    flops[r.cls] += r.flops;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  h->prof.clear();
  return AACLIP_OK;
}

size_t aaclip_score_workspace_bytes(int32_t B, int32_t n_levels, int32_t P) {
  if (B <= 0 || n_levels <= 0 || P <= 0) return 0;
  return align_up((size_t)B * n_levels * 2 * P * sizeof(float), 256);
}

static int score_maps_impl(bool seg_raw, const float* const* seg, int32_t n_levels, const float* text,
                           int32_t text_per_image, int32_t B, int32_t G, int32_t E, int32_t S, int32_t test_mode,
                           int32_t ksize, float sigma, void* workspace, size_t workspace_bytes, float* maps_out,
                           const float* det, float* score_out, void* stream) {
  if (!seg || !text || !workspace || !maps_out) return fail(AACLIP_EINVAL, "aaclip_score_maps: null argument");
  if (n_levels <= 0 || n_levels > AACLIP_MAX_LEVELS) return fail(AACLIP_EINVAL, "aaclip_score_maps: n_levels must be 1..4");
  if (B <= 0 || G <= 0 || E <= 0 || S <= 0) return fail(AACLIP_EINVAL, "aaclip_score_maps: bad shape");
  if (!test_mode && n_levels != 1) return fail(AACLIP_EINVAL, "aaclip_score_maps: test_mode=0 takes exactly one level");
  if (test_mode && (ksize < 1 || ksize > MAP_MAXK || (ksize & 1) == 0 || !(sigma > 0.f)))
    return fail(AACLIP_EINVAL, "aaclip_score_maps: ksize must be odd in 1..15 and sigma > 0");
  if (test_mode && ksize / 2 >= G) return fail(AACLIP_EINVAL, "aaclip_score_maps: reflect border needs ksize/2 < G");
  const int P = G * G;
  if (workspace_bytes < aaclip_score_workspace_bytes(B, n_levels, P)) return fail(AACLIP_EINVAL, "aaclip_score_maps: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* s = static_cast<float*>(workspace);
  SegPtrs sp;
  for (int l = 0; l < 4; ++l) sp.p[l] = (l < n_levels) ? seg[l] : nullptr;
  for (int l = 0; l < n_levels; ++l) if (!seg[l]) return fail(AACLIP_EINVAL, "aaclip_score_maps: seg[%d] is null", l);
  const long long tstride = text_per_image ? (long long)E * 2 : 0;
  {
    const long long rows = (long long)B * n_levels * P;
    const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
    if (seg_raw) launch_pdl(similarity_kernel<true>, dim3(grid), dim3(256), 0, st, sp, text, tstride, s, B, n_levels, P, E);
    else launch_pdl(similarity_kernel<false>, dim3(grid), dim3(256), 0, st, sp, text, tstride, s, B, n_levels, P, E);
    LAUNCH_CHECK("similarity_kernel");
  }
  MapTaps taps;
  memset(&taps, 0, sizeof(taps));
  if (test_mode) {
    // kornia/torchvision Gaussian taps: exp(-x^2 / (2 sigma^2)) on x = -(k-1)/2 .. (k-1)/2, normalised
    float sum = 0.f;
    for (int k = 0; k < ksize; ++k) {
      const float xk = static_cast<float>(k) - 0.5f * static_cast<float>(ksize - 1);
      taps.w[k] = expf(-0.5f * (xk / sigma) * (xk / sigma));
      sum += taps.w[k];
    }
    for (int k = 0; k < ksize; ++k) taps.w[k] /= sum;
  }
  {
    static PerDeviceOnce once;
    const size_t smem = 3 * (size_t)P * sizeof(float);
    if (smem > 200 * 1024) return fail(AACLIP_EINVAL, "aaclip_score_maps: grid %dx%d too large", G, G);
    if (int rc = opt_in_smem(once, anomaly_map_kernel, 200 * 1024)) return rc;
    dim3 grid((S + MAP_ROWS - 1) / MAP_ROWS, B);
    launch_pdl(anomaly_map_kernel, grid, dim3(256), smem, st, static_cast<const float*>(s), maps_out, n_levels, G, S, test_mode, ksize, taps);
    LAUNCH_CHECK("anomaly_map_kernel");
  }
  if (det != nullptr && score_out != nullptr) {
    launch_pdl(image_score_kernel, dim3((B + 7) / 8), dim3(256), 0, st, det, text, tstride, score_out, B, E);
    LAUNCH_CHECK("image_score_kernel");
  }
  return AACLIP_OK;
}

int aaclip_score_maps(const float* const* seg, int32_t n_levels, const float* text, int32_t text_per_image,
                      int32_t B, int32_t G, int32_t E, int32_t S, int32_t test_mode, int32_t ksize, float sigma,
                      void* workspace, size_t workspace_bytes, float* maps_out, const float* det, float* score_out,
                      void* stream) {
  return score_maps_impl(false, seg, n_levels, text, text_per_image, B, G, E, S, test_mode, ksize, sigma, workspace,
                         workspace_bytes, maps_out, det, score_out, stream);
}

int aaclip_score_maps_raw(const float* const* seg, int32_t n_levels, const float* text, int32_t text_per_image,
                          int32_t B, int32_t G, int32_t E, int32_t S, int32_t test_mode, int32_t ksize, float sigma,
                          void* workspace, size_t workspace_bytes, float* maps_out, const float* det, float* score_out,
                          void* stream) {
  return score_maps_impl(true, seg, n_levels, text, text_per_image, B, G, E, S, test_mode, ksize, sigma, workspace,
                         workspace_bytes, maps_out, det, score_out, stream);
}

int aaclip_seg_normalize(AaclipHandle* h, int32_t enable) {
  if (!h) return fail(AACLIP_EINVAL, "aaclip_seg_normalize: null handle");
  h->seg_normalize = enable != 0;
  return AACLIP_OK;
}

int aaclip_adapter(int32_t dtype, void* x_hi, void* x_lo, const void* W16, int32_t M, int32_t D, float weight,
                   void* workspace, size_t workspace_bytes, float* stats_out, void* stream) {
  if (!x_hi || !x_lo || !W16 || !workspace || !stats_out) return fail(AACLIP_EINVAL, "aaclip_adapter: null argument");
  if (dtype != AACLIP_DT_F16 && dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "aaclip_adapter: bad dtype %d", dtype);
  if (M <= 0 || (D != 768 && D != 1024)) return fail(AACLIP_EINVAL, "aaclip_adapter: D=%d unsupported (768 or 1024)", D);
  const int np = kStatPartials(D);
  const size_t need = static_cast<size_t>(M) * D * 2 + static_cast<size_t>(M) * np * 2 * sizeof(float);
  if (workspace_bytes < need) return fail(AACLIP_EINVAL, "aaclip_adapter: workspace too small (%zu < %zu)", workspace_bytes, need);
  BlockRun run;
  run.prof = nullptr; run.bf16 = dtype == AACLIP_DT_BF16; run.W = D; run.st = static_cast<cudaStream_t>(stream);
  run.x16 = static_cast<uint16_t*>(x_hi); run.xlo = static_cast<uint16_t*>(x_lo);
  run.abuf = static_cast<float*>(workspace); run.stats = stats_out;
  return run.adapter(static_cast<const uint16_t*>(W16), M, weight);
}

int aaclip_split_hilo(int32_t dtype, const float* x, void* x_hi, void* x_lo, int64_t n, void* stream) {
  if (!x || !x_hi || !x_lo || n < 0 || n % 4 != 0) return fail(AACLIP_EINVAL, "aaclip_split_hilo: bad argument");
  if (dtype != AACLIP_DT_F16 && dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "aaclip_split_hilo: bad dtype %d", dtype);
  if (n == 0) return AACLIP_OK;
  const unsigned grid = static_cast<unsigned>((n / 4 + 255) / 256 < 148 * 8 ? (n / 4 + 255) / 256 : 148 * 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == AACLIP_DT_BF16) f32_to_hilo_kernel<true><<<grid, 256, 0, st>>>(x, static_cast<uint16_t*>(x_hi), static_cast<uint16_t*>(x_lo), static_cast<size_t>(n / 4));
  else f32_to_hilo_kernel<false><<<grid, 256, 0, st>>>(x, static_cast<uint16_t*>(x_hi), static_cast<uint16_t*>(x_lo), static_cast<size_t>(n / 4));
  LAUNCH_CHECK("f32_to_hilo_kernel");
  return AACLIP_OK;
}

int aaclip_join_hilo(int32_t dtype, const void* x_hi, const void* x_lo, float* x, int64_t n, void* stream) {
  if (!x || !x_hi || !x_lo || n < 0 || n % 4 != 0) return fail(AACLIP_EINVAL, "aaclip_join_hilo: bad argument");
  if (dtype != AACLIP_DT_F16 && dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "aaclip_join_hilo: bad dtype %d", dtype);
  if (n == 0) return AACLIP_OK;
  const unsigned grid = static_cast<unsigned>((n / 4 + 255) / 256 < 148 * 8 ? (n / 4 + 255) / 256 : 148 * 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == AACLIP_DT_BF16) hilo_to_f32_kernel<true><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(x_hi), static_cast<const uint16_t*>(x_lo), x, static_cast<size_t>(n / 4));
  else hilo_to_f32_kernel<false><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(x_hi), static_cast<const uint16_t*>(x_lo), x, static_cast<size_t>(n / 4));
  LAUNCH_CHECK("hilo_to_f32_kernel");
  return AACLIP_OK;
}

int aaclip_gemm16_resid_hl(int32_t dtype, const void* A, const void* W, int32_t M, int32_t N, int32_t K,
                           const float* bias, void* x_hi, void* x_lo, float* stats_out, int32_t ldo, void* stream) {
  if (!A || !W || !bias || !x_hi || !x_lo) return fail(AACLIP_EINVAL, "aaclip_gemm16_resid_hl: null argument");
  if (dtype != AACLIP_DT_F16 && dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "aaclip_gemm16_resid_hl: bad dtype %d", dtype);
  if (ldo < N || ldo % 8 != 0) return fail(AACLIP_EINVAL, "aaclip_gemm16_resid_hl: ldo must be >= N and a multiple of 8");
  if (stats_out && N / kStatCols > 16) return fail(AACLIP_EINVAL, "aaclip_gemm16_resid_hl: statistics need N <= %d (16 partials)", 16 * kStatCols);
  GemmEpi e;
  e.bias = bias; e.out_16 = static_cast<uint16_t*>(x_hi); e.out_lo = static_cast<uint16_t*>(x_lo); e.stats_out = stats_out; e.ldo = ldo;
  return launch_gemm(EPI_BIAS_RESID_HL, dtype == AACLIP_DT_BF16, A, W, M, N, K, e, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// =========================================================================================== text tower
struct AaclipTextHandle {
  AaclipTextCfg cfg;
  bool bf16 = false;
  bool f32 = false;       // fp32-accuracy mode
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  float *tok_emb, *pos, *lnf_g, *lnf_b;
  std::vector<BlockW> blk;
  bool folded = false;
  unsigned long long* sat_counter = nullptr;
  bool sat_on = false;
  std::vector<uint16_t*> w_adapter;  // [W,W] each (SimpleAdapter)
  uint16_t* w_proj;                  // [E,W] final SimpleProj
  std::vector<std::string> missing;
};

struct TextWsPlan { CoreWs core; size_t eot, eot16, total; };
static TextWsPlan plan_text_ws(const AaclipTextHandle* h, int n) {
  const size_t W = h->cfg.width, F = h->cfg.mlp_dim;
  const size_t M = (size_t)n * h->cfg.context;
  ArenaPlan p;
  TextWsPlan w;
  w.core = plan_core_ws(p, M, W, F, h->f32 ? 4 : 2);
  w.eot = p.take<float>((size_t)n * W);
  w.eot16 = p.take<float>((size_t)n * W);      // 16-bit in the tensor-core modes, fp32 in fp32 mode
  w.total = p.off;
  return w;
}

extern "C" {

int aaclip_text_create(const AaclipTextCfg* cfg, AaclipTextHandle** out) {
  if (cfg == nullptr || out == nullptr) return fail(AACLIP_EINVAL, "aaclip_text_create: null argument");
  *out = nullptr;
  if (int rc = ensure_device()) return rc;
  const AaclipTextCfg& c = *cfg;
  if (c.width != 768 && c.width != 1024) return fail(AACLIP_EINVAL, "text width %d unsupported (768 or 1024)", c.width);
  if (c.heads * 64 != c.width) return fail(AACLIP_EINVAL, "text head dim must be 64 (width %d, heads %d)", c.width, c.heads);
  if (c.mlp_dim % 256 != 0 || c.embed_dim % 256 != 0) return fail(AACLIP_EINVAL, "text mlp_dim/embed_dim must be multiples of 256");
  if (c.context <= 0 || c.vocab <= 0 || c.layers <= 0 || c.adapt_until < 0 || c.adapt_until > c.layers || c.max_sentences <= 0)
    return fail(AACLIP_EINVAL, "bad text configuration");
  if (c.operand_dtype != AACLIP_DT_F16 && c.operand_dtype != AACLIP_DT_BF16 && c.operand_dtype != AACLIP_DT_F32)
    return fail(AACLIP_EINVAL, "operand_dtype must be AACLIP_DT_F16, AACLIP_DT_BF16 or AACLIP_DT_F32");
  AaclipTextHandle* h = new AaclipTextHandle();
  h->cfg = c;
  h->bf16 = (c.operand_dtype == AACLIP_DT_BF16);
  h->f32 = (c.operand_dtype == AACLIP_DT_F32);
  const size_t wb = h->f32 ? 4 : 2;
  const size_t W = c.width, F = c.mlp_dim, E = c.embed_dim;
  ArenaPlan plan;
  const size_t o_tok = plan.take<float>((size_t)c.vocab * W), o_pos = plan.take<float>((size_t)c.context * W);
  const size_t o_lg = plan.take<float>(W), o_lb = plan.take<float>(W);
  std::vector<BlockOff> bo(c.layers);
  for (int i = 0; i < c.layers; ++i) bo[i] = plan_block(plan, W, F, wb);
  std::vector<size_t> o_ad(c.adapt_until);
  for (int i = 0; i < c.adapt_until; ++i) o_ad[i] = plan.take<uint8_t>(W * W * wb);
  const size_t o_proj = plan.take<uint8_t>(E * W * wb);
  const size_t o_sat = plan.take<unsigned long long>(1);
  h->arena_bytes = plan.off;
  cudaError_t e = cudaMalloc(&h->arena, h->arena_bytes);
  if (e == cudaSuccess) e = cudaMemset(h->arena, 0, h->arena_bytes);
  if (e != cudaSuccess) {
    if (h->arena) cudaFree(h->arena);
    delete h;
    return fail(AACLIP_ECUDA, "text weight arena (%zu bytes): %s", plan.off, cudaGetErrorString(e));
  }
  uint8_t* a = h->arena;
  h->tok_emb = (float*)(a + o_tok); h->pos = (float*)(a + o_pos);
  h->lnf_g = (float*)(a + o_lg); h->lnf_b = (float*)(a + o_lb);
  h->blk.resize(c.layers);
  for (int i = 0; i < c.layers; ++i) bind_block(h->blk[i], a, bo[i]);
  for (int i = 0; i < c.adapt_until; ++i) h->w_adapter.push_back((uint16_t*)(a + o_ad[i]));
  h->w_proj = (uint16_t*)(a + o_proj);
  h->sat_counter = (unsigned long long*)(a + o_sat);
  auto need = [&](const std::string& k) { h->missing.push_back(k); };
  need("token_embedding.weight"); need("positional_embedding"); need("ln_final.weight"); need("ln_final.bias");
  for (int i = 0; i < c.layers; ++i)
    for (const char* t : kBlockTensors) need("blk." + std::to_string(i) + "." + t);
  for (int i = 0; i <= c.adapt_until; ++i) need("text_adapter." + std::to_string(i));
  *out = h;
  return AACLIP_OK;
}

void aaclip_text_destroy(AaclipTextHandle* h) {
  if (h == nullptr) return;
  free_block_staging(h->blk);
  if (h->arena) cudaFree(h->arena);
  delete h;
}

int aaclip_text_missing_weights(AaclipTextHandle* h) { return h ? static_cast<int>(h->missing.size()) : -1; }

int aaclip_text_set_weight(AaclipTextHandle* h, const char* name, const float* data, const int64_t* shape,
                           int32_t ndim, void* stream) {
  if (!h || !name || !data || !shape || ndim <= 0) return fail(AACLIP_EINVAL, "aaclip_text_set_weight: null/empty argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const AaclipTextCfg& c = h->cfg;
  const int64_t W = c.width, F = c.mlp_dim, E = c.embed_dim;
  const int64_t n = numel(shape, ndim);
  std::string k(name);
  if (k.rfind("clipmodel.", 0) == 0) k = k.substr(strlen("clipmodel."));
  auto copy_f32 = [&](float* dst, int64_t expect, const std::string& key) -> int {
    if (n != expect) return fail(AACLIP_EINVAL, "%s: expected %lld elements, got %lld", name, (long long)expect, (long long)n);
    CUDA_TRY(cudaMemcpyAsync(dst, data, expect * sizeof(float), cudaMemcpyDeviceToDevice, st));
    erase_key(h->missing, key);
    return AACLIP_OK;
  };
  auto pack = [&](uint16_t* dst, int64_t R, int64_t C, const std::string& key) -> int {
    if (n != R * C) return fail(AACLIP_EINVAL, "%s: expected %lld x %lld elements, got %lld", name, (long long)R, (long long)C, (long long)n);
    if (int rc = h->f32 ? copy_pad_f32(data, dst, R, C, C, st) : launch_pack(h->bf16, data, dst, R, C, C, st)) return rc;
    erase_key(h->missing, key);
    return AACLIP_OK;
  };
  if (k == "token_embedding.weight") return copy_f32(h->tok_emb, (int64_t)c.vocab * W, k);
  if (k == "positional_embedding") return copy_f32(h->pos, (int64_t)c.context * W, k);
  if (k == "ln_final.weight") return copy_f32(h->lnf_g, W, k);
  if (k == "ln_final.bias") return copy_f32(h->lnf_b, W, k);
  int idx = -1;
  char tail[96];
  if (sscanf(k.c_str(), "transformer.resblocks.%d.%95s", &idx, tail) == 2) {
    if (idx < 0 || idx >= c.layers) return fail(AACLIP_EINVAL, "%s: block index out of range", name);
    return set_block_tensor(h->blk[idx], tail, "blk." + std::to_string(idx) + "." + tail, h->missing, h->folded, h->bf16,
                            h->f32, W, F, name, data, n, st);
  }
  if (sscanf(k.c_str(), "text_adapter.%d.%95s", &idx, tail) == 2) {
    if (strcmp(tail, "fc.0.weight") != 0) return fail(AACLIP_EINVAL, "%s: unknown tensor", name);
    if (idx < 0 || idx > c.adapt_until) return fail(AACLIP_EINVAL, "%s: text_adapter index out of range", name);
    if (idx < c.adapt_until) return pack(h->w_adapter[idx], W, W, "text_adapter." + std::to_string(idx));
    return pack(h->w_proj, E, W, "text_adapter." + std::to_string(idx));
  }
  return fail(AACLIP_EINVAL, "%s: unknown tensor name", name);
}

size_t aaclip_text_workspace_bytes(AaclipTextHandle* h, int32_t n) {
  if (!h || n <= 0) return 0;
  return plan_text_ws(h, n).total;
}

int aaclip_encode_text(AaclipTextHandle* h, const int32_t* tokens, int32_t n, void* workspace, size_t workspace_bytes,
                       float* out, void* stream) {
  if (!h || !tokens || !workspace || !out) return fail(AACLIP_EINVAL, "aaclip_encode_text: null argument");
  if (n <= 0 || n > h->cfg.max_sentences) return fail(AACLIP_EINVAL, "aaclip_encode_text: %d sentences outside 1..%d", n, h->cfg.max_sentences);
  if (!h->missing.empty())
    return fail(AACLIP_ESTATE, "aaclip_encode_text: %zu weight tensors not set (first: %s)", h->missing.size(), h->missing[0].c_str());
  const TextWsPlan ws = plan_text_ws(h, n);
  if (workspace_bytes < ws.total) return fail(AACLIP_EINVAL, "aaclip_encode_text: workspace %zu < required %zu", workspace_bytes, ws.total);
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(AACLIP_EINVAL, "aaclip_encode_text: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const AaclipTextCfg& c = h->cfg;
  const int W = c.width, F = c.mlp_dim, E = c.embed_dim, T = c.context;
  const int M = n * T;
  const bool bf = h->bf16;
  int rc;
  uint8_t* wsb = static_cast<uint8_t*>(workspace);
  if (h->f32) {
    // fp32-accuracy mode: same sequence, un-folded LayerNorms, SIMT fp32 kernels
    BlockRunF32 r32;
    r32.prof = nullptr; r32.items = n; r32.tokens = T; r32.W = W; r32.F = F; r32.causal = 1; r32.act = c.act;
    r32.x = (float*)(wsb + ws.core.x); r32.xn = (float*)(wsb + ws.core.x16); r32.qkv = (float*)(wsb + ws.core.qkv);
    r32.att = (float*)(wsb + ws.core.att); r32.hid = (float*)(wsb + ws.core.hid); r32.abuf = (float*)(wsb + ws.core.abuf);
    r32.st = st;
    float* eot32 = (float*)(wsb + ws.eot);
    float* eotn = (float*)(wsb + ws.eot16);
    text_embed_kernel<<<M, 192, 0, st>>>(tokens, h->tok_emb, h->pos, r32.x, T, W, c.vocab);
    LAUNCH_CHECK("text_embed_kernel");
    for (int i = 0; i < c.layers; ++i)
      if ((rc = r32.run(h->blk[i], i < c.adapt_until ? h->w_adapter[i] : nullptr, c.adapt_weight))) return rc;
    text_gather_eot_kernel<<<n, 192, 0, st>>>(tokens, r32.x, eot32, T, W);
    LAUNCH_CHECK("text_gather_eot_kernel");
    if ((rc = r32.lnorm(eot32, h->lnf_g, h->lnf_b, eotn, n, 0))) return rc;
    GemmEpi e; e.out_f32 = out; e.ldo = E;
    return r32.gemm(EPI_LEAKY_F32, eotn, h->w_proj, n, E, W, e);
  }
  if ((rc = aaclip_text_finalize_weights(h, stream))) return rc;
  BlockRun run;
  if (h->sat_on) { run.sat.counter = h->sat_counter; run.sat.bf16 = bf; }
  run.prof = nullptr; run.bf16 = bf; run.items = n; run.tokens = T; run.W = W; run.F = F; run.causal = 1; run.act = c.act;
  run.x16 = (uint16_t*)(wsb + ws.core.x16); run.xlo = (uint16_t*)(wsb + ws.core.xlo); run.qkv = (uint16_t*)(wsb + ws.core.qkv);
  run.att = (uint16_t*)(wsb + ws.core.att); run.hid = (uint16_t*)(wsb + ws.core.hid);
  run.abuf = (float*)(wsb + ws.core.abuf); run.stats = (float*)(wsb + ws.core.stats); run.st = st;
  float* x32 = (float*)(wsb + ws.core.x);  // fp32 staging of the embedding
  float* eot = (float*)(wsb + ws.eot);
  uint16_t* eot16 = (uint16_t*)(wsb + ws.eot16);

  // token + positional embedding (model/adapter.py:277-281); an identity-LayerNorm pass (NORMALIZE=false) then
  // produces the (hi, lo) form of the stream and the LayerNorm partials block 0's folded ln_1 needs
  launch_pdl(text_embed_kernel, dim3(M), dim3(192), 0, st, tokens, static_cast<const float*>(h->tok_emb), static_cast<const float*>(h->pos), x32, T, W, c.vocab);
  LAUNCH_CHECK("text_embed_kernel");
  if ((rc = launch_layernorm(bf, x32, nullptr, nullptr, run.x16, nullptr, M, W, 0, /*normalize=*/false, st, run.stats, kStatPartials(W), run.xlo))) return rc;
  run.sat.scan(run.x16, static_cast<size_t>(M) * W, st);
  // 12 causal residual attention blocks + adapters (model/adapter.py:284-295)
  for (int i = 0; i < c.layers; ++i)
    if ((rc = run.run(h->blk[i], i < c.adapt_until ? h->w_adapter[i] : nullptr, c.adapt_weight))) return rc;
  // EOT row -> ln_final -> SimpleProj + LeakyReLU (model/adapter.py:296-298)
  if (bf) launch_pdl(text_gather_eot_hl_kernel<true>, dim3(n), dim3(192), 0, st, tokens, static_cast<const uint16_t*>(run.x16), static_cast<const uint16_t*>(run.xlo), eot, T, W);
  else launch_pdl(text_gather_eot_hl_kernel<false>, dim3(n), dim3(192), 0, st, tokens, static_cast<const uint16_t*>(run.x16), static_cast<const uint16_t*>(run.xlo), eot, T, W);
  LAUNCH_CHECK("text_gather_eot_hl_kernel");
  if ((rc = launch_layernorm(bf, eot, h->lnf_g, h->lnf_b, eot16, nullptr, n, W, 0, true, st))) return rc;
  run.sat.scan(eot16, static_cast<size_t>(n) * W, st);
  { GemmEpi e; e.out_f32 = out; e.ldo = E;
    if ((rc = launch_gemm(EPI_LEAKY_F32, bf, eot16, h->w_proj, n, E, W, e, st))) return rc; }
  return AACLIP_OK;
}

int aaclip_text_finalize_weights(AaclipTextHandle* h, void* stream) {
  if (!h) return fail(AACLIP_EINVAL, "aaclip_text_finalize_weights: null handle");
  if (h->f32 || h->folded) return AACLIP_OK;
  if (!h->missing.empty())
    return fail(AACLIP_ESTATE, "aaclip_text_finalize_weights: %zu weight tensors not set (first: %s)", h->missing.size(), h->missing[0].c_str());
  if (int rc = fold_blocks(h->blk, h->bf16, h->cfg.width, h->cfg.mlp_dim, static_cast<cudaStream_t>(stream))) return rc;
  h->folded = true;
  return AACLIP_OK;
}

int aaclip_text_saturation_check(AaclipTextHandle* h, int32_t enable) {
  if (!h) return fail(AACLIP_EINVAL, "aaclip_text_saturation_check: null handle");
  if (h->f32 && enable) return fail(AACLIP_EINVAL, "aaclip_text_saturation_check: fp32-accuracy mode stores no 16-bit activations");
  h->sat_on = enable != 0;
  return AACLIP_OK;
}

int aaclip_text_saturation_count(AaclipTextHandle* h, uint64_t* count_host, int32_t reset, void* stream) {
  if (!h || !count_host) return fail(AACLIP_EINVAL, "aaclip_text_saturation_count: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long v = 0;
  CUDA_TRY(cudaMemcpyAsync(&v, h->sat_counter, sizeof(v), cudaMemcpyDeviceToHost, st));
  if (reset) CUDA_TRY(cudaMemsetAsync(h->sat_counter, 0, sizeof(v), st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *count_host = v;
  return AACLIP_OK;
}

int aaclip_text_anchors(const float* emb, const int32_t* offsets, int32_t n_groups, int32_t E, float* anchors,
                        void* stream) {
  if (!emb || !offsets || !anchors) return fail(AACLIP_EINVAL, "aaclip_text_anchors: null argument");
  if (n_groups <= 0 || (n_groups & 1) || E <= 0) return fail(AACLIP_EINVAL, "aaclip_text_anchors: n_groups must be a positive even number");
  text_ensemble_kernel<<<n_groups, 256, (E + 32) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(emb, offsets, anchors, E);
  LAUNCH_CHECK("text_ensemble_kernel");
  return AACLIP_OK;
}

}  // extern "C"

extern "C" {

// ---- on-device evaluation metrics (metrics.cuh)
size_t aaclip_metrics_workspace_bytes(void) { return 2 * METRIC_PARTIALS * sizeof(float); }

int aaclip_minmax(const float* x, int64_t n, void* workspace, float* minmax, int32_t merge, void* stream) {
  if (!x || !workspace || !minmax || n <= 0) return fail(AACLIP_EINVAL, "aaclip_minmax: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long blocks_needed = (n + 1023) / 1024;
  const int grid = static_cast<int>(blocks_needed < METRIC_PARTIALS ? blocks_needed : METRIC_PARTIALS);
  float* partials = static_cast<float*>(workspace);
  minmax_partial_kernel<<<grid, 256, 0, st>>>(x, n, partials);
  LAUNCH_CHECK("minmax_partial_kernel");
  minmax_final_kernel<<<1, 256, 0, st>>>(partials, grid, minmax, merge);
  LAUNCH_CHECK("minmax_final_kernel");
  return AACLIP_OK;
}

int aaclip_score_histogram(const float* x, const uint8_t* labels, int64_t n, const float* minmax, int32_t n_bins,
                           uint32_t* hist, void* stream) {
  if (!x || !labels || !hist || n <= 0 || n_bins < 2) return fail(AACLIP_EINVAL, "aaclip_score_histogram: bad argument");
  const long long blocks_needed = (n + 255) / 256;
  const int grid = static_cast<int>(blocks_needed < 148 * 16 ? blocks_needed : 148 * 16);
  score_histogram_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, labels, n, minmax, minmax != nullptr, n_bins, hist);
  LAUNCH_CHECK("score_histogram_kernel");
  return AACLIP_OK;
}

int aaclip_auc_ap(const uint32_t* hist, int32_t n_bins, double* out, void* stream) {
  if (!hist || !out || n_bins < 2) return fail(AACLIP_EINVAL, "aaclip_auc_ap: bad argument");
  auc_ap_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(hist, n_bins, out);
  LAUNCH_CHECK("auc_ap_kernel");
  return AACLIP_OK;
}

int aaclip_row_max(const float* x, int32_t rows, int64_t cols, float* out, void* stream) {
  if (!x || !out || rows <= 0 || cols <= 0) return fail(AACLIP_EINVAL, "aaclip_row_max: bad argument");
  row_max_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, cols, out);
  LAUNCH_CHECK("row_max_kernel");
  return AACLIP_OK;
}

/* ---- image pre-processing (include/aaclip.h) ---------------------------------------------------------------------- */
int32_t aaclip_resample_ksize(int32_t in_size, int32_t out_size) {
  if (in_size <= 0 || out_size <= 0) return fail(AACLIP_EINVAL, "aaclip_resample_ksize: sizes must be positive");
  return resample_ksize_host(in_size, out_size);
}

int aaclip_resample_coeffs_host(int32_t in_size, int32_t out_size, int32_t* bounds_host, int32_t* kk_host,
                                int32_t kk_capacity) {
  if (in_size <= 0 || out_size <= 0 || !bounds_host || !kk_host)
    return fail(AACLIP_EINVAL, "aaclip_resample_coeffs_host: bad argument");
  const int ksize = resample_ksize_host(in_size, out_size);
  if (static_cast<int64_t>(kk_capacity) < static_cast<int64_t>(ksize) * out_size)
    return fail(AACLIP_EINVAL, "aaclip_resample_coeffs_host: kk_host holds %d entries, %lld needed", kk_capacity,
                static_cast<long long>(ksize) * out_size);
  resample_coeffs_host(in_size, out_size, ksize, bounds_host, kk_host);
  return AACLIP_OK;
}

size_t aaclip_preprocess_workspace_bytes(int32_t B, int32_t H, int32_t S) {
  if (B <= 0 || H <= 0 || S <= 0) return 0;
  // [B, H, S, 3] uint8 intermediate + the 3 x 256 fp32 normalisation table
  return ((static_cast<size_t>(B) * H * S * 3 + 255) & ~static_cast<size_t>(255)) + 4096;
}

int aaclip_preprocess_u8(const uint8_t* src, int32_t B, int32_t H, int32_t W, int32_t S, const int32_t* xbounds,
                         const int32_t* xkk, const int32_t* ybounds, const int32_t* ykk, const float* mean_host,
                         const float* std_host, void* workspace, size_t workspace_bytes, float* out, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (!src || !xbounds || !xkk || !ybounds || !ykk || !mean_host || !std_host || !out)
    return fail(AACLIP_EINVAL, "aaclip_preprocess_u8: null argument");
  if (B <= 0 || H <= 0 || W <= 0 || S <= 0 || B > 65535 || S > 32768)
    return fail(AACLIP_EINVAL, "aaclip_preprocess_u8: bad shape B=%d H=%d W=%d S=%d", B, H, W, S);
  if (!workspace || workspace_bytes < aaclip_preprocess_workspace_bytes(B, H, S))
    return fail(AACLIP_EINVAL, "aaclip_preprocess_u8: workspace too small (%zu < %zu)", workspace_bytes,
                aaclip_preprocess_workspace_bytes(B, H, S));
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(AACLIP_EINVAL, "aaclip_preprocess_u8: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* tmp = reinterpret_cast<uint8_t*>(workspace);
  float* lut = reinterpret_cast<float*>(tmp + ((static_cast<size_t>(B) * H * S * 3 + 255) & ~static_cast<size_t>(255)));
  const long long rows = static_cast<long long>(B) * H;
  if (rows > 2147483647ll) return fail(AACLIP_EINVAL, "aaclip_preprocess_u8: too many rows");
  const int xks = resample_ksize_host(W, S), yks = resample_ksize_host(H, S);

  resample_lut_kernel<<<3, 256, 0, st>>>(lut, mean_host[0], mean_host[1], mean_host[2], std_host[0], std_host[1], std_host[2]);
  LAUNCH_CHECK("resample_lut_kernel");

  // horizontal pass: unrolled kernel for the usual scale factors (ksize 5: up-scaling / same size, 9: ~2x, 13: ~3x
  // down-scaling) when the output row fits two columns per thread and 8 rows fit shared memory; generic kernel otherwise
  const size_t smem_x = static_cast<size_t>(HROWS_X) * W * 3 + 16 + 64;
  const int half = (S + 1) / 2, threads_x = ((half + 31) / 32) * 32;
  if (xks <= 13 && threads_x <= 1024 && smem_x <= 200 * 1024) {
    const unsigned grid = static_cast<unsigned>((rows + HROWS_X - 1) / HROWS_X);
#define AACLIP_RESAMPLE_H(KS)                                                                                          \
  do {                                                                                                                 \
    if (smem_x > 48 * 1024)                                                                                            \
      CUDA_TRY(cudaFuncSetAttribute(resample_horizontal_x_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                    static_cast<int>(smem_x)));                                                        \
    resample_horizontal_x_kernel<KS><<<grid, threads_x, smem_x, st>>>(src, tmp, W, S, xbounds, xkk, xks, rows,         \
                                                                      rows * W * 3);                                  \
  } while (0)
    if (xks <= 5) AACLIP_RESAMPLE_H(5);
    else if (xks <= 9) AACLIP_RESAMPLE_H(9);
    else AACLIP_RESAMPLE_H(13);
#undef AACLIP_RESAMPLE_H
  } else {
    const size_t smem_g = static_cast<size_t>(HROWS) * W * 3 + 32;
    if (smem_g > 200 * 1024) return fail(AACLIP_EINVAL, "aaclip_preprocess_u8: rows wider than 17 000 pixels are not supported");
    if (smem_g > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(resample_horizontal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem_g)));
    resample_horizontal_kernel<<<static_cast<unsigned>((rows + HROWS - 1) / HROWS), 256, smem_g, st>>>(
        src, tmp, W, S, xbounds, xkk, xks, rows, rows * W * 3);
  }
  LAUNCH_CHECK("resample_horizontal_kernel");

  const dim3 vgrid((S * S + 127) / 128, B);
#define AACLIP_RESAMPLE_V(KS) \
  resample_vertical_normalize_kernel<KS><<<vgrid, 128, 0, st>>>(tmp, out, H, S, ybounds, ykk, yks, lut)
  if (yks <= 5) AACLIP_RESAMPLE_V(5);
  else if (yks <= 9) AACLIP_RESAMPLE_V(9);
  else if (yks <= 13) AACLIP_RESAMPLE_V(13);
  else AACLIP_RESAMPLE_V(0);
#undef AACLIP_RESAMPLE_V
  LAUNCH_CHECK("resample_vertical_normalize_kernel");
  return AACLIP_OK;
}

int aaclip_gemm16(int32_t dtype, const void* A, const void* W, int32_t M, int32_t N, int32_t K, int32_t epi,
                  const float* bias, float* out_f32, void* out_16, int32_t ldo, void* stream) {
  if (!A || !W) return fail(AACLIP_EINVAL, "aaclip_gemm16: null operand");
  if (dtype != AACLIP_DT_F16 && dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "aaclip_gemm16: bad dtype %d", dtype);
  if (epi == EPI_PATCH_F32 || epi < 0 || epi > EPI_BIAS_F32) return fail(AACLIP_EINVAL, "aaclip_gemm16: epilogue %d is not available here (folded-LayerNorm epilogues: aaclip_gemm16_ln)", epi);
  const bool needs_bias = epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16 || epi == EPI_BIAS_QGELU_BF16 ||
                          epi == EPI_BIAS_RESID_F32 || epi == EPI_BIAS_F32;
  const bool out16 = epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16 || epi == EPI_BIAS_QGELU_BF16;
  if (needs_bias && !bias) return fail(AACLIP_EINVAL, "aaclip_gemm16: epilogue %d needs a bias", epi);
  if (out16 && !out_16) return fail(AACLIP_EINVAL, "aaclip_gemm16: epilogue %d needs out_16", epi);
  if (!out16 && !out_f32) return fail(AACLIP_EINVAL, "aaclip_gemm16: epilogue %d needs out_f32", epi);
  if (ldo < N || ldo % 8 != 0) return fail(AACLIP_EINVAL, "aaclip_gemm16: ldo must be >= N and a multiple of 8");
  GemmEpi e;
  e.bias = bias; e.out_f32 = out_f32; e.out_16 = reinterpret_cast<uint16_t*>(out_16); e.ldo = ldo;
  return launch_gemm(epi, dtype == AACLIP_DT_BF16, A, W, M, N, K, e, static_cast<cudaStream_t>(stream));
}

int aaclip_ln_fold(int32_t dtype, const float* W, const float* b, const float* gamma, const float* beta, int32_t N,
                   int32_t K, void* W16, float* colsum, float* bias_f, void* stream) {
  if (!W || !b || !gamma || !beta || !W16 || !colsum || !bias_f || N <= 0 || K <= 0)
    return fail(AACLIP_EINVAL, "aaclip_ln_fold: bad argument");
  if (dtype != AACLIP_DT_F16 && dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "aaclip_ln_fold: bad dtype %d", dtype);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint16_t* o = reinterpret_cast<uint16_t*>(W16);
  if (dtype == AACLIP_DT_BF16) ln_fold_kernel<true><<<(N + 7) / 8, 256, 0, st>>>(W, b, gamma, beta, o, colsum, bias_f, N, K);
  else ln_fold_kernel<false><<<(N + 7) / 8, 256, 0, st>>>(W, b, gamma, beta, o, colsum, bias_f, N, K);
  LAUNCH_CHECK("ln_fold_kernel");
  return AACLIP_OK;
}

int aaclip_gemm16_ln(int32_t dtype, const void* A, const void* W16, int32_t M, int32_t N, int32_t K, int32_t epi,
                     const float* bias_f, const float* colsum, const float* stats, int32_t n_partials,
                     int32_t ln_width, void* out_16, int32_t ldo, void* stream) {
  if (!A || !W16 || !bias_f || !colsum || !stats || !out_16) return fail(AACLIP_EINVAL, "aaclip_gemm16_ln: null argument");
  if (dtype != AACLIP_DT_F16 && dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "aaclip_gemm16_ln: bad dtype %d", dtype);
  if (epi != EPI_LN_BIAS_BF16 && epi != EPI_LN_BIAS_GELU_BF16 && epi != EPI_LN_BIAS_QGELU_BF16)
    return fail(AACLIP_EINVAL, "aaclip_gemm16_ln: epilogue %d is not a folded-LayerNorm epilogue", epi);
  if (n_partials < 1 || n_partials > 16 || ln_width <= 0) return fail(AACLIP_EINVAL, "aaclip_gemm16_ln: n_partials must be 1..16");
  if (ldo < N || ldo % 8 != 0) return fail(AACLIP_EINVAL, "aaclip_gemm16_ln: ldo must be >= N and a multiple of 8");
  GemmEpi e;
  e.bias = bias_f; e.colsum = colsum; e.stats = stats; e.n_partials = n_partials; e.inv_width = 1.0f / ln_width;
  e.out_16 = reinterpret_cast<uint16_t*>(out_16); e.ldo = ldo;
  return launch_gemm(epi, dtype == AACLIP_DT_BF16, A, W16, M, N, K, e, static_cast<cudaStream_t>(stream));
}

int aaclip_gemm16_ln_tok(int32_t dtype, const void* A, const void* W16, int32_t M, int32_t N, int32_t K, int32_t leaky,
                         const float* bias_f, const float* colsum, const float* stats, int32_t n_partials,
                         int32_t ln_width, int32_t patches, float* out_f32, int32_t ldo, void* stream) {
  if (!A || !W16 || !bias_f || !colsum || !stats || !out_f32) return fail(AACLIP_EINVAL, "aaclip_gemm16_ln_tok: null argument");
  if (dtype != AACLIP_DT_F16 && dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "aaclip_gemm16_ln_tok: bad dtype %d", dtype);
  if (n_partials < 1 || n_partials > 16 || ln_width <= 0) return fail(AACLIP_EINVAL, "aaclip_gemm16_ln_tok: n_partials must be 1..16");
  if (patches <= 0 || M % (patches + 1) != 0) return fail(AACLIP_EINVAL, "aaclip_gemm16_ln_tok: M must be a multiple of patches + 1");
  if (ldo < N || ldo % 4 != 0) return fail(AACLIP_EINVAL, "aaclip_gemm16_ln_tok: ldo must be >= N and a multiple of 4");
  GemmEpi e;
  e.bias = bias_f; e.colsum = colsum; e.stats = stats; e.n_partials = n_partials; e.inv_width = 1.0f / ln_width;
  e.out_f32 = out_f32; e.ldo = ldo; e.patches = patches;
  return launch_gemm(leaky ? EPI_LN_LEAKY_F32_TOK : EPI_LN_F32_TOK, dtype == AACLIP_DT_BF16, A, W16, M, N, K, e,
                     static_cast<cudaStream_t>(stream));
}

int aaclip_gemm16_resid(int32_t dtype, const void* A, const void* W, int32_t M, int32_t N, int32_t K,
                        const float* bias, float* x, void* x16_out, float* stats_out, int32_t ldo, void* stream) {
  if (!A || !W || !bias || !x) return fail(AACLIP_EINVAL, "aaclip_gemm16_resid: null argument");
  if (dtype != AACLIP_DT_F16 && dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "aaclip_gemm16_resid: bad dtype %d", dtype);
  if (ldo < N || ldo % 8 != 0) return fail(AACLIP_EINVAL, "aaclip_gemm16_resid: ldo must be >= N and a multiple of 8");
  if (stats_out && N / kStatCols > 16) return fail(AACLIP_EINVAL, "aaclip_gemm16_resid: statistics need N <= %d (16 partials)", 16 * kStatCols);
  GemmEpi e;
  e.bias = bias; e.out_f32 = x; e.out_16 = reinterpret_cast<uint16_t*>(x16_out); e.stats_out = stats_out; e.ldo = ldo;
  return launch_gemm(EPI_BIAS_RESID_F32, dtype == AACLIP_DT_BF16, A, W, M, N, K, e, static_cast<cudaStream_t>(stream));
}

int aaclip_attention(int32_t dtype, const void* qkv, void* out, int32_t B, int32_t N, int32_t D, int32_t causal,
                     void* stream) {
  if (!qkv || !out) return fail(AACLIP_EINVAL, "aaclip_attention: null argument");
  if (dtype != AACLIP_DT_F16 && dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "aaclip_attention: bad dtype %d", dtype);
  return launch_attention(dtype == AACLIP_DT_BF16, qkv, out, B, N, D, causal, static_cast<cudaStream_t>(stream));
}

int aaclip_layernorm(int32_t dtype, const float* x, const float* gamma, const float* beta, void* out_16,
                     float* out_f32, int32_t rows_out, int32_t D, int32_t skip_cls_P, void* stream) {
  if (!x || !gamma || !beta) return fail(AACLIP_EINVAL, "aaclip_layernorm: null argument");
  if (!out_16 && !out_f32) return fail(AACLIP_EINVAL, "aaclip_layernorm: no output requested");
  if (dtype != AACLIP_DT_F16 && dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "aaclip_layernorm: bad dtype %d", dtype);
  return launch_layernorm(dtype == AACLIP_DT_BF16, x, gamma, beta, out_16, out_f32, rows_out, D, skip_cls_P, true,
                          static_cast<cudaStream_t>(stream));
}

int aaclip_pack16(int32_t dtype, const float* src, void* dst_16, int64_t R, int64_t C, int64_t Cp, void* stream) {
  if (!src || !dst_16 || R <= 0 || C <= 0 || Cp < C) return fail(AACLIP_EINVAL, "aaclip_pack16: bad argument");
  if (dtype != AACLIP_DT_F16 && dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "aaclip_pack16: bad dtype %d", dtype);
  return launch_pack(dtype == AACLIP_DT_BF16, src, dst_16, R, C, Cp, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
