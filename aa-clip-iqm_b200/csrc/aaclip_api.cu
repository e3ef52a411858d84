// C-ABI + host orchestration of the B200-native AA-CLIP hot path. See include/aaclip.h for the contract.
// Compiled with: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/aaclip.h"
#include "attention_tcgen05.cuh"
#include "elementwise.cuh"
#include "gemm_tcgen05.cuh"

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

int g_num_sms = 0;
int ensure_device() {
  if (g_num_sms > 0) return AACLIP_OK;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(AACLIP_ENOSYS, "device '%s' is sm_%d%d; this library is sm_100a only", prop.name,
                                    prop.major, prop.minor);
  g_num_sms = prop.multiProcessorCount;
  return AACLIP_OK;
}

// ------------------------------------------------------------------------------------------ kernel launchers
#ifndef AACLIP_GEMM_PAIR
#define AACLIP_GEMM_PAIR 1        // 1: CTA-pair kernel (cta_group::2); 0: single-CTA kernel (kept for A/B measurement)
#endif
#ifndef AACLIP_GEMM_PAIR_BN128_MAX_N
#define AACLIP_GEMM_PAIR_BN128_MAX_N 0   // N <= this uses 256x128 pair tiles (less wave quantisation on narrow GEMMs)
#endif

template <int EPI, bool BF16>
int launch_gemm_t(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const GemmEpi& epi,
                  cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<EPI, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  gemm_cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = ((M + gemm_cfg::BM - 1) / gemm_cfg::BM) * (N / gemm_cfg::BN);
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  gemm_tcgen05_kernel<EPI, BF16><<<grid, gemm_cfg::THREADS, gemm_cfg::SMEM_BYTES, st>>>(ta, tb, M, N, K, epi);
  LAUNCH_CHECK("gemm_tcgen05_kernel");
  return AACLIP_OK;
}

template <int EPI, bool BF16, int BN>
int launch_gemm_pair_t(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const GemmEpi& epi,
                       cudaStream_t st) {
  using C = GemmPairCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(gemm_pair_kernel<EPI, BF16, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = ((M + 2 * C::BM - 1) / (2 * C::BM)) * (N / BN);
  const int pairs = tiles < g_num_sms / 2 ? tiles : g_num_sms / 2;
  gemm_pair_kernel<EPI, BF16, BN><<<2 * pairs, C::THREADS, C::SMEM_BYTES, st>>>(ta, tb, M, N, K, epi);
  LAUNCH_CHECK("gemm_pair_kernel");
  return AACLIP_OK;
}

int launch_gemm(int epi_mode, bool bf16, const void* A, const void* W, int M, int N, int K, const GemmEpi& epi,
                cudaStream_t st) {
  if (int rc = ensure_device()) return rc;
  if (M <= 0 || N <= 0 || K <= 0 || N % gemm_cfg::BN != 0 || K % gemm_cfg::BK != 0)
    return fail(AACLIP_EINVAL, "gemm: unsupported shape M=%d N=%d K=%d (need N%%256==0, K%%64==0)", M, N, K);
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return fail(AACLIP_EINVAL, "gemm: operands must be 16-byte aligned");
  const bool pair = AACLIP_GEMM_PAIR != 0;
  const bool bn128 = pair && N <= AACLIP_GEMM_PAIR_BN128_MAX_N;
  CUtensorMap ta, tb;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {gemm_cfg::BK, gemm_cfg::BM};
    if (int rc = make_tmap_16(&ta, A, 2, dims, str, box, bf16)) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t str[1] = {(uint64_t)K * 2};
    // single-CTA kernel: the whole 256-row W tile; pair kernel: this CTA's half of the (256 | 128)-row tile
    const uint32_t box[2] = {gemm_cfg::BK, (uint32_t)(pair ? (bn128 ? 64 : 128) : gemm_cfg::BN)};
    if (int rc = make_tmap_16(&tb, W, 2, dims, str, box, bf16)) return rc;
  }
#define AACLIP_GEMM_CASE(E)                                                                                       \
  case E:                                                                                                         \
    if (!pair) return bf16 ? launch_gemm_t<E, true>(ta, tb, M, N, K, epi, st) : launch_gemm_t<E, false>(ta, tb, M, N, K, epi, st); \
    if (bn128) return bf16 ? launch_gemm_pair_t<E, true, 128>(ta, tb, M, N, K, epi, st) : launch_gemm_pair_t<E, false, 128>(ta, tb, M, N, K, epi, st); \
    return bf16 ? launch_gemm_pair_t<E, true, 256>(ta, tb, M, N, K, epi, st) : launch_gemm_pair_t<E, false, 256>(ta, tb, M, N, K, epi, st);
  switch (epi_mode) {
    AACLIP_GEMM_CASE(EPI_BIAS_BF16)
    AACLIP_GEMM_CASE(EPI_BIAS_GELU_BF16)
    AACLIP_GEMM_CASE(EPI_BIAS_QGELU_BF16)
    AACLIP_GEMM_CASE(EPI_BIAS_RESID_F32)
    AACLIP_GEMM_CASE(EPI_F32)
    AACLIP_GEMM_CASE(EPI_LEAKY_F32)
    AACLIP_GEMM_CASE(EPI_PATCH_F32)
    AACLIP_GEMM_CASE(EPI_BIAS_F32)
#undef AACLIP_GEMM_CASE
    default: return fail(AACLIP_EINVAL, "gemm: unknown epilogue %d", epi_mode);
  }
}

template <bool BF16>
int launch_attention_t(const CUtensorMap& tq, void* out, int B, int N, int D, int causal, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(attention_tcgen05_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  attn_cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((N + attn_cfg::BQ * attn_cfg::QT - 1) / (attn_cfg::BQ * attn_cfg::QT), D / 64, B);
  const float scale_log2 = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
  attention_tcgen05_kernel<BF16><<<grid, attn_cfg::THREADS, attn_cfg::SMEM_BYTES, st>>>(
      tq, reinterpret_cast<uint16_t*>(out), N, D, causal, scale_log2);
  LAUNCH_CHECK("attention_tcgen05_kernel");
  return AACLIP_OK;
}

int launch_attention(bool bf16, const void* qkv, void* out, int B, int N, int D, int causal, cudaStream_t st) {
  if (int rc = ensure_device()) return rc;
  if (B <= 0 || N <= 0 || D <= 0 || D % 64 != 0) return fail(AACLIP_EINVAL, "attention: bad shape B=%d N=%d D=%d", B, N, D);
  CUtensorMap tq;
  const uint64_t dims[3] = {(uint64_t)3 * D, (uint64_t)N, (uint64_t)B};
  const uint64_t str[2] = {(uint64_t)3 * D * 2, (uint64_t)N * 3 * D * 2};
  const uint32_t box[3] = {attn_cfg::HD, attn_cfg::BKV, 1};
  if (int rc = make_tmap_16(&tq, qkv, 3, dims, str, box, bf16)) return rc;
  return bf16 ? launch_attention_t<true>(tq, out, B, N, D, causal, st) : launch_attention_t<false>(tq, out, B, N, D, causal, st);
}

template <int D, bool BF16>
void launch_layernorm_t(const float* x, const float* g, const float* b, void* out16, float* out_f32, int rows,
                        int skipP, bool normalize, cudaStream_t st) {
  const int threads = 256, grid = (rows + 7) / 8;
  uint16_t* o = reinterpret_cast<uint16_t*>(out16);
  if (normalize) layernorm_kernel<D, true, BF16><<<grid, threads, 0, st>>>(x, g, b, o, out_f32, rows, skipP, 1e-5f);
  else layernorm_kernel<D, false, BF16><<<grid, threads, 0, st>>>(x, g, b, o, out_f32, rows, skipP, 1e-5f);
}

int launch_layernorm(bool bf16, const float* x, const float* g, const float* b, void* out16, float* out_f32, int rows,
                     int D, int skipP, bool normalize, cudaStream_t st) {
  if (rows <= 0) return AACLIP_OK;
  if (D == 1024) {
    if (bf16) launch_layernorm_t<1024, true>(x, g, b, out16, out_f32, rows, skipP, normalize, st);
    else launch_layernorm_t<1024, false>(x, g, b, out16, out_f32, rows, skipP, normalize, st);
  } else if (D == 768) {
    if (bf16) launch_layernorm_t<768, true>(x, g, b, out16, out_f32, rows, skipP, normalize, st);
    else launch_layernorm_t<768, false>(x, g, b, out16, out_f32, rows, skipP, normalize, st);
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
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < D; i += blockDim.x) out[static_cast<size_t>(b) * D + i] = x[static_cast<size_t>(b) * tokens * D + i];
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

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// =========================================================================================== handle
struct BlockW {
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *b_qkv, *b_out, *b_fc, *b_proj;
  uint16_t *w_qkv, *w_out, *w_fc, *w_proj;
};

struct AaclipHandle {
  AaclipCfg cfg;
  bool bf16 = false;      // 16-bit operand dtype: fp16 (default) or bf16
  int G, P, N, KP;        // grid side, patches, tokens, padded patch K
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  // fp32 params
  float *cls, *pos, *lnpre_g, *lnpre_b, *lnpost_g, *lnpost_b;
  uint16_t* w_conv;  // [W, KP]
  std::vector<BlockW> blk;
  std::vector<uint16_t*> w_adapter;  // [W,W] each
  std::vector<uint16_t*> w_seg;      // [E,W] each
  uint16_t* w_det;                   // [E,W]
  std::vector<std::string> missing;
  int last_launches = 0;
  // optional per-kernel-class device timing (aaclip_profile_*)
  struct ProfRec { int cls; double flops; cudaEvent_t a, b; };
  bool prof_on = false;
  std::vector<ProfRec> prof;
};

// Records a CUDA-event pair around one launch on the launch stream when profiling is enabled.
struct ProfScope {
  AaclipHandle* h;
  cudaStream_t st;
  size_t idx = 0;
  bool on;
  ProfScope(AaclipHandle* h_, int cls, double flops, cudaStream_t st_) : h(h_), st(st_), on(h_->prof_on) {
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

struct ArenaPlan {
  size_t off = 0;
  template <typename T>
  size_t take(size_t n) {
    size_t o = off;
    off = align_up(off + n * sizeof(T), 256);
    return o;
  }
};

void mark_loaded(AaclipHandle* h, const std::string& key) {
  for (size_t i = 0; i < h->missing.size(); ++i)
    if (h->missing[i] == key) { h->missing.erase(h->missing.begin() + i); return; }
}

int64_t numel(const int64_t* shape, int ndim) {
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  return n;
}

}  // namespace

extern "C" {

const char* aaclip_version(void) { return "aaclip_b200 0.1 sm_100a"; }
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
  if (c.operand_dtype != AACLIP_DT_F16 && c.operand_dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "operand_dtype must be AACLIP_DT_F16 or AACLIP_DT_BF16");

  AaclipHandle* h = new AaclipHandle();
  h->cfg = c;
  h->bf16 = (c.operand_dtype == AACLIP_DT_BF16);
  h->G = c.img_size / c.patch;
  h->P = h->G * h->G;
  h->N = h->P + 1;
  const int k_raw = 3 * c.patch * c.patch;
  h->KP = (k_raw + 63) / 64 * 64;
  const size_t W = c.width, F = c.mlp_dim, E = c.embed_dim;

  ArenaPlan plan;
  const size_t o_cls = plan.take<float>(W), o_pos = plan.take<float>((size_t)h->N * W);
  const size_t o_lpg = plan.take<float>(W), o_lpb = plan.take<float>(W), o_log = plan.take<float>(W), o_lob = plan.take<float>(W);
  const size_t o_conv = plan.take<uint16_t>(W * h->KP);
  struct BO { size_t ln1g, ln1b, ln2g, ln2b, bqkv, bout, bfc, bproj, wqkv, wout, wfc, wproj; };
  std::vector<BO> bo(c.layers);
  for (int i = 0; i < c.layers; ++i) {
    bo[i].ln1g = plan.take<float>(W); bo[i].ln1b = plan.take<float>(W);
    bo[i].ln2g = plan.take<float>(W); bo[i].ln2b = plan.take<float>(W);
    bo[i].bqkv = plan.take<float>(3 * W); bo[i].bout = plan.take<float>(W);
    bo[i].bfc = plan.take<float>(F); bo[i].bproj = plan.take<float>(W);
    bo[i].wqkv = plan.take<uint16_t>(3 * W * W); bo[i].wout = plan.take<uint16_t>(W * W);
    bo[i].wfc = plan.take<uint16_t>(F * W); bo[i].wproj = plan.take<uint16_t>(W * F);
  }
  std::vector<size_t> o_ad(c.adapt_until), o_seg(c.n_levels);
  for (int i = 0; i < c.adapt_until; ++i) o_ad[i] = plan.take<uint16_t>(W * W);
  for (int l = 0; l < c.n_levels; ++l) o_seg[l] = plan.take<uint16_t>(E * W);
  const size_t o_det = plan.take<uint16_t>(E * W);

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
  for (int i = 0; i < c.layers; ++i) {
    BlockW& b = h->blk[i];
    b.ln1_g = (float*)(a + bo[i].ln1g); b.ln1_b = (float*)(a + bo[i].ln1b);
    b.ln2_g = (float*)(a + bo[i].ln2g); b.ln2_b = (float*)(a + bo[i].ln2b);
    b.b_qkv = (float*)(a + bo[i].bqkv); b.b_out = (float*)(a + bo[i].bout);
    b.b_fc = (float*)(a + bo[i].bfc); b.b_proj = (float*)(a + bo[i].bproj);
    b.w_qkv = (uint16_t*)(a + bo[i].wqkv); b.w_out = (uint16_t*)(a + bo[i].wout);
    b.w_fc = (uint16_t*)(a + bo[i].wfc); b.w_proj = (uint16_t*)(a + bo[i].wproj);
  }
  for (int i = 0; i < c.adapt_until; ++i) h->w_adapter.push_back((uint16_t*)(a + o_ad[i]));
  for (int l = 0; l < c.n_levels; ++l) h->w_seg.push_back((uint16_t*)(a + o_seg[l]));
  h->w_det = (uint16_t*)(a + o_det);

  // the set of tensors that must arrive through aaclip_set_weight before forward may run
  auto need = [&](const std::string& k) { h->missing.push_back(k); };
  need("conv1.weight"); need("class_embedding"); need("positional_embedding");
  need("ln_pre.weight"); need("ln_pre.bias"); need("ln_post.weight"); need("ln_post.bias");
  for (int i = 0; i < c.layers; ++i) {
    const std::string p = "blk." + std::to_string(i) + ".";
    for (const char* s : {"ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias", "attn.in_proj_weight",
                          "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias", "mlp.c_fc.weight",
                          "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias"})
      need(p + s);
  }
  for (int i = 0; i < c.adapt_until; ++i) need("adapter." + std::to_string(i));
  for (int l = 0; l < c.n_levels; ++l) need("seg." + std::to_string(l));
  need("det");
  *out = h;
  return AACLIP_OK;
}

void aaclip_destroy(AaclipHandle* h) {
  if (h == nullptr) return;
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
    mark_loaded(h, key);
    return AACLIP_OK;
  };
  auto pack = [&](uint16_t* dst, int64_t R, int64_t C, int64_t Cp, const std::string& key) -> int {
    if (n != R * C) return fail(AACLIP_EINVAL, "%s: expected %lld x %lld elements, got %lld", name, (long long)R, (long long)C, (long long)n);
    if (int rc = launch_pack(h->bf16, data, dst, R, C, Cp, st)) return rc;
    mark_loaded(h, key);
    return AACLIP_OK;
  };

  if (k == "conv1.weight") return pack(h->w_conv, W, 3 * c.patch * c.patch, h->KP, k);
  if (k == "class_embedding") return copy_f32(h->cls, W, k);
  if (k == "positional_embedding") return copy_f32(h->pos, (int64_t)h->N * W, k);
  if (k == "ln_pre.weight") return copy_f32(h->lnpre_g, W, k);
  if (k == "ln_pre.bias") return copy_f32(h->lnpre_b, W, k);
  if (k == "ln_post.weight") return copy_f32(h->lnpost_g, W, k);
  if (k == "ln_post.bias") return copy_f32(h->lnpost_b, W, k);

  int idx = -1;
  char tail[96];
  if (sscanf(k.c_str(), "transformer.resblocks.%d.%95s", &idx, tail) == 2) {
    if (idx < 0 || idx >= c.layers) return fail(AACLIP_EINVAL, "%s: block index out of range", name);
    BlockW& b = h->blk[idx];
    const std::string t(tail), key = "blk." + std::to_string(idx) + "." + t;
    if (t == "ln_1.weight") return copy_f32(b.ln1_g, W, key);
    if (t == "ln_1.bias") return copy_f32(b.ln1_b, W, key);
    if (t == "ln_2.weight") return copy_f32(b.ln2_g, W, key);
    if (t == "ln_2.bias") return copy_f32(b.ln2_b, W, key);
    if (t == "attn.in_proj_bias") return copy_f32(b.b_qkv, 3 * W, key);
    if (t == "attn.out_proj.bias") return copy_f32(b.b_out, W, key);
    if (t == "mlp.c_fc.bias") return copy_f32(b.b_fc, F, key);
    if (t == "mlp.c_proj.bias") return copy_f32(b.b_proj, W, key);
    if (t == "attn.in_proj_weight") return pack(b.w_qkv, 3 * W, W, W, key);
    if (t == "attn.out_proj.weight") return pack(b.w_out, W, W, W, key);
    if (t == "mlp.c_fc.weight") return pack(b.w_fc, F, W, W, key);
    if (t == "mlp.c_proj.weight") return pack(b.w_proj, W, F, F, key);
    return fail(AACLIP_EINVAL, "%s: unknown block tensor", name);
  }
  if (sscanf(k.c_str(), "image_adapter.layer_adapters.%d.%95s", &idx, tail) == 2) {
    if (strcmp(tail, "fc.0.weight") != 0) return fail(AACLIP_EINVAL, "%s: unknown tensor", name);
    if (idx < 0 || idx >= c.adapt_until) return fail(AACLIP_EINVAL, "%s: adapter index out of range", name);
    return pack(h->w_adapter[idx], W, W, W, "adapter." + std::to_string(idx));
  }
  if (sscanf(k.c_str(), "image_adapter.seg_proj.%d.fc.%95s", &idx, tail) == 2) {
    if (idx < 0 || idx >= c.n_levels) return fail(AACLIP_EINVAL, "%s: seg_proj index out of range", name);
    if (strcmp(tail, "weight") != 0 && strcmp(tail, "0.weight") != 0) return fail(AACLIP_EINVAL, "%s: unknown tensor", name);
    return pack(h->w_seg[idx], E, W, W, "seg." + std::to_string(idx));
  }
  if (k == "image_adapter.det_proj.fc.weight" || k == "image_adapter.det_proj.fc.0.weight") return pack(h->w_det, E, W, W, "det");
  return fail(AACLIP_EINVAL, "%s: unknown tensor name", name);
}

// workspace carve-up shared by aaclip_workspace_bytes and aaclip_forward
struct WsPlan {
  size_t x, xn, qkv, att, hid, abuf, patchA, tok, det, total;
};
static WsPlan plan_ws(const AaclipHandle* h, int B) {
  const size_t W = h->cfg.width, F = h->cfg.mlp_dim, E = h->cfg.embed_dim;
  const size_t M = (size_t)B * h->N, MP = (size_t)B * h->P;
  ArenaPlan p;
  WsPlan w;
  w.x = p.take<float>(M * W);
  w.xn = p.take<uint16_t>(M * W);
  w.qkv = p.take<uint16_t>(M * 3 * W);
  w.att = p.take<uint16_t>(M * W);
  w.hid = p.take<uint16_t>(M * F);
  w.abuf = p.take<float>(M * W);
  w.patchA = p.take<uint16_t>(MP * h->KP);
  w.tok = p.take<uint16_t>(MP * W);
  w.det = p.take<float>(MP * E);
  w.total = p.off;
  return w;
}

size_t aaclip_workspace_bytes(AaclipHandle* h, int32_t B) {
  if (!h || B <= 0) return 0;
  return plan_ws(h, B).total;
}

int aaclip_forward(AaclipHandle* h, const float* images, int32_t B, void* workspace, size_t workspace_bytes,
                   float* const* seg_out, float* det_out, float* const* tokens_out, float* cls_out, void* stream) {
  if (!h || !images || !workspace || !seg_out || !det_out) return fail(AACLIP_EINVAL, "aaclip_forward: null argument");
  if (B <= 0 || B > h->cfg.max_batch) return fail(AACLIP_EINVAL, "aaclip_forward: batch %d outside 1..%d", B, h->cfg.max_batch);
  if (!h->missing.empty())
    return fail(AACLIP_ESTATE, "aaclip_forward: %zu weight tensors not set (first: %s)", h->missing.size(), h->missing[0].c_str());
  const WsPlan ws = plan_ws(h, B);
  if (workspace_bytes < ws.total) return fail(AACLIP_EINVAL, "aaclip_forward: workspace %zu < required %zu", workspace_bytes, ws.total);
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(AACLIP_EINVAL, "aaclip_forward: workspace must be 256-byte aligned");
  for (int l = 0; l < h->cfg.n_levels; ++l)
    if (!seg_out[l]) return fail(AACLIP_EINVAL, "aaclip_forward: seg_out[%d] is null", l);

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const AaclipCfg& c = h->cfg;
  const int W = c.width, F = c.mlp_dim, E = c.embed_dim, N = h->N, P = h->P;
  const int M = B * N, MP = B * P;
  uint8_t* wsb = static_cast<uint8_t*>(workspace);
  float* x = (float*)(wsb + ws.x);
  uint16_t* xn = (uint16_t*)(wsb + ws.xn);
  uint16_t* qkv = (uint16_t*)(wsb + ws.qkv);
  uint16_t* att = (uint16_t*)(wsb + ws.att);
  uint16_t* hid = (uint16_t*)(wsb + ws.hid);
  float* abuf = (float*)(wsb + ws.abuf);
  uint16_t* patchA = (uint16_t*)(wsb + ws.patchA);
  uint16_t* tok = (uint16_t*)(wsb + ws.tok);
  float* detbuf = (float*)(wsb + ws.det);
  int launches = 0;
  int rc;
  const bool bf = h->bf16;

  // every launch goes through one of these wrappers: launch counting + optional per-class event timing
  auto gemm = [&](int epi, const void* A, const void* Wt, int m, int n, int k, const GemmEpi& e) -> int {
    ProfScope ps(h, AACLIP_PROF_GEMM, 2.0 * m * n * k, st);
    ++launches;
    return launch_gemm(epi, bf, A, Wt, m, n, k, e, st);
  };
  auto lnorm = [&](const float* in, const float* g, const float* bta, void* o16, float* o32, int rows, int skipP) -> int {
    ProfScope ps(h, AACLIP_PROF_LAYERNORM, 0.0, st);
    ++launches;
    return launch_layernorm(bf, in, g, bta, o16, o32, rows, W, skipP, true, st);
  };
  auto l2norm = [&](float* t, int rows) -> int {
    ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
    ++launches;
    const int grid = (rows + 7) / 8;
    if (E == 768) l2norm_rows_kernel<768><<<grid, 256, 0, st>>>(t, rows);
    else l2norm_rows_kernel<1024><<<grid, 256, 0, st>>>(t, rows);
    LAUNCH_CHECK("l2norm_rows_kernel");
    return AACLIP_OK;
  };

  // ---- patch embedding + CLS + positional embedding + ln_pre (model/adapter.py:139-156)
  {
    ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
    if (bf) patch_gather_kernel<true><<<dim3(h->G, B), 256, 0, st>>>(images, patchA, c.img_size, h->G, c.patch, h->KP);
    else patch_gather_kernel<false><<<dim3(h->G, B), 256, 0, st>>>(images, patchA, c.img_size, h->G, c.patch, h->KP);
    LAUNCH_CHECK("patch_gather_kernel"); ++launches;
  }
  {
    GemmEpi e; e.out_f32 = x; e.ldo = W; e.pos = h->pos; e.patches = P;
    if ((rc = gemm(EPI_PATCH_F32, patchA, h->w_conv, MP, W, h->KP, e))) return rc;
  }
  {
    ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
    cls_row_kernel<<<B, 256, 0, st>>>(x, h->cls, h->pos, N, W);
    LAUNCH_CHECK("cls_row_kernel"); ++launches;
  }
  if ((rc = lnorm(x, h->lnpre_g, h->lnpre_b, nullptr, x, M, 0))) return rc;

  // ---- residual attention blocks (model/transformer.py:239-258) + adapters (model/adapter.py:161-172)
  int level_idx = 0;
  for (int i = 0; i < c.layers; ++i) {
    const BlockW& b = h->blk[i];
    if ((rc = lnorm(x, b.ln1_g, b.ln1_b, xn, nullptr, M, 0))) return rc;
    { GemmEpi e; e.bias = b.b_qkv; e.out_16 = qkv; e.ldo = 3 * W;
      if ((rc = gemm(EPI_BIAS_BF16, xn, b.w_qkv, M, 3 * W, W, e))) return rc; }
    {
      ProfScope ps(h, AACLIP_PROF_ATTENTION, 4.0 * B * (W / 64) * (double)N * N * 64, st);
      ++launches;
      if ((rc = launch_attention(bf, qkv, att, B, N, W, 0, st))) return rc;
    }
    { GemmEpi e; e.bias = b.b_out; e.out_f32 = x; e.ldo = W;
      if ((rc = gemm(EPI_BIAS_RESID_F32, att, b.w_out, M, W, W, e))) return rc; }
    if ((rc = lnorm(x, b.ln2_g, b.ln2_b, xn, nullptr, M, 0))) return rc;
    { GemmEpi e; e.bias = b.b_fc; e.out_16 = hid; e.ldo = F;
      if ((rc = gemm(c.act == AACLIP_ACT_QUICK_GELU ? EPI_BIAS_QGELU_BF16 : EPI_BIAS_GELU_BF16, xn, b.w_fc, M, F, W, e))) return rc; }
    const bool adapt = i < c.adapt_until;
    { GemmEpi e; e.bias = b.b_proj; e.out_f32 = x; e.ldo = W; e.out_16 = adapt ? xn : nullptr;
      if ((rc = gemm(EPI_BIAS_RESID_F32, hid, b.w_proj, M, W, F, e))) return rc; }
    if (adapt) {
      { GemmEpi e; e.out_f32 = abuf; e.ldo = W;
        if ((rc = gemm(EPI_LEAKY_F32, xn, h->w_adapter[i], M, W, W, e))) return rc; }
      ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
      const int grid = (M + 7) / 8;
      if (W == 1024) adapter_blend_kernel<1024><<<grid, 256, 0, st>>>(x, abuf, M, c.adapt_weight, 1.0f - c.adapt_weight);
      else adapter_blend_kernel<768><<<grid, 256, 0, st>>>(x, abuf, M, c.adapt_weight, 1.0f - c.adapt_weight);
      LAUNCH_CHECK("adapter_blend_kernel");
      ++launches;
    }
    if (level_idx < c.n_levels && c.levels[level_idx] == i + 1) {
      float* tok_f32 = (tokens_out != nullptr) ? tokens_out[level_idx] : nullptr;
      if ((rc = lnorm(x, h->lnpost_g, h->lnpost_b, tok, tok_f32, MP, P))) return rc;
      const int proj_epi = c.proj_relu ? EPI_LEAKY_F32 : EPI_F32;
      { GemmEpi e; e.out_f32 = seg_out[level_idx]; e.ldo = E;
        if ((rc = gemm(proj_epi, tok, h->w_seg[level_idx], MP, E, W, e))) return rc; }
      if ((rc = l2norm(seg_out[level_idx], MP))) return rc;
      if (level_idx == c.n_levels - 1) {
        { GemmEpi e; e.out_f32 = detbuf; e.ldo = E;
          if ((rc = gemm(proj_epi, tok, h->w_det, MP, E, W, e))) return rc; }
        if ((rc = l2norm(detbuf, MP))) return rc;
        ProfScope ps(h, AACLIP_PROF_OTHER, 0.0, st);
        mean_over_patches_kernel<<<dim3((E + 127) / 128, B), 128, 0, st>>>(detbuf, det_out, P, E);
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

int aaclip_last_launch_count(AaclipHandle* h) { return h ? h->last_launches : -1; }

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

int aaclip_score_maps(const float* const* seg, int32_t n_levels, const float* text, int32_t text_per_image,
                      int32_t B, int32_t G, int32_t E, int32_t S, int32_t test_mode, int32_t ksize, float sigma,
                      void* workspace, size_t workspace_bytes, float* maps_out, const float* det, float* score_out,
                      void* stream) {
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
    similarity_kernel<<<grid, 256, 0, st>>>(sp, text, tstride, s, B, n_levels, P, E);
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
    static bool attr_set = false;
    const size_t smem = 3 * (size_t)P * sizeof(float);
    if (smem > 200 * 1024) return fail(AACLIP_EINVAL, "aaclip_score_maps: grid %dx%d too large", G, G);
    if (!attr_set) {
      CUDA_TRY(cudaFuncSetAttribute(anomaly_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    dim3 grid((S + MAP_ROWS - 1) / MAP_ROWS, B);
    anomaly_map_kernel<<<grid, 256, smem, st>>>(s, maps_out, n_levels, G, S, test_mode, ksize, taps);
    LAUNCH_CHECK("anomaly_map_kernel");
  }
  if (det != nullptr && score_out != nullptr) {
    image_score_kernel<<<(B + 7) / 8, 256, 0, st>>>(det, text, tstride, score_out, B, E);
    LAUNCH_CHECK("image_score_kernel");
  }
  return AACLIP_OK;
}

}  // extern "C"

// =========================================================================================== text tower
struct AaclipTextHandle {
  AaclipTextCfg cfg;
  bool bf16 = false;
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  float *tok_emb, *pos, *lnf_g, *lnf_b;
  std::vector<BlockW> blk;
  std::vector<uint16_t*> w_adapter;  // [W,W] each (SimpleAdapter)
  uint16_t* w_proj;                  // [E,W] final SimpleProj
  std::vector<std::string> missing;
};

struct TextWsPlan { size_t x, xn, qkv, att, hid, abuf, eot, eot16, total; };
static TextWsPlan plan_text_ws(const AaclipTextHandle* h, int n) {
  const size_t W = h->cfg.width, F = h->cfg.mlp_dim;
  const size_t M = (size_t)n * h->cfg.context;
  ArenaPlan p;
  TextWsPlan w;
  w.x = p.take<float>(M * W);
  w.xn = p.take<uint16_t>(M * W);
  w.qkv = p.take<uint16_t>(M * 3 * W);
  w.att = p.take<uint16_t>(M * W);
  w.hid = p.take<uint16_t>(M * F);
  w.abuf = p.take<float>(M * W);
  w.eot = p.take<float>((size_t)n * W);
  w.eot16 = p.take<uint16_t>((size_t)n * W);
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
  if (c.operand_dtype != AACLIP_DT_F16 && c.operand_dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "operand_dtype must be AACLIP_DT_F16 or AACLIP_DT_BF16");
  AaclipTextHandle* h = new AaclipTextHandle();
  h->cfg = c;
  h->bf16 = (c.operand_dtype == AACLIP_DT_BF16);
  const size_t W = c.width, F = c.mlp_dim, E = c.embed_dim;
  ArenaPlan plan;
  const size_t o_tok = plan.take<float>((size_t)c.vocab * W), o_pos = plan.take<float>((size_t)c.context * W);
  const size_t o_lg = plan.take<float>(W), o_lb = plan.take<float>(W);
  struct BO { size_t ln1g, ln1b, ln2g, ln2b, bqkv, bout, bfc, bproj, wqkv, wout, wfc, wproj; };
  std::vector<BO> bo(c.layers);
  for (int i = 0; i < c.layers; ++i) {
    bo[i].ln1g = plan.take<float>(W); bo[i].ln1b = plan.take<float>(W);
    bo[i].ln2g = plan.take<float>(W); bo[i].ln2b = plan.take<float>(W);
    bo[i].bqkv = plan.take<float>(3 * W); bo[i].bout = plan.take<float>(W);
    bo[i].bfc = plan.take<float>(F); bo[i].bproj = plan.take<float>(W);
    bo[i].wqkv = plan.take<uint16_t>(3 * W * W); bo[i].wout = plan.take<uint16_t>(W * W);
    bo[i].wfc = plan.take<uint16_t>(F * W); bo[i].wproj = plan.take<uint16_t>(W * F);
  }
  std::vector<size_t> o_ad(c.adapt_until);
  for (int i = 0; i < c.adapt_until; ++i) o_ad[i] = plan.take<uint16_t>(W * W);
  const size_t o_proj = plan.take<uint16_t>(E * W);
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
  for (int i = 0; i < c.layers; ++i) {
    BlockW& b = h->blk[i];
    b.ln1_g = (float*)(a + bo[i].ln1g); b.ln1_b = (float*)(a + bo[i].ln1b);
    b.ln2_g = (float*)(a + bo[i].ln2g); b.ln2_b = (float*)(a + bo[i].ln2b);
    b.b_qkv = (float*)(a + bo[i].bqkv); b.b_out = (float*)(a + bo[i].bout);
    b.b_fc = (float*)(a + bo[i].bfc); b.b_proj = (float*)(a + bo[i].bproj);
    b.w_qkv = (uint16_t*)(a + bo[i].wqkv); b.w_out = (uint16_t*)(a + bo[i].wout);
    b.w_fc = (uint16_t*)(a + bo[i].wfc); b.w_proj = (uint16_t*)(a + bo[i].wproj);
  }
  for (int i = 0; i < c.adapt_until; ++i) h->w_adapter.push_back((uint16_t*)(a + o_ad[i]));
  h->w_proj = (uint16_t*)(a + o_proj);
  auto need = [&](const std::string& k) { h->missing.push_back(k); };
  need("token_embedding.weight"); need("positional_embedding"); need("ln_final.weight"); need("ln_final.bias");
  for (int i = 0; i < c.layers; ++i) {
    const std::string p = "blk." + std::to_string(i) + ".";
    for (const char* s : {"ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias", "attn.in_proj_weight",
                          "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias", "mlp.c_fc.weight",
                          "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias"})
      need(p + s);
  }
  for (int i = 0; i <= c.adapt_until; ++i) need("text_adapter." + std::to_string(i));
  *out = h;
  return AACLIP_OK;
}

void aaclip_text_destroy(AaclipTextHandle* h) {
  if (h == nullptr) return;
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
  auto loaded = [&](const std::string& key) {
    for (size_t i = 0; i < h->missing.size(); ++i)
      if (h->missing[i] == key) { h->missing.erase(h->missing.begin() + i); return; }
  };
  auto copy_f32 = [&](float* dst, int64_t expect, const std::string& key) -> int {
    if (n != expect) return fail(AACLIP_EINVAL, "%s: expected %lld elements, got %lld", name, (long long)expect, (long long)n);
    CUDA_TRY(cudaMemcpyAsync(dst, data, expect * sizeof(float), cudaMemcpyDeviceToDevice, st));
    loaded(key);
    return AACLIP_OK;
  };
  auto pack = [&](uint16_t* dst, int64_t R, int64_t C, const std::string& key) -> int {
    if (n != R * C) return fail(AACLIP_EINVAL, "%s: expected %lld x %lld elements, got %lld", name, (long long)R, (long long)C, (long long)n);
    if (int rc = launch_pack(h->bf16, data, dst, R, C, C, st)) return rc;
    loaded(key);
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
    BlockW& b = h->blk[idx];
    const std::string t(tail), key = "blk." + std::to_string(idx) + "." + t;
    if (t == "ln_1.weight") return copy_f32(b.ln1_g, W, key);
    if (t == "ln_1.bias") return copy_f32(b.ln1_b, W, key);
    if (t == "ln_2.weight") return copy_f32(b.ln2_g, W, key);
    if (t == "ln_2.bias") return copy_f32(b.ln2_b, W, key);
    if (t == "attn.in_proj_bias") return copy_f32(b.b_qkv, 3 * W, key);
    if (t == "attn.out_proj.bias") return copy_f32(b.b_out, W, key);
    if (t == "mlp.c_fc.bias") return copy_f32(b.b_fc, F, key);
    if (t == "mlp.c_proj.bias") return copy_f32(b.b_proj, W, key);
    if (t == "attn.in_proj_weight") return pack(b.w_qkv, 3 * W, W, key);
    if (t == "attn.out_proj.weight") return pack(b.w_out, W, W, key);
    if (t == "mlp.c_fc.weight") return pack(b.w_fc, F, W, key);
    if (t == "mlp.c_proj.weight") return pack(b.w_proj, W, F, key);
    return fail(AACLIP_EINVAL, "%s: unknown block tensor", name);
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
  uint8_t* wsb = static_cast<uint8_t*>(workspace);
  float* x = (float*)(wsb + ws.x);
  uint16_t* xn = (uint16_t*)(wsb + ws.xn);
  uint16_t* qkv = (uint16_t*)(wsb + ws.qkv);
  uint16_t* att = (uint16_t*)(wsb + ws.att);
  uint16_t* hid = (uint16_t*)(wsb + ws.hid);
  float* abuf = (float*)(wsb + ws.abuf);
  float* eot = (float*)(wsb + ws.eot);
  uint16_t* eot16 = (uint16_t*)(wsb + ws.eot16);
  int rc;

  // token + positional embedding (model/adapter.py:277-281)
  text_embed_kernel<<<M, 192, 0, st>>>(tokens, h->tok_emb, h->pos, x, T, W, c.vocab);
  LAUNCH_CHECK("text_embed_kernel");
  // 12 causal residual attention blocks + adapters (model/adapter.py:284-295)
  for (int i = 0; i < c.layers; ++i) {
    const BlockW& b = h->blk[i];
    if ((rc = launch_layernorm(bf, x, b.ln1_g, b.ln1_b, xn, nullptr, M, W, 0, true, st))) return rc;
    { GemmEpi e; e.bias = b.b_qkv; e.out_16 = qkv; e.ldo = 3 * W;
      if ((rc = launch_gemm(EPI_BIAS_BF16, bf, xn, b.w_qkv, M, 3 * W, W, e, st))) return rc; }
    if ((rc = launch_attention(bf, qkv, att, n, T, W, /*causal=*/1, st))) return rc;
    { GemmEpi e; e.bias = b.b_out; e.out_f32 = x; e.ldo = W;
      if ((rc = launch_gemm(EPI_BIAS_RESID_F32, bf, att, b.w_out, M, W, W, e, st))) return rc; }
    if ((rc = launch_layernorm(bf, x, b.ln2_g, b.ln2_b, xn, nullptr, M, W, 0, true, st))) return rc;
    { GemmEpi e; e.bias = b.b_fc; e.out_16 = hid; e.ldo = F;
      if ((rc = launch_gemm(c.act == AACLIP_ACT_QUICK_GELU ? EPI_BIAS_QGELU_BF16 : EPI_BIAS_GELU_BF16, bf, xn, b.w_fc, M, F, W, e, st))) return rc; }
    const bool adapt = i < c.adapt_until;
    { GemmEpi e; e.bias = b.b_proj; e.out_f32 = x; e.ldo = W; e.out_16 = adapt ? xn : nullptr;
      if ((rc = launch_gemm(EPI_BIAS_RESID_F32, bf, hid, b.w_proj, M, W, F, e, st))) return rc; }
    if (adapt) {
      { GemmEpi e; e.out_f32 = abuf; e.ldo = W;
        if ((rc = launch_gemm(EPI_LEAKY_F32, bf, xn, h->w_adapter[i], M, W, W, e, st))) return rc; }
      const int grid = (M + 7) / 8;
      if (W == 1024) adapter_blend_kernel<1024><<<grid, 256, 0, st>>>(x, abuf, M, c.adapt_weight, 1.0f - c.adapt_weight);
      else adapter_blend_kernel<768><<<grid, 256, 0, st>>>(x, abuf, M, c.adapt_weight, 1.0f - c.adapt_weight);
      LAUNCH_CHECK("adapter_blend_kernel");
    }
  }
  // EOT row -> ln_final -> SimpleProj + LeakyReLU (model/adapter.py:296-298)
  text_gather_eot_kernel<<<n, 192, 0, st>>>(tokens, x, eot, T, W);
  LAUNCH_CHECK("text_gather_eot_kernel");
  if ((rc = launch_layernorm(bf, eot, h->lnf_g, h->lnf_b, eot16, nullptr, n, W, 0, true, st))) return rc;
  { GemmEpi e; e.out_f32 = out; e.ldo = E;
    if ((rc = launch_gemm(EPI_LEAKY_F32, bf, eot16, h->w_proj, n, E, W, e, st))) return rc; }
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

int aaclip_gemm16(int32_t dtype, const void* A, const void* W, int32_t M, int32_t N, int32_t K, int32_t epi,
                  const float* bias, float* out_f32, void* out_16, int32_t ldo, void* stream) {
  if (!A || !W) return fail(AACLIP_EINVAL, "aaclip_gemm16: null operand");
  if (dtype != AACLIP_DT_F16 && dtype != AACLIP_DT_BF16) return fail(AACLIP_EINVAL, "aaclip_gemm16: bad dtype %d", dtype);
  if (epi == EPI_PATCH_F32) return fail(AACLIP_EINVAL, "aaclip_gemm16: patch epilogue is internal");
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
