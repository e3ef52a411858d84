/*
 * aaclip.h -- C ABI of the B200-native AA-CLIP inference hot path (libaaclip_b200.so, sm_100a).
 *
 * The reference (liu20050510/AA-CLIP-IQM) has no FFI/plugin interface: its hot path sits behind plain Python
 * methods. Each entry point below names the reference code it replaces (file:line relative to the reference
 * root). The only caller is this repo's Python mirror of that surface (aa-clip-iqm_b200/model/adapter.py,
 * aa-clip-iqm_b200/forward_utils.py); INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every data pointer is CALLER-OWNED DEVICE memory (e.g. torch-allocated)
 *    unless the parameter name ends in _host;
 *  - all work is enqueued asynchronously on the `stream` argument (a cudaStream_t passed as void*); no hidden
 *    synchronisation, no allocation after aaclip_create(), so calls may be captured in a CUDA graph;
 *  - return 0 on success, a negative AACLIP_E* code on failure; never throws across the ABI;
 *    aaclip_last_error() returns a thread-local message for the last failure;
 *  - a handle is bound to the device that was current at aaclip_create() and is not thread-safe.
 */
#ifndef AACLIP_H_
#define AACLIP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AACLIP_OK 0
#define AACLIP_EINVAL (-1)   /* bad argument / unsupported shape */
#define AACLIP_ECUDA (-2)    /* CUDA runtime / driver error */
#define AACLIP_ESTATE (-3)   /* weights missing, wrong call order */
#define AACLIP_ENOSYS (-4)   /* not an sm_100 device / driver entry point missing */

#define AACLIP_MAX_LEVELS 4

/* activation of the MLP: the reference runs nn.GELU (erf) -- model/model.py:84; QuickGELU only with
 * force_quick_gelu=True -- model/clip.py:151-153, model/transformer.py:46-49 */
#define AACLIP_ACT_GELU_ERF 0
#define AACLIP_ACT_QUICK_GELU 1

/* 16-bit tensor-core operand type (fp32 accumulation in TMEM either way; same tensor throughput).
 * F16 is the default: on the reference's fp32 oracle it keeps anomaly maps within 1e-2 abs, which bf16 operands
 * (3 fewer mantissa bits) do not (DESIGN.md "Precision"). Stores saturate to +-65504 in F16 mode. */
#define AACLIP_DT_F16 0
#define AACLIP_DT_BF16 1
/* fp32-accuracy mode (handles only): the same launch sequence with fp32 SIMT GEMM / attention kernels and un-folded
 * LayerNorms; reproduces the reference's fp32 maps to <= 1e-4 (BASELINE config 4). A verification mode, ~30x slower. */
#define AACLIP_DT_F32 2

/* epilogue selectors of aaclip_gemm16 (unit-test / building-block entry point); "BF16" in these names means
 * "the 16-bit operand dtype" */
#define AACLIP_EPI_BIAS_BF16 0
#define AACLIP_EPI_BIAS_GELU_BF16 1
#define AACLIP_EPI_BIAS_QGELU_BF16 2
#define AACLIP_EPI_BIAS_RESID_F32 3
#define AACLIP_EPI_F32 4
#define AACLIP_EPI_LEAKY_F32 5
#define AACLIP_EPI_BIAS_F32 7
/* folded-LayerNorm epilogues of aaclip_gemm16_ln (ln_1 / ln_2 of model/transformer.py:245,252 live in the GEMMs) */
#define AACLIP_EPI_LN_BIAS_BF16 8
#define AACLIP_EPI_LN_BIAS_GELU_BF16 9
#define AACLIP_EPI_LN_BIAS_QGELU_BF16 10

/* Static description of the adapted visual tower (model/model_configs/ViT-L-14-336.json +
 * AdaptedCLIP.__init__, model/adapter.py:11-47). */
typedef struct AaclipCfg {
  int32_t img_size;      /* 518 (any multiple of patch) */
  int32_t patch;         /* 14 */
  int32_t width;         /* 1024 */
  int32_t heads;         /* 16 (head dim must be 64) */
  int32_t layers;        /* 24 */
  int32_t mlp_dim;       /* 4096 */
  int32_t embed_dim;     /* 768 */
  int32_t adapt_until;   /* residual adapters after blocks 0..adapt_until-1 (6) */
  int32_t n_levels;      /* 4 */
  int32_t levels[AACLIP_MAX_LEVELS]; /* 1-based tapped blocks: 6,12,18,24 (ascending; last one feeds det_proj) */
  int32_t act;           /* AACLIP_ACT_* */
  int32_t proj_relu;     /* SimpleProj(relu=...) : LeakyReLU after seg/det projections */
  float adapt_weight;    /* i_w = 0.1 */
  int32_t max_batch;     /* largest B accepted by aaclip_forward */
  int32_t operand_dtype; /* AACLIP_DT_F16 (default), AACLIP_DT_BF16, or AACLIP_DT_F32 (fp32-accuracy mode) */
  int32_t query_dim;     /* IQM branch: width of query_adapters (model/adapter.py:75-78: iqm_hidden_size, a multiple
                          * of 256); 0 = the tower carries no query adapters */
} AaclipCfg;

typedef struct AaclipHandle AaclipHandle;

/* Library / build info: "aaclip_b200 <version> sm_100a". */
const char* aaclip_version(void);
/* Thread-local description of the last error returned on this thread. */
const char* aaclip_last_error(void);

/* Allocates the packed-weight arena (bf16 GEMM operands, fp32 norms/biases) on the current device.
 * Replaces: building the nn.Module tree in create_model (model/clip.py:84-202) + AdaptedCLIP.__init__. */
int aaclip_create(const AaclipCfg* cfg, AaclipHandle** out);
void aaclip_destroy(AaclipHandle* h);

/* Hands one fp32 DEVICE tensor to the engine, which copies and repacks it (caller keeps ownership).
 * `name` is the reference state_dict key relative to AdaptedCLIP, e.g.
 *   "image_encoder.conv1.weight" [W,3,p,p]      "image_encoder.class_embedding" [W]
 *   "image_encoder.positional_embedding" [N,W]  "image_encoder.ln_pre.weight" / ".bias" / "ln_post.*"
 *   "image_encoder.transformer.resblocks.<i>.{ln_1,ln_2}.{weight,bias}"
 *   "image_encoder.transformer.resblocks.<i>.attn.{in_proj_weight,in_proj_bias}", ".attn.out_proj.{weight,bias}"
 *   "image_encoder.transformer.resblocks.<i>.mlp.{c_fc,c_proj}.{weight,bias}"
 *   "image_adapter.layer_adapters.<i>.fc.0.weight", "image_adapter.seg_proj.<l>.fc.weight" (or ".fc.0.weight"),
 *   "image_adapter.det_proj.fc.weight" (or ".fc.0.weight")
 *   "query_adapters.<l>.fc.weight" (or ".fc.0.weight") [query_dim, W] -- OPTIONAL (never counted as missing): the
 *   IQM branch's per-level SimpleProj (model/adapter.py:75-78), needed only when aaclip_forward is asked for query_out
 * Replaces: load_state_dict (model/clip.py:132) / checkpoint loading (test_last.py:231-247). */
int aaclip_set_weight(AaclipHandle* h, const char* name, const float* data, const int64_t* shape, int32_t ndim,
                      void* stream);
/* Number of tensors still missing before aaclip_forward may run (0 = ready). */
int aaclip_missing_weights(AaclipHandle* h);

/* Bytes of scratch aaclip_forward needs for a batch of B images (256-byte aligned pointer expected). */
size_t aaclip_workspace_bytes(AaclipHandle* h, int32_t B);

/* Adapted visual tower. Replaces AdaptedCLIP.forward lines model/adapter.py:137-184:
 * patch embed, CLS + positional embedding, ln_pre, 24 pre-LN residual attention blocks
 * (model/transformer.py:239-258) with residual adapters on the first `adapt_until`, token taps + ln_post,
 * seg/det projections, L2 normalisation, patch-mean of the det token.
 *   images      [B,3,S,S] fp32 NCHW
 *   seg_out[l]  [B,P,E] fp32, unit-norm rows            (l < n_levels)
 *   det_out     [B,E]   fp32
 *   tokens_out  optional (NULL or array of n_levels pointers, each NULL or [B,P,W] fp32): ln_post(tokens[l]),
 *               the input of the reference's IQM branch (model/adapter.py:202-205)
 *   cls_out     optional [B,W] fp32: x[:,0,:] after the last block, no ln_post (model/adapter.py:190)
 *   query_out   optional (NULL or array of n_levels pointers, each NULL or [B,P,query_dim] fp32):
 *               query_adapters[l](ln_post(tokens[l])) -- model/adapter.py:202-205, the 4 x (P x W x query_dim) GEMMs of
 *               the IQM branch run on the tensor cores from the 16-bit ln_post tokens the tower already holds
 *               (SURVEY.md section 8f rank 2); the host side then never needs tokens_out */
int aaclip_forward(AaclipHandle* h, const float* images, int32_t B, void* workspace, size_t workspace_bytes,
                   float* const* seg_out, float* det_out, float* const* tokens_out, float* cls_out,
                   float* const* query_out, void* stream);

/* Bytes of scratch aaclip_score_maps needs. */
size_t aaclip_score_workspace_bytes(int32_t B, int32_t n_levels, int32_t P);

/* Patch-text similarity maps. Replaces forward_utils.calculate_similarity_map (forward_utils.py:196-216), the
 * level sum of test_last.py:143,149 and the image score of test_last.py:90-92 (well-posed form).
 *   seg[l]      [B,P,E] fp32 (n_levels pointers; P = G*G)
 *   text        [E,2] (text_per_image = 0) or [B,E,2] (text_per_image = 1) fp32; column 0 normal, 1 abnormal
 *   test_mode=1: maps_out [B,S,S] = sum_l upsample(blur((s1 + 1 - s0)/2))      ksize/sigma: 7/1.0 Industrial,
 *                9/1.5 Medical (forward_utils.py:205-206); ksize odd <= 15
 *   test_mode=0: n_levels must be 1; maps_out [B,2,S,S] = softmax_c(upsample(s_c))  (training-time form)
 *   det/score_out optional: score_out[b] = (det[b,:] . text[:,1] + 1) / 2 */
int aaclip_score_maps(const float* const* seg, int32_t n_levels, const float* text, int32_t text_per_image,
                      int32_t B, int32_t G, int32_t E, int32_t S, int32_t test_mode, int32_t ksize, float sigma,
                      void* workspace, size_t workspace_bytes, float* maps_out, const float* det, float* score_out,
                      void* stream);

/* Columns covered by one LayerNorm partial written by aaclip_gemm16_resid (= accumulator columns per epilogue warp). */
/* The same call for seg tensors that are NOT unit-norm (a tower with aaclip_seg_normalize(h, 0)): the similarity
 * kernel applies F.normalize(dim=-1, eps=1e-12) (model/adapter.py:182) itself, in the one pass it makes over the
 * features, instead of a separate normalisation pass that reads and re-writes all of them. */
int aaclip_score_maps_raw(const float* const* seg, int32_t n_levels, const float* text, int32_t text_per_image,
                          int32_t B, int32_t G, int32_t E, int32_t S, int32_t test_mode, int32_t ksize, float sigma,
                          void* workspace, size_t workspace_bytes, float* maps_out, const float* det, float* score_out,
                          void* stream);

/* enable = 0: aaclip_forward leaves seg_out as the raw projections tokens W^T (no F.normalize); det_out is unchanged.
 * For callers that hand seg_out straight to aaclip_score_maps_raw. Default 1 = the reference's return values. */
int aaclip_seg_normalize(AaclipHandle* h, int32_t enable);

int aaclip_stat_columns(void);

/* Kernel launches enqueued by the last aaclip_forward on this handle (for bench.py's gpu_launches). */
int aaclip_last_launch_count(AaclipHandle* h);

/* Load-time finalisation: folds ln_1 / ln_2 (model/transformer.py:245,252) into the in_proj / c_fc GEMM weights
 * (kernels on `stream`, one stream synchronisation, release of the fp32 staging copies). Call it after the last
 * aaclip_set_weight(); aaclip_forward() then allocates nothing, never blocks the host, and can be stream-captured.
 * If it was not called, the first aaclip_forward() does it. Idempotent; AACLIP_ESTATE while tensors are missing.
 * After it the LayerNorm / in_proj / c_fc tensors of the blocks are frozen (aaclip_set_weight refuses them). */
int aaclip_finalize_weights(AaclipHandle* h, void* stream);

/* Run-time validation of the 16-bit operand range (there is no such check in the fp32 reference: this guards the
 * one thing the tensor-core path adds). While enabled, aaclip_forward follows every kernel that stores 16-bit
 * activations (patch operand, the raw residual copy x16, q|k|v, attention output, GELU output, ln_post tokens) with a
 * scan that counts stored elements sitting at the format's ceiling: fp16 +-65504 -- what the saturating stores write
 * for anything larger -- or non-finite; bf16: non-finite. A count of 0 proves that no activation of the run was clipped.
 * Off by default (the scans re-read every activation: ~1.5x the step time). aaclip_saturation_count copies the
 * accumulated count to the host (synchronises `stream`) and optionally resets it. */
int aaclip_saturation_check(AaclipHandle* h, int32_t enable);
int aaclip_saturation_count(AaclipHandle* h, uint64_t* count_host, int32_t reset, void* stream);

/* Per-kernel-class device timing of aaclip_forward, for bench.py's roofline: while enabled, every launch is bracketed
 * by a CUDA-event pair on the launch stream. aaclip_profile_read waits for the recorded events, returns the
 * accumulated milliseconds / launch counts / algorithmic FLOPs (2*M*N*K per GEMM, 4*B*H*N*N*64 per attention) per
 * class since the last enable/read, and clears the records. Arrays have AACLIP_PROF_CLASSES entries. */
#define AACLIP_PROF_GEMM 0        /* GEMMs outside the blocks: patch embed, adapters, seg / det projections */
#define AACLIP_PROF_ATTENTION 1
#define AACLIP_PROF_LAYERNORM 2
#define AACLIP_PROF_OTHER 3
#define AACLIP_PROF_GEMM_QKV 4    /* the four GEMMs of a residual attention block, timed separately */
#define AACLIP_PROF_GEMM_OUT 5
#define AACLIP_PROF_GEMM_FC 6
#define AACLIP_PROF_GEMM_PROJ 7
#define AACLIP_PROF_CLASSES 8
int aaclip_profile_enable(AaclipHandle* h, int32_t enable);
int aaclip_profile_read(AaclipHandle* h, float* ms, int32_t* launches, double* flops);

/* ---- text anchors (BASELINE config 3) --------------------------------------------------------------------- */

/* Static description of the adapted CLIP text tower (model/model_configs/ViT-L-14-336.json "text_cfg" +
 * AdaptedCLIP.__init__ text_adapter, model/adapter.py:51-54). */
typedef struct AaclipTextCfg {
  int32_t context;       /* 77 */
  int32_t vocab;         /* 49408 */
  int32_t width;         /* 768 (head dim must be 64) */
  int32_t heads;         /* 12 */
  int32_t layers;        /* 12 */
  int32_t mlp_dim;       /* 3072 */
  int32_t embed_dim;     /* 768: output of text_adapter[-1] = SimpleProj(768, 768, relu=True) */
  int32_t adapt_until;   /* residual adapters after blocks 0..adapt_until-1 (3) */
  int32_t act;           /* AACLIP_ACT_* */
  float adapt_weight;    /* t_w = 0.1 */
  int32_t max_sentences; /* largest n accepted by aaclip_encode_text */
  int32_t operand_dtype; /* AACLIP_DT_F16 (default) or AACLIP_DT_BF16 */
} AaclipTextCfg;

typedef struct AaclipTextHandle AaclipTextHandle;

/* Replaces: building the CLIP text tower (model/model.py:119-172) + text_adapter (model/adapter.py:51-54). */
int aaclip_text_create(const AaclipTextCfg* cfg, AaclipTextHandle** out);
void aaclip_text_destroy(AaclipTextHandle* h);

/* One fp32 DEVICE tensor of the text path; `name` is the AdaptedCLIP state_dict key:
 *   "clipmodel.token_embedding.weight" [vocab,W]   "clipmodel.positional_embedding" [ctx,W]
 *   "clipmodel.transformer.resblocks.<i>.{ln_1,ln_2}.{weight,bias}", ".attn.{in_proj_weight,in_proj_bias}",
 *   ".attn.out_proj.{weight,bias}", ".mlp.{c_fc,c_proj}.{weight,bias}"   "clipmodel.ln_final.{weight,bias}"
 *   "text_adapter.<i>.fc.0.weight" (i < adapt_until: SimpleAdapter; i == adapt_until: the final SimpleProj) */
int aaclip_text_set_weight(AaclipTextHandle* h, const char* name, const float* data, const int64_t* shape,
                           int32_t ndim, void* stream);
int aaclip_text_missing_weights(AaclipTextHandle* h);
size_t aaclip_text_workspace_bytes(AaclipTextHandle* h, int32_t n_sentences);
/* Text-tower counterparts of aaclip_finalize_weights / aaclip_saturation_check / aaclip_saturation_count. */
int aaclip_text_finalize_weights(AaclipTextHandle* h, void* stream);
int aaclip_text_saturation_check(AaclipTextHandle* h, int32_t enable);
int aaclip_text_saturation_count(AaclipTextHandle* h, uint64_t* count_host, int32_t reset, void* stream);

/* Adapted text encoder. Replaces AdaptedCLIP.encode_text(text, adapt_text=True), model/adapter.py:273-304:
 * token + positional embedding, 12 causal pre-LN residual attention blocks with residual adapters on the first
 * `adapt_until`, ln_final, EOT-row gather (argmax of the token ids), SimpleProj + LeakyReLU.
 *   tokens [n,ctx] int32 device; ids outside [0,vocab) are clamped by the kernel (the Python wrapper validates)
 *   out    [n,E] fp32 (NOT normalised, exactly what encode_text returns) */
int aaclip_encode_text(AaclipTextHandle* h, const int32_t* tokens, int32_t n_sentences, void* workspace,
                       size_t workspace_bytes, float* out, void* stream);

/* Prompt ensembling of forward_utils.get_adapted_single_class_text_embedding (forward_utils.py:152-161):
 * group g = rows offsets[g]..offsets[g+1]-1 of emb; each row is L2-normalised, the group is averaged and the mean
 * re-normalised. Groups 2c and 2c+1 are the normal / abnormal states of class c:
 *   anchors [n_groups/2, E, 2] fp32  (== torch.stack([normal, abnormal], dim=1) per class)
 *   offsets [n_groups+1] int32 DEVICE */
int aaclip_text_anchors(const float* emb, const int32_t* offsets, int32_t n_groups, int32_t E, float* anchors,
                        void* stream);

/* ---- image pre-processing on the device -------------------------------------------------- */
/* Replaces the per-image CPU transform of the reference's datasets (dataset/__init__.py:150-159):
 *   transforms.Resize((S, S), Image.BICUBIC) -> transforms.ToTensor() -> transforms.Normalize(mean, std)
 * for a batch of decoded RGB images of one size, BIT-EXACT with PIL + torchvision: Pillow's 8-bit fixed-point
 * antialiased bicubic resampling (horizontal pass, uint8, then vertical pass), float(u8) / 255, (x - mean) / std.
 *
 * The tap tables depend only on (input size, S) and are computed on the HOST in double precision, exactly like
 * Resample.c's precompute_coeffs + normalize_coeffs_8bpc; the caller uploads them once per image size:
 *   bounds [S, 2] int32 = (first tap, tap count) per output coordinate;  kk [S, ksize] int32 fixed-point (22 bit) taps */
int32_t aaclip_resample_ksize(int32_t in_size, int32_t out_size);
int aaclip_resample_coeffs_host(int32_t in_size, int32_t out_size, int32_t* bounds_host, int32_t* kk_host,
                                int32_t kk_capacity /* entries available in kk_host, >= S * ksize */);
size_t aaclip_preprocess_workspace_bytes(int32_t B, int32_t H, int32_t S);
/*   src [B, H, W, 3] uint8 (HWC, RGB) DEVICE;  xbounds / xkk: tables of (W -> S), ybounds / ykk: tables of (H -> S),
 *   DEVICE;  mean_host / std_host: 3 floats each, HOST;  out [B, 3, S, S] fp32 DEVICE (the layout aaclip_forward takes) */
int aaclip_preprocess_u8(const uint8_t* src, int32_t B, int32_t H, int32_t W, int32_t S, const int32_t* xbounds,
                         const int32_t* xkk, const int32_t* ybounds, const int32_t* ykk, const float* mean_host,
                         const float* std_host, void* workspace, size_t workspace_bytes, float* out, void* stream);

/* ---- on-device evaluation metrics ------------------------------------------------------ */
/* Replace the scikit-learn calls of forward_utils.metrics_eval (forward_utils.py:246-249,275,288-296): global min-max
 * normalisation, per-image maximum, AUROC and average precision, computed from two fine score histograms (positives,
 * negatives). min/max and histograms are additive across ranks: a data-parallel evaluation all-reduces them instead
 * of gathering the maps. */
size_t aaclip_metrics_workspace_bytes(void);
/* minmax[0..1] = (min, max) of x[0..n); merge != 0 also folds the previous content of minmax (streaming batches). */
int aaclip_minmax(const float* x, int64_t n, void* workspace, float* minmax, int32_t merge, void* stream);
/* hist[label * n_bins + bin] += 1, bin = floor(n_bins * (x - min) / (max - min)) (minmax == NULL: x is already in
 * [0,1]); labels uint8 in {0,1}; hist uint32 [2][n_bins], zeroed by the caller before the first batch. */
int aaclip_score_histogram(const float* x, const uint8_t* labels, int64_t n, const float* minmax, int32_t n_bins,
                           uint32_t* hist, void* stream);
/* out[0] = AUROC, out[1] = average precision, out[2] = #positives, out[3] = #negatives (doubles, device). */
int aaclip_auc_ap(const uint32_t* hist, int32_t n_bins, double* out, void* stream);
/* out[r] = max_c x[r, c]: the per-image maximum of the normalised maps (pmax, forward_utils.py:275). */
int aaclip_row_max(const float* x, int32_t rows, int64_t cols, float* out, void* stream);

/* ---- building blocks, exported for the parity tests (each is exactly what aaclip_forward launches) ---- */

/* C[M,N] = A[M,K] * W[N,K]^T with a fused epilogue (tcgen05/TMEM/TMA GEMM). A, W: `dtype` row-major.
 * N % 256 == 0, K % 64 == 0. out_f32 / out_16 as required by `epi`; ldo = leading dim of outputs. */
int aaclip_gemm16(int32_t dtype, const void* A, const void* W, int32_t M, int32_t N, int32_t K, int32_t epi,
                  const float* bias, float* out_f32, void* out_16, int32_t ldo, void* stream);

/* LayerNorm folding at weight-load time (what the first aaclip_forward does for every block):
 *   W16[n,k] = W[n,k]*gamma[k] (rounded to `dtype`), colsum[n] = sum_k W16[n,k], bias_f[n] = b[n] + sum_k W[n,k]*beta[k],
 * so that  LayerNorm(x) W^T + b  =  rstd * (x W16^T - mean * colsum) + bias_f  with x the RAW residual row. */
int aaclip_ln_fold(int32_t dtype, const float* W, const float* b, const float* gamma, const float* beta, int32_t N,
                   int32_t K, void* W16, float* colsum, float* bias_f, void* stream);

/* GEMM with the LayerNorm of the A rows folded into the epilogue (epi = AACLIP_EPI_LN_*):
 *   out_16 = act(rstd_m * (A W16^T - mean_m * colsum) + bias_f),
 * mean / rstd (eps 1e-5, width ln_width) finished from stats [M][n_partials][2] = per-row partial (sum, sum of squares)
 * over disjoint column ranges, n_partials <= 16 (written by aaclip_gemm16_resid / the ln_pre and adapter-blend kernels). */
int aaclip_gemm16_ln(int32_t dtype, const void* A, const void* W16, int32_t M, int32_t N, int32_t K, int32_t epi,
                     const float* bias_f, const float* colsum, const float* stats, int32_t n_partials,
                     int32_t ln_width, void* out_16, int32_t ldo, void* stream);

/* Token-tap projection with ln_post folded in (model/adapter.py:171-184), as aaclip_forward launches it for seg_proj /
 * det_proj / query_adapters: A = the raw 16-bit residual stream with ALL patches+1 rows of every image (M = B*(patches+1));
 * the rows of the CLS tokens are dropped and row b*(patches+1)+t lands at out_f32[b*patches + t - 1]:
 *   out_f32 = act(rstd_m * (A W16^T - mean_m * colsum) + bias_f),  act = LeakyReLU(0.01) if `leaky` else identity. */
int aaclip_gemm16_ln_tok(int32_t dtype, const void* A, const void* W16, int32_t M, int32_t N, int32_t K, int32_t leaky,
                         const float* bias_f, const float* colsum, const float* stats, int32_t n_partials,
                         int32_t ln_width, int32_t patches, float* out_f32, int32_t ldo, void* stream);

/* Residual GEMM on an fp32 stream: x[M,N] (fp32, in place) += A W^T + bias; optionally also the raw 16-bit copy of the
 * new x and its LayerNorm partials stats_out [M][N/aaclip_stat_columns()][2]. (Rounds 1-2a; aaclip_forward now uses the
 * hi / lo form, aaclip_gemm16_resid_hl. Kept as a building block.) */
int aaclip_gemm16_resid(int32_t dtype, const void* A, const void* W, int32_t M, int32_t N, int32_t K,
                        const float* bias, float* x, void* x16_out, float* stats_out, int32_t ldo, void* stream);

/* The residual stream of the 16-bit modes is kept as two 16-bit planes: hi = the value rounded to `dtype` (which is
 * also the A operand of the GEMMs that consume the stream) and lo = the rounded remainder, x = hi + lo to ~22 bits
 * (fp16) -- 8 instead of 10 bytes per element and update against an fp32 stream with a separate 16-bit copy.
 * aaclip_split_hilo / aaclip_join_hilo convert n elements (n % 4 == 0) between fp32 and that form. */
int aaclip_split_hilo(int32_t dtype, const float* x, void* x_hi, void* x_lo, int64_t n, void* stream);
int aaclip_join_hilo(int32_t dtype, const void* x_hi, const void* x_lo, float* x, int64_t n, void* stream);

/* Residual GEMM on that form, exactly what aaclip_forward launches for out_proj / c_proj (model/transformer.py:254-257):
 * x[M,N] (hi / lo planes, in place) += A W^T + bias; stats_out [M][N/aaclip_stat_columns()][2] = LayerNorm partials of
 * the new rows (the inputs of the next folded-LN GEMM), optional. */
int aaclip_gemm16_resid_hl(int32_t dtype, const void* A, const void* W, int32_t M, int32_t N, int32_t K,
                           const float* bias, void* x_hi, void* x_lo, float* stats_out, int32_t ldo, void* stream);

/* Residual adapter + blend of one block (model/adapter.py:163-170, model/adapter_modules.py:6-13), exactly the two
 * launches aaclip_forward issues:  a = LeakyReLU_0.01(x W^T);  a <- a |x|_2 / |a|_2 per row;  x <- weight a + (1-weight) x.
 *   x_hi / x_lo [M,D] `dtype`: the residual stream, in place (x_hi is the GEMM operand);
 *   W16 [D,D] `dtype` (aaclip_pack16 of fc.0.weight);  stats_out [M][D/aaclip_stat_columns()][2] LayerNorm partials of
 *   the new x;  workspace >= M*D*2 + M*(D/aaclip_stat_columns())*8 bytes.  D in {768, 1024}. */
int aaclip_adapter(int32_t dtype, void* x_hi, void* x_lo, const void* W16, int32_t M, int32_t D, float weight,
                   void* workspace, size_t workspace_bytes, float* stats_out, void* stream);

/* Fused multi-head attention: qkv `dtype` [B,N,3*D] -> out `dtype` [B,N,D]; head dim 64; softmax scale 1/8. */
int aaclip_attention(int32_t dtype, const void* qkv, void* out, int32_t B, int32_t N, int32_t D, int32_t causal,
                     void* stream);

/* LayerNorm(eps=1e-5) rows of fp32 x [rows_in, D] (D in {768,1024}); out_16 and/or out_f32 may be NULL.
 * skip_cls_P > 0 drops the first row of every (P+1)-row group (x[1:] of each image). */
int aaclip_layernorm(int32_t dtype, const float* x, const float* gamma, const float* beta, void* out_16,
                     float* out_f32, int32_t rows_out, int32_t D, int32_t skip_cls_P, void* stream);

/* fp32 [R,C] -> `dtype` [R,Cp] (Cp >= C, zero padded): the weight / operand packing used by aaclip_set_weight. */
int aaclip_pack16(int32_t dtype, const float* src, void* dst_16, int64_t R, int64_t C, int64_t Cp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AACLIP_H_ */
