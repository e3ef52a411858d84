// On-device evaluation metrics: pixel / image AUROC and average precision of
// forward_utils.metrics_eval (forward_utils.py:233-308), which the reference computes with scikit-learn on the host
// after copying every map back (test_last.py:151) -- a CPU sort over N_img * 268 324 pixels per class.
//
// Here the maps never leave the GPU: one min/max pass (the global min-max normalisation of forward_utils.py:246-249),
// one pass that bins the normalised scores into two fine histograms (positives / negatives, 2^20 bins) and a scan
// that evaluates the rank statistics from the histograms. Histograms and min/max are ADDITIVE across GPUs: a
// data-parallel run all-reduces 8 MB of counts instead of gathering 1.07 MB per image.
//   AUROC = sum_b neg_b * (pos_above_b + pos_b / 2) / (P * N)                  (Mann-Whitney, ties inside a bin = 1/2)
//   AP    = sum_b (pos_b / P) * TP_b / (TP_b + FP_b), bins swept from the top   (sklearn's step-wise definition)
// With 2^20 bins the discretisation error is <= 2^-21 (pairs that share a bin), far below the 3 decimals the
// reference reports (round(x, 4) * 100).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

namespace aaclip {

constexpr int METRIC_PARTIALS = 1024;

// stage 1 of min/max: grid-stride, one (min, max) pair per block into partials[2 * gridDim.x]
__global__ void minmax_partial_kernel(const float* __restrict__ x, long long n, float* __restrict__ partials) {
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  __shared__ float slo[32], shi[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    lo = threadIdx.x < (blockDim.x >> 5) ? slo[threadIdx.x] : FLT_MAX;
    hi = threadIdx.x < (blockDim.x >> 5) ? shi[threadIdx.x] : -FLT_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = lo; partials[2 * blockIdx.x + 1] = hi; }
  }
}
// stage 2: one block folds the partials; `merge` != 0 also folds the previous content of out (multi-batch streaming)
__global__ void minmax_final_kernel(const float* __restrict__ partials, int n_partials, float* __restrict__ out, int merge) {
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (int i = threadIdx.x; i < n_partials; i += blockDim.x) {
    lo = fminf(lo, partials[2 * i]);
    hi = fmaxf(hi, partials[2 * i + 1]);
  }
  __shared__ float slo[32], shi[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < (blockDim.x >> 5); ++w) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
    if (merge) { lo = fminf(lo, out[0]); hi = fmaxf(hi, out[1]); }
    out[0] = lo;
    out[1] = hi;
  }
}

// hist[label][bin] += 1 with bin = floor(K * (x - lo) / (hi - lo)) (the reference's min-max normalisation, fp32,
// same operation order: (x - min) / (max - min)); label in {0, 1}. normalize == 0 bins x itself (already in [0,1]).
__global__ void score_histogram_kernel(const float* __restrict__ x, const uint8_t* __restrict__ labels, long long n,
                                       const float* __restrict__ minmax, int normalize, int n_bins,
                                       unsigned int* __restrict__ hist) {
  const float lo = normalize ? minmax[0] : 0.f;
  const float range = normalize ? (minmax[1] - minmax[0]) : 1.f;
  const float K = static_cast<float>(n_bins);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = normalize ? (x[i] - lo) / range : x[i];
    int b = static_cast<int>(v * K);
    b = min(max(b, 0), n_bins - 1);
    atomicAdd(&hist[(labels[i] ? n_bins : 0) + b], 1u);
  }
}

// One block of 1024 threads; thread t owns bins [t * per, (t + 1) * per) counted from the TOP of the score range.
// out[0] = AUROC, out[1] = AP, out[2] = #positives, out[3] = #negatives.
__global__ void __launch_bounds__(1024)
auc_ap_kernel(const unsigned int* __restrict__ hist, int n_bins, double* __restrict__ out) {
  const int T = blockDim.x, t = threadIdx.x;
  const int per = (n_bins + T - 1) / T;
  const unsigned int* neg = hist;
  const unsigned int* pos = hist + n_bins;
  // pass 1: segment totals
  unsigned long long sp = 0, sn = 0;
  for (int k = 0; k < per; ++k) {
    const int b = n_bins - 1 - (t * per + k);
    if (b >= 0) { sp += pos[b]; sn += neg[b]; }
  }
  __shared__ unsigned long long cp[1024], cn[1024];
  cp[t] = sp; cn[t] = sn;
  __syncthreads();
  // exclusive prefix over threads (Hillis-Steele, 10 steps)
  for (int o = 1; o < T; o <<= 1) {
    const unsigned long long ap = (t >= o) ? cp[t - o] : 0, an = (t >= o) ? cn[t - o] : 0;
    __syncthreads();
    cp[t] += ap; cn[t] += an;
    __syncthreads();
  }
  const unsigned long long P = cp[T - 1], N = cn[T - 1];
  unsigned long long tp = cp[t] - sp, fp = cn[t] - sn;   // counts strictly above this thread's segment
  // pass 2: rank statistics of the segment
  double auc = 0.0, ap = 0.0;
  for (int k = 0; k < per; ++k) {
    const int b = n_bins - 1 - (t * per + k);
    if (b < 0) break;
    const unsigned long long pb = pos[b], nb = neg[b];
    auc += static_cast<double>(nb) * (static_cast<double>(tp) + 0.5 * static_cast<double>(pb));
    tp += pb; fp += nb;
    if (pb) ap += static_cast<double>(pb) * static_cast<double>(tp) / static_cast<double>(tp + fp);
  }
  __shared__ double sa[1024], sq[1024];
  sa[t] = auc; sq[t] = ap;
  __syncthreads();
  for (int o = T / 2; o > 0; o >>= 1) {
    if (t < o) { sa[t] += sa[t + o]; sq[t] += sq[t + o]; }
    __syncthreads();
  }
  if (t == 0) {
    out[0] = (P && N) ? sa[0] / (static_cast<double>(P) * static_cast<double>(N)) : 0.0;
    out[1] = P ? sq[0] / static_cast<double>(P) : 0.0;
    out[2] = static_cast<double>(P);
    out[3] = static_cast<double>(N);
  }
}

// per-image maximum of a [B, hw] array (pmax of forward_utils.py:275), one block per image
__global__ void row_max_kernel(const float* __restrict__ x, long long hw, float* __restrict__ out) {
  const float* r = x + static_cast<long long>(blockIdx.x) * hw;
  float hi = -FLT_MAX;
  for (long long i = threadIdx.x; i < hw; i += blockDim.x) hi = fmaxf(hi, r[i]);
  __shared__ float sh[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = hi;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < (blockDim.x >> 5); ++w) hi = fmaxf(hi, sh[w]);
    out[blockIdx.x] = hi;
  }
}

}  // namespace aaclip
