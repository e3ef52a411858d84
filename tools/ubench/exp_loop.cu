// Micro-benchmark: the attention softmax inner loop in isolation (TMEM row load -> P = exp2(S*scale - m) -> packed
// 16-bit -> TMEM store), clk per 128-element row per warp, for 1 and 2 warps per SM sub-partition and several
// formulations of the loop. nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_loop exp_loop.cu && ./exp_loop
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint64_t pk(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) { uint32_t r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

#define LD32(r, addr)                                                                                               \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                            \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                             \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"             \
               : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]), "=r"((r)[6]), "=r"((r)[7]),     \
                 "=r"((r)[8]), "=r"((r)[9]), "=r"((r)[10]), "=r"((r)[11]), "=r"((r)[12]), "=r"((r)[13]), "=r"((r)[14]),            \
                 "=r"((r)[15]), "=r"((r)[16]), "=r"((r)[17]), "=r"((r)[18]), "=r"((r)[19]), "=r"((r)[20]), "=r"((r)[21]),          \
                 "=r"((r)[22]), "=r"((r)[23]), "=r"((r)[24]), "=r"((r)[25]), "=r"((r)[26]), "=r"((r)[27]), "=r"((r)[28]),          \
                 "=r"((r)[29]), "=r"((r)[30]), "=r"((r)[31])                                                               \
               : "r"(addr)                                                                                           \
               : "memory")
#define ST32(r, addr)                                                                                               \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                      \
               "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                            \
               "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(addr),       \
               "r"((r)[0]), "r"((r)[1]), "r"((r)[2]), "r"((r)[3]), "r"((r)[4]), "r"((r)[5]), "r"((r)[6]), "r"((r)[7]), "r"((r)[8]),    \
               "r"((r)[9]), "r"((r)[10]), "r"((r)[11]), "r"((r)[12]), "r"((r)[13]), "r"((r)[14]), "r"((r)[15]), "r"((r)[16]),        \
               "r"((r)[17]), "r"((r)[18]), "r"((r)[19]), "r"((r)[20]), "r"((r)[21]), "r"((r)[22]), "r"((r)[23]), "r"((r)[24]),       \
               "r"((r)[25]), "r"((r)[26]), "r"((r)[27]), "r"((r)[28]), "r"((r)[29]), "r"((r)[30]), "r"((r)[31])                    \
               : "memory")

// degree-3 polynomial exp2 on the FMA pipe for a packed pair (Cody-Waite: x = n + f, f in [-0.5, 0.5])
__device__ __forceinline__ void poly_exp2_pair(float x0, float x1, float& p0, float& p1) {
  const float magic = 12582912.f;  // 1.5 * 2^23
  x0 = fmaxf(x0, -126.f); x1 = fmaxf(x1, -126.f);
  uint64_t x = pk(x0, x1);
  uint64_t n = add2(x, pk(magic, magic));                 // round to nearest integer in the low mantissa bits
  uint64_t nf = add2(n, pk(-magic, -magic));
  uint64_t f = add2(x, nf ^ 0x8000000080000000ull);       // f = x - n
  uint64_t r = fma2(f, pk(0.0555041f, 0.0555041f), pk(0.2402265f, 0.2402265f));
  r = fma2(r, f, pk(0.6931472f, 0.6931472f));
  r = fma2(r, f, pk(1.0f, 1.0f));
  float r0, r1, n0, n1;
  upk(r, r0, r1); upk(n, n0, n1);
  p0 = __uint_as_float(__float_as_uint(r0) + (__float_as_uint(n0) << 23));
  p1 = __uint_as_float(__float_as_uint(r1) + (__float_as_uint(n1) << 23));
}

// MODE 0: scalar FFMA / MUFU / FADD / F2FP    1: FFMA2 + FADD2 (current kernel)   2: MODE 1 without the row sum
// MODE 3: f16x2 MUFU (FFMA2, F2FP, MUFU.f16x2; no row sum)   4: MODE 1 with every 4th pair on the FMA-pipe polynomial
// MODE 5: MODE 2 (no row sum) with every 4th pair polynomial   6: every 2nd pair polynomial, no row sum
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float* out, int iters, long long* cycles, float scale, float m) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 256;
  uint32_t s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = __float_as_uint(-0.01f * (float)((threadIdx.x + i) & 255));
  ST32(&s[0], t); ST32(&s[32], t + 32); ST32(&s[64], t + 64); ST32(&s[96], t + 96);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  float l = 0.f;
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
    LD32(&s[0], t); LD32(&s[32], t + 32); LD32(&s[64], t + 64); LD32(&s[96], t + 96);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const uint64_t scale2 = pk(scale, scale), negm2 = pk(-m, -m);
    uint64_t rs2[2] = {0ull, 0ull};
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t w[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int kk = c * 64 + 2 * i;
        if (MODE == 0) {
          const float p0 = ex2(fmaf(__uint_as_float(s[kk]), scale, -m));
          const float p1 = ex2(fmaf(__uint_as_float(s[kk + 1]), scale, -m));
          rs[(2 * i) & 3] += p0; rs[(2 * i + 1) & 3] += p1;
          w[i] = pack_h2(p0, p1);
        } else if (MODE == 3) {
          float x0, x1;
          upk(fma2(pk(__uint_as_float(s[kk]), __uint_as_float(s[kk + 1])), scale2, negm2), x0, x1);
          w[i] = ex2_h2(pack_h2(x0, x1));
        } else {
          float x0, x1, p0, p1;
          upk(fma2(pk(__uint_as_float(s[kk]), __uint_as_float(s[kk + 1])), scale2, negm2), x0, x1);
          const bool poly = ((MODE == 4 || MODE == 5) && (i & 3) == 3) || (MODE == 6 && (i & 1) == 1);
          if (poly) poly_exp2_pair(x0, x1, p0, p1);
          else { p0 = ex2(x0); p1 = ex2(x1); }
          if (MODE == 1 || MODE == 4) rs2[i & 1] = add2(rs2[i & 1], pk(p0, p1));
          w[i] = pack_h2(p0, p1);
        }
      }
      ST32(w, t + 128 + c * 32);
    }
    float r0, r1, r2, r3;
    upk(rs2[0], r0, r1); upk(rs2[1], r2, r3);
    l += (r0 + r1) + (r2 + r3) + (rs[0] + rs[1]) + (rs[2] + rs[3]);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  const long long c1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l;
  if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = c1 - c0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

template <int MODE>
void run(const char* name, int warps) {
  const int sms = 148, iters = 2048;
  float* out;
  long long* cyc;
  cudaMalloc(&out, sms * warps * 32 * sizeof(float));
  cudaMallocManaged(&cyc, sizeof(long long));
  k<MODE><<<sms, warps * 32>>>(out, 8, cyc, 0.18f, 1.0f);
  k<MODE><<<sms, warps * 32>>>(out, iters, cyc, 0.18f, 1.0f);
  cudaDeviceSynchronize();
  printf("%-52s warps/SMSP=%d  %8.1f clk per 128-element row per warp  = %5.2f clk/element/SMSP (err=%s)\n", name,
         warps / 4, (double)*cyc / iters, (double)*cyc / iters / 128 / (warps / 4), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {4, 8}) {
    run<0>("0 scalar FFMA/MUFU/FADD/F2FP", w);
    run<1>("1 FFMA2 + MUFU + FADD2 + F2FP", w);
    run<2>("2 FFMA2 + MUFU + F2FP (no row sum)", w);
    run<3>("3 FFMA2 + F2FP + MUFU.f16x2 (no row sum)", w);
    run<4>("4 mode 1, every 4th pair polynomial", w);
    run<5>("5 mode 2, every 4th pair polynomial", w);
    run<6>("6 mode 2, every 2nd pair polynomial", w);
  }
  return 0;
}
