// Micro-benchmark (clean version): tcgen05.ld / tcgen05.st throughput per SM vs the number and placement of warps.
// Results are consumed through a static XOR of a few registers (no dynamic register indexing -> no local memory).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_rates2 tmem_rates2.cu && ./tmem_rates2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(r, addr)                                                                                               \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                            \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                             \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"             \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),            \
                 "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),          \
                 "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),          \
                 "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                               \
               : "r"(addr)                                                                                           \
               : "memory")
#define ST32(r, addr)                                                                                               \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                      \
               "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                            \
               "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(addr),       \
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),    \
               "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),        \
               "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),       \
               "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                    \
               : "memory")

// MODE 0: 4 x LDTM.x32 then wait (a 128-column row per iteration, like the softmax)  1: 4 x STTM.x32 then wait
// MODE 2: LDTM.x32, wait, LDTM.x32, wait ... (latency-exposed)   3: 128-col load + 128 FFMA on the values
template <int MODE>
__global__ void k(float* out, int iters, long long* cycles) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warps w and w+4 share a lane quarter (and an SM sub-partition); they use different column ranges
  const uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  uint32_t r0[32], r1[32], r2[32], r3[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { r0[i] = threadIdx.x + i; r1[i] = i; r2[i] = 2 * i; r3[i] = 3 * i; }
  uint32_t acc = 0;
  float facc = 0.f;
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 3) {
      LD32(r0, t); LD32(r1, t + 32); LD32(r2, t + 64); LD32(r3, t + 96);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else if (MODE == 1) {
      ST32(r0, t); ST32(r1, t + 32); ST32(r2, t + 64); ST32(r3, t + 96);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else {
      LD32(r0, t); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      LD32(r1, t + 32); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      LD32(r2, t + 64); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      LD32(r3, t + 96); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        facc += __uint_as_float(r0[i]) * 1.5f + __uint_as_float(r1[i]) * 0.5f + __uint_as_float(r2[i]) * 0.25f +
                __uint_as_float(r3[i]) * 0.125f;
    } else {
      acc ^= r0[0] ^ r1[5] ^ r2[17] ^ r3[31];
      if (MODE == 1) r0[0] = acc + it;
    }
  }
  const long long c1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc) + facc;
  if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = c1 - c0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

template <int MODE>
void run(const char* name, int warps) {
  const int sms = 148, iters = 4096;
  float* out;
  long long* cyc;
  cudaMalloc(&out, sms * warps * 32 * sizeof(float));
  cudaMallocManaged(&cyc, sizeof(long long));
  k<MODE><<<sms, warps * 32>>>(out, 8, cyc);
  k<MODE><<<sms, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  const double bytes_per_sm = (double)warps * iters * 4 * 32 * 32 * 4;
  printf("%-28s warps/SM=%2d  %8.1f clk per 128-col row per warp   %7.1f B/clk/SM  (err=%s)\n", name, warps,
         (double)*cyc / iters, bytes_per_sm / (double)*cyc, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {1, 2, 4, 8, 16}) {
    run<0>("LDTM 4 x .x32 + wait", w);
    run<2>("LDTM (.x32 + wait) x 4", w);
    run<1>("STTM 4 x .x32 + wait", w);
    run<3>("LDTM 128 cols + 128 FFMA", w);
  }
  return 0;
}
