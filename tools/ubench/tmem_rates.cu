// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_rates tmem_rates.cu && ./tmem_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>  // 0 = ld x32, 1 = st x32, 2 = ld x32 + 32 MUFU (overlap test)
__global__ void k(float* out, int iters) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (MODE == 0 || MODE == 2) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(t + c * 32)
            : "memory");
      } else {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(t + c * 32),
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
            "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
            "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
            "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
            : "memory");
      }
    }
    if (MODE == 0 || MODE == 2) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    else asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float f = __uint_as_float(r[i]);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f));
        acc += f;
      }
    } else {
      acc += __uint_as_float(r[it & 31]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

template <int MODE>
void run(const char* name, int warps) {
  const int sms = 148, iters = 2048;
  float* out;
  cudaMalloc(&out, sms * warps * 32 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<sms, warps * 32>>>(out, 8);
  cudaEventRecord(e0);
  k<MODE><<<sms, warps * 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes_per_sm = (double)warps * iters * 4 * 32 * 32 * 4;
  printf("%-16s warps/SM=%2d  %.3f ms  %.1f B/ns/SM = %.1f B/clk/SM @1.9GHz  (err=%s)\n", name, warps, ms,
         bytes_per_sm / (ms * 1e6), bytes_per_sm / (ms * 1e6) / 1.9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int w : {1, 4, 8}) {
    run<0>("LDTM x32", w);
    run<1>("STTM x32", w);
    run<2>("LDTM+128 MUFU", w);
  }
  return 0;
}
