// Micro-benchmark: per-SM instruction throughput of the pipes the attention softmax leans on (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
  unsigned u[4] = {0, 0, 0, 0};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 2) asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u[i & 3]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      if (OP == 3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
      if (OP == 4) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
      if (OP == 5) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i & 3]));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + u[0] + u[1] + u[2] + u[3];
}

template <int OP>
void run(const char* name, int warps_per_sm) {
  int sms = 148, iters = 4096;
  float* out;
  cudaMalloc(&out, sms * warps_per_sm * 32 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<OP><<<sms, warps_per_sm * 32>>>(out, 16, 0.5f);
  cudaEventRecord(e0);
  k<OP><<<sms, warps_per_sm * 32>>>(out, iters, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double lane_ops = (double)sms * warps_per_sm * 32 * iters * 8;
  printf("%-22s warps/SM=%2d  %.3f ms  %.1f lane-ops/ns/chip = %.2f lane-ops/clk/SM @1.9GHz\n", name, warps_per_sm, ms,
         lane_ops / (ms * 1e6), lane_ops / (ms * 1e6) / 148 / 1.9);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("MUFU.EX2 f32", w);
    run<5>("MUFU.EX2 f16x2", w);
    run<1>("FFMA", w);
    run<2>("F2FP f16x2 pack", w);
    run<3>("FMNMX3", w);
    run<4>("FADD", w);
  }
  return 0;
}
