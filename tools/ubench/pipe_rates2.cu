// Micro-benchmark: issue rate (clk per warp instruction per SM sub-partition) of the instructions in the attention
// softmax loop, alone and mixed with MUFU.EX2, timed with clock64 over 8 independent dependency chains.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates2 pipe_rates2.cu && ./pipe_rates2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// OP 0 MUFU.EX2 | 1 F2FP f16x2 | 2 FFMA2 | 3 FADD2 | 4 FFMA | 5 FMNMX3 | 6 MUFU+F2FP (2:1) | 7 MUFU+FFMA2+FADD2+F2FP (2:1:1:1)
// 8 F2FP bf16x2 | 9 MUFU f16x2 | 10 MUFU + 3 FFMA
template <int OP>
__global__ void k(float* out, int iters, float seed, long long* cyc) {
  float a[8];
  uint64_t d[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + threadIdx.x * 1e-3f + i * 0.01f;
    u[i] = threadIdx.x + i;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(d[i]) : "f"(a[i]));
  }
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0 || OP == 6 || OP == 7 || OP == 10) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 6 || OP == 7) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(i + 4) & 7]));
      if (OP == 1 || OP == 6 || OP == 7)
        asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i])));
      if (OP == 8) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i])));
      if (OP == 2 || OP == 7) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(d[i]));
      if (OP == 3 || OP == 7) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(d[(i + 1) & 7]));
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 10) {
        asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[(i + 1) & 7]));
        asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[(i + 2) & 7]));
        asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[(i + 3) & 7]));
      }
      if (OP == 5) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
      if (OP == 9) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
    }
  }
  const long long c1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + u[i] + (float)(d[i] & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = c1 - c0;
}

template <int OP>
void run(const char* name, int warps_per_sm, int instr_per_group) {
  int sms = 148, iters = 2048;
  float* out;
  long long* cyc;
  cudaMalloc(&out, sms * warps_per_sm * 32 * sizeof(float));
  cudaMallocManaged(&cyc, 8);
  k<OP><<<sms, warps_per_sm * 32>>>(out, 16, 0.5f, cyc);
  k<OP><<<sms, warps_per_sm * 32>>>(out, iters, 0.5f, cyc);
  cudaDeviceSynchronize();
  const double groups = (double)iters * 8 * (warps_per_sm / 4);   // groups issued per SM sub-partition
  printf("%-40s warps/SMSP=%d  %6.2f clk per group per SMSP  (%d instr/group -> %5.2f clk/instr)  %s\n", name,
         warps_per_sm / 4, *cyc / groups, instr_per_group, *cyc / groups / instr_per_group,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {4, 8}) {
    run<0>("MUFU.EX2 f32", w, 1);
    run<9>("MUFU.EX2 f16x2", w, 1);
    run<1>("F2FP satfinite f16x2", w, 1);
    run<8>("F2FP bf16x2", w, 1);
    run<2>("FFMA2", w, 1);
    run<3>("FADD2", w, 1);
    run<4>("FFMA", w, 1);
    run<5>("FMNMX3", w, 1);
    run<6>("2 MUFU + F2FP", w, 3);
    run<7>("2 MUFU + FFMA2 + FADD2 + F2FP", w, 5);
    run<10>("MUFU + 3 FFMA", w, 4);
  }
  return 0;
}
