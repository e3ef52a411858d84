// Micro-benchmark (round 2): issue rate (clk per warp instruction per SM sub-partition) of the packed-half and integer
// instructions a half-precision polynomial exp2 would be made of, next to the fp32 ones of pipe_rates2.cu.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates3 pipe_rates3.cu && ./pipe_rates3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// OP 0 HFMA2 | 1 HADD2 | 2 HMNMX2 (max.f16x2) | 3 LOP3 | 4 IADD3 | 5 SHF | 6 PRMT | 7 FFMA (reference) | 8 FMNMX (2-input)
// 9 HFMA2.BF16 | 10 cvt.rn.f16x2.f32 | 11 MUFU.EX2 f32 + 4 HFMA2 | 12 IMAD
template <int OP>
__global__ void k(float* out, int iters, float seed, long long* cyc) {
  float a[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + threadIdx.x * 1e-3f + i * 0.01f;
    u[i] = 0x3c003800u + threadIdx.x + i;     // (1.0, 0.5) as f16x2, perturbed
  }
  const uint32_t c1 = 0x38003800u;             // (0.5, 0.5)
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("fma.rn.f16x2 %0, %0, %1, %0;" : "+r"(u[i]) : "r"(c1));
      if (OP == 1) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(c1));
      if (OP == 2) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(c1));
      if (OP == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      if (OP == 5) asm volatile("shf.l.wrap.b32 %0, %0, %1, 10;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      if (OP == 7) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 8) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
      if (OP == 9) asm volatile("fma.rn.bf16x2 %0, %0, %1, %0;" : "+r"(u[i]) : "r"(c1));
      if (OP == 10) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      if (OP == 11) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %0;" : "+r"(u[i]) : "r"(c1));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %0;" : "+r"(u[(i + 1) & 7]) : "r"(c1));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %0;" : "+r"(u[(i + 2) & 7]) : "r"(c1));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %0;" : "+r"(u[(i + 3) & 7]) : "r"(c1));
      }
      if (OP == 12) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(c1));
    }
  }
  const long long c1c = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = c1c - c0;
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
  const double groups = (double)iters * 8 * (warps_per_sm / 4);
  printf("%-34s warps/SMSP=%d  %6.2f clk per group per SMSP  (%d instr/group -> %5.2f clk/instr)  %s\n", name,
         warps_per_sm / 4, *cyc / groups, instr_per_group, *cyc / groups / instr_per_group,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {4, 8}) {
    run<7>("FFMA (reference)", w, 1);
    run<0>("HFMA2 f16x2", w, 1);
    run<9>("HFMA2 bf16x2", w, 1);
    run<1>("HADD2", w, 1);
    run<2>("HMNMX2 (max.f16x2)", w, 1);
    run<8>("FMNMX (2-input max.f32)", w, 1);
    run<10>("F2FP cvt.rn.f16x2.f32", w, 1);
    run<3>("LOP3", w, 1);
    run<4>("IADD", w, 1);
    run<5>("SHF", w, 1);
    run<6>("PRMT", w, 1);
    run<12>("IMAD", w, 1);
    run<11>("MUFU.EX2 + 4 HFMA2", w, 5);
  }
  return 0;
}
