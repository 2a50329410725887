// Instruction-throughput microbenchmark for sm_100a: FP32 (scalar and packed f32x2), ALU
// selects/compares, MUFU and FP64, in warp-instructions per clock per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
// Measurement tool only; not part of the product library.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(256) bench(float* out, float seed) {
  float a[CHAINS], b = seed + 1.0001f, c = seed * 0.5f + 0.25f;
  unsigned long long p[CHAINS];
  double d[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    a[i] = seed + threadIdx.x * 1e-3f + i;
    float2 v = make_float2(a[i], a[i] + 1.f);
    p[i] = *reinterpret_cast<unsigned long long*>(&v);
    d[i] = a[i];
  }
  float2 bv = make_float2(b, b), cv = make_float2(c, c);
  unsigned long long pb = *reinterpret_cast<unsigned long long*>(&bv);
  unsigned long long pc = *reinterpret_cast<unsigned long long*>(&cv);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
      if (OP == 2) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
      if (OP == 3) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
      if (OP == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
      if (OP == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));
      if (OP == 6) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 7) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 8) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 9) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 10) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
      if (OP == 11) asm volatile("{.reg .pred q; setp.lt.f32 q, %0, %1; selp.f32 %0, %2, %0, q;}"
                                 : "+f"(a[i]) : "f"(b), "f"(c));
      if (OP == 12) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"((double)b), "d"((double)c));
      if (OP == 13) {  // 4 FFMA + 1 MUFU interleaved: do the pipes overlap?
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      }
      if (OP == 14) {  // 2 FFMA + 1 select pair
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
      }
      if (OP == 15) asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
      if (OP == 16) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 17) {  // FFMA2 + MUFU mix: 2 packed fma (=4 scalar) + 1 MUFU
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    float2 v = *reinterpret_cast<float2*>(&p[i]);
    s += a[i] + v.x + v.y + (float)d[i];
  }
  if (s == 12345.678f) out[0] = s;
}

template <int OP>
void run(const char* name, int ops_per_iter, float* out, int sm, double ghz) {
  const int blocks = sm * 8, threads = 256;
  bench<OP><<<blocks, threads>>>(out, 0.5f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<OP><<<blocks, threads>>>(out, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_inst = (double)blocks * (threads / 32) * ITERS * CHAINS * ops_per_iter;
  double per_smsp_clk = warp_inst / (ms * 1e-3) / (ghz * 1e9) / (sm * 4);
  printf("%-28s %8.3f ms  %6.3f warp-inst/clk/SMSP (at %.3f GHz)  err=%d\n", name, ms,
         per_smsp_clk, ghz, (int)cudaGetLastError());
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz * 1e-6;
  printf("%s, %d SMs, clockRate %.3f GHz\n", pr.name, pr.multiProcessorCount, ghz);
  float* out; cudaMalloc(&out, 4);
  int sm = pr.multiProcessorCount;
  run<0>("FFMA", 1, out, sm, ghz);
  run<1>("FFMA2 (f32x2)", 1, out, sm, ghz);
  run<2>("FMUL", 1, out, sm, ghz);
  run<3>("FADD", 1, out, sm, ghz);
  run<4>("FMUL2", 1, out, sm, ghz);
  run<5>("FADD2", 1, out, sm, ghz);
  run<15>("FFMA.SAT", 1, out, sm, ghz);
  run<6>("MUFU.RCP", 1, out, sm, ghz);
  run<7>("MUFU.EX2", 1, out, sm, ghz);
  run<8>("MUFU.LG2", 1, out, sm, ghz);
  run<9>("MUFU.RSQ", 1, out, sm, ghz);
  run<16>("MUFU.SQRT", 1, out, sm, ghz);
  run<10>("FMNMX", 1, out, sm, ghz);
  run<11>("FSETP+FSEL", 2, out, sm, ghz);
  run<12>("DFMA", 1, out, sm, ghz);
  run<13>("4 FFMA + 1 MUFU", 5, out, sm, ghz);
  run<14>("2 FFMA + 1 FMNMX", 3, out, sm, ghz);
  run<17>("2 FFMA2 + 1 MUFU", 3, out, sm, ghz);
  return 0;
}
