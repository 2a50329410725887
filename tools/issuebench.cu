// How fast can ONE producer thread issue cp.async.bulk copies?  One block per SM, a producer
// warp streaming NROWS row copies of ROWB bytes per stage, consumers only touch the data.
// Variants: V0 elected thread inside a divergent branch (pointers from shared memory),
//           V1 whole warp runs the loop uniformly, pointers from the constant bank, copies
//              predicated on one lane.
// Measurement tool only.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issuebench issuebench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define NARR 11
struct Src { const float* p[NARR]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

// stage = SPLIT sub-rows per array x NARR arrays, each ROWB bytes => NARR*SPLIT copies per stage
template <int ROWB, int SPLIT, int STAGES, int V, int NPROD>
__global__ void __launch_bounds__(480 + 32 * NPROD) k(const __grid_constant__ Src src, float* out, int64_t N, int T) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
  __shared__ const float* sptr[NARR];
  const int tid = threadIdx.x;
  constexpr uint32_t STAGEB = ROWB * SPLIT * NARR;
  const int64_t tile0 = (int64_t)blockIdx.x * (ROWB / 4);
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty), data0 = smem_u32(smem);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, NPROD); mbar_init(empty0 + 8 * s, 15); }
    for (int k2 = 0; k2 < NARR; ++k2) sptr[k2] = src.p[k2];
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid >= 480) {
    const int pw = (tid - 480) / 32;  // producer warp index
    if (V == 0) {
      if ((tid & 31) == 0) {
        for (int i = 0; i < T; ++i) {
          const int s = i % STAGES;
          if (i >= STAGES) mbar_wait(empty0 + 8 * s, ((i / STAGES) - 1) & 1);
          mbar_expect_tx(full0 + 8 * s, STAGEB / NPROD);
          for (int u = 0; u < SPLIT; ++u)
#pragma unroll 4
            for (int k2 = pw; k2 < NARR; k2 += NPROD)
              bulk_g2s(data0 + s * STAGEB + (u * NARR + k2) * ROWB, sptr[k2] + ((int64_t)i * SPLIT + u) * N + tile0, ROWB, full0 + 8 * s);
        }
      }
    } else {
      for (int i = 0; i < T; ++i) {
        const int s = i % STAGES;
        if (i >= STAGES) mbar_wait(empty0 + 8 * s, ((i / STAGES) - 1) & 1);
        if (elect_one()) {
          mbar_expect_tx(full0 + 8 * s, STAGEB / NPROD);
#pragma unroll
          for (int u = 0; u < SPLIT; ++u)
#pragma unroll
            for (int k2 = 0; k2 < NARR; ++k2)
              if (k2 % NPROD == pw)
                bulk_g2s(data0 + s * STAGEB + (u * NARR + k2) * ROWB, src.p[k2] + ((int64_t)i * SPLIT + u) * N + tile0, ROWB, full0 + 8 * s);
        }
        __syncwarp();
      }
    }
    return;
  }
  float acc = 0.f;
  for (int i = 0; i < T; ++i) {
    const int s = i % STAGES;
    mbar_wait(full0 + 8 * s, (i / STAGES) & 1);
    acc += reinterpret_cast<const float*>(smem + s * STAGEB)[tid];
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(empty0 + 8 * s);
  }
  out[(int64_t)blockIdx.x * 480 + tid] = acc;
}

template <int ROWB, int SPLIT, int STAGES, int V, int NPROD>
void run(const char* name, Src src, float* out, int64_t N, int Tsteps) {
  auto kk = k<ROWB, SPLIT, STAGES, V, NPROD>;
  const size_t smem = (size_t)STAGES * ROWB * SPLIT * NARR;
  cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int blocks = (int)(N / (ROWB / 4));
  const int T = Tsteps / SPLIT;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kk<<<blocks, 480 + 32 * NPROD, smem>>>(src, out, N, T);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double bytes = (double)NARR * N * T * SPLIT * 4;
    double copies = (double)blocks * T * SPLIT * NARR;
    if (rep == 1)
      printf("%-40s %8.3f ms %7.1f GB/s  %6.2f copies/us/producer-warp  err=%d\n", name, ms, bytes / ms * 1e-6,
             copies / 148.0 / NPROD / (ms * 1e3), (int)cudaGetLastError());
  }
}

int main() {
  const int64_t N = 5760000; const int T = 32;
  Src src; float* out;
  for (int k2 = 0; k2 < NARR; ++k2) { float* p; cudaMalloc(&p, N * T * 4); cudaMemset(p, 0, N * T * 4); src.p[k2] = p; }
  cudaMalloc(&out, 64 << 20);
  run<3840, 1, 4, 0, 1>("3840B x11/stage st4 V0 1prod", src, out, N, T);
  run<3840, 1, 4, 1, 1>("3840B x11/stage st4 V1 1prod", src, out, N, T);
  run<1920, 2, 5, 0, 1>("1920B x22/stage st5 V0 1prod", src, out, N, T);
  run<1920, 2, 5, 1, 1>("1920B x22/stage st5 V1 1prod", src, out, N, T);
  run<1920, 2, 5, 1, 2>("1920B x22/stage st5 V1 2prod", src, out, N, T);
  run<1920, 2, 5, 0, 2>("1920B x22/stage st5 V0 2prod", src, out, N, T);
  run<1920, 4, 2, 1, 1>("1920B x44/stage st2 V1 1prod", src, out, N, T);
  run<960, 4, 5, 1, 1>("960B x44/stage st5 V1 1prod", src, out, N, T);
  run<960, 4, 5, 1, 2>("960B x44/stage st5 V1 2prod", src, out, N, T);
  return 0;
}
