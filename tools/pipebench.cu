// Bandwidth of the bulk-copy (cp.async.bulk + mbarrier) pipeline alone, as the MOD16 forward
// kernel uses it: NARR separate (T, N) arrays, one block per pixel tile, STAGES stages, the
// consumer only touches the data.  Measurement tool only; not part of the product library.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipebench pipebench.cu
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
#ifndef WAITMODE
#define WAITMODE 0
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if WAITMODE == 0
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(bar), "r"(parity) : "memory");
#elif WAITMODE == 1   // suspend-time hint (ns)
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra D;\nbra W;\nD:\n}" ::"r"(bar), "r"(parity), "r"(20000u) : "memory");
#else                 // test_wait + nanosleep back-off
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nnanosleep.u32 200;\nbra W;\nD:\n}" ::"r"(bar), "r"(parity) : "memory");
#endif
}

// MODE 0: thread 0 of warp 0 is consumer + producer (as forward_pipe_kernel r01)
// MODE 1: dedicated producer warp (warp THREADS/32), consumers never issue
template <int TILE, int STAGES, int THREADS, int MODE, int WORK>
__global__ void __launch_bounds__(THREADS + (MODE == 1 ? 32 : 0))
pipe(Src src, float* out, int64_t N, int T) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
  __shared__ int released[STAGES];
  const int tid = threadIdx.x;
  const int64_t tile0 = (int64_t)blockIdx.x * TILE;
  constexpr uint32_t ROWB = TILE * 4, STAGEB = ROWB * NARR;
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty), data0 = smem_u32(smem);
  constexpr int CW = THREADS / 32;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, CW); released[s] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int s, int t) {
    mbar_expect_tx(full0 + 8 * s, STAGEB);
#pragma unroll 1
    for (int k = 0; k < NARR; ++k)
      bulk_g2s(data0 + s * STAGEB + k * ROWB, src.p[k] + (int64_t)t * N + tile0, ROWB, full0 + 8 * s);
  };
  if (MODE == 1 && tid >= THREADS) {  // producer warp
    if (tid == THREADS) {
      for (int i = 0; i < T; ++i) {
        const int s = i % STAGES;
        if (i >= STAGES) mbar_wait(empty0 + 8 * s, ((i / STAGES) - 1) & 1);
        issue(s, i);
      }
    }
    return;
  }
  if (MODE != 1 && MODE != 3 && tid == 0) for (int s = 0; s < STAGES && s < T; ++s) issue(s, s);
  float acc = 0.f;
  constexpr int PER = TILE / THREADS;  // floats per thread per row
  for (int i = 0; i < T; ++i) {
    const int s = i % STAGES;
    const uint32_t ph = (i / STAGES) & 1;
    if (MODE != 3) mbar_wait(full0 + 8 * s, ph);
    const float* sp = reinterpret_cast<const float*>(smem + s * STAGEB) + tid * PER;
#pragma unroll
    for (int k = 0; k < NARR; ++k)
#pragma unroll
      for (int j = 0; j < PER; ++j) acc += sp[k * TILE + j];
    if (MODE == 2) {  // release at once; the last warp to release re-fills
      __syncwarp();
      if ((tid & 31) == 0) {
        __threadfence_block();
        if (atomicAdd(&released[s], 1) == CW - 1) {
          released[s] = 0;
          __threadfence_block();
          if (i + STAGES < T) issue(s, i + STAGES);
        }
      }
      __syncwarp();
    }
    // synthetic compute: WORK issue-bound FMAs (8 independent chains) per thread and step
    {
      float c[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) c[j] = acc + j;
#pragma unroll 1
      for (int w = 0; w < WORK / 8; ++w) {
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = fmaf(c[j], 1.0000001f, 1e-9f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += c[j];
    }
    if (MODE != 2 && MODE != 3) {
      __syncwarp();
      if ((tid & 31) == 0) {
        mbar_arrive(empty0 + 8 * s);
        if (MODE == 0 && tid == 0 && i + STAGES < T) { mbar_wait(empty0 + 8 * s, ph); issue(s, i + STAGES); }
      }
      __syncwarp();
    }
  }
  out[(int64_t)blockIdx.x * THREADS + tid] = acc;
}

template <int TILE, int STAGES, int THREADS, int MODE, int WORK>
void run(const char* name, Src src, float* out, int64_t N, int T) {
  auto k = pipe<TILE, STAGES, THREADS, MODE, WORK>;
  const size_t smem = (size_t)STAGES * NARR * TILE * 4;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int blocks = (int)(N / TILE);
  const int threads = THREADS + (MODE == 1 ? 32 : 0);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, smem);
  k<<<blocks, threads, smem>>>(src, out, N, T);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<<<blocks, threads, smem>>>(src, out, N, T);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double bytes = (double)NARR * N * T * 4;
  printf("%-44s occ %d  %8.3f ms  %7.1f GB/s  err=%d\n", name, occ, ms, bytes / ms * 1e-6, (int)cudaGetLastError());
}

int main() {
  const int64_t N = 5760000; const int T = 32;
  Src src; float* out;
  for (int k = 0; k < NARR; ++k) { float* p; cudaMalloc(&p, N * T * 4); cudaMemset(p, 0, N * T * 4); src.p[k] = p; }
  cudaMalloc(&out, 64 << 20);
  run<512, 4, 256, 3, 400>("tile512 st4 COMPUTE ONLY work400", src, out, N, T);
  run<512, 4, 256, 3, 600>("tile512 st4 COMPUTE ONLY work600", src, out, N, T);
  run<512, 4, 256, 1, 400>("tile512 st4 mode1 work400", src, out, N, T);
  run<512, 4, 256, 1, 600>("tile512 st4 mode1 work600", src, out, N, T);
  run<512, 4, 256, 2, 400>("tile512 st4 mode2 work400", src, out, N, T);
  run<512, 3, 256, 1, 400>("tile512 st3 mode1 work400 occ3", src, out, N, T);
  run<1024, 4, 512, 1, 400>("tile1024 st4 t512 mode1 work400", src, out, N, T);
  run<256, 4, 128, 1, 400>("tile256 st4 t128 mode1 work400 occ5", src, out, N, T);
  return 0;
}
