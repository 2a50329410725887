// K3 -- per-pixel time aggregates of a (T, N) field, optionally grouped by a per-step index
// (annual sums) and optionally converted on the fly (x * scale / divisor, negative terms
// dropped), as the calibration's precipitation constraint forms them
// (/root/reference/mod16/calibration.py:783-790).  HBM-bound: each element is read once
// (4 or 8 bytes, + the divisor), accumulation is in double.
//
// Layout: one thread per pixel (a warp reads 128 contiguous bytes of every row), a slice
// of time steps per block row; partial sums are flushed with double atomics whenever the
// group index changes (group[t] is uniform across the block, so the flush is
// warp-uniform).  The domain total per group is a warp-shuffle reduction of the flushed
// values followed by one atomic per warp.
#include "mod16_common.cuh"

namespace mod16 {

template <typename R>
struct AggArgs {
  const R* x;
  const R* div;
  double scale;
  const int32_t* group;
  int n_groups;
  double* out;
  double* total;
  int64_t T, N;
  int t_chunk;
  int clamp_neg, skip_nan;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

template <typename R>
__global__ void __launch_bounds__(256)
time_aggregate_kernel(const __grid_constant__ AggArgs<R> a) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t t0 = (int64_t)blockIdx.y * a.t_chunk;
  int64_t t1 = t0 + a.t_chunk;
  if (t1 > a.T) t1 = a.T;
  const bool live = n < a.N;
  double acc = 0.0;
  int cur = -1;
  auto flush = [&](int g) {  // block-uniform call
    if (g >= 0 && g < a.n_groups) {
      if (live) atomicAdd(a.out + (int64_t)g * a.N + n, acc);
      if (a.total) {
        const double w = warp_sum(live ? acc : 0.0);
        if ((threadIdx.x & 31) == 0) atomicAdd(a.total + g, w);
      }
    }
    acc = 0.0;
  };
  for (int64_t t = t0; t < t1; ++t) {
    const int g = a.group ? __ldg(a.group + t) : 0;
    if (g != cur) {
      if (cur != -1) flush(cur);
      cur = g;
    }
    if (live && g >= 0 && g < a.n_groups) {
      double v = (double)__ldg(a.x + t * a.N + n);
      if (a.div) v = (v * a.scale) / (double)__ldg(a.div + t * a.N + n);  // :783
      else v = v * a.scale;
      if (a.clamp_neg) v = v < 0.0 ? 0.0 : v;                             // :784
      if (a.skip_nan) v = v != v ? 0.0 : v;
      acc += v;
    }
  }
  if (cur != -1) flush(cur);
}

template <typename R>
static int run_aggregate(const mod16_aggregate_t* g, int64_t T, int64_t N, uint32_t flags,
                         cudaStream_t st) {
  if (!g) return MOD16_ERR_NULL;
  if (T < 0 || N < 0 || g->n_groups < 1) return MOD16_ERR_SHAPE;
  if (flags & ~(MOD16_AGG_CLAMP_NEG | MOD16_AGG_SKIP_NAN)) return MOD16_ERR_FLAGS;
  if (!g->out) return MOD16_ERR_NULL;
  M16_CUDA_CHECK(cudaMemsetAsync(g->out, 0, sizeof(double) * (size_t)g->n_groups * (size_t)N, st));
  if (g->total) M16_CUDA_CHECK(cudaMemsetAsync(g->total, 0, sizeof(double) * g->n_groups, st));
  if (T == 0 || N == 0) return MOD16_OK;
  if (!g->x) return MOD16_ERR_NULL;
  AggArgs<R> a;
  a.x = (const R*)g->x; a.div = (const R*)g->divisor; a.scale = g->scale;
  a.group = g->group; a.n_groups = g->n_groups; a.out = g->out; a.total = g->total;
  a.T = T; a.N = N;
  a.clamp_neg = (flags & MOD16_AGG_CLAMP_NEG) ? 1 : 0;
  a.skip_nan = (flags & MOD16_AGG_SKIP_NAN) ? 1 : 0;
  const int threads = 256;
  const int64_t bx = (N + threads - 1) / threads;
  if (bx > 2147483647LL) return MOD16_ERR_SHAPE;
  // time slices: enough block rows for several waves of 148 SMs x 8 resident blocks
  int64_t want_rows = (148 * 8 * 2 + bx - 1) / bx;
  if (want_rows < 1) want_rows = 1;
  int64_t chunk = (T + want_rows - 1) / want_rows;
  if (chunk < 1) chunk = 1;
  int64_t by = (T + chunk - 1) / chunk;
  while (by > 65535) { chunk *= 2; by = (T + chunk - 1) / chunk; }
  a.t_chunk = (int)chunk;
  time_aggregate_kernel<R><<<dim3((unsigned)bx, (unsigned)by), threads, 0, st>>>(a);
  M16_CUDA_CHECK(cudaGetLastError());
  return MOD16_OK;
}

}  // namespace mod16

extern "C" {
int mod16_time_aggregate_f32(const mod16_aggregate_t* agg, int64_t T, int64_t N, uint32_t flags,
                             void* stream) {
  return mod16::run_aggregate<float>(agg, T, N, flags, (cudaStream_t)stream);
}
int mod16_time_aggregate_f64(const mod16_aggregate_t* agg, int64_t T, int64_t N, uint32_t flags,
                             void* stream) {
  return mod16::run_aggregate<double>(agg, T, N, flags, (cudaStream_t)stream);
}
}
