// K3 -- per-pixel time aggregates of a (T, N) field, optionally grouped by a per-step index
// (annual sums) and optionally converted on the fly (x * scale / divisor, negative terms
// dropped), as the calibration's precipitation constraint forms them
// (/root/reference/mod16/calibration.py:783-790).  HBM-bound: each element is read once
// (4 or 8 bytes, + the divisor), accumulation is in double.
//
// Layout: one thread per 16 bytes of a row (4 float / 2 double pixels; one pixel per thread
// for ragged or unaligned rows), a slice of time steps per block row walked in batches of 8
// steps whose loads are all issued before the first is used (memory-level parallelism
// instead of occupancy).  Determinism: a thread owns its pixels for ALL time steps (time is
// never split across blocks), so a partial sum is flushed with a plain read-modify-write of
// out[g][n] whenever the group index changes -- no floating-point atomics, the summation
// order is the time order, results are bit-reproducible run to run.  The domain total per
// group is a second, tiny launch: one block per group sums out[g][:] in a fixed order
// (per-thread strided partials, then a fixed shuffle / shared-memory tree).
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
  int clamp_neg, skip_nan;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// VEC pixels per thread (one 16-byte load per step and field), AGG_UNROLL steps in flight.
// 8 steps in flight per thread (16 without a divisor spills at 3 blocks/SM: 2.3 vs 4.3 TB/s)

template <typename R, int VEC>
struct alignas(sizeof(R) * VEC) Pack { R v[VEC]; };

template <typename R, int VEC, bool HAS_DIV>
__global__ void __launch_bounds__(256, HAS_DIV ? 2 : 3)
time_aggregate_kernel(const __grid_constant__ AggArgs<R> a) {
  typedef Pack<R, VEC> P;
  const int64_t n0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  const int64_t t0 = 0, t1 = a.T;
  constexpr int AGG_UNROLL = 8;
  const bool live = n0 < a.N;   // (VEC > 1 only when N % VEC == 0: the whole pack is live)
  double acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.0;
  int cur = -1;
  auto flush = [&](int g) {  // block-uniform call; this thread alone owns out[g][n0 ..]
    if (g >= 0 && g < a.n_groups && live) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) a.out[(int64_t)g * a.N + n0 + k] += acc[k];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.0;
  };
  for (int64_t tb = t0; tb < t1; tb += AGG_UNROLL) {
    // issue the loads of the whole batch first: AGG_UNROLL x (1 or 2) 16-byte loads in flight
    P xs[AGG_UNROLL], ds[HAS_DIV ? AGG_UNROLL : 1];
    int gs[AGG_UNROLL];
#pragma unroll
    for (int u = 0; u < AGG_UNROLL; ++u) {
      const int64_t t = tb + u;
      gs[u] = -1;  // (steps past the end are recognised by their index, not by a sentinel)
      if (t < t1) {
        gs[u] = a.group ? __ldg(a.group + t) : 0;
        if (live && gs[u] >= 0 && gs[u] < a.n_groups) {
          xs[u] = *reinterpret_cast<const P*>(a.x + t * a.N + n0);
          if (HAS_DIV) ds[u] = *reinterpret_cast<const P*>(a.div + t * a.N + n0);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < AGG_UNROLL; ++u) {
      if (tb + u >= t1) break;  // past the last step (block-uniform)
      int g = gs[u];
      if (g < 0 || g >= a.n_groups) g = -1;  // every out-of-range group: skipped
      if (g != cur) {
        if (cur != -1) flush(cur);
        cur = g;
      }
      if (live && g >= 0 && g < a.n_groups) {
        // the term in the arithmetic of the entry (numpy evaluates float32 arrays in
        // float32, :783), the accumulation in double
        const R scale = (R)a.scale;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          R v = xs[u].v[k];
          if (HAS_DIV) v = (v * scale) / ds[u].v[k];           // :783
          else v = v * scale;
          if (a.clamp_neg) v = v < R(0) ? R(0) : v;            // :784
          if (a.skip_nan) v = v != v ? R(0) : v;
          acc[k] += (double)v;
        }
      }
    }
  }
  if (cur != -1) flush(cur);
}

// total[g] = sum over n of out[g][n], fixed order: thread i adds elements i, i + 256, ...;
// then a fixed shuffle tree per warp and a serial sum of the eight warp results.
__global__ void __launch_bounds__(256)
aggregate_total_kernel(const double* __restrict__ out, double* __restrict__ total, int64_t N) {
  __shared__ double part[8];
  const double* row = out + (int64_t)blockIdx.x * N;
  double s = 0.0;
  for (int64_t n = threadIdx.x; n < N; n += 256) s += row[n];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += part[w];
    total[blockIdx.x] = t;
  }
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
  // 16-byte loads when the rows allow them (N a multiple of the pack, aligned bases)
  constexpr int VEC = 16 / (int)sizeof(R);
  // (small N: one pixel per thread, so that more threads share the serial walk over time)
  const bool vec = N % VEC == 0 && N >= 148 * threads * VEC &&
                   (reinterpret_cast<uintptr_t>(g->x) & 15) == 0 &&
                   (!g->divisor || (reinterpret_cast<uintptr_t>(g->divisor) & 15) == 0);
  const int64_t per_block = (int64_t)threads * (vec ? VEC : 1);
  const int64_t bx = (N + per_block - 1) / per_block;
  if (bx > 2147483647LL) return MOD16_ERR_SHAPE;
  // (time is not split across blocks: see the note on determinism above)
  const dim3 grid((unsigned)bx, 1);
  if (vec && a.div) time_aggregate_kernel<R, VEC, true><<<grid, threads, 0, st>>>(a);
  else if (vec) time_aggregate_kernel<R, VEC, false><<<grid, threads, 0, st>>>(a);
  else if (a.div) time_aggregate_kernel<R, 1, true><<<grid, threads, 0, st>>>(a);
  else time_aggregate_kernel<R, 1, false><<<grid, threads, 0, st>>>(a);
  M16_CUDA_CHECK(cudaGetLastError());
  if (g->total) {
    aggregate_total_kernel<<<g->n_groups, 256, 0, st>>>(g->out, g->total, N);
    M16_CUDA_CHECK(cudaGetLastError());
  }
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
