// Shared host/device declarations of the CUDA translation units: kernel argument blocks,
// broadcast-aware loads, error plumbing of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mod16_b200.h"
#include "mod16_math.cuh"

namespace mod16 {

// thread-local record of the last failed CUDA call (mod16_last_cuda_error())
extern thread_local int g_last_cuda_error;
// kernel chosen by the last forward call of this thread (mod16_last_forward_kernel())
extern thread_local int g_last_forward_kernel;

#define M16_CUDA_CHECK(expr)                                   \
  do {                                                         \
    cudaError_t e__ = (expr);                                  \
    if (e__ != cudaSuccess) {                                  \
      ::mod16::g_last_cuda_error = (int)e__;                   \
      return MOD16_ERR_CUDA;                                   \
    }                                                          \
  } while (0)

// One broadcastable input as the kernels see it.
template <typename R>
struct Arr {
  const R* ptr;      // NULL: scalar
  int64_t ts, ns;    // element strides over time / pixels (0 broadcasts)
  R scalar;
  int is_f32;        // (R = double only) ptr holds float32 elements: widened on load
  __device__ __forceinline__ R at(int64_t t, int64_t n) const {
    if (sizeof(R) == 8 && is_f32)
      return ptr ? (R)__ldg(reinterpret_cast<const float*>(ptr) + t * ts + n * ns) : scalar;
    return ptr ? __ldg(ptr + t * ts + n * ns) : scalar;
  }
};

template <typename R>
inline void fill_arr(Arr<R>& dst, const mod16_driver_t& src) {
  dst.ptr = (const R*)src.ptr;
  dst.ts = src.ptr ? src.t_stride : 0;
  dst.ns = src.ptr ? src.n_stride : 0;
  dst.scalar = (R)src.scalar;
  dst.is_f32 = 0;
}

// Storage format of one driver of mod16_forward_raw_f32 (float32 path only).
struct RawInfo {
  int dtype;      // MOD16_T_*
  int has_fill;
  float scale, offset;
  int fill;
};

template <typename R>
struct ForwardArgs {
  Arr<R> drv[MOD16_N_DRIVERS];
  // raw-format call: drv[k].ptr points at elements of raw[k].dtype (strides in elements)
  RawInfo raw[MOD16_N_DRIVERS];
  int raw_fmt;       // 0: every driver float (the plain entries); else see mod16_forward_pipe.cuh
  uint32_t prep;     // MOD16_PREP_*
  Arr<R> ps[2];      // MOD16_PREP_VPD_FROM_QV: surface pressure, day / night
  Arr<R> rc[2];
  int has_rc;
  // parameters
  const uint8_t* pft;
  const R* par_arr[MOD16_N_PARAMS];
  int64_t par_stride[MOD16_N_PARAMS];
  int par_mode;  // 0: table column 0 for everyone, 1: gather by pft, 2: per-pixel arrays
  double table[MOD16_N_PARAMS][MOD16_N_PFT];
  // outputs
  R* out[6];
  const int* group;   // MOD16_OUT_ANNUAL: group (year) of every time step; out[0] = uint64 sums
  int n_groups;
  const float* w_day;    // MOD16_OUT_COMPOSITE: float[T] weights of the day / night totals
  const float* w_night;
  float out_inv_scale;   // MOD16_OUT_SUM_I16: counts per physical unit
  int* gate_flag;
  unsigned long long* redo_ws;  // [0] = count, [1 ..] = queued linear indices t * N + n
  long long redo_cap;           // entries the queue can hold (0: no queue)
  int skip_if_flag;  // fix-up launch: return at once when *gate_flag != 0
  int track_gate;    // OR (g_surf > 0) into *gate_flag
  int gate;          // mod16::Gate
  int64_t T, N;
  int t_chunk;       // time steps per block row
};

// Builds Params<R> for one pixel from the argument block.
template <typename R, int SEM>
__device__ __forceinline__ Params<R> load_params(const ForwardArgs<R>& a, int64_t n,
                                                 const double (*tab)[MOD16_N_PFT]) {
  Params<R> q;
  double v[MOD16_N_PARAMS];
  if (a.par_mode == 2) {
#pragma unroll
    for (int k = 0; k < MOD16_N_PARAMS; ++k)
      v[k] = (double)__ldg(a.par_arr[k] + n * a.par_stride[k]);
  } else {
    int code = 0;
    if (a.par_mode == 1) {
      code = (int)__ldg(a.pft + n);
    }
    // numpy fancy indexing would raise on a code > 12; the kernels yield NaN parameters
#pragma unroll
    for (int k = 0; k < MOD16_N_PARAMS; ++k)
      v[k] = code < MOD16_N_PFT ? tab[k][code] : __longlong_as_double(0x7ff8000000000000LL);
  }
  q.tmin_close = (R)v[0]; q.tmin_open = (R)v[1]; q.vpd_open = (R)v[2];
  q.vpd_close = (R)v[3];  q.gl_sh = (R)v[4];     q.gl_wv = (R)v[5];
  q.g_cuticular = (R)v[6]; q.csl = (R)v[7];      q.rbl_min = (R)v[8];
  q.rbl_max = (R)v[9];    q.beta = (R)v[10];
  q.tann_lo = 273.15 + (SEM == SEM_STATIC ? v[1] : v[0]);
  return q;
}

}  // namespace mod16
