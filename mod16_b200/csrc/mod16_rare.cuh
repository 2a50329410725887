// Shared pieces of the float32 fast-path kernels (pipelined and strided): the shared-memory
// copy of the BPLUT, input gathering of ONE element, and the out-of-line handling of the rare
// elements the fast math flags -- the masked-driver rule and the op-for-op evaluation in
// double (mod16_math.cuh).  Everything here is off the hot path by construction: the
// functions are __noinline__ and self-contained (they re-read the element's inputs from
// global memory, which the fast pass has just pulled through L2), so that the double code
// costs the hot loops no registers.
#pragma once
#include "mod16_common.cuh"
#include "mod16_fast.cuh"

namespace mod16 {

using fast::PftConst;
using fast::PixConst;
using fast::StepIn;

// Shared-memory copy of the BPLUT: derived float constants for the fast math and the raw
// double table for the op-for-op fallback.
// Column MOD16_N_PFT is an all-NaN column for land-cover codes outside 0..12.
struct SharedLut {
  PftConst pc[MOD16_N_PFT + 1];
  double tab[MOD16_N_PARAMS][MOD16_N_PFT + 1];
};

template <int SEM>
__device__ __forceinline__ void stage_lut(const ForwardArgs<float>& a, SharedLut& s) {
  for (int i = threadIdx.x; i < MOD16_N_PARAMS * (MOD16_N_PFT + 1); i += blockDim.x) {
    const int k = i / (MOD16_N_PFT + 1), c = i % (MOD16_N_PFT + 1);
    s.tab[k][c] = c < MOD16_N_PFT ? a.table[k][c] : __longlong_as_double(0x7ff8000000000000LL);
  }
  __syncthreads();
  if (threadIdx.x < MOD16_N_PFT + 1) {
    double col[MOD16_N_PARAMS];
#pragma unroll
    for (int k = 0; k < MOD16_N_PARAMS; ++k) col[k] = s.tab[k][threadIdx.x];
    s.pc[threadIdx.x] = fast::make_pft_const(col, SEM);
  }
  __syncthreads();
}

// The BPLUT column of pixel n in double, for any parameter source.
__device__ __forceinline__ void load_column(const ForwardArgs<float>& a, int64_t n,
                                            double col[MOD16_N_PARAMS]) {
  if (a.par_mode == 2) {
#pragma unroll
    for (int k = 0; k < MOD16_N_PARAMS; ++k)
      col[k] = (double)__ldg(a.par_arr[k] + n * a.par_stride[k]);
  } else {
    int code = 0;
    if (a.par_mode == 1) code = (int)__ldg(a.pft + n);
    // numpy fancy indexing would raise on a code > 12; the kernels yield NaN parameters
#pragma unroll
    for (int k = 0; k < MOD16_N_PARAMS; ++k)
      col[k] = code < MOD16_N_PFT ? a.table[k][code] : __longlong_as_double(0x7ff8000000000000LL);
  }
}

__device__ __forceinline__ StepIn load_step(const ForwardArgs<float>& a, int64_t t, int64_t n) {
  StepIn in;
  in.lw_d = a.drv[MOD16_D_LW_NET_DAY].at(t, n);
  in.lw_n = a.drv[MOD16_D_LW_NET_NIGHT].at(t, n);
  in.sw_d = a.drv[MOD16_D_SW_RAD_DAY].at(t, n);
  in.sw_n = a.drv[MOD16_D_SW_RAD_NIGHT].at(t, n);
  in.alb = a.drv[MOD16_D_SW_ALBEDO].at(t, n);
  in.t_d = a.drv[MOD16_D_TEMP_DAY].at(t, n);
  in.t_n = a.drv[MOD16_D_TEMP_NIGHT].at(t, n);
  in.t_ann = a.drv[MOD16_D_TEMP_ANNUAL].at(t, n);
  in.tmin = a.drv[MOD16_D_TMIN].at(t, n);
  in.vpd_d = a.drv[MOD16_D_VPD_DAY].at(t, n);
  in.vpd_n = a.drv[MOD16_D_VPD_NIGHT].at(t, n);
  in.p = a.drv[MOD16_D_PRESSURE].at(t, n);
  in.fpar = a.drv[MOD16_D_FPAR].at(t, n);
  in.lai = a.drv[MOD16_D_LAI].at(t, n);
  in.rc_d = a.has_rc ? a.rc[0].at(t, n) : 0.0f;
  in.rc_n = a.has_rc ? a.rc[1].at(t, n) : 0.0f;
  return in;
}

template <int OUT>
__device__ __forceinline__ void store_exact(const ForwardArgs<float>& a, int64_t o,
                                            const fast::StepOut& r) {
  if (OUT == 0) {
    a.out[0][o] = r.tot;
  } else if (OUT == 1) {
    a.out[0][o] = r.tot_day;
    a.out[1][o] = r.tot_night;
  } else {
    a.out[0][o] = r.day.canopy;   a.out[1][o] = r.day.soil;   a.out[2][o] = r.day.trans;
    a.out[3][o] = r.night.canopy; a.out[4][o] = r.night.soil; a.out[5][o] = r.night.trans;
  }
}

// Exact re-evaluation of ONE element (op-for-op double code), out of line.
template <int SEM, int OUT>
__device__ __noinline__ bool fix_exact_one(const ForwardArgs<float>& a, int64_t t, int64_t n) {
  double col[MOD16_N_PARAMS];
  load_column(a, n, col);
  const StepIn in = load_step(a, t, n);
  fast::StepOut r;
  bool g_pos;
  fast::pixel_step_exact<SEM>(in, col, a.has_rc != 0, a.gate, r, g_pos);
  store_exact<OUT>(a, t * a.N + n, r);
  return g_pos;
}

// ONE element (pixel n, time step t) that step_fast() flagged: the masked-driver rule first
// (NaN totals without the double code: fast::nan_halves), else the op-for-op double code.
// Writes the outputs; returns the element's gate predicate g_surf > 0 (:343).
// `rule_only`: apply the rule, do NOT fall back to the double code (returns -1 then) -- the
// pipelined kernel queues such elements instead of stalling its ring on them.
template <int SEM, int OUT>
__device__ __noinline__ int rare_one(const ForwardArgs<float>& a, int64_t t, int64_t n,
                                     bool rule_only) {
  double col[MOD16_N_PARAMS];
  load_column(a, n, col);
  const StepIn in = load_step(a, t, n);
  const int64_t o = t * a.N + n;
  bool g_pos = false;
  if (OUT != 2) {
    const PftConst pc = fast::make_pft_const(col, SEM);
    const int h = fast::nan_halves<SEM>(in, pc, a.has_rc != 0, g_pos);
    if (OUT == 0 ? h != 0 : h == 3) {
      const float qnan = __int_as_float(0x7fc00000);
      a.out[0][o] = qnan;
      if (OUT == 1) a.out[1][o] = qnan;
      return g_pos ? 1 : 0;
    }
  }
  if (rule_only) return -1;
  fast::StepOut r;
  fast::pixel_step_exact<SEM>(in, col, a.has_rc != 0, a.gate, r, g_pos);
  store_exact<OUT>(a, o, r);
  return g_pos ? 1 : 0;
}

}  // namespace mod16
