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

// MOD16.air_pressure (:415-447): P = 101325 (1 - 0.0065 z / 288.15)^5.256, in double
__device__ __forceinline__ float pressure_from_elevation(float z) {
  return (float)(101325.0 * pow(1.0 - (0.0065 * (double)z) / 288.15, 5.256));
}

// A storage-format element as the float the fused kernel computes with (one FMA, like the
// hot path); `fill` -> NaN.
__device__ __forceinline__ float dequant(int q, const RawInfo& r) {
  if (q == r.fill) return __int_as_float(0x7fc00000);
  return fmaf((float)q, r.scale, r.offset);
}

// Driver k of one element for any call form: float arrays (plain entries) or storage
// formats with the loaders' conversions (mod16_forward_raw_f32).
__device__ __forceinline__ float drv_at(const ForwardArgs<float>& a, int k, int64_t t, int64_t n) {
  const Arr<float>& d = a.drv[k];
  if (a.raw_fmt == 0) return d.at(t, n);
  float v;
  if (!d.ptr) {
    v = d.scalar;
  } else {
    const RawInfo& r = a.raw[k];
    const int64_t i = t * d.ts + n * d.ns;
    if (r.dtype == MOD16_T_I16) v = dequant((int)__ldg((const short*)d.ptr + i), r);
    else if (r.dtype == MOD16_T_U8) v = dequant((int)__ldg((const unsigned char*)d.ptr + i), r);
    else v = __ldg(d.ptr + i);
  }
  if ((k == MOD16_D_VPD_DAY || k == MOD16_D_VPD_NIGHT) && (a.prep & MOD16_PREP_VPD_FROM_QV)) {
    // MOD16.vpd(qv10m, pressure, tmean) (:605-644), in double for the op-for-op pass
    const int h = k == MOD16_D_VPD_DAY ? 0 : 1;
    const double temp = (double)a.drv[h == 0 ? MOD16_D_TEMP_DAY : MOD16_D_TEMP_NIGHT].at(t, n);
    v = (float)vpd_from_qv<double>((double)v, (double)a.ps[h].at(t, n), temp);
  }
  if (k == MOD16_D_VPD_NIGHT && (a.prep & MOD16_PREP_VPD_NIGHT_CLAMP)) v = v < 0.0f ? 0.0f : v;
  if (k == MOD16_D_PRESSURE && (a.prep & MOD16_PREP_ELEVATION)) v = pressure_from_elevation(v);
  return v;
}

__device__ __forceinline__ StepIn load_step(const ForwardArgs<float>& a, int64_t t, int64_t n) {
  StepIn in;
  in.lw_d = drv_at(a, MOD16_D_LW_NET_DAY, t, n);
  in.lw_n = drv_at(a, MOD16_D_LW_NET_NIGHT, t, n);
  in.sw_d = drv_at(a, MOD16_D_SW_RAD_DAY, t, n);
  in.sw_n = drv_at(a, MOD16_D_SW_RAD_NIGHT, t, n);
  in.alb = drv_at(a, MOD16_D_SW_ALBEDO, t, n);
  in.t_d = drv_at(a, MOD16_D_TEMP_DAY, t, n);
  in.t_n = drv_at(a, MOD16_D_TEMP_NIGHT, t, n);
  in.t_ann = drv_at(a, MOD16_D_TEMP_ANNUAL, t, n);
  in.tmin = drv_at(a, MOD16_D_TMIN, t, n);
  in.vpd_d = drv_at(a, MOD16_D_VPD_DAY, t, n);
  in.vpd_n = drv_at(a, MOD16_D_VPD_NIGHT, t, n);
  in.p = drv_at(a, MOD16_D_PRESSURE, t, n);
  in.fpar = drv_at(a, MOD16_D_FPAR, t, n);
  in.lai = drv_at(a, MOD16_D_LAI, t, n);
  in.rc_d = a.has_rc ? a.rc[0].at(t, n) : 0.0f;
  in.rc_n = a.has_rc ? a.rc[1].at(t, n) : 0.0f;
  return in;
}

// ---- MOD16_OUT_ANNUAL: annual evapotranspiration sums in 2^-32 mm ------------------------------
constexpr unsigned long long kAnnualNaN = 1ULL << 54;   // what a NaN element adds
// max(LE, 0) * 86400 / lambda(tc_mean) [mm day-1] as fixed point (calibration.py:783-786,
// :838-842; lambda = (2.501 - 0.002361 Tc) 1e6, :121).  A float times 2^32 converts exactly.
__device__ __forceinline__ unsigned long long annual_fixed(float le, float tc_mean) {
  if (le != le) return kAnnualNaN;
  const float lam = fmaf(tc_mean, -2361.0f, 2.501e6f);
  const float mm = fminf(fmaxf(le, 0.0f) * (86400.0f * 4294967296.0f) / lam, 1.0e16f);
  return __float2ull_rn(mm);
}
__device__ __forceinline__ void annual_add(const ForwardArgs<float>& a, int64_t t, int64_t n,
                                           unsigned long long q) {
  const int g = __ldg(a.group + t);
  if (g >= 0 && g < a.n_groups && q != 0ULL)
    atomicAdd((unsigned long long*)a.out[0] + (int64_t)g * a.N + n, q);
}

// MOD16_OUT_SUM_I16: round(v / scale) saturated to +-32767; NaN -> -32768
__device__ __forceinline__ short to_i16(float v, float inv_scale) {
  if (v != v) return (short)-32768;
  int q;
  asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(q) : "f"(fmaxf(v * inv_scale, -32767.0f)));
  return (short)q;
}

template <int OUT>
__device__ __forceinline__ void store_exact(const ForwardArgs<float>& a, int64_t o,
                                            const fast::StepOut& r, const StepIn& in) {
  if (OUT == 3) {
    const float tcm = 0.5f * (fast::celsius<float>(in.t_d) + fast::celsius<float>(in.t_n));
    annual_add(a, o / a.N, o % a.N, annual_fixed(r.tot, tcm));
  } else if (OUT == 6) {
    ((short*)a.out[0])[o] = to_i16(r.tot, a.out_inv_scale);
  } else if (OUT == 5) {
    const int64_t t = o / a.N;
    a.out[0][o] = fmaf(r.tot_day, __ldg(a.w_day + t), r.tot_night * __ldg(a.w_night + t));
  } else if (OUT == 4) {
    a.out[0][o] = r.tot;        a.out[1][o] = r.day.pt;
    a.out[2][o] = r.day.sat;    a.out[3][o] = r.day.unsat;
    a.out[4][o] = r.night.sat;  a.out[5][o] = r.night.unsat;
  } else if (OUT == 0) {
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
  store_exact<OUT>(a, t * a.N + n, r, in);
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
  if (OUT != 2 && OUT != 4) {
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
  store_exact<OUT>(a, o, r, in);
  return g_pos ? 1 : 0;
}

}  // namespace mod16
