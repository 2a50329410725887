// K1, float64 entry, fast path: the rearranged algebra of mod16_fast.cuh evaluated in double.
//
// The op-for-op transcription (forward_strided_kernel<double>, mod16_math.cuh) spends its time
// in ~66 IEEE divisions and 14 libdevice exp / pow calls per pixel-timestep (FP64-issue-bound:
// 4.3 G units/s on the tile).  Multiplying every Penman-Monteith ratio through by the common
// factor leaves 7 divisions, 4 exp2, 2 log2 and 4 square roots; in double the rearrangement
// changes results by a few 1e-15 relative (well inside the 1e-10 the float64 path promises),
// except next to the model's discontinuities, where -- exactly like the float32 path, with
// guards five orders of magnitude tighter -- the element is handed to the op-for-op code, as
// is everything outside the physical domain (NaN / fill values, degenerate parameters).
// One thread per pixel walks a chunk of time steps; any broadcast layout.
#include "mod16_common.cuh"
#include "mod16_fast.cuh"

namespace mod16 {

template <int OUT>
__device__ __forceinline__ void store64(const ForwardArgs<double>& a, int64_t o, double dc,
                                        double ds, double dt, double nc, double ns, double nt,
                                        double tot_d, double tot_n) {
  if (OUT == 0) {
    a.out[0][o] = tot_d + tot_n;          // :193
  } else if (OUT == 1) {
    a.out[0][o] = tot_d;
    a.out[1][o] = tot_n;
  } else {
    a.out[0][o] = dc; a.out[1][o] = ds; a.out[2][o] = dt;
    a.out[3][o] = nc; a.out[4][o] = ns; a.out[5][o] = nt;
  }
}

// The op-for-op evaluation of ONE element, out of line (its registers are not the loop's).
template <int SEM, int OUT>
__device__ __noinline__ bool exact64_one(const ForwardArgs<double>& a, const Drivers<double>& d,
                                         const Params<double>& q, int64_t o) {
  Half<double> day, night;
  bool g_pos;
  pixel_step<double, SEM>(d, q, a.has_rc != 0, a.gate, day, night, g_pos);
  store64<OUT>(a, o, day.canopy, day.soil, day.trans, night.canopy, night.soil, night.trans,
               half_total<double, SEM>(day), half_total<double, SEM>(night));
  return g_pos;
}

// The masked-driver rule (fast::nan_halves, see mod16_fast.cuh) for ONE element of the float64
// entry: a NaN (or negative-temperature) driver that makes the total NaN whatever else the
// element holds is answered without the op-for-op code -- masked pixels of a real tile are
// masked at every step.  Returns true when the outputs were written (NaN totals).
template <int SEM, int OUT>
__device__ __noinline__ bool rule64_one(const ForwardArgs<double>& a, const Drivers<double>& d,
                                        const Params<double>& q, int64_t o, bool& g_pos) {
  if (OUT == 2) return false;
  fast::StepIn f;
  f.lw_d = (float)d.lw_d; f.lw_n = (float)d.lw_n; f.sw_d = (float)d.sw_d; f.sw_n = (float)d.sw_n;
  f.alb = (float)d.alb; f.t_d = (float)d.t_d; f.t_n = (float)d.t_n; f.t_ann = 0.0f;
  f.tmin = (float)d.tmin; f.vpd_d = (float)d.vpd_d; f.vpd_n = (float)d.vpd_n; f.p = (float)d.p;
  f.fpar = (float)d.fpar; f.lai = (float)d.lai; f.rc_d = 0.0f; f.rc_n = 0.0f;
  const double col[MOD16_N_PARAMS] = {q.tmin_close, q.tmin_open, q.vpd_open, q.vpd_close,
                                      q.gl_sh,      q.gl_wv,     q.g_cuticular, q.csl,
                                      q.rbl_min,    q.rbl_max,   q.beta};
  const fast::PftConst pc = fast::make_pft_const(col, SEM);
  const int h = fast::nan_halves<SEM>(f, pc, a.has_rc != 0, g_pos);
  if (!(OUT == 0 ? h != 0 : h == 3)) return false;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  a.out[0][o] = qnan;
  if (OUT == 1) a.out[1][o] = qnan;
  return true;
}

template <int SEM, int OUT>
__global__ void __launch_bounds__(128)
forward_fast64_kernel(const __grid_constant__ ForwardArgs<double> a) {
  if (a.skip_if_flag && *((volatile int*)a.gate_flag) != 0) return;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t t0 = (int64_t)blockIdx.y * a.t_chunk;
  int64_t t1 = t0 + a.t_chunk;
  if (t1 > a.T) t1 = a.T;
  bool any_pos = false;
  if (n < a.N) {
    const Params<double> q = load_params<double, SEM>(a, n, a.table);
    fast::PftConstV<double> pc;
    {
      const double col[MOD16_N_PARAMS] = {q.tmin_close, q.tmin_open, q.vpd_open, q.vpd_close,
                                          q.gl_sh,      q.gl_wv,     q.g_cuticular, q.csl,
                                          q.rbl_min,    q.rbl_max,   q.beta};
      fast::fill_pft_const<double>(pc, col);
    }
    const bool pix_static = a.drv[MOD16_D_PRESSURE].ts == 0 && a.drv[MOD16_D_TEMP_ANNUAL].ts == 0;
    fast::PixConstV<double> px = fast::make_pix_const_s<double>(
        a.drv[MOD16_D_PRESSURE].at(t0, n), a.drv[MOD16_D_TEMP_ANNUAL].at(t0, n), q.tann_lo,
        pc.ok != 0.0, SEM);
    const bool has_rc = a.has_rc != 0;
    for (int64_t t = t0; t < t1; ++t) {
      Drivers<double> d;
      d.lw_d = a.drv[MOD16_D_LW_NET_DAY].at(t, n);
      d.lw_n = a.drv[MOD16_D_LW_NET_NIGHT].at(t, n);
      d.sw_d = a.drv[MOD16_D_SW_RAD_DAY].at(t, n);
      d.sw_n = a.drv[MOD16_D_SW_RAD_NIGHT].at(t, n);
      d.alb = a.drv[MOD16_D_SW_ALBEDO].at(t, n);
      d.t_d = a.drv[MOD16_D_TEMP_DAY].at(t, n);
      d.t_n = a.drv[MOD16_D_TEMP_NIGHT].at(t, n);
      d.t_ann = a.drv[MOD16_D_TEMP_ANNUAL].at(t, n);
      d.tmin = a.drv[MOD16_D_TMIN].at(t, n);
      d.vpd_d = a.drv[MOD16_D_VPD_DAY].at(t, n);
      d.vpd_n = a.drv[MOD16_D_VPD_NIGHT].at(t, n);
      d.p = a.drv[MOD16_D_PRESSURE].at(t, n);
      d.fpar = a.drv[MOD16_D_FPAR].at(t, n);
      d.lai = a.drv[MOD16_D_LAI].at(t, n);
      d.rc_d = has_rc ? a.rc[0].at(t, n) : 0.0;
      d.rc_n = has_rc ? a.rc[1].at(t, n) : 0.0;
      if (!pix_static)
        px = fast::make_pix_const_s<double>(d.p, d.t_ann, q.tann_lo, pc.ok != 0.0, SEM);
      fast::StepInV<double> in;
      in.lw_d = d.lw_d; in.lw_n = d.lw_n; in.sw_d = d.sw_d; in.sw_n = d.sw_n; in.alb = d.alb;
      in.t_d = d.t_d; in.t_n = d.t_n; in.tmin = d.tmin; in.vpd_d = d.vpd_d; in.vpd_n = d.vpd_n;
      in.fpar = d.fpar; in.lai = d.lai; in.rc_d = d.rc_d; in.rc_n = d.rc_n;
      fast::StepOutV<double> r;
      bool g_pos;
      const bool need = has_rc ? fast::step_fast<double, SEM, true>(in, pc, px, a.gate, r, g_pos)
                               : fast::step_fast<double, SEM, false>(in, pc, px, a.gate, r, g_pos);
      const int64_t o = t * a.N + n;
      if (need) {
        bool gp = false;
        if (rule64_one<SEM, OUT>(a, d, q, o, gp)) any_pos |= gp;
        else any_pos |= exact64_one<SEM, OUT>(a, d, q, o);
      } else {
        any_pos |= g_pos;
        store64<OUT>(a, o, r.day.canopy, r.day.soil, r.day.trans, r.night.canopy, r.night.soil,
                     r.night.trans, r.tot_day, r.tot_night);
      }
    }
  }
  if (a.track_gate) {
    if (__any_sync(0xffffffffu, any_pos) && (threadIdx.x & 31) == 0) *a.gate_flag = 1;
  }
}

template <int SEM>
static int launch64_sem(const ForwardArgs<double>& a, uint32_t om, dim3 grid, dim3 block,
                        cudaStream_t st) {
  if (om == MOD16_OUT_SUM) forward_fast64_kernel<SEM, 0><<<grid, block, 0, st>>>(a);
  else if (om == MOD16_OUT_DAY_NIGHT) forward_fast64_kernel<SEM, 1><<<grid, block, 0, st>>>(a);
  else forward_fast64_kernel<SEM, 2><<<grid, block, 0, st>>>(a);
  M16_CUDA_CHECK(cudaGetLastError());
  return MOD16_OK;
}

int launch_fast_f64(ForwardArgs<double>& a, uint32_t flags, cudaStream_t st) {
  g_last_forward_kernel = MOD16_KERNEL_FAST64;
  if (a.T == 0 || a.N == 0) return MOD16_OK;
  const int threads = 128;
  const int64_t bx = (a.N + threads - 1) / threads;
  if (bx > 2147483647LL) return MOD16_ERR_SHAPE;
  // block rows: several waves of 148 SMs x resident blocks, each thread amortising its
  // per-pixel constants over a run of time steps
  int64_t want_rows = (148 * 8 * 4 + bx - 1) / bx;
  if (want_rows < 1) want_rows = 1;
  int64_t chunk = (a.T + want_rows - 1) / want_rows;
  if (chunk < 1) chunk = 1;
  if (chunk > 64) chunk = 64;
  int64_t by = (a.T + chunk - 1) / chunk;
  while (by > 65535) { chunk *= 2; by = (a.T + chunk - 1) / chunk; }
  a.t_chunk = (int)chunk;
  dim3 grid((unsigned)bx, (unsigned)by), block(threads);
  const uint32_t om = flags & MOD16_OUT_MASK;
  if ((flags & MOD16_SEM_MASK) == MOD16_SEM_STATIC)
    return launch64_sem<SEM_STATIC>(a, om, grid, block, st);
  return launch64_sem<SEM_INSTANCE>(a, om, grid, block, st);
}

}  // namespace mod16
