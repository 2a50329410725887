// K1 -- fused MOD16 forward model.  "Strided" kernels: one thread per pixel, a loop over
// a chunk of time steps, arbitrary broadcast strides, float or double, both semantics,
// op-for-op math (mod16_math.cuh).  This is the accuracy path (f64 within 1e-10 of the
// reference) and the fallback layout for the f32 path; the vectorised f32 fast path lives
// in mod16_forward_fast.cu.
#include "mod16_common.cuh"

namespace mod16 {

thread_local int g_last_cuda_error = 0;
thread_local int g_last_forward_kernel = 0;

template <typename R, int SEM, int OUT>
__global__ void __launch_bounds__(256)
forward_strided_kernel(const __grid_constant__ ForwardArgs<R> a) {
  if (a.skip_if_flag && *((volatile int*)a.gate_flag) != 0) return;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t t0 = (int64_t)blockIdx.y * a.t_chunk;
  int64_t t1 = t0 + a.t_chunk;
  if (t1 > a.T) t1 = a.T;
  bool any_pos = false;
  if (n < a.N) {
    const Params<R> q = load_params<R, SEM>(a, n, a.table);
    for (int64_t t = t0; t < t1; ++t) {
      Drivers<R> d;
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
      d.rc_d = a.has_rc ? a.rc[0].at(t, n) : R(0);
      d.rc_n = a.has_rc ? a.rc[1].at(t, n) : R(0);
      Half<R> day, night;
      bool g_pos;
      pixel_step<R, SEM>(d, q, a.has_rc != 0, a.gate, day, night, g_pos);
      any_pos |= g_pos;
      const int64_t o = t * a.N + n;
      if (OUT == 0) {
        a.out[0][o] = half_total<R, SEM>(day) + half_total<R, SEM>(night);  // :193
      } else if (OUT == 1) {
        a.out[0][o] = half_total<R, SEM>(day);
        a.out[1][o] = half_total<R, SEM>(night);
      } else {
        a.out[0][o] = day.canopy;   a.out[1][o] = day.soil;   a.out[2][o] = day.trans;
        a.out[3][o] = night.canopy; a.out[4][o] = night.soil; a.out[5][o] = night.trans;
      }
    }
  }
  if (a.track_gate) {
    if (__any_sync(0xffffffffu, any_pos) && (threadIdx.x & 31) == 0) *a.gate_flag = 1;
  }
}

// Validates the ABI structs and fills the kernel argument block (shared with the fast path).
template <typename R>
int build_forward_args(const mod16_drivers_t* drv, const mod16_params_t* par,
                       const mod16_out_t* out, int64_t T, int64_t N, uint32_t flags,
                       ForwardArgs<R>& a) {
  if (!drv || !par || !out) return MOD16_ERR_NULL;
  if (T < 0 || N < 0) return MOD16_ERR_SHAPE;
  const uint32_t sem = flags & MOD16_SEM_MASK, om = flags & MOD16_OUT_MASK,
                 gm = flags & MOD16_GATE_MASK;
  if (sem > MOD16_SEM_INSTANCE) return MOD16_ERR_FLAGS;
  if (om != MOD16_OUT_SUM && om != MOD16_OUT_DAY_NIGHT && om != MOD16_OUT_COMPONENTS &&
      om != MOD16_OUT_ANNUAL && om != MOD16_OUT_POTENTIAL && om != MOD16_OUT_COMPOSITE &&
      om != MOD16_OUT_SUM_I16)
    return MOD16_ERR_FLAGS;
  if ((om == MOD16_OUT_POTENTIAL || om == MOD16_OUT_COMPOSITE || om == MOD16_OUT_SUM_I16) &&
      sizeof(R) != sizeof(float))
    return MOD16_ERR_FLAGS;
  if (om == MOD16_OUT_SUM_I16 && (!(out->out_scale > 0.0) || !out->gate_flag))
    return MOD16_ERR_FLAGS;
  a.out_inv_scale = om == MOD16_OUT_SUM_I16 ? (float)(1.0 / out->out_scale) : 0.0f;
  if (om == MOD16_OUT_COMPOSITE && T > 0 && (!out->day_weight || !out->night_weight))
    return MOD16_ERR_NULL;
  a.w_day = out->day_weight;
  a.w_night = out->night_weight;
  if (om == MOD16_OUT_ANNUAL) {   // float32 static path; needs the group index of every step
    if (sizeof(R) != sizeof(float) || sem != MOD16_SEM_STATIC) return MOD16_ERR_FLAGS;
    if ((!out->group && T > 0) || out->n_groups < 1 || !out->gate_flag) return MOD16_ERR_NULL;
  }
  a.group = out->group;
  a.n_groups = out->n_groups;
  if (gm > MOD16_GATE_FIXUP) return MOD16_ERR_FLAGS;
  for (int k = 0; k < MOD16_N_DRIVERS; ++k) {
    if (drv->drv[k].t_stride < 0 || drv->drv[k].n_stride < 0) return MOD16_ERR_SHAPE;
    fill_arr<R>(a.drv[k], drv->drv[k]);
    if (sizeof(R) == 8 && (drv->f32_mask >> k & 1u) && drv->drv[k].ptr) a.drv[k].is_f32 = 1;
  }
  if (sizeof(R) != 8 && drv->f32_mask != 0) return MOD16_ERR_FLAGS;
  a.has_rc = drv->has_r_corr ? 1 : 0;
  for (int k = 0; k < 2; ++k) {
    if (a.has_rc) {
      if (drv->r_corr[k].t_stride < 0 || drv->r_corr[k].n_stride < 0) return MOD16_ERR_SHAPE;
      fill_arr<R>(a.rc[k], drv->r_corr[k]);
    } else {
      a.rc[k].ptr = nullptr; a.rc[k].ts = a.rc[k].ns = 0; a.rc[k].scalar = R(0);
      a.rc[k].is_f32 = 0;
    }
  }
  int n_arr = 0;
  for (int k = 0; k < MOD16_N_PARAMS; ++k) n_arr += par->arrays[k] != nullptr;
  if (n_arr != 0 && n_arr != MOD16_N_PARAMS) return MOD16_ERR_PARAMS;
  if (n_arr && par->pft) return MOD16_ERR_PARAMS;
  a.par_mode = n_arr ? 2 : (par->pft ? 1 : 0);
  a.pft = par->pft;
  for (int k = 0; k < MOD16_N_PARAMS; ++k) {
    a.par_arr[k] = (const R*)par->arrays[k];
    a.par_stride[k] = par->array_stride[k];
    if (n_arr && (a.par_stride[k] < 0 || a.par_stride[k] > 1)) return MOD16_ERR_SHAPE;
    for (int c = 0; c < MOD16_N_PFT; ++c) a.table[k][c] = par->table[k][c];
  }
  const int n_out = (om == MOD16_OUT_SUM || om == MOD16_OUT_ANNUAL || om == MOD16_OUT_COMPOSITE ||
                     om == MOD16_OUT_SUM_I16) ? 1
                    : (om == MOD16_OUT_DAY_NIGHT ? 2 : 6);
  for (int k = 0; k < 6; ++k) {
    a.out[k] = k < n_out ? (R*)out->out[k] : nullptr;
    if (k < n_out && !a.out[k] && T * N > 0) return MOD16_ERR_NULL;
  }
  a.gate_flag = out->gate_flag;
  a.redo_ws = nullptr;
  a.redo_cap = 0;
  if (out->workspace && out->workspace_bytes >= 16 &&
      (reinterpret_cast<uintptr_t>(out->workspace) & 7) == 0) {
    a.redo_ws = (unsigned long long*)out->workspace;
    a.redo_cap = out->workspace_bytes / 8 - 1;
  }
  const bool needs_flag = sem == MOD16_SEM_STATIC &&
                          (gm == MOD16_GATE_GLOBAL || gm == MOD16_GATE_OPEN_TRACK ||
                           gm == MOD16_GATE_FIXUP);
  if (needs_flag && !a.gate_flag) return MOD16_ERR_NULL;
  a.skip_if_flag = 0;
  a.track_gate = 0;
  a.gate = GATE_OPEN;
  a.T = T;
  a.N = N;
  a.t_chunk = 1;
  a.raw_fmt = 0;
  a.prep = 0;
  for (int k = 0; k < 2; ++k) {
    a.ps[k].ptr = nullptr; a.ps[k].ts = a.ps[k].ns = 0; a.ps[k].scalar = R(0); a.ps[k].is_f32 = 0;
  }
  for (int k = 0; k < MOD16_N_DRIVERS; ++k) {
    a.raw[k].dtype = MOD16_T_F32; a.raw[k].has_fill = 0;
    a.raw[k].scale = 1.0f; a.raw[k].offset = 0.0f; a.raw[k].fill = 0;
  }
  return MOD16_OK;
}

template <typename R, int SEM>
static int launch_strided_out(const ForwardArgs<R>& a, uint32_t om, dim3 grid, dim3 block,
                              cudaStream_t st) {
  if (om == MOD16_OUT_SUM)
    forward_strided_kernel<R, SEM, 0><<<grid, block, 0, st>>>(a);
  else if (om == MOD16_OUT_DAY_NIGHT)
    forward_strided_kernel<R, SEM, 1><<<grid, block, 0, st>>>(a);
  else
    forward_strided_kernel<R, SEM, 2><<<grid, block, 0, st>>>(a);
  M16_CUDA_CHECK(cudaGetLastError());
  return MOD16_OK;
}

// One pass of the strided kernel with the gate settings already placed in `a`.
template <typename R>
int launch_strided(ForwardArgs<R>& a, uint32_t flags, cudaStream_t st) {
  g_last_forward_kernel = MOD16_KERNEL_EXACT;
  if (a.T == 0 || a.N == 0) return MOD16_OK;
  const int threads = 256;
  const int64_t bx = (a.N + threads - 1) / threads;
  // enough block rows to fill 148 SMs x 8 resident blocks a few times over, while each
  // thread still amortises its parameter gather over several time steps
  int64_t want_rows = (148 * 8 * 4 + bx - 1) / bx;
  if (want_rows < 1) want_rows = 1;
  int64_t chunk = (a.T + want_rows - 1) / want_rows;
  if (chunk < 1) chunk = 1;
  if (chunk > 64) chunk = 64;
  int64_t by = (a.T + chunk - 1) / chunk;
  while (by > 65535) { chunk *= 2; by = (a.T + chunk - 1) / chunk; }
  if (bx > 2147483647LL) return MOD16_ERR_SHAPE;
  a.t_chunk = (int)chunk;
  dim3 grid((unsigned)bx, (unsigned)by), block(threads);
  const uint32_t om = flags & MOD16_OUT_MASK;
  if ((flags & MOD16_SEM_MASK) == MOD16_SEM_STATIC)
    return launch_strided_out<R, SEM_STATIC>(a, om, grid, block, st);
  return launch_strided_out<R, SEM_INSTANCE>(a, om, grid, block, st);
}

// Gate protocol shared by both precisions.  `pass` launches one kernel pass.
template <typename R, typename PassFn>
int run_with_gate(ForwardArgs<R>& a, uint32_t flags, cudaStream_t st, PassFn pass) {
  const uint32_t gm = flags & MOD16_GATE_MASK;
  const bool is_static = (flags & MOD16_SEM_MASK) == MOD16_SEM_STATIC;
  if (!is_static) {  // the instance path has no gate (:1234-1238)
    a.gate = GATE_OPEN;
    return pass(a);
  }
  int rc;
  switch (gm) {
    case MOD16_GATE_ELEMENT: a.gate = GATE_ELEMENT; return pass(a);
    case MOD16_GATE_OPEN: a.gate = GATE_OPEN; return pass(a);
    case MOD16_GATE_CLOSED: a.gate = GATE_CLOSED; return pass(a);
    case MOD16_GATE_OPEN_TRACK:
      a.gate = GATE_OPEN; a.track_gate = 1;
      return pass(a);
    case MOD16_GATE_FIXUP:
      a.gate = GATE_CLOSED; a.skip_if_flag = 1;
      return pass(a);
    case MOD16_GATE_GLOBAL:
      M16_CUDA_CHECK(cudaMemsetAsync(a.gate_flag, 0, sizeof(int), st));
      a.gate = GATE_OPEN; a.track_gate = 1; a.skip_if_flag = 0;
      rc = pass(a);
      if (rc != MOD16_OK) return rc;
      a.gate = GATE_CLOSED; a.track_gate = 0; a.skip_if_flag = 1;
      return pass(a);
  }
  return MOD16_ERR_FLAGS;
}

// defined in mod16_forward_fast.cu / mod16_forward_fast64.cu
int launch_fast_f32(ForwardArgs<float>& a, uint32_t flags, cudaStream_t st);
int launch_fast_f64(ForwardArgs<double>& a, uint32_t flags, cudaStream_t st);
// storage formats of mod16_forward_raw_f32 (mod16_forward_pipe.cuh)
constexpr int FMT_F32_PREP = 1, FMT_MODIS = 2, FMT_PACKED = 3;

template int build_forward_args<float>(const mod16_drivers_t*, const mod16_params_t*,
                                       const mod16_out_t*, int64_t, int64_t, uint32_t,
                                       ForwardArgs<float>&);
template int build_forward_args<double>(const mod16_drivers_t*, const mod16_params_t*,
                                        const mod16_out_t*, int64_t, int64_t, uint32_t,
                                        ForwardArgs<double>&);

}  // namespace mod16

extern "C" {

int mod16_abi_version(void) { return MOD16_ABI_VERSION; }
int64_t mod16_workspace_bytes(void) { return (int64_t)8 << 20; }
int mod16_last_cuda_error(void) { return mod16::g_last_cuda_error; }
int mod16_last_forward_kernel(void) { return mod16::g_last_forward_kernel; }

const char* mod16_error_string(int code) {
  switch (code) {
    case MOD16_OK: return "ok";
    case MOD16_ERR_NULL: return "a required pointer is NULL";
    case MOD16_ERR_SHAPE: return "bad shape or stride";
    case MOD16_ERR_FLAGS: return "unknown or inconsistent flags";
    case MOD16_ERR_CUDA: return "CUDA call failed (see mod16_last_cuda_error)";
    case MOD16_ERR_LAYOUT: return "storage-format layout not supported by the fused kernel";
    case MOD16_ERR_PARAMS: return "inconsistent parameter source";
    case MOD16_ERR_WORKSPACE: return "workspace too small";
  }
  return "unknown error";
}

int mod16_forward_f64(const mod16_drivers_t* drivers, const mod16_params_t* params,
                      const mod16_out_t* out, int64_t T, int64_t N, uint32_t flags,
                      void* stream) {
  mod16::ForwardArgs<double> a;
  int rc = mod16::build_forward_args<double>(drivers, params, out, T, N, flags, a);
  if (rc != MOD16_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool fast = !(flags & MOD16_FORCE_MIRROR);
  return mod16::run_with_gate<double>(a, flags, st, [&](mod16::ForwardArgs<double>& x) {
    return fast ? mod16::launch_fast_f64(x, flags, st)
                : mod16::launch_strided<double>(x, flags, st);
  });
}

int mod16_forward_f32(const mod16_drivers_t* drivers, const mod16_params_t* params,
                      const mod16_out_t* out, int64_t T, int64_t N, uint32_t flags,
                      void* stream) {
  mod16::ForwardArgs<float> a;
  int rc = mod16::build_forward_args<float>(drivers, params, out, T, N, flags, a);
  if (rc != MOD16_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool fast = !(flags & MOD16_FORCE_MIRROR);  // every layout has a fast-math kernel
  return mod16::run_with_gate<float>(a, flags, st, [&](mod16::ForwardArgs<float>& x) {
    return fast ? mod16::launch_fast_f32(x, flags, st)
                : mod16::launch_strided<float>(x, flags, st);
  });
}

int mod16_forward_raw_f32(const mod16_raw_drivers_t* drivers, const mod16_params_t* params,
                          const mod16_out_t* out, int64_t T, int64_t N, uint32_t flags,
                          void* stream) {
  if (!drivers) return MOD16_ERR_NULL;
  if (flags & (MOD16_FORCE_MIRROR | MOD16_FORCE_GENERIC)) return MOD16_ERR_FLAGS;
  // the float view of the drivers (pointers and strides; element types follow below)
  mod16_drivers_t plain;
  for (int k = 0; k < MOD16_N_DRIVERS; ++k) {
    const mod16_raw_driver_t& r = drivers->drv[k];
    if (r.dtype != MOD16_T_F32 && r.dtype != MOD16_T_I16 && r.dtype != MOD16_T_U8)
      return MOD16_ERR_FLAGS;
    plain.drv[k].ptr = r.ptr;
    plain.drv[k].t_stride = r.t_stride;
    plain.drv[k].n_stride = r.n_stride;
    plain.drv[k].scalar = r.scalar;
  }
  plain.has_r_corr = 0;
  plain.f32_mask = 0;
  for (int k = 0; k < 2; ++k) {
    plain.r_corr[k].ptr = nullptr; plain.r_corr[k].t_stride = plain.r_corr[k].n_stride = 0;
    plain.r_corr[k].scalar = 0.0;
  }
  mod16::ForwardArgs<float> a;
  int rc = mod16::build_forward_args<float>(&plain, params, out, T, N, flags, a);
  if (rc != MOD16_OK) return rc;
  // layout: all float32 (+ conversions), MODIS or PACKED (see the header)
  bool any_int = false, met_int = false;
  for (int k = 0; k < MOD16_N_DRIVERS; ++k) {
    const mod16_raw_driver_t& r = drivers->drv[k];
    a.raw[k].dtype = r.ptr ? r.dtype : MOD16_T_F32;
    a.raw[k].has_fill = r.has_fill != 0 && a.raw[k].dtype != MOD16_T_F32;
    a.raw[k].scale = (float)r.scale;
    a.raw[k].offset = (float)r.offset;
    // (no fill value: one no int16 / uint8 element can equal)
    a.raw[k].fill = a.raw[k].has_fill ? (int)r.fill : (int)0x80000000;
    if (a.raw[k].dtype != MOD16_T_F32) {
      any_int = true;
      if (k != MOD16_D_SW_ALBEDO && k != MOD16_D_FPAR && k != MOD16_D_LAI) met_int = true;
    }
  }
  a.prep = drivers->prep;
  if (a.prep & ~(MOD16_PREP_VPD_NIGHT_CLAMP | MOD16_PREP_ELEVATION | MOD16_PREP_VPD_FROM_QV))
    return MOD16_ERR_FLAGS;
  if (a.prep & MOD16_PREP_VPD_FROM_QV) {
    if (any_int) return MOD16_ERR_LAYOUT;   // (float32 arrays only)
    for (int k = 0; k < 2; ++k) {
      const mod16_raw_driver_t& r = drivers->ps[k];
      if (r.dtype != MOD16_T_F32 || r.t_stride < 0 || r.n_stride < 0) return MOD16_ERR_LAYOUT;
      mod16_driver_t d;
      d.ptr = r.ptr; d.t_stride = r.t_stride; d.n_stride = r.n_stride; d.scalar = r.scalar;
      mod16::fill_arr<float>(a.ps[k], d);
    }
  }
  a.raw_fmt = !any_int ? mod16::FMT_F32_PREP : (met_int ? mod16::FMT_PACKED : mod16::FMT_MODIS);
  cudaStream_t st = (cudaStream_t)stream;
  return mod16::run_with_gate<float>(a, flags, st, [&](mod16::ForwardArgs<float>& x) {
    return mod16::launch_fast_f32(x, flags, st);
  });
}

}  // extern "C"
