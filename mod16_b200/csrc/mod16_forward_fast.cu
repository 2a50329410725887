// K1, float32 fast path.
//
// forward_vec4_kernel -- the headline kernel.  Layout: (T, N) row-major structure-of-arrays,
//   one thread owns 4 consecutive pixels and walks a chunk of time steps.  Per step it
//   issues eleven 128-bit streaming loads (one per per-step driver; a warp reads 512
//   contiguous bytes of each array), evaluates the 4 pixel-timesteps with the rearranged
//   math of mod16_fast.cuh and writes ET once with a 128-bit streaming store.  The BPLUT
//   (13 land-cover columns x 16 derived constants) is staged in shared memory once per
//   block and gathered by the pixel's land-cover code; pressure terms and the annual
//   temperature window are hoisted out of the time loop.  Algorithmic traffic:
//   11 x 4 B in + 4 B out = 48 B per pixel-timestep (OUT_SUM).
//
// forward_strided_fast_kernel -- same math for layouts that do not allow 128-bit loads
//   (broadcast strides, per-pixel parameter arrays, N % 4 != 0, r_corr_list).
#include "mod16_common.cuh"
#include "mod16_fast.cuh"

namespace mod16 {

using fast::PftConst;
using fast::PixConst;
using fast::StepIn;

__device__ __forceinline__ float4 ld_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float comp(const float4& v, int j) {
  return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
}
__device__ __forceinline__ void setc(float4& v, int j, float x) {
  if (j == 0) v.x = x; else if (j == 1) v.y = x; else if (j == 2) v.z = x; else v.w = x;
}

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

template <int SEM, int OUT>
__global__ void __launch_bounds__(256, 2)
forward_vec4_kernel(const __grid_constant__ ForwardArgs<float> a) {
  __shared__ SharedLut lut;
  if (a.skip_if_flag && *((volatile int*)a.gate_flag) != 0) return;
  stage_lut<SEM>(a, lut);
  const int64_t n0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int64_t t0 = (int64_t)blockIdx.y * a.t_chunk;
  int64_t t1 = t0 + a.t_chunk;
  if (t1 > a.T) t1 = a.T;
  bool any_pos = false;
  if (n0 < a.N) {
    // per-pixel state: land-cover codes, pressure terms, annual-temperature window
    int code[4] = {0, 0, 0, 0};
    if (a.par_mode == 1) {
      const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(a.pft + n0));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = (w >> (8 * j)) & 0xff;
        code[j] = c < MOD16_N_PFT ? c : MOD16_N_PFT;  // out of range: the all-NaN column
      }
    }
    const Arr<float>& ap = a.drv[MOD16_D_PRESSURE];
    const Arr<float>& aa = a.drv[MOD16_D_TEMP_ANNUAL];
    float4 p4, ta4;
    if (ap.ptr) p4 = __ldg(reinterpret_cast<const float4*>(ap.ptr + n0));
    else p4 = make_float4(ap.scalar, ap.scalar, ap.scalar, ap.scalar);
    if (aa.ptr) ta4 = __ldg(reinterpret_cast<const float4*>(aa.ptr + n0));
    else ta4 = make_float4(aa.scalar, aa.scalar, aa.scalar, aa.scalar);
    PixConst px[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      px[j] = fast::make_pix_const(comp(p4, j), comp(ta4, j), lut.pc[code[j]], SEM);

    const Arr<float>& swn = a.drv[MOD16_D_SW_RAD_NIGHT];
    for (int64_t t = t0; t < t1; ++t) {
      const int64_t off = t * a.N + n0;  // per-step arrays are dense (checked on the host)
      const float4 lw_d = ld_stream(a.drv[MOD16_D_LW_NET_DAY].ptr + off);
      const float4 lw_n = ld_stream(a.drv[MOD16_D_LW_NET_NIGHT].ptr + off);
      const float4 sw_d = ld_stream(a.drv[MOD16_D_SW_RAD_DAY].ptr + off);
      const float4 alb = ld_stream(a.drv[MOD16_D_SW_ALBEDO].ptr + off);
      const float4 t_d = ld_stream(a.drv[MOD16_D_TEMP_DAY].ptr + off);
      const float4 t_n = ld_stream(a.drv[MOD16_D_TEMP_NIGHT].ptr + off);
      const float4 tmin = ld_stream(a.drv[MOD16_D_TMIN].ptr + off);
      const float4 vpd_d = ld_stream(a.drv[MOD16_D_VPD_DAY].ptr + off);
      const float4 vpd_n = ld_stream(a.drv[MOD16_D_VPD_NIGHT].ptr + off);
      const float4 fpar = ld_stream(a.drv[MOD16_D_FPAR].ptr + off);
      const float4 lai = ld_stream(a.drv[MOD16_D_LAI].ptr + off);
      float4 sw_n;
      if (swn.ptr) sw_n = ld_stream(swn.ptr + off);
      else sw_n = make_float4(swn.scalar, swn.scalar, swn.scalar, swn.scalar);
      float4 o0, o1, o2, o3, o4, o5;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        StepIn in;
        in.lw_d = comp(lw_d, j); in.lw_n = comp(lw_n, j); in.sw_d = comp(sw_d, j);
        in.sw_n = comp(sw_n, j); in.alb = comp(alb, j);   in.t_d = comp(t_d, j);
        in.t_n = comp(t_n, j);   in.t_ann = comp(ta4, j); in.tmin = comp(tmin, j);
        in.vpd_d = comp(vpd_d, j); in.vpd_n = comp(vpd_n, j); in.p = comp(p4, j);
        in.fpar = comp(fpar, j); in.lai = comp(lai, j);
        in.rc_d = 0.0f; in.rc_n = 0.0f;
        fast::StepOut r;
        bool g_pos;
        const int cj = code[j];
        fast::pixel_step_fast<SEM>(
            in, lut.pc[cj], px[j], [&](int k) { return lut.tab[k][cj]; }, false, a.gate, r,
            g_pos);
        any_pos |= g_pos;
        if (OUT == 0) {
          setc(o0, j, r.tot);
        } else if (OUT == 1) {
          setc(o0, j, r.tot_day);
          setc(o1, j, r.tot_night);
        } else {
          setc(o0, j, r.day.canopy);   setc(o1, j, r.day.soil);   setc(o2, j, r.day.trans);
          setc(o3, j, r.night.canopy); setc(o4, j, r.night.soil); setc(o5, j, r.night.trans);
        }
      }
      st_stream(a.out[0] + off, o0);
      if (OUT >= 1) st_stream(a.out[1] + off, o1);
      if (OUT == 2) {
        st_stream(a.out[2] + off, o2); st_stream(a.out[3] + off, o3);
        st_stream(a.out[4] + off, o4); st_stream(a.out[5] + off, o5);
      }
    }
  }
  if (a.track_gate) {
    if (__any_sync(0xffffffffu, any_pos) && (threadIdx.x & 31) == 0) *a.gate_flag = 1;
  }
}

template <int SEM, int OUT>
__global__ void __launch_bounds__(256)
forward_strided_fast_kernel(const __grid_constant__ ForwardArgs<float> a) {
  if (a.skip_if_flag && *((volatile int*)a.gate_flag) != 0) return;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t t0 = (int64_t)blockIdx.y * a.t_chunk;
  int64_t t1 = t0 + a.t_chunk;
  if (t1 > a.T) t1 = a.T;
  bool any_pos = false;
  if (n < a.N) {
    double col[MOD16_N_PARAMS];
    if (a.par_mode == 2) {
#pragma unroll
      for (int k = 0; k < MOD16_N_PARAMS; ++k)
        col[k] = (double)__ldg(a.par_arr[k] + n * a.par_stride[k]);
    } else {
      int code = 0;
      if (a.par_mode == 1) {
        code = (int)__ldg(a.pft + n);
      }
#pragma unroll
      for (int k = 0; k < MOD16_N_PARAMS; ++k)
        col[k] = code < MOD16_N_PFT ? a.table[k][code]
                                    : __longlong_as_double(0x7ff8000000000000LL);
    }
    const PftConst pc = fast::make_pft_const(col, SEM);
    const bool pix_static = a.drv[MOD16_D_PRESSURE].ts == 0 &&
                            a.drv[MOD16_D_TEMP_ANNUAL].ts == 0;
    PixConst px = fast::make_pix_const(a.drv[MOD16_D_PRESSURE].at(t0, n),
                                       a.drv[MOD16_D_TEMP_ANNUAL].at(t0, n), pc, SEM);
    for (int64_t t = t0; t < t1; ++t) {
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
      if (!pix_static) px = fast::make_pix_const(in.p, in.t_ann, pc, SEM);
      fast::StepOut r;
      bool g_pos;
      fast::pixel_step_fast<SEM>(
          in, pc, px, [&](int k) { return col[k]; }, a.has_rc != 0, a.gate, r, g_pos);
      any_pos |= g_pos;
      const int64_t o = t * a.N + n;
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
  }
  if (a.track_gate) {
    if (__any_sync(0xffffffffu, any_pos) && (threadIdx.x & 31) == 0) *a.gate_flag = 1;
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// The vectorised kernel needs: N % 4 == 0; the eleven per-step drivers dense (T, N) and
// 16-byte aligned; sw_rad_night a scalar or dense; pressure / temp_annual per pixel (or
// scalar); parameters from the table (uniform or gathered by land cover); no r_corr_list.
static bool vec4_ok(const ForwardArgs<float>& a) {
  if (a.N % 4 != 0 || a.has_rc || a.par_mode == 2) return false;
  static const int per_step[] = {
      MOD16_D_LW_NET_DAY, MOD16_D_LW_NET_NIGHT, MOD16_D_SW_RAD_DAY, MOD16_D_SW_ALBEDO,
      MOD16_D_TEMP_DAY,   MOD16_D_TEMP_NIGHT,   MOD16_D_TMIN,       MOD16_D_VPD_DAY,
      MOD16_D_VPD_NIGHT,  MOD16_D_FPAR,         MOD16_D_LAI};
  for (int k : per_step) {
    const Arr<float>& d = a.drv[k];
    if (!d.ptr || d.ns != 1 || !aligned16(d.ptr)) return false;
    if (a.T > 1 && d.ts != a.N) return false;
  }
  const Arr<float>& sn = a.drv[MOD16_D_SW_RAD_NIGHT];
  if (sn.ptr && (sn.ns != 1 || !aligned16(sn.ptr) || (a.T > 1 && sn.ts != a.N))) return false;
  for (int k : {MOD16_D_PRESSURE, MOD16_D_TEMP_ANNUAL}) {
    const Arr<float>& d = a.drv[k];
    if (d.ptr && (d.ts != 0 || d.ns != 1 || !aligned16(d.ptr))) return false;
  }
  if (a.par_mode == 1 && (reinterpret_cast<uintptr_t>(a.pft) & 3) != 0) return false;
  for (int k = 0; k < 6; ++k)
    if (a.out[k] && !aligned16(a.out[k])) return false;
  return true;
}

bool fast_f32_applicable(const ForwardArgs<float>& a, uint32_t flags) {
  (void)a; (void)flags;
  return true;  // every layout has a fast-math kernel; vec4_ok() picks which
}

template <int SEM>
static int launch_sem(const ForwardArgs<float>& a, uint32_t om, bool vec, dim3 grid,
                      dim3 block, cudaStream_t st) {
#define M16_LAUNCH(K, O) K<SEM, O><<<grid, block, 0, st>>>(a)
  if (vec) {
    if (om == MOD16_OUT_SUM) M16_LAUNCH(forward_vec4_kernel, 0);
    else if (om == MOD16_OUT_DAY_NIGHT) M16_LAUNCH(forward_vec4_kernel, 1);
    else M16_LAUNCH(forward_vec4_kernel, 2);
  } else {
    if (om == MOD16_OUT_SUM) M16_LAUNCH(forward_strided_fast_kernel, 0);
    else if (om == MOD16_OUT_DAY_NIGHT) M16_LAUNCH(forward_strided_fast_kernel, 1);
    else M16_LAUNCH(forward_strided_fast_kernel, 2);
  }
#undef M16_LAUNCH
  M16_CUDA_CHECK(cudaGetLastError());
  return MOD16_OK;
}

int launch_fast_f32(ForwardArgs<float>& a, uint32_t flags, cudaStream_t st) {
  if (a.T == 0 || a.N == 0) return MOD16_OK;
  const bool vec = !(flags & MOD16_FORCE_GENERIC) && vec4_ok(a);
  const int threads = 256;
  const int64_t cols = vec ? a.N / 4 : a.N;
  const int64_t bx = (cols + threads - 1) / threads;
  if (bx > 2147483647LL) return MOD16_ERR_SHAPE;
  // Block rows: split time so that the grid is several waves of 148 SMs x resident blocks
  // while each thread amortises its per-pixel setup over a run of time steps.
  int64_t want_rows = (148 * 2 * 8 + bx - 1) / bx;
  if (want_rows < 1) want_rows = 1;
  int64_t chunk = (a.T + want_rows - 1) / want_rows;
  if (chunk < 1) chunk = 1;
  if (chunk > 32) chunk = 32;
  int64_t by = (a.T + chunk - 1) / chunk;
  while (by > 65535) { chunk *= 2; by = (a.T + chunk - 1) / chunk; }
  a.t_chunk = (int)chunk;
  dim3 grid((unsigned)bx, (unsigned)by), block(threads);
  const uint32_t om = flags & MOD16_OUT_MASK;
  if ((flags & MOD16_SEM_MASK) == MOD16_SEM_STATIC)
    return launch_sem<SEM_STATIC>(a, om, vec, grid, block, st);
  return launch_sem<SEM_INSTANCE>(a, om, vec, grid, block, st);
}

}  // namespace mod16
