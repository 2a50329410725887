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
#include <cuda_pipeline.h>
#include <cstdlib>
#include <type_traits>
#include "mod16_common.cuh"
#include "mod16_fast.cuh"

namespace mod16 {

constexpr int kSrc = MOD16_N_DRIVERS + 2;   // the sources of an element: 14 drivers + r_corr (day, night)

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

// ---- the rare elements, out of line -------------------------------------------------------
// Both helpers are self-contained: they re-read the element's inputs from global memory (the
// loop has just pulled them through L2), so the loop's drivers never need a home in local
// memory -- a by-reference struct did, and its stores serialised the loop's loads.
__device__ __forceinline__ Drivers<double> load_drivers64(const ForwardArgs<double>& a, int64_t t,
                                                          int64_t n) {
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
  d.rc_d = a.has_rc ? a.rc[0].at(t, n) : 0.0;
  d.rc_n = a.has_rc ? a.rc[1].at(t, n) : 0.0;
  return d;
}

// ONE element the fast algebra flagged and the masked-driver rule did not answer: the op-for-op
// evaluation (mod16_math.cuh).  Writes the outputs, returns the gate predicate g_surf > 0 (:343).
template <int SEM, int OUT>
__device__ __noinline__ bool rare64_one(const ForwardArgs<double>& a, int64_t t, int64_t n) {
  const Drivers<double> d = load_drivers64(a, t, n);
  const Params<double> q = load_params<double, SEM>(a, n, a.table);
  bool g_pos = false;
  Half<double> day, night;
  pixel_step<double, SEM>(d, q, a.has_rc != 0, a.gate, day, night, g_pos);
  store64<OUT>(a, t * a.N + n, day.canopy, day.soil, day.trans, night.canopy, night.soil,
               night.trans, half_total<double, SEM>(day), half_total<double, SEM>(night));
  return g_pos;
}

// Width of the array drivers of a launch: all float64, all float32 (the reference's stock
// call: float32 drivers + the float64 BPLUT), or mixed.
enum { W_F64 = 0, W_F32 = 1, W_MIXED = 2 };

// byte strides over time of the 16 sources (0: scalar or time-constant)
struct Steps64 { int64_t bytes[kSrc]; };

template <int W>
__device__ __forceinline__ double load64(const char* p, int is_f32) {
  if (W == W_F64) return *reinterpret_cast<const double*>(p);
  if (W == W_F32) return (double)*reinterpret_cast<const float*>(p);
  return is_f32 ? (double)*reinterpret_cast<const float*>(p) : *reinterpret_cast<const double*>(p);
}

// The masked-driver rule of the float32 kernels (fast::nan_halves, see mod16_fast.cuh for the
// derivation) on the float64 element in registers: a NaN (or negative-temperature) driver that
// makes a half-day total NaN whatever else the element holds is answered without the
// op-for-op code -- masked pixels of a real tile are masked at every step.  Bit 0 / bit 1:
// the day / night total is NaN; `g_pos`: the element's gate predicate.
template <int SEM>
__device__ __forceinline__ int nan_halves64(const double (&v)[kSrc],
                                            const fast::PftConstV<double>& c, bool has_rc,
                                            bool& g_pos) {
  g_pos = false;
  if (has_rc) return 0;
  const double t_d = v[MOD16_D_TEMP_DAY], t_n = v[MOD16_D_TEMP_NIGHT], p = v[MOD16_D_PRESSURE];
  const double vpd_d = v[MOD16_D_VPD_DAY], tmin = v[MOD16_D_TMIN];
  const bool n_t_d = !(t_d >= 0.0), n_t_n = !(t_n >= 0.0);       // NaN or negative
  const bool n_vpd_d = vpd_d != vpd_d, n_p = p != p;
  const bool n_fpar = v[MOD16_D_FPAR] != v[MOD16_D_FPAR];
  const bool n_lai = SEM == SEM_STATIC && v[MOD16_D_LAI] != v[MOD16_D_LAI];
  const bool day = v[MOD16_D_LW_NET_DAY] != v[MOD16_D_LW_NET_DAY] ||
                   v[MOD16_D_SW_RAD_DAY] != v[MOD16_D_SW_RAD_DAY] ||
                   v[MOD16_D_SW_ALBEDO] != v[MOD16_D_SW_ALBEDO] || n_t_d || n_vpd_d || n_p ||
                   n_fpar || n_lai;
  const bool night = v[MOD16_D_LW_NET_NIGHT] != v[MOD16_D_LW_NET_NIGHT] || n_t_n ||
                     v[MOD16_D_VPD_NIGHT] != v[MOD16_D_VPD_NIGHT] || n_p || n_fpar || n_lai;
  if (!(day || night)) return 0;
  if (SEM == SEM_STATIC && !(n_t_d || n_p || n_vpd_d || tmin != tmin)) {
    if (c.ok == 0.0) return 0;
    const bool sane = t_d > 206.3 && t_d < 340.0 && p >= 3.0e4 && p <= 1.2e5 &&
                      vpd_d > -1000.0 && vpd_d < 1.0e5 && fabs(tmin) < 9000.0;
    if (!sane) return 0;
    // g_surf = csl m_tmin m_vpd / r_corr, r_corr > 0 here (:325-328)
    const double m_t = fast::vsat(fma(tmin, c.inv_dtmin, c.tmin_bias));
    const double m_v = fast::vsat(fma(-vpd_d, c.inv_dvpd, c.vpd_off));
    g_pos = (c.csl * m_t) * m_v > 0.0;
  }
  return (day ? 1 : 0) | (night ? 2 : 0);
}

// One element from its 16 source values: the fast algebra, or the out-of-line rare path.
// Returns the element's gate predicate.
template <int SEM, int OUT>
__device__ __forceinline__ bool step64(const ForwardArgs<double>& a, const double (&v)[kSrc],
                                       double tann_lo, bool pix_static, bool has_rc,
                                       const fast::PftConstV<double>& pc,
                                       fast::PixConstV<double>& px, int64_t t, int64_t n) {
  if (!pix_static)
    px = fast::make_pix_const_s<double>(v[MOD16_D_PRESSURE], v[MOD16_D_TEMP_ANNUAL], tann_lo,
                                        pc.ok != 0.0, SEM);
  fast::StepInV<double> in;
  in.lw_d = v[MOD16_D_LW_NET_DAY];  in.lw_n = v[MOD16_D_LW_NET_NIGHT];
  in.sw_d = v[MOD16_D_SW_RAD_DAY];  in.sw_n = v[MOD16_D_SW_RAD_NIGHT];
  in.alb = v[MOD16_D_SW_ALBEDO];    in.t_d = v[MOD16_D_TEMP_DAY];
  in.t_n = v[MOD16_D_TEMP_NIGHT];   in.tmin = v[MOD16_D_TMIN];
  in.vpd_d = v[MOD16_D_VPD_DAY];    in.vpd_n = v[MOD16_D_VPD_NIGHT];
  in.fpar = v[MOD16_D_FPAR];        in.lai = v[MOD16_D_LAI];
  in.rc_d = v[MOD16_N_DRIVERS];     in.rc_n = v[MOD16_N_DRIVERS + 1];
  fast::StepOutV<double> r;
  bool g_pos;
  const bool need = has_rc ? fast::step_fast<double, SEM, true>(in, pc, px, a.gate, r, g_pos)
                           : fast::step_fast<double, SEM, false>(in, pc, px, a.gate, r, g_pos);
  if (need) {
    if (OUT != 2) {
      bool gp;
      const int h = nan_halves64<SEM>(v, pc, has_rc, gp);
      if (OUT == 0 ? h != 0 : h == 3) {
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        a.out[0][t * a.N + n] = qnan;
        if (OUT == 1) a.out[1][t * a.N + n] = qnan;
        return gp;
      }
    }
    return rare64_one<SEM, OUT>(a, t, n);
  }
  store64<OUT>(a, t * a.N + n, r.day.canopy, r.day.soil, r.day.trans, r.night.canopy,
               r.night.soil, r.night.trans, r.tot_day, r.tot_night);
  return g_pos;
}

// One thread per pixel walks a chunk of time steps.  Every driver is one pointer that
// advances by its own stride; scalar drivers point at a shared-memory copy of their value
// (stride 0), so the 14-16 loads of a step are straight-line code, issued back to back, and
// the step's arithmetic starts when the last one lands.
template <int SEM, int OUT, int W>
#ifndef M16_F64_MINB
#define M16_F64_MINB 3
#endif
__global__ void __launch_bounds__(128, M16_F64_MINB)
forward_fast64_kernel(const __grid_constant__ ForwardArgs<double> a,
                      const __grid_constant__ Steps64 step) {
  __shared__ double s_scal[kSrc];     // (W_F32: float values in the low words)
  if (a.skip_if_flag && *((volatile int*)a.gate_flag) != 0) return;
  if (threadIdx.x < kSrc) {
    const int k = threadIdx.x;
    const double v = k < MOD16_N_DRIVERS ? a.drv[k].scalar : a.rc[k - MOD16_N_DRIVERS].scalar;
    if (W == W_F32) reinterpret_cast<float*>(s_scal + k)[0] = (float)v;   // exact: host checked
    else s_scal[k] = v;
  }
  __syncthreads();
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
    const bool has_rc = a.has_rc != 0;
    const char* src[kSrc];
#pragma unroll
    for (int k = 0; k < kSrc; ++k) {
      const Arr<double>& A = k < MOD16_N_DRIVERS ? a.drv[k] : a.rc[k - MOD16_N_DRIVERS];
      const int es = (W == W_F32 || (W == W_MIXED && A.is_f32)) ? 4 : 8;
      src[k] = A.ptr ? reinterpret_cast<const char*>(A.ptr) + (t0 * A.ts + n * A.ns) * es
                     : reinterpret_cast<const char*>(s_scal + k);
    }
    const bool pix_static = a.drv[MOD16_D_PRESSURE].ts == 0 && a.drv[MOD16_D_TEMP_ANNUAL].ts == 0;
    fast::PixConstV<double> px = fast::make_pix_const_s<double>(
        load64<W>(src[MOD16_D_PRESSURE], a.drv[MOD16_D_PRESSURE].is_f32),
        load64<W>(src[MOD16_D_TEMP_ANNUAL], a.drv[MOD16_D_TEMP_ANNUAL].is_f32), q.tann_lo,
        pc.ok != 0.0, SEM);
    for (int64_t t = t0; t < t1; ++t) {
      double v[kSrc];
#pragma unroll
      for (int k = 0; k < kSrc; ++k) {
        const Arr<double>& A = k < MOD16_N_DRIVERS ? a.drv[k] : a.rc[k - MOD16_N_DRIVERS];
        if (k >= MOD16_N_DRIVERS && !has_rc) { v[k] = 0.0; continue; }
        v[k] = load64<W>(src[k], A.is_f32);
        src[k] += step.bytes[k];      // (a kernel parameter: an operand, not a register)
      }
      any_pos |= step64<SEM, OUT>(a, v, q.tann_lo, pix_static, has_rc, pc, px, t, n);
    }
  }
  if (a.track_gate) {
    if (__any_sync(0xffffffffu, any_pos) && (threadIdx.x & 31) == 0) *a.gate_flag = 1;
  }
}

// ---- the canonical layout: every time-varying source is a (T, N) array of one width with one
// time stride (what the reference's loaders produce) -------------------------------------------
// The step's inputs are fetched one step AHEAD with cp.async into a two-stage shared-memory
// ring (each thread copies, waits for and reads only its own 4- or 8-byte slots: no block
// barrier), so the DRAM latency of a step is covered by the arithmetic of the step before;
// addresses are one per-thread byte offset added to the uniform base pointers, so no source
// owns registers between steps.
struct Canon64 {
  int64_t step_bytes;     // time stride of the streamed sources
  uint32_t streamed;      // bit k: source k is streamed (else time-constant: read once)
};

// cp.async of one 4- or 8-byte element into the thread's own slot, if bit K of `mask` is set.
// (Inline PTX: the predicate and the address are formed next to the copy -- written in C++
// they are hoisted out of the time loop as 16 loop invariants and spilled.)
template <int K, int BYTES>
__device__ __forceinline__ void fetch_if(uint32_t mask, uint32_t slot, const void* base,
                                         int64_t ob) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 u;\n\t.reg .b64 g;\n\t"
      "and.b32 u, %0, %1;\n\tsetp.ne.u32 p, u, 0;\n\t"
      "add.s64 g, %3, %4;\n\t"
      "@p cp.async.ca.shared.global [%2], [g], %5;\n\t}"
      :: "r"(mask), "n"(1u << K), "r"(slot), "l"(base), "l"(ob), "n"(BYTES) : "memory");
}
template <int W, int K>
__device__ __forceinline__ void fetch_all(const ForwardArgs<double>& a, uint32_t mask,
                                          uint32_t slot0, int64_t ob) {
  constexpr int B = W == W_F32 ? 4 : 8;
  if (K < MOD16_N_DRIVERS) fetch_if<K, B>(mask, slot0 + K * 128 * B, a.drv[K].ptr, ob);
  else fetch_if<K, B>(mask, slot0 + K * 128 * B, a.rc[K - MOD16_N_DRIVERS].ptr, ob);
  if (K + 1 < kSrc) fetch_all<W, (K + 1 < kSrc ? K + 1 : K)>(a, mask, slot0, ob);
}

template <int SEM, int OUT, int W>
__global__ void __launch_bounds__(128, M16_F64_MINB)
forward_canon64_kernel(const __grid_constant__ ForwardArgs<double> a,
                       const __grid_constant__ Canon64 c) {
  typedef typename std::conditional<W == W_F32, float, double>::type E;
  __shared__ E ring[2][kSrc][128];
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
    const bool has_rc = a.has_rc != 0;
    E* const mine = &ring[0][0][threadIdx.x];
    constexpr int kStage = kSrc * 128;      // elements per stage
    const uint32_t slot = (uint32_t)__cvta_generic_to_shared(mine);
    // (without r_corr its two bits are clear: the launcher only sets bits of sources in use)
    int64_t ob = (t0 * (c.step_bytes / (int64_t)sizeof(E)) + n) * (int64_t)sizeof(E);
#pragma unroll
    for (int k = 0; k < kSrc; ++k) {
      const Arr<double>& A = k < MOD16_N_DRIVERS ? a.drv[k] : a.rc[k - MOD16_N_DRIVERS];
      if (((c.streamed >> k) & 1u) == 0) {
        // scalar / time-constant / unused: both stages hold the value for good
        const E v = (k < MOD16_N_DRIVERS || has_rc) ? (E)A.at(t0, n) : (E)0;
        mine[k * 128] = v;
        mine[kStage + k * 128] = v;
      }
    }
    fetch_all<W, 0>(a, c.streamed, slot, ob);
    __pipeline_commit();
    const bool pix_static = ((c.streamed >> MOD16_D_PRESSURE) & 1u) == 0 &&
                            ((c.streamed >> MOD16_D_TEMP_ANNUAL) & 1u) == 0;
    fast::PixConstV<double> px = fast::make_pix_const_s<double>(
        a.drv[MOD16_D_PRESSURE].at(t0, n), a.drv[MOD16_D_TEMP_ANNUAL].at(t0, n), q.tann_lo,
        pc.ok != 0.0, SEM);
    int cur = 0;
    for (int64_t t = t0; t < t1; ++t) {
      ob += c.step_bytes;
      if (t + 1 < t1)
        fetch_all<W, 0>(a, c.streamed, slot + (uint32_t)((cur ^ kStage) * sizeof(E)), ob);
      __pipeline_commit();
      __pipeline_wait_prior(1);
      double v[kSrc];
#pragma unroll
      for (int k = 0; k < kSrc; ++k) v[k] = (double)mine[cur + k * 128];
      cur ^= kStage;
      any_pos |= step64<SEM, OUT>(a, v, q.tann_lo, pix_static, has_rc, pc, px, t, n);
    }
  }
  if (a.track_gate) {
    if (__any_sync(0xffffffffu, any_pos) && (threadIdx.x & 31) == 0) *a.gate_flag = 1;
  }
}

// Is the launch in the canonical layout?  Fills `c`.
static bool canonical_layout(const ForwardArgs<double>& a, int w, Canon64& c) {
  if (w == W_MIXED) return false;
  c.step_bytes = 0; c.streamed = 0;
  const int64_t es = w == W_F32 ? 4 : 8;
  for (int k = 0; k < kSrc; ++k) {
    if (k >= MOD16_N_DRIVERS && !a.has_rc) break;
    const Arr<double>& A = k < MOD16_N_DRIVERS ? a.drv[k] : a.rc[k - MOD16_N_DRIVERS];
    if (!A.ptr || A.ts == 0) continue;       // scalar / time-constant: read once
    if (A.ns != 1 || A.ts < 0) return false;
    if (c.streamed && A.ts * es != c.step_bytes) return false;
    c.step_bytes = A.ts * es;
    c.streamed |= 1u << k;
  }
  return c.streamed != 0;
}

template <int SEM, int W>
static int launch64_canon(const ForwardArgs<double>& a, const Canon64& c, uint32_t om, dim3 grid,
                          dim3 block, cudaStream_t st) {
  if (om == MOD16_OUT_SUM) forward_canon64_kernel<SEM, 0, W><<<grid, block, 0, st>>>(a, c);
  else if (om == MOD16_OUT_DAY_NIGHT) forward_canon64_kernel<SEM, 1, W><<<grid, block, 0, st>>>(a, c);
  else forward_canon64_kernel<SEM, 2, W><<<grid, block, 0, st>>>(a, c);
  M16_CUDA_CHECK(cudaGetLastError());
  return MOD16_OK;
}

template <int SEM, int W>
static int launch64_w(const ForwardArgs<double>& a, uint32_t om, dim3 grid, dim3 block,
                      cudaStream_t st) {
  Steps64 s;
  for (int k = 0; k < kSrc; ++k) {
    const Arr<double>& A = k < MOD16_N_DRIVERS ? a.drv[k] : a.rc[k - MOD16_N_DRIVERS];
    s.bytes[k] = A.ptr ? A.ts * (A.is_f32 ? 4 : 8) : 0;
  }
  if (om == MOD16_OUT_SUM) forward_fast64_kernel<SEM, 0, W><<<grid, block, 0, st>>>(a, s);
  else if (om == MOD16_OUT_DAY_NIGHT) forward_fast64_kernel<SEM, 1, W><<<grid, block, 0, st>>>(a, s);
  else forward_fast64_kernel<SEM, 2, W><<<grid, block, 0, st>>>(a, s);
  M16_CUDA_CHECK(cudaGetLastError());
  return MOD16_OK;
}

// W_F32 needs every array driver to be float32 and every scalar driver to be a float32 value
// (the kernel reads scalars through the same 4-byte loads); W_F64 needs no float32 array.
static int source_width(const ForwardArgs<double>& a) {
  int n32 = 0, n64 = 0, exact = 1;
  for (int k = 0; k < kSrc; ++k) {
    if (k >= MOD16_N_DRIVERS && !a.has_rc) break;
    const Arr<double>& A = k < MOD16_N_DRIVERS ? a.drv[k] : a.rc[k - MOD16_N_DRIVERS];
    if (!A.ptr) { if (!((double)(float)A.scalar == A.scalar)) exact = 0; }   // (NaN: not exact)
    else if (A.is_f32) ++n32;
    else ++n64;
  }
  if (n32 == 0) return W_F64;
  if (n64 == 0 && exact) return W_F32;
  return W_MIXED;
}

template <int SEM>
static int launch64_sem(const ForwardArgs<double>& a, uint32_t om, dim3 grid, dim3 block,
                        cudaStream_t st) {
  const int w = source_width(a);
  Canon64 c;
  if (canonical_layout(a, w, c) && !getenv("MOD16_F64_GENERAL")) {
    g_last_forward_kernel = MOD16_KERNEL_FAST64 | 0x100;
    return w == W_F64 ? launch64_canon<SEM, W_F64>(a, c, om, grid, block, st)
                      : launch64_canon<SEM, W_F32>(a, c, om, grid, block, st);
  }
  if (w == W_F64) return launch64_w<SEM, W_F64>(a, om, grid, block, st);
  if (w == W_F32) return launch64_w<SEM, W_F32>(a, om, grid, block, st);
  return launch64_w<SEM, W_MIXED>(a, om, grid, block, st);
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
