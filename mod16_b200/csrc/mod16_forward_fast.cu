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

#ifndef VEC4_THREADS
#define VEC4_THREADS 256
#endif
#ifndef VEC4_MIN_BLOCKS
#define VEC4_MIN_BLOCKS 2
#endif

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
// The same constants for every PAIR of land-cover codes, interleaved so that one shared
// load yields the packed (pixel 0, pixel 1) operand of an FFMA2 (14 x 14 x 96 B = 18.4 KB).
struct PairLut {
  fast::PftConstV<fast::F2> pp[MOD16_N_PFT + 1][MOD16_N_PFT + 1];
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

__device__ __forceinline__ void stage_pair_lut(const SharedLut& s, PairLut& p) {
  constexpr int NC = MOD16_N_PFT + 1, NF = sizeof(fast::PftConstV<float>) / sizeof(float);
  static_assert(sizeof(fast::PftConstV<fast::F2>) == NF * sizeof(fast::F2), "layout");
  for (int i = threadIdx.x; i < NC * NC * NF; i += blockDim.x) {
    const int f = i % NF, b = (i / NF) % NC, a = i / (NF * NC);
    const float* fa = reinterpret_cast<const float*>(static_cast<const fast::PftConstV<float>*>(&s.pc[a]));
    const float* fb = reinterpret_cast<const float*>(static_cast<const fast::PftConstV<float>*>(&s.pc[b]));
    reinterpret_cast<fast::F2*>(&p.pp[a][b])[f] = fast::F2{fa[f], fb[f]};
  }
  __syncthreads();
}

// Deferred exact re-evaluation of the elements of one float4 that step_fast() flagged.
// Out of line and self-contained (re-reads its inputs from global memory, which the fast
// pass has just pulled through L2), so that the double-precision code costs the hot loop
// no registers.  Returns the gate predicate OR-ed over the re-evaluated elements.
template <int SEM, int OUT>
__device__ __noinline__ bool fix_exact_vec4(const ForwardArgs<float>& a, const SharedLut& lut,
                                            int64_t t, int64_t n0, unsigned mask) {
  bool any_pos = false;
  for (int j = 0; j < 4; ++j) {
    if (!((mask >> j) & 1u)) continue;  // (the pipelined kernel passes pairs: j < 2)
    const int64_t n = n0 + j;
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
    in.rc_d = 0.0f; in.rc_n = 0.0f;
    int code = 0;
    if (a.par_mode == 1) {
      code = (int)__ldg(a.pft + n);
      code = code < MOD16_N_PFT ? code : MOD16_N_PFT;
    }
    double col[MOD16_N_PARAMS];
    for (int k = 0; k < MOD16_N_PARAMS; ++k) col[k] = lut.tab[k][code];
    fast::StepOut r;
    bool g_pos;
    fast::pixel_step_exact<SEM>(in, col, false, a.gate, r, g_pos);
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
  return any_pos;
}

__device__ __forceinline__ fast::F2 lo2(const float4& v) { return fast::F2{v.x, v.y}; }
__device__ __forceinline__ fast::F2 hi2(const float4& v) { return fast::F2{v.z, v.w}; }

template <int SEM, int OUT>
__global__ void __launch_bounds__(VEC4_THREADS, VEC4_MIN_BLOCKS)
forward_vec4_kernel(const __grid_constant__ ForwardArgs<float> a) {
  using fast::F2;
  using fast::M2;
  __shared__ SharedLut lut;
  __shared__ PairLut plut;
  if (a.skip_if_flag && *((volatile int*)a.gate_flag) != 0) return;
  stage_lut<SEM>(a, lut);
  stage_pair_lut(lut, plut);
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
    fast::PixConstV<F2> px[2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
      px[h] = fast::pair_pix(
          fast::make_pix_const(comp(p4, 2 * h), comp(ta4, 2 * h), lut.pc[code[2 * h]], SEM),
          fast::make_pix_const(comp(p4, 2 * h + 1), comp(ta4, 2 * h + 1),
                               lut.pc[code[2 * h + 1]], SEM));
    const fast::PftConstV<F2>* pc2[2] = {&plut.pp[code[0]][code[1]], &plut.pp[code[2]][code[3]]};

    const Arr<float>& swn = a.drv[MOD16_D_SW_RAD_NIGHT];
    const int gate = a.gate;
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
      F2 res[6][2];
      unsigned redo = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        fast::StepInV<F2> in;
        in.lw_d = h ? hi2(lw_d) : lo2(lw_d);   in.lw_n = h ? hi2(lw_n) : lo2(lw_n);
        in.sw_d = h ? hi2(sw_d) : lo2(sw_d);   in.sw_n = h ? hi2(sw_n) : lo2(sw_n);
        in.alb = h ? hi2(alb) : lo2(alb);      in.t_d = h ? hi2(t_d) : lo2(t_d);
        in.t_n = h ? hi2(t_n) : lo2(t_n);      in.tmin = h ? hi2(tmin) : lo2(tmin);
        in.vpd_d = h ? hi2(vpd_d) : lo2(vpd_d); in.vpd_n = h ? hi2(vpd_n) : lo2(vpd_n);
        in.fpar = h ? hi2(fpar) : lo2(fpar);   in.lai = h ? hi2(lai) : lo2(lai);
        in.rc_d = F2{0.f, 0.f}; in.rc_n = F2{0.f, 0.f};
        fast::StepOutV<F2> r;
        M2 g_pos;
        const M2 need = fast::step_fast<F2, SEM, false>(in, *pc2[h], px[h], gate, r, g_pos);
        redo |= (need.x ? 1u : 0u) << (2 * h) | (need.y ? 2u : 0u) << (2 * h);
        any_pos |= (g_pos.x && !need.x) || (g_pos.y && !need.y);
        if (OUT == 0) {
          res[0][h] = r.tot;
        } else if (OUT == 1) {
          res[0][h] = r.tot_day; res[1][h] = r.tot_night;
        } else {
          res[0][h] = r.day.canopy;   res[1][h] = r.day.soil;   res[2][h] = r.day.trans;
          res[3][h] = r.night.canopy; res[4][h] = r.night.soil; res[5][h] = r.night.trans;
        }
      }
      constexpr int n_out = OUT == 0 ? 1 : (OUT == 1 ? 2 : 6);
#pragma unroll
      for (int k = 0; k < n_out; ++k)
        st_stream(a.out[k] + off,
                  make_float4(res[k][0].x, res[k][0].y, res[k][1].x, res[k][1].y));
      if (redo) any_pos |= fix_exact_vec4<SEM, OUT>(a, lut, t, n0, redo);
    }
  }
  if (a.track_gate) {
    if (__any_sync(0xffffffffu, any_pos) && (threadIdx.x & 31) == 0) *a.gate_flag = 1;
  }
}

// ---- forward_pipe_kernel: the headline kernel, fed by bulk asynchronous copies ----------
// One block owns a tile of PIPE_TILE = 512 pixels (one packed pair per thread) and walks a
// chunk of time steps.  For every step one elected thread issues eleven (twelve with a
// sw_rad_night array) `cp.async.bulk` copies -- 2 KB of each driver row -- into one of
// PIPE_STAGES shared-memory stages and arms the stage's `full` mbarrier with the byte
// count.  Each warp waits on the barrier's phase, evaluates its 64 pixel-timesteps out of
// shared memory (64-bit conflict-free LDS straight into FFMA2 operand pairs)
// straight into registers and releases the stage at once (a shared counter); the last of
// the eight warps to release a stage re-fills it, so nobody ever waits for a free stage.
// Copies of the next PIPE_STAGES steps are in flight while a step is evaluated: HBM latency
// is hidden by the copy engine instead of by registers (the direct-load kernel holds 44
// registers of loads in flight and spills), there is no block-wide barrier and no address
// arithmetic in the time loop.
constexpr int PIPE_THREADS = 256;
constexpr int PIPE_WARPS = PIPE_THREADS / 32;
constexpr int PIPE_TILE = PIPE_THREADS * 2;
constexpr int PIPE_STAGES = 4;
constexpr int PIPE_MAX_ARR = 12;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// (all barrier / stage addresses are 32-bit shared-window addresses, computed once)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ fast::F2 lds2(uint32_t addr) {
  fast::F2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "M16_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra M16_DONE;\n"
      "bra M16_WAIT;\n"
      "M16_DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_stream2(float* p, fast::F2 v) {
  asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

struct PipeSrc {  // per-step arrays in stage order
  const float* ptr[PIPE_MAX_ARR];
  int n_arr;
};

// Producer step (one thread): arm the barrier, start the row copies.  The array base
// pointers are read from a shared-memory copy (a generic load from the kernel-parameter
// space costs hundreds of cycles, and the eleven copies issue back to back).
__device__ __forceinline__ void pipe_issue(const float* const* sptr, int n_arr, uint32_t dst,
                                           uint32_t bar, int64_t off, uint32_t row_bytes) {
  mbar_expect_tx(bar, row_bytes * (uint32_t)n_arr);
#pragma unroll 4
  for (int k = 0; k < n_arr; ++k)
    bulk_g2s(dst + (uint32_t)k * PIPE_TILE * (uint32_t)sizeof(float), sptr[k] + off, row_bytes,
             bar);
}

template <int SEM, int OUT>
__global__ void __launch_bounds__(PIPE_THREADS, 2)
forward_pipe_kernel(const __grid_constant__ ForwardArgs<float> a,
                    const __grid_constant__ PipeSrc src) {
  using fast::F2;
  using fast::M2;
  extern __shared__ __align__(128) unsigned char dyn_smem[];  // [stage][array][PIPE_TILE]
  __shared__ SharedLut lut;
  __shared__ PairLut plut;
  __shared__ __align__(8) uint64_t full[PIPE_STAGES];
  __shared__ int released[PIPE_STAGES];  // warps that have read the stage's current contents
  __shared__ const float* sptr[PIPE_MAX_ARR];
  if (a.skip_if_flag && *((volatile int*)a.gate_flag) != 0) return;
  const int tid = threadIdx.x;
  const int n_arr = src.n_arr;
  const int64_t tile0 = (int64_t)blockIdx.x * PIPE_TILE;
  const int64_t t0 = (int64_t)blockIdx.y * a.t_chunk;
  int64_t t1 = t0 + a.t_chunk;
  if (t1 > a.T) t1 = a.T;
  const int nsteps = (int)(t1 - t0);
  const int64_t left = a.N - tile0;
  const uint32_t row_bytes = (uint32_t)((left < PIPE_TILE ? left : PIPE_TILE) * sizeof(float));
  const uint32_t stage_bytes = (uint32_t)n_arr * PIPE_TILE * sizeof(float);

  const uint32_t full0 = smem_u32(&full[0]);
  const uint32_t data0 = smem_u32(dyn_smem);
  if (tid == 0) {
    for (int s = 0; s < PIPE_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      released[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int k = 0; k < PIPE_MAX_ARR; ++k) sptr[k] = src.ptr[k];
    for (int s = 0; s < PIPE_STAGES && s < nsteps; ++s)
      pipe_issue(sptr, n_arr, data0 + s * stage_bytes, full0 + 8 * s, (t0 + s) * a.N + tile0,
                 row_bytes);
  }
  stage_lut<SEM>(a, lut);       // (both contain __syncthreads: the barriers are visible)
  stage_pair_lut(lut, plut);

  const int64_t n0 = tile0 + (int64_t)tid * 2;
  const bool live = n0 < a.N;   // N is even: a pair is inside or outside as a whole
  bool any_pos = false;
  // per-pixel state: land-cover codes, pressure terms, annual-temperature window
  int code0 = 0, code1 = 0;
  fast::PixConstV<F2> px;
  {
    float2 p2 = make_float2(1e5f, 1e5f), ta2 = make_float2(0.f, 0.f);
    if (live) {
      if (a.par_mode == 1) {
        const uint32_t w = __ldg(reinterpret_cast<const uint16_t*>(a.pft + n0));
        code0 = w & 0xff; code1 = (w >> 8) & 0xff;
        code0 = code0 < MOD16_N_PFT ? code0 : MOD16_N_PFT;  // out of range: all-NaN column
        code1 = code1 < MOD16_N_PFT ? code1 : MOD16_N_PFT;
      }
      const Arr<float>& ap = a.drv[MOD16_D_PRESSURE];
      const Arr<float>& aa = a.drv[MOD16_D_TEMP_ANNUAL];
      if (ap.ptr) p2 = __ldg(reinterpret_cast<const float2*>(ap.ptr + n0));
      else p2 = make_float2(ap.scalar, ap.scalar);
      if (aa.ptr) ta2 = __ldg(reinterpret_cast<const float2*>(aa.ptr + n0));
      else ta2 = make_float2(aa.scalar, aa.scalar);
    }
    px = fast::pair_pix(fast::make_pix_const(p2.x, ta2.x, lut.pc[code0], SEM),
                        fast::make_pix_const(p2.y, ta2.y, lut.pc[code1], SEM));
  }
  const fast::PftConstV<F2>& pc2 = plut.pp[code0][code1];
  const float swn_scalar = a.drv[MOD16_D_SW_RAD_NIGHT].scalar;
  const bool swn_array = n_arr == PIPE_MAX_ARR;
  const int gate = a.gate;
  constexpr uint32_t ROWB = PIPE_TILE * sizeof(float);  // bytes per array row of a stage
  constexpr int n_out = OUT == 0 ? 1 : (OUT == 1 ? 2 : 6);
  const uint32_t my0 = data0 + (uint32_t)tid * 8u;  // this thread's pair within a row
  int64_t off = t0 * a.N + n0;                      // running output offset
  const int64_t t_stride = a.N;
  const bool warp_lead = (tid & 31) == 0;
  int stage = 0;
  uint32_t phase = 0;

#pragma unroll 1
  for (int i = 0; i < nsteps; ++i) {
    mbar_wait(full0 + 8 * stage, phase);
    // stage order == PipeSrc order (see launch_pipe).  The step's inputs are pulled into
    // registers at once so that the stage can be re-filled while the step is evaluated.
    fast::StepInV<F2> in;
    {
      const uint32_t sp = my0 + (uint32_t)stage * stage_bytes;
      in.lw_d = lds2(sp + 0 * ROWB);  in.lw_n = lds2(sp + 1 * ROWB);  in.sw_d = lds2(sp + 2 * ROWB);
      in.alb = lds2(sp + 3 * ROWB);   in.t_d = lds2(sp + 4 * ROWB);   in.t_n = lds2(sp + 5 * ROWB);
      in.tmin = lds2(sp + 6 * ROWB);  in.vpd_d = lds2(sp + 7 * ROWB); in.vpd_n = lds2(sp + 8 * ROWB);
      in.fpar = lds2(sp + 9 * ROWB);  in.lai = lds2(sp + 10 * ROWB);
      in.sw_n = swn_array ? lds2(sp + 11 * ROWB) : F2{swn_scalar, swn_scalar};
      in.rc_d = F2{0.f, 0.f}; in.rc_n = F2{0.f, 0.f};
    }
    __syncwarp();
    if (warp_lead) {
      // release the stage; the last of the eight warps to do so re-fills it (nobody waits)
      __threadfence_block();
      const int prior = atomicAdd(&released[stage], 1);
      if (prior == PIPE_WARPS - 1) {
        released[stage] = 0;
        __threadfence_block();
        if (i + PIPE_STAGES < nsteps)
          pipe_issue(sptr, n_arr, data0 + (uint32_t)stage * stage_bytes, full0 + 8 * stage,
                     (t0 + i + PIPE_STAGES) * a.N + tile0, row_bytes);
      }
    }
    __syncwarp();
    if (live) {
      fast::StepOutV<F2> r;
      M2 g_pos;
      const M2 need = fast::step_fast<F2, SEM, false>(in, pc2, px, gate, r, g_pos);
      any_pos |= (g_pos.x && !need.x) || (g_pos.y && !need.y);
      F2 res[6];
      if (OUT == 0) {
        res[0] = r.tot;
      } else if (OUT == 1) {
        res[0] = r.tot_day; res[1] = r.tot_night;
      } else {
        res[0] = r.day.canopy;   res[1] = r.day.soil;   res[2] = r.day.trans;
        res[3] = r.night.canopy; res[4] = r.night.soil; res[5] = r.night.trans;
      }
#pragma unroll
      for (int k = 0; k < n_out; ++k) st_stream2(a.out[k] + off, res[k]);
      if (need.x || need.y)
        any_pos |= fix_exact_vec4<SEM, OUT>(a, lut, t0 + i, n0, (need.x ? 1u : 0u) | (need.y ? 2u : 0u));
    }
    off += t_stride;
    if (++stage == PIPE_STAGES) { stage = 0; phase ^= 1u; }
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

template <int SEM, int OUT>
static int launch_pipe_one(const ForwardArgs<float>& a, const PipeSrc& src, dim3 grid,
                           size_t smem, cudaStream_t st) {
  static bool configured = false;  // per instantiation; the attribute is per function
  if (!configured) {
    M16_CUDA_CHECK(cudaFuncSetAttribute(forward_pipe_kernel<SEM, OUT>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        PIPE_STAGES * PIPE_MAX_ARR * PIPE_TILE * (int)sizeof(float)));
    configured = true;
  }
  forward_pipe_kernel<SEM, OUT><<<grid, PIPE_THREADS, smem, st>>>(a, src);
  M16_CUDA_CHECK(cudaGetLastError());
  return MOD16_OK;
}

// Grid of the pipelined kernel: one block per 1024-pixel tile and time chunk; the chunk is
// as long as possible (the per-block LUT staging is amortised over it) while the grid
// still holds several waves of 148 SMs x 2 resident blocks.
static int launch_pipe(ForwardArgs<float>& a, uint32_t flags, cudaStream_t st) {
  static const int order[] = {
      MOD16_D_LW_NET_DAY, MOD16_D_LW_NET_NIGHT, MOD16_D_SW_RAD_DAY, MOD16_D_SW_ALBEDO,
      MOD16_D_TEMP_DAY,   MOD16_D_TEMP_NIGHT,   MOD16_D_TMIN,       MOD16_D_VPD_DAY,
      MOD16_D_VPD_NIGHT,  MOD16_D_FPAR,         MOD16_D_LAI,        MOD16_D_SW_RAD_NIGHT};
  PipeSrc src;
  src.n_arr = a.drv[MOD16_D_SW_RAD_NIGHT].ptr ? PIPE_MAX_ARR : PIPE_MAX_ARR - 1;
  for (int k = 0; k < PIPE_MAX_ARR; ++k) src.ptr[k] = k < src.n_arr ? a.drv[order[k]].ptr : nullptr;
  const int64_t bx = (a.N + PIPE_TILE - 1) / PIPE_TILE;
  if (bx > 2147483647LL) return MOD16_ERR_SHAPE;
  int64_t want_rows = (148 * 2 * 6 + bx - 1) / bx;
  if (want_rows < 1) want_rows = 1;
  int64_t chunk = (a.T + want_rows - 1) / want_rows;
  if (chunk < 8) chunk = a.T < 8 ? a.T : 8;
  int64_t by = (a.T + chunk - 1) / chunk;
  while (by > 65535) { chunk *= 2; by = (a.T + chunk - 1) / chunk; }
  a.t_chunk = (int)chunk;
  dim3 grid((unsigned)bx, (unsigned)by);
  const size_t smem = (size_t)PIPE_STAGES * src.n_arr * PIPE_TILE * sizeof(float);
  const uint32_t om = flags & MOD16_OUT_MASK;
  const bool is_static = (flags & MOD16_SEM_MASK) == MOD16_SEM_STATIC;
#define M16_PIPE(S, O) return launch_pipe_one<S, O>(a, src, grid, smem, st)
  if (is_static) {
    if (om == MOD16_OUT_SUM) M16_PIPE(SEM_STATIC, 0);
    if (om == MOD16_OUT_DAY_NIGHT) M16_PIPE(SEM_STATIC, 1);
    M16_PIPE(SEM_STATIC, 2);
  }
  if (om == MOD16_OUT_SUM) M16_PIPE(SEM_INSTANCE, 0);
  if (om == MOD16_OUT_DAY_NIGHT) M16_PIPE(SEM_INSTANCE, 1);
  M16_PIPE(SEM_INSTANCE, 2);
#undef M16_PIPE
}

int launch_fast_f32(ForwardArgs<float>& a, uint32_t flags, cudaStream_t st) {
  if (a.T == 0 || a.N == 0) return MOD16_OK;
  const bool vec = !(flags & MOD16_FORCE_GENERIC) && vec4_ok(a);
  if (vec && !(flags & MOD16_FORCE_DIRECT)) return launch_pipe(a, flags, st);
  const int threads = vec ? VEC4_THREADS : 256;
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
