// K1, float32 fast path: forward_pipe_kernel -- the headline kernel.
//
// Layout: (T, N) row-major structure-of-arrays.  One block per SM-resident tile of 480
// pixels walks a chunk of time steps two at a time.  A producer warp streams the per-step
// driver rows of the next steps into a shared-memory ring with bulk asynchronous copies
// (cp.async.bulk, completion on an mbarrier; SASS UBLKCP); fifteen consumer warps evaluate
// the model with the packed-FP32 (FFMA2) math of mod16_fast.cuh -- the two lanes of every
// packed instruction are the same pixel at two consecutive time steps, so all
// BPLUT-derived and per-pixel constants are scalar registers -- and write ET exactly once.
// Algorithmic traffic: 11 x 4 B in + 4 B out = 48 B per pixel-timestep (OUT_SUM).
//
// Warp specialisation: warp 15 is the PRODUCER; for every pair of time steps it waits until
// a stage is free, arms the stage's `full` mbarrier with the byte count and issues one
// `cp.async.bulk` of 1920 B per driver row and step.  The fifteen CONSUMER warps wait on
// the barrier's phase, pull their inputs out of shared memory (conflict-free 32-bit LDS,
// the rows of step t and t + 1 landing in the two halves of an FFMA2 operand pair), release
// the stage at once (`empty` mbarrier) and then evaluate and store.  The copies of the next
// 2 x stages steps are in flight while a pair of steps is evaluated: HBM latency is hidden
// by the copy engine instead of by registers, there is no block-wide barrier, no table
// gather and no address arithmetic in the time loop.
//
// The set of streamed arrays is a template parameter (EXTRA), so every shared-memory offset
// of the loop is an immediate:
//   EXTRA 0: the eleven per-step drivers                                     (5 stages)
//   EXTRA 1: + sw_rad_night as a dense array (every reference caller passes 0) (4 stages)
//   EXTRA 2: + temp_annual and pressure varying in time -- the (T x N) annual means of
//            mod16/calibration.py:556-560; the per-pixel terms are then rebuilt per step
//   EXTRA 3: all fourteen
//   EXTRA 4: + surface pressure day / night; the two VPD slots hold the water vapour mixing
//            ratio QV10M and VPD = MOD16.vpd(qv10m, ps, temp) (mod16/__init__.py:605-644) is
//            formed on the fly -- the loaders' other way to the VPD (calibration.py:393-400)
// Parameters come from the BPLUT (one column, or gathered by land cover) or from per-pixel
// arrays -- the `params_dict[key][pft_map].reshape((1, N))` form of the reference's own
// callers (notebooks/MOD16_Forward_Run_example.ipynb) --: they are time-invariant, so the
// block prologue turns either into the same scalar registers.
//
// Elements that need the double-precision code are not evaluated in the loop -- one such
// element would hold its warp for ~20 us and, through the ring, the whole block -- but
// queued and evaluated afterwards, one element per thread: by redo_queue_kernel when the
// caller supplied a workspace, else by this block after its loop.
#pragma once
#include <type_traits>

#include "mod16_pipe.cuh"
#include "mod16_rare.cuh"

namespace mod16 {

#ifndef M16_PIPE_WARPS
#define M16_PIPE_WARPS 15
#endif
constexpr int PIPE_CONSUMER_WARPS = M16_PIPE_WARPS;
constexpr int PIPE_CONSUMERS = PIPE_CONSUMER_WARPS * 32;
constexpr int PIPE_THREADS = PIPE_CONSUMERS + 32;
constexpr int PIPE_TILE = PIPE_CONSUMERS;  // pixels per block
constexpr int PIPE_SUB = 2;                // time steps per stage (the two FFMA2 lanes)
constexpr int PIPE_MAX_ARR = 14;
#ifndef M16_PIPE_MAX_STAGES
#define M16_PIPE_MAX_STAGES 5
#endif
constexpr int PIPE_MAX_STAGES = M16_PIPE_MAX_STAGES;
constexpr int REDO_MARK = 0x7fc0dead;  // a NaN no arithmetic produces: "evaluate me exactly"
constexpr int PIPE_QCAP = 512;             // deferred exact re-evaluations per block
constexpr uint32_t PIPE_ROWB = PIPE_TILE * sizeof(float);  // bytes per array row

__host__ __device__ constexpr int pipe_n_arr(int extra) {
  return 11 + ((extra & 1) ? 1 : 0) + ((extra & 2) ? 2 : 0) + ((extra & 4) ? 2 : 0);
}
// Storage formats of the streamed arrays (FMT): 0 float32, no loader conversions (the plain
// entries); 1 float32 + the `prep` conversions; 2 MODIS (sw_albedo int16, fpar / lai uint8);
// 3 PACKED (every other per-step driver int16 as well).  Stage order of the arrays:
// lw_d lw_n sw_d alb t_d t_n tmin vpd_d vpd_n fpar lai [sw_n] [t_ann p].
constexpr int FMT_F32 = 0, FMT_F32_PREP = 1, FMT_MODIS = 2, FMT_PACKED = 3;
__host__ __device__ constexpr int pipe_esize(int fmt, int k) {
  return fmt < FMT_MODIS ? 4 : (k == 9 || k == 10) ? 1 : (k == 3 || fmt == FMT_PACKED) ? 2 : 4;
}
// element offset (in units of PIPE_TILE bytes) of array k inside one sub-step of a stage
__host__ __device__ constexpr int pipe_eoff(int fmt, int k) {
  int o = 0;
  for (int j = 0; j < k; ++j) o += pipe_esize(fmt, j);
  return o;
}
// as many stages as fit into 227 KB of shared memory next to ~8 KB of static data, at most
// PIPE_MAX_STAGES (more copies in flight than that buy nothing)
__host__ __device__ constexpr int pipe_n_stages(int extra, int fmt = 0) {
  const int fit = (int)((232448u - 8192u) /
                        ((uint32_t)(PIPE_SUB * pipe_eoff(fmt, pipe_n_arr(extra))) * PIPE_TILE));
  return fit < PIPE_MAX_STAGES ? fit : PIPE_MAX_STAGES;
}
// stage slot of an optional array (the first eleven are fixed, see launch_pipe)
__host__ __device__ constexpr int pipe_slot_swn(int extra) { return 11; }
__host__ __device__ constexpr int pipe_slot_tann(int extra) { return (extra & 1) ? 12 : 11; }
__host__ __device__ constexpr int pipe_slot_p(int extra) { return pipe_slot_tann(extra) + 1; }

struct PipeSrc {  // streamed arrays in stage order: base pointer and BYTES per time step
  const char* ptr[PIPE_MAX_ARR];
  int64_t ts[PIPE_MAX_ARR];
};

// Producer step, executed by the WHOLE producer warp so that addresses stay in uniform
// registers and the base pointers come straight from the constant bank; one elected lane
// arms the barrier and starts the row copies of `n_sub` time steps.  (Measured: this form
// issues > 50 copies/us, an elected thread inside a divergent branch only 18 --
// profiles/r01_issuebench.txt.)
template <int NARR, int FMT>
__device__ __forceinline__ void pipe_issue(const PipeSrc& src, int n_sub, uint32_t dst,
                                           uint32_t bar, int64_t t, int64_t tile0,
                                           uint32_t n_live) {
  constexpr uint32_t SUBE = pipe_eoff(FMT, NARR);   // bytes per pixel of one sub-step
  if (elect_one()) {
    mbar_expect_tx(bar, n_live * SUBE * (uint32_t)n_sub);
#pragma unroll
    for (int u = 0; u < PIPE_SUB; ++u) {
      if (u < n_sub) {
#pragma unroll
        for (int k = 0; k < NARR; ++k)
          bulk_g2s(dst + (uint32_t)(u * SUBE + pipe_eoff(FMT, k)) * PIPE_TILE,
                   src.ptr[k] + (t + u) * src.ts[k] + tile0 * pipe_esize(FMT, k),
                   n_live * (uint32_t)pipe_esize(FMT, k), bar);
      }
    }
  }
  __syncwarp();
}

// One streamed driver of this thread's pixel at the two time steps of a stage, as the floats
// the model computes with: storage formats are converted with one packed FMA; the fill value
// (an impossible one when the driver has none, see mod16_forward_raw_f32) becomes NaN.
template <int FMT, int K>
__device__ __forceinline__ fast::F2 pipe_load(uint32_t s0, uint32_t s1, uint32_t tid,
                                              const ForwardArgs<float>& a, int drv_index) {
  constexpr int ES = pipe_esize(FMT, K);
  constexpr uint32_t OFF = (uint32_t)pipe_eoff(FMT, K) * PIPE_TILE;
  const uint32_t a0 = s0 + OFF + tid * ES, a1 = s1 + OFF + tid * ES;
  if (ES == 4) return fast::F2{lds1(a0), lds1(a1)};
  int qx, qy;
  if (ES == 2) {
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(qx) : "r"(a0));
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(qy) : "r"(a1));
  } else {
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(qx) : "r"(a0));
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(qy) : "r"(a1));
  }
  const RawInfo& r = a.raw[drv_index];
  fast::F2 v = fast::vfma(fast::F2{(float)qx, (float)qy}, r.scale, r.offset);
  const float qnan = __int_as_float(0x7fc00000);
  v.x = qx == r.fill ? qnan : v.x;
  v.y = qy == r.fill ? qnan : v.y;
  return v;
}

// ---- rare-path helpers ---------------------------------------------------------------------
// Out of line, and they work out the element's indices themselves (block and thread index,
// `ts` = time step within the block's chunk): the hot loop keeps nothing alive for them.

// Defers ONE element to the double-precision pass: the caller's queue (redo_queue_kernel)
// or -- no workspace -- this block's own queue (evaluated after its loop); a full queue
// leaves a marker in the output instead, which redo_sweep_kernel / the block's own sweep
// finds (the counter past the capacity tells them to run).  One in-place double evaluation
// here would hold the whole block.  Inlined on purpose: a CALL inside the time loop keeps
// the loop's uniform values out of the uniform registers.
template <int SEM, int OUT>
__device__ __forceinline__ bool pipe_defer(const ForwardArgs<float>& a, int ts, uint32_t* redo_q,
                                           int* redo_n) {
  const int tid = (int)threadIdx.x;
  const int64_t n = (int64_t)blockIdx.x * PIPE_TILE + tid;
  const int64_t t = (int64_t)blockIdx.y * a.t_chunk + ts;
  bool queued;
  if (a.redo_cap > 0) {
    const unsigned long long slot = atomicAdd(a.redo_ws, 1ULL);
    queued = slot < (unsigned long long)a.redo_cap;
    if (queued) a.redo_ws[1 + slot] = (unsigned long long)(t * a.N + n);
  } else {
    const int slot = atomicAdd(redo_n, 1);
    queued = slot < PIPE_QCAP;
    if (queued) redo_q[slot] = (uint32_t)ts << 10 | (uint32_t)tid;
  }
  if (queued) return false;
  if (OUT == 3 || OUT == 6) {
    // the annual sums have no per-step output to carry a marker: with a workspace the element
    // goes to the block's own queue; when that is full too the call reports it (bit 30 of
    // the gate word: the sums are incomplete, see MOD16_OUT_ANNUAL in the header)
    if (a.redo_cap > 0) {
      const int slot = atomicAdd(redo_n, 1);
      if (slot < PIPE_QCAP) {
        redo_q[slot] = (uint32_t)ts << 10 | (uint32_t)tid;
        return false;
      }
    }
    atomicOr(a.gate_flag, MOD16_ANNUAL_OVERFLOW);
    return false;
  }
  st_stream1(a.out[0] + t * a.N + n, __int_as_float(REDO_MARK));
  return false;
}

// The masked-driver rule (fast::nan_halves) for ONE element of the modes that cannot decide
// it from registers: re-reads the element's inputs.  Writes the NaN outputs and returns the
// gate predicate (0 / 1) when the rule decides the element, else -1.
template <int SEM, int OUT>
__device__ __forceinline__ int pipe_rule(const ForwardArgs<float>& a, int ts, float ok, float csl,
                                      float inv_dtmin, float tmin_bias, float vpd_off,
                                      float inv_dvpd) {
  const int64_t n = (int64_t)blockIdx.x * PIPE_TILE + (int)threadIdx.x;
  const int64_t t = (int64_t)blockIdx.y * a.t_chunk + ts;
  StepIn in = load_step(a, t, n);
  in.t_ann = 0.0f;
  fast::PftConstV<float> c;
  c.ok = ok; c.csl = csl; c.inv_dtmin = inv_dtmin; c.tmin_bias = tmin_bias;
  c.vpd_off = vpd_off; c.inv_dvpd = inv_dvpd;
  bool g_pos;
  const int h = fast::nan_halves<SEM>(in, c, a.has_rc != 0, g_pos);
  if (OUT == 2 || OUT == 4 ||
      !((OUT == 0 || OUT == 3 || OUT == 5 || OUT == 6) ? h != 0 : h == 3))
    return -1;
  if (OUT == 3) {
    annual_add(a, t, n, kAnnualNaN);
    return g_pos ? 1 : 0;
  }
  if (OUT == 6) {
    ((short*)a.out[0])[t * a.N + n] = (short)-32768;
    return g_pos ? 1 : 0;
  }
  const float qnan = __int_as_float(0x7fc00000);
  st_stream1(a.out[0] + t * a.N + n, qnan);
  if (OUT == 1) st_stream1(a.out[1] + t * a.N + n, qnan);
  return g_pos ? 1 : 0;
}

// (a value the compiler must compute once and keep: shared-window addresses would otherwise
// be re-derived from special registers in every iteration)
__device__ __forceinline__ uint32_t keep_u32(uint32_t x) {
  uint32_t y;
  asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

// EG: the instantiation for the per-element gate of the static path, with the guard at the
// zeros of the Tmin / VPD ramps (fast::ramp_zero_need).  A compile-time flag, so that the
// instantiations every other launch uses -- the headline one among them -- do not carry it:
// at this kernel's register limit the few extra live values cost ten instructions per pair
// of steps on every path.
template <int SEM, int OUT, int EXTRA, int FMT, bool EG = false>
__global__ void __launch_bounds__(PIPE_THREADS, 1)
forward_pipe_kernel(const __grid_constant__ ForwardArgs<float> a,
                    const __grid_constant__ PipeSrc src) {
  using fast::F2;
  using fast::M2;
  static_assert(FMT < FMT_MODIS || EXTRA == 0, "storage formats: the eleven per-step arrays");
  constexpr int NARR = pipe_n_arr(EXTRA);
  constexpr int NST = pipe_n_stages(EXTRA, FMT);
  constexpr bool SWN = (EXTRA & 1) != 0;   // sw_rad_night streamed
  constexpr bool PXT = (EXTRA & 2) != 0;   // temp_annual / pressure streamed
  constexpr bool QV = (EXTRA & 4) != 0;    // VPD from QV10M and surface pressure
  static_assert(!QV || EXTRA == 4, "VPD from QV10M: the eleven arrays + two pressures");
  constexpr uint32_t SUBB = pipe_eoff(FMT, NARR) * PIPE_TILE;   // bytes per sub-step
  constexpr uint32_t STAGEB = PIPE_SUB * SUBB;                  // bytes per stage
  // [stage][sub-step][array][PIPE_TILE]
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ SharedLut lut;
  __shared__ __align__(8) uint64_t full[PIPE_MAX_STAGES], empty[PIPE_MAX_STAGES];
  __shared__ uint32_t redo_q[PIPE_QCAP];  // (step within the chunk) << 10 | consumer thread
  __shared__ int redo_n;
  if (a.skip_if_flag && *((volatile int*)a.gate_flag) != 0) return;
  const int tid = threadIdx.x;
  const int64_t tile0 = (int64_t)blockIdx.x * PIPE_TILE;
  const int64_t t0 = (int64_t)blockIdx.y * a.t_chunk;
  int64_t t1 = t0 + a.t_chunk;
  if (t1 > a.T) t1 = a.T;
  const int nsteps = (int)(t1 - t0);
  const int niter = (nsteps + PIPE_SUB - 1) / PIPE_SUB;  // stage uses; the last may hold 1 step
  const int64_t left = a.N - tile0;
  const int n_live = (int)(left < PIPE_TILE ? left : PIPE_TILE);
  const uint32_t full0 = smem_u32(&full[0]), empty0 = smem_u32(&empty[0]);
  const uint32_t data0 = smem_u32(dyn_smem);
  const bool producer = tid >= PIPE_CONSUMERS;

  if (producer) {  // barriers, first fills (before the table staging: latency overlaps)
    if (tid == PIPE_CONSUMERS) {
      for (int s = 0; s < NST; ++s) {
        mbar_init(full0 + 8 * s, 1);
        mbar_init(empty0 + 8 * s, PIPE_CONSUMER_WARPS);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      redo_n = 0;
    }
    __syncwarp();
    for (int s = 0; s < NST && s < niter; ++s)
      pipe_issue<NARR, FMT>(src, nsteps - s * PIPE_SUB >= PIPE_SUB ? PIPE_SUB : 1,
                       data0 + s * STAGEB, full0 + 8 * s, t0 + s * PIPE_SUB, tile0, (uint32_t)n_live);
  }
  if (a.par_mode != 2) stage_lut<SEM>(a, lut);  // (contains __syncthreads)
  else __syncthreads();                         // (the barriers are visible to everyone)

  if (producer) {
    int stage = 0;
    uint32_t phase = 0;  // parity of the stage's previous use
#pragma unroll 1
    for (int i = NST; i < niter; ++i) {
      mbar_wait(empty0 + 8 * stage, phase);  // all fifteen warps have read the stage
      pipe_issue<NARR, FMT>(src, nsteps - i * PIPE_SUB >= PIPE_SUB ? PIPE_SUB : 1,
                       data0 + (uint32_t)stage * STAGEB, full0 + 8 * stage,
                       t0 + (int64_t)i * PIPE_SUB, tile0, (uint32_t)n_live);
      if (++stage == NST) { stage = 0; phase ^= 1u; }
    }
    return;
  }

  // ---- consumers ---------------------------------------------------------------------------
  // Per-thread element state of an iteration, one bit per lane (x: step t, y: step t + 1):
  //   DONE    the element needs nothing more from this thread (a pixel past the end of the
  //           tile, the missing second step of an odd chunk, an element already deferred)
  //   MASKED  static `_et` total of an element with a masked (NaN) driver: NaN goes out
  //           after the math, with the hot path's own gate predicate
  constexpr uint32_t DONE_X = 1u, DONE_Y = 2u, MASKED_X = 4u, MASKED_Y = 8u;
  const bool live = tid < n_live;
  const uint32_t flags0 = live ? 0u : (DONE_X | DONE_Y);
  bool any_pos = false;
  float gs_max = 0.0f;  // max of the gate quantity g_surf over this thread's elements (:343)
  // per-pixel state, all scalar: BPLUT column constants, pressure terms, annual window
  fast::PftConstV<F2> pc;
  fast::PixConstV<F2> px;
  PftConst c1;
  {
    const int64_t n = tile0 + tid;
    if (a.par_mode == 2) {  // per-pixel parameter arrays: same constants, built per thread
      double col[MOD16_N_PARAMS];
      load_column(a, live ? n : 0, col);
      c1 = fast::make_pft_const(col, SEM);
    } else {
      int code = 0;
      if (live && a.par_mode == 1) {
        code = (int)__ldg(a.pft + n);
        code = code < MOD16_N_PFT ? code : MOD16_N_PFT;  // out of range: the all-NaN column
      }
      c1 = lut.pc[code];
    }
    pc = fast::pair_pft(c1, c1);
    if (!PXT) {
      float p = 1e5f, ta = 0.f;
      if (live) {
        p = a.drv[MOD16_D_PRESSURE].at(0, n);   // per pixel or scalar (checked on the host)
        if (FMT != FMT_F32 && (a.prep & MOD16_PREP_ELEVATION)) p = pressure_from_elevation(p);
        ta = a.drv[MOD16_D_TEMP_ANNUAL].at(0, n);
      }
      const PixConst x1 = fast::make_pix_const(p, ta, c1, SEM);
      px = fast::pair_pix(x1, x1);
    }
  }
  // the masked-driver rule is decided from registers for the static total of a sane column
  const bool rule_in_regs = (OUT == 0 || OUT == 3 || OUT == 6) && SEM == SEM_STATIC &&
                            c1.ok != 0.0f;
  // MOD16_OUT_ANNUAL: this pixel's running sum of the current group (year), 2^-32 mm
  unsigned long long acc_run = 0ULL;
  int acc_g = -1;
  const float swn_scalar = a.drv[MOD16_D_SW_RAD_NIGHT].scalar;
  const int gate = a.gate;
  constexpr int n_out = (OUT == 3 || OUT == 6) ? 0 : ((OUT == 0 || OUT == 5) ? 1 : (OUT == 1 ? 2 : 6));
  constexpr bool NOMARK = OUT == 3 || OUT == 6;   // no float per-step output to carry a marker
  constexpr bool RULE = OUT != 2 && OUT != 4;   // outputs a NaN total decides without the double code
  // this thread's pixel in a row (float32 rows; storage formats index per array)
  const uint32_t my0 = keep_u32(FMT < FMT_MODIS ? data0 + (uint32_t)tid * 4u : data0);
  const bool clamp_vpd_n = FMT != FMT_F32 && (a.prep & MOD16_PREP_VPD_NIGHT_CLAMP) != 0;
  const uint32_t full_k = keep_u32(full0), empty_k = keep_u32(empty0);
  const int64_t t_stride = a.N;
  int64_t off = t0 * t_stride + tile0 + tid;        // running output offset

  // One use of a stage.  TWO: both time steps of the pair exist (always, except for the last
  // use of an odd chunk, which is peeled off the loop below so that the loop carries no
  // tail handling).  Stage order == PipeSrc order (see launch_pipe).  The inputs are pulled
  // into registers at once so that the stage can be re-filled while the steps are
  // evaluated.  Lane x of every pair is step t, lane y is step t + 1 (step t again when the
  // chunk is odd).
  auto consume = [&](auto two_tag, const int i) {
    constexpr bool TWO = decltype(two_tag)::value;
    const uint32_t stage = (uint32_t)i % (uint32_t)NST;
    const uint32_t phase = ((uint32_t)i / (uint32_t)NST) & 1u;
    mbar_wait(full_k + 8 * stage, phase);
    fast::StepInV<F2> in;
    {
      const uint32_t s0 = my0 + stage * STAGEB;
      const uint32_t s1 = TWO ? s0 + SUBB : s0;
#define M16_IN(k) F2{lds1(s0 + (k) * PIPE_ROWB), lds1(s1 + (k) * PIPE_ROWB)}
      if (FMT < FMT_MODIS) {
        in.lw_d = M16_IN(0);  in.lw_n = M16_IN(1);  in.sw_d = M16_IN(2);  in.alb = M16_IN(3);
        in.t_d = M16_IN(4);   in.t_n = M16_IN(5);   in.tmin = M16_IN(6);  in.vpd_d = M16_IN(7);
        in.vpd_n = M16_IN(8); in.fpar = M16_IN(9);  in.lai = M16_IN(10);
      } else {
#define M16_RAW(k, d) pipe_load<FMT, k>(s0, s1, (uint32_t)tid, a, d)
        in.lw_d = M16_RAW(0, MOD16_D_LW_NET_DAY);  in.lw_n = M16_RAW(1, MOD16_D_LW_NET_NIGHT);
        in.sw_d = M16_RAW(2, MOD16_D_SW_RAD_DAY);  in.alb = M16_RAW(3, MOD16_D_SW_ALBEDO);
        in.t_d = M16_RAW(4, MOD16_D_TEMP_DAY);     in.t_n = M16_RAW(5, MOD16_D_TEMP_NIGHT);
        in.tmin = M16_RAW(6, MOD16_D_TMIN);        in.vpd_d = M16_RAW(7, MOD16_D_VPD_DAY);
        in.vpd_n = M16_RAW(8, MOD16_D_VPD_NIGHT);  in.fpar = M16_RAW(9, MOD16_D_FPAR);
        in.lai = M16_RAW(10, MOD16_D_LAI);
#undef M16_RAW
      }
      if (QV) {   // MOD16.vpd (:639-644): 610.7 exp(17.38 Tc / (239 + Tc)) - qv ps / (0.622 + 0.379 qv)
        const F2 ps_d = M16_IN(11), ps_n = M16_IN(12);
        const F2 td = fast::celsius<F2>(in.t_d), tn = fast::celsius<F2>(in.t_n);
        const float k2 = (float)(17.38 * 1.4426950408889634);
        const F2 es_d = 610.7f * fast::vex2((k2 * td) * fast::vrcp(td + 239.0f));
        const F2 es_n = 610.7f * fast::vex2((k2 * tn) * fast::vrcp(tn + 239.0f));
        in.vpd_d = es_d - (in.vpd_d * ps_d) * fast::vrcp(fast::vfma(in.vpd_d, 0.379f, 0.622f));
        in.vpd_n = es_n - (in.vpd_n * ps_n) * fast::vrcp(fast::vfma(in.vpd_n, 0.379f, 0.622f));
      }
      if (FMT != FMT_F32 && clamp_vpd_n) {   // where(vpd_night < 0, 0, vpd_night): NaN stays
        in.vpd_n.x = in.vpd_n.x < 0.0f ? 0.0f : in.vpd_n.x;
        in.vpd_n.y = in.vpd_n.y < 0.0f ? 0.0f : in.vpd_n.y;
      }
      if (SWN) in.sw_n = M16_IN(pipe_slot_swn(EXTRA));
      else in.sw_n = F2{swn_scalar, swn_scalar};
      if (PXT) {  // per-step annual mean temperature and pressure: rebuild the pixel terms
        const F2 ta = M16_IN(pipe_slot_tann(EXTRA)), p = M16_IN(pipe_slot_p(EXTRA));
        px = fast::pair_pix(fast::make_pix_const(p.x, ta.x, c1, SEM),
                            fast::make_pix_const(p.y, ta.y, c1, SEM));
      }
#undef M16_IN
      in.rc_d = F2{0.f, 0.f}; in.rc_n = F2{0.f, 0.f};
    }
    __syncwarp();
    if (elect_one()) mbar_arrive(empty_k + 8 * stage);  // this warp has read the stage
    uint32_t flags = TWO ? flags0 : (flags0 | DONE_Y);

    // -- the domain of the fast formulas, tested while every input is still in a register
    const F2 tc_d = fast::celsius_b<F2>(in.t_d), tc_n = fast::celsius_b<F2>(in.t_n);
    F2 nan_probe;
    const M2 dom = fast::mand(px.ok, fast::in_domain<F2>(in, tc_d, tc_n, nan_probe));
    if (!(dom.x && dom.y)) {  // rare
#pragma unroll
      for (int j = 0; j < PIPE_SUB; ++j) {
        const uint32_t done_j = j == 0 ? DONE_X : DONE_Y;
        if ((j == 0 ? dom.x : dom.y) || (flags & done_j)) continue;
        // A NaN probe (or pressure) means a NaN driver of the masked-driver rule; a negative
        // (fill) temperature acts the same (its r_corr is NaN).  A masked pixel is masked at
        // every step: its warp must not stall on re-reads or on the double code.
        const float probe = j == 0 ? nan_probe.x : nan_probe.y;
        const float td = j == 0 ? in.t_d.x : in.t_d.y, tn = j == 0 ? in.t_n.x : in.t_n.y;
        const float pa = j == 0 ? px.pa.x : px.pa.y;
        if (RULE && (probe != probe || pa != pa || td < 0.0f || tn < 0.0f)) {
          if (rule_in_regs) {
            flags |= j == 0 ? MASKED_X : MASKED_Y;
            continue;
          }
          const int r = pipe_rule<SEM, OUT>(a, i * PIPE_SUB + j, c1.ok, c1.csl, c1.inv_dtmin,
                                            c1.tmin_bias, c1.vpd_off, c1.inv_dvpd);
          flags |= done_j;
          if (r >= 0) {
            any_pos |= r > 0;
            continue;
          }
        }
        flags |= done_j;
        any_pos |= pipe_defer<SEM, OUT>(a, i * PIPE_SUB + j, redo_q, &redo_n);
      }
    }

    fast::StepOutV<F2> r;
    M2 g_pos;
    F2 gs;
    // (before the step: a branch inside its arithmetic would change which products are fused)
    M2 rz = M2{false, false};
    if (EG) rz = fast::ramp_zero_need<F2, SEM>(in, pc, gate);
    M2 late = fast::step_core<F2, SEM, false>(in, tc_d, tc_n, pc, px, gate, r, g_pos, gs);
    if (EG) late = fast::mor(late, rz);
    F2 res[6];
    unsigned long long q_x = 0ULL, q_y = 0ULL;   // OUT_ANNUAL: the two steps' mass, fixed point
    int g_x = -1, g_y = -1;
    if (OUT == 3) {
      const F2 tcm = 0.5f * (tc_d + tc_n);
      const F2 lam = fast::vfma(tcm, -2361.0f, (float)(2.501e6 + 2361.0 * fast::kTcBias));  // :121
      const F2 mm = (fast::vmax(r.tot, 0.0f) * (86400.0f * 4294967296.0f)) * fast::vrcp(lam);
      q_x = __float2ull_rn(fminf(mm.x, 1.0e16f));
      q_y = __float2ull_rn(fminf(mm.y, 1.0e16f));
      g_x = __ldg(a.group + (int64_t)blockIdx.y * a.t_chunk + i * PIPE_SUB);
      g_y = TWO ? __ldg(a.group + (int64_t)blockIdx.y * a.t_chunk + i * PIPE_SUB + 1) : g_x;
    } else if (OUT == 5) {   // day-length weighted composite (verify2.py:112-115)
      const float* wd = a.w_day + (int64_t)blockIdx.y * a.t_chunk + i * PIPE_SUB;
      const float* wn = a.w_night + (int64_t)blockIdx.y * a.t_chunk + i * PIPE_SUB;
      const F2 w_d = F2{__ldg(wd), TWO ? __ldg(wd + 1) : 0.0f};
      const F2 w_n = F2{__ldg(wn), TWO ? __ldg(wn + 1) : 0.0f};
      res[0] = fast::vfma(r.tot_day, w_d, r.tot_night * w_n);
    } else if (OUT == 6) {
      res[0] = r.tot * a.out_inv_scale;
    } else if (OUT == 4) {
      res[0] = r.tot;       res[1] = r.day.pt;
      res[2] = r.day.sat;   res[3] = r.day.unsat;
      res[4] = r.night.sat; res[5] = r.night.unsat;
    } else if (OUT == 0) {
      res[0] = r.tot;
    } else if (OUT == 1) {
      res[0] = r.tot_day; res[1] = r.tot_night;
    } else {
      res[0] = r.day.canopy;   res[1] = r.day.soil;   res[2] = r.day.trans;
      res[3] = r.night.canopy; res[4] = r.night.soil; res[5] = r.night.trans;
    }
    if (OUT == 3 && flags == 0u && !(late.x || late.y) && g_x == acc_g && g_y == acc_g) {
      gs_max = fmaxf(gs_max, fmaxf(gs.x, gs.y));
      acc_run += q_x + q_y;
    } else if (OUT == 6 && flags == 0u && !(late.x || late.y)) {
      gs_max = fmaxf(gs_max, fmaxf(gs.x, gs.y));
      short* o16 = (short*)a.out[0] + off;   // (in the domain: finite)
      int q0, q1;
      asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(q0) : "f"(fmaxf(res[0].x, -32767.0f)));
      asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(q1) : "f"(fmaxf(res[0].y, -32767.0f)));
      o16[0] = (short)q0;
      if (TWO) o16[t_stride] = (short)q1;
    } else if (OUT != 3 && flags == 0u && !(late.x || late.y)) {
      gs_max = fmaxf(gs_max, fmaxf(gs.x, gs.y));
#pragma unroll
      for (int k = 0; k < n_out; ++k) {
        st_stream1(a.out[k] + off, res[k].x);
        if (TWO) st_stream1(a.out[k] + off + t_stride, res[k].y);
      }
    } else {  // rare
#pragma unroll
      for (int j = 0; j < PIPE_SUB; ++j) {
        if (flags & (j == 0 ? DONE_X : DONE_Y)) continue;
        unsigned long long q_j = j == 0 ? q_x : q_y;
        if (flags & (j == 0 ? MASKED_X : MASKED_Y)) {
          // (g_surf > 0 from Tmin, VPD, T, P of the hot path; NaN > 0 is false)
          if (OUT == 3) q_j = kAnnualNaN;
          else if (OUT == 6) ((short*)a.out[0])[off + j * t_stride] = (short)-32768;
          else st_stream1(a.out[0] + off + j * t_stride, __int_as_float(0x7fc00000));
          any_pos |= j == 0 ? g_pos.x : g_pos.y;
        } else if (j == 0 ? late.x : late.y) {
          any_pos |= pipe_defer<SEM, OUT>(a, i * PIPE_SUB + j, redo_q, &redo_n);
          continue;
        } else {
          any_pos |= j == 0 ? g_pos.x : g_pos.y;
          if (OUT == 6)
            ((short*)a.out[0])[off + j * t_stride] = to_i16(j == 0 ? res[0].x : res[0].y, 1.0f);
#pragma unroll
          for (int k = 0; k < n_out; ++k)
            st_stream1(a.out[k] + off + j * t_stride, j == 0 ? res[k].x : res[k].y);
        }
        if (OUT == 3) {   // into the running sum; a new group flushes the old one
          const int g_j = j == 0 ? g_x : g_y;
          if (g_j != acc_g) {
            if (acc_g >= 0 && acc_g < a.n_groups && acc_run != 0ULL)
              atomicAdd((unsigned long long*)a.out[0] + (int64_t)acc_g * a.N + tile0 + tid,
                        acc_run);
            acc_g = g_j;
            acc_run = 0ULL;
          }
          acc_run += q_j;
        }
      }
    }
    off += PIPE_SUB * t_stride;
  };
  const int nfull = nsteps / PIPE_SUB;
#pragma unroll 1
  for (int i = 0; i < nfull; ++i) consume(std::true_type(), i);
  if (nsteps & 1) consume(std::false_type(), nfull);
  any_pos |= gs_max > 0.0f;
  if (OUT == 3 && live && acc_g >= 0 && acc_g < a.n_groups && acc_run != 0ULL)
    atomicAdd((unsigned long long*)a.out[0] + (int64_t)acc_g * a.N + tile0 + tid, acc_run);
  // the block-local deferred pass: consumers only (named barrier 1), one element each; when
  // the block's queue overflowed, every thread also looks for markers in its own column
  if (a.redo_cap == 0 || NOMARK) {
    asm volatile("bar.sync 1, %0;" ::"n"(PIPE_CONSUMERS) : "memory");
    const int total = redo_n < PIPE_QCAP ? redo_n : PIPE_QCAP;
    for (int e = tid; e < total; e += PIPE_CONSUMERS) {
      const uint32_t q = redo_q[e];
      any_pos |= fix_exact_one<SEM, OUT>(a, t0 + (int64_t)(q >> 10), tile0 + (q & 1023u));
    }
    if (!NOMARK && redo_n > PIPE_QCAP && live) {
      for (int ts = 0; ts < nsteps; ++ts) {
        const int64_t o = (t0 + ts) * a.N + tile0 + tid;
        if (__float_as_int(((volatile float*)a.out[0])[o]) == REDO_MARK)
          any_pos |= fix_exact_one<SEM, OUT>(a, t0 + ts, tile0 + tid);
      }
    }
  }
  if (a.track_gate) {
    if (__any_sync(0xffffffffu, any_pos) && (threadIdx.x & 31) == 0) *a.gate_flag = 1;
  }
}

template <int SEM, int OUT, int EXTRA, int FMT, bool EG = false>
static int launch_pipe_inst(const ForwardArgs<float>& a, const PipeSrc& src, dim3 grid,
                            cudaStream_t st) {
  constexpr size_t smem = (size_t)pipe_n_stages(EXTRA, FMT) * PIPE_SUB *
                          pipe_eoff(FMT, pipe_n_arr(EXTRA)) * PIPE_TILE;
  // (the attribute is per function AND per device: set before every launch, it is cheap)
  M16_CUDA_CHECK(cudaFuncSetAttribute(forward_pipe_kernel<SEM, OUT, EXTRA, FMT, EG>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  forward_pipe_kernel<SEM, OUT, EXTRA, FMT, EG><<<grid, PIPE_THREADS, smem, st>>>(a, src);
  M16_CUDA_CHECK(cudaGetLastError());
  return MOD16_OK;
}

// Instantiated per semantics and part in mod16_forward_pipe_{static,instance}[_b].cu (parallel
// builds).
// EXTRA 2 / 3 (time-varying annual temperature / pressure) exist for the summed and the
// day / night outputs; `separate=True` with such drivers takes the strided kernel.
// Storage formats (FMT > 0) exist for the summed and the day / night outputs of the eleven
// per-step arrays.
// PART splits the instantiations over translation units (build time): 0 = the float32
// layouts with the three classic output modes, 1 = storage formats, loader conversions and
// the epilogue output modes, 2 = part 0 again with the ramp-zero guard (EG; static semantics,
// per-element gate).  Returns MOD16_ERR_FLAGS for a variant of another part.
template <int SEM, int PART>
int launch_pipe_sem(const ForwardArgs<float>& a, const PipeSrc& src, int out_mode, int extra,
                    int fmt, dim3 grid, cudaStream_t st) {
#define M16_PIPE(O, E, F) \
  if (out_mode == O && extra == E && fmt == F) \
    return launch_pipe_inst<SEM, O, E, F>(a, src, grid, st)
  if constexpr (PART == 0) {
    M16_PIPE(0, 0, 0); M16_PIPE(0, 1, 0); M16_PIPE(0, 2, 0); M16_PIPE(0, 3, 0);
    M16_PIPE(1, 0, 0); M16_PIPE(1, 1, 0); M16_PIPE(1, 2, 0); M16_PIPE(1, 3, 0);
    M16_PIPE(2, 0, 0); M16_PIPE(2, 1, 0);
  } else if constexpr (PART == 2) {
#define M16_PIPE_EG(O, E) \
  if (out_mode == O && extra == E && fmt == 0) \
    return launch_pipe_inst<SEM, O, E, 0, true>(a, src, grid, st)
    M16_PIPE_EG(0, 0); M16_PIPE_EG(0, 1); M16_PIPE_EG(0, 2); M16_PIPE_EG(0, 3);
    M16_PIPE_EG(1, 0); M16_PIPE_EG(1, 1); M16_PIPE_EG(1, 2); M16_PIPE_EG(1, 3);
    M16_PIPE_EG(2, 0); M16_PIPE_EG(2, 1);
#undef M16_PIPE_EG
  } else {
    M16_PIPE(0, 0, 1); M16_PIPE(0, 0, 2); M16_PIPE(0, 0, 3);
    M16_PIPE(1, 0, 1); M16_PIPE(1, 0, 2); M16_PIPE(1, 0, 3);
    M16_PIPE(0, 4, 1); M16_PIPE(1, 4, 1);
    if constexpr (SEM == SEM_STATIC) { M16_PIPE(3, 0, 0); }
    M16_PIPE(4, 0, 0); M16_PIPE(5, 0, 0); M16_PIPE(6, 0, 0);
  }
#undef M16_PIPE
  return MOD16_ERR_FLAGS;
}

inline bool pipe_has_variant(int out_mode, int extra, int fmt = 0) {
  // annual sums, potential terms: the plain tile layout
  if (out_mode >= 3) return extra == 0 && fmt == 0;
  if (extra == 4) return fmt == 1 && out_mode < 2;   // VPD from QV10M: float32 + conversions
  return fmt == 0 ? (out_mode < 2 || extra < 2) : (out_mode < 2 && extra == 0);
}

}  // namespace mod16
