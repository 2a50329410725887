/*
 * mod16_b200 -- C ABI of the B200-native MOD16 evapotranspiration forward path.
 *
 * The reference (arthur-e/MOD16) is pure Python/numpy and has NO foreign-function
 * interface; the boundary it exposes for this path is the Python callable contract
 *
 *     MOD16._evapotranspiration(params, *14 drivers, f_wet, tiny, r_corr_list)
 *                                              mod16/__init__.py:196-201
 *     MOD16._et(params, *14 drivers, ...)      mod16/__init__.py:163-168
 *     MOD16.evapotranspiration(self, *14 drivers, f_wet, separate)
 *                                              mod16/__init__.py:675-682
 *
 * and the objective loops built on `_et` (mod16/sensitivity.py:94-96,
 * mod16/calibration.py:224-227).  The entry points below are what a binding for those
 * calls needs: plain pointers, sizes and flag words; no Python, numpy or torch types.
 * `mod16_b200/_abi.py` binds them with ctypes; INTEGRATION.md shows the stub a maintainer
 * of the reference would add.
 *
 * Conventions
 *  - every data pointer is a DEVICE pointer (cudaMalloc / torch CUDA tensor storage)
 *    unless the field comment says "host";
 *  - per-step arrays are (T, N) row-major: element (t, n) of driver k lives at
 *    ptr[t * t_stride + n * n_stride] (strides in ELEMENTS; 0 broadcasts);
 *  - all calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL is the
 *    legacy default stream); nothing is allocated behind the caller's back;
 *  - return value: MOD16_OK or a negative error code; mod16_last_cuda_error() returns the
 *    cudaError_t of the last failed CUDA call on this thread's library state.
 */
#ifndef MOD16_B200_H
#define MOD16_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOD16_ABI_VERSION 12

#define MOD16_N_PARAMS 11  /* MOD16.required_parameters, mod16/__init__.py:152-155 */
#define MOD16_N_PFT 13     /* BPLUT arrays are indexed by PFT code 0..12, mod16/utils.py:104 */
#define MOD16_N_DRIVERS 14 /* positional drivers of the three entry points above */

/* parameter row order == MOD16.required_parameters */
enum {
  MOD16_P_TMIN_CLOSE = 0, MOD16_P_TMIN_OPEN, MOD16_P_VPD_OPEN, MOD16_P_VPD_CLOSE,
  MOD16_P_GL_SH, MOD16_P_GL_WV, MOD16_P_G_CUTICULAR, MOD16_P_CSL, MOD16_P_RBL_MIN,
  MOD16_P_RBL_MAX, MOD16_P_BETA
};

/* driver order == positional order of the reference signatures */
enum {
  MOD16_D_LW_NET_DAY = 0, MOD16_D_LW_NET_NIGHT, MOD16_D_SW_RAD_DAY, MOD16_D_SW_RAD_NIGHT,
  MOD16_D_SW_ALBEDO, MOD16_D_TEMP_DAY, MOD16_D_TEMP_NIGHT, MOD16_D_TEMP_ANNUAL,
  MOD16_D_TMIN, MOD16_D_VPD_DAY, MOD16_D_VPD_NIGHT, MOD16_D_PRESSURE, MOD16_D_FPAR,
  MOD16_D_LAI
};

/* error codes */
#define MOD16_OK 0
#define MOD16_ERR_NULL (-1)     /* a required pointer is NULL */
#define MOD16_ERR_SHAPE (-2)    /* T or N < 0, or a stride is negative */
#define MOD16_ERR_FLAGS (-3)    /* unknown / inconsistent flag word */
#define MOD16_ERR_CUDA (-4)     /* a CUDA call failed: see mod16_last_cuda_error() */
#define MOD16_ERR_PARAMS (-5)   /* parameter source is inconsistent */
#define MOD16_ERR_WORKSPACE (-6)/* workspace too small */
#define MOD16_ERR_LAYOUT (-7)   /* raw-format layout the fused kernel cannot stream: convert
                                   the drivers to float32 first (no silent slow path) */

/* flag word of mod16_forward_* ------------------------------------------------------- */
/* semantics: which of the reference's two forward paths is reproduced */
#define MOD16_SEM_STATIC 0x0u    /* _evapotranspiration/_et: W m-2, no clamps (:196-382) */
#define MOD16_SEM_INSTANCE 0x1u  /* evapotranspiration(): kg m-2 s-1, clamps (:675-793) */
#define MOD16_SEM_MASK 0xFu
/* outputs written (each a dense (T, N) array) */
#define MOD16_OUT_SUM 0x00u        /* out[0] = day + night                    (_et) */
#define MOD16_OUT_DAY_NIGHT 0x10u  /* out[0] = day, out[1] = night */
#define MOD16_OUT_COMPONENTS 0x20u /* out[0..2] = day canopy, soil, transpiration;
                                      out[3..5] = night  (separate=True, :789-790) */
#define MOD16_OUT_ANNUAL 0x30u     /* float32 static path only, no per-step output: out[0] is
                                      a zero-initialised uint64 (n_groups, N) array that
                                      receives, per pixel and group of time steps (year),
                                      the evapotranspiration sum [mm] of the calibration's
                                      precipitation constraint -- max(LE, 0) * 86400 /
                                      lambda((T_day + T_night) / 2), mod16/calibration.py:
                                      783-790, :838-842 -- in units of 2^-32 mm.  Integer
                                      accumulation: the sums do not depend on the launch
                                      geometry or on the order of the atomic adds.  A NaN
                                      element adds 2^54: a sum >= 2^54 reads as NaN, like
                                      the reference's `x.sum()`.  Groups of at most 1000
                                      steps.  gate_flag is required; if more elements need
                                      the double-precision pass than the workspace (plus 512
                                      per block) holds, MOD16_ANNUAL_OVERFLOW is OR-ed into
                                      *gate_flag and the sums are incomplete: rerun with a
                                      larger workspace. */
#define MOD16_ANNUAL_OVERFLOW 0x40000000
#define MOD16_OUT_POTENTIAL 0x40u  /* float32, streamable layouts (else MOD16_ERR_LAYOUT):
                                      out[0] = day + night total as MOD16_OUT_SUM, and the
                                      potential terms in W m-2: out[1] potential
                                      transpiration (day, mod16/__init__.py:547-602); out[2],
                                      out[3] evaporation from the saturated soil fraction and
                                      POTENTIAL evaporation from the unsaturated fraction,
                                      day (potential_soil_evaporation, :450-545); out[4],
                                      out[5] the same at night */
#define MOD16_OUT_COMPOSITE 0x50u  /* float32, streamable layouts: out[0] = day * day_weight[t] +
                                      night * night_weight[t] -- the day-length weighted
                                      composite of tests/verification/verify2.py:112-115,
                                      `pred_day * updown * 8 * 3600 + pred_night * (24 -
                                      updown) * 8 * 3600` [kg m-2 (8 days)-1 from the instance
                                      path's kg m-2 s-1]; the weights are float32[T] */
#define MOD16_OUT_SUM_I16 0x60u    /* float32, streamable layouts: out[0] is int16 (T, N),
                                      round(total / out_scale) saturated to +-32767; a NaN
                                      total is -32768 (the way the MOD16 product stores its
                                      layers: scaled integers with a fill value).  Needs
                                      gate_flag; queue overflow is reported like
                                      MOD16_OUT_ANNUAL (MOD16_ANNUAL_OVERFLOW) */
#define MOD16_OUT_MASK 0xF0u
/* the static path's batch-global transpiration gate `if np.any(g_surf > 0)` (:343) */
#define MOD16_GATE_GLOBAL 0x000u     /* call-exact: optimistic pass + device-side fix-up */
#define MOD16_GATE_ELEMENT 0x100u    /* predicate per element (a loop of scalar calls).  With it
                                        g_surf = 0+ is a switch of each element (the cuticular
                                        transpiration appears): the fast paths hand elements
                                        within rounding of a zero of the Tmin / VPD ramps to
                                        the op-for-op code.  float32, static path: the
                                        pipelined kernel has that guard for plain float32
                                        drivers and OUT_SUM / _DAY_NIGHT / _COMPONENTS; with
                                        this gate mod16_forward_raw_f32 (convert, then
                                        mod16_forward_f32) and the outputs MOD16_OUT_ANNUAL /
                                        _POTENTIAL / _COMPOSITE / _SUM_I16 (epilogues of the
                                        batch-global `_et`) answer MOD16_ERR_LAYOUT */
#define MOD16_GATE_OPEN 0x200u       /* always compute transpiration */
#define MOD16_GATE_CLOSED 0x300u     /* transpiration = 0 */
#define MOD16_GATE_OPEN_TRACK 0x400u /* GATE_OPEN, and OR the predicate into *gate_flag */
#define MOD16_GATE_FIXUP 0x500u      /* only if *gate_flag == 0: recompute with GATE_CLOSED
                                        (multi-GPU: all-reduce(max) the flag in between) */
#define MOD16_GATE_MASK 0xF00u
/* kernel selection (testing / profiling): default picks the vectorised kernel when the
   layout allows 128-bit loads, else the strided one */
#define MOD16_FORCE_GENERIC 0x1000u
#define MOD16_FORCE_MIRROR 0x2000u  /* op-for-op transcription instead of the rearranged math */

typedef struct mod16_driver {
  const void* ptr;   /* device array of float / double, or NULL: use `scalar` */
  int64_t t_stride;  /* elements between consecutive time steps (0: constant in time) */
  int64_t n_stride;  /* elements between consecutive pixels (0: constant in space) */
  double scalar;     /* value used when ptr == NULL */
} mod16_driver_t;

typedef struct mod16_drivers {
  mod16_driver_t drv[MOD16_N_DRIVERS];
  mod16_driver_t r_corr[2]; /* optional `r_corr_list` (day, night), :291-294 */
  int32_t has_r_corr;
  /* mod16_forward_f64 only: bit k set = drv[k].ptr is a FLOAT32 array (strides in float32
     elements), widened exactly inside the kernel.  The reference's stock call -- float32
     drivers from the archive with the float64 BPLUT, which numpy promotes to float64 -- then
     needs no float64 copies of the drivers.  0 for every other entry. */
  uint32_t f32_mask;
} mod16_drivers_t;

/*
 * Parameter source, one of
 *  (a) pft != NULL: per-pixel land-cover code (0..12) gathers a column of `table`
 *      (the BPLUT dict of mod16/utils.py:81-117 as a dense 11 x 13 array);
 *  (b) arrays[k] != NULL for all k: per-pixel parameter vectors, the `(1 x N)` arrays of
 *      the static path's `params` list (:175-178); array_stride[k] is 1, or 0 for a scalar
 *      held in device memory;
 *  (c) neither: one parameter set for every pixel, table[k][0].
 */
typedef struct mod16_params {
  const uint8_t* pft;                         /* device, N codes, or NULL */
  double table[MOD16_N_PARAMS][MOD16_N_PFT];  /* host values, copied at launch */
  const void* arrays[MOD16_N_PARAMS];         /* device, float / double as the entry */
  int64_t array_stride[MOD16_N_PARAMS];
} mod16_params_t;

typedef struct mod16_out {
  void* out[6];       /* device (T, N) row-major, dense; see MOD16_OUT_* */
  int32_t* gate_flag; /* device int; required by GATE_GLOBAL / OPEN_TRACK / FIXUP */
  /* Optional device scratch of the float32 path (8-byte aligned): the rare elements that must
     be re-evaluated in double precision are queued here (8 bytes each after an 8-byte
     counter) and evaluated by a second launch.  mod16_workspace_bytes() is the minimum worth
     passing; 8 bytes per 64 elements of the call holds every data set measured so far
     (physical drivers queue 5e-5 of their elements, unmasked fill values or negative VPD
     more).  A queue that fills up costs speed only: the surplus is marked in out[0] and
     found again by a third launch.  NULL / tiny: a small per-block queue, then evaluation
     inside the main kernel (same results, slow).  Contents are scratch. */
  void* workspace;
  int64_t workspace_bytes;
  /* MOD16_OUT_ANNUAL: device int32[T], group (year) index of every time step (indices
     outside [0, n_groups) are skipped) */
  const int32_t* group;
  int32_t n_groups;
  int32_t reserved;
  /* MOD16_OUT_COMPOSITE: device float32[T], seconds of daylight / of night each time step
     stands for (e.g. updown * 8 * 3600 and (24 - updown) * 8 * 3600) */
  const float* day_weight;
  const float* night_weight;
  double out_scale;   /* MOD16_OUT_SUM_I16: physical value of one count (e.g. 0.1 W m-2) */
} mod16_out_t;

/*
 * Drivers in their storage formats: the conversions the reference's
 * loaders perform on the host before the model is called are applied inside the fused kernel
 * instead, so that a pixel-timestep costs 24-40 bytes of traffic instead of 48:
 *   fPAR  uint8 percent        `fpar /= 100`                 (mod16/calibration.py:422)
 *   LAI   uint8 tenths         `lai /= 10`                   (mod16/calibration.py:423)
 *   albedo, and optionally every other per-step driver, int16 with scale / offset
 *   vpd_night = where(vpd_night < 0, 0, vpd_night)           (mod16/calibration.py:401)
 *   pressure = MOD16.air_pressure(elevation)                 (mod16/__init__.py:415-447,
 *                                                             mod16/calibration.py:405-408)
 * physical value = raw * scale + offset, evaluated in float32 (one fused multiply-add); a raw
 * value equal to `fill` (when has_fill) is a missing value: NaN.  Two layouts are streamed:
 *   MODIS : sw_albedo int16, fpar uint8, lai uint8, the eight other per-step drivers float32
 *   PACKED: sw_albedo and the eight other per-step drivers int16, fpar uint8, lai uint8
 * (all-float32 drivers are accepted too: the plain kernel plus the `prep` conversions).
 * Requirements: static or instance semantics, OUT_SUM or OUT_DAY_NIGHT, per-step arrays
 * (T, N) with unit pixel stride and 16-byte aligned rows (N % 16 == 0 for the uint8 rows),
 * sw_rad_night a scalar, temp_annual / pressure per pixel or scalar (float32), no
 * r_corr_list.  Anything else returns MOD16_ERR_LAYOUT.
 */
#define MOD16_T_F32 0
#define MOD16_T_I16 1
#define MOD16_T_U8 2
#define MOD16_PREP_VPD_NIGHT_CLAMP 0x1u /* vpd_night < 0 -> 0 (a NaN stays NaN) */
#define MOD16_PREP_ELEVATION 0x2u       /* drv[MOD16_D_PRESSURE] holds elevation in metres */
#define MOD16_PREP_VPD_FROM_QV 0x4u     /* drv[MOD16_D_VPD_DAY / _NIGHT] hold the 10 m water
                                           vapour mixing ratio QV10M, `ps` the surface
                                           pressure: VPD = MOD16.vpd(qv10m, ps, temp)
                                           = 610.7 exp(17.38 Tc / (239 + Tc)) - qv ps /
                                           (0.622 + 0.379 qv) (mod16/__init__.py:605-644,
                                           mod16/calibration.py:393-400); float32 arrays */

typedef struct mod16_raw_driver {
  const void* ptr;   /* device array of `dtype`, or NULL: use `scalar` (a physical value) */
  int64_t t_stride;  /* ELEMENTS of `dtype` between consecutive time steps */
  int64_t n_stride;  /* ... between consecutive pixels */
  double scalar;
  int32_t dtype;     /* MOD16_T_* */
  int32_t has_fill;  /* integer types: `fill` marks a missing value */
  double scale, offset;
  int64_t fill;
} mod16_raw_driver_t;

typedef struct mod16_raw_drivers {
  mod16_raw_driver_t drv[MOD16_N_DRIVERS];
  uint32_t prep;     /* MOD16_PREP_* */
  uint32_t reserved;
  mod16_raw_driver_t ps[2]; /* MOD16_PREP_VPD_FROM_QV: surface pressure, day and night */
} mod16_raw_drivers_t;

int mod16_abi_version(void);
int64_t mod16_workspace_bytes(void); /* recommended size of mod16_out_t.workspace */
int mod16_last_cuda_error(void);
/* Which kernel the last mod16_forward_* call of this thread launched (for tests and
   profiling): 0 none yet; MOD16_KERNEL_PIPE | (variant << 8) the bulk-copy pipeline
   (variant bit 0: sw_rad_night streamed, bit 1: temp_annual / pressure streamed);
   MOD16_KERNEL_STRIDED the float32 strided kernel; MOD16_KERNEL_FAST64 the float64 entry's
   kernel; MOD16_KERNEL_EXACT the op-for-op kernel (MOD16_FORCE_MIRROR, either precision). */
#define MOD16_KERNEL_PIPE 1
#define MOD16_KERNEL_STRIDED 2
#define MOD16_KERNEL_EXACT 3
#define MOD16_KERNEL_FAST64 4  /* float64 entry: rearranged algebra in double */
/* (MOD16_KERNEL_FAST64: bit 8 of the value is set when the launch took the canonical-layout
   kernel -- every time-varying source a (T, N) array of one width and one time stride, which
   is what the reference's loaders produce -- and clear for the general kernel, which accepts
   any broadcast layout and arrays of both widths.  Same results, bit for bit.) */
int mod16_last_forward_kernel(void);
const char* mod16_error_string(int code);

/* K1 on storage-format drivers (see mod16_raw_drivers_t); float32 arithmetic and outputs. */
int mod16_forward_raw_f32(const mod16_raw_drivers_t* drivers, const mod16_params_t* params,
                          const mod16_out_t* out, int64_t T, int64_t N, uint32_t flags,
                          void* stream);

/* K1 -- fused forward model over a (T, N) block of pixel-timesteps. */
int mod16_forward_f32(const mod16_drivers_t* drivers, const mod16_params_t* params,
                      const mod16_out_t* out, int64_t T, int64_t N, uint32_t flags,
                      void* stream);
int mod16_forward_f64(const mod16_drivers_t* drivers, const mod16_params_t* params,
                      const mod16_out_t* out, int64_t T, int64_t N, uint32_t flags,
                      void* stream);

/* K2 -- objective batch: D parameter draws evaluated on the SAME (T, N) block of drivers,
 * reduced on the device to the sufficient statistics of the callers' objectives.
 * Replaces the loops `for X in param_sweep: yhat = MOD16._et(X, *drivers); nash_sutcliffe(
 * yhat, obs)` (mod16/sensitivity.py:94-96) and mod17's BlackBoxLikelihood calling
 * `model(params, *drivers)` once per MCMC proposal (mod16/calibration.py:224-227, 907-909).
 *
 *   pred(d; t, n) = `_et(draws[d], *drivers)` (MOD16_SEM_STATIC; SEM_INSTANCE: day + night
 *                   of evapotranspiration())
 *   sse[d]   = sum over valid (t, n) of ((pred - observed) * weights)^2
 *   count[d] = number of valid (t, n): observed, weights and pred all non-NaN
 * so that RMSD = sqrt(sse / count) (= sqrt(nanmean(((pred - obs) w)^2))) and
 * NSE = 1 - sse / sum((obs - mean(obs))^2) with unit weights.
 *
 * The static path's batch-global transpiration gate `if np.any(g_surf > 0)` (:343) is
 * evaluated PER DRAW over EVERY element of the (T, N) block, with or without an observation or
 * weight, as one reference call per draw does: `_et` sees whatever drivers it is handed, and
 * missing observations are masked afterwards in the score (mod16/calibration.py:224-227; the
 * default calibration run leaves them in place, :355-370, the k-fold branch :895-901 and
 * mod16/sensitivity.py:179-186 remove those rows first) (MOD16_GATE_GLOBAL); MOD16_GATE_OPEN /
 * MOD16_GATE_CLOSED fix it.  Elements whose drivers make the prediction NaN for every draw
 * (masked pixels) enter no sum, like missing observations.  r_corr_list is not supported.
 * All reductions are deterministic (no floating-point atomics).
 */
typedef struct mod16_objective {
  const double* draws;     /* device (D, 11) row-major, order of required_parameters */
  mod16_driver_t observed; /* (T, N) broadcastable like a driver; NaN = no observation */
  mod16_driver_t weights;  /* residual weights; ptr NULL: the scalar (1.0 = unweighted) */
  double* sse;             /* device (D,), overwritten */
  double* count;           /* device (D,), overwritten */
  int32_t* gate;           /* device (D,) or NULL: 1 where any g_surf > 0 for the draw
                              (over every element, see above) */
  void* workspace;         /* device scratch, 256-byte aligned */
  int64_t workspace_bytes; /* >= mod16_objective_workspace_bytes(T, N, D, f32) */
  double* raw;             /* device (D, 5) or NULL: the un-gated statistics sse_open, n_open,
                              sse_closed, n_closed, any(g_surf > 0) of THIS call's elements.
                              A site-sharded multi-GPU batch all-reduces them (sum) and then
                              applies each draw's gate over all shards (any > 0). */
} mod16_objective_t;

int64_t mod16_objective_workspace_bytes(int64_t T, int64_t N, int64_t D, int is_f32);
int mod16_objective_batch_f32(const mod16_drivers_t* drivers, const mod16_objective_t* obj,
                              int64_t T, int64_t N, int64_t D, uint32_t flags, void* stream);
int mod16_objective_batch_f64(const mod16_drivers_t* drivers, const mod16_objective_t* obj,
                              int64_t T, int64_t N, int64_t D, uint32_t flags, void* stream);

/* K4 -- component methods and physics helpers of the reference's public API, elementwise.
 * `in[k]` are the positional arguments of the reference signature (file:line under
 * mod16/__init__.py), each broadcastable to (T, N); bit k of `present` says that argument k
 * was supplied -- required arguments must be present, optional ones (marked [..]) fall back
 * to the reference's defaults.  `params` are the instance parameters of MOD16(params), in
 * required_parameters order (unused by the free functions).  Outputs are dense (T, N).
 *
 *   op                              arguments                                       out
 *   SVP                    :1340    temp_k                                          Pa
 *   SVP_SLOPE              :1370    temp_k, [s]                                     Pa K-1
 *   LHV                    :121     temp_k                                          J kg-1
 *   PSYCHROMETRIC          :1261    pressure, temp_k                                Pa K-1
 *   AIR_DENSITY            :385     temp_k, pressure, rhumidity                     kg m-3
 *   RHUMIDITY              :647     temp_k, vpd                                     [0, 1]
 *   VPD                    :605     qv10m, pressure, tmean                          Pa
 *   AIR_PRESSURE           :415     elevation_m                                     Pa
 *   RADIATION_NET          :1293    sw_rad, sw_albedo, temp_k                       W m-2
 *   SOIL_HEAT_FLUX         :1055    rad_net_day, rad_net_night, temp_day,
 *                                   temp_night, temp_annual                         2 x W m-2
 *   RADIATION_SOIL         :963     lw_net_day, lw_net_night, sw_rad_day,
 *                                   sw_rad_night, sw_albedo, temp_day, temp_night,
 *                                   temp_annual, fpar                               2 x W m-2
 *   POTENTIAL_SOIL_EVAPORATION :450 pressure, temp_k, vpd, fpar, rad_soil, [r_corr],
 *                                   [lhv], [rhumidity], [f_wet]                     2 x W m-2
 *   EVAPORATION_SOIL       :795     (same)                                          kg m-2 s-1
 *   EVAPORATION_WET_CANOPY :866     pressure, temp_k, vpd, lai, fpar, rad_canopy,
 *                                   [lhv], [rhumidity], [f_wet]; opt[1] = tiny      kg m-2 s-1
 *   SURFACE_CONDUCTANCE    :1121    tmin, vpd_day                                   m s-1
 *   TRANSPIRATION          :1152    pressure, temp_k, vpd, lai, fpar, rad_canopy, tmin,
 *                                   [r_corr], [lhv], [rhumidity], [f_wet];
 *                                   opt[0] = daytime (0 / 1), opt[1] = tiny         kg m-2 s-1
 *   POTENTIAL_TRANSPIRATION :547    lw_net, sw_rad, sw_albedo, pressure, temp_k, vpd,
 *                                   fpar, [rhumidity], [f_wet]; opt[0] = alpha      W m-2
 *   DAYLENGTH_COMPOSITE  tests/verification/verify2.py:112-115
 *                                   pred_day, pred_night [kg m-2 s-1], daylength [h];
 *                                   opt[0] = days per composite (8):
 *                                   day dl n 3600 + night (24 - dl) n 3600          kg m-2 (n d)-1
 */
enum {
  MOD16_OP_SVP = 0, MOD16_OP_SVP_SLOPE, MOD16_OP_LHV, MOD16_OP_PSYCHROMETRIC,
  MOD16_OP_AIR_DENSITY, MOD16_OP_RHUMIDITY, MOD16_OP_VPD, MOD16_OP_AIR_PRESSURE,
  MOD16_OP_RADIATION_NET, MOD16_OP_SOIL_HEAT_FLUX, MOD16_OP_RADIATION_SOIL,
  MOD16_OP_POTENTIAL_SOIL_EVAPORATION, MOD16_OP_EVAPORATION_SOIL,
  MOD16_OP_EVAPORATION_WET_CANOPY, MOD16_OP_SURFACE_CONDUCTANCE, MOD16_OP_TRANSPIRATION,
  MOD16_OP_POTENTIAL_TRANSPIRATION, MOD16_OP_DAYLENGTH_COMPOSITE, MOD16_OP_COUNT
};
#define MOD16_MAX_OP_INPUTS 12

typedef struct mod16_elementwise {
  int32_t op;        /* MOD16_OP_* */
  uint32_t present;  /* bit k: in[k] was supplied */
  mod16_driver_t in[MOD16_MAX_OP_INPUTS];
  double params[MOD16_N_PARAMS];
  double opt[2];
  void* out[2];      /* device, dense (T, N), float / double as the entry; out[1] may be NULL */
} mod16_elementwise_t;

int mod16_elementwise_f32(const mod16_elementwise_t* ew, int64_t T, int64_t N, void* stream);
int mod16_elementwise_f64(const mod16_elementwise_t* ew, int64_t T, int64_t N, void* stream);

/* K3 -- per-pixel time aggregates (annual sums of the calibration's precipitation
 * constraint, mod16/calibration.py:783-790: mass = x * 86400 / lhv, negatives dropped,
 * summed per year).  term(t, n) = x * scale / divisor (divisor NULL: x * scale);
 * out[g][n] = sum of the terms of the steps whose group index is g; total[g] = sum over n.
 * Accumulation is in double, in time order, by the one thread that owns the pixel (no
 * floating-point atomics; `total` is a fixed-order second pass): results are bit-reproducible
 * run to run.  NaN terms propagate like numpy's sum unless SKIP_NAN. */
#define MOD16_AGG_CLAMP_NEG 0x1u /* negative terms count as 0 (calibration.py:784) */
#define MOD16_AGG_SKIP_NAN 0x2u  /* NaN terms count as 0 (nansum) */

typedef struct mod16_aggregate {
  const void* x;        /* device (T, N) dense, float / double as the entry */
  const void* divisor;  /* device (T, N) dense, or NULL */
  double scale;
  const int32_t* group; /* device (T,): group of each step, outside [0, n_groups) = skipped;
                           NULL: one group */
  int32_t n_groups;
  int32_t reserved;
  double* out;          /* device (n_groups, N), overwritten */
  double* total;        /* device (n_groups,), overwritten; or NULL */
} mod16_aggregate_t;

int mod16_time_aggregate_f32(const mod16_aggregate_t* agg, int64_t T, int64_t N, uint32_t flags,
                             void* stream);
int mod16_time_aggregate_f64(const mod16_aggregate_t* agg, int64_t T, int64_t N, uint32_t flags,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOD16_B200_H */
