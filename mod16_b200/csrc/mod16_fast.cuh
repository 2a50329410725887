// float32 fast path of the MOD16 forward model.
//
// Same physics as mod16_math.cuh, rearranged for the B200 instruction budget (SURVEY.md
// H4): every chain of divisions is merged into one reciprocal, the aerodynamic
// resistances are carried as conductances (1/r_a = 1/r_r + g, so no resistance is ever
// formed and the reference's inf/NaN detours through r_h = 1/0 disappear), powers become
// MUFU sequences (x^1.75 via rsqrt/sqrt, rh^(vpd/beta) via lg2/ex2, rh^4 by squaring) and
// everything that depends only on the pixel (pressure terms, the annual-temperature
// window, BPLUT-derived ramp slopes) is hoisted out of the time loop.
//
// Parity with the float64 reference on the same float32 inputs is kept by three guards,
// each a rare, divergent detour:
//   * domain guard  -- an element whose temperatures or pressure are outside the physical
//     range (NaN, -9999 fill, ...) is evaluated by the op-for-op code in DOUBLE, so fill /
//     NaN masks are the reference's by construction;
//   * RH guard      -- near the f_wet jump (rh = 0.7) and near rh = 0+ (where
//     rh^(vpd/beta) is ill-conditioned) rh is recomputed in double;
//   * G guard       -- when |G| and 0.39|A| are within rounding of each other and of
//     opposite sign (the only discontinuous switch of the soil heat flux), the predicate
//     is re-evaluated in double.
// Citations are file:line under /root/reference/mod16/__init__.py.
#pragma once
#include "mod16_math.cuh"

namespace mod16 {
namespace fast {

// ---- MUFU wrappers (approximate, flush-to-zero; host versions for tests/hostmath) -------
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float rcp(float x) {
  float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float ex2(float x) {
  float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float rsq(float x) {
  float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float sqr(float x) {
  float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
#else
inline float rcp(float x) { return 1.0f / x; }
inline float ex2(float x) { return exp2f(x); }
inline float lg2(float x) { return log2f(x); }
inline float rsq(float x) { return 1.0f / sqrtf(x); }
inline float sqr(float x) { return sqrtf(x); }
#endif

// ---- constants derived from one BPLUT column (computed in double, stored as float) -----
struct PftConst {
  float tmin_close, tmin_open, inv_dtmin;  // ramp up   (:325)
  float vpd_open, vpd_close, inv_dvpd;     // ramp down (:326)
  float gl_sh, gl_wv, g_cut, csl;
  float rbl_min, rbl_max, rbl_slope;       // r_tot = rbl_max - rbl_slope*(vpd_close - vpd)
  float inv_beta;
  float tann_lo_hi, tann_lo_lo;            // 273.15 + tmin_{open|close} as a float pair
};

M16_HD PftConst make_pft_const(const double v[11], int sem) {
  PftConst c;
  c.tmin_close = (float)v[0]; c.tmin_open = (float)v[1];
  c.inv_dtmin = (float)(1.0 / (v[1] - v[0]));
  c.vpd_open = (float)v[2]; c.vpd_close = (float)v[3];
  c.inv_dvpd = (float)(1.0 / (v[3] - v[2]));
  c.gl_sh = (float)v[4]; c.gl_wv = (float)v[5]; c.g_cut = (float)v[6]; c.csl = (float)v[7];
  c.rbl_min = (float)v[8]; c.rbl_max = (float)v[9];
  c.rbl_slope = (float)((v[9] - v[8]) / (v[3] - v[2]));
  c.inv_beta = (float)(1.0 / v[10]);
  double lo = 273.15 + (sem == SEM_STATIC ? v[1] : v[0]);
  c.tann_lo_hi = (float)lo;
  c.tann_lo_lo = (float)(lo - (double)c.tann_lo_hi);
  return c;
}

// Params<double> for the op-for-op fallback, rebuilt from the float constants is NOT
// possible without loss, so the fallback receives the original double column.

// ---- per-pixel state hoisted out of the time loop --------------------------------------
struct PixConst {
  float pa;        // 0.348444 * (P / 100)                     (:410)
  float p101;      // 101300 / P                               (:292)
  float inv_p101;  // P / 101300
  float cpp;       // Cp * P / eps  (gamma = cpp / lambda)     (:1289)
  bool window;     // T_ann < 298.15 and T_ann >(=) 273.15 + tmin_{open|close}  (:230-233)
  bool p_ok;       // pressure inside the physical domain
};

M16_HD PixConst make_pix_const(float p, float t_ann, const PftConst& c, int sem) {
  PixConst x;
  x.pa = 0.348444f * (p * 0.01f);
  x.p101 = 101300.0f * rcp(p);
  x.inv_p101 = p * (1.0f / 101300.0f);
  x.cpp = p * (float)(kCp / kEps);
  // the window is decided in double: a float32 T_ann must fall on the same side of
  // 298.15 and of 273.15 + tmin as it does in the float64 reference
  const double ta = (double)t_ann;
  const double lo = (double)c.tann_lo_hi + (double)c.tann_lo_lo;
  x.window = (ta < 273.15 + 25) && (sem == SEM_STATIC ? (ta > lo) : (ta >= lo));
  x.p_ok = (p >= 1.0e3f) && (p <= 2.0e5f);
  return x;
}

// rare: relative humidity in double from float inputs (:279-281 / :670-673)
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
static float rh_exact(float t, float vpd, int sem, bool& wet) {
  double es = svp<double>((double)t);
  double avp = es - (double)vpd;
  double rh = avp / es;
  if (sem == SEM_STATIC) rh = rh < 0.0 ? 0.0 : rh;
  else rh = avp < 0.0 ? 0.0 : (rh > 1.0 ? 1.0 : rh);
  wet = !(rh < 0.7);
  return (float)rh;
}

// rare: the |G| > 0.39 |A| predicate in double (:239)
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
static bool g_exceeds_exact(float t, float sw, float alb, float lw, bool night) {
  double a = night ? (double)lw : (double)sw * (1.0 - (double)alb) + (double)lw;
  double g = (4.73 * ((double)t - 273.15)) - 20.87;
  return fabs(g) > (0.39 * fabs(a));
}

// t - 273.15 with the constant split into a float pair, so that the Celsius temperature
// carries no 6e-6 K bias from rounding 273.15 to float
M16_HD float celsius(float t) {
  const float hi = 273.15f;
  const float lo = (float)(273.15 - (double)hi);
  return (t - hi) - lo;
}

struct HalfIn {
  float t, vpd, a_can, a_soil, rc;
};

// One half-day, fast form.  STATIC night never transpires (its gate is never open).
template <int SEM, bool DAY>
M16_HD Half<float> half_fast(const HalfIn& h, float tmin, float fpar, float lai,
                             const PftConst& c, const PixConst& x, bool has_rc, int gate,
                             bool& g_pos) {
  const float tc = celsius(h.t);
  // saturation vapour pressure and its reciprocal (:1365-1367)
  const float xarg = (17.27f * tc) * rcp(tc + 237.3f);
  const float es = 610.8f * ex2(xarg * 1.4426950408889634f);
  const float inv_es = rcp(es);
  float rh;
  if (SEM == SEM_STATIC) {
    rh = (es - h.vpd) * inv_es;      // :280
    rh = rh < 0.0f ? 0.0f : rh;      // :281
  } else {
    const float avp = es - h.vpd;    // :671-673
    rh = avp * inv_es;
    rh = avp < 0.0f ? 0.0f : (rh > 1.0f ? 1.0f : rh);
  }
  bool wet = !(rh < 0.7f);
  if (fabsf(rh - 0.7f) < 3.0e-5f || (rh > 0.0f && rh < 2.0e-3f))
    rh = rh_exact(h.t, h.vpd, SEM, wet);
  const float rh2 = rh * rh;
  const float fw = wet ? rh2 * rh2 : 0.0f;                       // :282
  const float d2 = 239.0f + tc;
  const float s = (17.38f * 239.0f) * es * rcp(d2 * d2);         // :1397
  const float lam = fmaf(-2361.0f, tc, 2.501e6f);                // :121
  const float inv_lam = rcp(lam);
  const float gamma = x.cpp * inv_lam;                           // :1289
  // (T / 293.15)^(+-1.75) from one rsqrt and one sqrt (:292)
  const float q = h.t * (1.0f / 293.15f);
  const float rs = rsq(q);          // q^-1/2
  const float r4 = sqr(rs);         // q^-1/4
  float r_corr, inv_rc;
  if (has_rc) {
    r_corr = h.rc;
    inv_rc = rcp(h.rc);
  } else {
    r_corr = x.p101 * (q * q * r4);
    inv_rc = x.inv_p101 * (rs * rs * rs * r4);
  }
  // air density numerator X (rho = X / T, :409-412) and 1/r_r = 4 sigma T^4 / (Cp X)
  const float xr = fmaf(-100.0f * rh, fmaf(0.00252f, tc, -0.020582f), x.pa);
  const float inv_t = rs * rs * (1.0f / 293.15f);
  const float rho_cp = (float)kCp * xr * inv_t;
  const float t2 = h.t * h.t;
  const float inv_rr = (float)(4.0 * kSigma / kCp) * (t2 * t2) * rcp(xr);
  const float omfw = 1.0f - fw;
  const float rcv = rho_cp * h.vpd;  // rho Cp VPD, shared by the three components

  Half<float> o;
  // -- wet canopy: e = fw (s Ac + rho Cp fpar VPD / r_aw) / (s + gamma r_e / r_aw),
  //    1/r_aw = 1/r_r + gl_sh lai fw, r_e / r_aw = (1/r_aw) / (gl_wv lai fw)   (:305-320)
  if (SEM == SEM_STATIC) {
    const float lf = lai * fw;
    const float ira = fmaf(c.gl_sh, lf, inv_rr);
    const float ce = c.gl_wv * lf;
    const float num = fmaf(s, h.a_can, rcv * fpar * ira);
    const float den = fmaf(s, ce, gamma * ira);
    const float e = (fw * num) * ce * rcp(den);
    o.canopy = (lf <= 1e-7f) ? 0.0f : e;
  } else {
    const float fw2 = fw == 0.0f ? 1e-7f : fw;                   // :934
    const float lai2 = lai == 0.0f ? 1e-7f : lai;                // :935
    const float lf = lai2 * fw2;
    const float ira = fmaf(c.gl_sh, lf, inv_rr);
    const float ce = c.gl_wv * lf;
    const float num = fw2 * fmaf(s, h.a_can, rcv * fpar * ira);
    const float den = fmaf(s, ce, gamma * ira);
    const float e = num < 0.0f ? 0.0f : (num * ce * rcp(den)) * inv_lam;  // :959
    o.canopy = (fw2 <= 1e-7f || lai2 <= 1e-7f) ? 0.0f : e;               // :961
  }

  // -- transpiration (:323-349, :1233-1258)
  g_pos = false;
  if (DAY || SEM == SEM_INSTANCE) {
    float gs = 0.0f;
    if (DAY) {
      const float tm = celsius(tmin);
      const float m_t = tm >= c.tmin_open ? 1.0f
                        : (tm < c.tmin_close ? 0.0f : (tm - c.tmin_close) * c.inv_dtmin);
      const float m_v = h.vpd >= c.vpd_close ? 0.0f
                        : (h.vpd < c.vpd_open ? 1.0f
                                              : 1.0f - (h.vpd - c.vpd_open) * c.inv_dvpd);
      gs = (c.csl * m_t * m_v) * inv_rc;
    } else {
      gs = 0.0f * inv_rc;
    }
    g_pos = gs > 0.0f;
    const float sc = gs + c.g_cut * inv_rc;
    const float gbl = c.gl_sh * lai * omfw;
    const bool ok = (lai > 0.0f) && (omfw > 0.0f);
    const float prod = ok ? gbl * sc : 1e-7f;    // g_canopy = prod / tot (:334-337)
    const float tot = ok ? gbl + sc : 1.0f;
    const float ird = inv_rr + c.gl_sh;          // 1 / r_a_dry (:340)
    const float a_can = (SEM == SEM_INSTANCE && h.a_can < 0.0f) ? 0.0f : h.a_can;  // :1251
    const float num = omfw * fmaf(s, a_can, rcv * fpar * ird);
    const float den = fmaf(s + gamma, prod, gamma * ird * tot);
    float tr = num * prod * rcp(den);
    if (SEM == SEM_STATIC) {
      if (gate == GATE_CLOSED) tr = 0.0f;
      else if (gate == GATE_ELEMENT) tr = g_pos ? tr : 0.0f;
      o.trans = tr;
    } else {
      const float g_can = ok ? prod * rcp(tot) : 1e-7f;
      o.trans = g_can <= 1e-7f ? 0.0f : tr * inv_lam;            // :1258
    }
  } else {
    o.trans = 0.0f;
  }

  // -- bare soil (:357-376): 1/r_as = 1/r_tot + 1/r_r, r_tot/r_as = 1 + r_tot/r_r
  {
    const float rbl = h.vpd <= c.vpd_open
                          ? c.rbl_min
                          : (h.vpd >= c.vpd_close
                                 ? c.rbl_max
                                 : fmaf(-c.rbl_slope, c.vpd_close - h.vpd, c.rbl_max));
    const float r_t = rbl * inv_rc;
    const float ira = fmaf(r_corr, rcp(rbl), inv_rr);
    const float num = fmaf(s, h.a_soil, rcv * (1.0f - fpar) * ira);
    const float den = fmaf(gamma, fmaf(r_t, inv_rr, 1.0f), s);
    const float qn = num * rcp(den);
    const float k = h.vpd * c.inv_beta;
    float pw = ex2(k * lg2(rh));                                  // rh^(vpd/beta)
    pw = (k == 0.0f || rh == 1.0f) ? 1.0f : pw;                   // pow(x, 0) = pow(1, y) = 1
    if (SEM == SEM_STATIC) {
      o.soil = qn * fmaf(omfw, pw, fw);
    } else {
      const float e_sat = qn * fw, e_uns = qn * omfw;
      float e = e_sat < 0.0f ? 0.0f : e_sat;                      // :858
      e = e + (e_uns < 0.0f ? 0.0f : e_uns * pw);                 // :859-861
      o.soil = e * inv_lam;                                       // :864
    }
  }
  return o;
}

struct StepIn {  // one pixel-timestep of float drivers
  float lw_d, lw_n, sw_d, sw_n, alb, t_d, t_n, t_ann, tmin, vpd_d, vpd_n, p, fpar, lai;
  float rc_d, rc_n;
};

struct StepOut {  // components and totals of one pixel-timestep
  Half<float> day, night;
  float tot_day, tot_night, tot;  // :380 / :792 per half, and day + night (:193)
};

// op-for-op evaluation in double (domain-guard fallback); `col` is the BPLUT column
template <int SEM>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
static void pixel_step_exact(const StepIn& in, const double col[11], bool has_rc, int gate,
                             StepOut& o, bool& g_pos) {
  Drivers<double> d;
  d.lw_d = in.lw_d; d.lw_n = in.lw_n; d.sw_d = in.sw_d; d.sw_n = in.sw_n; d.alb = in.alb;
  d.t_d = in.t_d; d.t_n = in.t_n; d.t_ann = in.t_ann; d.tmin = in.tmin;
  d.vpd_d = in.vpd_d; d.vpd_n = in.vpd_n; d.p = in.p; d.fpar = in.fpar; d.lai = in.lai;
  d.rc_d = in.rc_d; d.rc_n = in.rc_n;
  Params<double> q;
  q.tmin_close = col[0]; q.tmin_open = col[1]; q.vpd_open = col[2]; q.vpd_close = col[3];
  q.gl_sh = col[4]; q.gl_wv = col[5]; q.g_cuticular = col[6]; q.csl = col[7];
  q.rbl_min = col[8]; q.rbl_max = col[9]; q.beta = col[10];
  q.tann_lo = 273.15 + (SEM == SEM_STATIC ? col[1] : col[0]);
  Half<double> dd, nn;
  pixel_step<double, SEM>(d, q, has_rc, gate, dd, nn, g_pos);
  o.day.canopy = (float)dd.canopy; o.day.soil = (float)dd.soil; o.day.trans = (float)dd.trans;
  o.night.canopy = (float)nn.canopy; o.night.soil = (float)nn.soil;
  o.night.trans = (float)nn.trans;
  // totals are formed in double: garbage-in elements (finite fills) cancel catastrophically
  const double td = half_total<double, SEM>(dd), tn = half_total<double, SEM>(nn);
  o.tot_day = (float)td; o.tot_night = (float)tn; o.tot = (float)(td + tn);
}

// NaN-propagating 3-input minimum (FMNMX3.NAN on sm_100a)
M16_HD float min3_nan(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  float r; asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
  if (a != a || b != b || c != c) return NAN;
  return fminf(a, fminf(b, c));
#endif
}

// Domain of the fast path: every per-step driver is a number above the fill range
// (> -9000: rejects NaN and -9999 / -32768 style fills) and the temperatures are such
// that every denominator of the rearranged formulas is positive (40 K < T < 1000 K).
// Pressure is checked once per pixel (PixConst::p_ok).
M16_HD bool in_domain(const StepIn& in) {
  float m = min3_nan(in.lw_d, in.lw_n, in.sw_d);
  m = min3_nan(m, in.alb, in.tmin);
  m = min3_nan(m, in.vpd_d, in.vpd_n);
  m = min3_nan(m, in.fpar, in.lai);
  m = min3_nan(m, in.sw_n, in.sw_n);
  const float tl = min3_nan(in.t_d, in.t_n, in.t_n);
  return (m > -9000.0f) && (tl > 40.0f) && (in.t_d < 1000.0f) && (in.t_n < 1000.0f);
}

// One pixel-timestep, both halves, with the guards.
// `col(k)` returns BPLUT parameter k of this pixel in double (only called on the fallback).
template <int SEM, typename ColFn>
M16_HD void pixel_step_fast(const StepIn& in, const PftConst& c, const PixConst& x,
                            ColFn col, bool has_rc, int gate, StepOut& o, bool& g_pos) {
  if (!(x.p_ok && in_domain(in))) {
    double cc[11];
    for (int k = 0; k < 11; ++k) cc[k] = col(k);
    pixel_step_exact<SEM>(in, cc, has_rc, gate, o, g_pos);
    return;
  }
  const float oma = 1.0f - in.alb;
  const float a_d = fmaf(in.sw_d, oma, in.lw_d);            // :225
  const float a_n = in.lw_n;                                 // :226
  const float a_n_can = fmaf(in.sw_n, oma, in.lw_n);         // :272 (night: sw_rad_night)
  // soil heat flux (:229-241)
  const bool cond = x.window && ((in.t_d - in.t_n) >= 5.0f);
  float g_d = cond ? fmaf(4.73f, celsius(in.t_d), -20.87f) : 0.0f;
  float g_n = cond ? fmaf(4.73f, celsius(in.t_n), -20.87f) : 0.0f;
  const float lim_d = 0.39f * a_d, lim_n = 0.39f * a_n;
  bool ex_d = fabsf(g_d) > fabsf(lim_d), ex_n = fabsf(g_n) > fabsf(lim_n);
  if (cond) {
    // the switch is discontinuous only when G and A have opposite signs
    if (g_d * lim_d < 0.0f &&
        fabsf(fabsf(g_d) - fabsf(lim_d)) <= fmaf(1e-5f, fabsf(lim_d), 2e-5f))
      ex_d = g_exceeds_exact(in.t_d, in.sw_d, in.alb, in.lw_d, false);
    if (g_n * lim_n < 0.0f &&
        fabsf(fabsf(g_n) - fabsf(lim_n)) <= fmaf(1e-5f, fabsf(lim_n), 2e-5f))
      ex_n = g_exceeds_exact(in.t_n, 0.0f, 0.0f, in.lw_n, true);
  }
  g_d = ex_d ? lim_d : g_d;                                                   // :239
  g_n = ex_n ? lim_n : g_n;
  g_d = ((a_d - g_d < 0.0f) && (a_d > 0.0f)) ? a_d : g_d;                     // :245-247
  g_n = ((a_d > 0.0f) && ((a_n - g_n) < (-0.5f * a_d))) ? fmaf(0.5f, a_d, a_n) : g_n;
  const float omf = 1.0f - in.fpar;
  HalfIn hd, hn;
  hd.t = in.t_d; hd.vpd = in.vpd_d; hd.a_can = in.fpar * a_d; hd.rc = in.rc_d;
  hd.a_soil = omf * (a_d - g_d);                                              // :253
  hn.t = in.t_n; hn.vpd = in.vpd_n; hn.a_can = in.fpar * a_n_can; hn.rc = in.rc_n;
  hn.a_soil = omf * (a_n - g_n);                                              // :254
  bool unused;
  o.day = half_fast<SEM, true>(hd, in.tmin, in.fpar, in.lai, c, x, has_rc, gate, g_pos);
  o.night = half_fast<SEM, false>(hn, in.tmin, in.fpar, in.lai, c, x, has_rc, gate, unused);
  o.tot_day = half_total<float, SEM>(o.day);
  o.tot_night = half_total<float, SEM>(o.night);
  o.tot = o.tot_day + o.tot_night;
}

}  // namespace fast
}  // namespace mod16
