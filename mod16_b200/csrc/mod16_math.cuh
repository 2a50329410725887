// Per-element physics of the MOD16 forward model, op-for-op ("mirror") form.
//
// One function, templated on the real type and on the semantics, evaluates both
// half-days of one pixel-timestep.  It follows the reference expression by expression,
// including the association order and every `np.where` direction, so that inf / NaN
// intermediates (e.g. r_h = 1/0 when f_wet == 0) arise and are masked exactly where the
// reference's do.  Citations are file:line under /root/reference/mod16/__init__.py.
//
// The header is __host__ __device__ clean so that tests/hostmath can compile the very
// same source with g++ and compare it with the oracle in a container without a GPU; the
// product only ever runs it on the device.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define M16_HD __host__ __device__ __forceinline__
#else
#define M16_HD inline
#endif

namespace mod16 {

enum Semantics { SEM_STATIC = 0, SEM_INSTANCE = 1 };
enum Gate { GATE_OPEN = 0, GATE_ELEMENT = 1, GATE_CLOSED = 2 };

// :107-110
constexpr double kSigma = 5.67e-8;
constexpr double kCp = 1013.0;
constexpr double kEps = 0.622;
constexpr double kTiny = 1e-7;

template <typename R>
struct Params {  // one column of the BPLUT, order of MOD16.required_parameters (:152-155)
  R tmin_close, tmin_open, vpd_open, vpd_close, gl_sh, gl_wv, g_cuticular, csl, rbl_min,
      rbl_max, beta;
  // lower bound of the annual-temperature window, always in double so that the f32
  // kernels take the same side of the threshold as the f64 reference on the same inputs:
  // static  273.15 + params[1] (tmin_OPEN, strict >)   :233
  // instance 273.15 + tmin_close (>=)                  :1106
  double tann_lo;
};

template <typename R>
struct Drivers {  // one pixel-timestep, positional order of the reference signatures
  R lw_d, lw_n, sw_d, sw_n, alb, t_d, t_n, t_ann, tmin, vpd_d, vpd_n, p, fpar, lai;
  R rc_d, rc_n;  // optional r_corr_list entries (:291-294)
};

template <typename R>
struct Half {
  R canopy, soil, trans;
  // MOD16_OUT_POTENTIAL [W m-2]: potential transpiration (:547-602, daytime; 0 at night) and
  // the two terms of potential_soil_evaporation (:505-545 = :371, :373): evaporation from the
  // saturated fraction and POTENTIAL evaporation from the unsaturated fraction
  R pt, sat, unsat;
};

template <typename R> M16_HD R m_exp(R x);
template <> M16_HD double m_exp<double>(double x) { return exp(x); }
template <> M16_HD float m_exp<float>(float x) { return expf(x); }
template <typename R> M16_HD R m_pow(R x, R y);
template <> M16_HD double m_pow<double>(double x, double y) { return pow(x, y); }
template <> M16_HD float m_pow<float>(float x, float y) { return powf(x, y); }
template <typename R> M16_HD R m_abs(R x) { return x < R(0) ? -x : (x == R(0) ? R(0) : x); }
template <> M16_HD double m_abs<double>(double x) { return fabs(x); }
template <> M16_HD float m_abs<float>(float x) { return fabsf(x); }

// ---- physics helpers (also exported one by one through mod16_helper_*) -----------------

// :121
template <typename R> M16_HD R latent_heat(R t) {
  return (R(2.501) - R(0.002361) * (t - R(273.15))) * R(1e6);
}
// :1365-1367
template <typename R> M16_HD R svp(R t) {
  R tc = t - R(273.15);
  return R(1e3 * 0.6108) * m_exp<R>((R(17.27) * tc) / (tc + R(237.3)));
}
// :1397
template <typename R> M16_HD R svp_slope(R t, R es) {
  R d = R(239.0) + t - R(273.15);
  return R(17.38 * 239.0) * es / (d * d);
}
// :1288-1290
template <typename R> M16_HD R psychrometric(R p, R t) {
  return (R(kCp) * p) / (latent_heat<R>(t) * R(kEps));
}
// :409-412
template <typename R> M16_HD R air_density(R t, R p, R rh) {
  return (R(0.348444) * (p / R(100)) -
          (rh * R(100)) * (R(0.00252) * (t - R(273.15)) - R(0.020582))) / t;
}
// :670-673 (instance path: clamped to [0, 1])
template <typename R> M16_HD R rhumidity_clamped(R t, R vpd) {
  R es = svp<R>(t);
  R avp = es - vpd;
  R rh = avp / es;
  return avp < R(0) ? R(0) : (rh > R(1) ? R(1) : rh);
}
// :292
template <typename R> M16_HD R r_correction(R p, R t) {
  return (R(101300) / p) * m_pow<R>(t / R(293.15), R(1.75));
}
// :282
template <typename R> M16_HD R f_wet_of(R rh) {
  R r2 = rh * rh;
  return rh < R(0.7) ? R(0) : r2 * r2;
}
// mod17.linear_constraint (EXTERNAL; call sites :325, :1148)
template <typename R> M16_HD R ramp_up(R x, R lo, R hi) {
  return x >= hi ? R(1) : (x < lo ? R(0) : (x - lo) / (hi - lo));
}
// mod17.linear_constraint(..., 'reversed') (call sites :326, :1149)
template <typename R> M16_HD R ramp_down(R x, R lo, R hi) {
  return x >= hi ? R(0) : (x < lo ? R(1) : R(1) - (x - lo) / (hi - lo));
}
// :357-361, :527-531
template <typename R> M16_HD R r_total(R vpd, const Params<R>& q) {
  return vpd <= q.vpd_open
             ? q.rbl_min
             : (vpd >= q.vpd_close
                    ? q.rbl_max
                    : q.rbl_max - ((q.rbl_max - q.rbl_min) * (q.vpd_close - vpd)) /
                                      (q.vpd_close - q.vpd_open));
}
// :298-299
template <typename R> M16_HD R r_radiative(R rho, R t) {
  return (rho * R(kCp)) / (R(4 * kSigma) * (t * t * t));
}
// :639-644 driver preparation
template <typename R> M16_HD R vpd_from_qv(R qv, R p, R t) {
  R tc = t - R(273.15);
  R avp = (qv * p) / (R(0.622) + (R(0.379) * qv));
  return R(610.7) * m_exp<R>((R(17.38) * tc) / (R(239) + tc)) - avp;
}
// :446-447 driver preparation
template <typename R> M16_HD R air_pressure(R z) {
  const double rate = 9.80665 / (0.0065 * (8.3143 / 28.9644e-3));
  R ratio = R(1) - ((R(0.0065) * z) / R(288.15));
  return R(101325.0) * m_pow<R>(ratio, R(rate));
}

// ---- soil heat flux and radiation reaching the soil (:225-254 == :1033-1053) ------------
template <typename R, int SEM>
M16_HD void soil_radiation(const Drivers<R>& d, const Params<R>& q, R& as_d, R& as_n) {
  R a_d = d.sw_d * (R(1) - d.alb) + d.lw_d;  // :225
  R a_n = d.lw_n;                             // :226
  double tann = (double)d.t_ann;
  bool window = SEM == SEM_STATIC ? (tann > q.tann_lo) : (tann >= q.tann_lo);
  bool cond = (tann < 273.15 + 25) && window && ((d.t_d - d.t_n) >= R(5));  // :230-234
  R g_d = cond ? (R(4.73) * (d.t_d - R(273.15))) - R(20.87) : R(0);        // :237
  R g_n = cond ? (R(4.73) * (d.t_n - R(273.15))) - R(20.87) : R(0);
  g_d = m_abs<R>(g_d) > (R(0.39) * m_abs<R>(a_d)) ? R(0.39) * a_d : g_d;   // :239
  g_n = m_abs<R>(g_n) > (R(0.39) * m_abs<R>(a_n)) ? R(0.39) * a_n : g_n;
  g_d = ((a_d - g_d < R(0)) && (a_d > R(0))) ? a_d : g_d;                  // :245-247
  g_n = ((a_d > R(0)) && ((a_n - g_n) < (R(-0.5) * a_d))) ? a_n + (R(0.5) * a_d)
                                                           : g_n;          // :248-252
  as_d = (R(1) - d.fpar) * (a_d - g_d);                                    // :253
  as_n = (R(1) - d.fpar) * (a_n - g_n);                                    // :254
}

// ---- one half-day ---------------------------------------------------------------------
// Returns the three components; `g_pos` reports the static gate predicate g_surf > 0.
template <typename R, int SEM>
M16_HD Half<R> half_day(bool daytime, R t, R vpd, R sw, R lw, R a_soil, R rc_in,
                        bool has_rc, const Drivers<R>& d, const Params<R>& q, int gate,
                        bool& g_pos) {
  const R cp = R(kCp), eps = R(kEps), tiny = R(kTiny);
  R a_net = sw * (R(1) - d.alb) + lw;  // :272 / :756
  R a_can = d.fpar * a_net;            // :274 / :758
  R es = svp<R>(t);                    // :279
  R rh;
  if (SEM == SEM_STATIC) {
    rh = (es - vpd) / es;              // :280
    rh = rh < R(0) ? R(0) : rh;        // :281  (no upper clamp)
  } else {
    R avp = es - vpd;                  // :671-673
    rh = avp / es;
    rh = avp < R(0) ? R(0) : (rh > R(1) ? R(1) : rh);
  }
  R fw = f_wet_of<R>(rh);              // :282 / :764 (caller's f_wet is ignored)
  R s = svp_slope<R>(t, es);           // :284
  R lam = latent_heat<R>(t);           // :286
  R gamma = (cp * d.p) / (lam * eps);  // :288
  R r_corr = has_rc ? rc_in : r_correction<R>(d.p, t);  // :291-294
  R rho = air_density<R>(t, d.p, rh);  // :296
  R r_r = r_radiative<R>(rho, t);      // :298-299

  Half<R> o;
  o.pt = daytime ? (R(1.26) * (s * a_can) * (R(1) - fw)) / (s + gamma) : R(0);  // :602
  // -- wet canopy evaporation
  if (SEM == SEM_STATIC) {
    R r_h = R(1) / (q.gl_sh * d.lai * fw);  // :305
    R r_e = R(1) / (q.gl_wv * d.lai * fw);  // :308
    R r_aw = (r_h * r_r) / (r_h + r_r);     // :311
    R e = (fw * ((s * a_can) + ((rho * cp * d.fpar * vpd * R(1)) / r_aw))) /
          (s + (((d.p * cp * r_e) * R(1)) / (lam * eps * r_aw)));  // :314-319
    o.canopy = (d.lai * fw <= tiny) ? R(0) : e;                   // :320
  } else {
    R fw2 = fw == R(0) ? fw + tiny : fw;            // :934
    R lai2 = d.lai == R(0) ? d.lai + tiny : d.lai;  // :935
    R r_h = R(1) / (q.gl_sh * lai2 * fw2);          // :943
    R r_e = R(1) / (q.gl_wv * lai2 * fw2);          // :945
    R r_aw = (r_h * r_r) / (r_h + r_r);             // :951
    R numer = fw2 * ((s * a_can) + ((rho * cp * d.fpar * vpd * R(1)) / r_aw));  // :952-953
    R denom = s + (((d.p * cp * r_e) * R(1)) / (lam * eps * r_aw));             // :954-955
    R evap = numer < R(0) ? R(0) : (numer / denom) / lam;                       // :959
    o.canopy = (fw2 <= tiny || lai2 <= tiny) ? R(0) : evap;                     // :961
  }

  // -- transpiration
  R g_surf = R(0);  // :323
  if (daytime)
    g_surf = q.csl * ramp_up<R>(d.tmin - R(273.15), q.tmin_close, q.tmin_open) *
             ramp_down<R>(vpd, q.vpd_open, q.vpd_close);  // :325-327
  g_surf = g_surf / r_corr;                               // :328
  g_pos = g_surf > R(0);
  R g_cut = q.g_cuticular / r_corr;                       // :329
  R gl_sh = q.gl_sh * d.lai * (R(1) - fw);                // :333
  R g = (gl_sh * (g_surf + g_cut)) / (gl_sh + g_surf + g_cut);       // :334-335
  R g_can = (d.lai > R(0) && (R(1) - fw) > R(0)) ? g : tiny;         // :336-337
  R r_ad = (R(1) / q.gl_sh * r_r) / (R(1) / q.gl_sh + r_r);          // :340
  if (SEM == SEM_STATIC) {
    R tr = (R(1) - fw) * ((s * a_can) + (rho * cp * d.fpar * (vpd / r_ad)));  // :344-345
    tr = tr / (s + gamma * (R(1) + (R(1) / g_can) / r_ad));                   // :346
    if (gate == GATE_CLOSED) tr = R(0);                                       // :348
    else if (gate == GATE_ELEMENT) tr = g_pos ? tr : R(0);
    o.trans = tr;
  } else {
    R a_can2 = a_can < R(0) ? R(0) : a_can;                                   // :1251
    R tr = (R(1) - fw) * ((s * a_can2) + (rho * cp * d.fpar * (vpd / r_ad))); // :1253
    tr = tr / (s + gamma * (R(1) + (R(1) / g_can) / r_ad));                   // :1255
    o.trans = g_can <= tiny ? R(0) : tr / lam;                                // :1258
  }

  // -- bare soil evaporation
  R r_tot = r_total<R>(vpd, q) / r_corr;                                      // :357-363
  R r_as = (r_tot * r_r) / (r_tot + r_r);                                     // :365
  R numer = (s * a_soil) + (rho * cp * (R(1) - d.fpar) * (vpd / r_as));       // :367-368
  R denom = s + gamma * (r_tot / r_as);                                       // :369
  R e_sat = (numer * fw) / denom;                                             // :371
  R e_unsat = (numer * (R(1) - fw)) / denom;                                  // :373
  o.sat = e_sat;
  o.unsat = e_unsat;
  R moist = m_pow<R>(rh, vpd / q.beta);                                       // :376 / :861
  if (SEM == SEM_STATIC) {
    o.soil = e_sat + e_unsat * moist;
  } else {
    R e = e_sat < R(0) ? R(0) : e_sat;                                        // :858
    e = e + (e_unsat < R(0) ? R(0) : e_unsat * moist);                        // :859-861
    o.soil = e / lam;                                                         // :864
  }
  return o;
}

// ---- one pixel-timestep: both halves -----------------------------------------------------
// Static semantics: night transpiration is identically 0 (g_surf = 0 at night, so the
// batch gate `np.any(g_surf > 0)` is never open, :323-348).
template <typename R, int SEM>
M16_HD void pixel_step(const Drivers<R>& d, const Params<R>& q, bool has_rc, int gate,
                       Half<R>& day, Half<R>& night, bool& g_pos) {
  R as_d, as_n;
  soil_radiation<R, SEM>(d, q, as_d, as_n);
  bool unused;
  day = half_day<R, SEM>(true, d.t_d, d.vpd_d, d.sw_d, d.lw_d, as_d, d.rc_d, has_rc, d, q,
                         gate, g_pos);
  night = half_day<R, SEM>(false, d.t_n, d.vpd_n, d.sw_n, d.lw_n, as_n, d.rc_n, has_rc, d,
                           q, SEM == SEM_STATIC ? (int)GATE_CLOSED : gate, unused);
}

// total of one half in the reference's association order (:380 static, :792 instance)
template <typename R, int SEM> M16_HD R half_total(const Half<R>& h) {
  return SEM == SEM_STATIC ? (h.trans + h.canopy) + h.soil : (h.canopy + h.soil) + h.trans;
}

}  // namespace mod16
