// K2 math: many parameter draws evaluated on the same elements (pixel-timesteps).
//
// The objective loops of the reference call `MOD16._et(params_vector, *drivers)` once per
// parameter draw on the SAME drivers (mod16/sensitivity.py:94-96, mod16/calibration.py:
// 224-227 through mod17's BlackBoxLikelihood), so nearly all of the model is
// draw-independent: saturation vapour pressure, slope, psychrometric constant, air density,
// r_corr, net radiation, soil heat flux, f_wet and every product of them.  make_elem_rec()
// evaluates that part ONCE per element with the formulas of mod16_fast.cuh (half_fast) and
// stores the coefficients; objective_step() is what remains per (draw, element): the two
// ramps, the conductance algebra, three to five reciprocals and two exponentials.
//
// objective_step() is written over a value type V: `float`, or `F2` = two DRAWS in one
// packed-FP32 register pair.  The element coefficients are then scalars shared by both
// lanes (SASS: FFMA2 with an .F32 broadcast operand), and every property of the element
// (f_wet == 0: no wet canopy) is a warp-uniform branch.
//
// Static semantics only (`_et`): that is what both callers pass.  Elements for which the
// rearranged float algebra is not trusted for EVERY draw (fast::element_needs_exact) are not
// given a record; the kernels evaluate them with the
// op-for-op double code for every draw.
// Citations are file:line under /root/reference/mod16/__init__.py.
#pragma once
#include "mod16_fast.cuh"

namespace mod16 {
namespace obj {

using fast::F2;
using fast::M2;

// ---- per-element record (40 floats = 10 quads, read with uniform 128-bit loads) ---------
// The element's coefficients are pre-multiplied as far as the draw-dependent algebra allows:
//   transpiration   num = gl_sh KfX + KfBSA,  G Yd = gl_sh GX + GB          (Yd = gl_sh X + B)
//   bare soil       q = (r P1X + P2X) / (r (r Q1 + Q2)),  r = r_tot / r_corr
//                   P1 = S As + Ko B, P2 = Ko X, Q1 = G B, Q2 = (S + G) X    (W = r B + X)
//   wet canopy      num = gl_sh Kfu + KfBSA,  G Y = gl_sh Gu + GB            (Y = gl_sh u + B)
// P1X0 / P1X1: with the soil heat flux G = 0 / G computed; which one applies depends on the
// draw (annual-temperature window, :230-233).
struct alignas(16) ElemRec {
  float obs_w, w, tann_eff, tminc;          // obs * w, w, window temperature, Tmin - 273.15
  float d_vpd, d_inv_rc, d_lo, d_KfX;       // day, always needed
  float d_KfBSA, d_GX, d_GB, d_SGX;
  float d_oX, d_P1X0, d_P1X1, d_P2X;
  float d_Q1, d_vl, d_omfw, d_fw;           // d_fw doubles as the wet-canopy branch condition
  float d_Kfu, d_Gu, d_Su, d_fwu;           // day wet canopy (f_wet > 0 only)
  float n_vpd, n_inv_rc, n_P1X0, n_P1X1;    // night, always needed
  float n_P2X, n_Q1, n_Q2, n_vl;
  float n_omfw, n_fw, n_KfBSA, n_GB;        // (the last two: night wet canopy)
  float n_Kfu, n_Gu, n_Su, n_fwu;           // night wet canopy
};
static_assert(sizeof(ElemRec) == 160, "ElemRec layout");

// Element coefficients of one half-day; same expressions as fast::half_fast.
struct HalfCoef {
  float S, G, Kv, X, B, inv_rc, fw, omfw, vl, rh;
};

M16_HD HalfCoef half_coef(float t, float vpd, float p) {
  using namespace fast;
  HalfCoef h;
  const float tc = celsius<float>(t);
  const float xe = ((float)(17.27 * 1.4426950408889634) * tc) * s_rcp(tc + 237.3f);
  const float e = s_ex2(xe);
  const float ies = (float)(1.0 / 610.8) * s_ex2(-xe);
  const float rh = fmaxf(fmaf(-vpd, ies, 1.0f), 0.0f);            // :280-281
  const float rh2 = rh * rh;
  h.rh = rh;
  h.fw = rh >= 0.7f ? rh2 * rh2 : 0.0f;                            // :282
  h.omfw = 1.0f - h.fw;
  const float q = t * (float)(1.0 / kT0);
  const float d2 = tc + 239.0f;
  const float dd = d2 * d2;
  const float lam_n = fmaf(tc, (float)(-2361.0 / kLam0), (float)(2.501e6 / kLam0));
  h.S = (((float)(17.38 * 239.0 * 610.8 / kDD0) * e) * lam_n) * q;
  const float cpp_n = p * (float)(kCp / kEps / (kLam0 * kDD0));
  h.G = (cpp_n * q) * dd;
  h.Kv = (((float)(kCp / (kT0 * kDD0)) * lam_n) * dd) * vpd;
  const float pa = 0.348444f * (p * 0.01f);
  h.X = fmaf(rh, fmaf(tc, -0.252f, 2.0582f), pa);
  const float q2 = q * q;
  h.B = (float)(4.0 * kSigma * kT0 * kT0 * kT0 * kT0 / kCp) * (q2 * q2);
  h.inv_rc = (p * (float)(1.0 / 101300.0)) * s_ex2(-1.75f * s_lg2(q));
  h.vl = vpd * s_lg2(rh);  // rh^(vpd / beta) = 2^(vl / beta)
  return h;
}

// Builds the record of one element.  Returns false when the element must take the exact
// path for every draw (the record is then not used).
M16_HD bool make_elem_rec(const fast::StepIn& in, float obs, float w, ElemRec& r) {
  using namespace fast;
  const StepInV<float> v = to_v(in);
  if (element_needs_exact(v, in.p)) return false;
  const HalfCoef d = half_coef(in.t_d, in.vpd_d, in.p);
  const HalfCoef n = half_coef(in.t_n, in.vpd_n, in.p);
  const float oma = 1.0f - in.alb;
  const float a_d = fmaf(in.sw_d, oma, in.lw_d);   // :225
  const float a_n = in.lw_n;                        // :226
  const float a_nc = fmaf(in.sw_n, oma, in.lw_n);   // :272 (night)
  const float omf = 1.0f - in.fpar;
  // radiation reaching the soil for both states of the annual-temperature window
  float as_d[2], as_n[2];
  const bool dt_ok = (in.t_d - in.t_n) >= 5.0f;     // :234
  const bool below = (double)in.t_ann < 273.15 + 25;  // :231
  for (int k = 0; k < 2; ++k) {
    const bool cond = k == 1 && dt_ok && below;
    float g_d = cond ? fmaf(celsius<float>(in.t_d), 4.73f, -20.87f) : 0.0f;   // :237
    float g_n = cond ? fmaf(celsius<float>(in.t_n), 4.73f, -20.87f) : 0.0f;
    const float lim_d = 0.39f * a_d, lim_n = 0.39f * a_n;
    g_d = fabsf(g_d) > fabsf(lim_d) ? lim_d : g_d;                             // :239
    g_n = fabsf(g_n) > fabsf(lim_n) ? lim_n : g_n;
    float r_d = a_d - g_d, r_n = a_n - g_n;
    const bool a_pos = a_d > 0.0f;
    r_d = (a_pos && r_d < 0.0f) ? 0.0f : r_d;                                  // :245-247
    const float floor_n = -0.5f * a_d;
    r_n = (a_pos && r_n < floor_n) ? floor_n : r_n;                            // :248-252
    as_d[k] = omf * r_d;
    as_n[k] = omf * r_n;
  }
  r.obs_w = obs * w; r.w = w;
  // the window's draw-dependent side is t_ann > 273.15 + tmin_open (:233): elements whose
  // other two conditions fail can never satisfy it
  r.tann_eff = (dt_ok && below) ? in.t_ann : -1.0e30f;
  r.tminc = celsius<float>(in.tmin);
  const float fpar = in.fpar;
  // day
  const float d_Kf = d.Kv * fpar, d_Ko = d.Kv * omf, d_SA = d.S * (fpar * a_d);
  r.d_vpd = in.vpd_d; r.d_inv_rc = d.inv_rc; r.d_lo = in.lai * d.omfw;
  r.d_KfX = d_Kf * d.X; r.d_KfBSA = fmaf(d_Kf, d.B, d_SA);
  r.d_GX = d.G * d.X; r.d_GB = d.G * d.B; r.d_SGX = (d.S + d.G) * d.X;
  r.d_oX = d.omfw * d.X;
  r.d_P1X0 = fmaf(d_Ko, d.B, d.S * as_d[0]) * d.X;
  r.d_P1X1 = fmaf(d_Ko, d.B, d.S * as_d[1]) * d.X;
  r.d_P2X = (d_Ko * d.X) * d.X;
  r.d_Q1 = r.d_GB; r.d_vl = d.vl; r.d_omfw = d.omfw; r.d_fw = d.fw;
  const float d_u = (in.lai * d.fw) * d.X;
  r.d_Kfu = d_Kf * d_u; r.d_Gu = d.G * d_u; r.d_Su = d.S * d_u; r.d_fwu = d.fw * d_u;
  // night
  const float n_Kf = n.Kv * fpar, n_Ko = n.Kv * omf, n_SA = n.S * (fpar * a_nc);
  r.n_vpd = in.vpd_n; r.n_inv_rc = n.inv_rc;
  r.n_P1X0 = fmaf(n_Ko, n.B, n.S * as_n[0]) * n.X;
  r.n_P1X1 = fmaf(n_Ko, n.B, n.S * as_n[1]) * n.X;
  r.n_P2X = (n_Ko * n.X) * n.X;
  r.n_Q1 = n.G * n.B; r.n_Q2 = (n.S + n.G) * n.X; r.n_vl = n.vl;
  r.n_omfw = n.omfw; r.n_fw = n.fw;
  r.n_KfBSA = fmaf(n_Kf, n.B, n_SA); r.n_GB = r.n_Q1;
  const float n_u = (in.lai * n.fw) * n.X;
  r.n_Kfu = n_Kf * n_u; r.n_Gu = n.G * n_u; r.n_Su = n.S * n_u; r.n_fwu = n.fw * n_u;
  return true;
}

// ---- per-draw constants -------------------------------------------------------------------
template <class V>
struct DrawConstV {
  V tmin_close, inv_dtmin, vpd_close, inv_dvpd;  // ramps (:325-326)
  V gl_sh, gl_wv, g_cut, csl;
  V rbl_max, rbl_delta, inv_beta;
  V tann_lo;  // largest float <= 273.15 + tmin_open: x > lo in double <=> x > tann_lo in float
};

// Largest float that is <= x.
M16_HD float float_round_down(double x) {
  float f = (float)x;
  if ((double)f > x) f = nextafterf(f, -3.0e38f);
  return f;
}

// Returns false when the draw is outside the domain of the rearranged algebra (the kernels
// then evaluate the whole draw with the op-for-op double code): same test as make_pft_const.
M16_HD bool make_draw_const(const double v[11], DrawConstV<float>& c) {
  const fast::PftConst p = fast::make_pft_const(v, SEM_STATIC);
  c.tmin_close = (float)v[0]; c.inv_dtmin = p.inv_dtmin;
  c.vpd_close = (float)v[3];  c.inv_dvpd = p.inv_dvpd;
  c.gl_sh = p.gl_sh; c.gl_wv = p.gl_wv; c.g_cut = p.g_cut; c.csl = p.csl;
  c.rbl_max = p.rbl_max; c.rbl_delta = p.rbl_delta; c.inv_beta = p.inv_beta;
  c.tann_lo = float_round_down(273.15 + v[1]);
  return p.ok != 0.0f;
}

M16_HD DrawConstV<F2> pair_draws(const DrawConstV<float>& a, const DrawConstV<float>& b) {
  DrawConstV<F2> r;
  r.tmin_close = F2{a.tmin_close, b.tmin_close}; r.inv_dtmin = F2{a.inv_dtmin, b.inv_dtmin};
  r.vpd_close = F2{a.vpd_close, b.vpd_close};    r.inv_dvpd = F2{a.inv_dvpd, b.inv_dvpd};
  r.gl_sh = F2{a.gl_sh, b.gl_sh}; r.gl_wv = F2{a.gl_wv, b.gl_wv};
  r.g_cut = F2{a.g_cut, b.g_cut}; r.csl = F2{a.csl, b.csl};
  r.rbl_max = F2{a.rbl_max, b.rbl_max}; r.rbl_delta = F2{a.rbl_delta, b.rbl_delta};
  r.inv_beta = F2{a.inv_beta, b.inv_beta}; r.tann_lo = F2{a.tann_lo, b.tann_lo};
  return r;
}

// scalar-on-the-left forms used below (the element coefficient is the scalar)
M16_HD float bsel(bool m, float a, float b) { return m ? a : b; }
M16_HD F2 bsel(M2 m, float a, float b) { return F2{m.x ? a : b, m.y ? a : b}; }
M16_HD bool bgt(float a, float b) { return a > b; }
M16_HD M2 bgt(float a, F2 b) { return M2{a > b.x, a > b.y}; }

// One element for one draw (or draw pair).  Returns day + night with the transpiration
// gate OPEN; `tr` is the day transpiration (the gate-closed total is the result minus tr)
// and `gs` the gate quantity csl m_tmin m_vpd (> 0 <=> g_surf > 0, :343).
// (Measured and rejected: sharing one reciprocal between the day's transpiration and soil
// terms, a / b + c / d = (a d + c b) / (b d), is slower -- 413 vs 461 G draw-elements/s --
// the FMA pipe, not the MUFU pipe, is what the loop saturates.)
template <class V>
M16_HD V objective_step(const ElemRec& e, const DrawConstV<V>& c, V& tr, V& gs) {
  using namespace fast;
  typedef typename MaskOf<V>::type M;
  const M win = bgt(e.tann_eff, c.tann_lo);                                  // :233
  // ---- day
  const V m_t = vsat((e.tminc - c.tmin_close) * c.inv_dtmin);                // :325
  const V m_vd = vsat((c.vpd_close - e.d_vpd) * c.inv_dvpd);                 // :326
  gs = (c.csl * m_t) * m_vd;                                                 // :327
  V tot;
  {
    const V sc = (c.g_cut + gs) * e.d_inv_rc;                                // :328-329
    const V gbl = c.gl_sh * e.d_lo;                                          // :333
    // lai == 0 or f_wet == 1: g_canopy is the constant 1e-7 (:336-337); a property of the
    // element, hence a uniform select
    const bool lo_pos = e.d_lo > 0.0f;
    const V prod = lo_pos ? gbl * sc : splat<V>(1.0e-7f);                    // :334-335
    const V sum = lo_pos ? gbl + sc : splat<V>(1.0f);
    const V num = vfma(c.gl_sh, e.d_KfX, e.d_KfBSA);                         // :344-345
    const V gyd = vfma(c.gl_sh, e.d_GX, e.d_GB);
    const V den = vfma(e.d_SGX, prod, gyd * sum);                            // :346
    const V nt = num * (e.d_oX * prod);
    // bare soil (:357-376)
    const V r_t = vfma(-c.rbl_delta, m_vd, c.rbl_max) * e.d_inv_rc;
    const V nx = vfma(r_t, bsel(win, e.d_P1X1, e.d_P1X0), e.d_P2X);
    const V dens = r_t * vfma(r_t, e.d_Q1, e.d_SGX);
    const V pw = vex2(c.inv_beta * e.d_vl);
    const V mix = vfma(e.d_omfw, pw, e.d_fw);
    tr = nt * vrcp(den);
    tot = vfma(nx * vrcp(dens), mix, tr);
  }
  if (e.d_fw > 0.0f) {  // wet canopy (:305-320); uniform branch
    const V num = vfma(c.gl_sh, e.d_Kfu, e.d_KfBSA);
    const V den = vfma(c.gl_wv, e.d_Su, vfma(c.gl_sh, e.d_Gu, e.d_GB));
    tot = vfma(num * (c.gl_wv * e.d_fwu), vrcp(den), tot);
  }
  // ---- night (static: no transpiration, :323-348)
  {
    const V m_vn = vsat((c.vpd_close - e.n_vpd) * c.inv_dvpd);
    const V r_t = vfma(-c.rbl_delta, m_vn, c.rbl_max) * e.n_inv_rc;
    const V nx = vfma(r_t, bsel(win, e.n_P1X1, e.n_P1X0), e.n_P2X);
    const V dens = r_t * vfma(r_t, e.n_Q1, e.n_Q2);
    const V pw = vex2(c.inv_beta * e.n_vl);
    const V mix = vfma(e.n_omfw, pw, e.n_fw);
    tot = vfma(nx * vrcp(dens), mix, tot);
  }
  if (e.n_fw > 0.0f) {
    const V num = vfma(c.gl_sh, e.n_Kfu, e.n_KfBSA);
    const V den = vfma(c.gl_wv, e.n_Su, vfma(c.gl_sh, e.n_Gu, e.n_GB));
    tot = vfma(num * (c.gl_wv * e.n_fwu), vrcp(den), tot);
  }
  return tot;
}

}  // namespace obj
}  // namespace mod16
