// TEST HARNESS ONLY -- compiles the kernels' per-element math (mod16_b200/csrc/*.cuh) for
// the host with g++, so that the math can be checked against the oracle in a container
// that has no GPU.  Nothing in the product loads this library.
#include <stdint.h>
#include <vector>
#include "../../mod16_b200/csrc/mod16_math.cuh"
#ifdef HM_WITH_FAST
#include "../../mod16_b200/csrc/mod16_fast.cuh"
#include "../../mod16_b200/csrc/mod16_objective.cuh"
#endif

using namespace mod16;

template <typename R>
static void run(int sem, int gate, const R* drv, const R* par, const R* rc, int has_rc,
                int64_t n, R* out, uint8_t* gpos) {
  for (int64_t i = 0; i < n; ++i) {
    Drivers<R> d;
    const R* c = drv;
    d.lw_d = c[0 * n + i]; d.lw_n = c[1 * n + i]; d.sw_d = c[2 * n + i];
    d.sw_n = c[3 * n + i]; d.alb = c[4 * n + i];  d.t_d = c[5 * n + i];
    d.t_n = c[6 * n + i];  d.t_ann = c[7 * n + i]; d.tmin = c[8 * n + i];
    d.vpd_d = c[9 * n + i]; d.vpd_n = c[10 * n + i]; d.p = c[11 * n + i];
    d.fpar = c[12 * n + i]; d.lai = c[13 * n + i];
    d.rc_d = has_rc ? rc[i] : R(0);
    d.rc_n = has_rc ? rc[n + i] : R(0);
    Params<R> q;
    q.tmin_close = par[0 * n + i]; q.tmin_open = par[1 * n + i];
    q.vpd_open = par[2 * n + i];   q.vpd_close = par[3 * n + i];
    q.gl_sh = par[4 * n + i];      q.gl_wv = par[5 * n + i];
    q.g_cuticular = par[6 * n + i]; q.csl = par[7 * n + i];
    q.rbl_min = par[8 * n + i];    q.rbl_max = par[9 * n + i];
    q.beta = par[10 * n + i];
    q.tann_lo = 273.15 + (double)(sem == SEM_STATIC ? q.tmin_open : q.tmin_close);
    Half<R> day, night;
    bool g;
    if (sem == SEM_STATIC) pixel_step<R, SEM_STATIC>(d, q, has_rc != 0, gate, day, night, g);
    else pixel_step<R, SEM_INSTANCE>(d, q, has_rc != 0, gate, day, night, g);
    out[0 * n + i] = day.canopy;   out[1 * n + i] = day.soil;   out[2 * n + i] = day.trans;
    out[3 * n + i] = night.canopy; out[4 * n + i] = night.soil; out[5 * n + i] = night.trans;
    gpos[i] = g ? 1 : 0;
  }
}

#ifdef HM_WITH_FAST
static fast::StepIn load_in(const float* c, const float* rc, int has_rc, int64_t n, int64_t i) {
  fast::StepIn in;
  in.lw_d = c[0 * n + i]; in.lw_n = c[1 * n + i]; in.sw_d = c[2 * n + i];
  in.sw_n = c[3 * n + i]; in.alb = c[4 * n + i];  in.t_d = c[5 * n + i];
  in.t_n = c[6 * n + i];  in.t_ann = c[7 * n + i]; in.tmin = c[8 * n + i];
  in.vpd_d = c[9 * n + i]; in.vpd_n = c[10 * n + i]; in.p = c[11 * n + i];
  in.fpar = c[12 * n + i]; in.lai = c[13 * n + i];
  in.rc_d = has_rc ? rc[i] : 0.f;
  in.rc_n = has_rc ? rc[n + i] : 0.f;
  return in;
}

static void store_out(const fast::StepOut& r, bool g, bool guarded, int64_t n, int64_t i,
                      float* out, uint8_t* gpos) {
  out[0 * n + i] = r.day.canopy;   out[1 * n + i] = r.day.soil;
  out[2 * n + i] = r.day.trans;    out[3 * n + i] = r.night.canopy;
  out[4 * n + i] = r.night.soil;   out[5 * n + i] = r.night.trans;
  out[6 * n + i] = r.tot_day;      out[7 * n + i] = r.tot_night;
  out[8 * n + i] = r.tot;
  gpos[i] = (g ? 1 : 0) | (guarded ? 2 : 0);
}

// pair != 0: elements are evaluated two at a time through the packed (F2) instantiation
static void run_fast(int sem, int gate, const float* drv, const double* par, const float* rc,
                     int has_rc, int64_t n, float* out, uint8_t* gpos, int pair) {
  for (int64_t i = 0; i < n; i += (pair && i + 1 < n && !has_rc) ? 2 : 1) {
    const bool two = pair && i + 1 < n && !has_rc;
    fast::StepIn in[2];
    double col[2][11];
    fast::PftConst pc[2];
    fast::PixConst px[2];
    for (int l = 0; l < (two ? 2 : 1); ++l) {
      in[l] = load_in(drv, rc, has_rc, n, i + l);
      for (int k = 0; k < 11; ++k) col[l][k] = par[k * n + i + l];
      pc[l] = fast::make_pft_const(col[l], sem);
      px[l] = fast::make_pix_const(in[l].p, in[l].t_ann, pc[l], sem);
    }
    fast::StepOut r[2];
    bool g[2] = {false, false}, guarded[2] = {false, false};
    auto c0 = [&](int k) { return col[0][k]; };
    auto c1 = [&](int k) { return col[1][k]; };
    if (two) {
      if (sem == SEM_STATIC)
        fast::pair_step_fast<SEM_STATIC>(in[0], in[1], pc[0], pc[1], px[0], px[1], c0, c1, gate,
                                         r[0], r[1], g[0], g[1]);
      else
        fast::pair_step_fast<SEM_INSTANCE>(in[0], in[1], pc[0], pc[1], px[0], px[1], c0, c1,
                                           gate, r[0], r[1], g[0], g[1]);
    } else {
      if (sem == SEM_STATIC)
        fast::pixel_step_fast<SEM_STATIC>(in[0], pc[0], px[0], c0, has_rc != 0, gate, r[0], g[0],
                                          &guarded[0]);
      else
        fast::pixel_step_fast<SEM_INSTANCE>(in[0], pc[0], px[0], c0, has_rc != 0, gate, r[0],
                                            g[0], &guarded[0]);
    }
    for (int l = 0; l < (two ? 2 : 1); ++l) store_out(r[l], g[l], guarded[l], n, i + l, out, gpos);
  }
}
#endif

#ifdef HM_WITH_FAST
// K2 math: per (draw, element) prediction with the gate open, the day transpiration and the
// gate quantity; cls[i] = 1 when the element has a fast record, 2 when it needs the exact path
static void run_objective(const float* drv, const double* draws, int64_t D, int64_t n, int pair,
                          float* tot, float* tr, float* gs, uint8_t* cls, uint8_t* draw_ok) {
  std::vector<obj::ElemRec> recs(n);
  for (int64_t i = 0; i < n; ++i) {
    fast::StepIn in = load_in(drv, nullptr, 0, n, i);
    cls[i] = obj::make_elem_rec(in, 0.0f, 1.0f, recs[i]) ? 1 : 2;
  }
  std::vector<obj::DrawConstV<float>> dc(D);
  for (int64_t d = 0; d < D; ++d) draw_ok[d] = obj::make_draw_const(draws + 11 * d, dc[d]) ? 1 : 0;
  for (int64_t d = 0; d < D; d += pair ? 2 : 1) {
    const int64_t d1 = (pair && d + 1 < D) ? d + 1 : d;
    const obj::DrawConstV<fast::F2> pc = obj::pair_draws(dc[d], dc[d1]);
    for (int64_t i = 0; i < n; ++i) {
      if (cls[i] != 1) continue;
      if (pair) {
        fast::F2 t2, g2;
        const fast::F2 r = obj::objective_step<fast::F2>(recs[i], pc, t2, g2);
        tot[d * n + i] = r.x; tr[d * n + i] = t2.x; gs[d * n + i] = g2.x;
        tot[d1 * n + i] = r.y; tr[d1 * n + i] = t2.y; gs[d1 * n + i] = g2.y;
      } else {
        float t1, g1;
        tot[d * n + i] = obj::objective_step<float>(recs[i], dc[d], t1, g1);
        tr[d * n + i] = t1; gs[d * n + i] = g1;
      }
    }
  }
}
#endif

#ifdef HM_WITH_FAST
// the masked-driver rule: per element, bit 0 / 1 = day / night total NaN by the rule (0: not
// decided), and the gate predicate the rule reports
static void run_nan_halves(int sem, const float* drv, const double* par, int64_t n, uint8_t* h,
                           uint8_t* gpos) {
  for (int64_t i = 0; i < n; ++i) {
    fast::StepIn in = load_in(drv, nullptr, 0, n, i);
    double col[11];
    for (int k = 0; k < 11; ++k) col[k] = par[k * n + i];
    const fast::PftConst pc = fast::make_pft_const(col, sem);
    bool g = false;
    h[i] = (uint8_t)(sem == SEM_STATIC ? fast::nan_halves<SEM_STATIC>(in, pc, false, g)
                                       : fast::nan_halves<SEM_INSTANCE>(in, pc, false, g));
    gpos[i] = g ? 1 : 0;
  }
}
#endif

#ifdef HM_WITH_FAST
// The float64 entry's fast path (forward_fast64_kernel): the rearranged algebra in double,
// the op-for-op code for the elements it flags.  out rows as in run_fast.
static void run_fast64(int sem, int gate, const double* drv, const double* par, const double* rc,
                       int has_rc, int64_t n, double* out, uint8_t* gpos) {
  for (int64_t i = 0; i < n; ++i) {
    Drivers<double> d;
    const double* c = drv;
    d.lw_d = c[0 * n + i]; d.lw_n = c[1 * n + i]; d.sw_d = c[2 * n + i];
    d.sw_n = c[3 * n + i]; d.alb = c[4 * n + i];  d.t_d = c[5 * n + i];
    d.t_n = c[6 * n + i];  d.t_ann = c[7 * n + i]; d.tmin = c[8 * n + i];
    d.vpd_d = c[9 * n + i]; d.vpd_n = c[10 * n + i]; d.p = c[11 * n + i];
    d.fpar = c[12 * n + i]; d.lai = c[13 * n + i];
    d.rc_d = has_rc ? rc[i] : 0.0;
    d.rc_n = has_rc ? rc[n + i] : 0.0;
    double col[11];
    for (int k = 0; k < 11; ++k) col[k] = par[k * n + i];
    Params<double> q;
    q.tmin_close = col[0]; q.tmin_open = col[1]; q.vpd_open = col[2]; q.vpd_close = col[3];
    q.gl_sh = col[4]; q.gl_wv = col[5]; q.g_cuticular = col[6]; q.csl = col[7];
    q.rbl_min = col[8]; q.rbl_max = col[9]; q.beta = col[10];
    q.tann_lo = 273.15 + (sem == SEM_STATIC ? col[1] : col[0]);
    fast::PftConstV<double> pc;
    fast::fill_pft_const<double>(pc, col);
    const fast::PixConstV<double> px =
        fast::make_pix_const_s<double>(d.p, d.t_ann, q.tann_lo, pc.ok != 0.0, sem);
    fast::StepInV<double> in;
    in.lw_d = d.lw_d; in.lw_n = d.lw_n; in.sw_d = d.sw_d; in.sw_n = d.sw_n; in.alb = d.alb;
    in.t_d = d.t_d; in.t_n = d.t_n; in.tmin = d.tmin; in.vpd_d = d.vpd_d; in.vpd_n = d.vpd_n;
    in.fpar = d.fpar; in.lai = d.lai; in.rc_d = d.rc_d; in.rc_n = d.rc_n;
    fast::StepOutV<double> r;
    bool g = false, need;
    if (sem == SEM_STATIC)
      need = has_rc ? fast::step_fast<double, SEM_STATIC, true>(in, pc, px, gate, r, g)
                    : fast::step_fast<double, SEM_STATIC, false>(in, pc, px, gate, r, g);
    else
      need = has_rc ? fast::step_fast<double, SEM_INSTANCE, true>(in, pc, px, gate, r, g)
                    : fast::step_fast<double, SEM_INSTANCE, false>(in, pc, px, gate, r, g);
    if (need) {
      Half<double> day, night;
      if (sem == SEM_STATIC) pixel_step<double, SEM_STATIC>(d, q, has_rc != 0, gate, day, night, g);
      else pixel_step<double, SEM_INSTANCE>(d, q, has_rc != 0, gate, day, night, g);
      r.day.canopy = day.canopy; r.day.soil = day.soil; r.day.trans = day.trans;
      r.night.canopy = night.canopy; r.night.soil = night.soil; r.night.trans = night.trans;
      if (sem == SEM_STATIC) {
        r.tot_day = half_total<double, SEM_STATIC>(day);
        r.tot_night = half_total<double, SEM_STATIC>(night);
      } else {
        r.tot_day = half_total<double, SEM_INSTANCE>(day);
        r.tot_night = half_total<double, SEM_INSTANCE>(night);
      }
      r.tot = r.tot_day + r.tot_night;
    }
    out[0 * n + i] = r.day.canopy;   out[1 * n + i] = r.day.soil;
    out[2 * n + i] = r.day.trans;    out[3 * n + i] = r.night.canopy;
    out[4 * n + i] = r.night.soil;   out[5 * n + i] = r.night.trans;
    out[6 * n + i] = r.tot_day;      out[7 * n + i] = r.tot_night;
    out[8 * n + i] = r.tot;
    gpos[i] = (g ? 1 : 0) | (need ? 2 : 0);
  }
}
#endif

extern "C" {
#ifdef HM_WITH_FAST
void hm_fast_f64(int sem, int gate, const double* drv, const double* par, const double* rc,
                 int has_rc, int64_t n, double* out, uint8_t* gpos) {
  run_fast64(sem, gate, drv, par, rc, has_rc, n, out, gpos);
}
// the float64 entry's own exp2 / log2 (mod16_fast.cuh: d_exp2, d_log2), element by element
void hm_exp2_log2(const double* x, int64_t n, double* ex, double* lg) {
  for (int64_t i = 0; i < n; ++i) {
    ex[i] = mod16::fast::d_exp2(x[i]);
    lg[i] = mod16::fast::d_log2(x[i]);
  }
}
void hm_nan_halves(int sem, const float* drv, const double* par, int64_t n, uint8_t* h,
                   uint8_t* gpos) {
  run_nan_halves(sem, drv, par, n, h, gpos);
}
void hm_objective_f32(const float* drv, const double* draws, int64_t D, int64_t n, int pair,
                      float* tot, float* tr, float* gs, uint8_t* cls, uint8_t* draw_ok) {
  run_objective(drv, draws, D, n, pair, tot, tr, gs, cls, draw_ok);
}
void hm_fast_f32(int sem, int gate, const float* drv, const double* par, const float* rc,
                 int has_rc, int64_t n, float* out, uint8_t* gpos) {
  run_fast(sem, gate, drv, par, rc, has_rc, n, out, gpos, 0);
}
void hm_fastpair_f32(int sem, int gate, const float* drv, const double* par, const float* rc,
                     int has_rc, int64_t n, float* out, uint8_t* gpos) {
  run_fast(sem, gate, drv, par, rc, has_rc, n, out, gpos, 1);
}
#endif
void hm_forward_f64(int sem, int gate, const double* drv, const double* par, const double* rc,
                    int has_rc, int64_t n, double* out, uint8_t* gpos) {
  run<double>(sem, gate, drv, par, rc, has_rc, n, out, gpos);
}
void hm_forward_f32(int sem, int gate, const float* drv, const float* par, const float* rc,
                    int has_rc, int64_t n, float* out, uint8_t* gpos) {
  run<float>(sem, gate, drv, par, rc, has_rc, n, out, gpos);
}
}
