// TEST HARNESS ONLY -- compiles the kernels' per-element math (mod16_b200/csrc/*.cuh) for
// the host with g++, so that the math can be checked against the oracle in a container
// that has no GPU.  Nothing in the product loads this library.
#include <stdint.h>
#include "../../mod16_b200/csrc/mod16_math.cuh"
#ifdef HM_WITH_FAST
#include "../../mod16_b200/csrc/mod16_fast.cuh"
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

extern "C" {
void hm_forward_f64(int sem, int gate, const double* drv, const double* par, const double* rc,
                    int has_rc, int64_t n, double* out, uint8_t* gpos) {
  run<double>(sem, gate, drv, par, rc, has_rc, n, out, gpos);
}
void hm_forward_f32(int sem, int gate, const float* drv, const float* par, const float* rc,
                    int has_rc, int64_t n, float* out, uint8_t* gpos) {
  run<float>(sem, gate, drv, par, rc, has_rc, n, out, gpos);
}
}
