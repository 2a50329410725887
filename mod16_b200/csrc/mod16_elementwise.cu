// K4 -- the component methods and physics helpers of the reference's public API, one
// element per thread.  These are the pieces the fused forward kernel is made of
// (/root/reference/mod16/__init__.py: latent_heat_vaporization :121, MOD16.air_density :385,
// air_pressure :415, potential_soil_evaporation :450, potential_transpiration :547, vpd :605,
// rhumidity :647, evaporation_soil :795, evaporation_wet_canopy :866, radiation_soil :963,
// soil_heat_flux :1055, surface_conductance :1121, transpiration :1152,
// psychrometric_constant :1261, radiation_net :1293, svp :1340, svp_slope :1370), exposed one
// by one because the reference's callers and its own unit tests (tests/tests.py:92-232) call
// them directly.  Inputs are the positional arguments of the reference signature, each
// broadcastable; optional arguments (lhv, rhumidity, f_wet, r_corr) are marked in `present`.
//
// All arithmetic is in double, expression by expression as the reference writes it, for
// both entries: the float32 entry converts its inputs on load and rounds the result on
// store (these calls are a few values to a few million; nothing here is throughput-bound).
#include "mod16_common.cuh"

namespace mod16 {

struct ElemArgs {
  Arr<float> in32[MOD16_MAX_OP_INPUTS];
  Arr<double> in64[MOD16_MAX_OP_INPUTS];
  uint32_t present;
  int op;
  double par[MOD16_N_PARAMS];
  double opt[2];
  void* out[2];
  int64_t T, N;
};

template <typename R>
__device__ __forceinline__ double load_in(const ElemArgs& a, int k, int64_t t, int64_t n) {
  if (sizeof(R) == 4) return (double)a.in32[k].at(t, n);
  return a.in64[k].at(t, n);
}

// optional (lhv, rhumidity, f_wet) of the component methods, defaults as :855-858, :926-931
struct Moist {
  double lam, rh, fw;
};
template <typename R>
__device__ __forceinline__ Moist moisture(const ElemArgs& a, int k_lam, int k_rh, int k_fw,
                                          double temp_k, double vpd, int64_t t, int64_t n) {
  Moist m;
  m.lam = (k_lam >= 0 && (a.present >> k_lam & 1)) ? load_in<R>(a, k_lam, t, n)
                                                   : latent_heat<double>(temp_k);
  m.rh = (a.present >> k_rh & 1) ? load_in<R>(a, k_rh, t, n)
                                 : rhumidity_clamped<double>(temp_k, vpd);
  m.fw = (a.present >> k_fw & 1) ? load_in<R>(a, k_fw, t, n) : f_wet_of<double>(m.rh);
  return m;
}

// :504-544 -- returns (evap_sat, evap_unsat) in W m-2
__device__ __forceinline__ void soil_potential(double p, double temp_k, double vpd, double fpar,
                                               double rad_soil, double r_corr, const Moist& m,
                                               const Params<double>& q, double& e_sat,
                                               double& e_unsat) {
  const double s = svp_slope<double>(temp_k, svp<double>(temp_k));
  const double rho = air_density<double>(temp_k, p, m.rh);
  const double gamma = psychrometric<double>(p, temp_k);
  const double r_r = r_radiative<double>(rho, temp_k);
  double r_tot = r_total<double>(vpd, q);
  r_tot = r_tot / r_corr;
  const double r_as = (r_tot * r_r) / (r_tot + r_r);
  const double numer = (s * rad_soil) + (rho * kCp * (1.0 - fpar) * (vpd / r_as));
  const double denom = s + gamma * (r_tot / r_as);
  e_sat = (numer * m.fw) / denom;
  e_unsat = (numer * (1.0 - m.fw)) / denom;
}

template <typename R>
__global__ void __launch_bounds__(256) elementwise_kernel(const __grid_constant__ ElemArgs a) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.T * a.N) return;
  const int64_t t = e / a.N, n = e - t * a.N;
#define IN(k) load_in<R>(a, (k), t, n)
  Params<double> q;
  q.tmin_close = a.par[0]; q.tmin_open = a.par[1]; q.vpd_open = a.par[2];
  q.vpd_close = a.par[3];  q.gl_sh = a.par[4];     q.gl_wv = a.par[5];
  q.g_cuticular = a.par[6]; q.csl = a.par[7];      q.rbl_min = a.par[8];
  q.rbl_max = a.par[9];    q.beta = a.par[10];
  q.tann_lo = 273.15 + a.par[0];
  double o0 = 0.0, o1 = 0.0;
  switch (a.op) {
    case MOD16_OP_SVP: o0 = svp<double>(IN(0)); break;
    case MOD16_OP_SVP_SLOPE: {
      const double tk = IN(0);
      o0 = svp_slope<double>(tk, (a.present >> 1 & 1) ? IN(1) : svp<double>(tk));
      break;
    }
    case MOD16_OP_LHV: o0 = latent_heat<double>(IN(0)); break;
    case MOD16_OP_PSYCHROMETRIC: o0 = psychrometric<double>(IN(0), IN(1)); break;
    case MOD16_OP_AIR_DENSITY: o0 = air_density<double>(IN(0), IN(1), IN(2)); break;
    case MOD16_OP_RHUMIDITY: o0 = rhumidity_clamped<double>(IN(0), IN(1)); break;
    case MOD16_OP_VPD: o0 = vpd_from_qv<double>(IN(0), IN(1), IN(2)); break;
    case MOD16_OP_AIR_PRESSURE: o0 = air_pressure<double>(IN(0)); break;
    case MOD16_OP_RADIATION_NET: {  // :1333-1337 (deprecated in the reference, still tested)
      const double sw = IN(0), alb = IN(1), tk = IN(2);
      const double tc = tk - 273.15;
      const double emis_atmos = 1.0 - 0.26 * exp(-7.77e-4 * (tc * tc));
      o0 = sw * (1.0 - alb) + kSigma * (emis_atmos - 0.97) * ((tk * tk) * (tk * tk));
      break;
    }
    case MOD16_OP_SOIL_HEAT_FLUX:     // :1103-1119
    case MOD16_OP_RADIATION_SOIL: {   // :1033-1053
      double rad_d, rad_n, t_d, t_n, t_ann, fpar = 0.0;
      if (a.op == MOD16_OP_SOIL_HEAT_FLUX) {
        rad_d = IN(0); rad_n = IN(1); t_d = IN(2); t_n = IN(3); t_ann = IN(4);
      } else {
        rad_d = IN(2) * (1.0 - IN(4)) + IN(0);
        rad_n = IN(1);
        t_d = IN(5); t_n = IN(6); t_ann = IN(7); fpar = IN(8);
      }
      const bool cond = (t_ann < 273.15 + 25) && (t_ann >= 273.15 + q.tmin_close) &&
                        ((t_d - t_n) >= 5.0);
      double g_d = cond ? (4.73 * (t_d - 273.15)) - 20.87 : 0.0;
      double g_n = cond ? (4.73 * (t_n - 273.15)) - 20.87 : 0.0;
      g_d = fabs(g_d) > (0.39 * fabs(rad_d)) ? 0.39 * rad_d : g_d;
      g_n = fabs(g_n) > (0.39 * fabs(rad_n)) ? 0.39 * rad_n : g_n;
      if (a.op == MOD16_OP_SOIL_HEAT_FLUX) {
        o0 = g_d; o1 = g_n;
      } else {
        g_d = ((rad_d - g_d < 0.0) && (rad_d > 0.0)) ? rad_d : g_d;
        g_n = ((rad_d > 0.0) && ((rad_n - g_n) < (-0.5 * rad_d))) ? rad_n + (0.5 * rad_d) : g_n;
        o0 = (1.0 - fpar) * (rad_d - g_d);
        o1 = (1.0 - fpar) * (rad_n - g_n);
      }
      break;
    }
    case MOD16_OP_POTENTIAL_SOIL_EVAPORATION:   // :450-544
    case MOD16_OP_EVAPORATION_SOIL: {           // :795-864
      const double p = IN(0), tk = IN(1), vpd = IN(2), fpar = IN(3), rad_soil = IN(4);
      const Moist m = moisture<R>(a, 6, 7, 8, tk, vpd, t, n);
      const double r_corr = (a.present >> 5 & 1) ? IN(5) : r_correction<double>(p, tk);
      double e_sat, e_unsat;
      soil_potential(p, tk, vpd, fpar, rad_soil, r_corr, m, q, e_sat, e_unsat);
      if (a.op == MOD16_OP_POTENTIAL_SOIL_EVAPORATION) {
        o0 = e_sat; o1 = e_unsat;
      } else {
        double ev = e_sat < 0.0 ? 0.0 : e_sat;
        ev = ev + (e_unsat < 0.0 ? 0.0 : e_unsat * pow(m.rh, vpd / q.beta));
        o0 = ev / m.lam;
      }
      break;
    }
    case MOD16_OP_EVAPORATION_WET_CANOPY: {     // :866-961
      const double p = IN(0), tk = IN(1), vpd = IN(2), fpar = IN(4), rad_c = IN(5);
      double lai = IN(3);
      const double tiny = a.opt[1];
      const Moist m = moisture<R>(a, 6, 7, 8, tk, vpd, t, n);
      const double fw = m.fw == 0.0 ? m.fw + tiny : m.fw;   // :934
      lai = lai == 0.0 ? lai + tiny : lai;                   // :935
      const double s = svp_slope<double>(tk, svp<double>(tk));
      const double rho = air_density<double>(tk, p, m.rh);
      const double r_h = 1.0 / (q.gl_sh * lai * fw);
      const double r_e = 1.0 / (q.gl_wv * lai * fw);
      const double r_r = r_radiative<double>(rho, tk);
      const double r_aw = (r_h * r_r) / (r_h + r_r);
      const double numer = fw * ((s * rad_c) + ((rho * kCp * fpar * vpd * 1.0) / r_aw));
      const double denom = s + (((p * kCp * r_e) * 1.0) / (m.lam * kEps * r_aw));
      const double evap = numer < 0.0 ? 0.0 : (numer / denom) / m.lam;   // :959
      o0 = (fw <= tiny || lai <= tiny) ? 0.0 : evap;                     // :961
      break;
    }
    case MOD16_OP_SURFACE_CONDUCTANCE:          // :1147-1150
      o0 = q.csl * ramp_up<double>(IN(0) - 273.15, q.tmin_close, q.tmin_open) *
           ramp_down<double>(IN(1), q.vpd_open, q.vpd_close);
      break;
    case MOD16_OP_TRANSPIRATION: {              // :1152-1258
      const double p = IN(0), tk = IN(1), vpd = IN(2), lai = IN(3), fpar = IN(4);
      double rad_c = IN(5);
      const double tmin = IN(6);
      const bool daytime = a.opt[0] != 0.0;
      const double tiny = a.opt[1];
      const Moist m = moisture<R>(a, 8, 9, 10, tk, vpd, t, n);
      const double r_corr = (a.present >> 7 & 1) ? IN(7) : r_correction<double>(p, tk);
      const double s = svp_slope<double>(tk, svp<double>(tk));
      const double rho = air_density<double>(tk, p, m.rh);
      const double gamma = psychrometric<double>(p, tk);
      const double r_r = r_radiative<double>(rho, tk);
      double g_surf = 0.0;
      if (daytime)
        g_surf = (q.csl * ramp_up<double>(tmin - 273.15, q.tmin_close, q.tmin_open) *
                  ramp_down<double>(vpd, q.vpd_open, q.vpd_close)) / r_corr;
      const double g_cut = q.g_cuticular / r_corr;
      const double gl_sh = q.gl_sh * lai * (1.0 - m.fw);
      const double g = (gl_sh * (g_surf + g_cut)) / (gl_sh + g_surf + g_cut);
      const double g_can = (lai > 0.0 && (1.0 - m.fw) > 0.0) ? g : tiny;
      const double r_ad = (1.0 / q.gl_sh * r_r) / (1.0 / q.gl_sh + r_r);
      rad_c = rad_c < 0.0 ? 0.0 : rad_c;                                  // :1251
      double tr = (1.0 - m.fw) * ((s * rad_c) + (rho * kCp * fpar * (vpd / r_ad)));
      tr = tr / (s + gamma * (1.0 + (1.0 / g_can) / r_ad));
      o0 = g_can <= tiny ? 0.0 : tr / m.lam;                              // :1258
      break;
    }
    case MOD16_OP_POTENTIAL_TRANSPIRATION: {    // :547-602
      const double lw = IN(0), sw = IN(1), alb = IN(2), p = IN(3), tk = IN(4), vpd = IN(5),
                   fpar = IN(6);
      const Moist m = moisture<R>(a, -1, 7, 8, tk, vpd, t, n);
      const double rad_net = sw * (1.0 - alb) + lw;
      const double rad_c = fpar * rad_net;
      const double s = svp_slope<double>(tk, svp<double>(tk));
      const double gamma = psychrometric<double>(p, tk);
      o0 = (a.opt[0] * (s * rad_c) * (1.0 - m.fw)) / (s + gamma);
      break;
    }
    case MOD16_OP_DAYLENGTH_COMPOSITE: {        // tests/verification/verify2.py:112-115
      const double day = IN(0), night = IN(1), dl = IN(2), nd = a.opt[0];
      o0 = (day * dl * nd * 60.0 * 60.0) + (night * (24.0 - dl) * nd * 60.0 * 60.0);
      break;
    }
  }
#undef IN
  if (a.out[0]) ((R*)a.out[0])[e] = (R)o0;
  if (a.out[1]) ((R*)a.out[1])[e] = (R)o1;
}

static const int kOpInputs[MOD16_OP_COUNT][2] = {
    // required, total (required + optional)
    {1, 1}, {1, 2}, {1, 1}, {2, 2}, {3, 3}, {2, 2}, {3, 3}, {1, 1}, {3, 3},
    {5, 5}, {9, 9}, {5, 9}, {5, 9}, {6, 9}, {2, 2}, {7, 11}, {7, 9}, {3, 3}};

template <typename R>
static int run_elementwise(const mod16_elementwise_t* ew, int64_t T, int64_t N,
                           cudaStream_t st) {
  if (!ew) return MOD16_ERR_NULL;
  if (T < 0 || N < 0) return MOD16_ERR_SHAPE;
  if (ew->op < 0 || ew->op >= MOD16_OP_COUNT) return MOD16_ERR_FLAGS;
  const int need = kOpInputs[ew->op][0], total = kOpInputs[ew->op][1];
  if ((ew->present & ((1u << need) - 1u)) != ((1u << need) - 1u)) return MOD16_ERR_NULL;
  if (ew->present >> total) return MOD16_ERR_FLAGS;
  if (T * N == 0) return MOD16_OK;
  if (!ew->out[0]) return MOD16_ERR_NULL;
  ElemArgs a;
  for (int k = 0; k < MOD16_MAX_OP_INPUTS; ++k) {
    mod16_driver_t d = ew->in[k];
    if (!(ew->present >> k & 1)) { d.ptr = nullptr; d.t_stride = d.n_stride = 0; d.scalar = 0.0; }
    if (d.t_stride < 0 || d.n_stride < 0) return MOD16_ERR_SHAPE;
    fill_arr<float>(a.in32[k], d);
    fill_arr<double>(a.in64[k], d);
  }
  a.present = ew->present;
  a.op = ew->op;
  for (int k = 0; k < MOD16_N_PARAMS; ++k) a.par[k] = ew->params[k];
  a.opt[0] = ew->opt[0]; a.opt[1] = ew->opt[1];
  a.out[0] = ew->out[0]; a.out[1] = ew->out[1];
  a.T = T; a.N = N;
  const int64_t blocks = (T * N + 255) / 256;
  if (blocks > 2147483647LL) return MOD16_ERR_SHAPE;
  elementwise_kernel<R><<<(unsigned)blocks, 256, 0, st>>>(a);
  M16_CUDA_CHECK(cudaGetLastError());
  return MOD16_OK;
}

}  // namespace mod16

extern "C" {
int mod16_elementwise_f32(const mod16_elementwise_t* ew, int64_t T, int64_t N, void* stream) {
  return mod16::run_elementwise<float>(ew, T, N, (cudaStream_t)stream);
}
int mod16_elementwise_f64(const mod16_elementwise_t* ew, int64_t T, int64_t N, void* stream) {
  return mod16::run_elementwise<double>(ew, T, N, (cudaStream_t)stream);
}
}
