// float32 fast path of the MOD16 forward model.
//
// Same physics as mod16_math.cuh, rearranged for the B200 instruction budget (SURVEY.md
// H4; measured pipe rates in profiles/r01_microbench.txt: FP32 one warp-instruction per
// clock per SM sub-partition, packed FFMA2/FMUL2/FADD2 one per two clocks for twice the
// work, MUFU one per 8 clocks, FSETP/FSEL one per 2 clocks):
//
//   * every Penman-Monteith term is a ratio (s A + rho Cp f VPD / r_a) / (s c + gamma k / r_a).
//     Numerator and denominator are multiplied by the common positive factor
//         m = lambda(T) T (239 + Tc)^2 X          X = rho T  (air-density numerator, :409-412)
//     which cancels every division the reference performs on the way (1/lambda in gamma,
//     1/(239+Tc)^2 in the SVP slope, 1/T in rho, 1/X in r_r).  What remains per half-day is
//     one reciprocal for SVP's exponent, two exponentials (SVP and 1/SVP), lg2/ex2 for
//     (T/293.15)^-1.75, lg2/ex2 for rh^(vpd/beta) and one reciprocal per component;
//   * resistances are carried as conductances (1/r_a = 1/r_r + g), so the reference's
//     inf/NaN detours through r_h = 1/0 (f_wet == 0) never arise;
//   * the two linear ramps and the r_tot interpolation share one saturated multiply;
//   * everything that depends only on the pixel (pressure terms, the annual-temperature
//     window, BPLUT-derived slopes) is hoisted out of the time loop;
//   * the math is written once over a value type V: `float` (one pixel; strided kernel and
//     host tests) or `F2` (two neighbouring pixels in one 64-bit register pair, evaluated
//     with the sm_100 packed-FP32 instructions; the vectorised kernel).
//
// The fast formulas are used only inside the physical domain (finite drivers, 200 K < T <
// 340 K, VPD / fPAR / LAI / albedo >= 0, 30 kPa <= P <= 120 kPa, sane BPLUT column) and away
// from the model's discontinuities.  step_fast() flags an element that is outside the
// domain or within rounding of a discontinuous switch (rh = 0.7, rh = 0+, |G| = 0.39 |A|,
// A_day = 0); the kernels then re-evaluate that element with the op-for-op code of
// mod16_math.cuh in DOUBLE (pixel_step_exact), so NaN / fill masks and branch decisions
// are the float64 reference's by construction.
// Citations are file:line under /root/reference/mod16/__init__.py.
#pragma once
#include <stdint.h>
#include <string.h>
#include "mod16_math.cuh"

namespace mod16 {
namespace fast {

// ======================= value types and their operations ================================
struct alignas(8) F2 { float x, y; };  // two pixels
struct M2 { bool x, y; };              // two predicates

#if defined(__CUDA_ARCH__)
#define M16_DEV 1
#else
#define M16_DEV 0
#endif

// ---- scalar ---------------------------------------------------------------------------
#if M16_DEV
__device__ __forceinline__ float s_rcp(float x) {
  float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float s_ex2(float x) {
  float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float s_lg2(float x) {
  float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float s_sat(float x) { return __saturatef(x); }
#else
inline float s_rcp(float x) { return 1.0f / x; }
inline float s_ex2(float x) { return exp2f(x); }
inline float s_lg2(float x) { return log2f(x); }
inline float s_sat(float x) { return x != x ? 0.0f : (x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x)); }
#endif

M16_HD float vfma(float a, float b, float c) { return fmaf(a, b, c); }
M16_HD float vrcp(float a) { return s_rcp(a); }
M16_HD float vex2(float a) { return s_ex2(a); }
M16_HD float vlg2(float a) { return s_lg2(a); }
M16_HD float vsat(float a) { return s_sat(a); }
M16_HD float vabs(float a) { return fabsf(a); }
M16_HD float vmax(float a, float b) { return fmaxf(a, b); }
M16_HD float vmin(float a, float b) { return fminf(a, b); }
M16_HD bool vlt(float a, float b) { return a < b; }
M16_HD bool vle(float a, float b) { return a <= b; }
M16_HD bool vgt(float a, float b) { return a > b; }
M16_HD bool vge(float a, float b) { return a >= b; }
M16_HD bool veq(float a, float b) { return a == b; }
M16_HD bool mor(bool a, bool b) { return a || b; }
M16_HD bool mand(bool a, bool b) { return a && b; }
M16_HD bool mnot(bool a) { return !a; }
M16_HD float vsel(bool m, float a, float b) { return m ? a : b; }
M16_HD bool many(bool m) { return m; }
// the scalar type constants are written in
template <class V> struct ScalarOf { typedef float type; };
template <> struct ScalarOf<double> { typedef double type; };
template <class V> struct Splat;
template <> struct Splat<float> { static M16_HD float of(float c) { return c; } };
template <class V> M16_HD V splat(typename ScalarOf<V>::type c) { return Splat<V>::of(c); }

// ---- packed pair -----------------------------------------------------------------------
#if M16_DEV
__device__ __forceinline__ unsigned long long f2_bits(F2 a) {
  return *reinterpret_cast<unsigned long long*>(&a);
}
__device__ __forceinline__ F2 f2_from(unsigned long long b) { return *reinterpret_cast<F2*>(&b); }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return f2_from(r);
}
__device__ __forceinline__ F2 operator+(F2 a, F2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return f2_from(r);
}
__device__ __forceinline__ F2 vfma(F2 a, F2 b, F2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)),
      "l"(f2_bits(c)));
  return f2_from(r);
}
#else
inline F2 operator*(F2 a, F2 b) { return F2{a.x * b.x, a.y * b.y}; }
inline F2 operator+(F2 a, F2 b) { return F2{a.x + b.x, a.y + b.y}; }
inline F2 vfma(F2 a, F2 b, F2 c) { return F2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#endif
M16_HD F2 operator-(F2 a) { return F2{-a.x, -a.y}; }
M16_HD F2 operator-(F2 a, F2 b) { return a + (-b); }
template <> struct Splat<F2> { static M16_HD F2 of(float c) { return F2{c, c}; } };
M16_HD F2 operator*(F2 a, float b) { return a * splat<F2>(b); }
M16_HD F2 operator*(float a, F2 b) { return splat<F2>(a) * b; }
M16_HD F2 operator+(F2 a, float b) { return a + splat<F2>(b); }
M16_HD F2 operator+(float a, F2 b) { return splat<F2>(a) + b; }
M16_HD F2 operator-(F2 a, float b) { return a + splat<F2>(-b); }
M16_HD F2 operator-(float a, F2 b) { return splat<F2>(a) + (-b); }
M16_HD F2 vfma(F2 a, float b, F2 c) { return vfma(a, splat<F2>(b), c); }
M16_HD F2 vfma(float a, F2 b, F2 c) { return vfma(splat<F2>(a), b, c); }
M16_HD F2 vfma(F2 a, F2 b, float c) { return vfma(a, b, splat<F2>(c)); }
M16_HD F2 vfma(float a, F2 b, float c) { return vfma(splat<F2>(a), b, splat<F2>(c)); }
M16_HD F2 vfma(F2 a, float b, float c) { return vfma(a, splat<F2>(b), splat<F2>(c)); }
M16_HD F2 vrcp(F2 a) { return F2{s_rcp(a.x), s_rcp(a.y)}; }
M16_HD F2 vex2(F2 a) { return F2{s_ex2(a.x), s_ex2(a.y)}; }
M16_HD F2 vlg2(F2 a) { return F2{s_lg2(a.x), s_lg2(a.y)}; }
M16_HD F2 vsat(F2 a) { return F2{s_sat(a.x), s_sat(a.y)}; }
// sat(a b + c): two scalar FFMA.SAT instead of one packed FMA and two saturating adds
M16_HD float vsat_fma(float a, float b, float c) { return s_sat(fmaf(a, b, c)); }
M16_HD F2 vsat_fma(F2 a, F2 b, F2 c) {
  return F2{s_sat(fmaf(a.x, b.x, c.x)), s_sat(fmaf(a.y, b.y, c.y))};
}
M16_HD F2 vabs(F2 a) { return F2{fabsf(a.x), fabsf(a.y)}; }
M16_HD F2 vmax(F2 a, F2 b) { return F2{fmaxf(a.x, b.x), fmaxf(a.y, b.y)}; }
M16_HD F2 vmin(F2 a, F2 b) { return F2{fminf(a.x, b.x), fminf(a.y, b.y)}; }
M16_HD F2 vmax(F2 a, float b) { return F2{fmaxf(a.x, b), fmaxf(a.y, b)}; }
M16_HD M2 vlt(F2 a, F2 b) { return M2{a.x < b.x, a.y < b.y}; }
M16_HD M2 vle(F2 a, F2 b) { return M2{a.x <= b.x, a.y <= b.y}; }
M16_HD M2 vgt(F2 a, F2 b) { return M2{a.x > b.x, a.y > b.y}; }
M16_HD M2 vge(F2 a, F2 b) { return M2{a.x >= b.x, a.y >= b.y}; }
M16_HD M2 vlt(F2 a, float b) { return M2{a.x < b, a.y < b}; }
M16_HD M2 vle(F2 a, float b) { return M2{a.x <= b, a.y <= b}; }
M16_HD M2 vgt(F2 a, float b) { return M2{a.x > b, a.y > b}; }
M16_HD M2 vge(F2 a, float b) { return M2{a.x >= b, a.y >= b}; }
M16_HD M2 veq(F2 a, F2 b) { return M2{a.x == b.x, a.y == b.y}; }
M16_HD M2 mor(M2 a, M2 b) { return M2{a.x || b.x, a.y || b.y}; }
M16_HD M2 mand(M2 a, M2 b) { return M2{a.x && b.x, a.y && b.y}; }
M16_HD M2 mnot(M2 a) { return M2{!a.x, !a.y}; }
M16_HD F2 vsel(M2 m, F2 a, F2 b) { return F2{m.x ? a.x : b.x, m.y ? a.y : b.y}; }
M16_HD F2 vsel(M2 m, F2 a, float b) { return F2{m.x ? a.x : b, m.y ? a.y : b}; }
M16_HD F2 vsel(M2 m, float a, F2 b) { return F2{m.x ? a : b.x, m.y ? a : b.y}; }
M16_HD bool many(M2 m) { return m.x || m.y; }
// comparison results as 1.0f / 0.0f (device: FSET.BF, one instruction instead of a compare
// and a select); a NaN compares false
M16_HD float vstep_ge(float a, float b) { return a >= b ? 1.0f : 0.0f; }
M16_HD F2 vstep_ge(F2 a, float b) { return F2{a.x >= b ? 1.0f : 0.0f, a.y >= b ? 1.0f : 0.0f}; }
M16_HD float vstep_ge_and(float a, float b, bool m) { return (a >= b && m) ? 1.0f : 0.0f; }
M16_HD F2 vstep_ge_and(F2 a, float b, M2 m) {
  return F2{(a.x >= b && m.x) ? 1.0f : 0.0f, (a.y >= b && m.y) ? 1.0f : 0.0f};
}

// ---- double (the float64 entry: same rearranged algebra, IEEE division and libdevice
// exp2 / log2 instead of the approximations) -----------------------------------------------
M16_HD double vfma(double a, double b, double c) { return fma(a, b, c); }
#if M16_DEV
// 1 / a and 1 / sqrt(a) for normal positive-or-negative a: the 20-bit hardware seed and two
// Newton steps (error below 2 ulp; an IEEE division costs four times the instructions)
__device__ __forceinline__ double d_rcp(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ double d_rsqrt(double a) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double h = 0.5 * a, e = fma(-h * r, r, 0.5);
  r = fma(r, e, r);
  e = fma(-h * r, r, 0.5);
  return fma(r, e, r);
}
#else
inline double d_rcp(double a) { return 1.0 / a; }
inline double d_rsqrt(double a) { return 1.0 / sqrt(a); }
#endif
M16_HD double vrcp(double a) { return d_rcp(a); }

// ---- exp2 / log2 of the float64 entry, for the arguments the fast algebra produces ----------
// Inside the fast domain exp2 sees the saturation-pressure exponent (|x| < 10) and the
// exponent of rh^(vpd/beta), which is <= 0 and may be -inf (rh clamped to exactly 0, :280-281);
// log2 sees rh in [2e-5, ~3e3] or exactly 0 (0 < rh < 2e-5 is guarded, see guard_rh0).  Lanes
// outside the domain are flagged and their values discarded, so nothing else has to be right.
// Restricted to that, both are branch-free, the polynomials are evaluated pairwise (Estrin:
// a dependent chain of 5 instead of 12 -- the kernel is latency-bound) and
// the coefficients are compile-time constants.  They are 2-3 ulp functions, not libm's
// < 1 ulp: measured maximum relative error 4.8e-16 (exp2, against long double) and 5.8e-16
// (log2) over those ranges; tests/test_hostmath.py holds both below 1e-15, where the float64
// entry promises 1e-10.  -DM16_F64_LIBM keeps the library functions (A/B builds).
#if M16_DEV
__device__ __forceinline__ int d_hi(double a) { return __double2hiint(a); }
__device__ __forceinline__ int d_lo(double a) { return __double2loint(a); }
__device__ __forceinline__ double d_make(int hi, int lo) { return __hiloint2double(hi, lo); }
#else
inline int d_hi(double a) { int64_t b; memcpy(&b, &a, 8); return (int)(b >> 32); }
inline int d_lo(double a) { int64_t b; memcpy(&b, &a, 8); return (int)(b & 0xffffffffLL); }
inline double d_make(int hi, int lo) {
  const int64_t b = ((int64_t)hi << 32) | (int64_t)(uint32_t)lo;
  double a; memcpy(&a, &b, 8); return a;
}
#endif
constexpr double kLn2 = 0.6931471805599453094;
constexpr double ex2_c(int k) {        // ln2^k / k!
  double c = 1.0;
  for (int i = 1; i <= k; ++i) c = c * kLn2 / (double)i;
  return c;
}
// 2^x for x <= ~1000 (results below 2^-1021, -inf included, are returned as 0)
M16_HD double d_exp2(double x) {
  const double xc = fmax(x, -1021.0);
  const double big = 6755399441055744.0;          // 1.5 * 2^52: rint(xc) in the low word
  const double t = xc + big;
  const double f = xc - (t - big);                // exact, |f| <= 1/2
  constexpr double c0 = ex2_c(0), c1 = ex2_c(1), c2 = ex2_c(2), c3 = ex2_c(3), c4 = ex2_c(4),
                   c5 = ex2_c(5), c6 = ex2_c(6), c7 = ex2_c(7), c8 = ex2_c(8), c9 = ex2_c(9),
                   c10 = ex2_c(10), c11 = ex2_c(11), c12 = ex2_c(12);
  const double f2 = f * f, f4 = f2 * f2, f8 = f4 * f4;
  const double p01 = fma(f, c1, c0), p23 = fma(f, c3, c2), p45 = fma(f, c5, c4);
  const double p67 = fma(f, c7, c6), p89 = fma(f, c9, c8), pab = fma(f, c11, c10);
  const double q0 = fma(p23, f2, p01), q1 = fma(p67, f2, p45), q2 = fma(pab, f2, p89);
  const double r0 = fma(q1, f4, q0), r1 = fma(f4, c12, q2);
  const double p = fma(r1, f8, r0);               // 2^f in [0.707, 1.415]
  // * 2^n, n = rint(xc) in the low word of t (two's complement): normal for n >= -1021
  const double r = d_make((int)((unsigned)d_hi(p) + ((unsigned)d_lo(t) << 20)), d_lo(p));
  return x < -1021.0 ? 0.0 : r;
}
constexpr double lg2_c(int k) { return (2.0 / kLn2) / (double)(2 * k + 1); }
// log2(r) for normal positive r; log2(0) = -inf
M16_HD double d_log2(double r) {
  int hi = d_hi(r);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;            // mantissa in [1, 2)
  const bool up = hi > 0x3ff6a09e;                // > sqrt(2) (top word): halve, e + 1
  hi -= up ? 0x00100000 : 0;
  e += up ? 1 : 0;
  const double m = d_make(hi, d_lo(r));           // [0.7071, 1.4143)
  const double s = (m - 1.0) * d_rcp(m + 1.0);    // |s| < 0.1716; m - 1 is exact
  const double z = s * s, z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
  // atanh series: log2(m) = s (2 / ln2) (1 + z/3 + z^2/5 + ...), terms to z^10 (2e-17)
  constexpr double l0 = lg2_c(0), l1 = lg2_c(1), l2 = lg2_c(2), l3 = lg2_c(3), l4 = lg2_c(4),
                   l5 = lg2_c(5), l6 = lg2_c(6), l7 = lg2_c(7), l8 = lg2_c(8), l9 = lg2_c(9),
                   l10 = lg2_c(10);
  const double p01 = fma(z, l1, l0), p23 = fma(z, l3, l2), p45 = fma(z, l5, l4);
  const double p67 = fma(z, l7, l6), p89 = fma(z, l9, l8);
  const double q0 = fma(p23, z2, p01), q1 = fma(p67, z2, p45), q2 = fma(z2, l10, p89);
  const double pl = fma(q2, z8, fma(q1, z4, q0));
  const double v = fma(s, pl, (double)e);
  return r == 0.0 ? d_make((int)0xfff00000u, 0) : v;     // -inf
}
#ifdef M16_F64_LIBM
M16_HD double vex2(double a) { return exp2(a); }
M16_HD double vlg2(double a) { return log2(a); }
#else
M16_HD double vex2(double a) { return d_exp2(a); }
M16_HD double vlg2(double a) { return d_log2(a); }
#endif
M16_HD double vsat(double a) { return a != a ? 0.0 : (a < 0.0 ? 0.0 : (a > 1.0 ? 1.0 : a)); }
M16_HD double vsat_fma(double a, double b, double c) { return vsat(fma(a, b, c)); }
M16_HD double vabs(double a) { return fabs(a); }
M16_HD double vmax(double a, double b) { return fmax(a, b); }
M16_HD double vmin(double a, double b) { return fmin(a, b); }
M16_HD bool vlt(double a, double b) { return a < b; }
M16_HD bool vle(double a, double b) { return a <= b; }
M16_HD bool vgt(double a, double b) { return a > b; }
M16_HD bool vge(double a, double b) { return a >= b; }
M16_HD bool veq(double a, double b) { return a == b; }
M16_HD double vsel(bool m, double a, double b) { return m ? a : b; }
template <> struct Splat<double> { static M16_HD double of(double c) { return c; } };
M16_HD double vstep_ge(double a, double b) { return a >= b ? 1.0 : 0.0; }
M16_HD double vstep_ge_and(double a, double b, bool m) { return (a >= b && m) ? 1.0 : 0.0; }

template <class V> struct MaskOf;
template <> struct MaskOf<float> { typedef bool type; };
template <> struct MaskOf<F2> { typedef M2 type; };
template <> struct MaskOf<double> { typedef bool type; };
template <class V> M16_HD typename MaskOf<V>::type mask_splat(bool b);
template <> M16_HD bool mask_splat<float>(bool b) { return b; }
template <> M16_HD M2 mask_splat<F2>(bool b) { return M2{b, b}; }
template <> M16_HD bool mask_splat<double>(bool b) { return b; }

// normalisation of the common factor m (keeps every scaled quantity O(1..1e3))
constexpr double kLam0 = 2.45e6;          // lambda / kLam0 ~ 1
constexpr double kT0 = 293.15;            // T / kT0 = q ~ 1            (:292)
constexpr double kDD0 = 264.0 * 264.0;    // (239 + Tc)^2 / kDD0 ~ 1    (:1397)

// Every Penman-Monteith ratio below is homogeneous of degree one in (S, G, K) -- the scaled
// slope, psychrometric constant and vapour-pressure term --, so a common factor cancels.  It
// is chosen such that S = e lam q and K = lam dd VPD carry no constant of their own and
// B = 4 sigma T^4 / Cp is (q^2)^2: q = kKappa T / kT0, dd = (kR (239 + Tc))^2.
constexpr double csqrt(double x) {
  double g = x > 1.0 ? x : 1.0;
  for (int i = 0; i < 100; ++i) g = 0.5 * (g + x / g);
  return g;
}
constexpr double kCB = 4.0 * kSigma * kT0 * kT0 * kT0 * kT0 / kCp;   // B = kCB (T / kT0)^4
constexpr double kKappa = csqrt(csqrt(kCB));
constexpr double kKappa175 = kKappa * csqrt(kKappa) * csqrt(csqrt(kKappa));   // kKappa^1.75
constexpr double kC1 = 17.38 * 239.0 * 610.8;      // slope: kC1 e / (239 + Tc)^2   (:1397)
constexpr double kR = csqrt(kKappa * kCp / (kT0 * kC1));

// ======================= per-PFT and per-pixel constants ==================================
// constants derived from one BPLUT column (computed in double, stored as float)
template <class V>
struct alignas(16) PftConstV {
  // ramps (:325-326) in one-FMA form: m_tmin = sat(Tmin * inv_dtmin + tmin_bias),
  // m_vpd = sat(vpd_off - VPD * inv_dvpd)
  V tmin_bias, inv_dtmin, vpd_off, inv_dvpd;
  V gl_sh, gl_wv, g_cut, csl;
  V rbl_max, rbl_delta, inv_beta, ok;  // r_tot = rbl_max - rbl_delta * m_vpd (:357-361)
};
struct alignas(16) PftConst : PftConstV<float> {
  float tann_lo_hi, tann_lo_lo, pad0, pad1;  // 273.15 + tmin_{open|close} as a float pair
};

template <class S>
M16_HD void fill_pft_const(PftConstV<S>& c, const double v[11]) {
  c.inv_dtmin = (S)(1.0 / (v[1] - v[0]));
  c.tmin_bias = (S)(-(273.15 + v[0]) / (v[1] - v[0]));  // Tmin in kelvin goes in (:325)
  c.inv_dvpd = (S)(1.0 / (v[3] - v[2]));
  c.vpd_off = (S)(v[3] / (v[3] - v[2]));
  c.gl_sh = (S)v[4]; c.gl_wv = (S)v[5]; c.g_cut = (S)v[6]; c.csl = (S)v[7];
  c.rbl_max = (S)v[9];
  c.rbl_delta = (S)(v[9] - v[8]);
  c.inv_beta = (S)(1.0 / v[10]);
  // the rearranged algebra assumes the column is physically sane; anything else (NaN
  // columns of PFT 0 / 11, zero-width ramps, non-positive conductances) goes to the
  // op-for-op code
  bool ok = true;
  for (int k = 0; k < 11; ++k) ok = ok && (v[k] == v[k]) && (v[k] > -1e30) && (v[k] < 1e30);
  ok = ok && (v[1] - v[0] > 1e-3) && (v[3] - v[2] > 1e-3) && (v[4] > 1e-12) && (v[5] > 1e-12) &&
       (v[6] >= 0.0) && (v[7] >= 0.0) && (v[8] > 1e-6) && (v[9] > 1e-6) && (v[10] > 1e-6);
  c.ok = ok ? (S)1 : (S)0;
}

M16_HD PftConst make_pft_const(const double v[11], int sem) {
  PftConst c;
  fill_pft_const<float>(c, v);
  double lo = 273.15 + (sem == SEM_STATIC ? v[1] : v[0]);
  c.tann_lo_hi = (float)lo;
  c.tann_lo_lo = (float)(lo - (double)c.tann_lo_hi);
  c.pad0 = c.pad1 = 0.0f;
  return c;
}

// per-pixel state hoisted out of the time loop
template <class V>
struct PixConstV {
  V pa;        // 0.348444 * (P / 100)                          (:410)
  V ip;        // (P / 101300) kKappa^1.75: 1 / r_corr = ip q^-1.75     (:292)
  V gp;        // scaled psychrometric term: G = gp q dd                (:1289)
  typename MaskOf<V>::type window;  // annual-temperature window (:230-233)
  typename MaskOf<V>::type ok;      // pressure and BPLUT column inside the fast domain
};
typedef PixConstV<float> PixConst;

// `lo` = 273.15 + tmin_open (static, strict) or + tmin_close (instance), in double
template <class S>
M16_HD PixConstV<S> make_pix_const_s(S p, double t_ann, double lo, bool column_ok, int sem) {
  PixConstV<S> x;
  x.pa = (S)0.348444 * (p * (S)0.01);
  x.ip = p * (S)(kKappa175 / 101300.0);
  x.gp = p * (S)(kT0 / (kEps * kLam0 * kKappa));
  // the window is decided in double: a float32 T_ann must fall on the same side of
  // 298.15 and of 273.15 + tmin as it does in the float64 reference
  x.window = (t_ann < 273.15 + 25) && (sem == SEM_STATIC ? (t_ann > lo) : (t_ann >= lo));
  x.ok = (p >= (S)3.0e4) && (p <= (S)1.2e5) && column_ok;
  return x;
}
M16_HD PixConst make_pix_const(float p, float t_ann, const PftConst& c, int sem) {
  return make_pix_const_s<float>(p, (double)t_ann,
                                 (double)c.tann_lo_hi + (double)c.tann_lo_lo, c.ok != 0.0f, sem);
}

M16_HD PixConstV<F2> pair_pix(const PixConst& a, const PixConst& b) {
  PixConstV<F2> r;
  r.pa = F2{a.pa, b.pa}; r.ip = F2{a.ip, b.ip};
  r.gp = F2{a.gp, b.gp};
  r.window = M2{a.window, b.window}; r.ok = M2{a.ok, b.ok};
  return r;
}
M16_HD PftConstV<F2> pair_pft(const PftConstV<float>& a, const PftConstV<float>& b) {
  PftConstV<F2> r;
  r.tmin_bias = F2{a.tmin_bias, b.tmin_bias}; r.inv_dtmin = F2{a.inv_dtmin, b.inv_dtmin};
  r.vpd_off = F2{a.vpd_off, b.vpd_off};    r.inv_dvpd = F2{a.inv_dvpd, b.inv_dvpd};
  r.gl_sh = F2{a.gl_sh, b.gl_sh}; r.gl_wv = F2{a.gl_wv, b.gl_wv};
  r.g_cut = F2{a.g_cut, b.g_cut}; r.csl = F2{a.csl, b.csl};
  r.rbl_max = F2{a.rbl_max, b.rbl_max}; r.rbl_delta = F2{a.rbl_delta, b.rbl_delta};
  r.inv_beta = F2{a.inv_beta, b.inv_beta}; r.ok = F2{a.ok, b.ok};
  return r;
}

// ======================= the model ========================================================
// t - 273.15 with the constant split into a float pair, so that the Celsius temperature
// carries no 6e-6 K bias from rounding 273.15 to float
template <class V> M16_HD V celsius(V t) {
  const float hi = 273.15f;
  const float lo = (float)(273.15 - (double)hi);
  return (t - hi) - lo;
}

template <> M16_HD double celsius<double>(double t) { return t - 273.15; }   // (:216 ...)

// The hot path subtracts the float 273.15f only: its Celsius temperature is kTcBias too high,
// and every constant next to it absorbs the bias exactly (tc enters through fused
// multiply-adds and one sum only), which saves the second subtraction of celsius().
constexpr double kTcBias = 273.15 - (double)273.15f;
template <class K> M16_HD double tc_bias();
template <> M16_HD double tc_bias<float>() { return kTcBias; }
template <> M16_HD double tc_bias<double>() { return 0.0; }
template <class V> M16_HD V celsius_b(V t) { return t - 273.15f; }
template <> M16_HD double celsius_b<double>(double t) { return t - 273.15; }

// 2^-x next to e = 2^x: a second ex2 on the special-function unit (independent of the first:
// no added latency), or the reciprocal of e in double (a libdevice exp2 costs four of those)
M16_HD float vex2_neg(float x, float e) { (void)e; return vex2(-x); }
M16_HD F2 vex2_neg(F2 x, F2 e) { (void)e; return vex2(-x); }
M16_HD double vex2_neg(double x, double e) { (void)x; return d_rcp(e); }

// q^-1.75 (:292): lg2 / ex2 on the special-function unit, or two square roots in double
M16_HD float vpow_m175(float q) { return vex2(-1.75f * vlg2(q)); }
M16_HD F2 vpow_m175(F2 q) { return vex2(-1.75f * vlg2(q)); }
M16_HD double vpow_m175(double q) {
  const double y = d_rsqrt(q);          // q^-1/2
  const double y2 = y * y;
  return (y2 * y2) * d_rsqrt(y);        // q^-2 q^1/4
}

template <class V>
struct StepInV {  // one pixel-timestep (or pair) of float drivers
  V lw_d, lw_n, sw_d, sw_n, alb, t_d, t_n, tmin, vpd_d, vpd_n, fpar, lai;
  V rc_d, rc_n;
};

template <class V>
struct HalfV { V canopy, soil, trans, pt, sat, unsat; };  // (pt, sat, unsat: see Half<R>)

template <class V>
struct StepOutV {  // components and totals of one pixel-timestep (or pair)
  HalfV<V> day, night;
  V tot_day, tot_night, tot;  // :380 / :792 per half, and day + night (:193)
};

// min of three / of the magnitudes of three (device: FMNMX3)
M16_HD float vmin3(float a, float b, float c) { return fminf(fminf(a, b), c); }
M16_HD F2 vmin3(F2 a, F2 b, F2 c) {
  return F2{fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y)};
}
M16_HD float vminabs3(float a, float b, float c) {
  return fminf(fminf(fabsf(a), fabsf(b)), fabsf(c));
}
M16_HD double vmin3(double a, double b, double c) { return fmin(fmin(a, b), c); }
M16_HD double vminabs3(double a, double b, double c) {
  return fmin(fmin(fabs(a), fabs(b)), fabs(c));
}
M16_HD F2 vminabs3(F2 a, F2 b, F2 c) {
  return F2{fminf(fminf(fabsf(a.x), fabsf(b.x)), fabsf(c.x)),
            fminf(fminf(fabsf(a.y), fabsf(b.y)), fabsf(c.y))};
}

// Width of every discontinuity guard after scaling: an element is re-evaluated in double
// when one of its scaled distances to a switch is below kGuard in magnitude.
constexpr float kGuard = 1.0e-4f;
// Per scalar type: the guard width, and the centre c of the guarded interval (0, 2 c) of rh
// next to the clamp at 0, where rh = 1 - VPD / SVP has lost its leading digits to
// cancellation and rh^(VPD / beta) inherits the relative error.
template <class S> M16_HD S guard_width();
template <> M16_HD float guard_width<float>() { return kGuard; }
template <> M16_HD double guard_width<double>() { return 1.0e-9; }
// Relative width of the guard around a ramp's zero (per-element gate only, ramp_zero_need): the
// ramp's two terms are rounded coefficients times inputs, ~3 roundings of their common size
template <class K> M16_HD K ramp_guard();
template <> M16_HD float ramp_guard<float>() { return 1.0e-6f; }
template <> M16_HD double ramp_guard<double>() { return 1.0e-12; }
template <class S> M16_HD S guard_rh0();
template <> M16_HD float guard_rh0<float>() { return 1.0e-3f; }
template <> M16_HD double guard_rh0<double>() { return 1.0e-5; }

// One half-day, fast form.  STATIC night never transpires (its gate is never open).
// Returns the three components.  `guard` receives the smaller of the half's two scaled
// distances to a discontinuity in rh (see below); the caller folds all of them into ONE
// compare.  `need` is only touched for a supplied r_corr outside its sane range.
template <class V, int SEM, bool DAY, bool HAS_RC>
M16_HD HalfV<V> half_fast(V t, V tc, V vpd, V a_can, V a_soil, V rc, V tmin, V fpar, V omf,
                          V lai, V gl, V gw, const PftConstV<V>& c, const PixConstV<V>& x,
                          int gate,
                          typename MaskOf<V>::type& g_pos, typename MaskOf<V>::type& need,
                          V& gs_out, V& guard_a, V& guard_b) {
  typedef typename MaskOf<V>::type M;
  typedef typename ScalarOf<V>::type K;
  // saturation vapour pressure 610.8 exp(17.27 Tc / (Tc + 237.3)) and its reciprocal (:1365)
  // (tc carries tc_bias<K>(), see celsius_b)
  const double tb = tc_bias<K>();
  const V xe = vfma(tc, K(17.27 * 1.4426950408889634), K(-17.27 * 1.4426950408889634 * tb)) *
               vrcp(tc + K(237.3 - tb));
  const V e = vex2(xe);                             // SVP / 610.8
  const V ies = K(1.0 / 610.8) * vex2_neg(xe, e);   // 1 / SVP
  V rh;
  if (SEM == SEM_STATIC) rh = vmax(vfma(-vpd, ies, K(1.0)), splat<V>(K(0.0)));  // :280-281 (no upper clamp)
  else rh = vsat_fma(-vpd, ies, splat<V>(K(1.0)));                              // :671-673
  // discontinuities in rh: the f_wet jump at 0.7 (:282), guarded for |rh - 0.7| < 4e-6, and
  // rh -> 0+, where rh^(vpd/beta) is ill-conditioned (:376), guarded for 0 < rh < 2e-3
  // (rh == 0 exactly -- the clamp -- is NOT guarded: |0 - 1e-3| * 0.1 == kGuard).
  guard_a = vfma(rh, K(25.0), K(-17.5));       // (rh - 0.7) * 25      : 4e-6 -> kGuard
  // (rh - c) * (kGuard / c): |.| < kGuard  <=>  0 < rh < 2 c   (float: c = 1e-3)
  guard_b = vfma(rh, guard_width<K>() / guard_rh0<K>(), -guard_width<K>());
  const V rh2 = rh * rh;
  const V fw = vstep_ge(rh, K(0.7)) * (rh2 * rh2);                          // :282
  const V omfw = K(1.0) - fw;
  // scaled slope S, psychrometric constant G and vapour-pressure term K (see kKappa)
  const V q = t * K(kKappa / kT0);
  const V d2 = vfma(tc, K(kR), K((239.0 - tb) * kR));
  const V dd = d2 * d2;
  const V lam_n = vfma(tc, K(-2361.0 / kLam0), K((2.501e6 + 2361.0 * tb) / kLam0));    // :121
  const V S = (e * lam_n) * q;                                                       // :1397
  const V G = (x.gp * q) * dd;                                                       // :1289
  const V Kv = (lam_n * dd) * vpd;
  // X = rho T (:409-412); B = X / r_r = 4 sigma T^4 / Cp (:298-299)
  const V X = vfma(rh, vfma(tc, K(-0.252), K(2.0582 + 0.252 * tb)), x.pa);
  const V q2 = q * q;
  const V B = q2 * q2;
  // 1 / r_corr = (P / 101300) (T / 293.15)^-1.75 (:292), or the supplied r_corr_list entry
  V inv_rc;
  if (HAS_RC) {
    inv_rc = vrcp(rc);
    need = mor(need, mnot(mand(vgt(rc, splat<V>(K(1e-6))), vlt(rc, splat<V>(K(1e6))))));
  } else {
    inv_rc = x.ip * vpow_m175(q);
  }
  const V m_v = vsat_fma(-vpd, c.inv_dvpd, c.vpd_off);  // reversed ramp on VPD (:326)
  const V Kf = Kv * fpar;
  const V SA = S * a_can;
  V il = splat<V>(K(0.0));
  if (SEM == SEM_INSTANCE) il = K(1.0 / kLam0) * vrcp(lam_n);  // 1 / lambda (:959 ...)

  HalfV<V> o;
  // potential transpiration (:602): alpha (1 - f_wet) s A_c / (s + gamma), daytime
  o.pt = DAY ? (K(1.26) * (SA * omfw)) * vrcp(S + G) : splat<V>(K(0.0));
  // -- wet canopy (:305-320): e = fw ce X (S Ac + Kf Y) / (S ce X + G Y), Y = X / r_aw
  {
    const V ceX = (gw * fw) * X;
    const V Y = vfma(gl * fw, X, B);
    const V num = vfma(Kf, Y, SA);
    const V den = vfma(S, ceX, G * Y);
    const V r = (fw * ceX) * vrcp(den);
    if (SEM == SEM_STATIC) {
      o.canopy = num * r;  // the (lai f_wet <= 1e-7) mask changes less than 1e-6 W m-2
    } else {
      const V ev = (vmax(num, splat<V>(K(0.0))) * r) * il;                        // :959
      o.canopy = vsel(vle(lai, splat<V>(K(1e-7))), splat<V>(K(0.0)), ev);           // :934-935, :961
    }
  }
  // -- transpiration (:323-349, :1233-1258)
  g_pos = mask_splat<V>(false);
  gs_out = splat<V>(K(0.0));
  if (DAY || SEM == SEM_INSTANCE) {
    V gs = splat<V>(K(0.0));
    if (DAY) {
      const V m_t = vsat_fma(tmin, c.inv_dtmin, c.tmin_bias);                 // :325
      gs = ((c.csl * m_t) * m_v) * inv_rc;                                      // :327-328
    }
    g_pos = vgt(gs, splat<V>(K(0.0)));
    gs_out = gs;
    const V sc = vfma(c.g_cut, inv_rc, gs);                                     // :329
    const V gbl = gl * omfw;                                                    // :333
    // lai > 0 and 1 - f_wet > 0 (:336): with gl_sh > 0 (PftConst::ok) and lai >= 0 (domain)
    // that is gbl > 0 -- rh > 1 (negative VPD, static) makes 1 - f_wet and gbl negative
    const M ok = vgt(gbl, splat<V>(K(0.0)));
    const V prod = vsel(ok, gbl * sc, splat<V>(K(1e-7)));  // g_canopy = prod / tot (:334-337)
    const V tot = vsel(ok, gbl + sc, splat<V>(K(1.0)));
    const V Yd = vfma(c.gl_sh, X, B);                    // X / r_a_dry (:340)
    const V ac = (SEM == SEM_INSTANCE) ? vmax(SA, splat<V>(K(0.0))) : SA;         // :1251
    const V num = vfma(Kf, Yd, ac);
    const V den = vfma((S + G) * prod, X, (G * Yd) * tot);
    V tr = (num * ((omfw * prod) * X)) * vrcp(den);
    if (SEM == SEM_STATIC) {
      // the gate without a branch: open -> keep; closed -> drop; per element -> g_surf > 0
      const K g_el = gate == GATE_ELEMENT ? K(1.0) : K(0.0);
      const K g_op = (gate == GATE_CLOSED || gate == GATE_ELEMENT) ? K(0.0) : K(1.0);
      o.trans = vsel(vgt(vfma(gs, g_el, g_op), splat<V>(K(0.0))), tr, splat<V>(K(0.0)));
    } else {
      const M dead = mor(mnot(ok), vle(prod, K(1e-7) * tot));                     // :1258
      o.trans = vsel(dead, splat<V>(K(0.0)), tr * il);
    }
  } else {
    o.trans = splat<V>(K(0.0));
  }
  // -- bare soil (:357-376): q = X (S As r_t + Kv (1 - fpar) W) / (r_t (S X + G W)),
  //    W = X + r_t B, r_t = r_tot / r_corr
  {
    const V r_t = vfma(-c.rbl_delta, m_v, c.rbl_max) * inv_rc;                  // :357-363
    const V W = vfma(r_t, B, X);
    const V num = vfma(S * a_soil, r_t, (Kv * omf) * W);
    const V den = r_t * vfma(G, W, S * X);
    V qn = (num * X) * vrcp(den);
    const V pw = vex2((vpd * c.inv_beta) * vlg2(rh));  // rh^(vpd/beta), 1 when rh == 1
    const V mix = vfma(omfw, pw, fw);
    o.sat = qn * fw;        // :371 / :542   (W m-2, before the instance path's clamps)
    o.unsat = qn * omfw;    // :373 / :544
    if (SEM == SEM_STATIC) {
      o.soil = qn * mix;
    } else {
      qn = vmax(qn, splat<V>(K(0.0)));                                            // :858-861
      o.soil = (qn * mix) * il;                                                 // :864
    }
  }
  return o;
}

// Physical domain of the fast formulas: finite drivers; radiation terms and Tmin not a fill
// value (-9999, -32768, 1e20 ...); fPAR, LAI, albedo >= 0; VPD > -1000 Pa (a negative VPD --
// reanalysis dew point above the air temperature -- is plain arithmetic in both semantics:
// static rh > 1, :279-281; instance rh clamped to 1, :671-673); 206.3 K < T < 340 K.
// Written as sums of magnitudes (packed adds) and three min / max so that the whole test
// costs five compares per element; a NaN or an infinity anywhere fails it (every compare
// with a NaN sum is false, fmin / fmax drop a NaN operand only next to a finite one that is
// then caught by its sum).
// `nan_probe` receives a value that is NaN exactly when one of the ten per-step drivers of
// the masked-driver rule (every per-step driver except sw_rad_night and Tmin, see
// nan_halves below) is NaN: a sum of absolute values, so infinities cannot cancel into NaN.
template <class V>
M16_HD typename MaskOf<V>::type in_domain(const StepInV<V>& in, V tc_d, V tc_n, V& nan_probe) {
  typedef typename ScalarOf<V>::type K;
  const V sa = (vabs(in.lw_d) + vabs(in.lw_n)) + vabs(in.sw_d);
  const V sb = vabs(in.sw_n) + vabs(in.tmin);
  const V s2 = ((vabs(in.lai) + vabs(in.fpar)) + (vabs(in.vpd_d) + vabs(in.vpd_n))) +
               vabs(in.alb);
  nan_probe = (sa + s2) + (vabs(in.t_d) + vabs(in.t_n));
  const V lo = vmin3(in.lai, in.fpar, in.alb);
  const V vlo = vmin(in.vpd_d, in.vpd_n);
  const V tmag = vmax(vabs(tc_d), vabs(tc_n));
  // ONE magnitude test: radiation terms and Tmin weigh twelve times (|.| < 9000 each: a fill
  // value fails), everything else once (< 1.1e5); the sum sees every driver, so a NaN or an
  // infinity anywhere fails it (fmin / fmax below would drop a NaN operand)
  const V mag = vfma(sa + sb, K(11.0), nan_probe + sb);
  return mand(mand(vlt(mag, splat<V>(K(1.1e5))), vge(lo, splat<V>(K(0.0)))),
              mand(vgt(vlo, splat<V>(K(-1000.0))), vlt(tmag, splat<V>(K(66.85)))));
}
template <class V>
M16_HD typename MaskOf<V>::type in_domain(const StepInV<V>& in) {
  V probe;
  return in_domain<V>(in, celsius<V>(in.t_d), celsius<V>(in.t_n), probe);
}

// Draw-independent part of step_fast()'s guards, for batches that evaluate MANY parameter
// sets on the same element (K2): true when the element is outside the fast domain or within
// rounding of a discontinuity for ANY parameter set -- the soil-heat-flux guard is taken for
// both states of the (parameter-dependent) annual-temperature window.
M16_HD bool element_needs_exact(const StepInV<float>& in, float p) {
  // (K2's element records assume 0 <= rh <= 1: negative VPD takes its exact path)
  bool need = !((p >= 3.0e4f) && (p <= 1.2e5f) && in_domain<float>(in)) ||
              !(in.vpd_d >= 0.0f) || !(in.vpd_n >= 0.0f);
  const float a_d = fmaf(in.sw_d, 1.0f - in.alb, in.lw_d);
  const float a_n = in.lw_n;
  const float g_d = fmaf(celsius<float>(in.t_d), 4.73f, -20.87f);
  const float g_n = fmaf(celsius<float>(in.t_n), 4.73f, -20.87f);
  const float lim_d = fabsf(0.39f * a_d), lim_n = fabsf(0.39f * a_n);
  need = need || (fabsf(fabsf(g_d) - lim_d) <= 1.0e-4f) || (fabsf(fabsf(g_n) - lim_n) <= 1.0e-4f) ||
         (lim_d <= 1.0e-4f) || (lim_n <= 1.0e-4f) || (fabsf(a_d) <= 1.0e-4f);
  // (t_d - t_n) >= 5 is exact in float for inputs within a factor of two (Sterbenz)
  const float ts[2] = {in.t_d, in.t_n}, vs[2] = {in.vpd_d, in.vpd_n};
  for (int h = 0; h < 2; ++h) {
    const float tc = celsius<float>(ts[h]);
    const float xe = ((float)(17.27 * 1.4426950408889634) * tc) * s_rcp(tc + 237.3f);
    const float ies = (float)(1.0 / 610.8) * s_ex2(-xe);
    const float rh = fmaxf(fmaf(-vs[h], ies, 1.0f), 0.0f);
    need = need || (fabsf(rh - 0.7f) <= 4.0e-6f) || (fabsf(rh - 1.0e-3f) < 1.0e-3f);
  }
  return need;
}

// One pixel-timestep (or pair), both halves.  Returns the mask of elements that must be
// re-evaluated by pixel_step_exact (outside the domain, or within rounding of a
// discontinuity); their entries of `o` are unspecified.
// `gs_day` receives the day's gate quantity g_surf (> 0 <=> g_pos), so that a caller can
// track `any(g_surf > 0)` with one max instead of predicate logic.
// step_core(): everything but the domain test (the pipelined kernel tests the domain FIRST
// and deals with the elements outside it while all their inputs are still in registers);
// returns the mask of elements within rounding of a discontinuity.
template <class V, int SEM, bool HAS_RC>
M16_HD typename MaskOf<V>::type step_core(const StepInV<V>& in, V tc_d, V tc_n,
                                          const PftConstV<V>& c, const PixConstV<V>& x,
                                          int gate, StepOutV<V>& o,
                                          typename MaskOf<V>::type& g_pos, V& gs_day) {
  typedef typename MaskOf<V>::type M;
  typedef typename ScalarOf<V>::type K;
  M need = mask_splat<V>(false);
  const V oma = K(1.0) - in.alb;
  const V a_d = vfma(in.sw_d, oma, in.lw_d);            // :225
  const V a_n = in.lw_n;                                 // :226
  const V a_nc = vfma(in.sw_n, oma, in.lw_n);            // :272 (night: sw_rad_night)
  // soil heat flux (:229-241)
  const V cond = vstep_ge_and(in.t_d - in.t_n, K(5.0), x.window);  // 1 inside the window, else 0
  const double tb = tc_bias<K>();    // (tc_d, tc_n carry it, see celsius_b)
  V g_d = cond * vfma(tc_d, K(4.73), K(-20.87 - 4.73 * tb));
  V g_n = cond * vfma(tc_n, K(4.73), K(-20.87 - 4.73 * tb));
  const V lim_d = K(0.39) * a_d, lim_n = K(0.39) * a_n;
  // |G| > 0.39 |A| switches G to 0.39 A (:239); within kGuard of equality the decision is
  // the double code's.  A_day > 0 gates both limits below (:245-252).
  const V gd_d = vabs(g_d) - vabs(lim_d), gd_n = vabs(g_n) - vabs(lim_n);
  g_d = vsel(vgt(gd_d, splat<V>(K(0.0))), lim_d, g_d);
  g_n = vsel(vgt(gd_n, splat<V>(K(0.0))), lim_n, g_n);
  const M a_pos = vgt(a_d, splat<V>(K(0.0)));
  V r_d = a_d - g_d;                                     // A - G, day
  V r_n = a_n - g_n;
  r_d = vsel(mand(a_pos, vlt(r_d, splat<V>(K(0.0)))), splat<V>(K(0.0)), r_d);  // :245-247 (G := A)
  const V floor_n = K(-0.5) * a_d;
  r_n = vsel(mand(a_pos, vlt(r_n, floor_n)), floor_n, r_n);  // :248-252 (G := A_n + A_d / 2)
  const V omf = K(1.0) - in.fpar;
  const V gl = c.gl_sh * in.lai, gw = c.gl_wv * in.lai;  // shared by the two half-days
  M unused = mask_splat<V>(false);
  V gs_unused, ga_d, gb_d, ga_n, gb_n;
  o.day = half_fast<V, SEM, true, HAS_RC>(in.t_d, tc_d, in.vpd_d, in.fpar * a_d, omf * r_d,
                                          in.rc_d, in.tmin, in.fpar, omf, in.lai, gl, gw, c, x,
                                          gate, g_pos, need, gs_day, ga_d, gb_d);
  o.night = half_fast<V, SEM, false, HAS_RC>(in.t_n, tc_n, in.vpd_n, in.fpar * a_nc, omf * r_n,
                                             in.rc_n, in.tmin, in.fpar, omf, in.lai, gl, gw, c, x,
                                             gate, unused, need, gs_unused, ga_n, gb_n);
  // every discontinuity guard in one compare: the smallest scaled distance to a switch
  const V dist = vminabs3(vminabs3(gd_d, gd_n, a_d), vminabs3(ga_d, gb_d, ga_n), gb_n);
  need = mor(need, vlt(dist, splat<V>(guard_width<K>())));
  if (SEM == SEM_STATIC) {  // :380
    o.tot_day = (o.day.trans + o.day.canopy) + o.day.soil;
    o.tot_night = (o.night.trans + o.night.canopy) + o.night.soil;
  } else {                  // :792
    o.tot_day = (o.day.canopy + o.day.soil) + o.day.trans;
    o.tot_night = (o.night.canopy + o.night.soil) + o.night.trans;
  }
  o.tot = o.tot_day + o.tot_night;
  return need;
}
// Evaluated per element, the gate (:343) makes g_surf = 0+ a switch of its own: the cuticular
// transpiration appears as soon as both ramps (:325-326) are positive.  Their one-FMA forms
// are differences of numbers of the size of tmin_bias / vpd_off, so within rounding of a
// ramp's zero the sign -- and with it the element's gate -- is the double code's to decide.
// (Under a batch-global gate an element's own outputs are continuous at these points; only
// its vote in any(g_surf > 0) is not: DESIGN.md 3.2.)
// A function of the raw inputs, called BEFORE the step's arithmetic: a branch inside
// step_core splits its basic block, changes which products are fused with their sums, and
// the kernels that have the guard then differ from the one that has not by an ulp.
template <class V, int SEM>
M16_HD typename MaskOf<V>::type ramp_zero_need(const StepInV<V>& in, const PftConstV<V>& c,
                                               int gate) {
  typedef typename ScalarOf<V>::type K;
  if (SEM != SEM_STATIC || gate != GATE_ELEMENT) return mask_splat<V>(false);
  const V ut = vfma(in.tmin, c.inv_dtmin, c.tmin_bias);
  const V uv = vfma(-in.vpd_d, c.inv_dvpd, c.vpd_off);
  return mor(vlt(vabs(ut), ramp_guard<K>() * vabs(c.tmin_bias)),
             vlt(vabs(uv), ramp_guard<K>() * vabs(c.vpd_off)));
}

// (The pipelined kernel calls step_core itself and carries the ramp-zero guard only in the
// instantiations a per-element-gate launch runs on -- forward_pipe_kernel<..., EG = true> --:
// at its register limit the extra live values cost ten instructions per pair of steps on
// every path, which the headline instantiation must not pay.)
template <class V, int SEM, bool HAS_RC>
M16_HD typename MaskOf<V>::type step_fast(const StepInV<V>& in, const PftConstV<V>& c,
                                          const PixConstV<V>& x, int gate, StepOutV<V>& o,
                                          typename MaskOf<V>::type& g_pos, V& gs_day,
                                          V& nan_probe) {
  const typename MaskOf<V>::type rz = ramp_zero_need<V, SEM>(in, c, gate);
  const V tc_d = celsius_b<V>(in.t_d), tc_n = celsius_b<V>(in.t_n);
  const typename MaskOf<V>::type bad = mnot(mand(x.ok, in_domain<V>(in, tc_d, tc_n, nan_probe)));
  return mor(mor(bad, rz), step_core<V, SEM, HAS_RC>(in, tc_d, tc_n, c, x, gate, o, g_pos, gs_day));
}
template <class V, int SEM, bool HAS_RC>
M16_HD typename MaskOf<V>::type step_fast(const StepInV<V>& in, const PftConstV<V>& c,
                                          const PixConstV<V>& x, int gate, StepOutV<V>& o,
                                          typename MaskOf<V>::type& g_pos) {
  V gs_day, probe;
  return step_fast<V, SEM, HAS_RC>(in, c, x, gate, o, g_pos, gs_day, probe);
}
template <class V, int SEM, bool HAS_RC>
M16_HD typename MaskOf<V>::type step_fast(const StepInV<V>& in, const PftConstV<V>& c,
                                          const PixConstV<V>& x, int gate, StepOutV<V>& o,
                                          typename MaskOf<V>::type& g_pos, V& gs_day) {
  V probe;
  return step_fast<V, SEM, HAS_RC>(in, c, x, gate, o, g_pos, gs_day, probe);
}

// ---- scalar interface used by the strided kernel, the fix-up and the host harness -------
struct StepIn {  // one pixel-timestep of float drivers
  float lw_d, lw_n, sw_d, sw_n, alb, t_d, t_n, t_ann, tmin, vpd_d, vpd_n, p, fpar, lai;
  float rc_d, rc_n;
};

struct StepOut {
  Half<float> day, night;
  float tot_day, tot_night, tot;
};

// op-for-op evaluation in double (the fallback); `col` is the BPLUT column
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
  o.day.pt = (float)dd.pt; o.day.sat = (float)dd.sat; o.day.unsat = (float)dd.unsat;
  o.night.pt = 0.0f; o.night.sat = (float)nn.sat; o.night.unsat = (float)nn.unsat;
  // totals are formed in double: garbage-in elements (finite fills) cancel catastrophically
  const double td = half_total<double, SEM>(dd), tn = half_total<double, SEM>(nn);
  o.tot_day = (float)td; o.tot_night = (float)tn; o.tot = (float)(td + tn);
}

M16_HD StepInV<float> to_v(const StepIn& in) {
  StepInV<float> v;
  v.lw_d = in.lw_d; v.lw_n = in.lw_n; v.sw_d = in.sw_d; v.sw_n = in.sw_n; v.alb = in.alb;
  v.t_d = in.t_d; v.t_n = in.t_n; v.tmin = in.tmin; v.vpd_d = in.vpd_d; v.vpd_n = in.vpd_n;
  v.fpar = in.fpar; v.lai = in.lai; v.rc_d = in.rc_d; v.rc_n = in.rc_n;
  return v;
}

// ---- masked (NaN) drivers ------------------------------------------------------------------
// Real tiles carry masked pixels (water, fill) as NaN drivers at every step; sending each of
// them through the double-precision code just to obtain NaN costs two orders of magnitude
// (measured: 1 % masked pixels took the tile from 95 to 6 G units/s).  Reading the op-for-op
// code (mod16_math.cuh), a NaN in one of these drivers makes the half-day TOTAL NaN whatever
// the other drivers and (finite) parameters are, through a term no mask can remove:
//   static   day:   lw_net_day, sw_rad_day, sw_albedo, temp_day, vpd_day -> soil numerator
//                   (:253, :367); pressure -> gamma, rho; fpar -> soil; lai -> r_h of the
//                   wet canopy, whose mask `lai * f_wet <= tiny` is False for NaN (:305-320)
//            night: lw_net_night, temp_night, vpd_night, pressure, fpar, lai
//   instance day:   the same without lai (NaN lai zeroes canopy and transpiration, :961,
//                   :1258: finite result); night: lw_net_night, temp_night, vpd_night,
//                   pressure, fpar
// (sw_rad_night, tmin and temp_annual are not in the lists: their NaN can be masked away.)
// A NEGATIVE temperature (the -9999 / -32768 fill values of real tiles) acts like a NaN one:
// r_corr = (101300 / P) (T / 293.15)^1.75 is NaN for T < 0 (:292, :771) and poisons the soil
// term of its half-day (:363, :533) and g_surf (:328).
// nan_halves() returns bit 0 / bit 1 when the day / night total is NaN by this rule (the other
// drivers then do not matter for that half: NaN + x = NaN).  `g_pos` is the element's gate
// predicate g_surf > 0 (:343), which a NaN total does not prevent; it is formed from Tmin,
// VPD_day, T_day and pressure alone: a NaN among them makes g_surf NaN (False), otherwise
// they must be inside the fast domain for the float ramps to decide like the double code --
// if not, 0 is returned and the double code evaluates the element.
template <int SEM>
M16_HD int nan_halves(const StepIn& in, const PftConstV<float>& c, bool has_rc, bool& g_pos) {
  g_pos = false;
  if (has_rc) return 0;
  const bool n_t_d = !(in.t_d >= 0.0f), n_t_n = !(in.t_n >= 0.0f);  // NaN or negative
  const bool n_vpd_d = in.vpd_d != in.vpd_d, n_p = in.p != in.p, n_fpar = in.fpar != in.fpar;
  const bool n_lai = SEM == SEM_STATIC && in.lai != in.lai;
  const bool day = in.lw_d != in.lw_d || in.sw_d != in.sw_d || in.alb != in.alb || n_t_d ||
                   n_vpd_d || n_p || n_fpar || n_lai;
  const bool night = in.lw_n != in.lw_n || n_t_n || in.vpd_n != in.vpd_n || n_p ||
                     n_fpar || n_lai;
  if (!(day || night)) return 0;
  if (SEM == SEM_STATIC && !(n_t_d || n_p || n_vpd_d || in.tmin != in.tmin)) {
    if (c.ok == 0.0f) return 0;
    const bool sane = in.t_d > 206.3f && in.t_d < 340.0f && in.p >= 3.0e4f && in.p <= 1.2e5f &&
                      in.vpd_d > -1000.0f && in.vpd_d < 1.0e5f && fabsf(in.tmin) < 9000.0f;
    if (!sane) return 0;
    // g_surf = csl m_tmin m_vpd / r_corr, r_corr > 0 here (:325-328)
    const float m_t = vsat(fmaf(in.tmin, c.inv_dtmin, c.tmin_bias));
    const float m_v = vsat(fmaf(-in.vpd_d, c.inv_dvpd, c.vpd_off));
    g_pos = (c.csl * m_t) * m_v > 0.0f;
  }
  return (day ? 1 : 0) | (night ? 2 : 0);
}

// Fast evaluation of one pixel-timestep with the exact re-evaluation done in place.
// `guarded` (optional) reports whether the element took the double path.
template <int SEM, typename ColFn>
M16_HD void pixel_step_fast(const StepIn& in, const PftConst& c, const PixConst& x,
                            ColFn col, bool has_rc, int gate, StepOut& o, bool& g_pos,
                            bool* guarded = nullptr, int out_mode = 2) {
  StepOutV<float> r;
  const StepInV<float> v = to_v(in);
  const bool need = has_rc ? step_fast<float, SEM, true>(v, c, x, gate, r, g_pos)
                           : step_fast<float, SEM, false>(v, c, x, gate, r, g_pos);
  if (guarded) *guarded = need;
  if (need && out_mode != 2) {  // masked drivers: NaN totals without the double code
    const int h = nan_halves<SEM>(in, c, has_rc, g_pos);
    if (out_mode == 0 ? h != 0 : h == 3) {
      const float qnan = (float)NAN;
      o.day.canopy = o.day.soil = o.day.trans = qnan;
      o.night.canopy = o.night.soil = o.night.trans = qnan;
      o.tot_day = o.tot_night = o.tot = qnan;
      return;
    }
  }
  if (need) {
    double cc[11];
    for (int k = 0; k < 11; ++k) cc[k] = col(k);
    pixel_step_exact<SEM>(in, cc, has_rc, gate, o, g_pos);
    return;
  }
  o.day.canopy = r.day.canopy; o.day.soil = r.day.soil; o.day.trans = r.day.trans;
  o.night.canopy = r.night.canopy; o.night.soil = r.night.soil; o.night.trans = r.night.trans;
  o.tot_day = r.tot_day; o.tot_night = r.tot_night; o.tot = r.tot;
}

// Pair evaluation with in-place exact re-evaluation (host harness: exercises the F2 code).
template <int SEM, typename ColFn0, typename ColFn1>
M16_HD void pair_step_fast(const StepIn& in0, const StepIn& in1, const PftConst& c0,
                           const PftConst& c1, const PixConst& x0, const PixConst& x1,
                           ColFn0 col0, ColFn1 col1, int gate, StepOut& o0, StepOut& o1,
                           bool& g0, bool& g1) {
  StepInV<F2> v;
  v.lw_d = F2{in0.lw_d, in1.lw_d}; v.lw_n = F2{in0.lw_n, in1.lw_n};
  v.sw_d = F2{in0.sw_d, in1.sw_d}; v.sw_n = F2{in0.sw_n, in1.sw_n};
  v.alb = F2{in0.alb, in1.alb};    v.t_d = F2{in0.t_d, in1.t_d};
  v.t_n = F2{in0.t_n, in1.t_n};    v.tmin = F2{in0.tmin, in1.tmin};
  v.vpd_d = F2{in0.vpd_d, in1.vpd_d}; v.vpd_n = F2{in0.vpd_n, in1.vpd_n};
  v.fpar = F2{in0.fpar, in1.fpar}; v.lai = F2{in0.lai, in1.lai};
  v.rc_d = F2{0.f, 0.f}; v.rc_n = F2{0.f, 0.f};
  StepOutV<F2> r;
  M2 gp;
  const M2 need = step_fast<F2, SEM, false>(v, pair_pft(c0, c1), pair_pix(x0, x1), gate, r, gp);
  const StepIn* ins[2] = {&in0, &in1};
  StepOut* outs[2] = {&o0, &o1};
  bool* gs[2] = {&g0, &g1};
  for (int l = 0; l < 2; ++l) {
    if (l == 0 ? need.x : need.y) {
      double cc[11];
      for (int k = 0; k < 11; ++k) cc[k] = l == 0 ? col0(k) : col1(k);
      pixel_step_exact<SEM>(*ins[l], cc, false, gate, *outs[l], *gs[l]);
    } else {
#define M16_L(v) (l == 0 ? (v).x : (v).y)
      StepOut& o = *outs[l];
      o.day.canopy = M16_L(r.day.canopy); o.day.soil = M16_L(r.day.soil);
      o.day.trans = M16_L(r.day.trans);   o.night.canopy = M16_L(r.night.canopy);
      o.night.soil = M16_L(r.night.soil); o.night.trans = M16_L(r.night.trans);
      o.tot_day = M16_L(r.tot_day); o.tot_night = M16_L(r.tot_night); o.tot = M16_L(r.tot);
      *gs[l] = M16_L(gp);
#undef M16_L
    }
  }
}

}  // namespace fast
}  // namespace mod16
