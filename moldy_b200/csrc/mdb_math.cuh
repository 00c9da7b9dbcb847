// mdb_math.cuh -- FP64 building blocks of the pair kernel.
//
// The pair loop is bound by the FP64 pipe (64 DFMA/clk/SM on sm_100), so the
// transcendental pieces are written as short DFMA chains around the MUFU
// seeds instead of calling libdevice (whose sqrt / division / exp carry
// slow-path branches and IEEE rounding we do not need: 1e-13 relative is
// plenty for the 1e-10 force tolerance, we keep ~2e-16).
#pragma once
#include <cuda_runtime.h>

#ifndef MDB_LIBM_MATH
#define MDB_LIBM_MATH 0        /* 1: fall back to libdevice (debugging) */
#endif

// Polynomial / reduction constants live in __constant__ memory so that DFMA takes
// them as c[bank][offset] operands (a 64-bit immediate would cost two UMOVs per use).
__constant__ double c_exp[16] = {
   2.08767569878680989792e-09,  // 1/12!
   2.50521083854417187751e-08,  // 1/11!
   2.75573192239858906526e-07,  // 1/10!
   2.75573192239858906526e-06,  // 1/9!
   2.48015873015873015873e-05,  // 1/8!
   1.98412698412698412698e-04,  // 1/7!
   1.38888888888888888889e-03,  // 1/6!
   8.33333333333333333333e-03,  // 1/5!
   4.16666666666666666667e-02,  // 1/4!
   1.66666666666666666667e-01,  // 1/3!
   1.4426950408889634074,       // [10] log2(e)
   6755399441055744.0,          // [11] 1.5 * 2^52
   -6.93147180369123816490e-01, // [12] -ln2 (high part)
   -1.90821492927058770002e-10, // [13] -ln2 (low part)
   0.0, 0.0};
__constant__ double c_as[6] = {0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429, 0.3275911};

// 1/sqrt(x), x > 0 normal.  MUFU.RSQ64H seed + Newton: two quadratic steps (default)
// or one cubic step (MDB_RSQRT_CUBIC, 5 instead of 7 FP64 ops).
#ifndef MDB_RSQRT_CUBIC
#define MDB_RSQRT_CUBIC 0
#endif
__device__ __forceinline__ double mdb_rsqrt(double x)
{
#if MDB_LIBM_MATH
   return rsqrt(x);
#else
   double y;
   asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if MDB_RSQRT_CUBIC
   double e = fma(-x * y, y, 1.0);
   double p = fma(0.375, e, 0.5);
   return fma(y * e, p, y);
#else
   double hx = 0.5 * x;
   double e = fma(-hx, y * y, 0.5);
   y = fma(y, e, y);
   e = fma(-hx, y * y, 0.5);
   y = fma(y, e, y);
   return y;
#endif
#endif
}

// 1/x, x normal.  MUFU.RCP64H seed + two Newton steps.
__device__ __forceinline__ double mdb_rcp(double x)
{
#if MDB_LIBM_MATH
   return 1.0 / x;
#else
   double y;
   asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
   double e = fma(-x, y, 1.0);
   y = fma(y, e, y);
   e = fma(-x, y, 1.0);
   y = fma(y, e, y);
   return y;
#endif
}

// exp(x) for x <= ~700: n = rint(x log2e), r = x - n ln2 in two pieces, degree-12
// Taylor on |r| <= 0.3466 (remainder 1.7e-16), scale by adding n to the exponent
// field.  x < -708 returns 0 through an integer test on the high word (no FP64
// min/max); large positive x is the caller's responsibility (never occurs for
// -a^2 r^2, -p r).
__device__ __forceinline__ double mdb_exp(double x)
{
#if MDB_LIBM_MATH
   return exp(x);
#else
   double t = fma(x, c_exp[10], c_exp[11]);
   int n = __double2loint(t);
   double nf = t - c_exp[11];
   double r = fma(nf, c_exp[12], x);
   r = fma(nf, c_exp[13], r);
   double p = c_exp[0];
#pragma unroll
   for (int k = 1; k < 10; k++) p = fma(p, r, c_exp[k]);
   p = fma(p, r, 0.5);
   p = fma(p, r, 1.0);
   p = fma(p, r, 1.0);
   int hi = __double2hiint(p) + (n << 20), lo = __double2loint(p);
   const bool tiny = (unsigned)__double2hiint(x) > 0xc0862000u;     // x < -708
   return __hiloint2double(tiny ? 0 : hi, tiny ? 0 : lo);
#endif
}

// ---- pair potentials --------------------------------------------------------
// -phi'(r)/r and phi(r) for one site pair, the arithmetic of src/kernel.c:182-461
// (one instantiation per potential type x {Coulomb on, off}; no run-time branch
// inside).  `p` points at the 8 parameters of the (type_i,type_j) entry as
// Moldy stores them in pot_mt.p (LJ: derived values, see PairTable).  The
// Coulomb part is the reference's erfc-screened term with its own A&S 7.1.26
// polynomial (src/kernel.c:88-96,195-201) -- NOT erfc().
enum { PT_LJ = 0, PT_E6 = 1, PT_MCY = 2, PT_GEN = 3, PT_HIW = 4, PT_RSV = 5, PT_MOR = 6 };

struct PairOut { double fij, phi; };

template <int PT, bool COUL>
__device__ __forceinline__ PairOut mdb_pair_eval(double r2, double qq, const double *__restrict__ p,
                                                 double alpha, double norm)
{
   PairOut o;
   double r_r, r, r_sqr_r, erfc_term = 0.0, t = 0.0;
   if (COUL || PT != PT_LJ) {
      r_r = mdb_rsqrt(r2);
      r = r2 * r_r;
      r_sqr_r = r_r * r_r;
   } else {
      r_sqr_r = mdb_rcp(r2);
      r_r = r = 0.0;
   }
   if (COUL) {
      double ar = alpha * r;
      double tt = mdb_rcp(fma(c_as[5], ar, 1.0));
      double e = qq * mdb_exp(-(ar * ar));
      double poly = tt * fma(tt, fma(tt, fma(tt, fma(tt, c_as[4], c_as[3]), c_as[2]), c_as[1]), c_as[0]);
      t = poly * e * r_r;
      erfc_term = fma(norm, e, t);
   }
   if (PT == PT_LJ) {                       // p[0]=eps, p[1]=sigma^2, p[2]=6 eps
      double r_6_r = p[1] * r_sqr_r;
      r_6_r = r_6_r * r_6_r * r_6_r;
      double r_12_r = r_6_r * r_6_r;
      o.phi = t + p[0] * (r_12_r - r_6_r);
      o.fij = r_sqr_r * fma(p[2], fma(2.0, r_12_r, -r_6_r), erfc_term);
   } else if (PT == PT_E6) {                // -p0/r^6 + p1 exp(-p2 r)
      double exp_f1 = p[1] * mdb_exp(-p[2] * r);
      double r_6_r = p[0] * r_sqr_r * r_sqr_r * r_sqr_r;
      o.phi = t - r_6_r + exp_f1;
      o.fij = fma(r_sqr_r, fma(-6.0, r_6_r, erfc_term), p[2] * exp_f1 * r_r);
   } else if (PT == PT_MCY) {               // p0 exp(-p1 r) - p2 exp(-p3 r)
      double exp_f1 = p[0] * mdb_exp(-p[1] * r);
      double exp_f2 = -p[2] * mdb_exp(-p[3] * r);
      o.phi = t + exp_f1 + exp_f2;
      o.fij = fma(fma(p[1], exp_f1, p[3] * exp_f2), r_r, erfc_term * r_sqr_r);
   } else if (PT == PT_GEN) {               // p0 exp(-p1 r) + p2/r^12 - p3/r^4 - p4/r^6 - p5/r^8
      double exp_f1 = p[0] * mdb_exp(-p[1] * r);
      double r_4_r = r_sqr_r * r_sqr_r;
      double r_6_r = r_sqr_r * r_4_r;
      double r_8_r = p[5] * r_4_r * r_4_r;
      double r_12_r = p[2] * r_6_r * r_6_r;
      r_4_r *= p[3];
      r_6_r *= p[4];
      o.phi = t + exp_f1 + r_12_r - r_4_r - r_6_r - r_8_r;
      o.fij = fma(r_sqr_r, 12.0 * r_12_r - 4.0 * r_4_r - 6.0 * r_6_r - 8.0 * r_8_r + erfc_term,
                  p[1] * exp_f1 * r_r);
   } else if (PT == PT_HIW) {               // p0/r^4 + p1/r^6 + p2/r^12
      double r_4_r = r_sqr_r * r_sqr_r;
      double r_6_r = r_sqr_r * r_4_r;
      double r_12_r = p[2] * r_6_r * r_6_r;
      r_6_r *= p[1];
      r_4_r *= p[0];
      o.phi = t + r_4_r + r_6_r + r_12_r;
      o.fij = r_sqr_r * (4.0 * r_4_r + 6.0 * r_6_r + 12.0 * r_12_r + erfc_term);
   } else {                                 // PT_MOR: Busing-Ida-Gilbert + Morse
      double exp_f1 = p[0] * mdb_exp((p[1] - r) * p[2]);
      double r_6_r = p[3] * r_sqr_r * r_sqr_r * r_sqr_r;
      double exp_f2 = p[4] * mdb_exp(-2.0 * p[5] * (r - p[6]));
      double exp_f3 = -p[4] * 2.0 * mdb_exp(-p[5] * (r - p[6]));
      o.phi = t + exp_f1 - r_6_r + exp_f2 + exp_f3;
      o.fij = fma(r_sqr_r, fma(-6.0, r_6_r, erfc_term),
                  r_r * (p[2] * exp_f1 + (2.0 * p[5]) * exp_f2 + p[5] * exp_f3));
   }
   return o;
}
