// fastfn.cuh — FP64 exp / log / sin-cos(2 pi u) with constant-bank coefficients (fdlibm kernels, < 1 ulp), used by the RNG (rng.cuh:
// Box-Muller) and by the fused kernels (fastmath.cuh).  No dependencies; the __constant__ tables are per translation unit.
// Define MCU_FASTMATH_ESTRIN to 1 before including for Estrin-scheme polynomials (measured: no gain).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#ifndef MCU_FASTMATH_ESTRIN
#define MCU_FASTMATH_ESTRIN 0
#endif
#ifndef MCU_D
#define MCU_D __device__ __forceinline__
#endif

namespace mcu {
namespace {

struct Pair { double a, b; };

// ---- FP64 exp / log with constant-bank coefficients ------------------------------------------------------
__constant__ double kExpC[12] = {   // 1/n!, n = 13 .. 2 (Horner order); |r| <= ln2/2 ⇒ truncation < 5e-18
  1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.755731922398589e-06,
  2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03, 4.1666666666666664e-02,
  1.6666666666666666e-01, 0.5};
__constant__ double kLogC[7] = {   // fdlibm e_log.c Lg7 .. Lg1
  1.479819860511658591e-01, 1.531383769920937332e-01, 1.818357216161805012e-01, 2.222219843214978396e-01,
  2.857142874366239149e-01, 3.999999999940941908e-01, 6.666666666666735130e-01};

// exp(x) for |x| < 700 (callers clamp): x = k ln2 + r, exp(r) by a degree-13 Taylor polynomial, 2^k through the exponent field
MCU_D double fast_exp(double x) {
  const double kf = rint(x * 1.4426950408889634074);
  double r = fma(kf, -6.93147180369123816490e-01, x);
  r = fma(kf, -1.90821492927058770002e-10, r);
#if MCU_FASTMATH_ESTRIN
  // Estrin's scheme: dependency depth 6 instead of 12 (kExpC[11 - i] is the coefficient of r^i)
  const double r2 = r * r, r4 = r2 * r2;
  const double b0 = fma(kExpC[10], r, kExpC[11]), b1 = fma(kExpC[8], r, kExpC[9]), b2 = fma(kExpC[6], r, kExpC[7]);
  const double b3 = fma(kExpC[4], r, kExpC[5]), b4 = fma(kExpC[2], r, kExpC[3]), b5 = fma(kExpC[0], r, kExpC[1]);
  const double c0 = fma(b1, r2, b0), c1 = fma(b3, r2, b2), c2 = fma(b5, r2, b4);
  double p = fma(fma(c2, r4, c1), r4, c0);
  p = fma(p, r2, r) + 1.0;                              // 1 + r + r^2 (1/2 + r (1/6 + ...))
#else
  double p = kExpC[0];
#pragma unroll
  for (int i = 1; i < 12; ++i) p = fma(p, r, kExpC[i]);
  p = fma(p * r, r, r) + 1.0;                           // 1 + r + r^2 (1/2 + r (1/6 + ...))
#endif
  const int k = (int)kf;
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
// log(x) for normal positive x (fdlibm e_log.c): x = 2^k m, m in [sqrt(1/2), sqrt(2)), f = m - 1, s = f / (2 + f)
MCU_D double fast_log(double x) {
  int hx = __double2hiint(x);
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int adj = (hx + 0x95f64) & 0x100000;            // mantissa above sqrt(2): halve m, k += 1
  k += adj >> 20;
  const double m = __hiloint2double(hx | (adj ^ 0x3ff00000), __double2loint(x));
  const double f = m - 1.0;
  const double dnm = 2.0 + f;
  double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(dnm));
  y = fma(fma(-dnm, y, 1.0), y, y);
  y = fma(fma(-dnm, y, 1.0), y, y);                     // 1 / (2 + f) to full precision
  const double sq = f * y;
  const double z = sq * sq;
#if MCU_FASTMATH_ESTRIN
  // fdlibm's own even/odd split (kLogC[7 - i] = Lg_i): two chains of depth 3-4 instead of one of depth 7
  const double w = z * z;
  const double t1 = w * fma(w, fma(w, kLogC[1], kLogC[3]), kLogC[5]);                      // w (Lg2 + w (Lg4 + w Lg6))
  const double t2 = z * fma(w, fma(w, fma(w, kLogC[0], kLogC[2]), kLogC[4]), kLogC[6]);    // z (Lg1 + w (Lg3 + w (Lg5 + w Lg7)))
  const double R = t2 + t1;
#else
  double R = kLogC[0];
#pragma unroll
  for (int i = 1; i < 7; ++i) R = fma(R, z, kLogC[i]);
  R *= z;
#endif
  const double hfsq = 0.5 * f * f;
  const double dk = (double)k;
  // log(1+f) = f - (hfsq - s (hfsq + R));  result = k ln2_hi - ((hfsq - (s (hfsq + R) + k ln2_lo)) - f)
  return fma(dk, 6.93147180369123816490e-01, -((hfsq - fma(sq, hfsq + R, dk * 1.90821492927058770002e-10)) - f));
}

// sqrt(x) for x >= 0 below ~1e300, branch-free (the library sqrt carries a slow-path call that ends the basic block, so the scheduler cannot
// interleave two Box-Muller transforms): y ~ 1/sqrt(x) from MUFU.RSQ64H (~2^-22), two Newton steps, then one correction of s = x y with
// the exact residual x - s^2 (FMA).  <= 1 ulp (tools/check_fasttab.cpp tests the same arithmetic on the host).  x == 0 returns 0.
MCU_D double fast_sqrt(double x) {
  double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = fma(y, fma(-hx * y, y, 0.5), y);                   // y (1 + (1/2 - x y^2 / 2))
  y = fma(y, fma(-hx * y, y, 0.5), y);
  double sq = x * y;
  sq = fma(fma(-sq, sq, x), 0.5 * y, sq);                // s + (x - s^2) / (2 s)
  return x > 0.0 ? sq : 0.0;
}

// 1 / sqrt(x) and 1 / x for normal positive x, branch-free, <= 2 ulp: MUFU seed + two Newton steps (the library forms are ~35 / ~25
// instructions with a slow-path call; these are 8 / 5).
MCU_D double fast_rsqrt(double x) {
  double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = fma(y, fma(-hx * y, y, 0.5), y);
  y = fma(y, fma(-hx * y, y, 0.5), y);
  return y;
}
MCU_D double fast_rcp(double x) {
  double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  y = fma(fma(-x, y, 1.0), y, y);
  y = fma(fma(-x, y, 1.0), y, y);
  return y;
}

// sin and cos of 2 pi u, u in [0, 1): quadrant reduction is exact (u - q/4), then fdlibm's sin/cos kernels on |x| <= pi/4
__constant__ double kSinC[6] = {1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
                                -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01};
__constant__ double kCosC[6] = {-1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
                                2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};
MCU_D Pair fast_sincos2pi(double u) {   // returns (sin, cos) of 2 pi u
  const double qf = rint(4.0 * u);
  const int q = (int)qf & 3;
  const double x = 6.283185307179586476925286766559 * fma(qf, -0.25, u);   // |x| <= pi/4
  const double z = x * x;
  double sp = kSinC[0], cp = kCosC[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) { sp = fma(sp, z, kSinC[i]); cp = fma(cp, z, kCosC[i]); }
  const double sn = fma(x * z, sp, x);
  const double cs = fma(z * z, cp, fma(z, -0.5, 1.0));
  // angle x + q pi/2:  cos: q = 0 → cos, 1 → -sin, 2 → -cos, 3 → sin ;  sin: q = 0 → sin, 1 → cos, 2 → -sin, 3 → -cos
  const double cv = (q & 1) ? sn : cs, sv = (q & 1) ? cs : sn;
  return {(q & 2) ? -sv : sv, ((q + 1) & 2) ? -cv : cv};
}

}  // namespace
}  // namespace mcu
