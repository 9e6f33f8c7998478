// FP64 elementary functions sized for the covariance build: the FP64 pipe is the
// bound of the small-object kernels, and libdevice's exp() costs ~49 issued
// instructions per call (ncu, round 1) against 15 FP64-pipe ops here.
#pragma once
#include <cuda_runtime.h>

namespace cgp {

// Literal doubles cost two 32-bit immediate moves per use; from constant memory they are a
// direct DFMA operand (c[bank][off]).
static __constant__ double kExp[16] = {
    1.4426950408889634, 6755399441055744.0, -6.93147180559945286e-01, -2.31904681384629956e-17,
    2.5110049204818658e-08, 2.763265472252779e-07, 2.755724088722987e-06, 2.4801485441561313e-05,
    0.00019841269890076403, 0.0013888888952352863, 0.008333333333319589, 0.04166666666648795,
    0.1666666666666668, 0.5000000000000019, 1.0, 0.0};

// exp(x) for x <= ~700 (the RBF exponent is <= 0).  Range reduction x = k ln2 + r with
// the 1.5*2^52 rounding trick, r by two FMAs against a hi/lo split of ln2, degree-11
// polynomial from Chebyshev-node interpolation on |r| <= ln2/2 (tools/fit_exp_poly.py:
// truncation 4e-18, evaluated error <= 1 ulp), scaling by an integer add on the
// exponent field.  x < -700 returns 0 (true value < 1e-304); NaN propagates.
__device__ __forceinline__ double cgp_exp(double x) {
  double t = fma(x, kExp[0], kExp[1]);
  const int k = __double2loint(t);
  t -= kExp[1];
  double r = fma(t, kExp[2], x);
  r = fma(t, kExp[3], r);
  double p = fma(kExp[4], r, kExp[5]);
  p = fma(p, r, kExp[6]);
  p = fma(p, r, kExp[7]);
  p = fma(p, r, kExp[8]);
  p = fma(p, r, kExp[9]);
  p = fma(p, r, kExp[10]);
  p = fma(p, r, kExp[11]);
  p = fma(p, r, kExp[12]);
  p = fma(p, r, kExp[13]);
  p = fma(p, r, kExp[14]);
  p = fma(p, r, kExp[14]);
  const int hi = __double2hiint(p) + (k << 20);
  const double res = __hiloint2double(hi, __double2loint(p));
  // -inf <= x < -700  <=>  high word in (0xC085E000, 0xFFF00000]; tested on the integer pipe
  return ((unsigned)__double2hiint(x) - 0xC085E001u <= 0xFFF00000u - 0xC085E001u) ? 0.0 : res;
}

// 1/sqrt(d), d > 0 normal: MUFU.RSQ64H seed (~2^-22) + one cubically convergent step,
// y1 = y + (y e)(1/2 + 3/8 e), e = 1 - d y^2, arranged 4 dependent FP64 ops deep (it sits on the
// serial pivot chain of the diagonal-tile factorisation: 64 times per 60-point object).
__device__ __forceinline__ double cgp_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double t = d * y;
  const double e = fma(-t, y, 1.0);
  const double ye = y * e;
  const double p = fma(0.375, e, 0.5);
  return fma(ye, p, y);
}

}  // namespace cgp
