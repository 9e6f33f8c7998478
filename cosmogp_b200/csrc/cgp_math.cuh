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
#ifdef CGP_EXP_FAKE_SHORT      /* timing experiment only: the cost of a degree-5 polynomial (what a table-driven exp would need) */
  double p = kExp[10];
#else
  double p = fma(kExp[4], r, kExp[5]);
  p = fma(p, r, kExp[6]);
  p = fma(p, r, kExp[7]);
  p = fma(p, r, kExp[8]);
  p = fma(p, r, kExp[9]);
  p = fma(p, r, kExp[10]);
#endif
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

// Table-driven exp for kernels that can spare 512 bytes of shared memory (the likelihood / factor kernel, where the 72
// exp expansions per lane and object are a third of the FP64 work): x = (64 k + j) ln2/64 + r with |r| <= ln2/128,
// exp(x) = 2^k * T[j] * (1 + expm1(r)), expm1 by a degree-5 Taylor polynomial (truncation 3.5e-17): 10 FP64 operations
// instead of 15, one shared-memory load.  Within 1 ulp of the polynomial version (2.2e-16 against libm over [-700, 0]).
// `tab` = T[j] = 2^(j/64), j < 64, copied to shared memory by the caller (exp_table_to_shared).
static __constant__ double kExpTab[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};
static __constant__ double kExpT[9] = {92.33248261689366, 6755399441055744.0, -0.01083042469326756, -2.9815858269852933e-12,
                                       1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0};
__device__ __forceinline__ void exp_table_to_shared(double* tab, int lane) {
  tab[lane] = kExpTab[lane]; tab[lane + 32] = kExpTab[lane + 32];
}
__device__ __forceinline__ double cgp_exp_tab(double x, const double* tab) {
  double t = fma(x, kExpT[0], kExpT[1]);
  const int ki = __double2loint(t);
  t -= kExpT[1];
  double r = fma(t, kExpT[2], x);
  r = fma(t, kExpT[3], r);
  double q = fma(r, kExpT[4], kExpT[5]);
  q = fma(q, r, kExpT[6]);
  q = fma(q, r, kExpT[7]);
  q = fma(q, r, kExpT[8]);
  const double tj = tab[ki & 63];
  const double p = fma(tj, q * r, tj);
  const int hi = __double2hiint(p) + ((ki >> 6) << 20);
  const double res = __hiloint2double(hi, __double2loint(p));
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
