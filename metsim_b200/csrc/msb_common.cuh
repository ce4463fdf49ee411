// Shared device/host helpers of the metsim_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "metsim_b200.h"

namespace msb {

// ---- physical constants: values of metsim/constants.py:38-80 (low-precision literals are part
// of the numerical contract and are used verbatim) ----
constexpr double kDegPerRev = 360.0;
constexpr double kSecPerRad = 13750.9871;
constexpr double kRadPerDay = 0.017214;
constexpr double kRadPerDeg = 0.01745329;
constexpr double kMinDecl = -0.4092797;
constexpr double kDaysOff = 11.25;
constexpr double kSwRadDt = 30.0;
constexpr double kMA = 28.9644e-3;
constexpr double kR = 8.3143;
constexpr double kRDry = 287.0;
constexpr double kGStd = 9.80665;
constexpr double kPStd = 101325.0;
constexpr double kTStd = 288.15;
constexpr double kCp = 1010.0;
constexpr double kKelvin = 273.15;
constexpr double kEps = 0.62196351;
constexpr double kDaysPerYear = 365.25;
constexpr double kSecPerDay = 86400.0;
constexpr int kMinPerDay = 1440;
constexpr double kStefanB = 5.669e-8;
constexpr double kB0 = 0.031, kB1 = 0.201, kB2 = 0.185;
constexpr double kTBase = 0.870;
constexpr double kABase = -6.1e-5;
constexpr int kRows = 366;           // ceil(DAYS_PER_YEAR)
constexpr int kTiny = MSB_TINY_PER_DAY;
// window-major table (metsim_b200.h, msb_solar_tables.q_win): C = 2 ts phases x q_win_w(ts) window
// slots (tpd + 1 entries, padded to a multiple of 4 = whole 32-byte sectors), then the row total
__host__ __device__ inline int q_win_w(int ts) { return (1440 / ts + 1 + 3) & ~3; }
__host__ __device__ inline int q_win_ld(int ts) { return 2 * ts * q_win_w(ts) + 4; }   // pitch: whole sectors too
// prefetch of data a later day / event / step will load (the streaming stores keep evicting it)
__device__ __forceinline__ void prefetch_l2(const void *ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}
__device__ __forceinline__ void prefetch_l1(const void *ptr) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
}
constexpr int kNoon = 1440;          // tiny step of hour angle 0: (12*3600 + 0)/30
constexpr int kQLd = MSB_Q_LD;
constexpr int kTT = MSB_TT_TERMS;
constexpr int kState = MSB_STATE_DAYS;

// error plumbing (api.cu)
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
#define MSB_CUDA(expr)                                            \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return ::msb::cuda_fail(_e, #expr);    \
  } while (0)

#ifdef __CUDACC__
// ---- fp64 building blocks --------------------------------------------------------------------
// The sub-daily kernel is co-limited by the fp64 pipe (64 lanes/clk/SM), the issue slots and HBM
// stores, so divide / exp / sqrt are written out with exactly the DFMAs the 1e-9 contract needs
// (results are good to ~1e-15) instead of the IEEE-rounded library sequences with slow paths.
// fp64 literals cost two UMOVs each on sm_100 unless their low 32 bits are zero (then they are
// instruction immediates), so low-order-insensitive coefficients are written in that form.

// 1/x for normal, finite x: MUFU.RCP64H seed (rel. error < 2^-20, measured) + one third-order
// step y*(1 + e + e^2), e = 1 - x*y: 3 DFMA, rel. error ~e^3 < 1e-17 + rounding.
__device__ __forceinline__ double rcp_seed(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double rcp_nr(double x) {
  const double y = rcp_seed(x);
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}
__device__ __forceinline__ double div_nr(double a, double b) {
  double y = rcp_nr(b);
  double q = a * y;
  double r = fma(-b, q, a);   // one residual step: q is then within 1 ulp
  return fma(r, y, q);
}

// exp(x) for |x| < 700: k = rint(x*256/ln2), 256-entry 2^(j/256) table in shared memory,
// polynomial on |r| <= ln2/512.
//   exp_tab: one-term reduction + cubic, rel. error < 2e-13 -- for values that are only OUTPUT
//            (the 1e-9 contract leaves three orders of margin);
//   exp_acc: two-term reduction + quartic, rel. error < 4e-16 -- for the dew-point fixed point,
//            whose evaluation COUNT must match the reference (an error of 1e-13 flips the stop
//            test of mtclim.py:63 on roughly one cell in 1e5).
constexpr int kExpTab = 256;
__device__ __forceinline__ void exp_table_init(double *tab) {
  for (int j = threadIdx.x; j < kExpTab; j += blockDim.x) tab[j] = exp2((double)j / (double)kExpTab);
}
__device__ __forceinline__ double exp_scale(double p, int k, const double *tab) {
  double s = tab[k & (kExpTab - 1)];
  const int hi = __double2hiint(s) + ((k >> 8) << 20);
  s = __hiloint2double(hi, __double2loint(s));
  return p * s;
}
__device__ __forceinline__ double exp_tab(double x, const double *tab) {
  const double kMagic = 6755399441055744.0;       // 1.5 * 2^52
  double kd = fma(x, 369.3299304675746, kMagic);  // 256/ln2
  const int k = __double2loint(kd);
  kd -= kMagic;
  // k*ln2/256 carries < 2e-15 absolute error for |k| < 2^17: one fma is enough at the 1e-13 level
  const double r = fma(-kd, 0.0027076061740622863, x);
  // 1 + r + r^2/2 + r^3/6; 1/6 rounded to a 32-bit-high literal (error 1e-7 * r^3 < 1e-16)
  const double kSixth = __hiloint2double(0x3fc55555, 0);
  double p = fma(r, kSixth, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return exp_scale(p, k, tab);
}
__device__ __forceinline__ double exp_acc(double x, const double *tab) {
  const double kMagic = 6755399441055744.0;
  double kd = fma(x, 369.3299304675746, kMagic);
  const int k = __double2loint(kd);
  kd -= kMagic;
  double r = fma(-kd, 0.0027076061740622863, x);     // fl(ln2/256)
  r = fma(-kd, 9.058776616587108e-20, r);            // ln2/256 - fl(ln2/256)
  double p = fma(r, 1.0 / 24.0, 1.0 / 6.0);           // r^5/120 < 4e-17
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return exp_scale(p, k, tab);
}

// sqrt(x), x > 0 normal: MUFU.RSQ64H seed + one third-order rsqrt step (e = 1 - x*y^2;
// y *= 1 + e/2 + 3e^2/8).  sqrt_fast stops there (rel. error ~ e^3 < 1e-17 + 3 roundings),
// sqrt_nr adds one residual step (< 1 ulp).
__device__ __forceinline__ double rsqrt_step(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double xy = x * y;
  const double e = fma(-xy, y, 1.0);
  const double t = fma(e, 0.375, 0.5) * e;
  return fma(y, t, y);
}
__device__ __forceinline__ double sqrt_fast(double x) { return x * rsqrt_step(x); }
__device__ __forceinline__ double sqrt_nr(double x) {
  const double y = rsqrt_step(x);
  const double g = x * y;                // ~sqrt(x)
  const double r = fma(-g, g, x);
  return fma(r, 0.5 * y, g);
}

// physics.py:124-152 in Pa, library-accurate version
__device__ __forceinline__ double svp_ref(double t) {
  double s = 0.61078 * exp((17.269 * t) / (237.3 + t));
  if (t < 0.0) s *= 1.0 + .00972 * t + .000042 * (t * t);
  return s * 1000.0;
}

// math.remainder(lon, 360) (disaggregate.py:67) for |lon| < 1e9 without the library's bit-serial
// loop: n = rint(lon / 360) on the IEEE quotient (exact for the ties at odd multiples of 180, which
// rint then sends to the even n exactly like IEEE remainder), r = lon - 360 n is exact in one fma,
// and a quotient that rounded across a half-integer is put back by the range fix-up.
__device__ __forceinline__ double wrap_lon(double lon) {
  const double n = rint(lon / 360.0);
  double r = fma(-360.0, n, lon);
  if (r > 180.0) r -= 360.0;
  else if (r < -180.0) r += 360.0;
  return r;
}

__device__ __forceinline__ double sgn(double x) { return (double)((x > 0.0) - (x < 0.0)); }
#endif  // __CUDACC__

}  // namespace msb
