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
// The sub-daily kernel is co-limited by the fp64 pipe (64 lanes/clk/SM) and HBM stores, so the
// divide / exp / sqrt sequences are written out with exactly the number of DFMAs the 1e-9
// contract needs instead of the IEEE-rounded library sequences (which add slow-path branches).

// 1/x for normal, finite x: MUFU.RCP64H seed + 2 Newton steps (4 DFMA).  rel. error < 2^-50.
__device__ __forceinline__ double rcp_nr(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}
__device__ __forceinline__ double div_nr(double a, double b) {
  double y = rcp_nr(b);
  double q = a * y;
  double r = fma(-b, q, a);   // one residual step: q is then within 1 ulp
  return fma(r, y, q);
}

// exp(x) for |x| < 700: k = rint(x*64/ln2), 64-entry 2^(j/64) table in shared memory,
// degree-5 polynomial on |r| <= ln2/128.  rel. error < 3e-16.
struct ExpTable { double t[64]; };
__device__ __forceinline__ void exp_table_init(double *tab) {
  for (int j = threadIdx.x; j < 64; j += blockDim.x) tab[j] = exp2((double)j / 64.0);
}
__device__ __forceinline__ double exp_tab(double x, const double *tab) {
  const double kInvL = 92.33248261689366;         // 64/ln2
  const double kLhi = 0.010830424696249145;      // fl(ln2/64)
  const double kLlo = 3.623510646634843e-19;     // ln2/64 - fl(ln2/64)
  const double kMagic = 6755399441055744.0;       // 1.5 * 2^52
  double kd = fma(x, kInvL, kMagic);
  int k = __double2loint(kd);
  kd -= kMagic;
  double r = fma(-kd, kLhi, x);
  r = fma(-kd, kLlo, r);
  double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  double s = tab[k & 63];
  int hi = __double2hiint(s) + ((k >> 6) << 20);
  s = __hiloint2double(hi, __double2loint(s));
  return p * s;
}

// sqrt(x), x > 0 normal: MUFU.RSQ64H seed + Newton on rsqrt + one residual step.
__device__ __forceinline__ double sqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double h = 0.5 * y;
  double e = fma(-x * y, h, 0.5);  // 0.5 - 0.5*x*y^2
  y = fma(y, e, y);
  h = 0.5 * y;
  e = fma(-x * y, h, 0.5);
  y = fma(y, e, y);
  double g = x * y;                // ~sqrt(x)
  double r = fma(-g, g, x);
  return fma(r, 0.5 * y, g);
}

// physics.py:124-152 in Pa, library-accurate version used by the daily pass
__device__ __forceinline__ double svp_ref(double t) {
  double s = 0.61078 * exp((17.269 * t) / (237.3 + t));
  if (t < 0.0) s *= 1.0 + .00972 * t + .000042 * (t * t);
  return s * 1000.0;
}

__device__ __forceinline__ double sgn(double x) { return (double)((x > 0.0) - (x < 0.0)); }
#endif  // __CUDACC__

}  // namespace msb
