// Sub-daily disaggregation, all variables fused in one streaming kernel.
//
// Reference: metsim/disaggregate.py:34-130 (disaggregate) and the kernels it calls --
// set_min_max_hour :133-201, temp :204-262 (scipy PCHIP), prec :265-352, wind :355-380,
// pressure :383-403, specific_humidity :406-424, relative_humidity :427-447,
// vapor_pressure :450-488 (scipy interp1d), longwave :491-590, tskc :593-615,
// shortwave :618-680; unit converters metsim/units.py:3-59.
//
// The reference materialises per-minute / per-30-second vectors per cell and reduces them with
// numpy.  Here one thread walks one cell through a tile of days with every intermediate in
// registers; lanes are consecutive cells, so each of the (up to) 10 output streams is written
// as full 128-byte lines.  What the reference gets from vectors is obtained in closed form:
//   * np.roll + reshape(-1, ts).sum/mean  -> split of each output window at the (shifted) day
//     boundary (tskc, wind, prec) -- integer arithmetic, bit-exact placement;
//   * the per-minute triangle             -> differences of a cumulative arithmetic series whose
//     positions are clamped in INTEGER arithmetic (exact zeros outside the limbs);
//   * bincount over 2*ts table entries    -> difference of two entries of the signed prefix table
//     of the latitude row (solar.cu);
//   * PchipInterpolator / interp1d        -> the current cubic / linear segment in registers.
//
// Loop structure: the kernel is EVENT DRIVEN.  Everything that happens less than once per step --
// a PCHIP or vapour-pressure knot is crossed, a source day or table row rolls over, the radiation
// window crosses solar noon -- is a state update done by a cold handler when the step counter
// reaches the next scheduled event.  Between events the warp runs a tight, call-free, branch-free
// inner loop of straight-line fp64 arithmetic; the run length is the warp-wide minimum of the
// lanes' next events (one REDUX per run), so lanes never lose lockstep on the step index.
// Values that only the handler needs between events live in per-thread shared-memory slots, and
// the one table load of the hot path is software-pipelined one step ahead.
#include <cstdlib>
#include <mutex>

#include "msb_common.cuh"

namespace msb {

constexpr int kDisaggThreads = 128;
constexpr int kNever = 0x7fffffff;
constexpr int kDayWin = 40;    // staged day-of-year / month window (tile + margins)

struct DisaggArgs {
  msb_params p;
  msb_forcing f;
  msb_solar_tables t;
  const double *vp_d, *sw_d, *tskc_d;  // daily [D][ld] from the MTCLIM pass
  int32_t *status;
  msb_subdaily out;
  int tpd, tile_days;
  unsigned var_mask;   // bit v set = out.var[v] requested
  double *rec;         // day records [MSB_REC_FIELDS][n_rec][ld], record r = padded day e0 + r
  double *lw_raw;      // passthrough with a supplied daily longwave: unscaled float64 longwave
                       // [(day1-day0)*tpd][out.ld] for the rescale pass, else NULL
  double inv_sw_den;   // 1 / (3600 * ts / 60), disaggregate.py:645
  int prep_days;       // padded days per thread of disagg_prepare_kernel
  int prep_r0, prep_r1;  // records [prep_r0, prep_r1) are (re)built by this launch
  bool sw_full_day;    // passthrough without sw_averaging='daylight': daylength := 86400 (disaggregate.py:78-80)
  // work range of the interior-day kernels (thermo_knot_kernel, streams_day_kernel)
  struct { int e0, e1, tile, i_lo, i_hi, pf; } k;   // pf: L1 prefetch distance of the table stream (steps)
  int e0, n_rec;       // e runs over [-1, D]: -1 / D are the padded PCHIP ends
};

// fields of a day record
enum { kR_XMIN = 0, kR_DMIN,            // PCHIP knot at t_Tmin: time, derivative (its value is the
       kR_XMAX, kR_DMAX,                 //   input t_min / t_max itself); same for the knot at t_Tmax
       kR_PPACK, kR_HSP, kR_HSQ, kR_PF, // PrecDay: packed integers, stepp/2, stepq/2, P/sum
       kR_RAD };                         // shortwave per tiny step, disaggregate.py:645

// ---- per-thread view of one cell (cold: only the event handler and the set-up read it) --------
struct Cell {
  int cell, D, ts, tpd, nk;
  long ld;
  int off_min;               // np.roll offset in minutes, trunc toward zero
  double utc_min;            // sunrise/sunset shift, disaggregate.py:184-189
  double frac;
  const double *rise, *set;  // latitude-row tables [366]
  const int32_t *doy;        // global [D]
  const int *sdoy;           // shared copy of doy[win0 .. win0+nwin)
  int win0, nwin;
  const double *t_min, *t_max, *st_t_min, *st_t_max, *t_end;
};
__device__ __forceinline__ int doy_at(const Cell &c, int d) {
#ifdef MSB_BOUNDS
  if ((unsigned)d >= (unsigned)c.D) { asm volatile("trap;"); }
#endif
  const unsigned r = (unsigned)(d - c.win0);
  return r < (unsigned)c.nwin ? c.sdoy[r] : c.doy[d];
}

// Knot times are built with explicitly rounded operations (numpy never contracts a*b+c, and the
// same knot must come out bit-identical from every place that evaluates it).
// knot j of the padded PCHIP sequence (disaggregate.py:236-257); j in [0, 2D+4)
__device__ __forceinline__ double knot_x(const Cell &c, int j) {
  const int e = (j >> 1) - 1;
  const int de = min(max(e, 0), c.D - 1);
  const int yd = doy_at(c, de) - 1;
  const double off = (double)de * 1440.0;
  const double rt = __dadd_rn(__dadd_rn(c.rise[yd], -c.utc_min), off);   // :188-195
  double x;
  if (j & 1) {
    const double st = __dadd_rn(__dadd_rn(c.set[yd], -c.utc_min), off);
    x = __dadd_rn(__dadd_rn(__dmul_rn(c.frac, __dadd_rn(st, -rt)), rt), -(double)c.ts);  // t_t_max :198-199
  } else {
    x = __dadd_rn(rt, -(double)c.ts);                                    // t_t_min :200
  }
  return __dadd_rn(x, 1440.0 * (double)(e - de));                        // padded ends :242-243
}
__device__ __forceinline__ double knot_y(const Cell &c, int j) {
  const int e = (j >> 1) - 1;
  if (e < 0)
    return (j & 1) ? c.st_t_max[(size_t)(kState - 1) * c.ld + c.cell]
                   : c.st_t_min[(size_t)(kState - 1) * c.ld + c.cell];  // t_begin, metsim.py:753
  if (e >= c.D) {
    if (c.t_end) return c.t_end[(size_t)(j & 1) * c.ld + c.cell];
    return (j & 1) ? c.t_max[(size_t)(c.D - 1) * c.ld + c.cell]
                   : c.t_min[(size_t)(c.D - 1) * c.ld + c.cell];
  }
  return (j & 1) ? c.t_max[(size_t)e * c.ld + c.cell] : c.t_min[(size_t)e * c.ld + c.cell];
}

// scipy _cubic.py:279-291
__device__ __forceinline__ double pchip_edge(double h0, double h1, double m0, double m1) {
  const double num = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(2.0, h0), h1), m0), -__dmul_rn(h0, m1));
  const double d = div_nr(num, __dadd_rn(h0, h1));
  if (sgn(d) != sgn(m0)) return 0.0;
  if (sgn(m0) != sgn(m1) && fabs(d) > 3.0 * fabs(m0)) return 3.0 * m0;
  return d;
}
// scipy _cubic.py:320-333 (weighted harmonic mean, one division)
__device__ __forceinline__ double pchip_mid(double h0, double h1, double m0, double m1) {
  if (sgn(m1) != sgn(m0) || m1 == 0.0 || m0 == 0.0) return 0.0;
  const double w1 = __dadd_rn(__dmul_rn(2.0, h1), h0), w2 = __dadd_rn(h1, __dmul_rn(2.0, h0));
  const double num = __dmul_rn(__dadd_rn(w1, w2), __dmul_rn(m0, m1));
  const double den = __dadd_rn(__dmul_rn(w1, m1), __dmul_rn(w2, m0));
  return div_nr(num, den);
}

struct Cubic {
  double x0, x1;         // knot times of the interval
  double c0, c1, c2, c3; // scipy power basis (CubicHermiteSpline, _cubic.py:158-165)
};
// what the next interval can reuse: knot k+1's value, knot k+2, the slope k+1 -> k+2 and the
// derivative at knot k+1
struct CubicNext { double y1, x2, y2, m2, d1; };

__device__ __forceinline__ void cubic_coeffs(Cubic &q, double xb, double xc, double yb, double h1,
                                             double m1, double d0, double d1) {
  const double inv = rcp_nr(h1);
  const double tt = __dmul_rn(__dadd_rn(__dadd_rn(d0, d1), -__dmul_rn(2.0, m1)), inv);
  q.x0 = xb;
  q.x1 = xc;
  q.c0 = __dmul_rn(tt, inv);
  q.c1 = __dadd_rn(__dmul_rn(__dadd_rn(m1, -d0), inv), -tt);
  q.c2 = d0;
  q.c3 = yb;
}

// coefficients of interval k (knots k, k+1) from the four knots k-1 .. k+2; returns whether `nx`
// was filled (k+2 exists)
__device__ __forceinline__ bool cubic_setup(const Cell &c, int k, Cubic &q, CubicNext &nx,
                                            int32_t *status) {
  const bool has_a = k > 0, has_d = k + 2 < c.nk;
  const double xb = knot_x(c, k), xc = knot_x(c, k + 1);
  const double xa = has_a ? knot_x(c, k - 1) : 0.0, xd = has_d ? knot_x(c, k + 2) : 0.0;
  const double yb = knot_y(c, k), yc = knot_y(c, k + 1);
  const double ya = has_a ? knot_y(c, k - 1) : 0.0, yd = has_d ? knot_y(c, k + 2) : 0.0;
  const double h1 = __dadd_rn(xc, -xb);
  const double m1 = div_nr(__dadd_rn(yc, -yb), h1);
  bool bad = !(h1 > 0.0);
  double h0 = 0, m0 = 0, h2 = 0, m2 = 0;
  if (has_a) {
    h0 = __dadd_rn(xb, -xa);
    m0 = div_nr(__dadd_rn(yb, -ya), h0);
    bad |= !(h0 > 0.0);
  }
  if (has_d) {
    h2 = __dadd_rn(xd, -xc);
    m2 = div_nr(__dadd_rn(yd, -yc), h2);
    bad |= !(h2 > 0.0);
  }
  const double d0 = has_a ? pchip_mid(h0, h1, m0, m1) : pchip_edge(h1, h2, m1, m2);
  const double d1 = has_d ? pchip_mid(h1, h2, m1, m2) : pchip_edge(h1, h0, m1, m0);
  if (bad) status[2] = 1;
  cubic_coeffs(q, xb, xc, yb, h1, m1, d0, d1);
  nx.y1 = yc; nx.x2 = xd; nx.y2 = yd; nx.m2 = m2; nx.d1 = d1;
  return has_d;
}
__device__ __forceinline__ double cubic_eval(const Cubic &q, double x) {
  const double s = x - q.x0;
  return fma(fma(fma(q.c0, s, q.c1), s, q.c2), s, q.c3);
}

// PPoly interval search: clip(searchsorted_right(x_knots, x) - 1, 0, nk-2), starting from a guess
// at or left of the answer and walking right.
__device__ __forceinline__ int cubic_walk(const Cell &c, int k, double x) {
  while (k < c.nk - 2) {
    if (x >= knot_x(c, k + 1)) ++k; else break;
  }
  return k;
}

// smallest step index i with ts*i >= xk  /  ts*i > xk (kNever beyond the int range)
__device__ __forceinline__ int first_step_ge(double xk, int ts, double inv_ts) {
  const double q = ceil(xk * inv_ts);
  if (!(q < 2.0e9)) return kNever;
  if (q < -2.0e9) return -kNever;
  int i = (int)q;
  while ((double)ts * (double)i < xk) ++i;
  while ((double)ts * (double)(i - 1) >= xk) --i;
  return i;
}
__device__ __forceinline__ int first_step_gt(double xk, int ts, double inv_ts) {
  const double q = floor(xk * inv_ts) + 1.0;
  if (!(q < 2.0e9)) return kNever;
  if (q < -2.0e9) return -kNever;
  int i = (int)q;
  while (!((double)ts * (double)i > xk)) ++i;
  while ((double)ts * (double)(i - 1) > xk) --i;
  return i;
}

// L2 prefetch of data a later event / step will load (the streaming stores keep evicting it)
__device__ __forceinline__ void prefetch_l2(const void *ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}
__device__ __forceinline__ void prefetch_l1(const void *ptr) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
}
// a load the scheduler must not sink towards its use (issued where it is written)
__device__ __forceinline__ double load_early(const double *ptr) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(ptr));
  return v;
}

// t_t_min of day d: knot of the vapour-pressure interpolant (disaggregate.py:480) = PCHIP knot 2d+2
__device__ __forceinline__ double vp_knot_x(const Cell &c, int d) { return knot_x(c, 2 * d + 2); }

// physics.py:124-152 on the fast fp64 path; returns svp in kPa (svp/MBAR_PER_BAR)
__device__ __forceinline__ double svp_kpa(double t, const double *etab) {
  const double r = rcp_nr(237.3 + t);
  // 17.269*t/(237.3+t) = 17.269 - 17.269*237.3/(237.3+t)
  const double arg = fma(-17.269 * 237.3, r, 17.269);
  const double s = 0.61078 * exp_tab(arg, etab);
  const double corr = fma(fma(.000042, t, .00972), t, 1.0);
  return s * (t < 0.0 ? corr : 1.0);
}

// ---- hot-loop constants ----------------------------------------------------------------------
// ptxas re-materialises an fp64 literal as two 32-bit moves at EVERY use (it never keeps an
// immediate in a register across a loop), which cost ~45 instructions per step.  Values loaded
// from constant memory are opaque to it, so they are loaded once and stay in (uniform) registers.
__constant__ double c_hot[12] = {
    237.3,                          // svp denominator offset, physics.py:149
    -17.269 * 237.3,                // 17.269*t/(237.3+t) = 17.269 - 17.269*237.3/(237.3+t)
    17.269 - 0.49301845011612494,   // + ln(0.61078): the factor folded into the exponent
    .000042, .00972,                // ice correction, physics.py:150-151
    369.3299304675746,              // 256/ln2
    0.0027076061740622863,          // ln2/256
    4.618333172514372,              // ln(P_STD/1000): air pressure in kPa, disaggregate.py:401-403
    273.15,                         // KELVIN
    1.2,                            // PRATA, disaggregate.py:557-559
    0.16666666666666666,            // exp polynomial
    0.375                           // rsqrt step
};
struct HotK { double k237, ka, kb, kc2, kc1, inv_l, l, ln_p, kelvin, k12, sixth, c375; };
__device__ __forceinline__ HotK load_hot() {
  HotK k;
  k.k237 = c_hot[0]; k.ka = c_hot[1]; k.kb = c_hot[2]; k.kc2 = c_hot[3]; k.kc1 = c_hot[4];
  k.inv_l = c_hot[5]; k.l = c_hot[6]; k.ln_p = c_hot[7]; k.kelvin = c_hot[8]; k.k12 = c_hot[9];
  k.sixth = c_hot[10]; k.c375 = c_hot[11];
  return k;
}
// exp_tab with its two non-immediate constants taken from HotK
__device__ __forceinline__ double exp_hot(double x, const double *tab, const HotK &K) {
  const double kMagic = 6755399441055744.0;       // 1.5 * 2^52
  double kd = fma(x, K.inv_l, kMagic);
  const int k = __double2loint(kd);
  kd -= kMagic;
  const double r = fma(-kd, K.l, x);
  double p = fma(r, K.sixth, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return exp_scale(p, k, tab);
}
// sqrt_fast with its non-immediate constant taken from HotK
__device__ __forceinline__ double sqrt_hot(double x, const HotK &K) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double xy = x * y;
  const double e = fma(-xy, y, 1.0);
  const double t = fma(e, K.c375, 0.5) * e;
  return x * fma(y, t, y);
}

// ---- precipitation: one day's per-minute shape in closed form (disaggregate.py:294-329) ------
// The day vector is: rising limb np.linspace(0, 1, n+) on minutes [p0, p0+n+), falling limb
// np.linspace(1, 0, n-) on [q0, q0+n-), the falling assignment second (it overwrites a shared
// peak minute), everything scaled by P / sum.  linspace is i*step with the LAST element forced to
// exactly 1 (rising) or 0 (falling).  The cumulative sum of the unscaled shape up to minute m is
//     S(m) = [m >= pth] + j + hsp * (i-1) i + hsq * (j-1) j,
//     i = clamp(m - p0, 0, nps),  j = clamp(m - q0, 0, nqs)       (integers)
// and a window is f * (S(end) - S(start)).  Both ends of a window that misses the limbs clamp to
// the same integers, so its value is an exact 0: zero / non-zero placement is bit-exact.  A
// uniform day (prec_UNIFORM, :332-343) is the same form with q0 = 0, nqs = 1440, hsq = 0.
struct PrecDay {
  int p0, nps;   // rising series minutes [p0, p0+nps): value i*stepp
  int pth;       // first minute after the forced-1.0 element, kNever if none
  int q0, nqs;   // falling series minutes [q0, q0+nqs): value 1 + j*stepq (forced 0.0 end excluded)
  double hsp, hsq, f;   // stepp/2, stepq/2, P/sum (NaN for a 0/0 day, :327)
};

// exact int -> double for 0 <= n < 2^31 on the fp64 pipe (I2F.F64 runs on the quarter-rate
// conversion unit, which the hot loop already loads with its eight F2F and five MUFU per step)
__device__ __forceinline__ double u2d(int n) {
  return __hiloint2double(0x43300000, n) - 4503599627370496.0;   // (2^52 + n) - 2^52
}
__device__ __forceinline__ double prec_cum(const PrecDay &pd, int m) {
  const int i = min(max(m - pd.p0, 0), pd.nps);
  const int j = min(max(m - pd.q0, 0), pd.nqs);
  const int u = (m >= pd.pth) ? 1 : 0;
  const int tr = i * i - i;   // 2 * sum_{k<i} k
  const int tf = j * j - j;
  return fma(pd.hsq, u2d(tf), fma(pd.hsp, u2d(tr), u2d(u + j)));
}

// m = calendar month of day d (1..12)
__device__ __forceinline__ void prec_day_setup(const msb_params &p, const msb_forcing &f,
                                               int cell, int d, int m, PrecDay &pd) {
  const size_t o = (size_t)d * f.ld + cell;
  const double P = f.prec[o];
  const bool uniform = p.prec_type == MSB_PREC_UNIFORM ||
                       (p.prec_type == MSB_PREC_MIX && f.t_min[o] < 0.0);
  if (uniform) {
    pd.p0 = 0; pd.nps = 0; pd.pth = kNever; pd.q0 = 0; pd.nqs = kMinPerDay;
    pd.hsp = 0.0; pd.hsq = 0.0;
    pd.f = P * (1.0 / (double)kMinPerDay);
    return;
  }
  // metsim.py:278-284: month m (1..12) takes domain entry m (0..11); December keeps prec
  const double t_pk = (m <= 11) ? f.t_pk12[(size_t)m * f.ld + cell] : P;
  const double t_dur = (m <= 11) ? f.dur12[(size_t)m * f.ld + cell] : P;
  const double t_start = t_pk - 0.5 * t_dur, t_stop = t_pk + 0.5 * t_dur;
  // integer minutes m with t_start <= m <= t_pk, and t_pk <= m <= t_stop, clipped to the day
  const double lo_p = fmax(ceil(t_start), 0.0), hi_p = fmin(floor(t_pk), 1439.0);
  const double lo_q = fmax(ceil(t_pk), 0.0), hi_q = fmin(floor(t_stop), 1439.0);
  const int npf = (hi_p >= lo_p) ? (int)(hi_p - lo_p) + 1 : 0;
  const int nq = (hi_q >= lo_q) ? (int)(hi_q - lo_q) + 1 : 0;
  pd.p0 = (npf > 0) ? (int)lo_p : 0;
  pd.q0 = (nq > 0) ? (int)lo_q : 0;
  const double stepp = (npf > 1) ? rcp_nr((double)(npf - 1)) : 0.0;   // np.linspace(0, 1, npf)
  const double stepq = (nq > 1) ? -rcp_nr((double)(nq - 1)) : 0.0;    // np.linspace(1, 0, nq)
  // the falling assignment comes second and overwrites the shared peak minute (:322-323)
  int np = npf;
  if (nq > 0 && npf > 0 && pd.p0 + npf - 1 >= pd.q0) np = max(pd.q0 - pd.p0, 0);
  const bool unit = np == npf && npf > 1;                  // forced 1.0 survives the overwrite
  pd.nps = unit ? npf - 1 : np;
  pd.pth = unit ? pd.p0 + npf : kNever;
  pd.nqs = nq > 1 ? nq - 1 : nq;                           // forced 0.0 end contributes nothing
  pd.hsp = 0.5 * stepp;
  pd.hsq = 0.5 * stepq;
  const double S = prec_cum(pd, kMinPerDay);               // np.sum(prec_day), :327
  pd.f = (S > 0.0 && P == P) ? div_nr(P, S) : nan("");     // 0/0 or x/0 * 0 -> NaN day (:327)
}

// value of output step i for the rolled per-minute record (disaggregate.py:347-351); NaN when the
// window touches a NaN day.  Generic form used by the NaN-day fix-up pass only.
__device__ __noinline__ double prec_step_generic(const msb_params &p, const msb_forcing &f,
                                                 const Cell &c, long i) {
  const long M = (long)c.D * kMinPerDay;
  long m0 = (i * c.ts + c.off_min) % M;
  if (m0 < 0) m0 += M;
  const int dA = (int)(m0 / kMinPerDay), rem = (int)(m0 - (long)dA * kMinPerDay);
  const int n1 = min(c.ts, kMinPerDay - rem);
  PrecDay pd;
  prec_day_setup(p, f, c.cell, dA, f.month[dA], pd);
  double v = pd.f * (prec_cum(pd, rem + n1) - prec_cum(pd, rem));
  if (n1 < c.ts) {
    const int dB = (dA + 1 == c.D) ? 0 : dA + 1;
    prec_day_setup(p, f, c.cell, dB, f.month[dB], pd);
    v += pd.f * prec_cum(pd, c.ts - n1);
  }
  return v;
}

// value of the last valid step before step i (what pandas ffill repeats); NaN if there is none
__device__ __noinline__ double prec_lookback(const msb_params &p, const msb_forcing &f, const Cell &c,
                                             long i) {
  for (long j = i - 1; j >= 0; --j) {
    double v = prec_step_generic(p, f, c, j);
    if (v == v) return v;
  }
  return nan("");
}

// vapour pressure of ONE step at an end of the interpolant's support: the value pandas
// back-fills into the leading NaNs (lo = true: knots 0, 1) or forward-fills into the trailing
// ones (knots D-2, D-1) -- disaggregate.py:480-487, :99-101.
__device__ __noinline__ double vp_endpoint(const Cell &c, const double *vp_d, int i, bool lo,
                                           const double *etab, int32_t *status) {
  const double x = (double)c.ts * (double)i;
  Cubic q;
  CubicNext nx;
  const int k = cubic_walk(c, max(0, 2 * (i / c.tpd + 1) - 4), x);
  cubic_setup(c, k, q, nx, status);
  const double tt = cubic_eval(q, x);
  const int d0 = lo ? 0 : c.D - 2;
  const double x0 = vp_knot_x(c, d0), x1 = vp_knot_x(c, d0 + 1);
  const double y0 = vp_d[(size_t)d0 * c.ld + c.cell] / 1000.0;
  const double y1 = vp_d[(size_t)(d0 + 1) * c.ld + c.cell] / 1000.0;
  const double v = ((x - x0) / (x1 - x0)) * y1 + ((x1 - x) / (x1 - x0)) * y0;
  return fmin(v, svp_kpa(tt, etab));
}

// disaggregate.py:491-590 without the daily_lw rescale (mtclim method): every lw_type / lw_cloud
// option with library math (the default PRATA + CLOUD_DEARDORFF pair is inlined in the hot loop)
__device__ __forceinline__ double longwave(const msb_params &p, double temp_c, double vp_kpa,
                                           double tskc) {
  const double air_temp = temp_c + kKelvin;
  const double vp = vp_kpa * 10;
  double e;
  switch (p.lw_type) {
    case MSB_LW_ANDERSON: e = 0.68 + 0.036 * sqrt(vp); break;
    case MSB_LW_BRUTSAERT: e = 1.24 * pow(vp / air_temp, 0.14285714); break;
    case MSB_LW_SATTERLUND: e = 1.08 * (1 - exp(-1 * pow(vp, (air_temp / 2016)))); break;
    case MSB_LW_IDSO: e = 0.7 + 5.95e-5 * vp * exp(1500 / air_temp); break;
    case MSB_LW_PRATA: {
      const double x = 46.5 * vp / air_temp;
      e = (1 - (1 + x) * exp(-sqrt((1.2 + 3. * x))));
      break;
    }
    default: e = 0.74 + 0.0049 * vp; break;
  }
  const double emis = (p.lw_cloud == MSB_CLOUD_DEARDORFF) ? tskc + (1 - tskc) * e
                                                          : (1.0 + (0.17 * (tskc * tskc))) * e;
  const double t2 = air_temp * air_temp;
  return emis * kStefanB * (t2 * t2);
}

// variable sets compiled into the kernel: kVarsExample = the 8 variables of
// examples/example_nc.conf (every pointer known non-NULL, no per-step tests), kVarsAny = runtime
// The fixed sets come as one kernel (kVarsExample) or as two that split the per-thread state --
// kVarsThermo (temperature, humidity, pressure, longwave: the fp64-heavy half) and kVarsStreams
// (precipitation, shortwave, wind: the source-stream half) -- so that each half keeps fewer
// registers and more warps are resident to cover the 8-cycle dependent DFMA latency.
constexpr int kVarsAny = 0, kVarsExample = 1, kVarsThermo = 2, kVarsStreams = 3;
constexpr unsigned kThermoBits = (1u << MSB_V_TEMP) | (1u << MSB_V_LONGWAVE) | (1u << MSB_V_VAPOR_PRESSURE) |
                                 (1u << MSB_V_REL_HUMID) | (1u << MSB_V_AIR_PRESSURE) | (1u << MSB_V_SPEC_HUMID);
constexpr unsigned kExampleMask = (1u << MSB_V_TEMP) | (1u << MSB_V_PREC) | (1u << MSB_V_SHORTWAVE) |
                                  (1u << MSB_V_LONGWAVE) | (1u << MSB_V_VAPOR_PRESSURE) |
                                  (1u << MSB_V_REL_HUMID) | (1u << MSB_V_AIR_PRESSURE) |
                                  (1u << MSB_V_WIND);

// per-thread shared-memory slots for values that only the event handler touches between events
enum { kS_TKN = 0, kS_WDN,   // tskc / wind of the source day just rolled in
       kS_ROWTOT,            // total of the current radiation row (noon event)
       kS_VPTAIL,            // forward-fill value of the vapour pressure
       kS_COUNT };

// ---- day records ------------------------------------------------------------------------------
// Everything an event needs about a (cell, day) -- which the per-cell reference code re-derives
// from whole-record vectors -- is computed once by a thread-per-(cell, day) kernel and read back
// by the streaming kernel with single coalesced loads: the PCHIP knots (time, value and the
// derivative scipy's PchipInterpolator would assign, _cubic.py:279-340), the vapour-pressure knot,
// the closed-form precipitation shape and the radiation per tiny step.
__device__ __forceinline__ double pack_prec(const PrecDay &pd) {
  const unsigned lo = (unsigned)pd.p0 | ((unsigned)pd.nps << 16);
  const unsigned hi = (unsigned)pd.q0 | ((unsigned)pd.nqs << 16) | (pd.pth != kNever ? 0x80000000u : 0u);
  return __hiloint2double((int)hi, (int)lo);
}
__device__ __forceinline__ void unpack_prec(double w, PrecDay &pd) {
  const unsigned lo = (unsigned)__double2loint(w), hi = (unsigned)__double2hiint(w);
  pd.p0 = (int)(lo & 0xffffu); pd.nps = (int)(lo >> 16);
  pd.q0 = (int)(hi & 0xffffu); pd.nqs = (int)((hi >> 16) & 0x7fffu);
  pd.pth = (hi & 0x80000000u) ? pd.p0 + pd.nps + 1 : kNever;   // forced 1.0 follows the series
}

// One thread = one (cell, padded day e): the two knots of the day from the four knots around them
// (interior: weighted harmonic mean of the neighbouring slopes; record ends: scipy's three-point
// edge rule, _cubic.py:279-291, :320-333).
constexpr int kPrepDays = 8;   // padded days per thread of the record kernel
__global__ void __launch_bounds__(128)
disagg_prepare_kernel(const __grid_constant__ DisaggArgs a) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= a.f.n_cells) return;
  const msb_params &p = a.p;
  const msb_forcing &f = a.f;
  const int D = f.n_days, ts = p.time_step;
  const int r_lo = a.prep_r0 + blockIdx.y * a.prep_days, r_hi = min(r_lo + a.prep_days, a.prep_r1);
  Cell c;
  c.cell = cell; c.D = D; c.ts = ts; c.tpd = a.tpd; c.nk = 2 * D + 4; c.ld = f.ld;
  const int nk = c.nk;
  const double lon = wrap_lon(f.lon[cell]);                            // disaggregate.py:67
  c.off_min = p.utc_offset ? (int)((lon / kDegPerRev) * kMinPerDay) : 0;
  c.utc_min = p.utc_offset ? (double)(int)(((lon - 0.0) / kDegPerRev) * kMinPerDay) : 0.0;  // :186
  c.frac = p.tmax_daylength_fraction;
  const int row = f.lat_row[cell];
  c.rise = a.t.rise_min + (size_t)row * kRows;
  c.set = a.t.set_min + (size_t)row * kRows;
  c.doy = f.doy; c.sdoy = nullptr; c.win0 = 0; c.nwin = 0;
  c.t_min = f.t_min; c.t_max = f.t_max; c.st_t_min = f.st_t_min; c.st_t_max = f.st_t_max;
  c.t_end = f.t_end;
  const size_t plane = (size_t)a.n_rec * f.ld;
  const double *const dayl = a.t.daylength + (size_t)row * kRows;
  const double inv_sw_den = a.inv_sw_den;
  // the per-cell constants above are shared by kPrepDays consecutive padded days
#pragma unroll 1
  for (int r = r_lo; r < r_hi; ++r) {
  const int e = a.e0 + r;
  double *out = a.rec + (size_t)r * f.ld + cell;

  // knots j0-1 .. j0+2 (j0 = t_Tmin knot of padded day e); all loads are independent
  const int j0 = 2 * (e + 1);
  const int ja = max(j0 - 1, 0), jd = min(j0 + 2, nk - 1);
  const double xa = knot_x(c, ja), xb = knot_x(c, j0), xc = knot_x(c, j0 + 1), xd = knot_x(c, jd);
  const double ya = knot_y(c, ja), yb = knot_y(c, j0), yc = knot_y(c, j0 + 1), yd = knot_y(c, jd);
  const double h1 = __dadd_rn(xc, -xb);
  const double m1 = div_nr(__dadd_rn(yc, -yb), h1);
  bool bad = !(h1 > 0.0);
  double d_min, d_max;
  if (j0 > 0 && j0 + 1 < nk - 1) {            // both knots interior
    const double h0 = __dadd_rn(xb, -xa), h2 = __dadd_rn(xd, -xc);
    const double m0 = div_nr(__dadd_rn(yb, -ya), h0), m2 = div_nr(__dadd_rn(yd, -yc), h2);
    bad |= !(h0 > 0.0) || !(h2 > 0.0);
    d_min = pchip_mid(h0, h1, m0, m1);
    d_max = pchip_mid(h1, h2, m1, m2);
  } else if (j0 == 0) {                       // first padded day: knot 0 takes the edge rule
    const double h2 = __dadd_rn(xd, -xc);
    const double m2 = div_nr(__dadd_rn(yd, -yc), h2);
    bad |= !(h2 > 0.0);
    d_min = pchip_edge(h1, h2, m1, m2);
    d_max = pchip_mid(h1, h2, m1, m2);
  } else {                                    // last padded day: knot nk-1 takes the edge rule
    const double h0 = __dadd_rn(xb, -xa);
    const double m0 = div_nr(__dadd_rn(yb, -ya), h0);
    bad |= !(h0 > 0.0);
    d_min = pchip_mid(h0, h1, m0, m1);
    d_max = pchip_edge(h1, h0, m1, m0);
  }
  if (bad) a.status[2] = 1;                   // scipy: `x` must be strictly increasing
  out[kR_XMIN * plane] = xb;
  out[kR_DMIN * plane] = d_min;
  out[kR_XMAX * plane] = xc;
  out[kR_DMAX * plane] = d_max;
  if (e < 0 || e >= D) continue;                    // padded ends carry knots only
  const size_t o = (size_t)e * f.ld + cell;
  PrecDay pd;
  pd.p0 = pd.nps = pd.q0 = pd.nqs = 0; pd.pth = kNever; pd.hsp = pd.hsq = pd.f = 0.0;
  if (a.out.var[MSB_V_PREC]) prec_day_setup(p, f, cell, e, f.month[e], pd);
  out[kR_PPACK * plane] = pack_prec(pd);
  out[kR_HSP * plane] = pd.hsp;
  out[kR_HSQ * plane] = pd.hsq;
  out[kR_PF * plane] = pd.f;
  const double dl = a.sw_full_day ? kSecPerDay : dayl[f.doy[e] - 1];
  out[kR_RAD * plane] = (a.sw_d[o] * dl) * inv_sw_den;
  }
}

// One thread = one cell x one tile of days.  np.roll by a whole number of minutes (tskc, wind,
// prec) or 30 s steps (shortwave) turns every daily series into a contiguous SOURCE STREAM that
// is read ts minutes per output step; the stream positions and the current source day live in
// registers and roll to the next source day once per day (an event).
constexpr int disagg_min_blocks(int vars) { return vars == kVarsStreams ? 6 : (vars == kVarsThermo ? 7 : 4); }

template <typename OutT, bool FAST_LW, bool UNITS, int VARS>
__global__ void __launch_bounds__(kDisaggThreads, disagg_min_blocks(VARS))
disagg_kernel(const __grid_constant__ DisaggArgs a) {
  __shared__ double etab[kExpTab];
  __shared__ double slot[kS_COUNT][kDisaggThreads];
  __shared__ int s_doy[kDayWin], s_month[kDayWin];
  const msb_params &p = a.p;
  const msb_forcing &f = a.f;
  const int D = f.n_days, ts = p.time_step, tpd = a.tpd;
  const int d_lo = a.out.day0 + blockIdx.y * a.tile_days;
  const int d_hi = min(a.out.day1, d_lo + a.tile_days);
  // day-of-year / month of the days the handler can touch, staged once per CTA
  const int win0 = max(d_lo - 4, 0), nwin = min(min(d_hi + 4, D) - win0, kDayWin);
  exp_table_init(etab);
  for (int j = threadIdx.x; j < nwin; j += blockDim.x) {
    s_doy[j] = f.doy[win0 + j];
    s_month[j] = f.month[win0 + j];
  }
  __syncthreads();

  const int tid = threadIdx.x;
  const int cell = blockIdx.x * blockDim.x + tid;
  if (cell >= f.n_cells || d_lo >= d_hi) return;
  const unsigned mask = VARS == kVarsExample ? kExampleMask
                      : VARS == kVarsThermo ? (kExampleMask & kThermoBits)
                      : VARS == kVarsStreams ? (kExampleMask & ~kThermoBits) : a.var_mask;
  auto want = [mask](int v) { return (mask >> v) & 1u; };
  const bool want_prec = want(MSB_V_PREC);
  const bool want_wind = want(MSB_V_WIND) && (VARS != kVarsAny || f.wind != nullptr);
  const bool want_sw = want(MSB_V_SHORTWAVE);
  const bool thermo = (mask & kThermoBits) != 0;   // temperature / humidity / pressure / longwave

  // ---- cold per-cell constants ----------------------------------------------------------------
  Cell c;
  c.cell = cell; c.D = D; c.ts = ts; c.tpd = tpd; c.nk = 2 * D + 4; c.ld = f.ld;
  const int nk = 2 * D + 4;
  const double lon = wrap_lon(f.lon[cell]);                            // disaggregate.py:67
  const int off_min = p.utc_offset ? (int)((lon / kDegPerRev) * kMinPerDay) : 0;  // :349,:376,:611
  const int off_tiny = p.utc_offset ? (int)((lon / kDegPerRev) * kTiny) : 0;      // :655
  c.off_min = off_min;
  c.utc_min = p.utc_offset ? (double)(int)(((lon - 0.0) / kDegPerRev) * kMinPerDay) : 0.0;  // :186
  c.frac = p.tmax_daylength_fraction;
  const int row = f.lat_row[cell];
  c.rise = a.t.rise_min + (size_t)row * kRows;
  c.set = a.t.set_min + (size_t)row * kRows;
  c.doy = f.doy; c.sdoy = s_doy; c.win0 = win0; c.nwin = nwin;
  c.t_min = f.t_min; c.t_max = f.t_max; c.st_t_min = f.st_t_min; c.st_t_max = f.st_t_max;
  c.t_end = f.t_end;
  const double *const dayl = a.t.daylength + (size_t)row * kRows;
  const double *const qrow = a.t.q + (size_t)row * 365 * kQLd;
  const long ld = f.ld;
  const int T = D * tpd;
  const int C = 2 * ts;                                     // tiny steps per output step (:660)
  const double inv_sw_den = 1.0 / (3600.0 * ((double)ts / 60.0));   // :645
  const double inv_ts = 1.0 / (double)ts;
  const int i_lo = d_lo * tpd, i_hi = d_hi * tpd;
  auto month_at = [&](int d) {
    const unsigned r = (unsigned)(d - win0);
    return r < (unsigned)nwin ? s_month[r] : f.month[d];
  };
  // day records (disagg_prepare_kernel): field fld of padded day e
  const size_t plane = (size_t)a.n_rec * (size_t)ld;
  const double *const rec0 = a.rec + cell - (long)a.e0 * ld;
  auto rec = [&](int fld, int e) {
#ifdef MSB_BOUNDS   // debug build (python -m metsim_b200.build --bounds): every record access checked
    if ((unsigned)(e - a.e0) >= (unsigned)a.n_rec) { a.status[2] = 1; e = a.e0; }
#endif
    return rec0[(size_t)fld * plane + (long)e * ld];
  };
  auto has_rec = [&](int e) { return (unsigned)(e - a.e0) < (unsigned)a.n_rec; };
  auto rkx = [&](int j) { return rec((j & 1) ? kR_XMAX : kR_XMIN, (j >> 1) - 1); };
  auto rky = [&](int j) {      // knot value: the input column, or the padded end (disaggregate.py:242-257)
    const int e = (j >> 1) - 1;
    if ((unsigned)e < (unsigned)D) return ((j & 1) ? f.t_max : f.t_min)[(size_t)e * ld + cell];
    return knot_y(c, j);
  };
  auto rvy = [&](int d) { return a.vp_d[(size_t)d * ld + cell] / 1000.0; };   // kPa, :480
  auto rad_of = [&](int d) {   // tmp_rad of day d, disaggregate.py:645
    if (has_rec(d)) return rec(kR_RAD, d);
    const double dl = a.sw_full_day ? kSecPerDay : dayl[doy_at(c, d) - 1];
    return (a.sw_d[(size_t)d * ld + cell] * dl) * inv_sw_den;                        // wrapped day
  };
  auto prec_of = [&](int d, PrecDay &o) {
    if (has_rec(d)) {
      unpack_prec(rec(kR_PPACK, d), o);
      o.hsp = rec(kR_HSP, d); o.hsq = rec(kR_HSQ, d); o.f = rec(kR_PF, d);
    } else {
      prec_day_setup(p, f, cell, d, month_at(d), o);                                  // wrapped day
    }
  };
  auto q_of = [&](int y) {     // table row y rolled over the table's own 366 rows (:656)
    y = y < 0 ? y + kRows : (y >= kRows ? y - kRows : y);
    return qrow + (size_t)min(y, 364) * kQLd;               // row 365 is a copy of row 364
  };

  // ---- hot state ------------------------------------------------------------------------------
  const double elev = f.elev[cell];
  const double pr_num = -(elev * kGStd) / kRDry;            // pressure :401-402
  const double pr_corr = kKelvin + 0.5 * elev * -p.lapse_rate;
  Cubic cub;                                                 // temperature segment
  cub.x0 = cub.x1 = cub.c0 = cub.c1 = cub.c2 = cub.c3 = 0.0;
  double xvl = 0.0, yvl = 0.0, vslope = 0.0;                 // vapour-pressure segment
  bool v_interior = false;                                   // clamp to saturation inside the knots
  double tk_cur = 0.0, wd_cur = 0.0;                         // minute stream: tskc, wind
  double tk_a = 0.0, tk_b = 0.0;                             // STEFAN_B * tskc, STEFAN_B * (1 - tskc)
  OutT wd_o = 0, tk_o = 0;                                   // wind / tskc as they are stored
  const HotK K = load_hot();
  PrecDay pd;                                                // minute stream: precipitation
  pd.p0 = pd.nps = pd.q0 = pd.nqs = 0; pd.pth = kNever; pd.hsp = pd.hsq = pd.f = 0.0;
  double s_prev = 0.0, p_carry = 0.0;
  double pr = 0.0;                 // precipitation of the previous step (what a NaN day repeats)
  double p_base = 0.0;             // what every step starts from: 0, or the fill value on a NaN day
  bool nan_cur = false;            // the current source day is a 0/0 day (disaggregate.py:327)
  double rd_cur = 0.0, q_prev = 0.0, q_pref = 0.0, q_pref2 = 0.0, sw_carry = 0.0;   // tiny-step stream
  const double *q_cur = qrow;

  // ---- cold state: stream positions and the event schedule ------------------------------------
  // Both streams start "one day early, one day too far": the roll event of the first step then
  // installs the true first source day through the same code as every later roll.
  auto wrap = [D](int d) { return d < 0 ? d + D : (d >= D ? d - D : d); };
  const int qm = off_min < 0 ? -1 : 0;          // floor(off_min / 1440), |off_min| <= 720
  int md = wrap(wrap(d_lo + qm) - 1);
  int sm = off_min - qm * kMinPerDay + kMinPerDay;
  const int qs = off_tiny < 0 ? -1 : 0;
  int sd = wrap(wrap(d_lo + qs) - 1);
  int ss = off_tiny - qs * kTiny + kTiny;
  const double *q_next = q_cur + ss + 3 * C;   // table entry the hot loop requests next (it only
                                               // advances this pointer; ss is recovered from it)
  int kc = max(0, 2 * (d_lo + 1) - 4) - 1;      // PCHIP interval: the first event walks right from kc + 1
  int vd = 0;                                   // upper knot of the vapour-pressure segment (event: from vd + 1)
  bool need_fixup = false;                      // leading NaN run: needs the back-fill pass
  int e_cub = thermo ? i_lo : kNever, e_day = want_sw ? i_lo : kNever;
  int e_mroll = i_lo, e_mfol = kNever;
  int e_troll = want_sw ? i_lo : kNever;
  int e_noon = kNever;
  if (want_sw) {
    const int s0 = ss - kTiny;                  // first window [s0, s0 + C)
    e_noon = i_lo + (s0 <= kNoon ? (kNoon - s0) / C : (kNoon + kTiny - s0) / C);
  }

  // vapour pressure interpolant: knots (t_t_min[d], vp_daily[d]/1000), NaN outside, the NaN ends
  // filled by pandas bfill / ffill (disaggregate.py:99-101)
  int e_vp = kNever;
  if (thermo) {
    const double xv_first = vp_knot_x(c, 0), xv_last = vp_knot_x(c, D - 1);
    const int i_head = max(first_step_ge(xv_first, ts, inv_ts), 0);   // first step >= x_v[0]
    const int i_tail = first_step_gt(xv_last, ts, inv_ts) - 1;        // last step <= x_v[D-1]
    const bool all_nan = (i_head > i_tail) || i_head >= T || i_tail < 0;
    double vp_tail = 0.0;
    if (all_nan) {
      yvl = nan(""); e_vp = kNever;
    } else {
      if (i_hi - 1 > i_tail) vp_tail = vp_endpoint(c, a.vp_d, i_tail, false, etab, a.status);
      if (i_lo < i_head) {             // head: constant until the first in-range step
        yvl = vp_endpoint(c, a.vp_d, i_head, true, etab, a.status);
        e_vp = i_head;
      } else if (i_lo > i_tail) {      // tail
        yvl = vp_tail; e_vp = kNever;
      } else {                         // interior: the handler installs the segment at i_lo
        const double x0 = (double)ts * (double)i_lo;
        vd = min(max(d_lo, 1), D - 1);
        while (vd > 1 && !(vp_knot_x(c, vd - 1) < x0)) --vd;
        --vd;                          // the event handler starts its search at vd + 1
        e_vp = i_lo;
      }
    }
    slot[kS_VPTAIL][tid] = vp_tail;
  }

  int i_event = i_lo;
  double x = (double)ts * (double)i_lo;
  const double dx = (double)ts;
  unsigned ooff = (unsigned)cell;                           // element offset inside the tile
  const unsigned ostride = (unsigned)a.out.ld;
  const size_t obase = (size_t)(i_lo - a.out.day0 * tpd) * (size_t)a.out.ld;
  OutT *const o_temp = reinterpret_cast<OutT *>(a.out.var[MSB_V_TEMP]) + obase;
  OutT *const o_prec = reinterpret_cast<OutT *>(a.out.var[MSB_V_PREC]) + obase;
  OutT *const o_sw = reinterpret_cast<OutT *>(a.out.var[MSB_V_SHORTWAVE]) + obase;
  OutT *const o_lw = reinterpret_cast<OutT *>(a.out.var[MSB_V_LONGWAVE]) + obase;
  OutT *const o_vp = reinterpret_cast<OutT *>(a.out.var[MSB_V_VAPOR_PRESSURE]) + obase;
  OutT *const o_rh = reinterpret_cast<OutT *>(a.out.var[MSB_V_REL_HUMID]) + obase;
  OutT *const o_ap = reinterpret_cast<OutT *>(a.out.var[MSB_V_AIR_PRESSURE]) + obase;
  OutT *const o_sh = reinterpret_cast<OutT *>(a.out.var[MSB_V_SPEC_HUMID]) + obase;
  OutT *const o_tk = reinterpret_cast<OutT *>(a.out.var[MSB_V_TSKC]) + obase;
  OutT *const o_wd = reinterpret_cast<OutT *>(a.out.var[MSB_V_WIND]) + obase;
  auto conv = [&](int v, double val) {
    if (UNITS) val = fma(val, a.out.scale[v], a.out.offset[v]);
    return (OutT)val;
  };
  auto put = [&](OutT *ptr, int v, double val) {
    __stcs(ptr + ooff, conv(v, val));   // streaming store: outputs are never re-read by this kernel
  };

  // Lanes of a warp stay in lockstep on the step index (coalesced stores), so the event-free run
  // that follows the handler ends at the earliest event of ANY lane of the warp.
  const unsigned lanes = __activemask();
  int i = i_lo;
  while (i < i_hi) {
    if (want_sw) ss = (int)(q_next - q_cur) - 3 * C;   // the hot loop only advances the pointer
    if (i == i_event) {
      // ================= event handler (cold): pure state updates =============================
      if (i == e_mfol) {   // step after a straddling roll: tskc / wind settle on the new day
        tk_cur = slot[kS_TKN][tid];
        wd_cur = slot[kS_WDN][tid];
        tk_a = kStefanB * tk_cur; tk_b = kStefanB * (1.0 - tk_cur);
        wd_o = conv(MSB_V_WIND, wd_cur); tk_o = conv(MSB_V_TSKC, tk_cur);
        e_mfol = kNever;
      }
      if (i == e_mroll) {  // minute stream reaches the end of its source day (:347-351, :376, :611)
        // at most two passes: a window that STARTS in the next day (clean switch; also how the
        // first source day of a tile is installed) and / or one that straddles the boundary
        while (sm + ts > kMinPerDay) {
          const int mdn = (md + 1 == D) ? 0 : md + 1;
          const size_t on = (size_t)mdn * ld + cell;
          const double tk_n = a.tskc_d[on];
          const double wd_n = want_wind ? f.wind[on] : 0.0;
          PrecDay pn = pd;
          bool nan_new = false;
          double fill = 0.0;               // pandas ffill: a NaN step repeats the last valid value
          if (want_prec) {
            prec_of(mdn, pn);
            nan_new = !(pn.f == pn.f);
            if (nan_cur) fill = p_base;                      // already inside a NaN run
            else if (nan_new) {
              if (i > i_lo) fill = pr;                       // the previous step of this tile
              else {                                         // the tile starts at a NaN day: look back
                fill = prec_lookback(p, f, c, i);
                if (!(fill == fill)) need_fixup = true;      // nothing valid before: back-fill later
              }
            }
          }
          int pos;                         // minute of the new day where the cumulative sum restarts
          if (sm >= kMinPerDay) {          // clean switch
            sm -= kMinPerDay;
            pos = sm;
            tk_cur = tk_n; wd_cur = wd_n;
            p_carry = nan_new ? fill : 0.0;
          } else {                         // straddle: n1 minutes of the old day, n2 of the new one
            const int n2 = sm + ts - kMinPerDay, n1 = ts - n2;
            tk_cur = ((double)n1 * tk_cur + (double)n2 * tk_n) * inv_ts;
            wd_cur = ((double)n1 * wd_cur + (double)n2 * wd_n) * inv_ts;
            slot[kS_TKN][tid] = tk_n;
            slot[kS_WDN][tid] = wd_n;
            e_mfol = i + 1;
            if (want_prec)   // old-day part + new-day part (NaN if either day is); the hot path
                             // adds f_new * 0 to it
              p_carry = (nan_cur || nan_new)
                            ? fill
                            : pd.f * (prec_cum(pd, kMinPerDay) - s_prev) + pn.f * prec_cum(pn, n2);
            pos = n2;
            sm = n2 - ts;                  // the hot path advances it to n2
          }
          md = mdn;
          if (want_prec) {
            pd = pn;
            if (nan_new) pd.f = 0.0;       // every step of a NaN day is p_base = the fill value
            p_base = nan_new ? fill : 0.0;
            nan_cur = nan_new;
            s_prev = prec_cum(pd, pos);
          }
        }
        tk_a = kStefanB * tk_cur; tk_b = kStefanB * (1.0 - tk_cur);
        wd_o = conv(MSB_V_WIND, wd_cur); tk_o = conv(MSB_V_TSKC, tk_cur);
        e_mroll = i + 1 + (kMinPerDay - (sm + ts)) / ts;
      }
      // tiny-step stream reaches the end of its source day / table row (:656-678); the clean
      // switch comes before, the straddle after the day-start row selection (reference order)
      if (i == e_troll && ss >= kTiny) {
        sd = (sd + 1 == D) ? 0 : sd + 1;
        rd_cur = rad_of(sd);
        // row of the "hi" part of the windows of the output day that holds the previous step
        q_cur = q_of(doy_at(c, max(i - 1, 0) / tpd) - 1 + qs + 1);
        ss -= kTiny;
        q_prev = q_cur[ss];
        sw_carry = 0.0;
        slot[kS_ROWTOT][tid] = q_cur[kTiny + 1];
      }
      if (i == e_day) {    // output day starts: table row of the "lo" part of its windows (:656)
        const int d = i / tpd;
        const double *qn = q_of(doy_at(c, d) - 1 + qs);
        if (qn != q_cur) {               // only across a year boundary (or at the tile start)
          q_cur = qn;
          q_prev = q_cur[ss];
          slot[kS_ROWTOT][tid] = q_cur[kTiny + 1];
        }
        // the roll inside a day installs the row of the next day's "lo" part unless the day of
        // year jumps: only those day starts need an event
        e_day = kNever;
        for (int dd = d; dd + 1 < d_hi; ++dd)
          if (doy_at(c, dd + 1) != doy_at(c, dd) + 1) { e_day = (dd + 1) * tpd; break; }
      }
      if (i == e_troll) {
        if (ss + C > kTiny) {            // straddle: [ss, 2880) of the old row + [0, k2) of the new
          const int k2 = ss + C - kTiny;
          sd = (sd + 1 == D) ? 0 : sd + 1;
          const double rd_n = rad_of(sd);
          const double *qn = q_of(doy_at(c, i / tpd) - 1 + qs + 1);
          const double qk = qn[k2];
          sw_carry = fma(rd_n, qk - qn[0], rd_cur * (q_cur[kTiny] - q_prev));  // ss > noon as C <= 720
          q_prev = qk;                   // the hot path then adds rd_new * (q[k2] - q[k2]) = 0
          ss = k2 - C;                   // the hot path advances it to k2
          rd_cur = rd_n;
          q_cur = qn;
          slot[kS_ROWTOT][tid] = qn[kTiny + 1];
        }
        e_troll = i + 1 + (kTiny - (ss + C)) / C;
      }
      if (thermo && i >= e_vp) {     // vapour-pressure knot crossed (searchsorted side='left')
        int vn = min(vd + 1, D - 1);     // usual case: the next knot is the new upper end
        double xl = rec(kR_XMIN, vn - 1), yl = rvy(vn - 1);
        double xh = rec(kR_XMIN, vn), yh = rvy(vn);
        while (vn < D - 1 && xh < x) { ++vn; xl = xh; yl = yh; xh = rec(kR_XMIN, vn); yh = rvy(vn); }
        vd = vn;
        if (xh < x) {      // beyond the last knot: forward fill
          yvl = slot[kS_VPTAIL][tid]; vslope = 0.0; v_interior = false; e_vp = kNever;
        } else {
          xvl = xl; yvl = yl;
          vslope = div_nr(yh - yl, xh - xl);                // scipy _interpolate.py:513-516
          v_interior = true;
          e_vp = first_step_gt(xh, ts, inv_ts);
        }
      }
      if (thermo && i >= e_cub) {    // PCHIP knot crossed (searchsorted side='right')
        int kn = kc + 1;                 // usual case: the next interval; all six loads in flight
        double x0 = rkx(kn), x1 = rkx(kn + 1);
        double y0 = rky(kn);
        double d0 = rec((kn & 1) ? kR_DMAX : kR_DMIN, (kn >> 1) - 1);
        double y1 = rky(kn + 1);
        double d1 = rec((kn & 1) ? kR_DMIN : kR_DMAX, ((kn + 1) >> 1) - 1);
        while (kn < nk - 2 && x >= x1) {   // walk right (tile start; several knots inside one step)
          ++kn;
          x0 = x1; y0 = y1; d0 = d1;
          x1 = rkx(kn + 1);
          y1 = rky(kn + 1);
          d1 = rec((kn & 1) ? kR_DMIN : kR_DMAX, ((kn + 1) >> 1) - 1);
        }
        kc = kn;
        // Hermite -> scipy power basis (CubicHermiteSpline, _cubic.py:158-165)
        const double inv = rcp_nr(x1 - x0);
        const double m = (y1 - y0) * inv;
        const double tt = ((d0 + d1) - 2.0 * m) * inv;
        cub.x0 = x0; cub.x1 = x1;
        cub.c0 = tt * inv;
        cub.c1 = (m - d0) * inv - tt;
        cub.c2 = d0;
        cub.c3 = y0;
        e_cub = (kc >= nk - 2) ? kNever : first_step_ge(x1, ts, inv_ts);   // last interval extrapolates
      }
      if (i == e_noon) {   // this step's window [ss, ss+C) contains tiny step 1440: add the row total
        q_prev -= slot[kS_ROWTOT][tid];
        e_noon += tpd;
      }
#ifdef MSB_BOUNDS
      if (want_sw && (ss + C < 0 || ss + 3 * C > kQLd + MSB_Q_PAD / 2 || q_cur < qrow ||
                      q_cur > qrow + (size_t)364 * kQLd)) asm volatile("trap;");
#endif
      if (want_sw) {       // re-prime the pipelined table loads: the entries of this step and the next
        q_pref = q_cur[ss + C];
        q_pref2 = q_cur[ss + 2 * C];
        q_next = q_cur + ss + 3 * C;
      }
      i_event = min(min(min(e_mfol, e_day), min(e_vp, e_cub)), min(min(e_noon, e_troll), e_mroll));
      i_event = max(i_event, i + 1);
    }
    const int i_stop = min(__reduce_min_sync(lanes, i_event), i_hi);

    // ================= hot path: event-free run [i, i_stop) ===================================
    const unsigned ooff_stop = ooff + (unsigned)(i_stop - i) * ostride;
    i = i_stop;
#pragma unroll 1
    for (; ooff != ooff_stop; x += dx, ooff += ostride) {
      // -- tiny-step stream: this step's table entry was requested one step ago; request the next
      //    one now and pull the same position of tomorrow's row into L2 --
      const double qn = q_pref;
      if (want_sw) {     // two loads in flight: the entry of the step after next is requested now
        q_pref = q_pref2;
        q_pref2 = load_early(q_next);
        prefetch_l2(q_next + kQLd);
        q_next += C;
        prefetch_l1(q_next + C);
      }

      double temp = 0.0, vp = 0.0, rh = 0.0, ap = 0.0, lw = 0.0;
      if (thermo) {
        // -- temperature (PCHIP), saturation vapour pressure in kPa (physics.py:124-152) --
        temp = cubic_eval(cub, x);
        const double es0 = exp_hot(fma(K.ka, rcp_nr(K.k237 + temp), K.kb), etab, K);
        const double corr = fma(fma(K.kc2, temp, K.kc1), temp, 1.0);
        const double es = es0 * (temp < 0.0 ? corr : 1.0);

        // -- vapour pressure (linear, clamped to saturation inside the knots) :480-487 --
        vp = fma(x - xvl, vslope, yvl);
        vp = (v_interior && es < vp) ? es : vp;

        // -- relative humidity :444-446, air pressure :401-403 --
        rh = fmin(100.0 * (vp * rcp_nr(es)), 100.0);   // 100*1000*vp/(svp in Pa)
        ap = exp_hot(fma(pr_num, rcp_nr(pr_corr + temp), K.ln_p), etab, K);

        // -- longwave :491-590 (PRATA + CLOUD_DEARDORFF inline; tk_a / tk_b carry STEFAN_B) --
        if (FAST_LW) {
          if (want(MSB_V_LONGWAVE)) {
            const double at = temp + K.kelvin;
            const double xx = (465.0 * vp) * rcp_nr(at);             // 46.5 * (10 vp) / T
            const double ee = exp_hot(-sqrt_hot(fma(3., xx, K.k12), K), etab, K);
            const double e1 = fma(-(1.0 + xx), ee, 1.0);
            const double t2 = at * at;
            lw = fma(tk_b, e1, tk_a) * (t2 * t2);
          }
        } else if (want(MSB_V_LONGWAVE)) {
          lw = longwave(p, temp, vp, tk_cur);
        }
      }

      // -- minute stream: precipitation :265-352 (tskc / wind are constant between events) --
      sm += ts;
      if (want_prec) {
        const double s1 = prec_cum(pd, sm);
        pr = fma(pd.f, s1 - s_prev, p_carry);
        s_prev = s1;
        p_carry = p_base;
      }

      // -- shortwave :618-680 --
      double sw = 0.0;
      if (want_sw) {
        sw = fma(rd_cur, qn - q_prev, sw_carry);
        q_prev = qn;
        sw_carry = 0.0;
      }

      if (want(MSB_V_TEMP)) put(o_temp, MSB_V_TEMP, temp);
      if (want(MSB_V_VAPOR_PRESSURE)) put(o_vp, MSB_V_VAPOR_PRESSURE, vp);
      if (want(MSB_V_REL_HUMID)) put(o_rh, MSB_V_REL_HUMID, rh);
      if (want(MSB_V_AIR_PRESSURE)) put(o_ap, MSB_V_AIR_PRESSURE, ap);
      if (want(MSB_V_TSKC)) __stcs(o_tk + ooff, tk_o);
      if (want_wind) __stcs(o_wd + ooff, wd_o);
      if (want(MSB_V_SPEC_HUMID)) {   // specific humidity :423-424
        const double mix = (kEps * vp) * rcp_nr(ap - vp);
        put(o_sh, MSB_V_SPEC_HUMID, mix * rcp_nr(1 + mix));
      }
      if (want(MSB_V_LONGWAVE)) {
        if (VARS == kVarsAny && a.lw_raw) a.lw_raw[obase + ooff] = lw;   // rescaled by lw_rescale_kernel
        else put(o_lw, MSB_V_LONGWAVE, lw);
      }
      if (want_prec) put(o_prec, MSB_V_PREC, pr);
      if (want_sw) put(o_sw, MSB_V_SHORTWAVE, sw);
    }
  }

  // ---- a NaN run (0/0 triangle days) that reaches back to the start of the record: pandas ffill
  //      found nothing, bfill takes the first valid step after it (disaggregate.py:130) --
  if (want_prec && need_fixup) {
    // every step before this tile is NaN, so the tile opens with the leading NaN run of the
    // record: find the first valid step after it and write its value over the run
    double fillv = nan("");
    int j = i_lo;
    for (; j < T; ++j) {
      fillv = prec_step_generic(p, f, c, j);
      if (fillv == fillv) break;
    }
    ooff = (unsigned)cell;
    for (int jj = i_lo; jj < min(j, i_hi); ++jj, ooff += ostride) put(o_prec, MSB_V_PREC, fillv);
  }
}


// ---- interior days of the fast path: knot-aligned thermodynamic kernel ---------------------------
// Days [2, D-2) of a record are "interior": every PCHIP / vapour-pressure knot a step needs exists
// in the day records, the vapour-pressure interpolant has no NaN ends there, and the rolled minute
// stream does not wrap.  For those days the work is cut at the KNOTS instead of at midnight: one
// thread handles the steps between the t_Tmin knot of day e and the t_Tmin knot of day e+1 -- exactly
// two PCHIP intervals and one vapour-pressure segment -- so what the event handler of
// disagg_kernel does under divergence (segment set-up, next-event search) happens here once per
// (cell, e) with the warp converged and all loads independent.  The step loop runs over the
// warp's union of step ranges with the lane's own range as a predicate, so stores stay full
// 128-byte rows.  The first / last two days of a record go through disagg_kernel.
enum { kT_C0 = 0, kT_C1, kT_C2, kT_C3, kT_X0, kT_TKA, kT_TKB, kT_WDA, kT_WDB,
       kT_KX0, kT_KY0, kT_KD0, kT_KX1, kT_KY1, kT_KM01, kT_COUNT };

// 2^(j/1024), j < 1024: the exp table of the knot kernel, built once per device (exp_tables_ready)
constexpr int kExpBig = 1024;
constexpr int kExpBigShift = 10;   // 20 - log2(kExpBig): where the index bits sit in the high word
__device__ double g_exp2_big[kExpBig];
__global__ void exp_big_init_kernel() {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < kExpBig) {
    // 2^(j/1024) with (j << 10) taken off the high word: exp_big() adds (k << 10) as a whole, whose
    // low 20 bits are exactly j << 10 and whose upper bits are the binary exponent k >> 10
    const double s = exp2((double)j / (double)kExpBig);
    g_exp2_big[j] = __hiloint2double(__double2hiint(s) - (j << kExpBigShift), __double2loint(s));
  }
}
__constant__ double c_knot[10] = {
    237.3,                          // svp denominator offset, physics.py:149
    -17.269 * 237.3,                // 17.269*t/(237.3+t) = 17.269 - 17.269*237.3/(237.3+t)
    17.269 - 0.49301845011612494,   // + ln(0.61078): the factor folded into the exponent
    .000042, .00972,                // ice correction, physics.py:150-151
    1477.3197218702985,             // 1024/ln2
    0.0006769015435155716,          // ln2/1024
    4.618333172514372,              // ln(P_STD/1000): air pressure in kPa, disaggregate.py:401-403
    273.15,                         // KELVIN
    1.2                             // PRATA, disaggregate.py:557-559
};
struct KnotK { double k237, ka, kb, kc2, kc1, inv_l, l, ln_p, kelvin, k12; };
__device__ __forceinline__ KnotK load_knot() {
  KnotK k;
  k.k237 = c_knot[0]; k.ka = c_knot[1]; k.kb = c_knot[2]; k.kc2 = c_knot[3]; k.kc1 = c_knot[4];
  k.inv_l = c_knot[5]; k.l = c_knot[6]; k.ln_p = c_knot[7]; k.kelvin = c_knot[8]; k.k12 = c_knot[9];
  return k;
}
// The knot kernel's outputs are only written, never fed back, so its fp64 sequences are cut to
// what the 1e-9 contract needs with two orders of margin:
//   1/x      MUFU.RCP64H seed (2^-20) + one second-order step        rel. error < 1e-12
//   sqrt(x)  MUFU.RSQ64H seed + one second-order step                 rel. error < 4e-13
//   exp(x)   1024-entry 2^(j/1024) table + quadratic on |r| < 3.4e-4  rel. error < 7e-12
__device__ __forceinline__ double rcp_2(double x) {
  const double y = rcp_seed(x);
  const double e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ double sqrt_2(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double xy = x * y;
  const double e = fma(-xy, y, 1.0);
  return fma(xy, 0.5 * e, xy);
}
__device__ __forceinline__ double exp_big(double x, const double *tab, const KnotK &K) {
  const double kMagic = 6755399441055744.0;       // 1.5 * 2^52
  // 1024/ln2 truncated to its high word (an instruction immediate): k may differ by one from
  // rint(x * 1024/ln2) at a rounding boundary, r = x - k ln2/1024 stays consistent with it
  double kd = fma(x, __hiloint2double(0x40971547, 0), kMagic);
  const int k = __double2loint(kd);
  kd -= kMagic;
  const double r = fma(-kd, K.l, x);
  const double q = fma(r * 0.5, r, r);            // exp(r) - 1 = r + r^2/2 (+ r^3/6 < 7e-12)
  double s = tab[k & (kExpBig - 1)];
  s = __hiloint2double(__double2hiint(s) + (k << kExpBigShift), __double2loint(s));
  return fma(q, s, s);
}

template <typename OutT, int MINB>
__global__ void __launch_bounds__(kDisaggThreads, MINB)
thermo_knot_kernel(const __grid_constant__ DisaggArgs a) {
  __shared__ double etab[kExpBig];
  __shared__ double slot[kT_COUNT][kDisaggThreads];
  __shared__ int s_doy[kDayWin];
  const int e_lo = a.k.e0 + blockIdx.y * a.k.tile;
  const int e_hi = min(e_lo + a.k.tile, a.k.e1);
  const int win0 = max(e_lo - 2, 0), nwin = min(min(e_hi + 3, a.f.n_days) - win0, kDayWin);
  for (int j = threadIdx.x; j < kExpBig; j += blockDim.x) etab[j] = g_exp2_big[j];
  for (int j = threadIdx.x; j < nwin; j += blockDim.x) s_doy[j] = a.f.doy[win0 + j];
  __syncthreads();
  const int tid = threadIdx.x;
  const int cell = blockIdx.x * blockDim.x + tid;
  if (cell >= a.f.n_cells || e_lo >= e_hi) return;
  const msb_params &p = a.p;
  const msb_forcing &f = a.f;
  const int ts = p.time_step;
  const long ld = f.ld;
  const double inv_ts = 1.0 / (double)ts;
  const double lon = wrap_lon(f.lon[cell]);                            // disaggregate.py:67
  const int off_min = p.utc_offset ? (int)((lon / kDegPerRev) * kMinPerDay) : 0;  // :611
  const double elev = f.elev[cell];
  const double pr_num = -(elev * kGStd) / kRDry;            // pressure :401-402
  const double pr_corr = kKelvin + 0.5 * elev * -p.lapse_rate;
  // the knots of the padded PCHIP sequence (disaggregate.py:236-257) straight from the inputs:
  // what disagg_prepare_kernel records per (cell, day) is derived here once per (cell, e) with the
  // warp converged, so the interior days need no record traffic
  Cell c;
  c.cell = cell; c.D = f.n_days; c.ts = ts; c.tpd = a.tpd; c.nk = 2 * f.n_days + 4; c.ld = ld;
  c.off_min = off_min;
  c.utc_min = p.utc_offset ? (double)(int)(((lon - 0.0) / kDegPerRev) * kMinPerDay) : 0.0;  // :186
  c.frac = p.tmax_daylength_fraction;
  {
    const int row = f.lat_row[cell];
    c.rise = a.t.rise_min + (size_t)row * kRows;
    c.set = a.t.set_min + (size_t)row * kRows;
  }
  c.doy = f.doy; c.sdoy = s_doy; c.win0 = win0; c.nwin = nwin;
  c.t_min = f.t_min; c.t_max = f.t_max; c.st_t_min = f.st_t_min; c.st_t_max = f.st_t_max;
  c.t_end = f.t_end;
  const KnotK K = load_knot();
  const unsigned lanes = __activemask();

  const size_t ostride = (size_t)a.out.ld;
  const long i_base = (long)a.out.day0 * a.tpd;
  OutT *const o_temp = reinterpret_cast<OutT *>(a.out.var[MSB_V_TEMP]) + cell;
  OutT *const o_lw = reinterpret_cast<OutT *>(a.out.var[MSB_V_LONGWAVE]) + cell;
  OutT *const o_vp = reinterpret_cast<OutT *>(a.out.var[MSB_V_VAPOR_PRESSURE]) + cell;
  OutT *const o_rh = reinterpret_cast<OutT *>(a.out.var[MSB_V_REL_HUMID]) + cell;
  OutT *const o_ap = reinterpret_cast<OutT *>(a.out.var[MSB_V_AIR_PRESSURE]) + cell;
  OutT *const o_wd = reinterpret_cast<OutT *>(a.out.var[MSB_V_WIND]) + cell;

  // tskc and wind: repeated per minute, rolled by off_min, block mean over ts minutes
  // (:593-615, :355-380); wind rides along here because it shares this stream's roll exactly
  int i_tk = kNever, tk_day = -4;
  bool tk_started = false;
  double tk_a = 0.0, tk_b = 0.0;
  OutT wd_o = 0;
  const double dx = (double)ts;

  // knots t_Tmax(e-1), t_Tmin(e), t_Tmax(e) of the first e: times, values, slopes and scipy's
  // derivative at t_Tmin(e) (_cubic.py:320-333); they roll forward through shared-memory slots
  {
    const int j0 = 2 * (e_lo + 1);
    const double xa = knot_x(c, j0 - 1), xb = knot_x(c, j0), xc = knot_x(c, j0 + 1);
    const double ya = knot_y(c, j0 - 1), yb = knot_y(c, j0), yc = knot_y(c, j0 + 1);
    const double h0 = __dadd_rn(xb, -xa), h1 = __dadd_rn(xc, -xb);
    const double m0 = div_nr(__dadd_rn(yb, -ya), h0), m1 = div_nr(__dadd_rn(yc, -yb), h1);
    if (!(h0 > 0.0) || !(h1 > 0.0)) a.status[2] = 1;      // scipy: `x` must be strictly increasing
    slot[kT_KX0][tid] = xb; slot[kT_KY0][tid] = yb; slot[kT_KD0][tid] = pchip_mid(h0, h1, m0, m1);
    slot[kT_KX1][tid] = xc; slot[kT_KY1][tid] = yc; slot[kT_KM01][tid] = m1;
  }

  for (int e = e_lo; e < e_hi; ++e) {
    // knots at t_Tmin / t_Tmax of day e and t_Tmin of day e+1 (disaggregate.py:198-200) with
    // scipy's derivatives, and the daily vapour pressure in kPa at the two t_Tmin knots (:480)
    const size_t o0 = (size_t)e * ld + cell, o1 = o0 + ld;
    const int j2 = 2 * (e + 2);
    const double x2 = knot_x(c, j2), x3 = knot_x(c, j2 + 1);
    const double y2 = knot_y(c, j2), y3 = knot_y(c, j2 + 1);
    const double x0 = slot[kT_KX0][tid], y0 = slot[kT_KY0][tid], d0 = slot[kT_KD0][tid];
    const double x1 = slot[kT_KX1][tid], y1 = slot[kT_KY1][tid], m01 = slot[kT_KM01][tid];
    const double h01 = __dadd_rn(x1, -x0), h12 = __dadd_rn(x2, -x1), h23 = __dadd_rn(x3, -x2);
    const double m12 = div_nr(__dadd_rn(y2, -y1), h12), m23 = div_nr(__dadd_rn(y3, -y2), h23);
    if (!(h12 > 0.0) || !(h23 > 0.0)) a.status[2] = 1;
    const double d1 = pchip_mid(h01, h12, m01, m12), d2 = pchip_mid(h12, h23, m12, m23);
    slot[kT_KX0][tid] = x2; slot[kT_KY0][tid] = y2; slot[kT_KD0][tid] = d2;
    slot[kT_KX1][tid] = x3; slot[kT_KY1][tid] = y3; slot[kT_KM01][tid] = m23;
    const double v0 = a.vp_d[o0] / 1000.0, v1 = a.vp_d[o1] / 1000.0;
    const int ia = first_step_ge(x0, ts, inv_ts), ib = first_step_ge(x2, ts, inv_ts);
    int isw = first_step_ge(x1, ts, inv_ts);
    // Hermite -> scipy power basis (CubicHermiteSpline, _cubic.py:158-165), both intervals
    Cubic cub;
    {
      const double inv = rcp_nr(x1 - x0);
      const double m = (y1 - y0) * inv;
      const double tt = ((d0 + d1) - 2.0 * m) * inv;
      cub.x0 = x0; cub.x1 = x1;
      cub.c0 = tt * inv;
      cub.c1 = (m - d0) * inv - tt;
      cub.c2 = d0;
      cub.c3 = y0;
    }
    {
      const double inv = rcp_nr(x2 - x1);
      const double m = (y2 - y1) * inv;
      const double tt = ((d1 + d2) - 2.0 * m) * inv;
      slot[kT_X0][tid] = x1;
      slot[kT_C0][tid] = tt * inv;
      slot[kT_C1][tid] = (m - d1) * inv - tt;
      slot[kT_C2][tid] = d1;
      slot[kT_C3][tid] = y1;
    }
    const double xvl = x0, yvl = v0;
    const double vslope = div_nr(v1 - v0, x2 - x0);                  // scipy _interpolate.py:513-516
    const int ja = max(ia, a.k.i_lo), jb = min(ib, a.k.i_hi);
    const bool any = ja < jb;
    if (any && isw <= ja) {          // the range starts beyond the t_Tmax knot
      cub.x0 = slot[kT_X0][tid]; cub.c0 = slot[kT_C0][tid]; cub.c1 = slot[kT_C1][tid];
      cub.c2 = slot[kT_C2][tid]; cub.c3 = slot[kT_C3][tid];
      isw = kNever;
    }
    if (any && !tk_started) { i_tk = ja; tk_started = true; }
    int i_next = min(isw, i_tk);
    const int lo_w = __reduce_min_sync(lanes, any ? ja : kNever);
    const int hi_w = __reduce_max_sync(lanes, any ? jb : -kNever);
    double x = (double)ts * (double)lo_w;
    size_t ooff = (size_t)((long)lo_w - i_base) * ostride;
#pragma unroll 1
    for (int i = lo_w; i < hi_w; ++i, x += dx, ooff += ostride) {
      if (i < ja || i >= jb) continue;
      if (i == i_next) {
        if (i == isw) {              // t_Tmax knot crossed (searchsorted side='right')
          cub.x0 = slot[kT_X0][tid]; cub.c0 = slot[kT_C0][tid]; cub.c1 = slot[kT_C1][tid];
          cub.c2 = slot[kT_C2][tid]; cub.c3 = slot[kT_C3][tid];
          isw = kNever;
        }
        if (i == i_tk) {             // source day of the minute stream rolls (:611-614)
          const int m = i * ts + off_min;         // interior days: 0 < m < (D-1)*1440, no wrap
          const int dA = m / kMinPerDay, r = m - dA * kMinPerDay;
          if (dA != tk_day) {
            const size_t oA = (size_t)dA * ld + cell;
            const bool next = dA == tk_day + 1;
            slot[kT_TKA][tid] = next ? slot[kT_TKB][tid] : a.tskc_d[oA];
            slot[kT_WDA][tid] = next ? slot[kT_WDB][tid] : f.wind[oA];
            slot[kT_TKB][tid] = a.tskc_d[oA + ld];   // used a day later
            slot[kT_WDB][tid] = f.wind[oA + ld];
            tk_day = dA;
          }
          const int n2 = r + ts - kMinPerDay;
          double tk = slot[kT_TKA][tid], wd = slot[kT_WDA][tid];
          if (n2 > 0) {              // window straddles the day boundary
            const double w1 = (double)(ts - n2), w2 = (double)n2;
            tk = (w1 * tk + w2 * slot[kT_TKB][tid]) * inv_ts;
            wd = (w1 * wd + w2 * slot[kT_WDB][tid]) * inv_ts;
            i_tk = i + 1;
          } else {
            i_tk = i + (kMinPerDay - r) / ts;
          }
          tk_a = kStefanB * tk; tk_b = kStefanB * (1.0 - tk);
          wd_o = (OutT)wd;
        }
        i_next = min(isw, i_tk);
      }
      // -- temperature (PCHIP), saturation vapour pressure in kPa (physics.py:124-152) --
      const double temp = cubic_eval(cub, x);
      const double es0 = exp_big(fma(K.ka, rcp_2(K.k237 + temp), K.kb), etab, K);
      const double corr = fma(fma(K.kc2, temp, K.kc1), temp, 1.0);
      const double es = es0 * (temp < 0.0 ? corr : 1.0);
      // -- vapour pressure (linear, clamped to saturation) :480-487 --
      double vp = fma(x - xvl, vslope, yvl);
      vp = es < vp ? es : vp;
      // -- relative humidity :444-446, air pressure :401-403 --
      const double rh_raw = 100.0 * (vp * rcp_2(es));
      const double rh = rh_raw > 100.0 ? 100.0 : rh_raw;      // np.minimum; vp, es are finite here
      const double ap = exp_big(fma(pr_num, rcp_2(pr_corr + temp), K.ln_p), etab, K);
      // -- longwave :491-590 (PRATA + CLOUD_DEARDORFF; tk_a / tk_b carry STEFAN_B) --
      const double at = temp + K.kelvin;
      const double xx = (465.0 * vp) * rcp_2(at);              // 46.5 * (10 vp) / T
      const double ee = exp_big(-sqrt_2(fma(3., xx, K.k12)), etab, K);
      const double e1 = fma(-(1.0 + xx), ee, 1.0);
      const double t2 = at * at;
      const double lw = fma(tk_b, e1, tk_a) * (t2 * t2);
      __stcs(o_temp + ooff, (OutT)temp);
      __stcs(o_vp + ooff, (OutT)vp);
      __stcs(o_rh + ooff, (OutT)rh);
      __stcs(o_ap + ooff, (OutT)ap);
      __stcs(o_lw + ooff, (OutT)lw);
      __stcs(o_wd + ooff, wd_o);
    }
  }
}

// ---- interior days of the fast path: source-stream kernel with a fixed daily schedule -------------
// Precipitation (minute stream) and shortwave (30 s stream) of the interior days [2, D-2) -- wind
// is written by thermo_knot_kernel, which already follows the minute stream's roll for tskc --
// where neither rolled stream wraps around the record.  Same closed forms as disagg_kernel, but the
// daily schedule is made explicit: for a given cell the roll of the minute stream, the roll of the
// 30 s stream and the noon crossing happen at the SAME step of every output day (they depend on
// the longitude offsets only), so they are four per-cell constants instead of a searched event
// queue, and what a roll needs about the next source day is fetched at the start of the output
// day into per-thread shared-memory slots.  The handlers are short, load-free state updates.
enum { kY_PPACK = 0, kY_HSP, kY_HSQ, kY_PF, kY_RAD, kY_QB0, kY_QBK, kY_QAEND, kY_ROWTOT,
       kY_ROWTOT_CUR, kY_COUNT };

template <typename OutT, int MINB>
__global__ void __launch_bounds__(kDisaggThreads, MINB)
streams_day_kernel(const __grid_constant__ DisaggArgs a) {
  __shared__ double slot[kY_COUNT][kDisaggThreads];
  __shared__ int s_doy[kDayWin], s_month[kDayWin];
  const msb_params &p = a.p;
  const msb_forcing &f = a.f;
  const int D = f.n_days, ts = p.time_step, tpd = a.tpd;
  const int d_lo = a.k.e0 + blockIdx.y * a.k.tile;
  const int d_hi = min(a.k.e1, d_lo + a.k.tile);
  const int win0 = max(d_lo - 4, 0), nwin = min(min(d_hi + 4, D) - win0, kDayWin);
  for (int j = threadIdx.x; j < nwin; j += blockDim.x) {
    s_doy[j] = f.doy[win0 + j];
    s_month[j] = f.month[win0 + j];
  }
  __syncthreads();
  const int tid = threadIdx.x;
  const int cell = blockIdx.x * blockDim.x + tid;
  if (cell >= f.n_cells || d_lo >= d_hi) return;
  const long ld = f.ld;
  const int C = 2 * ts;                                     // tiny steps per output step (:660)
  const double lon = wrap_lon(f.lon[cell]);                            // disaggregate.py:67
  const int off_min = p.utc_offset ? (int)((lon / kDegPerRev) * kMinPerDay) : 0;  // :349,:376
  const int off_tiny = p.utc_offset ? (int)((lon / kDegPerRev) * kTiny) : 0;      // :655
  const double *const qrow = a.t.q + (size_t)f.lat_row[cell] * 365 * kQLd;
  auto q_of = [&](int y) {     // table row y rolled over the table's own 366 rows (:656)
    y = y < 0 ? y + kRows : (y >= kRows ? y - kRows : y);
    return qrow + (size_t)min(y, 364) * kQLd;               // row 365 is a copy of row 364
  };
  // what disagg_prepare_kernel would put into the record of source day d: done here, once per
  // (cell, day) with the warp converged, so the interior days need no record traffic at all
  const double *const dayl = a.t.daylength + (size_t)f.lat_row[cell] * kRows;
  auto rad_of = [&](int d) {   // tmp_rad of day d, disaggregate.py:645
    const double dl = a.sw_full_day ? kSecPerDay : dayl[s_doy[d - win0] - 1];
    return (a.sw_d[(size_t)d * ld + cell] * dl) * a.inv_sw_den;
  };
  auto lookback = [&](int i, bool &none) {   // what pandas ffill repeats at step i (cold)
    Cell c;
    c.cell = cell; c.D = D; c.ts = ts; c.tpd = tpd; c.nk = 2 * D + 4; c.ld = ld;
    c.off_min = off_min; c.utc_min = 0.0; c.frac = 0.0; c.rise = c.set = nullptr;
    c.doy = f.doy; c.sdoy = nullptr; c.win0 = 0; c.nwin = 0;
    c.t_min = f.t_min; c.t_max = f.t_max; c.st_t_min = f.st_t_min; c.st_t_max = f.st_t_max;
    c.t_end = f.t_end;
    const double v = prec_lookback(p, f, c, i);
    if (!(v == v)) none = true;
    return v;
  };

  // ---- the cell's daily schedule (steps of the output day at which something happens) ----------
  // minute stream: output day d reads source day d + mq - 1 up to the roll at step t_s, then
  // source day d + mq; n2m minutes of the straddling window belong to the new day (0 = clean)
  const int qm = off_min < 0 ? -1 : 0;
  const int r0 = off_min - qm * kMinPerDay;                 // [0, 1440)
  const int mq = r0 == 0 ? qm : qm + 1;
  const int t_s = r0 == 0 ? 0 : (kMinPerDay - r0) / ts;
  const int n1m = r0 == 0 ? 0 : kMinPerDay - r0 - t_s * ts;
  const int n2m = n1m > 0 ? ts - n1m : 0;
  // 30 s stream: same for the table row / radiation day
  const int qs = off_tiny < 0 ? -1 : 0;
  const int c0 = off_tiny - qs * kTiny;                     // [0, 2880)
  const int tq = c0 == 0 ? qs : qs + 1;
  const int t_r = c0 == 0 ? 0 : (kTiny - c0) / C;
  const int k2 = c0 == 0 ? 0 : (c0 + (t_r + 1) * C - kTiny) % C;        // 0 = clean switch
  const int t_n = ((c0 <= kNoon ? kNoon : kNoon + kTiny) - c0) / C;     // window that contains noon
  const unsigned ev = (unsigned)t_s | ((unsigned)t_r << 8) | ((unsigned)t_n << 16);
  auto next_event = [ev](int t) {      // first scheduled step after t (255 = none left today)
    unsigned best = 255u;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const unsigned b = (ev >> (8 * k)) & 255u;
      if ((int)b > t && b < best) best = b;
    }
    return (int)best;
  };

  // ---- state at the first step of the chunk -----------------------------------------------------
  const int i_lo = d_lo * tpd;
  int sm = r0 == 0 ? kMinPerDay : r0;                       // start of the next window in its source day
  PrecDay pd;
  {
    const int md0 = d_lo + mq - 1;
    prec_day_setup(p, f, cell, md0, s_month[md0 - win0], pd);
  }
  bool nan_cur = !(pd.f == pd.f), need_fixup = false;
  double p_base = 0.0, pr = 0.0;
  if (nan_cur && r0 != 0) {            // the chunk opens inside a NaN day
    p_base = lookback(i_lo, need_fixup);
    pd.f = 0.0;
  }
  if (r0 == 0) nan_cur = false;        // the roll at the first step decides (and looks back itself)
  double p_carry = p_base;
  double s_prev = prec_cum(pd, sm);
  double rd_cur = rad_of(d_lo + tq - 1);
  double q_prev, q_pref, sw_carry = 0.0;
  const double *q_ptr;                 // table entry the step after next will need
  {
    const double *q0 = q_of(s_doy[d_lo - win0] - 1 + tq - 1);
    const int ss = c0 == 0 ? kTiny : c0;
    q_prev = q0[ss];
    q_pref = q0[ss + C];
    q_ptr = q0 + ss + 2 * C;
    slot[kY_ROWTOT_CUR][tid] = q0[kTiny + 1];
  }

  const size_t ostride = (size_t)a.out.ld;
  size_t ooff = (size_t)((long)i_lo - (long)a.out.day0 * tpd) * ostride;
  OutT *const o_prec = reinterpret_cast<OutT *>(a.out.var[MSB_V_PREC]) + cell;
  OutT *const o_sw = reinterpret_cast<OutT *>(a.out.var[MSB_V_SHORTWAVE]) + cell;
  const unsigned lanes = __activemask();
  const int pf_off = a.k.pf * C;

  for (int d = d_lo; d < d_hi; ++d) {
    // ---- output day d starts (warp converged): fetch what today's rolls need --------------------
    {
      const int y = s_doy[d - win0] - 1;
      const double *qa = q_of(y + tq - 1);
      if (d > d_lo && c0 != 0 && s_doy[d - win0] != s_doy[d - 1 - win0] + 1) {
        // the day of year jumped (year boundary): row of the "lo" part of today's windows (:656)
        q_prev = qa[c0];
        q_pref = qa[c0 + C];
        q_ptr = qa + c0 + 2 * C;
        slot[kY_ROWTOT_CUR][tid] = qa[kTiny + 1];
      }
      const double *qb = q_of(y + tq);
      const int mdn = d + mq, sdn = d + tq;
      const double w5 = rad_of(sdn), w6 = qb[0], w7 = qb[k2], w8 = qa[kTiny], w9 = qb[kTiny + 1];
      PrecDay pn;
      prec_day_setup(p, f, cell, mdn, s_month[mdn - win0], pn);
      const double w0 = pack_prec(pn), w1 = pn.hsp, w2 = pn.hsq, w3 = pn.f;
      prefetch_l2(q_of(y + tq + 1) + (c0 & ~15));   // tomorrow's row, where this cell starts
      slot[kY_PPACK][tid] = w0; slot[kY_HSP][tid] = w1; slot[kY_HSQ][tid] = w2;
      slot[kY_PF][tid] = w3; slot[kY_RAD][tid] = w5;
      slot[kY_QB0][tid] = w6; slot[kY_QBK][tid] = w7; slot[kY_QAEND][tid] = w8;
      slot[kY_ROWTOT][tid] = w9;
    }
    int t_next = next_event(-1);
    int t = 0;
    while (t < tpd) {
      if (t == t_next) {   // ---- scheduled state updates (cold, lanes with an event at this step) ----
        if (t == t_s) {      // minute stream reaches the end of its source day (:347-351, :376)
          PrecDay pn;
          unpack_prec(slot[kY_PPACK][tid], pn);
          pn.hsp = slot[kY_HSP][tid]; pn.hsq = slot[kY_HSQ][tid]; pn.f = slot[kY_PF][tid];
          const bool nan_new = !(pn.f == pn.f);
          double fill = 0.0;               // pandas ffill: a NaN step repeats the last valid value
          if (nan_cur) fill = p_base;      // already inside a NaN run
          else if (nan_new) {
            if (d > d_lo || t > 0) fill = pr;      // the previous step of this chunk
            else fill = lookback(i_lo, need_fixup); // the chunk starts at a NaN day
          }
          int pos;
          if (n1m == 0) {                  // clean switch
            pos = 0;
            sm = 0;
            p_carry = nan_new ? fill : 0.0;
          } else {                         // straddle: n1m minutes of the old day, n2m of the new one
            p_carry = (nan_cur || nan_new)
                          ? fill
                          : pd.f * (prec_cum(pd, kMinPerDay) - s_prev) + pn.f * prec_cum(pn, n2m);
            pos = n2m;
            sm = n2m - ts;                 // the step advances it to n2m
          }
          pd = pn;
          if (nan_new) pd.f = 0.0;         // every step of a NaN day is p_base = the fill value
          p_base = nan_new ? fill : 0.0;
          nan_cur = nan_new;
          s_prev = prec_cum(pd, pos);
        }
        if (t == t_r) {      // 30 s stream reaches the end of its table row / radiation day (:656-678)
          const double rd_n = slot[kY_RAD][tid];
          const double *qb = q_of(s_doy[d - win0] - 1 + tq);
          int ss;
          if (k2 > 0) {                    // straddle: [ss, 2880) of the old row + [0, k2) of the new
            const double qk = slot[kY_QBK][tid];
            sw_carry = fma(rd_n, qk - slot[kY_QB0][tid], rd_cur * (slot[kY_QAEND][tid] - q_prev));
            q_prev = qk;
            ss = k2 - C;                   // the step advances it to k2
          } else {
            q_prev = slot[kY_QB0][tid];
            ss = 0;
          }
          rd_cur = rd_n;
          q_pref = qb[ss + C];
          q_ptr = qb + ss + 2 * C;
          slot[kY_ROWTOT_CUR][tid] = slot[kY_ROWTOT][tid];
        }
        if (t == t_n) q_prev -= slot[kY_ROWTOT_CUR][tid];   // window contains tiny step 1440: + row total
        t_next = next_event(t);
      }
      // ---- event-free run up to the earliest event of any lane of the warp ---------------------
      const int t_stop = min(__reduce_min_sync(lanes, t_next), tpd);
#pragma unroll 2
      for (; t < t_stop; ++t, ooff += ostride) {
        // step: precipitation :265-352, shortwave :618-680, wind :355-380
        const double qn = q_pref;
        q_pref = load_early(q_ptr);
        prefetch_l1(q_ptr + pf_off);
        q_ptr += C;
        sm += ts;
        const double s1 = prec_cum(pd, sm);
        pr = fma(pd.f, s1 - s_prev, p_carry);
        s_prev = s1;
        p_carry = p_base;
        const double sw = fma(rd_cur, qn - q_prev, sw_carry);
        q_prev = qn;
        sw_carry = 0.0;
        __stcs(o_prec + ooff, (OutT)pr);
        __stcs(o_sw + ooff, (OutT)sw);
      }
    }
  }

  // a NaN run that reaches back to the start of the record: bfill (disaggregate.py:130)
  if (need_fixup) {
    Cell c;
    c.cell = cell; c.D = D; c.ts = ts; c.tpd = tpd; c.nk = 2 * D + 4; c.ld = ld;
    c.off_min = off_min; c.utc_min = 0.0; c.frac = 0.0; c.rise = c.set = nullptr;
    c.doy = f.doy; c.sdoy = nullptr; c.win0 = 0; c.nwin = 0;
    c.t_min = f.t_min; c.t_max = f.t_max; c.st_t_min = f.st_t_min; c.st_t_max = f.st_t_max;
    c.t_end = f.t_end;
    const int T = D * tpd, i_hi = d_hi * tpd;
    double fillv = nan("");
    int j = i_lo;
    for (; j < T; ++j) {
      fillv = prec_step_generic(p, f, c, j);
      if (fillv == fillv) break;
    }
    ooff = (size_t)((long)i_lo - (long)a.out.day0 * tpd) * ostride;
    for (int jj = i_lo; jj < min(j, i_hi); ++jj, ooff += ostride) __stcs(o_prec + ooff, (OutT)fillv);
  }
}

// disaggregate.py:583-589 (passthrough with a daily longwave forcing): every day of the sub-daily
// longwave is scaled so that its mean equals the supplied daily value.  One thread = one
// (cell, day); reads the float64 staging plane the streaming kernel wrote, applies the unit
// conversion and the output cast.
template <typename OutT>
__global__ void __launch_bounds__(128)
lw_rescale_kernel(const __grid_constant__ DisaggArgs a) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  const int d = a.out.day0 + blockIdx.y;
  if (cell >= a.f.n_cells || d >= a.out.day1) return;
  const int tpd = a.tpd;
  const size_t base = (size_t)blockIdx.y * tpd * (size_t)a.out.ld + cell;
  double s = 0.0;
  for (int t = 0; t < tpd; ++t) s += a.lw_raw[base + (size_t)t * a.out.ld];
  const double factor = a.f.given_longwave[(size_t)d * a.f.ld + cell] / (s / (double)tpd);
  OutT *o = reinterpret_cast<OutT *>(a.out.var[MSB_V_LONGWAVE]);
  const double sc = a.out.scale[MSB_V_LONGWAVE], of = a.out.offset[MSB_V_LONGWAVE];
  for (int t = 0; t < tpd; ++t) {
    const size_t k = base + (size_t)t * a.out.ld;
    o[k] = (OutT)fma(a.lw_raw[k] * factor, sc, of);
  }
}

// accuracy probe of the hand-written fp64 sequences (tests only touch it through the C ABI)
__global__ void math_selftest_kernel(int n, const double *__restrict__ x, double *__restrict__ out) {
  __shared__ double etab[kExpTab];
  __shared__ double ebig[kExpBig];
  exp_table_init(etab);
  for (int j = threadIdx.x; j < kExpBig; j += blockDim.x) ebig[j] = g_exp2_big[j];
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  const KnotK K = load_knot();
  out[8 * (size_t)n + i] = wrap_lon(v);
  out[9 * (size_t)n + i] = rcp_2(v);
  out[10 * (size_t)n + i] = sqrt_2(v);
  out[11 * (size_t)n + i] = exp_big(v, ebig, K);
  out[i] = rcp_nr(v);
  out[n + i] = exp_tab(v, etab);
  out[2 * n + i] = sqrt_fast(v);
  out[3 * n + i] = div_nr(1.2345678901234567, v);
  out[4 * n + i] = svp_kpa(v, etab);
  out[5 * n + i] = rcp_seed(v);
  out[6 * n + i] = exp_acc(v, etab);
  out[7 * n + i] = sqrt_nr(v);
}

}  // namespace msb

using namespace msb;

// g_exp2_big of the current device, filled on the first call per device (stream-ordered before use)
static int exp_tables_ready(cudaStream_t st) {
  static std::mutex mu;
  static bool done[64] = {false};
  int dev = 0;
  MSB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= 64) { set_error("msb_disaggregate: device index %d out of range", dev); return MSB_E_ARG; }
  if (!done[dev]) {
    exp_big_init_kernel<<<kExpBig / 256, 256, 0, st>>>();
    MSB_CUDA(cudaGetLastError());
    MSB_CUDA(cudaStreamSynchronize(st));   // later calls may come on other streams
    done[dev] = true;
  }
  return MSB_OK;
}

extern "C" int msb_selftest_math(int n, const double *x_dev, double *out_dev, void *stream) {
  if (n <= 0 || !x_dev || !out_dev) { set_error("msb_selftest_math: bad argument"); return MSB_E_ARG; }
  const int rc_tab = exp_tables_ready(reinterpret_cast<cudaStream_t>(stream));
  if (rc_tab) return rc_tab;
  math_selftest_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(n, x_dev, out_dev);
  MSB_CUDA(cudaGetLastError());
  return MSB_OK;
}

extern "C" size_t msb_disagg_workspace_bytes(int64_t ld, int day0, int day1) {
  if (ld <= 0 || day1 <= day0) return 0;
  const size_t n_rec = (size_t)(day1 - day0) + 2 * MSB_REC_MARGIN;
  return (size_t)MSB_REC_FIELDS * n_rec * (size_t)ld * sizeof(double);
}

// tuning knob read once from the environment (experiments only; the defaults are the tuned values)
static int env_int(const char *name, int dflt, int lo, int hi) {
  const char *v = getenv(name);
  if (!v || !*v) return dflt;
  const int x = atoi(v);
  return x < lo ? lo : (x > hi ? hi : x);
}

extern "C" size_t msb_disagg_workspace_bytes_lw(int64_t ld, int day0, int day1, int time_step) {
  if (ld <= 0 || day1 <= day0 || time_step <= 0 || time_step >= kMinPerDay) return 0;
  const size_t rec = (msb_disagg_workspace_bytes(ld, day0, day1) + 255) & ~(size_t)255;
  return rec + (size_t)(day1 - day0) * (size_t)(kMinPerDay / time_step) * (size_t)ld * sizeof(double);
}

extern "C" int msb_disaggregate(const msb_params *p, const msb_forcing *f,
                                const msb_solar_tables *t, const msb_daily *daily,
                                const msb_subdaily *out, void *stream) {
  if (!p || !f || !t || !daily || !out) { set_error("msb_disaggregate: NULL argument"); return MSB_E_ARG; }
  if (f->n_cells <= 0 || f->n_days < 2 || f->ld < f->n_cells || out->ld < f->n_cells) {
    set_error("msb_disaggregate: bad shapes");
    return MSB_E_ARG;
  }
  const int ts = p->time_step;
  if (ts <= 0 || ts >= kMinPerDay || kMinPerDay % ts || ts > 360) {
    set_error("Time step must be evenly divisible into 1440 minutes (24 hours) and less than 360 "
              "minutes (6 hours). Got %d.", ts);   // metsim.py:641-646
    return MSB_E_ARG;
  }
  if (p->prec_type != MSB_PREC_UNIFORM && out->var[MSB_V_PREC] && (!f->t_pk12 || !f->dur12)) {
    set_error("msb_disaggregate: triangle/mix precipitation needs t_pk12 and dur12");
    return MSB_E_ARG;
  }
  if (out->dtype != 4 && out->dtype != 8) { set_error("msb_disaggregate: out dtype must be 4 or 8"); return MSB_E_ARG; }
  if (out->day0 < 0 || out->day1 > f->n_days || out->day0 >= out->day1) {
    set_error("msb_disaggregate: bad day range [%d, %d)", out->day0, out->day1);
    return MSB_E_ARG;
  }
  const double *vp = daily->var[MSB_D_VAPOR_PRESSURE], *sw = daily->var[MSB_D_SHORTWAVE],
               *tk = daily->var[MSB_D_TSKC];
  if (!vp || !sw || !tk || !daily->status) {
    set_error("msb_disaggregate: daily vapor_pressure / shortwave / tskc / status are required");
    return MSB_E_ARG;
  }
  DisaggArgs a;
  a.p = *p; a.f = *f; a.t = *t; a.vp_d = vp; a.sw_d = sw; a.tskc_d = tk; a.status = daily->status;
  a.out = *out;
  a.tpd = kMinPerDay / ts;
  const int n_days = out->day1 - out->day0;
  // Day tiles: long enough to amortise the per-tile set-up, short enough that (cell blocks x
  // tiles) fills the 148 SMs x 4 resident CTAs several waves deep, and small enough that the
  // element offset inside a tile fits the 32-bit store index.
  const int cell_blocks = (f->n_cells + kDisaggThreads - 1) / kDisaggThreads;
  int tile = 8;
  while (tile > 1 && (long)cell_blocks * ((n_days + tile - 1) / tile) < 148L * 4 * 4) tile >>= 1;
  while (tile > 1 && (double)tile * a.tpd * (double)out->ld >= 4.0e9) tile >>= 1;
  if ((double)tile * a.tpd * (double)out->ld >= 4.0e9) {
    set_error("msb_disaggregate: one day of output (%d steps x ld %lld) exceeds the 32-bit tile index",
              a.tpd, (long long)out->ld);
    return MSB_E_ARG;
  }
  a.tile_days = tile;
  dim3 grid(cell_blocks, (n_days + tile - 1) / tile);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // day records of the padded days around the range (-1 and D are the padded PCHIP ends)
  a.e0 = out->day0 - MSB_REC_MARGIN < -1 ? -1 : out->day0 - MSB_REC_MARGIN;
  const int e1 = out->day1 + MSB_REC_MARGIN > f->n_days + 1 ? f->n_days + 1 : out->day1 + MSB_REC_MARGIN;
  a.n_rec = e1 - a.e0;
  a.rec = reinterpret_cast<double *>(out->workspace);
  if (!out->workspace ||
      out->workspace_bytes < (size_t)MSB_REC_FIELDS * a.n_rec * (size_t)f->ld * sizeof(double)) {
    set_error("msb_disaggregate: workspace of msb_disagg_workspace_bytes() bytes is required");
    return MSB_E_ARG;
  }
  a.sw_full_day = p->method == MSB_METHOD_PASSTHROUGH && !p->sw_daylight;
  a.lw_raw = nullptr;
  if (p->method == MSB_METHOD_PASSTHROUGH && f->given_longwave && out->var[MSB_V_LONGWAVE]) {
    const size_t rec_bytes = (msb_disagg_workspace_bytes(f->ld, out->day0, out->day1) + 255) & ~(size_t)255;
    if (out->workspace_bytes < rec_bytes + (size_t)n_days * a.tpd * (size_t)out->ld * sizeof(double)) {
      set_error("msb_disaggregate: the longwave rescale needs a workspace of "
                "msb_disagg_workspace_bytes_lw() bytes");
      return MSB_E_ARG;
    }
    a.lw_raw = reinterpret_cast<double *>(reinterpret_cast<char *>(out->workspace) + rec_bytes);
  }
  const bool fast_lw = p->lw_type == MSB_LW_PRATA && p->lw_cloud == MSB_CLOUD_DEARDORFF;
  bool units = false;
  a.var_mask = 0;
  for (int v = 0; v < MSB_N_SUBVARS; ++v) {
    if (!out->var[v] || (v == MSB_V_WIND && !f->wind)) continue;
    a.var_mask |= 1u << v;
    if (out->scale[v] != 1.0 || out->offset[v] != 0.0) units = true;
  }
  if (!a.var_mask) { set_error("msb_disaggregate: no output variable requested"); return MSB_E_ARG; }
  // the default variable set (examples/example_nc.conf) without unit conversion: specialised kernels
  const bool example = a.var_mask == kExampleMask && fast_lw && !units && !a.lw_raw;
  const bool f4 = out->dtype == 4;
  a.inv_sw_den = 1.0 / (3600.0 * ((double)ts / 60.0));
  static const int prep_days = env_int("MSB_PREP_DAYS", kPrepDays, 1, 16);
  a.prep_days = prep_days;
  // day records of the padded days [e_a, e_b) (only disagg_kernel reads them)
  auto launch_prepare = [&](int e_a, int e_b) {
    a.prep_r0 = (e_a > a.e0 ? e_a : a.e0) - a.e0;
    a.prep_r1 = (e_b < e1 ? e_b : e1) - a.e0;
    if (a.prep_r0 >= a.prep_r1) return;
    disagg_prepare_kernel<<<dim3(cell_blocks, (a.prep_r1 - a.prep_r0 + prep_days - 1) / prep_days), 128, 0, st>>>(a);
  };
  {
    // the fast path sends only the record's first / last two days through disagg_kernel
    const int D = f->n_days;
    const int a0 = out->day0 > 2 ? out->day0 : 2, b0 = out->day1 < D - 2 ? out->day1 : D - 2;
    if (example && a0 < b0) {
      if (out->day0 < a0) launch_prepare(out->day0 - MSB_REC_MARGIN, a0 + MSB_REC_MARGIN);
      if (b0 < out->day1) launch_prepare(b0 - MSB_REC_MARGIN, out->day1 + MSB_REC_MARGIN);
    } else {
      launch_prepare(a.e0, e1);
    }
  }
#define MSB_LAUNCH(T, F, U, V) disagg_kernel<T, F, U, V><<<grid, kDisaggThreads, 0, st>>>(a)
  if (example) {   // reference defaults, 8 variables: the thermodynamic half and the stream half
    // sub-range [s0, s1) of the call through disagg_kernel<VARS> (output pointers shifted)
    auto launch_old = [&](int vars, int s0, int s1) {
      if (s0 >= s1) return;
      DisaggArgs b = a;
      b.out.day0 = s0; b.out.day1 = s1;
      for (int v = 0; v < MSB_N_SUBVARS; ++v)
        if (b.out.var[v])
          b.out.var[v] = reinterpret_cast<char *>(b.out.var[v]) +
                         (size_t)(s0 - out->day0) * a.tpd * (size_t)out->ld * (size_t)out->dtype;
      int tl = a.tile_days;
      while (tl > 1 && (long)cell_blocks * ((s1 - s0 + tl - 1) / tl) < 148L * 4 * 4) tl >>= 1;
      b.tile_days = tl;
      dim3 g(cell_blocks, (s1 - s0 + tl - 1) / tl);
      if (!f4) {         // float64 outputs: the edge days take the general kernel, all variables at once
        if (vars == kVarsThermo) disagg_kernel<double, true, true, kVarsAny><<<g, kDisaggThreads, 0, st>>>(b);
      }
      else if (vars == kVarsThermo) disagg_kernel<float, true, false, kVarsThermo><<<g, kDisaggThreads, 0, st>>>(b);
      else disagg_kernel<float, true, false, kVarsStreams><<<g, kDisaggThreads, 0, st>>>(b);
    };
    // interior days [2, D-2): knot-aligned kernel; the record's first / last two days: disagg_kernel
    const int D = f->n_days;
    static const int knot_tile = env_int("MSB_KNOT_TILE", 16, 1, 64);
    static const int day_tile = env_int("MSB_DAY_TILE", 16, 1, 32);
    const int a0 = out->day0 > 2 ? out->day0 : 2, b0 = out->day1 < D - 2 ? out->day1 : D - 2;
    if (a0 < b0) {
      const int rc_tab = exp_tables_ready(st);
      if (rc_tab) return rc_tab;
      launch_old(kVarsThermo, out->day0, a0);
      launch_old(kVarsThermo, b0, out->day1);
      a.k.i_lo = a0 * a.tpd; a.k.i_hi = b0 * a.tpd;
      a.k.e0 = a0 - 2 > 0 ? a0 - 2 : 0;
      a.k.e1 = (b0 < D - 2 ? b0 : D - 2) + 1;
      // small shards: shorter chunks so that the grid still fills the 148 SMs several waves deep
      a.k.tile = knot_tile;
      while (a.k.tile > 2 && (long)cell_blocks * ((a.k.e1 - a.k.e0 + a.k.tile - 1) / a.k.tile) < 148L * 7 * 6)
        a.k.tile >>= 1;
      dim3 gk(cell_blocks, (a.k.e1 - a.k.e0 + a.k.tile - 1) / a.k.tile);
      if (f4) thermo_knot_kernel<float, 7><<<gk, kDisaggThreads, 0, st>>>(a);
      else thermo_knot_kernel<double, 7><<<gk, kDisaggThreads, 0, st>>>(a);
      launch_old(kVarsStreams, out->day0, a0);
      launch_old(kVarsStreams, b0, out->day1);
      a.k.i_lo = a.k.i_hi = 0;
      a.k.e0 = a0; a.k.e1 = b0;            // streams_day_kernel: output days, 8 per thread
      a.k.tile = day_tile;
      while (a.k.tile > 2 && (long)cell_blocks * ((b0 - a0 + a.k.tile - 1) / a.k.tile) < 148L * 6 * 6)
        a.k.tile >>= 1;
      static const int day_pf = env_int("MSB_DAY_PF", 4, 0, 16);
      a.k.pf = day_pf;
      dim3 gs(cell_blocks, (b0 - a0 + a.k.tile - 1) / a.k.tile);
      if (f4) streams_day_kernel<float, 6><<<gs, kDisaggThreads, 0, st>>>(a);
      else streams_day_kernel<double, 6><<<gs, kDisaggThreads, 0, st>>>(a);
    } else if (f4) {
      MSB_LAUNCH(float, true, false, kVarsThermo);
      MSB_LAUNCH(float, true, false, kVarsStreams);
    } else {
      MSB_LAUNCH(double, true, true, kVarsAny);
    }
  }
  else if (out->dtype == 4) {
    if (fast_lw) MSB_LAUNCH(float, true, true, kVarsAny); else MSB_LAUNCH(float, false, true, kVarsAny);
  } else {
    if (fast_lw) MSB_LAUNCH(double, true, true, kVarsAny); else MSB_LAUNCH(double, false, true, kVarsAny);
  }
#undef MSB_LAUNCH
  if (a.lw_raw) {
    if (out->dtype == 4) lw_rescale_kernel<float><<<dim3(cell_blocks, n_days), 128, 0, st>>>(a);
    else lw_rescale_kernel<double><<<dim3(cell_blocks, n_days), 128, 0, st>>>(a);
  }
  MSB_CUDA(cudaGetLastError());
  return MSB_OK;
}
