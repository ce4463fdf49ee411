// Shared device helpers and launch plumbing of the sub-daily disaggregation (disagg*.cu).
//
// Reference: metsim/disaggregate.py:34-130 (disaggregate) and the kernels it calls --
// set_min_max_hour :133-201, temp :204-262 (scipy PCHIP), prec :265-352, wind :355-380,
// pressure :383-403, specific_humidity :406-424, relative_humidity :427-447,
// vapor_pressure :450-488 (scipy interp1d), longwave :491-590, tskc :593-615,
// shortwave :618-680; unit converters metsim/units.py:3-59.
//
// The reference materialises per-minute / per-30-second vectors per cell and reduces them with
// numpy.  Every kernel here has one thread walk one cell through a range of days with the
// intermediates in registers; lanes are consecutive cells, so each output stream is written as
// full 128-byte lines.  What the reference gets from vectors is obtained in closed form:
//   * np.roll + reshape(-1, ts).sum/mean  -> split of each output window at the (shifted) day
//     boundary (tskc, wind, prec) -- integer arithmetic, bit-exact placement;
//   * the per-minute triangle             -> differences of a cumulative arithmetic series whose
//     positions are clamped in INTEGER arithmetic (exact zeros outside the limbs);
//   * bincount over 2*ts table entries    -> difference of two entries of the signed prefix table
//     of the latitude row (solar.cu);
//   * PchipInterpolator / interp1d        -> the current cubic / linear segment in registers.
//
//   disagg.cu           C ABI entry points: argument checks, kernel selection
//   disagg_event.cu     day records + the event-driven kernel (any variable set, any day)
//   disagg_interior.cu  the two specialised kernels of the interior days [2, D-2)
#pragma once

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
  const int32_t *cand_sel;             // single-pass daily form: vp_d / sw_d are candidate plane 0 and
  size_t cand_stride;                  //   a cell reads plane cand_sel[cell], cand_stride elements apart
  int32_t *status;
  msb_subdaily out;
  int tpd, tile_days;
  unsigned var_mask;   // bit v set = out.var[v] requested
  bool units;          // some requested variable has scale != 1 or offset != 0
  bool fast_lw;        // lw_type PRATA + lw_cloud CLOUD_DEARDORFF (the defaults): inlined sequence
  double *rec;         // day records [MSB_REC_FIELDS][n_rec][ld], record r = padded day e0 + r
  double *lw_raw;      // passthrough with a supplied daily longwave: unscaled float64 longwave
                       // [(day1-day0)*tpd][out.ld] for the rescale pass, else NULL
  double inv_sw_den;   // 1 / (3600 * ts / 60), disaggregate.py:645
  int prep_days;       // padded days per thread of disagg_prepare_kernel
  int prep_r0, prep_r1;  // records [prep_r0, prep_r1) are (re)built by this launch
  int fill[4];           // padded days whose records are valid: [fill[0], fill[1]) and [fill[2], fill[3])
                         // (the specialised path only builds the records around the edge days)
  bool sw_full_day;    // passthrough without sw_averaging='daylight': daylength := 86400 (disaggregate.py:78-80)
  // work range of the interior-day kernels (thermo_knot_kernel, sw_day_kernel, prec_day_kernel)
  struct { int e0, e1, tile, i_lo, i_hi, pf; } k;   // pf: L1 prefetch distance of the table stream (steps)
  int e0, n_rec;       // e runs over [-1, D]: -1 / D are the padded PCHIP ends
};

// the cell's daily vapour pressure / shortwave column base (msb_daily.cand, daily.cu)
__device__ __forceinline__ const double *daily_plane(const DisaggArgs &a, const double *plane0, int cell) {
  return a.cand_sel ? plane0 + (size_t)a.cand_sel[cell] * a.cand_stride : plane0;
}

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
// (total: optionally, the unscaled sum of the day's shape = prec_cum(pd, 1440))
__device__ __forceinline__ void prec_day_setup(const msb_params &p, const msb_forcing &f,
                                               int cell, int d, int m, PrecDay &pd,
                                               double *total = nullptr) {
  const size_t o = (size_t)d * f.ld + cell;
  const double P = f.prec[o];
  const bool uniform = p.prec_type == MSB_PREC_UNIFORM ||
                       (p.prec_type == MSB_PREC_MIX && f.t_min[o] < 0.0);
  if (uniform) {
    pd.p0 = 0; pd.nps = 0; pd.pth = kNever; pd.q0 = 0; pd.nqs = kMinPerDay;
    pd.hsp = 0.0; pd.hsq = 0.0;
    pd.f = P * (1.0 / (double)kMinPerDay);
    if (total) *total = (double)kMinPerDay;
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
  if (total) *total = S;
}

// value of output step i for the rolled per-minute record (disaggregate.py:347-351); NaN when the
// window touches a NaN day.  Generic form used by the NaN-day fix-up pass only.
static __device__ __noinline__ double prec_step_generic(const msb_params &p, const msb_forcing &f,
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
static __device__ __noinline__ double prec_lookback(const msb_params &p, const msb_forcing &f, const Cell &c,
                                             long i) {
  for (long j = i - 1; j >= 0; --j) {
    double v = prec_step_generic(p, f, c, j);
    if (v == v) return v;
  }
  return nan("");
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

// ---- launch plumbing (host) --------------------------------------------------------------------
// disagg_event.cu
enum EventKernel { kEvThermoF4, kEvStreamsF4, kEvAnyF4, kEvAnyF8 };
// day records of the padded days [e_a, e_b) clipped to [a.e0, e1) (only disagg_kernel reads them)
int launch_day_records(DisaggArgs &a, int e_a, int e_b, int e1, int cell_blocks, cudaStream_t st);
// disagg_kernel over the call's day range; fast_lw = PRATA + CLOUD_DEARDORFF inlined
int launch_event_kernel(const DisaggArgs &a, EventKernel which, bool fast_lw, int cell_blocks, cudaStream_t st);
int launch_lw_rescale(const DisaggArgs &a, int cell_blocks, cudaStream_t st);

// disagg_interior.cu
int exp_tables_ready(cudaStream_t st);
// interior days [a0, b0): thermo_knot_kernel
int launch_interior_thermo(DisaggArgs &a, int a0, int b0, int cell_blocks, cudaStream_t st);
// disagg_streams.cu: sw_day_kernel + prec_day_kernel over the interior days [a0, b0)
int launch_interior_streams(DisaggArgs &a, int a0, int b0, int cell_blocks, cudaStream_t st);

}  // namespace msb
