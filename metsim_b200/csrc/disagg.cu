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
//   * the per-minute triangle             -> arithmetic series over the clipped limb ranges;
//   * bincount over 2*ts table entries    -> two loads from the signed prefix table of the
//     latitude row (solar.cu), one per window edge;
//   * PchipInterpolator / interp1d        -> a sliding 4-knot window, coefficients rebuilt only
//     when the output time crosses a knot.
#include "msb_common.cuh"

namespace msb {

constexpr int kDisaggThreads = 128;

struct DisaggArgs {
  msb_params p;
  msb_forcing f;
  msb_solar_tables t;
  const double *vp_d, *sw_d, *tskc_d;  // daily [D][ld] from the MTCLIM pass
  int32_t *status;
  msb_subdaily out;
  int tpd, tile_days;
  unsigned var_mask;   // bit v set = out.var[v] requested
};

// ---- per-thread view of one cell -------------------------------------------------------------
struct Cell {
  int cell, D, ts, tpd, nk;
  long ld;
  int off_min, off_tiny;     // np.roll offsets (minutes / 30 s steps), trunc toward zero
  double utc_min;            // sunrise/sunset shift, disaggregate.py:184-189
  double frac;
  const double *rise, *set, *dayl;  // latitude-row tables [366]
  const double *qrow;               // latitude-row base of the signed prefix table
  const int32_t *doy, *month;
  const double *t_min, *t_max, *st_t_min, *st_t_max, *t_end;
};

__device__ __forceinline__ int floordiv(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// knot j of the padded PCHIP sequence (disaggregate.py:236-257); j in [0, 2D+4)
__device__ __forceinline__ void knot(const Cell &c, int j, double &x, double &y) {
  const int e = (j >> 1) - 1;
  const int de = min(max(e, 0), c.D - 1);
  const int yd = c.doy[de] - 1;
  const double off = (double)de * 1440.0;
  const double rt = (c.rise[yd] - c.utc_min) + off;   // :188-195
  const double st = (c.set[yd] - c.utc_min) + off;
  if (j & 1) x = (c.frac * (st - rt) + rt) - c.ts;    // t_t_max :198-199
  else x = rt - c.ts;                                  // t_t_min :200
  x += 1440.0 * (double)(e - de);                      // padded ends :242-243
  if (e < 0) {
    y = (j & 1) ? c.st_t_max[(size_t)(kState - 1) * c.ld + c.cell]
                : c.st_t_min[(size_t)(kState - 1) * c.ld + c.cell];  // t_begin, metsim.py:753
  } else if (e >= c.D) {
    if (c.t_end) y = c.t_end[(size_t)(j & 1) * c.ld + c.cell];
    else y = (j & 1) ? c.t_max[(size_t)(c.D - 1) * c.ld + c.cell]
                     : c.t_min[(size_t)(c.D - 1) * c.ld + c.cell];
  } else {
    y = (j & 1) ? c.t_max[(size_t)e * c.ld + c.cell] : c.t_min[(size_t)e * c.ld + c.cell];
  }
}

// scipy _cubic.py:279-291
__device__ __forceinline__ double pchip_edge(double h0, double h1, double m0, double m1) {
  double d = ((2 * h0 + h1) * m0 - h0 * m1) / (h0 + h1);
  if (sgn(d) != sgn(m0)) return 0.0;
  if (sgn(m0) != sgn(m1) && fabs(d) > 3.0 * fabs(m0)) return 3.0 * m0;
  return d;
}
// scipy _cubic.py:320-333 (weighted harmonic mean, one division)
__device__ __forceinline__ double pchip_mid(double h0, double h1, double m0, double m1) {
  if (sgn(m1) != sgn(m0) || m1 == 0.0 || m0 == 0.0) return 0.0;
  double w1 = 2 * h1 + h0, w2 = h1 + 2 * h0;
  return ((w1 + w2) * (m0 * m1)) / (w1 * m1 + w2 * m0);
}

struct Cubic {
  int k;                 // interval index: knots k, k+1
  double x0, x1;         // knot times
  double c0, c1, c2, c3; // scipy power basis (CubicHermiteSpline, _cubic.py:158-165)
};

__device__ __noinline__ void cubic_setup(const Cell &c, int k, Cubic &q, int32_t *status) {
  double xa, ya, xb, yb, xc, yc, xd, yd;
  knot(c, k, xb, yb);
  knot(c, k + 1, xc, yc);
  const double h1 = xc - xb;
  const double m1 = (yc - yb) / h1;
  bool bad = !(h1 > 0.0);
  double d0, d1;
  double h0 = 0, m0 = 0, h2 = 0, m2 = 0;
  if (k > 0) {
    knot(c, k - 1, xa, ya);
    h0 = xb - xa;
    m0 = (yb - ya) / h0;
    bad |= !(h0 > 0.0);
  }
  if (k + 2 < c.nk) {
    knot(c, k + 2, xd, yd);
    h2 = xd - xc;
    m2 = (yd - yc) / h2;
    bad |= !(h2 > 0.0);
  }
  d0 = (k > 0) ? pchip_mid(h0, h1, m0, m1) : pchip_edge(h1, h2, m1, m2);
  d1 = (k + 2 < c.nk) ? pchip_mid(h1, h2, m1, m2) : pchip_edge(h1, h0, m1, m0);
  if (bad) status[2] = 1;
  const double inv = 1.0 / h1;
  const double tt = (d0 + d1 - 2 * m1) * inv;
  q.k = k;
  q.x0 = xb;
  q.x1 = xc;
  q.c0 = tt * inv;
  q.c1 = (m1 - d0) * inv - tt;
  q.c2 = d0;
  q.c3 = yb;
}

__device__ __forceinline__ double cubic_eval(const Cubic &q, double x) {
  const double s = x - q.x0;
  return fma(fma(fma(q.c0, s, q.c1), s, q.c2), s, q.c3);
}

// PPoly interval search: clip(searchsorted_right(x_knots, x) - 1, 0, nk-2), starting two days
// before `day` (always at or left of the answer) and walking right.
__device__ __noinline__ void cubic_locate(const Cell &c, int day, double x, Cubic &q,
                                          int32_t *status) {
  int k = max(0, 2 * (day + 1) - 4);
  double xn, yn;
  while (k < c.nk - 2) {
    knot(c, k + 1, xn, yn);
    if (x >= xn) ++k; else break;
  }
  cubic_setup(c, k, q, status);
}

// t_t_min of day d: knot of the vapour-pressure interpolant (disaggregate.py:480)
__device__ __forceinline__ double vp_knot_x(const Cell &c, int d) {
  const int yd = c.doy[d] - 1;
  return ((c.rise[yd] - c.utc_min) + (double)d * 1440.0) - c.ts;
}

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
// exactly 1 (rising) or 0 (falling).  Any window sum is then
//     f * ( units + stepp * tri(ia..ib) + count(ja..jb) + stepq * tri(ja..jb) )
// with every count / triangular number formed in INTEGER arithmetic, so a window that misses the
// limbs is an exact 0 and zero / non-zero placement is bit-exact.
struct PrecDay {
  int p0, nps;   // rising series minutes [p0, p0+nps): value i*stepp
  int pu;        // minute index (relative to p0) after the forced-1.0 element, INT_MAX if none
  int q0, nqs;   // falling series minutes [q0, q0+nqs): value 1 + j*stepq (forced 0.0 end excluded)
  double stepp, stepq, f;   // uniform day: nps < 0 and f = P/1440
  bool nan;
};
struct PrecPos { int i, j, u; };   // clamped positions of a minute inside the two limbs

__device__ __forceinline__ PrecPos prec_pos(const PrecDay &pd, int m) {
  PrecPos q;
  q.i = min(max(m - pd.p0, 0), pd.nps);
  q.j = min(max(m - pd.q0, 0), pd.nqs);
  q.u = (m - pd.p0 >= pd.pu) ? 1 : 0;
  return q;
}
// sum over minutes [a, b) given the positions of a and b (a <= b)
__device__ __forceinline__ double prec_window(const PrecDay &pd, const PrecPos &a, const PrecPos &b,
                                              int n_minutes) {
  if (pd.nps < 0) return (double)n_minutes * pd.f;
  const int tr = (b.i - 1) * b.i - (a.i - 1) * a.i;   // 2 * sum_{i=a.i}^{b.i-1} i
  const int tf = (b.j - 1) * b.j - (a.j - 1) * a.j;
  const double s = (double)(b.u - a.u + b.j - a.j) + 0.5 * (pd.stepp * (double)tr + pd.stepq * (double)tf);
  return pd.f * s;
}

__device__ __forceinline__ void prec_day_setup(const msb_params &p, const msb_forcing &f,
                                               int cell, int d, PrecDay &pd) {
  const size_t o = (size_t)d * f.ld + cell;
  const double P = f.prec[o];
  pd.nan = false;
  const bool uniform = p.prec_type == MSB_PREC_UNIFORM ||
                       (p.prec_type == MSB_PREC_MIX && f.t_min[o] < 0.0);
  if (uniform) {
    pd.nps = -1;
    pd.f = P / (double)kMinPerDay;
    return;
  }
  // metsim.py:278-284: month m (1..12) takes domain entry m (0..11); December keeps prec
  const int m = f.month[d];
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
  pd.stepp = (npf > 1) ? 1.0 / (double)(npf - 1) : 0.0;    // np.linspace(0, 1, npf)
  pd.stepq = (nq > 1) ? -1.0 / (double)(nq - 1) : 0.0;     // np.linspace(1, 0, nq)
  // the falling assignment comes second and overwrites the shared peak minute (:322-323)
  int np = npf;
  if (nq > 0 && npf > 0 && pd.p0 + npf - 1 >= pd.q0) np = max(pd.q0 - pd.p0, 0);
  const bool unit = np == npf && npf > 1;                  // forced 1.0 survives the overwrite
  pd.nps = unit ? npf - 1 : np;
  pd.pu = unit ? npf : 0x7fffffff;
  pd.nqs = nq > 1 ? nq - 1 : nq;                           // forced 0.0 end contributes nothing
  pd.f = 1.0;
  const PrecPos z = {0, 0, 0}, e = prec_pos(pd, kMinPerDay);
  const double S = prec_window(pd, z, e, kMinPerDay);      // np.sum(prec_day), :327
  pd.nan = !(S != 0.0) || !(P == P);   // 0/0 or x/0 * 0 -> NaN day (:327)
  pd.f = P / S;
}

// value of output step i for the rolled per-minute record (disaggregate.py:347-351); NaN when the
// window touches a NaN day.  Generic (slow) form used for the fill look-back only.
__device__ __noinline__ double prec_step_generic(const msb_params &p, const msb_forcing &f,
                                                 const Cell &c, long i) {
  const long M = (long)c.D * kMinPerDay;
  long m0 = (i * c.ts + c.off_min) % M;
  if (m0 < 0) m0 += M;
  const int dA = (int)(m0 / kMinPerDay), rem = (int)(m0 - (long)dA * kMinPerDay);
  const int n1 = min(c.ts, kMinPerDay - rem);
  PrecDay pd;
  prec_day_setup(p, f, c.cell, dA, pd);
  if (pd.nan) return nan("");
  double v = prec_window(pd, prec_pos(pd, rem), prec_pos(pd, rem + n1), n1);
  if (n1 < c.ts) {
    const int dB = (dA + 1 == c.D) ? 0 : dA + 1;
    prec_day_setup(p, f, c.cell, dB, pd);
    if (pd.nan) return nan("");
    v += prec_window(pd, prec_pos(pd, 0), prec_pos(pd, c.ts - n1), c.ts - n1);
  }
  return v;
}

// pandas ffill then bfill of one NaN step (disaggregate.py:130)
__device__ __noinline__ double prec_fill(const msb_params &p, const msb_forcing &f, const Cell &c,
                                         long i) {
  const long T = (long)c.D * c.tpd;
  for (long j = i - 1; j >= 0; --j) {
    double v = prec_step_generic(p, f, c, j);
    if (v == v) return v;
  }
  for (long j = i + 1; j < T; ++j) {
    double v = prec_step_generic(p, f, c, j);
    if (v == v) return v;
  }
  return nan("");
}

// window sum of the normalised radiation row over tiny steps [ca, cb) of table row `yr`
__device__ __forceinline__ double row_window(const Cell &c, int yr, int ca, int cb) {
  const double *q = c.qrow + (size_t)min(yr, 364) * kQLd;
  double w = q[cb] - q[ca];
  if (ca <= kNoon && cb > kNoon) w += q[kTiny + 1];
  return w;
}

// disaggregate.py:491-590 without the daily_lw rescale (mtclim method)
template <bool FAST>
__device__ __forceinline__ double longwave(const msb_params &p, double temp_c, double vp_kpa,
                                           double tskc, const double *etab) {
  const double air_temp = temp_c + kKelvin;
  const double vp = vp_kpa * 10;
  double e;
  if (FAST) {  // PRATA + CLOUD_DEARDORFF, the reference defaults (metsim.py:158-159)
    const double x = 46.5 * vp * rcp_nr(air_temp);
    e = 1 - (1 + x) * exp_tab(-sqrt_nr(fma(3., x, 1.2)), etab);
    const double emis = tskc + (1 - tskc) * e;
    const double t2 = air_temp * air_temp;
    return emis * kStefanB * (t2 * t2);
  }
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

template <typename OutT, bool UNITS>
__device__ __forceinline__ void put(const msb_subdaily &o, int v, size_t idx, double val) {
  OutT *ptr = reinterpret_cast<OutT *>(o.var[v]);
  if (UNITS) val = fma(val, o.scale[v], o.offset[v]);
  __stcs(ptr + idx, (OutT)val);   // streaming store: outputs are never re-read by this kernel
}

// variable sets compiled into the kernel: kVarsExample = the 8 variables of
// examples/example_nc.conf (every pointer known non-NULL, no per-step tests), kVarsAny = runtime
constexpr int kVarsAny = 0, kVarsExample = 1;
constexpr unsigned kExampleMask = (1u << MSB_V_TEMP) | (1u << MSB_V_PREC) | (1u << MSB_V_SHORTWAVE) |
                                  (1u << MSB_V_LONGWAVE) | (1u << MSB_V_VAPOR_PRESSURE) |
                                  (1u << MSB_V_REL_HUMID) | (1u << MSB_V_AIR_PRESSURE) |
                                  (1u << MSB_V_WIND);

// One thread = one cell x one tile of days.  np.roll by a whole number of minutes (tskc, wind,
// prec) or 30 s steps (shortwave) turns every daily series into a contiguous SOURCE STREAM that
// is read ts minutes per output step; the stream position and the current source day live in
// registers and "roll" to the next source day once per day.
template <typename OutT, bool FAST_LW, bool UNITS, int VARS>
__global__ void __launch_bounds__(kDisaggThreads, 4)
disagg_kernel(const __grid_constant__ DisaggArgs a) {
  __shared__ double etab[kExpTab];
  exp_table_init(etab);
  __syncthreads();

  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= a.f.n_cells) return;
  const msb_params &p = a.p;
  const msb_forcing &f = a.f;
  const int D = f.n_days, ts = p.time_step, tpd = a.tpd;
  const int d_lo = a.out.day0 + blockIdx.y * a.tile_days;
  const int d_hi = min(a.out.day1, d_lo + a.tile_days);
  if (d_lo >= d_hi) return;
  const unsigned mask = VARS == kVarsExample ? kExampleMask : a.var_mask;
  auto want = [mask](int v) { return (mask >> v) & 1u; };

  Cell c;
  c.cell = cell; c.D = D; c.ts = ts; c.tpd = tpd; c.nk = 2 * D + 4; c.ld = f.ld;
  const int nk = 2 * D + 4;
  const double lon = remainder(f.lon[cell], 360.0);                      // disaggregate.py:67
  const int off_min = p.utc_offset ? (int)((lon / kDegPerRev) * kMinPerDay) : 0;  // :349,:376,:611
  const int off_tiny = p.utc_offset ? (int)((lon / kDegPerRev) * kTiny) : 0;      // :655
  c.off_min = off_min; c.off_tiny = off_tiny;
  c.utc_min = p.utc_offset ? (double)(int)(((lon - 0.0) / kDegPerRev) * kMinPerDay) : 0.0;  // :186
  c.frac = p.tmax_daylength_fraction;
  const int row = f.lat_row[cell];
  c.rise = a.t.rise_min + (size_t)row * kRows;
  c.set = a.t.set_min + (size_t)row * kRows;
  c.dayl = a.t.daylength + (size_t)row * kRows;
  c.qrow = a.t.q + (size_t)row * 365 * kQLd;
  c.doy = f.doy; c.month = f.month;
  c.t_min = f.t_min; c.t_max = f.t_max; c.st_t_min = f.st_t_min; c.st_t_max = f.st_t_max;
  c.t_end = f.t_end;

  const double elev = f.elev[cell];
  const double pr_num = -(elev * kGStd) / kRDry;            // pressure :401-402
  const double pr_corr = kKelvin + 0.5 * elev * -p.lapse_rate;
  const long ld = f.ld;
  const int T = D * tpd;
  const int C = 2 * ts;                                     // tiny steps per output step (:660)
  const double sw_den = 3600.0 * ((double)ts / 60.0);       // :645
  const double inv_ts = 1.0 / (double)ts;
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);

  // ---- vapour pressure interpolant: knots (t_t_min[d], vp_daily[d]/1000), NaN outside --------
  const double xv_first = vp_knot_x(c, 0), xv_last = vp_knot_x(c, D - 1);
  const int i_head = (xv_first > 0.0) ? (int)ceil(xv_first / ts) : 0;   // first step >= x_v[0]
  const int i_tail = (int)floor(xv_last / ts);                          // last step <= x_v[D-1]
  const int i_lo = d_lo * tpd, i_hi = d_hi * tpd;
  int vd = min(max(d_lo, 1), D - 1);                         // upper knot index, clip [1, D-1]
  {
    const double x0 = (double)ts * (double)i_lo;
    while (vd > 1 && !(vp_knot_x(c, vd - 1) < x0)) --vd;
    while (vd < D - 1 && vp_knot_x(c, vd) < x0) ++vd;
  }
  double xvl = vp_knot_x(c, vd - 1), xvh = vp_knot_x(c, vd);
  double yvl = a.vp_d[(size_t)(vd - 1) * ld + cell] / 1000.0, yvh = a.vp_d[(size_t)vd * ld + cell] / 1000.0;
  double vinv = 1.0 / (xvh - xvl);

  Cubic cub;
  cubic_locate(c, d_lo, (double)ts * (double)i_lo, cub, a.status);
  if (cub.k >= nk - 2) cub.x1 = kInf;   // last interval: extrapolate, never advance

  // head / tail fill values (pandas bfill / ffill of the NaN ends, disaggregate.py:99-101)
  double vp_head = 0.0, vp_tail = 0.0;
  if (i_lo < i_head && i_head < T) {
    const double x = (double)ts * (double)i_head;
    Cubic q2;
    cubic_locate(c, i_head / tpd, x, q2, a.status);
    const double tt = cubic_eval(q2, x);
    const double x0 = vp_knot_x(c, 0), x1 = vp_knot_x(c, 1);
    const double y0 = a.vp_d[cell] / 1000.0, y1 = a.vp_d[(size_t)ld + cell] / 1000.0;
    double v = ((x - x0) / (x1 - x0)) * y1 + ((x1 - x) / (x1 - x0)) * y0;
    vp_head = fmin(v, svp_kpa(tt, etab));
  }
  if (i_hi - 1 > i_tail && i_tail >= 0) {
    const double x = (double)ts * (double)i_tail;
    Cubic q2;
    cubic_locate(c, i_tail / tpd, x, q2, a.status);
    const double tt = cubic_eval(q2, x);
    const double x0 = vp_knot_x(c, D - 2), x1 = vp_knot_x(c, D - 1);
    const double y0 = a.vp_d[(size_t)(D - 2) * ld + cell] / 1000.0;
    const double y1 = a.vp_d[(size_t)(D - 1) * ld + cell] / 1000.0;
    double v = ((x - x0) / (x1 - x0)) * y1 + ((x1 - x) / (x1 - x0)) * y0;
    vp_tail = fmin(v, svp_kpa(tt, etab));
  }
  const bool vp_all_nan = (i_head > i_tail) || i_head >= T || i_tail < 0;
  // steps with an interpolated vapour pressure: [i_sp0, i_sp1]; before: head fill, after: tail
  const int i_sp0 = vp_all_nan ? i_hi : i_head;
  const int i_sp1 = vp_all_nan ? i_lo - 1 : i_tail;
  if (vp_all_nan) vp_head = vp_tail = nan("");

  // ---- source streams ------------------------------------------------------------------------
  const bool want_prec = want(MSB_V_PREC);
  const bool want_wind = want(MSB_V_WIND) && f.wind != nullptr;
  const bool want_sw = want(MSB_V_SHORTWAVE);
  auto wrap = [D](int d) { return d < 0 ? d + D : (d >= D ? d - D : d); };
  auto rad_of = [&](int d) {   // tmp_rad of day d, disaggregate.py:645
    return (a.sw_d[(size_t)d * ld + cell] * c.dayl[f.doy[d] - 1]) / sw_den;
  };
  // minute stream (tskc, wind, prec): source day md, minute sm of that day
  const int qm = off_min < 0 ? -1 : 0;          // floor(off_min / 1440), |off_min| <= 720
  int md = wrap(d_lo + qm);
  int sm = off_min - qm * kMinPerDay;
  PrecDay pd;
  double tk_cur = a.tskc_d[(size_t)md * ld + cell];
  double wd_cur = want_wind ? f.wind[(size_t)md * ld + cell] : 0.0;
  PrecPos ppos = {0, 0, 0};
  if (want_prec) { prec_day_setup(p, f, cell, md, pd); ppos = prec_pos(pd, sm); }
  double prec_last = nan("");   // last valid precipitation step (ffill carry)
  bool prec_last_known = false;
  // tiny-step stream (shortwave): source day sd, tiny step ss of that day, table row per output day
  const int qs = off_tiny < 0 ? -1 : 0;
  int sd = wrap(d_lo + qs);
  int ss = off_tiny - qs * kTiny;
  double rd_cur = want_sw ? rad_of(sd) : 0.0;
  const double *q_cur = c.qrow;
  double qprev = 0.0;

  int i = i_lo;
  size_t oidx = (size_t)(i_lo - a.out.day0 * tpd) * (size_t)a.out.ld + cell;
  double x = (double)ts * (double)i_lo;
  const double dx = (double)ts;
  const size_t ostride = (size_t)a.out.ld;

  for (int d = d_lo; d < d_hi; ++d) {
    const int y = f.doy[d] - 1;
    if (want_sw) {
      // every output day starts in the "lo" part of its window: table row y + qs, rolled over the
      // table's own 366 rows (:656); row 365 is a copy of row 364
      int yl = y + qs;
      yl = yl < 0 ? yl + kRows : yl;
      q_cur = c.qrow + (size_t)min(yl, 364) * kQLd;
      qprev = q_cur[ss];
    }
    for (int t = 0; t < tpd; ++t, ++i, x += dx, oidx += ostride) {
      // -- temperature (PCHIP) --
      if (x >= cub.x1) {
        int k = cub.k + 1;
        double xn, yn;
        while (k < nk - 2) { knot(c, k + 1, xn, yn); if (x >= xn) ++k; else break; }
        cubic_setup(c, k, cub, a.status);
        if (k >= nk - 2) cub.x1 = kInf;
      }
      const double temp = cubic_eval(cub, x);
      const double es = svp_kpa(temp, etab);       // saturation vapour pressure, kPa

      // -- vapour pressure (linear, clamped to saturation, NaN ends filled) --
      double vp;
      if (i >= i_sp0 && i <= i_sp1) {
        while (vd < D - 1 && xvh < x) {     // searchsorted(side='left')
          ++vd;
          xvl = xvh; yvl = yvh;
          xvh = vp_knot_x(c, vd);
          yvh = a.vp_d[(size_t)vd * ld + cell] / 1000.0;
          vinv = 1.0 / (xvh - xvl);
        }
        vp = ((x - xvl) * vinv) * yvh + ((xvh - x) * vinv) * yvl;   // scipy _interpolate.py:513-516
        vp = fmin(vp, es);                                         // disaggregate.py:486-487
      } else {
        vp = i < i_sp0 ? vp_head : vp_tail;
      }

      // -- relative humidity :444-446, air pressure :401-403 --
      const double rh = fmin(100.0 * (vp * rcp_nr(es)), 100.0);   // 100*1000*vp/(svp in Pa)
      const double ap = (kPStd / 1000.0) * exp_tab(pr_num * rcp_nr(pr_corr + temp), etab);

      // -- minute stream: tskc / wind block means :355-380, :593-615; precipitation :265-352 --
      double tskc, wind, pr = 0.0;
      sm += ts;
      if (sm < kMinPerDay) {                      // whole window inside the current source day
        tskc = tk_cur;
        wind = wd_cur;
        if (want_prec) {
          const PrecPos e = prec_pos(pd, sm);
          pr = pd.nan ? nan("") : prec_window(pd, ppos, e, ts);
          ppos = e;
        }
      } else {                                    // window reaches the end of the source day: roll
        const int n2 = sm - kMinPerDay, n1 = ts - n2;   // minutes from the current / next day
        md = (md + 1 == D) ? 0 : md + 1;
        const double tk_n = a.tskc_d[(size_t)md * ld + cell];
        tskc = n2 == 0 ? tk_cur : ((double)n1 * tk_cur + (double)n2 * tk_n) * inv_ts;
        tk_cur = tk_n;
        if (want_wind) {
          const double wd_n = f.wind[(size_t)md * ld + cell];
          wind = n2 == 0 ? wd_cur : ((double)n1 * wd_cur + (double)n2 * wd_n) * inv_ts;
          wd_cur = wd_n;
        } else wind = 0.0;
        if (want_prec) {
          pr = pd.nan ? nan("") : prec_window(pd, ppos, prec_pos(pd, kMinPerDay), n1);
          prec_day_setup(p, f, cell, md, pd);
          const PrecPos z = {0, 0, 0};
          ppos = prec_pos(pd, n2);
          if (n2 > 0) pr = pd.nan ? nan("") : pr + prec_window(pd, z, ppos, n2);
        }
        sm = n2;
      }
      if (want_prec) {
        if (pr != pr) {    // NaN day: pandas ffill, then bfill for a leading run (:130)
          if (!prec_last_known) { prec_last = prec_fill(p, f, c, i); prec_last_known = true; }
          pr = prec_last;
        } else {
          prec_last = pr;
          prec_last_known = true;
        }
      }

      // -- tiny-step stream: shortwave :618-680 --
      double sw = 0.0;
      if (want_sw) {
        const int ca = ss;
        ss += C;
        if (ss < kTiny) {
          const double qn = q_cur[ss];
          double w = qn - qprev;
          if (ca <= kNoon && ss > kNoon) w += q_cur[kTiny + 1];
          qprev = qn;
          sw = rd_cur * w;
        } else {                                  // roll to the next source day / table row
          const int k2 = ss - kTiny;
          double w = q_cur[kTiny] - qprev;        // [ca, 2880)
          if (ca <= kNoon) w += q_cur[kTiny + 1];
          sw = rd_cur * w;
          sd = (sd + 1 == D) ? 0 : sd + 1;
          rd_cur = rad_of(sd);
          int yh = y + qs + 1;
          yh = yh >= kRows ? yh - kRows : yh;
          q_cur = c.qrow + (size_t)min(yh, 364) * kQLd;
          qprev = q_cur[k2];
          if (k2 > 0) sw = fma(rd_cur, qprev - q_cur[0], sw);   // k2 <= 720 < noon: plain prefix
          ss = k2;
        }
      }

      if (want(MSB_V_TEMP)) put<OutT, UNITS>(a.out, MSB_V_TEMP, oidx, temp);
      if (want(MSB_V_VAPOR_PRESSURE)) put<OutT, UNITS>(a.out, MSB_V_VAPOR_PRESSURE, oidx, vp);
      if (want(MSB_V_REL_HUMID)) put<OutT, UNITS>(a.out, MSB_V_REL_HUMID, oidx, rh);
      if (want(MSB_V_AIR_PRESSURE)) put<OutT, UNITS>(a.out, MSB_V_AIR_PRESSURE, oidx, ap);
      if (want(MSB_V_TSKC)) put<OutT, UNITS>(a.out, MSB_V_TSKC, oidx, tskc);
      if (want_wind) put<OutT, UNITS>(a.out, MSB_V_WIND, oidx, wind);
      if (want(MSB_V_SPEC_HUMID)) {   // specific humidity :423-424
        const double mix = (kEps * vp) * rcp_nr(ap - vp);
        put<OutT, UNITS>(a.out, MSB_V_SPEC_HUMID, oidx, mix * rcp_nr(1 + mix));
      }
      if (want(MSB_V_LONGWAVE))
        put<OutT, UNITS>(a.out, MSB_V_LONGWAVE, oidx, longwave<FAST_LW>(p, temp, vp, tskc, etab));
      if (want_prec) put<OutT, UNITS>(a.out, MSB_V_PREC, oidx, pr);
      if (want_sw) put<OutT, UNITS>(a.out, MSB_V_SHORTWAVE, oidx, sw);
    }
  }
}

// accuracy probe of the hand-written fp64 sequences (tests only touch it through the C ABI)
__global__ void math_selftest_kernel(int n, const double *__restrict__ x, double *__restrict__ out) {
  __shared__ double etab[kExpTab];
  exp_table_init(etab);
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  out[i] = rcp_nr(v);
  out[n + i] = exp_tab(v, etab);
  out[2 * n + i] = sqrt_nr(v);
  out[3 * n + i] = div_nr(1.2345678901234567, v);
  out[4 * n + i] = svp_kpa(v, etab);
  out[5 * n + i] = rcp_seed(v);
}

}  // namespace msb

using namespace msb;

extern "C" int msb_selftest_math(int n, const double *x_dev, double *out_dev, void *stream) {
  if (n <= 0 || !x_dev || !out_dev) { set_error("msb_selftest_math: bad argument"); return MSB_E_ARG; }
  math_selftest_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(n, x_dev, out_dev);
  MSB_CUDA(cudaGetLastError());
  return MSB_OK;
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
  // enough (cell block) x (day tile) CTAs to cover 148 SMs several waves deep
  int cell_blocks = (f->n_cells + kDisaggThreads - 1) / kDisaggThreads;
  int tile = 8;
  while (tile > 1 && (long)cell_blocks * ((n_days + tile - 1) / tile) < 148L * 16) tile >>= 1;
  a.tile_days = tile;
  dim3 grid(cell_blocks, (n_days + tile - 1) / tile);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool fast_lw = p->lw_type == MSB_LW_PRATA && p->lw_cloud == MSB_CLOUD_DEARDORFF;
  bool units = false;
  a.var_mask = 0;
  for (int v = 0; v < MSB_N_SUBVARS; ++v) {
    if (!out->var[v] || (v == MSB_V_WIND && !f->wind)) continue;
    a.var_mask |= 1u << v;
    if (out->scale[v] != 1.0 || out->offset[v] != 0.0) units = true;
  }
  if (!a.var_mask) { set_error("msb_disaggregate: no output variable requested"); return MSB_E_ARG; }
  const bool example = a.var_mask == kExampleMask && fast_lw && !units && out->dtype == 4;
#define MSB_LAUNCH(T, F, U, V) disagg_kernel<T, F, U, V><<<grid, kDisaggThreads, 0, st>>>(a)
  if (example) MSB_LAUNCH(float, true, false, kVarsExample);   // reference defaults, 8 variables
  else if (out->dtype == 4) {
    if (fast_lw) MSB_LAUNCH(float, true, true, kVarsAny); else MSB_LAUNCH(float, false, true, kVarsAny);
  } else {
    if (fast_lw) MSB_LAUNCH(double, true, true, kVarsAny); else MSB_LAUNCH(double, false, true, kVarsAny);
  }
#undef MSB_LAUNCH
  MSB_CUDA(cudaGetLastError());
  return MSB_OK;
}
