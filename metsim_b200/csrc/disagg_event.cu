// The event-driven disaggregation kernel and its day records: every variable set, every day of a
// record (NaN ends of the vapour-pressure interpolant, padded PCHIP ends, wrapped streams).  On the
// default variable set it only takes the first / last two days; disagg_interior.cu has the rest.
//
// Loop structure: the kernel is EVENT DRIVEN.  Everything that happens less than once per step --
// a PCHIP or vapour-pressure knot is crossed, a source day or table row rolls over, the radiation
// window crosses solar noon -- is a state update done by a cold handler when the step counter
// reaches the next scheduled event.  Between events the warp runs a tight, call-free, branch-free
// inner loop of straight-line fp64 arithmetic; the run length is the warp-wide minimum of the
// lanes' next events (one REDUX per run), so lanes never lose lockstep on the step index.
// Values that only the handler needs between events live in per-thread shared-memory slots, and
// the one table load of the hot path is software-pipelined one step ahead.
#include "disagg_common.cuh"

namespace msb {

// ---- hot-loop constants ----------------------------------------------------------------------
// ptxas re-materialises an fp64 literal as two 32-bit moves at EVERY use (it never keeps an
// immediate in a register across a loop), which cost ~45 instructions per step.  Values loaded
// from constant memory are opaque to it, so they are loaded once and stay in (uniform) registers.
static __constant__ double c_hot[12] = {
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

// vapour pressure of ONE step at an end of the interpolant's support: the value pandas
// back-fills into the leading NaNs (lo = true: knots 0, 1) or forward-fills into the trailing
// ones (knots D-2, D-1) -- disaggregate.py:480-487, :99-101.
static __device__ __noinline__ double vp_endpoint(const Cell &c, const double *vp_d, int i, bool lo,
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

// per-thread shared-memory slots for values that only the event handler touches between events
enum { kS_TKN = 0, kS_WDN,   // tskc / wind of the source day just rolled in
       kS_ROWTOT,            // total of the current radiation row (noon event)
       kS_VPTAIL,            // forward-fill value of the vapour pressure
       kS_COUNT };

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
  out[kR_RAD * plane] = (daily_plane(a, a.sw_d, cell)[o] * dl) * inv_sw_den;
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
  auto has_rec = [&](int e) {   // a valid record exists for padded day e (else: derive it on the spot)
    return (e >= a.fill[0] && e < a.fill[1]) || (e >= a.fill[2] && e < a.fill[3]);
  };
  auto rkx = [&](int j) { return rec((j & 1) ? kR_XMAX : kR_XMIN, (j >> 1) - 1); };
  auto rky = [&](int j) {      // knot value: the input column, or the padded end (disaggregate.py:242-257)
    const int e = (j >> 1) - 1;
    if ((unsigned)e < (unsigned)D) return ((j & 1) ? f.t_max : f.t_min)[(size_t)e * ld + cell];
    return knot_y(c, j);
  };
  const double *const vp_d = daily_plane(a, a.vp_d, cell), *const sw_d = daily_plane(a, a.sw_d, cell);
  auto rvy = [&](int d) { return vp_d[(size_t)d * ld + cell] / 1000.0; };   // kPa, :480
  auto rad_of = [&](int d) {   // tmp_rad of day d, disaggregate.py:645
    if (has_rec(d)) return rec(kR_RAD, d);
    const double dl = a.sw_full_day ? kSecPerDay : dayl[doy_at(c, d) - 1];
    return (sw_d[(size_t)d * ld + cell] * dl) * inv_sw_den;                        // wrapped day
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
      if (i_hi - 1 > i_tail) vp_tail = vp_endpoint(c, vp_d, i_tail, false, etab, a.status);
      if (i_lo < i_head) {             // head: constant until the first in-range step
        yvl = vp_endpoint(c, vp_d, i_head, true, etab, a.status);
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

int launch_day_records(DisaggArgs &a, int e_a, int e_b, int e1, int cell_blocks, cudaStream_t st) {
  const int prep_days = a.out.tune.prep_days > 0 ? (a.out.tune.prep_days < 16 ? a.out.tune.prep_days : 16)
                                                 : kPrepDays;
  a.prep_days = prep_days;
  a.prep_r0 = (e_a > a.e0 ? e_a : a.e0) - a.e0;
  a.prep_r1 = (e_b < e1 ? e_b : e1) - a.e0;
  if (a.prep_r0 >= a.prep_r1) return MSB_OK;
  // remember what is valid: disagg_kernel derives a (wrapped) day without a record on the spot
  const int slot = (a.fill[0] >= a.fill[1]) ? 0 : 2;
  a.fill[slot] = a.e0 + a.prep_r0;
  a.fill[slot + 1] = a.e0 + a.prep_r1;
  disagg_prepare_kernel<<<dim3(cell_blocks, (a.prep_r1 - a.prep_r0 + prep_days - 1) / prep_days), 128, 0, st>>>(a);
  MSB_CUDA(cudaGetLastError());
  return MSB_OK;
}

int launch_event_kernel(const DisaggArgs &a0, EventKernel which, bool fast_lw, int cell_blocks, cudaStream_t st) {
  DisaggArgs a = a0;
  const int n_days = a.out.day1 - a.out.day0;
  if (n_days <= 0) return MSB_OK;
  // Day tiles: long enough to amortise the per-tile set-up, short enough that (cell blocks x
  // tiles) fills the 148 SMs x 4 resident CTAs several waves deep, and small enough that the
  // element offset inside a tile fits the 32-bit store index.
  int tile = 8;
  while (tile > 1 && (long)cell_blocks * ((n_days + tile - 1) / tile) < 148L * 4 * 4) tile >>= 1;
  while (tile > 1 && (double)tile * a.tpd * (double)a.out.ld >= 4.0e9) tile >>= 1;
  if ((double)tile * a.tpd * (double)a.out.ld >= 4.0e9) {
    set_error("msb_disaggregate: one day of output (%d steps x ld %lld) exceeds the 32-bit tile index",
              a.tpd, (long long)a.out.ld);
    return MSB_E_ARG;
  }
  a.tile_days = tile;
  const dim3 grid(cell_blocks, (n_days + tile - 1) / tile);
#define MSB_LAUNCH(T, F, U, V) disagg_kernel<T, F, U, V><<<grid, kDisaggThreads, 0, st>>>(a)
  switch (which) {
    case kEvThermoF4: MSB_LAUNCH(float, true, false, kVarsThermo); break;
    case kEvStreamsF4: MSB_LAUNCH(float, true, false, kVarsStreams); break;
    case kEvAnyF4:
      if (fast_lw) MSB_LAUNCH(float, true, true, kVarsAny); else MSB_LAUNCH(float, false, true, kVarsAny);
      break;
    case kEvAnyF8:
      if (fast_lw) MSB_LAUNCH(double, true, true, kVarsAny); else MSB_LAUNCH(double, false, true, kVarsAny);
      break;
  }
#undef MSB_LAUNCH
  MSB_CUDA(cudaGetLastError());
  return MSB_OK;
}

int launch_lw_rescale(const DisaggArgs &a, int cell_blocks, cudaStream_t st) {
  const dim3 grid(cell_blocks, a.out.day1 - a.out.day0);
  if (a.out.dtype == 4) lw_rescale_kernel<float><<<grid, 128, 0, st>>>(a);
  else lw_rescale_kernel<double><<<grid, 128, 0, st>>>(a);
  MSB_CUDA(cudaGetLastError());
  return MSB_OK;
}

}  // namespace msb
