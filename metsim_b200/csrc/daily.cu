// Daily MTCLIM pass: rolling means, cloud transmittance, the dew-point fixed point and PET.
//
// Reference: metsim/methods/mtclim.py:28-80 (run) with its helpers :83-256,
// metsim/physics.py:28-152 (calc_pet, atm_pres, svp) and the driver's rolling means
// metsim/metsim.py:589-618.
//
// The reference iterates vapor_pressure -> shortwave -> pet -> tdew on a cell's whole record
// until the RMS change over ALL days drops below tdew_tol.  Per day the map is independent; days
// couple only through the number of evaluations.  B200 design: one thread per (cell, day chunk),
// cells on the lanes so every [D][N] access is coalesced;
//   pass 1  runs 8 evaluations per day in registers and accumulates sum_d (tdew_k - tdew_{k-1})^2
//           per k into a [chunk][k][cell] scratch array (deterministic, no atomics);
//   reduce  adds the chunks in fixed order and picks the evaluation count n* exactly as the
//           reference's while-loop would;
//   pass 2  replays n* evaluations and writes the daily outputs.
// No per-iteration state ever touches HBM.
//
// Single-pass form (msb_daily.cand != NULL, the benchmarked configuration): what the
// disaggregation needs from the daily pass is tskc (independent of n*), vapour pressure and
// shortwave.  n* is 4, 5 or 6 on almost every cell, and evaluation k+1 starts by computing exactly
// vp(tdew_k) and sw(tdew_k) -- so daily_fused_kernel runs six evaluations ONCE, stores the three
// candidate (vp, sw) pairs and the stop-rule partials, the reduce kernel turns n* into a plane
// index per cell, and msb_disaggregate reads that plane.  The replay (7.4 of 13 ms on CONUS) and
// its second read of the inputs disappear; cells with n* outside the window are appended to a list
// and finished by the two-pass kernels running over that list only.
#include <cstring>

#include "msb_common.cuh"

// software prefetch of the fused daily kernel's input rows: 0 = none, 1 = the next day's into L1,
// 2 = the day after next's into L2 (measured per launch on CONUS: 6.12 / 6.10 / 5.91 ms; L1 is nearly
// all shared memory here -- six CTAs of 34 KB -- so only the L2 variant pays)
#ifndef MSB_DAILY_PF
#define MSB_DAILY_PF 2
#endif

namespace msb {

constexpr int kEvalsPerRound = 6;   // the fixed point needs 4-6 evaluations on realistic forcings
constexpr int kDailyThreads = 128;

struct DailyLaunch {
  int chunk_len, n_chunks;
};

static DailyLaunch daily_launch(int n_cells, int n_days, int forced_chunk_len = 0) {
  DailyLaunch l;
  if (forced_chunk_len > 0) {   // msb_tuning.daily_chunk_len: the tests pin the benchmarked geometry
    l.chunk_len = forced_chunk_len < n_days ? forced_chunk_len : n_days;
    l.n_chunks = (n_days + l.chunk_len - 1) / l.chunk_len;
    return l;
  }
  // aim for >= ~2M threads, chunks of at least 8 days
  long want = (2L << 20) / (n_cells > 0 ? n_cells : 1) + 1;
  long max_chunks = (n_days + 7) / 8;
  long nc = want < 1 ? 1 : want;
  if (nc > max_chunks) nc = max_chunks;
  if (nc > 64) nc = 64;
  if (nc < 1) nc = 1;
  l.chunk_len = (int)((n_days + nc - 1) / nc);
  l.n_chunks = (n_days + l.chunk_len - 1) / l.chunk_len;
  return l;
}

// Day-invariant factors of one evaluation of the reference's loop body (mtclim.py:65-71):
//   vp = svp(tdew); t_tmax = max(tt_max + ABASE*vp, 1e-4); sw = potrad*t_tmax*tfmax
//   pet = 1.26*s/(s+gamma)*(0.72*sw)*daylength/lhvap        (physics.py:62-96, x/10*10)
//   ratio = pet/max(seasonal_prec, 80)
//   tdew' = (t_min+K)*(-0.127 + 1.121*(1.003 - 1.444r + 12.312r^2 - 32.766r^3) + 0.0006*dtr) - K
struct DayInv {
  double tk;      // t_min + KELVIN
  double base;    // -0.127 + 1.121*1.003 + 0.0006*dtr
  double g;       // potrad * tfmax
  double tt_max;
  double b;       // ratio per unit shortwave: 1.26*s/(s+gamma)*0.72*daylength/lhvap/sp80
};

// svp in Pa (physics.py:124-152) with the accurate exp of msb_common.cuh (the evaluation count of
// the fixed point must match the reference exactly)
__device__ __forceinline__ double svp_pa(double t, const double *etab) {
  const double r = rcp_nr(237.3 + t);
  const double arg = fma(-17.269 * 237.3, r, 17.269);   // 17.269*t/(237.3+t)
  double s = 610.78 * exp_acc(arg, etab);
  if (t < 0.0) s *= fma(fma(.000042, t, .00972), t, 1.0);
  return s;
}

__device__ __forceinline__ double tdew_step(const DayInv &v, double tdew_old, const double *etab) {
  const double vp = svp_pa(tdew_old, etab);
  const double t_raw = fma(kABase, vp, v.tt_max);
  const double t_tmax = t_raw > 0.0001 ? t_raw : 0.0001;   // np.maximum; t_raw is never NaN here
  const double r = (v.g * t_tmax) * v.b;
  const double poly = r * fma(r, fma(r, -32.766, 12.312), -1.444);
  return fma(v.tk, fma(1.121, poly, v.base), -kKelvin);
}

struct CellConst {
  double pa, px;       // air pressure (Pa), p - p0 for the tt_max0 series
  const double *tt, *dayl, *potrad;  // per-latitude-row table bases
};

__device__ __forceinline__ CellConst cell_const(const msb_params &p, const msb_forcing &f,
                                                const msb_solar_tables &t, int cell) {
  CellConst c;
  double elev = f.elev[cell];
  double t1 = 1.0 - (p.lapse_rate * elev) / kTStd;        // physics.py:118-121, :206-208
  double t2 = kGStd / (p.lapse_rate * (kR / kMA));
  double pr = pow(t1, t2);
  c.pa = kPStd * pr;
  c.px = pr - MSB_TT_P0;
  if (f.given_daylength == nullptr) {
    int row = f.lat_row[cell];
    c.tt = t.tt_coef + (size_t)row * 365 * kTT;
    c.dayl = t.daylength + (size_t)row * kRows;
    c.potrad = t.potrad + (size_t)row * kRows;
  } else {
    c.tt = c.dayl = c.potrad = nullptr;
  }
  return c;
}

// window sums for day d (inclusive), reading the 90-day state for negative indices
__device__ __forceinline__ double hist(const double *cur, const double *st, long ld, int cell,
                                       int k) {
  return (k < 0) ? st[(size_t)(kState + k) * ld + cell] : cur[(size_t)k * ld + cell];
}

template <int PASS>
__global__ void __launch_bounds__(kDailyThreads)
daily_kernel(msb_params p, msb_forcing f, msb_solar_tables t, msb_daily out, int chunk_len,
             int k_off, int max_evals, const int32_t *__restrict__ list) {
  __shared__ double etab[kExpTab];
  exp_table_init(etab);
  __syncthreads();
  // list != NULL: the cells the single-pass form could not finish (list[0] of them, from list[1])
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (list) {
    if (cell >= list[0]) return;
    cell = list[1 + cell];
  }
  if (cell >= f.n_cells) return;
  if (PASS == 1 && k_off > 0 && out.status[3] == 0) return;   // every cell converged earlier
  const int chunk = blockIdx.y;
  const int D = f.n_days;
  const long ld = f.ld;
  const int d_lo = chunk * chunk_len;
  const int d_hi = min(D, d_lo + chunk_len);
  int n_evals = 0;
  if (PASS == 1) {
    if (k_off > 0 && out.n_iter[cell] != 0) return;  // converged in an earlier round
  } else if (PASS == 3) {
    if (chunk == 0) out.n_iter[cell] = 1;            // passthrough: one evaluation, no fixed point
  } else {
    n_evals = out.n_iter[cell];
    if (n_evals == 0) {                               // never converged: flag, use the cap
      n_evals = max_evals;
      out.status[1] = 1;
    }
    if (list && chunk == 0) out.cand_sel[cell] = 0;   // the replay writes candidate plane 0
  }
  const CellConst cc = cell_const(p, f, t, cell);

  // rolling windows (metsim.py:601-616): 90-day precipitation, 30-day temperature range
  const bool roll = f.given_seasonal_prec == nullptr;
  const bool solar = f.given_daylength == nullptr;
  double s_prec = 0.0, s_dtr = 0.0;
  if (roll) {
    for (int k = d_lo - 89; k < d_lo; ++k) s_prec += hist(f.prec, f.st_prec, ld, cell, k);
    for (int k = d_lo - 29; k < d_lo; ++k)
      s_dtr += hist(f.t_max, f.st_t_max, ld, cell, k) - hist(f.t_min, f.st_t_min, ld, cell, k);
  }

  double acc[kEvalsPerRound];
#pragma unroll
  for (int k = 0; k < kEvalsPerRound; ++k) acc[k] = 0.0;

  for (int d = d_lo; d < d_hi; ++d) {
    const size_t o = (size_t)d * ld + cell;
    DayInv v;
    const double t_min = f.t_min[o], t_max = f.t_max[o], prec = f.prec[o];
    const double dtr = t_max - t_min;
    if (dtr < 0.0) out.status[0] = 1;                   // ValueError, metsim.py:612-614
    double seasonal, sm_dtr;
    if (roll) {
      s_prec += prec;
      s_dtr += dtr;
      seasonal = kDaysPerYear * (s_prec * (1.0 / 90.0));
      sm_dtr = s_dtr * (1.0 / 30.0);
      // slide the windows for the next day (d is uniform across the block: no divergence)
      if (d >= 89) s_prec -= f.prec[o - (size_t)89 * ld];
      else s_prec -= f.st_prec[(size_t)(kState + d - 89) * ld + cell];
      if (d >= 29) s_dtr -= f.t_max[o - (size_t)29 * ld] - f.t_min[o - (size_t)29 * ld];
      else s_dtr -= f.st_t_max[(size_t)(kState + d - 29) * ld + cell] -
                    f.st_t_min[(size_t)(kState + d - 29) * ld + cell];
    } else {
      seasonal = f.given_seasonal_prec[o];
      sm_dtr = f.given_smoothed_dtr[o];
    }
    const double sp80 = seasonal < 80.0 ? 80.0 : seasonal;   // mtclim.py:187-188
    double daylength, potrad;

    if (solar) {
      const int y = f.doy[d] - 1;
      daylength = cc.dayl[y];
      potrad = cc.potrad[y];
      // tt_max0(p): Horner in (p - p0), coefficients per (latitude, yday)
      const double *c = cc.tt + (size_t)min(y, 364) * kTT;
      double s = c[kTT - 1];
#pragma unroll
      for (int m = kTT - 2; m >= 0; --m) s = fma(s, cc.px, c[m]);
      v.tt_max = s;
    } else {
      daylength = f.given_daylength[o];
      potrad = f.given_potrad[o];
      v.tt_max = f.given_tt_max[o];
    }
    const double t_mean = (t_min + t_max) / 2;                          // mtclim.py:103-104
    const double t_day = ((t_max - t_mean) * p.tday_coef) + t_mean;
    const double bb = kB0 + kB1 * exp_acc(-kB2 * sm_dtr, etab);        // mtclim.py:129-133
    double tf = 1.0 - 0.9 * exp_acc(-bb * (dtr * sqrt_nr(fmax(dtr, 1e-300))), etab);  // dtr**1.5
    if (prec > p.sw_prec_thresh) tf *= p.rain_scalar;
    // day-invariant part of calc_pet (physics.py:62-92)
    const double lhvap = 2.5023e6 - 2430.54 * t_day;
    const double inv_lhvap = rcp_nr(lhvap);
    const double gamma = (kCp * cc.pa) * (inv_lhvap * (1.0 / kEps));
    const double t1 = t_day + 0.2, t2 = t_day - 0.2;
    const double sl = div_nr(svp_pa(t1, etab) - svp_pa(t2, etab), t1 - t2);
    const double c1 = 1.26 * div_nr(sl, sl + gamma);
    v.tk = t_min + kKelvin;
    v.base = -0.127 + 1.121 * 1.003 + 0.0006 * dtr;
    v.g = potrad * tf;
    v.b = (((c1 * 0.72) * daylength) * inv_lhvap) * rcp_nr(sp80);

    double tdew = t_min;                                                // mtclim.py:54
    if (PASS == 1) {
      for (int k = 0; k < k_off; ++k) tdew = tdew_step(v, tdew, etab);
#pragma unroll
      for (int k = 0; k < kEvalsPerRound; ++k) {
        const double nxt = tdew_step(v, tdew, etab);
        const double df = nxt - tdew;
        acc[k] = fma(df, df, acc[k]);
        tdew = nxt;
      }
    } else {
      double vp, sw, tskc = (p.lw_cloud == MSB_CLOUD_DEARDORFF) ? 1. - tf : sqrt((1. - tf) / 0.65);
      if (PASS == 3) {
        // methods/passthrough.py:28-46: shortwave is a forcing; pet, tdew and vapour pressure
        // follow from it in one evaluation; supplied tskc / vapour pressure columns win
        sw = f.given_shortwave[o];
        if (f.given_tskc) tskc = f.given_tskc[o];
        const double r = sw * v.b;                                      // pet / max(seasonal_prec, 80)
        const double poly = r * fma(r, fma(r, -32.766, 12.312), -1.444);
        tdew = fma(v.tk, fma(1.121, poly, v.base), -kKelvin);           // mtclim.py:190-194
        vp = f.given_vapor_pressure ? f.given_vapor_pressure[o] : svp_pa(tdew, etab);
      } else {
        for (int k = 0; k < n_evals; ++k) tdew = tdew_step(v, tdew, etab);
        vp = svp_pa(tdew, etab);                                        // mtclim.py:73-79
        const double t_tmax = fmax(v.tt_max + (kABase * vp), 0.0001);
        sw = potrad * t_tmax * tf;
      }
      const double pet = ((c1 * (sw * 0.72) * daylength) * inv_lhvap) / 10. * 10.0;
      double *const *o_ = out.var;
      if (o_[MSB_D_T_DAY]) o_[MSB_D_T_DAY][o] = t_day;
      if (o_[MSB_D_TFMAX]) o_[MSB_D_TFMAX][o] = tf;
      if (o_[MSB_D_TSKC]) o_[MSB_D_TSKC][o] = tskc;
      if (o_[MSB_D_TDEW]) o_[MSB_D_TDEW][o] = tdew;
      if (o_[MSB_D_VAPOR_PRESSURE]) o_[MSB_D_VAPOR_PRESSURE][o] = vp;
      if (o_[MSB_D_SHORTWAVE]) o_[MSB_D_SHORTWAVE][o] = sw;
      if (o_[MSB_D_PET]) o_[MSB_D_PET][o] = pet;
      if (o_[MSB_D_DAYLENGTH]) o_[MSB_D_DAYLENGTH][o] = daylength;
      if (o_[MSB_D_POTRAD]) o_[MSB_D_POTRAD][o] = potrad;
      if (o_[MSB_D_TT_MAX]) o_[MSB_D_TT_MAX][o] = v.tt_max;
      if (o_[MSB_D_SEASONAL_PREC]) o_[MSB_D_SEASONAL_PREC][o] = seasonal;
      if (o_[MSB_D_SMOOTHED_DTR]) o_[MSB_D_SMOOTHED_DTR][o] = sm_dtr;
      // daily-output mode: daytime flux -> daily average flux (metsim.py:779-781)
      if (o_[MSB_D_SHORTWAVE_DAILY]) o_[MSB_D_SHORTWAVE_DAILY][o] = sw * (daylength / kSecPerDay);
    }
  }
  if (PASS == 1) {
#pragma unroll
    for (int k = 0; k < kEvalsPerRound; ++k)
      out.partial[((size_t)chunk * kEvalsPerRound + k) * f.n_cells + cell] = acc[k];
  }
}

// while np.sqrt(np.mean((tdew_temp - tdew_old)**2)) > tdew_tol  (mtclim.py:63)
__global__ void daily_reduce_kernel(msb_params p, int n_cells, int n_days, int n_chunks, int k_off,
                                    const double *__restrict__ partial, int32_t *n_iter,
                                    int32_t *status, const int32_t *__restrict__ list) {
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (list) {
    if (cell >= list[0]) return;
    cell = list[1 + cell];
  }
  if (cell >= n_cells) return;
  if (k_off == 0) n_iter[cell] = 0;
  else if (n_iter[cell] != 0) return;
  for (int k = 0; k < kEvalsPerRound; ++k) {
    double s = 0.0;
    for (int c = 0; c < n_chunks; ++c) s += partial[((size_t)c * kEvalsPerRound + k) * n_cells + cell];
    const double rms = sqrt(s / (double)n_days);
    if (!(rms > p.tdew_tol)) { n_iter[cell] = k_off + k + 1; return; }
  }
  atomicAdd(&status[3], 1);   // still iterating: the next round has work to do
}


// ---- single-pass form -----------------------------------------------------------------------
// One thread = (cell, chunk of days), as above, for the product configuration only (mtclim
// method, rolling means from the state, solar tables, no extra daily outputs).  Differences from
// daily_kernel<1>: the 30-day temperature-range window lives in a per-thread shared-memory ring
// (no second read of t_min / t_max rows), and the (vp, sw) pairs after 4, 5 and 6 evaluations go
// to the candidate planes together with tskc.
constexpr int kDtrWin = 30;
__global__ void __launch_bounds__(kDailyThreads)
daily_fused_kernel(msb_params p, msb_forcing f, msb_solar_tables t, msb_daily out, int chunk_len) {
  __shared__ double etab[kExpTab];
  __shared__ double ring[kDtrWin][kDailyThreads];
  exp_table_init(etab);
  __syncthreads();
  const int tid = threadIdx.x;
  const int cell = blockIdx.x * blockDim.x + tid;
  if (cell >= f.n_cells) return;
  const int chunk = blockIdx.y;
  const int D = f.n_days;
  const long ld = f.ld;
  const int d_lo = chunk * chunk_len;
  const int d_hi = min(D, d_lo + chunk_len);
  const CellConst cc = cell_const(p, f, t, cell);

  // rolling windows (metsim.py:601-616): sums over the 89 / 29 days before d_lo, the ring holds
  // the temperature ranges of days d-29 .. d-1 at slot (day mod 30)
  double s_prec = 0.0, s_dtr = 0.0;
  for (int k = d_lo - 89; k < d_lo; ++k) s_prec += hist(f.prec, f.st_prec, ld, cell, k);
  int slot = (d_lo + 1) % kDtrWin;      // slot of day d_lo - 29
  for (int k = d_lo - 29; k < d_lo; ++k) {
    const double v = hist(f.t_max, f.st_t_max, ld, cell, k) - hist(f.t_min, f.st_t_min, ld, cell, k);
    s_dtr += v;
    ring[slot][tid] = v;
    slot = slot + 1 == kDtrWin ? 0 : slot + 1;
  }
  // slot now names day d_lo: (d_lo + 1 + 29) mod 30

  double acc[kEvalsPerRound];
#pragma unroll
  for (int k = 0; k < kEvalsPerRound; ++k) acc[k] = 0.0;
  const size_t plane = (size_t)D * (size_t)ld;
  double *const o_tskc = out.var[MSB_D_TSKC];
  double *const o_cand = out.cand;
  const bool deardorff = p.lw_cloud == MSB_CLOUD_DEARDORFF;

  for (int d = d_lo; d < d_hi; ++d) {
    const size_t o = (size_t)d * ld + cell;
    DayInv v;
#if MSB_DAILY_PF == 1     // the input rows of the next day into L1 while this one computes
    if (d + 1 < d_hi) {
      prefetch_l1(f.t_min + o + ld); prefetch_l1(f.t_max + o + ld); prefetch_l1(f.prec + o + ld);
      if (d + 1 >= 89) prefetch_l1(f.prec + o + ld - (size_t)89 * ld);
    }
#elif MSB_DAILY_PF == 2   // ... of the day after next into L2
    if (d + 2 < d_hi) {
      prefetch_l2(f.t_min + o + 2 * ld); prefetch_l2(f.t_max + o + 2 * ld); prefetch_l2(f.prec + o + 2 * ld);
      if (d + 2 >= 89) prefetch_l2(f.prec + o + 2 * ld - (size_t)89 * ld);
    }
#endif
    const double t_min = f.t_min[o], t_max = f.t_max[o], prec = f.prec[o];
    // the value leaving the 90-day window (d is uniform across the block: no divergence)
    const double prec_out = d >= 89 ? f.prec[o - (size_t)89 * ld]
                                    : f.st_prec[(size_t)(kState + d - 89) * ld + cell];
    const double dtr = t_max - t_min;
    if (dtr < 0.0) out.status[0] = 1;                   // ValueError, metsim.py:612-614
    s_prec += prec;
    s_dtr += dtr;
    const double seasonal = kDaysPerYear * (s_prec * (1.0 / 90.0));
    const double sm_dtr = s_dtr * (1.0 / 30.0);
    s_prec -= prec_out;
    {   // day d-29 leaves the 30-day window; its slot is the one after day d's
      const int nxt = slot + 1 == kDtrWin ? 0 : slot + 1;
      ring[slot][tid] = dtr;
      s_dtr -= ring[nxt][tid];
      slot = nxt;
    }
    const double sp80 = seasonal < 80.0 ? 80.0 : seasonal;   // mtclim.py:187-188
    const int y = f.doy[d] - 1;
    const double daylength = cc.dayl[y];
    const double potrad = cc.potrad[y];
    {   // tt_max0(p): Horner in (p - p0), coefficients per (latitude, yday)
      const double *c = cc.tt + (size_t)min(y, 364) * kTT;
      double s = c[kTT - 1];
#pragma unroll
      for (int m = kTT - 2; m >= 0; --m) s = fma(s, cc.px, c[m]);
      v.tt_max = s;
    }
    const double t_mean = (t_min + t_max) / 2;                          // mtclim.py:103-104
    const double t_day = ((t_max - t_mean) * p.tday_coef) + t_mean;
    const double bb = kB0 + kB1 * exp_acc(-kB2 * sm_dtr, etab);        // mtclim.py:129-133
    double tf = 1.0 - 0.9 * exp_acc(-bb * (dtr * sqrt_nr(fmax(dtr, 1e-300))), etab);  // dtr**1.5
    if (prec > p.sw_prec_thresh) tf *= p.rain_scalar;
    // day-invariant part of calc_pet (physics.py:62-92)
    const double lhvap = 2.5023e6 - 2430.54 * t_day;
    const double inv_lhvap = rcp_nr(lhvap);
    const double gamma = (kCp * cc.pa) * (inv_lhvap * (1.0 / kEps));
    const double t1 = t_day + 0.2, t2 = t_day - 0.2;
    const double sl = div_nr(svp_pa(t1, etab) - svp_pa(t2, etab), t1 - t2);
    const double c1 = 1.26 * div_nr(sl, sl + gamma);
    v.tk = t_min + kKelvin;
    v.base = -0.127 + 1.121 * 1.003 + 0.0006 * dtr;
    v.g = potrad * tf;
    v.b = (((c1 * 0.72) * daylength) * inv_lhvap) * rcp_nr(sp80);
    o_tskc[o] = deardorff ? 1. - tf : sqrt((1. - tf) / 0.65);           // mtclim.py:238-256

    double tdew = t_min;                                                // mtclim.py:54
#pragma unroll
    for (int k = 0; k <= kEvalsPerRound; ++k) {
      // evaluation k+1 of the loop body (mtclim.py:65-71) opens with vp and sw of tdew_k -- the
      // daily outputs of a cell that stops after k evaluations (mtclim.py:73-79)
      const double vp = svp_pa(tdew, etab);
      const double t_raw = fma(kABase, vp, v.tt_max);
      const double t_tmax = t_raw > 0.0001 ? t_raw : 0.0001;
      const double sw = v.g * t_tmax;
      if (k >= MSB_DAILY_CAND_K0) {
        const int j = k - MSB_DAILY_CAND_K0;
        __stcs(o_cand + (size_t)(2 * j) * plane + o, vp);
        __stcs(o_cand + (size_t)(2 * j + 1) * plane + o, sw);
      }
      if (k < kEvalsPerRound) {
        const double r = sw * v.b;
        const double poly = r * fma(r, fma(r, -32.766, 12.312), -1.444);
        const double nxt = fma(v.tk, fma(1.121, poly, v.base), -kKelvin);
        const double df = nxt - tdew;
        acc[k] = fma(df, df, acc[k]);
        tdew = nxt;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kEvalsPerRound; ++k)
    out.partial[((size_t)chunk * kEvalsPerRound + k) * f.n_cells + cell] = acc[k];
}

// the stop rule for every cell after the single pass: n* inside the candidate window picks the
// plane, anything else (n* <= 3, or more than 6 evaluations needed) goes onto the replay list
__global__ void daily_fused_reduce_kernel(msb_params p, int n_cells, int n_days, int n_chunks,
                                          const double *__restrict__ partial, int32_t *n_iter,
                                          int32_t *status, int32_t *sel, int32_t *list) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  int n = 0;
  for (int k = 0; k < kEvalsPerRound; ++k) {
    double s = 0.0;
    for (int c = 0; c < n_chunks; ++c) s += partial[((size_t)c * kEvalsPerRound + k) * n_cells + cell];
    const double rms = sqrt(s / (double)n_days);
    if (!(rms > p.tdew_tol)) { n = k + 1; break; }
  }
  n_iter[cell] = n;
  const int j = n - MSB_DAILY_CAND_K0;
  if (j >= 0 && j < MSB_DAILY_CAND) {
    sel[cell] = j;
  } else {
    sel[cell] = 0;
    list[1 + atomicAdd(&list[0], 1)] = cell;
    if (n == 0) atomicAdd(&status[3], 1);   // still iterating: the next round has work to do
  }
}

}  // namespace msb

using namespace msb;

extern "C" size_t msb_daily_scratch_bytes_chunk(int n_cells, int n_days, int daily_chunk_len) {
  if (n_cells <= 0 || n_days <= 0) return 0;
  DailyLaunch l = daily_launch(n_cells, n_days, daily_chunk_len);
  return (size_t)l.n_chunks * kEvalsPerRound * (size_t)n_cells * sizeof(double);
}
extern "C" size_t msb_daily_scratch_bytes(int n_cells, int n_days) {
  return msb_daily_scratch_bytes_chunk(n_cells, n_days, 0);
}

extern "C" int msb_daily_single_pass(const msb_params *p, const msb_forcing *f, const msb_daily *out) {
  if (!p || !f || !out) return 0;
  if (!out->cand || !out->cand_sel || !out->cand_list || out->tune.daily_two_pass) return 0;
  if (p->method != MSB_METHOD_MTCLIM || f->given_seasonal_prec || f->given_daylength) return 0;
  if (!out->var[MSB_D_TSKC]) return 0;
  for (int k = 0; k < MSB_N_DAILYVARS; ++k)   // the single pass writes tskc and the candidate planes only
    if (k != MSB_D_TSKC && out->var[k]) return 0;
  return 1;
}

static int check_forcing(const msb_forcing *f, const char *who) {
  if (!f || f->n_cells <= 0 || f->n_days < 2 || f->ld < f->n_cells) {
    set_error("%s: bad forcing shape (n_cells=%d n_days=%d ld=%lld)", who, f ? f->n_cells : 0,
              f ? f->n_days : 0, f ? (long long)f->ld : 0LL);
    return MSB_E_ARG;
  }
  const bool roll = f->given_seasonal_prec == nullptr, solar = f->given_daylength == nullptr;
  if (!roll && !f->given_smoothed_dtr) { set_error("%s: given_smoothed_dtr missing", who); return MSB_E_ARG; }
  if (!solar && (!f->given_potrad || !f->given_tt_max)) { set_error("%s: given_potrad/given_tt_max missing", who); return MSB_E_ARG; }
  if (!f->elev || !f->t_min || !f->t_max || !f->prec ||
      (roll && (!f->st_t_min || !f->st_t_max || !f->st_prec)) ||
      (solar && (!f->lat_row || !f->doy))) {
    set_error("%s: a required forcing pointer is NULL", who);
    return MSB_E_ARG;
  }
  return MSB_OK;
}

extern "C" int msb_mtclim_daily(const msb_params *p, const msb_forcing *f,
                                const msb_solar_tables *t, const msb_daily *out, void *stream) {
  if (!p || !out) { set_error("msb_mtclim_daily: NULL argument"); return MSB_E_ARG; }
  int rc = check_forcing(f, "msb_mtclim_daily");
  if (rc) return rc;
  msb_solar_tables no_tables;
  memset(&no_tables, 0, sizeof(no_tables));
  if (!t) {
    if (!f->given_daylength) { set_error("msb_mtclim_daily: solar tables required"); return MSB_E_ARG; }
    t = &no_tables;
  }
  if (!out->n_iter || !out->partial || !out->status) {
    set_error("msb_mtclim_daily: n_iter / partial / status buffers are required");
    return MSB_E_ARG;
  }
  if (!(p->lapse_rate != 0.0)) { set_error("msb_mtclim_daily: lapse_rate must be non-zero"); return MSB_E_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DailyLaunch l = daily_launch(f->n_cells, f->n_days, out->tune.daily_chunk_len);
  dim3 grid((f->n_cells + kDailyThreads - 1) / kDailyThreads, l.n_chunks);
  if (p->method == MSB_METHOD_PASSTHROUGH) {
    if (!f->given_shortwave) {
      set_error("passthrough method: 'shortwave' must be among the forcings");   // passthrough.py:29
      return MSB_E_ARG;
    }
    MSB_CUDA(cudaMemsetAsync(out->status, 0, 4 * sizeof(int32_t), st));
    daily_kernel<3><<<grid, kDailyThreads, 0, st>>>(*p, *f, *t, *out, l.chunk_len, 0, 0, nullptr);
    MSB_CUDA(cudaGetLastError());
    return MSB_OK;
  } else if (p->method != MSB_METHOD_MTCLIM) {
    set_error("msb_mtclim_daily: unknown method %d", p->method);
    return MSB_E_ARG;
  }
  int max_iter = p->max_iter > 0 ? p->max_iter : 24;
  int rounds = (max_iter + kEvalsPerRound - 1) / kEvalsPerRound;
  if (rounds > 8) rounds = 8;
  MSB_CUDA(cudaMemsetAsync(out->status, 0, 4 * sizeof(int32_t), st));
  const int rgrid = (f->n_cells + 255) / 256;
  if (msb_daily_single_pass(p, f, out)) {
    // six evaluations once; candidate planes + stop-rule partials; plane index per cell
    MSB_CUDA(cudaMemsetAsync(out->cand_list, 0, sizeof(int32_t), st));
    daily_fused_kernel<<<grid, kDailyThreads, 0, st>>>(*p, *f, *t, *out, l.chunk_len);
    daily_fused_reduce_kernel<<<rgrid, 256, 0, st>>>(*p, f->n_cells, f->n_days, l.n_chunks, out->partial,
                                                     out->n_iter, out->status, out->cand_sel,
                                                     out->cand_list);
    // the cells outside the candidate window (normally none or a few): further rounds for those
    // that need more than six evaluations, then the replay into candidate plane 0
    for (int r = 1; r < rounds; ++r) {
      int k_off = r * kEvalsPerRound;
      daily_kernel<1><<<grid, kDailyThreads, 0, st>>>(*p, *f, *t, *out, l.chunk_len, k_off, 0, out->cand_list);
      MSB_CUDA(cudaMemsetAsync(out->status + 3, 0, sizeof(int32_t), st));
      daily_reduce_kernel<<<rgrid, 256, 0, st>>>(*p, f->n_cells, f->n_days, l.n_chunks, k_off,
                                                 out->partial, out->n_iter, out->status, out->cand_list);
    }
    msb_daily replay = *out;
    for (int k = 0; k < MSB_N_DAILYVARS; ++k) replay.var[k] = nullptr;   // tskc is already final
    replay.var[MSB_D_VAPOR_PRESSURE] = out->cand;
    replay.var[MSB_D_SHORTWAVE] = out->cand + (size_t)f->n_days * (size_t)f->ld;
    daily_kernel<2><<<grid, kDailyThreads, 0, st>>>(*p, *f, *t, replay, l.chunk_len, 0,
                                                    rounds * kEvalsPerRound, out->cand_list);
    MSB_CUDA(cudaGetLastError());
    return MSB_OK;
  }
  for (int r = 0; r < rounds; ++r) {
    int k_off = r * kEvalsPerRound;
    daily_kernel<1><<<grid, kDailyThreads, 0, st>>>(*p, *f, *t, *out, l.chunk_len, k_off, 0, nullptr);
    MSB_CUDA(cudaMemsetAsync(out->status + 3, 0, sizeof(int32_t), st));   // unconverged-cell count
    daily_reduce_kernel<<<rgrid, 256, 0, st>>>(*p, f->n_cells, f->n_days, l.n_chunks, k_off, out->partial,
                                               out->n_iter, out->status, nullptr);
  }
  daily_kernel<2><<<grid, kDailyThreads, 0, st>>>(*p, *f, *t, *out, l.chunk_len, 0,
                                                  rounds * kEvalsPerRound, nullptr);
  MSB_CUDA(cudaGetLastError());
  return MSB_OK;
}

extern "C" int msb_daily_check(const msb_forcing *f, const msb_daily *out, void *stream) {
  if (!f || !out || !out->status) { set_error("msb_daily_check: NULL argument"); return MSB_E_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int32_t h[4] = {0, 0, 0, 0};
  MSB_CUDA(cudaMemcpyAsync(h, out->status, sizeof(h), cudaMemcpyDeviceToHost, st));
  MSB_CUDA(cudaStreamSynchronize(st));
  if (h[0]) {
    set_error("Daily maximum temperature lower than daily minimum temperature!");
    return MSB_E_DTR_NEGATIVE;
  }
  if (h[1]) {
    set_error("tdew fixed point did not reach tdew_tol within max_iter evaluations");
    return MSB_E_NOT_CONVERGED;
  }
  if (h[2]) {
    set_error("`x` must be strictly increasing sequence. (temperature knots)");
    return MSB_E_KNOTS;
  }
  return MSB_OK;
}
