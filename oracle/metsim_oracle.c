/* TEST INFRASTRUCTURE ONLY -- see metsim_oracle.h.  Plain scalar C, glibc libm,
 * compiled with -O2 -ffp-contract=off (numba/LLVM and numpy do not contract a*b+c). */
#include "metsim_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- metsim/constants.py:38-80, verbatim values ---- */
#define DEG_PER_REV 360.0
#define SEC_PER_RAD 13750.9871
#define RAD_PER_DAY 0.017214
#define RAD_PER_DEG 0.01745329
#define MIN_DECL (-0.4092797)
#define DAYS_OFF 11.25
#define SW_RAD_DT 30.0
#define MA 28.9644e-3
#define R_GAS 8.3143
#define R_DRY 287.0
#define G_STD 9.80665
#define P_STD 101325.0
#define T_STD 288.15
#define CP 1010.0
#define KELVIN 273.15
#define EPS 0.62196351
#define DAYS_PER_YEAR 365.25
#define SEC_PER_DAY 86400.0
#define MIN_PER_DAY 1440
#define STEFAN_B 5.669e-8
#define MBAR_PER_BAR 1000.0
#define C_RAD 1.5
#define B0 0.031
#define B1 0.201
#define B2 0.185
#define TBASE 0.870
#define ABASE (-6.1e-5)

#define NROW 366
#define NTINY 2880

void mso_default_params(mso_params *p) {
  /* metsim/metsim.py:140-171 */
  p->time_step = 60;
  p->prec_type = MSO_PREC_UNIFORM;
  p->utc_offset = 0;
  p->lw_type = MSO_LW_PRATA;
  p->lw_cloud = MSO_CLOUD_DEARDORFF;
  p->max_iter = 1000;
  p->sw_prec_thresh = 0.0;
  p->rain_scalar = 0.75;
  p->tdew_tol = 1e-6;
  p->tmax_daylength_fraction = 0.67;
  p->tday_coef = 0.45;
  p->lapse_rate = 0.0065;
}

/* physics.py:124-152 (returns Pa) */
static double svp(double t) {
  double s = 0.61078 * exp((17.269 * t) / (237.3 + t));
  if (t < 0.0) s *= 1.0 + .00972 * t + .000042 * (t * t);
  return s * 1000.0;
}

/* physics.py:99-121 */
static double atm_pres(double elev, double lr) {
  double t1 = 1.0 - (lr * elev) / T_STD;
  double t2 = G_STD / (lr * (R_GAS / MA));
  return P_STD * pow(t1, t2);
}

/* physics.py:28-96 (cm/day) */
static double calc_pet(double rad, double ta, double dayl, double pa) {
  const double dt = 0.2;
  double rnet = rad * 0.72;
  double lhvap = 2.5023e6 - 2430.54 * ta;
  double gamma = CP * pa / (lhvap * EPS);
  double t1 = ta + dt, t2 = ta - dt;
  double pvs1 = svp(t1), pvs2 = svp(t2);
  double s = (pvs1 - pvs2) / (t1 - t2);
  double pet = (1.26 * (s / (s + gamma)) * rnet * dayl) / lhvap;
  return pet / 10.;
}

/* physics.py:177-285 */
void mso_solar_geom(double elev, double lat, double lr, double *trf, double *daylength,
                    double *flat_potrad, double *tt_max0) {
  static const double OPTAM[21] = {2.90, 3.05, 3.21, 3.39, 3.69, 3.82, 4.07, 4.37, 4.72, 5.12, 5.60,
                                   6.18, 6.88, 7.77, 8.90, 10.39, 12.44, 15.36, 19.79, 26.96, 30.00};
  const int dayperyear = NROW;
  memset(trf, 0, sizeof(double) * NROW * NTINY);
  memset(daylength, 0, sizeof(double) * NROW);
  memset(flat_potrad, 0, sizeof(double) * NROW);
  memset(tt_max0, 0, sizeof(double) * NROW);

  double t1 = 1.0 - (lr * elev) / T_STD;
  double t2 = G_STD / (lr * (R_GAS / MA));
  double trans = pow(TBASE, pow(t1, t2));

  double latr = fmin(fmax(lat * RAD_PER_DEG, -M_PI / 2.), M_PI / 2.0);
  double coslat = cos(latr), sinlat = sin(latr);
  const double dt = SW_RAD_DT;
  const double dh = dt / SEC_PER_RAD;
  const int tiny_step_per_day = NTINY;

  for (int i = 0; i < dayperyear - 1; ++i) {
    double decl = MIN_DECL * cos((i + DAYS_OFF) * RAD_PER_DAY);
    double cosdecl = cos(decl), sindecl = sin(decl);
    double cosegeom = coslat * cosdecl;
    double sinegeom = sinlat * sindecl;
    double coshss = fmin(fmax(-sinegeom / cosegeom, -1.0), 1.0);
    double hss = acos(coshss);
    daylength[i] = fmin(2.0 * hss * SEC_PER_RAD, SEC_PER_DAY);
    double dir_beam_topa = (1368.0 + 45.5 * sin((2.0 * M_PI * i / DAYS_PER_YEAR) + 1.7)) * dt;
    double sum_trans = 0, sum_flat_potrad = 0;
    double *row = trf + (size_t)i * NTINY;
    /* numba np.arange(-hss, hss, dh): n = ceil((stop-start)/step), v = start + k*step */
    double start = -hss;
    long n = (long)ceil((hss - start) / dh);
    if (n < 0) n = 0;
    for (long k = 0; k < n; ++k) {
      double h = start + (double)k * dh;
      double cosh_ = cos(h);
      double cza = cosegeom * cosh_ + sinegeom;
      double dir_flat_topa;
      if (cza > 0) {
        dir_flat_topa = dir_beam_topa * cza;
        double am = 1.0 / (cza + 0.0000001);
        if (am > 2.9) {
          int ami = (int)(acos(cza) / RAD_PER_DEG) - 69;
          if (ami < 0) ami = 0;
          if (ami > 20) ami = 20;
          am = OPTAM[ami];
        }
        sum_trans += pow(trans, am) * dir_flat_topa;
        sum_flat_potrad += dir_flat_topa;
      } else {
        dir_flat_topa = 0;
      }
      double v = (12 * 3600 + h * SEC_PER_RAD) / dt;
      v = fmin(fmax(v, 0.0), (double)(tiny_step_per_day - 1));
      int tinystep = (int)v;
      row[tinystep] = dir_flat_topa;
    }
    if (daylength[i] != 0.0 && sum_flat_potrad > 0)
      for (int j = 0; j < NTINY; ++j) row[j] /= sum_flat_potrad;
    if (daylength[i] != 0.0) {
      tt_max0[i] = sum_trans / sum_flat_potrad;
      flat_potrad[i] = sum_flat_potrad / daylength[i];
    } else {
      tt_max0[i] = 0.;
      flat_potrad[i] = 0.;
    }
  }
  tt_max0[dayperyear - 1] = tt_max0[dayperyear - 2];
  flat_potrad[dayperyear - 1] = flat_potrad[dayperyear - 2];
  daylength[dayperyear - 1] = daylength[dayperyear - 2];
  memcpy(trf + (size_t)(dayperyear - 1) * NTINY, trf + (size_t)(dayperyear - 2) * NTINY,
         sizeof(double) * NTINY);
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const mso_params *p;
  int n_days;
  double lat, lon, elev;
  const int *doy, *month;
  /* contiguous per-cell copies */
  double *t_min, *t_max, *prec, *wind; /* [D]; wind may be NULL */
  double *st_tmin, *st_tmax, *st_prec; /* [90] */
  const double *t_pk12, *dur12;        /* [12] or NULL */
  const double *t_end;                 /* [2] or NULL */
  /* passthrough method: supplied daily columns [D] (pt_sw != NULL selects the method) */
  int pt_sw_daylight;
  const double *pt_sw, *pt_vp, *pt_tskc, *pt_lw;
} cell_in;

typedef struct {
  double *trf, daylength366[NROW], potrad366[NROW], ttmax366[NROW];
  /* daily [D] */
  double *seasonal, *dtr, *smdtr, *daylength, *potrad, *tt_max, *t_day, *tfmax, *tskc, *tdew,
      *tdew_old, *vp, *sw, *pet, *t_pk, *dur;
  /* sub-daily [T] */
  double *temp, *sprec, *ssw, *lw, *svp_, *rh, *ap, *sh, *stskc, *swind;
  /* scratch */
  double *minute;        /* [D*1440] */
  double *kx, *ky, *kd;  /* PCHIP knots [2D+4] */
  double *t_tmin, *t_tmax;
  int n_iter;
} cell_ws;

static double *dalloc(size_t n) { return (double *)malloc(sizeof(double) * (n ? n : 1)); }

static cell_ws *ws_new(int D, int tpd) {
  cell_ws *w = (cell_ws *)calloc(1, sizeof(cell_ws));
  size_t T = (size_t)D * tpd;
  w->trf = dalloc((size_t)NROW * NTINY);
  double **dd[] = {&w->seasonal, &w->dtr, &w->smdtr, &w->daylength, &w->potrad, &w->tt_max,
                   &w->t_day, &w->tfmax, &w->tskc, &w->tdew, &w->tdew_old, &w->vp, &w->sw,
                   &w->pet, &w->t_pk, &w->dur, &w->t_tmin, &w->t_tmax};
  for (size_t i = 0; i < sizeof(dd) / sizeof(dd[0]); ++i) *dd[i] = dalloc(D);
  double **ss[] = {&w->temp, &w->sprec, &w->ssw, &w->lw, &w->svp_, &w->rh, &w->ap, &w->sh,
                   &w->stskc, &w->swind};
  for (size_t i = 0; i < sizeof(ss) / sizeof(ss[0]); ++i) *ss[i] = dalloc(T);
  w->minute = dalloc((size_t)D * MIN_PER_DAY);
  w->kx = dalloc(2 * (size_t)D + 4);
  w->ky = dalloc(2 * (size_t)D + 4);
  w->kd = dalloc(2 * (size_t)D + 4);
  return w;
}

static void ws_free(cell_ws *w) {
  double *all[] = {w->trf, w->seasonal, w->dtr, w->smdtr, w->daylength, w->potrad, w->tt_max,
                   w->t_day, w->tfmax, w->tskc, w->tdew, w->tdew_old, w->vp, w->sw, w->pet,
                   w->t_pk, w->dur, w->t_tmin, w->t_tmax, w->temp, w->sprec, w->ssw, w->lw,
                   w->svp_, w->rh, w->ap, w->sh, w->stskc, w->swind, w->minute, w->kx, w->ky,
                   w->kd};
  for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); ++i) free(all[i]);
  free(w);
}

/* metsim.py:589-618 -- plain windowed means (see header: xarray rolling is not pinned) */
static int rolling_means(const cell_in *c, cell_ws *w) {
  int D = c->n_days;
  for (int d = 0; d < D; ++d) {
    w->dtr[d] = c->t_max[d] - c->t_min[d];
    if (w->dtr[d] < 0) return -2;
  }
  for (int d = 0; d < D; ++d) {
    double s = 0;
    for (int k = d - 89; k <= d; ++k) s += (k < 0) ? c->st_prec[90 + k] : c->prec[k];
    w->seasonal[d] = DAYS_PER_YEAR * (s / 90.0);
    double q = 0;
    for (int k = d - 29; k <= d; ++k)
      q += (k < 0) ? (c->st_tmax[90 + k] - c->st_tmin[90 + k]) : w->dtr[k];
    w->smdtr[d] = q / 30.0;
  }
  return 0;
}

/* mtclim.py:217-235 */
static double mt_shortwave(double tfmax, double vp, double tt_max, double potrad) {
  double t_tmax = fmax(tt_max + (ABASE * vp), 0.0001);
  return potrad * t_tmax * tfmax;
}

/* mtclim.py:164-194 */
static double mt_tdew(double pet, double t_min, double seasonal_prec, double dtr) {
  if (seasonal_prec < 80.0) seasonal_prec = 80.0;
  double ratio = pet / seasonal_prec;
  return (t_min + KELVIN) * (-0.127 + 1.121 * (1.003 - 1.444 * ratio + 12.312 * (ratio * ratio) -
                                               32.766 * pow(ratio, 3)) + 0.0006 * dtr) - KELVIN;
}

/* mtclim.py:28-80 */
static int mtclim_run(const cell_in *c, cell_ws *w) {
  const mso_params *p = c->p;
  int D = c->n_days;
  double pa = atm_pres(c->elev, p->lapse_rate); /* mtclim.py:160 */
  for (int d = 0; d < D; ++d) {
    /* t_day mtclim.py:83-104 */
    double t_mean = (c->t_min[d] + c->t_max[d]) / 2;
    w->t_day[d] = ((c->t_max[d] - t_mean) * p->tday_coef) + t_mean;
    /* tfmax mtclim.py:107-133 */
    double b = B0 + B1 * exp(-B2 * w->smdtr[d]);
    double tf = 1.0 - 0.9 * exp(-b * pow(w->dtr[d], C_RAD));
    if (c->prec[d] > p->sw_prec_thresh) tf *= p->rain_scalar;
    w->tfmax[d] = tf;
    /* tskc mtclim.py:238-256 */
    w->tskc[d] = (p->lw_cloud == MSO_CLOUD_DEARDORFF) ? 1. - tf : sqrt((1. - tf) / 0.65);
  }
  if (c->pt_sw) { /* methods/passthrough.py:28-46: one evaluation, supplied columns win */
    for (int d = 0; d < D; ++d) {
      if (c->pt_tskc) w->tskc[d] = c->pt_tskc[d];
      w->sw[d] = c->pt_sw[d];
      w->pet[d] = calc_pet(w->sw[d], w->t_day[d], w->daylength[d], pa) * 10.0;
      w->tdew[d] = mt_tdew(w->pet[d], c->t_min[d], w->seasonal[d], w->dtr[d]);
      w->vp[d] = c->pt_vp ? c->pt_vp[d] : svp(w->tdew[d]);
    }
    w->n_iter = 1;
    return 0;
  }
  for (int d = 0; d < D; ++d) w->tdew_old[d] = c->t_min[d];
  int n_eval = 0;
  for (;;) {
    /* one evaluation of vapor_pressure -> shortwave -> pet -> tdew on the whole record */
    double ss = 0;
    for (int d = 0; d < D; ++d) {
      double vp = svp(w->tdew_old[d]);
      double sw = mt_shortwave(w->tfmax[d], vp, w->tt_max[d], w->potrad[d]);
      double pet = calc_pet(sw, w->t_day[d], w->daylength[d], pa) * 10.0;
      w->tdew[d] = mt_tdew(pet, c->t_min[d], w->seasonal[d], w->dtr[d]);
      double df = w->tdew[d] - w->tdew_old[d];
      ss += df * df;
    }
    ++n_eval;
    if (!(sqrt(ss / D) > p->tdew_tol)) break;
    if (n_eval >= p->max_iter) return -4;
    memcpy(w->tdew_old, w->tdew, sizeof(double) * D);
  }
  w->n_iter = n_eval;
  for (int d = 0; d < D; ++d) {
    w->vp[d] = svp(w->tdew[d]);
    w->sw[d] = mt_shortwave(w->tfmax[d], w->vp[d], w->tt_max[d], w->potrad[d]);
    w->pet[d] = calc_pet(w->sw[d], w->t_day[d], w->daylength[d], pa) * 10.0;
  }
  return 0;
}

static long py_mod(long a, long n) {
  long r = a % n;
  return r < 0 ? r + n : r;
}

/* disaggregate.py:618-680 -- np.roll expressed as index arithmetic, products summed in order */
static void dg_shortwave(const cell_in *c, cell_ws *w, double lon) {
  const mso_params *p = c->p;
  int D = c->n_days, ts = p->time_step, tpd = MIN_PER_DAY / ts;
  double ts_hourly = (double)ts / 60.0;
  long off = p->utc_offset ? (long)((lon / DEG_PER_REV) * NTINY) : 0;
  int chunk = (int)(ts * (60.0 / SW_RAD_DT));
  long L = (long)NROW * NTINY, LR = (long)D * NTINY;
  for (int d = 0; d < D; ++d) {
    for (int t = 0; t < tpd; ++t) {
      double acc = 0;
      for (int k = 0; k < chunk; ++k) {
        long it = py_mod((long)(c->doy[d] - 1) * NTINY + (long)t * chunk + k + off, L);
        long ir = py_mod((long)d * NTINY + (long)t * chunk + k + off, LR) / NTINY;
        double dayl = (c->pt_sw && !c->pt_sw_daylight) ? 86400.0 : w->daylength[ir];
        double tmp_rad = (w->sw[ir] * dayl) / (3600 * ts_hourly);
        acc += w->trf[it] * tmp_rad;
      }
      w->ssw[(size_t)d * tpd + t] = acc;
    }
  }
}

/* disaggregate.py:133-201 */
static void dg_set_min_max_hour(const cell_in *c, cell_ws *w, double lon) {
  const mso_params *p = c->p;
  int D = c->n_days, ts = p->time_step;
  double rise_l[NROW], set_l[NROW];
  for (int i = 0; i < NROW; ++i) {
    const double *row = w->trf + (size_t)i * NTINY;
    double r = 0, s = 0;
    for (int j = 0; j < NTINY - 1; ++j) {
      int m = (row[j + 1] > 0) - (row[j] > 0);
      if (m > 0) r = j * (SW_RAD_DT / 60.0);
      if (m < 0) s = j * (SW_RAD_DT / 60.0);
    }
    double day_tot = (SEC_PER_DAY / SW_RAD_DT) * (SW_RAD_DT / 60.0);
    if (r == 0) r = day_tot / 6;
    if (s == 0) s = 4 * day_tot / 6;
    rise_l[i] = r;
    set_l[i] = s;
  }
  if (p->utc_offset) {
    int utc = (int)(((lon - 0) / DEG_PER_REV) * MIN_PER_DAY);
    for (int i = 0; i < NROW; ++i) { rise_l[i] -= utc; set_l[i] -= utc; }
  }
  for (int d = 0; d < D; ++d) {
    int y = c->doy[d] - 1;
    double off = (double)d * 1440.0;
    double rt = rise_l[y] + off, st = set_l[y] + off;
    w->t_tmax[d] = (p->tmax_daylength_fraction * (st - rt) + rt) - ts;
    w->t_tmin[d] = rt - ts;
  }
}

static double sgn(double x) { return (x > 0) - (x < 0); }

/* scipy _cubic.py:279-291 */
static double pchip_edge(double h0, double h1, double m0, double m1) {
  double d = ((2 * h0 + h1) * m0 - h0 * m1) / (h0 + h1);
  if (sgn(d) != sgn(m0)) return 0.;
  if (sgn(m0) != sgn(m1) && fabs(d) > 3. * fabs(m0)) return 3. * m0;
  return d;
}

/* disaggregate.py:204-262 + scipy PchipInterpolator(extrapolate=True) */
static void dg_temp(const cell_in *c, cell_ws *w) {
  int D = c->n_days, ts = c->p->time_step, tpd = MIN_PER_DAY / ts;
  int nk = 2 * D + 4;
  double *x = w->kx, *y = w->ky, *dk = w->kd;
  for (int d = 0; d < D; ++d) {
    x[2 + 2 * d] = w->t_tmin[d];
    x[3 + 2 * d] = w->t_tmax[d];
    y[2 + 2 * d] = c->t_min[d];
    y[3 + 2 * d] = c->t_max[d];
  }
  x[0] = x[2] - 1440;
  x[1] = x[3] - 1440;
  x[nk - 2] = x[nk - 4] + 1440;
  x[nk - 1] = x[nk - 3] + 1440;
  y[0] = c->st_tmin[89]; /* t_begin, metsim.py:753-754 */
  y[1] = c->st_tmax[89];
  if (c->t_end) { y[nk - 2] = c->t_end[0]; y[nk - 1] = c->t_end[1]; }
  else { y[nk - 2] = y[nk - 4]; y[nk - 1] = y[nk - 3]; }
  /* derivatives, _cubic.py:294-340 */
  for (int k = 1; k < nk - 1; ++k) {
    double h0 = x[k] - x[k - 1], h1 = x[k + 1] - x[k];
    double m0 = (y[k] - y[k - 1]) / h0, m1 = (y[k + 1] - y[k]) / h1;
    if (sgn(m1) != sgn(m0) || m1 == 0 || m0 == 0) dk[k] = 0.0;
    else {
      double w1 = 2 * h1 + h0, w2 = h1 + 2 * h0;
      double whmean = (w1 / m0 + w2 / m1) / (w1 + w2);
      dk[k] = 1.0 / whmean;
    }
  }
  {
    double h0 = x[1] - x[0], h1 = x[2] - x[1];
    double m0 = (y[1] - y[0]) / h0, m1 = (y[2] - y[1]) / h1;
    dk[0] = pchip_edge(h0, h1, m0, m1);
    double g0 = x[nk - 1] - x[nk - 2], g1 = x[nk - 2] - x[nk - 3];
    double n0 = (y[nk - 1] - y[nk - 2]) / g0, n1 = (y[nk - 2] - y[nk - 3]) / g1;
    dk[nk - 1] = pchip_edge(g0, g1, n0, n1);
  }
  /* evaluation: PPoly, interval = clip(searchsorted_right(x, t) - 1, 0, nk-2) */
  int k = 0;
  for (long i = 0; i < (long)D * tpd; ++i) {
    double t = (double)ts * (double)i;
    while (k < nk - 2 && t >= x[k + 1]) ++k;
    double dx = x[k + 1] - x[k];
    double slope = (y[k + 1] - y[k]) / dx;
    double tt = (dk[k] + dk[k + 1] - 2 * slope) / dx;
    double c0 = tt / dx, c1 = (slope - dk[k]) / dx - tt, c2 = dk[k], c3 = y[k];
    double s = t - x[k];
    w->temp[i] = ((c0 * s + c1) * s + c2) * s + c3;
  }
}

/* disaggregate.py:450-488 + the ffill/bfill of :96-101 */
static void dg_vapor_pressure(const cell_in *c, cell_ws *w) {
  int D = c->n_days, ts = c->p->time_step, tpd = MIN_PER_DAY / ts;
  long T = (long)D * tpd;
  const double *x = w->t_tmin;
  for (long i = 0; i < T; ++i) {
    double t = (double)ts * (double)i;
    double v;
    if (t < x[0] || t > x[D - 1]) v = NAN;
    else {
      /* searchsorted(side='left') clipped to [1, D-1] */
      int idx = 0;
      while (idx < D && x[idx] < t) ++idx;
      if (idx < 1) idx = 1;
      if (idx > D - 1) idx = D - 1;
      double xl = x[idx - 1], xh = x[idx];
      double yl = w->vp[idx - 1] / MBAR_PER_BAR, yh = w->vp[idx] / MBAR_PER_BAR;
      v = ((t - xl) / (xh - xl)) * yh + ((xh - t) / (xh - xl)) * yl;
    }
    double vp_sat = svp(w->temp[i]) / MBAR_PER_BAR;
    if (vp_sat < v) v = vp_sat;
    w->svp_[i] = v;
  }
}

static void ffill_bfill(double *a, long n) {
  long first = -1;
  for (long i = 0; i < n; ++i) {
    if (isnan(a[i])) { if (i > 0 && !isnan(a[i - 1])) a[i] = a[i - 1]; }
    if (first < 0 && !isnan(a[i])) first = i;
  }
  if (first > 0) for (long i = 0; i < first; ++i) a[i] = a[first];
}

/* disaggregate.py:265-352 */
static void dg_prec(const cell_in *c, cell_ws *w, double lon) {
  const mso_params *p = c->p;
  int D = c->n_days, ts = p->time_step, tpd = MIN_PER_DAY / ts;
  double *pm = w->minute;
  long M = (long)D * MIN_PER_DAY;
  if (p->prec_type == MSO_PREC_UNIFORM) {
    for (int d = 0; d < D; ++d) {
      double v = c->prec[d] / MIN_PER_DAY;
      for (int m = 0; m < MIN_PER_DAY; ++m) pm[(long)d * MIN_PER_DAY + m] = v;
    }
  } else {
    int do_mix = p->prec_type == MSO_PREC_MIX;
    for (int d = 0; d < D; ++d) {
      double *day = pm + (long)d * MIN_PER_DAY;
      double P = c->prec[d];
      if (do_mix && c->t_min[d] < 0) {
        for (int m = 0; m < MIN_PER_DAY; ++m) day[m] = P * 1.0 / MIN_PER_DAY;
        continue;
      }
      for (int m = 0; m < MIN_PER_DAY; ++m) day[m] = 0.0;
      double t_pk = w->t_pk[d], t_dur = w->dur[d];
      double t_start = t_pk - 0.5 * t_dur, t_stop = t_pk + 0.5 * t_dur;
      /* boolean masks over minutes 0..1439; both index sets are contiguous */
      int plus0 = -1, nplus = 0, minus0 = -1, nminus = 0;
      for (int m = 0; m < MIN_PER_DAY; ++m) {
        if ((double)m <= t_pk && (double)m >= t_start) { if (plus0 < 0) plus0 = m; ++nplus; }
        if ((double)m >= t_pk && (double)m <= t_stop) { if (minus0 < 0) minus0 = m; ++nminus; }
      }
      /* np.linspace(0, 1.0, n): arange(n)*step + start, last := stop; n==1 -> [start] */
      if (nplus == 1) day[plus0] = 0.0;
      else if (nplus > 1) {
        double step = (1.0 - 0.0) / (double)(nplus - 1);
        for (int i = 0; i < nplus; ++i) day[plus0 + i] = (double)i * step + 0.0;
        day[plus0 + nplus - 1] = 1.0;
      }
      if (nminus == 1) day[minus0] = 1.0;
      else if (nminus > 1) {
        double step = (0.0 - 1.0) / (double)(nminus - 1);
        for (int i = 0; i < nminus; ++i) day[minus0 + i] = (double)i * step + 1.0;
        day[minus0 + nminus - 1] = 0.0;
      }
      double s = 0;
      for (int m = 0; m < MIN_PER_DAY; ++m) s += day[m];
      double f = P / s; /* 0/0 -> NaN day, disaggregate.py:327 */
      for (int m = 0; m < MIN_PER_DAY; ++m) day[m] = f * day[m];
    }
  }
  long off = p->utc_offset ? (long)((lon / DEG_PER_REV) * MIN_PER_DAY) : 0;
  for (long i = 0; i < (long)D * tpd; ++i) {
    double s = 0;
    for (int k = 0; k < ts; ++k) s += pm[py_mod(i * ts + k + off, M)];
    w->sprec[i] = s;
  }
}

/* disaggregate.py:355-380 (wind) and :593-615 (tskc): repeat per minute, roll, block mean */
static void dg_repeat_mean(const cell_in *c, const double *daily, double *outv, double lon) {
  const mso_params *p = c->p;
  int D = c->n_days, ts = p->time_step, tpd = MIN_PER_DAY / ts;
  long M = (long)D * MIN_PER_DAY;
  long off = p->utc_offset ? (long)((lon / DEG_PER_REV) * MIN_PER_DAY) : 0;
  for (long i = 0; i < (long)D * tpd; ++i) {
    double s = 0;
    for (int k = 0; k < ts; ++k) s += daily[py_mod(i * ts + k + off, M) / MIN_PER_DAY];
    outv[i] = s / ts;
  }
}

/* disaggregate.py:491-590 (no daily_lw rescale: mtclim method) */
static double dg_longwave(const mso_params *p, double air_temp_c, double vp_kpa, double tskc) {
  double air_temp = air_temp_c + KELVIN;
  double vp = vp_kpa * 10;
  double e;
  switch (p->lw_type) {
    case MSO_LW_ANDERSON: e = 0.68 + 0.036 * sqrt(vp); break;
    case MSO_LW_BRUTSAERT: e = 1.24 * pow(vp / air_temp, 0.14285714); break;
    case MSO_LW_SATTERLUND: e = 1.08 * (1 - exp(-1 * pow(vp, (air_temp / 2016)))); break;
    case MSO_LW_IDSO: e = 0.7 + 5.95e-5 * vp * exp(1500 / air_temp); break;
    case MSO_LW_PRATA: {
      double x = 46.5 * vp / air_temp;
      e = (1 - (1 + x) * exp(-sqrt((1.2 + 3. * x))));
      break;
    }
    default: e = 0.74 + 0.0049 * vp; break;
  }
  double emis = (p->lw_cloud == MSO_CLOUD_DEARDORFF) ? tskc + (1 - tskc) * e
                                                      : (1.0 + (0.17 * (tskc * tskc))) * e;
  return emis * STEFAN_B * pow(air_temp, 4);
}

static int run_one(const cell_in *c, cell_ws *w) {
  const mso_params *p = c->p;
  int D = c->n_days, ts = p->time_step;
  int rc = rolling_means(c, w);
  if (rc) return rc;
  /* metsim.py:278-284 (month off-by-one: December keeps the prec values) */
  if (p->prec_type != MSO_PREC_UNIFORM && c->t_pk12 && c->dur12) {
    for (int d = 0; d < D; ++d) {
      int m = c->month[d]; /* 1..12 */
      w->t_pk[d] = (m <= 11) ? c->t_pk12[m] : c->prec[d];
      w->dur[d] = (m <= 11) ? c->dur12[m] : c->prec[d];
    }
  }
  /* metsim.py:733-739 */
  mso_solar_geom(c->elev, c->lat, p->lapse_rate, w->trf, w->daylength366, w->potrad366,
                 w->ttmax366);
  for (int d = 0; d < D; ++d) {
    int y = c->doy[d] - 1;
    w->daylength[d] = w->daylength366[y];
    w->potrad[d] = w->potrad366[y];
    w->tt_max[d] = w->ttmax366[y];
  }
  rc = mtclim_run(c, w);
  if (rc) return rc;
  if (ts >= MIN_PER_DAY) return 0;

  int tpd = MIN_PER_DAY / ts;
  long T = (long)D * tpd;
  double lon = remainder(c->lon, 360.0); /* disaggregate.py:67 */
  dg_shortwave(c, w, lon);
  dg_set_min_max_hour(c, w, lon);
  dg_temp(c, w);
  dg_vapor_pressure(c, w);
  ffill_bfill(w->svp_, T);
  for (long i = 0; i < T; ++i) {
    /* relative_humidity :427-447 */
    double rh = 100.0 * MBAR_PER_BAR * (w->svp_[i] / svp(w->temp[i]));
    if (rh > 100.0) rh = 100.0;
    w->rh[i] = rh;
    /* pressure :383-403 */
    double temp_corr = KELVIN + w->temp[i] + 0.5 * c->elev * -p->lapse_rate;
    double ratio = -(c->elev * G_STD) / (R_DRY * temp_corr);
    w->ap[i] = P_STD * exp(ratio) / MBAR_PER_BAR;
    /* specific_humidity :406-424 */
    double mix = (EPS * w->svp_[i]) / (w->ap[i] - w->svp_[i]);
    w->sh[i] = mix / (1 + mix);
  }
  dg_repeat_mean(c, w->tskc, w->stskc, lon);
  for (long i = 0; i < T; ++i) w->lw[i] = dg_longwave(p, w->temp[i], w->svp_[i], w->stskc[i]);
  if (c->pt_lw) { /* disaggregate.py:583-589: per-day rescale to the supplied daily longwave */
    for (int d = 0; d < D; ++d) {
      double m = 0;
      for (int t = 0; t < tpd; ++t) m += w->lw[(size_t)d * tpd + t];
      double factor = c->pt_lw[d] / (m / tpd);
      for (int t = 0; t < tpd; ++t) w->lw[(size_t)d * tpd + t] *= factor;
    }
  }
  dg_prec(c, w, lon);
  if (c->wind) dg_repeat_mean(c, c->wind, w->swind, lon);
  /* final frame-wide ffill/bfill, disaggregate.py:130 */
  ffill_bfill(w->sprec, T);
  ffill_bfill(w->ssw, T);
  ffill_bfill(w->temp, T);
  ffill_bfill(w->lw, T);
  ffill_bfill(w->rh, T);
  ffill_bfill(w->ap, T);
  ffill_bfill(w->sh, T);
  ffill_bfill(w->stskc, T);
  if (c->wind) ffill_bfill(w->swind, T);
  return 0;
}

typedef struct {
  const mso_params *p;
  int n_cells, n_days, c0, c1, rc;
  const double *lat, *lon, *elev;
  const int *doy, *month;
  const double *t_min, *t_max, *prec, *wind, *st_tmin, *st_tmax, *st_prec, *t_pk12, *dur12, *t_end;
  const mso_outputs *out;
  const mso_passthrough *pt;
} job;

static void gather(double *dst, const double *src, int n, int N, int cidx) {
  for (int i = 0; i < n; ++i) dst[i] = src[(size_t)i * N + cidx];
}
static void scatter(double *dst, const double *src, long n, int N, int cidx) {
  if (!dst) return;
  for (long i = 0; i < n; ++i) dst[(size_t)i * N + cidx] = src[i];
}

static void *worker(void *arg) {
  job *j = (job *)arg;
  int D = j->n_days, N = j->n_cells;
  int ts = j->p->time_step;
  int tpd = ts >= MIN_PER_DAY ? 1 : MIN_PER_DAY / ts;
  cell_ws *w = ws_new(D, tpd);
  double *b_tmin = dalloc(D), *b_tmax = dalloc(D), *b_prec = dalloc(D), *b_wind = dalloc(D);
  double *b_psw = dalloc(D), *b_pvp = dalloc(D), *b_ptk = dalloc(D), *b_plw = dalloc(D);
  double s_tmin[90], s_tmax[90], s_prec[90], tpk[12], dur[12], tend[2];
  for (int cidx = j->c0; cidx < j->c1; ++cidx) {
    cell_in c;
    memset(&c, 0, sizeof(c));
    c.p = j->p; c.n_days = D;
    c.lat = j->lat[cidx]; c.lon = j->lon[cidx]; c.elev = j->elev[cidx];
    c.doy = j->doy; c.month = j->month;
    gather(b_tmin, j->t_min, D, N, cidx); gather(b_tmax, j->t_max, D, N, cidx);
    gather(b_prec, j->prec, D, N, cidx);
    c.t_min = b_tmin; c.t_max = b_tmax; c.prec = b_prec;
    if (j->wind) { gather(b_wind, j->wind, D, N, cidx); c.wind = b_wind; }
    gather(s_tmin, j->st_tmin, 90, N, cidx); gather(s_tmax, j->st_tmax, 90, N, cidx);
    gather(s_prec, j->st_prec, 90, N, cidx);
    c.st_tmin = s_tmin; c.st_tmax = s_tmax; c.st_prec = s_prec;
    if (j->t_pk12 && j->dur12) {
      gather(tpk, j->t_pk12, 12, N, cidx); gather(dur, j->dur12, 12, N, cidx);
      c.t_pk12 = tpk; c.dur12 = dur;
    }
    if (j->t_end) { gather(tend, j->t_end, 2, N, cidx); c.t_end = tend; }
    if (j->pt && j->pt->shortwave) {
      c.pt_sw_daylight = j->pt->sw_daylight;
      gather(b_psw, j->pt->shortwave, D, N, cidx); c.pt_sw = b_psw;
      if (j->pt->vapor_pressure) { gather(b_pvp, j->pt->vapor_pressure, D, N, cidx); c.pt_vp = b_pvp; }
      if (j->pt->tskc) { gather(b_ptk, j->pt->tskc, D, N, cidx); c.pt_tskc = b_ptk; }
      if (j->pt->longwave) { gather(b_plw, j->pt->longwave, D, N, cidx); c.pt_lw = b_plw; }
    }
    int rc = run_one(&c, w);
    if (rc) { j->rc = rc; break; }
    const mso_outputs *o = j->out;
    scatter(o->t_day, w->t_day, D, N, cidx); scatter(o->tfmax, w->tfmax, D, N, cidx);
    scatter(o->tskc_daily, w->tskc, D, N, cidx); scatter(o->tdew, w->tdew, D, N, cidx);
    scatter(o->vp_daily, w->vp, D, N, cidx); scatter(o->sw_daily, w->sw, D, N, cidx);
    scatter(o->pet, w->pet, D, N, cidx); scatter(o->daylength, w->daylength, D, N, cidx);
    scatter(o->potrad, w->potrad, D, N, cidx); scatter(o->tt_max, w->tt_max, D, N, cidx);
    scatter(o->seasonal_prec, w->seasonal, D, N, cidx);
    scatter(o->smoothed_dtr, w->smdtr, D, N, cidx);
    if (o->n_iter) o->n_iter[cidx] = w->n_iter;
    if (ts < MIN_PER_DAY) {
      long T = (long)D * tpd;
      scatter(o->temp, w->temp, T, N, cidx); scatter(o->prec, w->sprec, T, N, cidx);
      scatter(o->shortwave, w->ssw, T, N, cidx); scatter(o->longwave, w->lw, T, N, cidx);
      scatter(o->vapor_pressure, w->svp_, T, N, cidx); scatter(o->rel_humid, w->rh, T, N, cidx);
      scatter(o->air_pressure, w->ap, T, N, cidx); scatter(o->spec_humid, w->sh, T, N, cidx);
      scatter(o->tskc, w->stskc, T, N, cidx);
      if (j->wind) scatter(o->wind, w->swind, T, N, cidx);
    }
  }
  free(b_tmin); free(b_tmax); free(b_prec); free(b_wind);
  free(b_psw); free(b_pvp); free(b_ptk); free(b_plw);
  ws_free(w);
  return NULL;
}

int mso_run(const mso_params *p, int n_cells, int n_days, const double *lat, const double *lon,
            const double *elev, const int *doy, const int *month, const double *t_min,
            const double *t_max, const double *prec, const double *wind, const double *st_tmin,
            const double *st_tmax, const double *st_prec, const double *t_pk12,
            const double *dur12, const double *t_end, const mso_outputs *out, int n_threads) {
  return mso_run_pt(p, NULL, n_cells, n_days, lat, lon, elev, doy, month, t_min, t_max, prec, wind,
                    st_tmin, st_tmax, st_prec, t_pk12, dur12, t_end, out, n_threads);
}

int mso_run_pt(const mso_params *p, const mso_passthrough *pt, int n_cells, int n_days,
               const double *lat, const double *lon, const double *elev, const int *doy,
               const int *month, const double *t_min, const double *t_max, const double *prec,
               const double *wind, const double *st_tmin, const double *st_tmax,
               const double *st_prec, const double *t_pk12, const double *dur12,
               const double *t_end, const mso_outputs *out, int n_threads) {
  if (!p || n_cells <= 0 || n_days < 2 || !out) return -3;
  if (p->time_step <= 0 || MIN_PER_DAY % p->time_step) return -3;
  if (p->prec_type != MSO_PREC_UNIFORM && (!t_pk12 || !dur12)) return -3;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_cells) n_threads = n_cells;
  job *jobs = (job *)calloc(n_threads, sizeof(job));
  pthread_t *th = (pthread_t *)calloc(n_threads, sizeof(pthread_t));
  int per = (n_cells + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    job *j = &jobs[t];
    j->p = p; j->n_cells = n_cells; j->n_days = n_days;
    j->c0 = t * per; j->c1 = (t + 1) * per > n_cells ? n_cells : (t + 1) * per;
    if (j->c0 > j->c1) j->c0 = j->c1;
    j->lat = lat; j->lon = lon; j->elev = elev; j->doy = doy; j->month = month;
    j->t_min = t_min; j->t_max = t_max; j->prec = prec; j->wind = wind;
    j->st_tmin = st_tmin; j->st_tmax = st_tmax; j->st_prec = st_prec;
    j->t_pk12 = t_pk12; j->dur12 = dur12; j->t_end = t_end; j->out = out; j->pt = pt;
    if (n_threads == 1) worker(j);
    else pthread_create(&th[t], NULL, worker, j);
  }
  int rc = 0;
  for (int t = 0; t < n_threads; ++t) {
    if (n_threads > 1) pthread_join(th[t], NULL);
    if (jobs[t].rc) rc = jobs[t].rc;
  }
  free(jobs); free(th);
  return rc;
}
