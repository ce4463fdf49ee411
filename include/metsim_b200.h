/* metsim_b200 -- C ABI of the B200-native MTCLIM hot path.
 *
 * This is the drop-in boundary: every entry point below replaces a piece of the reference's
 * per-cell Python path (citations are path:line in the UW-Hydro/MetSim checkout).  Signatures are
 * plain C: pointers, sizes, POD structs.  No torch / pandas / xarray types cross this line.
 *
 * Conventions
 *   - every array is cell-fastest: daily [D][ld], sub-daily [D*tpd][ld], state [90][ld],
 *     monthly [12][ld]; "ld" (leading dimension, in elements) >= n_cells.
 *   - pointers inside msb_forcing / msb_daily / msb_subdaily / msb_solar_tables are DEVICE
 *     pointers owned by the caller; `stream` is a cudaStream_t passed as void* (0 = default
 *     stream).  Those entry points allocate nothing, keep no state between calls and only enqueue
 *     work on `stream`: they are thread-safe per stream.  msb_run_host_ctx() takes HOST buffers and
 *     uses the device arena of a caller-owned msb_host_ctx.
 *   - every function returns 0 on success, <0 on error (MSB_E_*); msb_last_error() gives the text.
 *     Errors mirror the reference's exceptions (e.g. t_max < t_min -> MSB_E_DTR_NEGATIVE,
 *     metsim/metsim.py:612-614).  There is no CPU fallback anywhere in this library.
 */
#ifndef METSIM_B200_H
#define METSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSB_VERSION 201

enum {
  MSB_OK = 0,
  MSB_E_ARG = -1,          /* bad argument / unsupported option */
  MSB_E_CUDA = -2,         /* CUDA runtime error */
  MSB_E_DTR_NEGATIVE = -3, /* t_max < t_min somewhere (ValueError, metsim.py:612-614) */
  MSB_E_NOT_CONVERGED = -4,/* tdew fixed point exceeded max_iter (the reference would spin) */
  MSB_E_KNOTS = -5,        /* PCHIP knots not strictly increasing (scipy raises ValueError) */
  MSB_E_NOMEM = -6
};

enum { MSB_PREC_UNIFORM = 0, MSB_PREC_TRIANGLE = 1, MSB_PREC_MIX = 2 };
enum { MSB_LW_DEFAULT = 0, MSB_LW_TVA = 1, MSB_LW_ANDERSON = 2, MSB_LW_BRUTSAERT = 3,
       MSB_LW_SATTERLUND = 4, MSB_LW_IDSO = 5, MSB_LW_PRATA = 6 };
enum { MSB_CLOUD_DEFAULT = 0, MSB_CLOUD_DEARDORFF = 1 };
/* MetSim.methods registry keys (metsim/metsim.py:139): the iterated MTCLIM daily pass, or
 * 'passthrough' (metsim/methods/passthrough.py:28-46: daily shortwave supplied, one evaluation) */
enum { MSB_METHOD_MTCLIM = 0, MSB_METHOD_PASSTHROUGH = 1 };

/* sub-daily output variables, metsim/metsim.py:660-666 (out_var_check) */
enum {
  MSB_V_TEMP = 0, MSB_V_PREC, MSB_V_SHORTWAVE, MSB_V_LONGWAVE, MSB_V_VAPOR_PRESSURE,
  MSB_V_REL_HUMID, MSB_V_AIR_PRESSURE, MSB_V_SPEC_HUMID, MSB_V_TSKC, MSB_V_WIND,
  MSB_N_SUBVARS
};
/* daily variables produced by the MTCLIM pass, metsim/methods/mtclim.py:50-79 (+ driver columns) */
enum {
  MSB_D_T_DAY = 0, MSB_D_TFMAX, MSB_D_TSKC, MSB_D_TDEW, MSB_D_VAPOR_PRESSURE, MSB_D_SHORTWAVE,
  MSB_D_PET, MSB_D_DAYLENGTH, MSB_D_POTRAD, MSB_D_TT_MAX, MSB_D_SEASONAL_PREC,
  MSB_D_SMOOTHED_DTR,
  MSB_D_SHORTWAVE_DAILY,   /* daily-output mode: shortwave * daylength / 86400, metsim/metsim.py:779-781 */
  MSB_N_DAILYVARS
};

/* the keys of MetSim.params that the hot path reads (metsim/metsim.py:140-171) */
typedef struct msb_params {
  int32_t time_step;   /* minutes; divides 1440; 1440 = daily mode */
  int32_t prec_type;   /* MSB_PREC_* */
  int32_t utc_offset;  /* bool */
  int32_t lw_type;     /* MSB_LW_* */
  int32_t lw_cloud;    /* MSB_CLOUD_* */
  int32_t max_iter;    /* cap on tdew evaluations (<= 64 supported per call chain) */
  int32_t method;      /* MSB_METHOD_* */
  int32_t sw_daylight; /* passthrough only: params['sw_averaging'] == 'daylight'; otherwise the
                          supplied shortwave is a 24 h mean (disaggregate.py:78-80) */
  double sw_prec_thresh, rain_scalar, tdew_tol, tmax_daylength_fraction, tday_coef, lapse_rate;
} msb_params;

/* Launch-geometry / path overrides, embedded in msb_daily, msb_subdaily and msb_host_run.
 * Every field: 0 = let the library choose (what bench.py times on the full grids).  On a small
 * shard the library shortens the work per thread so that the grid still fills the GPU; the parity
 * tests pin the values of the benchmarked configuration instead, so that the oracle comparisons
 * run exactly the geometry the performance number is earned with.  Nothing here changes results
 * beyond floating-point summation order (daily_chunk_len: partial sums of the tdew stop rule). */
typedef struct msb_tuning {
  int32_t daily_chunk_len;  /* days per thread of the daily pass */
  int32_t daily_two_pass;   /* != 0: two-pass daily form even when candidate planes are supplied */
  int32_t knot_tile;        /* knot-days per thread of the interior thermodynamic kernel, used as given */
  int32_t day_tile;         /* output days per thread of the interior shortwave / precipitation kernels, used as given */
  int32_t disagg_path;      /* 0 = specialised interior-day kernels + event kernel on the edge days,
                               1 = event-driven kernel for every day */
  int32_t day_prefetch;     /* shortwave kernel, L1 prefetch of table entries: 0 = the next segment's (default),
                               1 = the next day's, -1 = none */
  int32_t prep_days;        /* padded days per thread of the day-record kernel */
  int32_t sw_table;         /* msb_run_host only (elsewhere the caller decides by what it puts into
                               msb_solar_tables.q_win): 0 = window-major copy of the table when
                               neighbouring cells are far apart in longitude, 1 = never, 2 = always */
} msb_tuning;

/* Geometry of the per-latitude solar tables (replaces the per-cell 366x2880 tiny_rad_fract of
 * physics.py:177-285: only tt_max0 depends on elevation, everything else on latitude alone). */
#define MSB_TINY_PER_DAY 2880
#define MSB_Q_LD 2888        /* row pitch of the signed prefix table: entries 0..2880, [2881]=row total */
#define MSB_Q_PAD 8192       /* zeroed elements after the last row (the kernel may read one step past a row) */
#define MSB_TT_TERMS 28      /* Taylor terms of tt_max0 in the pressure ratio */
#define MSB_TT_P0 0.68       /* expansion point of the pressure ratio p = t1**t2 */
#define MSB_STATE_DAYS 90

typedef struct msb_solar_tables {
  int32_t n_lat;        /* number of unique latitude rows */
  double *daylength;    /* [n_lat][366]  s            physics.py:236 */
  double *potrad;       /* [n_lat][366]  W m-2        physics.py:276 */
  double *rise_min;     /* [n_lat][366]  minutes, 240 default applied  disaggregate.py:163-182 */
  double *set_min;      /* [n_lat][366]  minutes, 960 default applied */
  double *q;            /* [n_lat][365][MSB_Q_LD] (+ MSB_Q_PAD) signed prefix sums of the normalised row */
  double *tt_coef;      /* [n_lat][365][MSB_TT_TERMS] Taylor coefficients of tt_max0(p) */
  /* Optional second copy of q for ONE time step (NULL = none), [n_lat][365][msb_solar_q_win_ld(step)]
   * (+ MSB_Q_PAD): the entries regrouped by their phase inside an output window of C = 2*step
   * tiny steps -- tiny step s sits at [(s % C) * W + s / C], the row total at [C * W], with
   * W = 1440/step + 1 rounded up to a multiple of 4 (whole 32-byte sectors per phase).  The windows of an output day read every C-th entry of a row
   * (disaggregate.py:660-667): here those are CONSECUTIVE, so a cell's day touches ~7 sectors
   * instead of 25.  It pays where neighbouring cells do not share table entries anyway (coarse or
   * sparsely masked grids: global 0.5 deg, masked 1/8 deg); on dense fine grids (CONUS 1/16 deg:
   * 16 consecutive entries per warp load) the plain table coalesces better and q_win stays NULL. */
  double *q_win;
  int32_t q_win_step;   /* minutes; the shortwave kernel uses q_win only when it equals params.time_step */
  int32_t reserved_;
} msb_solar_tables;

/* inputs of one shard (contiguous chunk of active cells), all device pointers */
typedef struct msb_forcing {
  int32_t n_cells, n_days;
  int64_t ld;              /* leading dimension of every [*][ld] array */
  const double *lon, *elev;          /* [n_cells]; lon in degrees, any range */
  const int32_t *lat_row;            /* [n_cells] index into the solar tables */
  const int32_t *doy, *month;        /* [n_days] 1-based day of year / month */
  const double *t_min, *t_max, *prec;/* [n_days][ld] */
  const double *wind;                /* [n_days][ld] or NULL */
  const double *st_t_min, *st_t_max, *st_prec; /* [90][ld] state, metsim.py:589-618 */
  const double *t_pk12, *dur12;      /* [12][ld] or NULL (needed for triangle / mix) */
  const double *t_end;               /* [2][ld] or NULL  (disaggregate.py:250-253) */
  /* Optional columns the reference driver hands to method.run(df, params) ready-made
   * (metsim.py:736-739, :601-616).  When non-NULL the daily pass uses them instead of deriving
   * them from the state / solar tables -- this is the exact contract of the per-cell callable. */
  const double *given_seasonal_prec, *given_smoothed_dtr;   /* [n_days][ld] or NULL (both) */
  const double *given_daylength, *given_potrad, *given_tt_max; /* [n_days][ld] or NULL (all three) */
  /* 'passthrough' method (methods/passthrough.py:28-46): forcing columns that replace computed
   * ones.  shortwave is required by that method; the others are optional.  longwave rescales the
   * sub-daily longwave to the supplied daily mean (disaggregate.py:115-118, :583-589). */
  const double *given_shortwave, *given_vapor_pressure, *given_tskc, *given_longwave; /* [n_days][ld] */
} msb_forcing;

#define MSB_DAILY_CAND 3      /* candidate evaluation counts kept by the single-pass daily form */
#define MSB_DAILY_CAND_K0 4   /* ... after 4, 5 and 6 evaluations of the tdew map */
typedef struct msb_daily {
  double *var[MSB_N_DAILYVARS];      /* [n_days][ld] each; TSKC, VAPOR_PRESSURE, SHORTWAVE are
                                        required by the disaggregation, the rest may be NULL
                                        (VAPOR_PRESSURE / SHORTWAVE may be NULL when `cand` is given) */
  int32_t *n_iter;                   /* [n_cells] tdew evaluations used (0 = not converged) */
  double *partial;                   /* scratch, msb_daily_scratch_bytes() */
  int32_t *status;                   /* [4] device error flags (dtr<0, knots, ...) */
  /* Single-pass form of the tdew fixed point (optional).  The number of evaluations n* of a cell is
   * only known once the RMS over its whole record is (mtclim.py:63).  With `cand` the daily kernel
   * runs the evaluations once, keeps vapour pressure and shortwave after MSB_DAILY_CAND_K0 ..
   * MSB_DAILY_CAND_K0 + MSB_DAILY_CAND - 1 evaluations in candidate planes and msb_disaggregate reads
   * the plane cand_sel[cell] = n* - MSB_DAILY_CAND_K0 picks; the (rare) cells outside that window are
   * finished by a compacted replay into plane 0.  Used when cand != NULL, the method is mtclim, no
   * given_* column replaces a computed one and no daily variable beyond TSKC is requested; otherwise
   * the two-pass form fills var[] as before.
   *   cand      [2*MSB_DAILY_CAND][n_days][ld]: plane 2j = vapour pressure, 2j+1 = shortwave
   *   cand_sel  [n_cells] plane index per cell (written by msb_mtclim_daily)
   *   cand_list [n_cells + 1] scratch: count, then the cells finished by the replay */
  double *cand;
  int32_t *cand_sel;
  int32_t *cand_list;
  msb_tuning tune;
} msb_daily;

typedef struct msb_subdaily {
  void *var[MSB_N_SUBVARS];          /* [n_steps][ld] each, NULL = not requested */
  double scale[MSB_N_SUBVARS];       /* out = value*scale + offset  (metsim/units.py:3-59) */
  double offset[MSB_N_SUBVARS];
  int32_t dtype;                     /* 4 = float32 (reference default out_precision), 8 = float64 */
  int64_t ld;
  int32_t day0, day1;                /* produce days [day0, day1); var[] points at step day0*tpd */
  void *workspace;                   /* device scratch of msb_disagg_workspace_bytes() bytes: the
                                        per-(cell, day) records of the days around [day0, day1) */
  size_t workspace_bytes;
  msb_tuning tune;
} msb_subdaily;

int msb_version(void);
const char *msb_last_error(void);
void msb_default_params(msb_params *p);                          /* metsim.py:140-171 */

/* ---- solar geometry: physics.py:177-285 + disaggregate.py:133-201 (rise/set) ---------------
 * Host prelude (glibc libm, decides the sunrise sign bit-for-bit like the reference) + table
 * kernel.  lat_deg is a HOST array of the unique latitudes; tables are DEVICE buffers sized by
 * msb_solar_table_bytes().  host_ws must hold msb_solar_prelude_bytes(n_lat) bytes (pinned
 * memory makes the upload asynchronous), dev_ws the same number of bytes on the device. */
size_t msb_solar_prelude_bytes(int n_lat);
int msb_solar_table_sizes(int n_lat, size_t *small_elems, size_t *q_elems, size_t *tt_elems);
/* row pitch / element count (pad included) of msb_solar_tables.q_win for a time step that divides
 * 1440 and is below it; 0 otherwise */
size_t msb_solar_q_win_ld(int time_step);
size_t msb_solar_q_win_elems(int n_lat, int time_step);
int msb_solar_tables_build(int n_lat, const double *lat_deg_host, void *host_ws, void *dev_ws,
                           const msb_solar_tables *tables, int n_host_threads, void *stream);

/* ---- daily MTCLIM: mtclim.py:28-80 (run), metsim.py:589-618 (rolling means) ---------------- */
size_t msb_daily_scratch_bytes(int n_cells, int n_days);
/* the same for a pinned msb_tuning.daily_chunk_len (0 = the library's choice, as above) */
size_t msb_daily_scratch_bytes_chunk(int n_cells, int n_days, int daily_chunk_len);
/* 1 when msb_mtclim_daily would take the single-pass form for these arguments, else 0 */
int msb_daily_single_pass(const msb_params *p, const msb_forcing *f, const msb_daily *out);
int msb_mtclim_daily(const msb_params *p, const msb_forcing *f, const msb_solar_tables *t,
                     const msb_daily *out, void *stream);
/* reads back status/n_iter flags (synchronises the stream); returns MSB_OK or the first error */
int msb_daily_check(const msb_forcing *f, const msb_daily *out, void *stream);

/* ---- sub-daily disaggregation: disaggregate.py:34-680, units.py:3-59 -----------------------
 * Days [2, D-2) of a record go through two specialised kernels that need no scratch.  The
 * record's first / last two days go through an
 * event-driven kernel that reads per-(cell, day) records (PCHIP knots with their derivatives, the
 * closed-form precipitation shape, radiation per tiny step) built by a first kernel in
 * caller-provided scratch.  Results do not depend on how a record is cut into calls. */
#define MSB_REC_FIELDS 9
#define MSB_REC_MARGIN 3
size_t msb_disagg_workspace_bytes(int64_t ld, int day0, int day1);
/* the same plus the float64 staging plane [(day1-day0)*tpd][ld] that the longwave rescale of the
 * passthrough method needs (forcing.given_longwave != NULL and longwave requested) */
size_t msb_disagg_workspace_bytes_lw(int64_t ld, int day0, int day1, int time_step);
int msb_disaggregate(const msb_params *p, const msb_forcing *f, const msb_solar_tables *t,
                     const msb_daily *daily, const msb_subdaily *out, void *stream);

/* accuracy probe of the library's fp64 divide / exp / sqrt sequences: out_dev is [12][n]
 * (1/x, fast exp(x), fast sqrt(x), 1.2345678901234567/x, svp(x) in kPa, raw reciprocal seed,
 * accurate exp(x), accurate sqrt(x), remainder(x, 360), and the second-order 1/x, sqrt(x), exp(x)
 * of the interior-day kernel).  Used by the parity tests. */
int msb_selftest_math(int n, const double *x_dev, double *out_dev, void *stream);

/* ---- whole path with HOST buffers (what a reference-side binding calls) ---------------------
 * Replaces the cell loop of MetSim.run_slice (metsim.py:477-493): unique-latitude tables,
 * H2D of the forcings, daily pass, time-tiled disaggregation with double-buffered D2H into
 * the caller's host arrays.  lat is per cell (degrees).  out_host[v] is [n_days*tpd][n_cells]
 * (dtype per out_dtype) or NULL.  daily_host[v] is [n_days][n_cells] float64 or NULL. */
typedef struct msb_host_run {
  int32_t n_cells, n_days, device;
  const double *lat, *lon, *elev;
  const int32_t *doy, *month;
  const double *t_min, *t_max, *prec, *wind;
  const double *st_t_min, *st_t_max, *st_prec;
  const double *t_pk12, *dur12, *t_end;
  const double *given_shortwave, *given_vapor_pressure, *given_tskc, *given_longwave; /* passthrough, or NULL */
  void *out_host[MSB_N_SUBVARS];
  double scale[MSB_N_SUBVARS], offset[MSB_N_SUBVARS];
  int32_t out_dtype;                 /* 4 or 8 */
  double *daily_host[MSB_N_DAILYVARS];
  int32_t *n_iter_host;              /* [n_cells] or NULL */
  int32_t tile_days;                 /* 0 = choose */
  int32_t out_is_ring;               /* !=0: out_host[v] holds only 2*tile_days*tpd steps and is
                                        reused as a ring (a streaming writer drains it) */
  /* Optional consumer hook (the streaming writer of metsim/metsim.py:363-454): called on the calling
   * thread once per output tile, in order, after the tile's device-to-host copies have completed,
   * while the NEXT tile is already computing / copying.  The tile holds days [day0, day1); with
   * out_is_ring its steps start at slot*tile_days*tpd of every out_host[v] (slot = tile & 1) and the
   * slot is not overwritten before the callback returns; without it slot = -1 and the steps are at
   * their final position day0*tpd.  A non-zero return aborts the run (MSB_E_ARG). */
  int32_t (*on_tile)(void *user, int32_t tile, int32_t day0, int32_t day1, int32_t slot);
  void *user;
  double ms_solar, ms_h2d, ms_daily, ms_disagg_total; /* filled on return (CUDA events) */
  msb_tuning tune;                   /* all zero = the library's choices */
} msb_host_run;

/* Caller-owned context of the host-buffer path: the device arena (grow-only, reused from call to
 * call), the pinned staging of the solar prelude, a compute and a copy stream and their events, all
 * on one device.  The library keeps NO process-wide mutable state: two contexts -- on the same
 * device or on different ones -- run concurrently from two threads (the reference drives its cell
 * loop from a thread pool, scheduler='threading', metsim/metsim.py:63,195); calls on ONE context
 * are serialised by it.  Every call returns with nothing of the context still in flight, on
 * success and on error alike. */
typedef struct msb_host_ctx msb_host_ctx;
int msb_host_ctx_create(int device, msb_host_ctx **ctx);
void msb_host_ctx_destroy(msb_host_ctx *ctx);
size_t msb_host_ctx_device_bytes(const msb_host_ctx *ctx);   /* current size of the device arena */
/* workspace-size query: device / pinned bytes a context holds after running `run` */
int msb_run_host_bytes(const msb_params *p, const msb_host_run *run, size_t *device_bytes,
                       size_t *pinned_bytes);
int msb_run_host_ctx(msb_host_ctx *ctx, const msb_params *p, msb_host_run *run);
/* convenience: a context created for this one call on run->device and destroyed afterwards */
int msb_run_host(const msb_params *p, msb_host_run *run);

/* ---- native consumer of output tiles (host threads; metsim/metsim.py:363-454, :581-586) ------
 * The reference scatters every cell's series into NaN-filled (time, lat, lon) arrays and writes them
 * through xarray / netCDF4.  A sink does that for the rows of one output variable as they arrive in
 * host memory from msb_run_host_ctx()'s on_tile hook: scatter [step][cell] -> [step][n_grid] (NaN where
 * no cell maps), optional byte swap to the big-endian order of the classic NetCDF format, pwrite()
 * at the given file offset -- on a pool of threads, each with its own staging buffer, so that a tile
 * is consumed at memory speed while the next one is being copied.  Host code only. */
typedef struct msb_sink msb_sink;
typedef struct msb_sink_desc {
  int32_t n_cells;            /* cells per row of the source */
  int32_t elem_bytes;         /* 4 (float32) or 8 (float64) */
  int64_t n_grid;             /* lat * lon positions per step in the file, >= n_cells */
  const int64_t *cell_index;  /* [n_cells] grid position of every cell; NULL = identity (n_grid == n_cells) */
  int32_t byteswap;           /* != 0: big-endian in the file */
  int32_t n_threads;          /* 0 = all hardware threads */
  int32_t fd;                 /* open file descriptor to pwrite() to; -1 = no file (scatter + swap only) */
  int32_t reserved;
} msb_sink_desc;
int msb_sink_create(const msb_sink_desc *desc, msb_sink **sink);
/* rows [0, n_steps) of src ([n_steps][n_cells]) -> file_offset (bytes; where the first row's grid
 * starts).  Returns when everything has been handed to the kernel's page cache. */
int msb_sink_write(msb_sink *sink, const void *src, int64_t n_steps, int64_t file_offset);
uint64_t msb_sink_checksum(const msb_sink *sink);      /* sum of the raw bit patterns consumed so far */
uint64_t msb_sink_bytes_written(const msb_sink *sink);
void msb_sink_destroy(msb_sink *sink);

#ifdef __cplusplus
}
#endif
#endif
