// C ABI plumbing: errors, defaults, and the host-buffer entry points msb_run_host_ctx() /
// msb_run_host().
//
// msb_run_host_ctx() replaces the per-cell loop of MetSim.run_slice (metsim/metsim.py:477-493 ->
// wrap_run_cell :693-796) for one shard of cells: it deduplicates latitudes, uploads the
// forcings, runs solar tables -> daily MTCLIM -> tiled disaggregation and streams every output
// tile back to host memory while the next tile is being computed (two device tile buffers, a
// compute stream and a copy stream; nothing larger than two tiles of outputs ever lives in HBM).
//
// There is no process-wide mutable state: everything a run needs between calls (device arena,
// pinned staging, the two streams, events) belongs to a caller-owned msb_host_ctx.  The reference
// calls its per-cell path from concurrent threads (scheduler='threading', metsim/metsim.py:63,
// :195): give each thread its own context and the runs overlap, on one device or on several.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <unordered_map>
#include <vector>

#include "msb_common.cuh"

namespace msb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return MSB_E_CUDA;
}

struct Carver {
  char *base;
  size_t off = 0;
  explicit Carver(void *b) : base(reinterpret_cast<char *>(b)) {}
  template <typename T>
  T *take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T *p = reinterpret_cast<T *>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

}  // namespace msb

// grow-only device / pinned arenas, streams and events of one caller-owned context
struct msb_host_ctx {
  int device = -1;
  void *dev = nullptr;
  size_t dev_bytes = 0;
  void *pin = nullptr;
  size_t pin_bytes = 0;
  cudaStream_t s_main = nullptr, s_copy = nullptr;
  cudaEvent_t ev[6] = {}, ev_k[2] = {}, ev_c[2] = {};
  std::mutex mu;   // one run at a time per context
};

using namespace msb;

static int ctx_reserve(msb_host_ctx &a, size_t dev_bytes, size_t pin_bytes) {
  if (a.dev_bytes < dev_bytes) {
    if (a.dev) cudaFree(a.dev);
    a.dev = nullptr;
    a.dev_bytes = 0;
    cudaError_t e = cudaMalloc(&a.dev, dev_bytes);
    if (e != cudaSuccess) {
      set_error("cudaMalloc of %zu bytes failed: %s", dev_bytes, cudaGetErrorString(e));
      cudaGetLastError();
      return MSB_E_NOMEM;
    }
    a.dev_bytes = dev_bytes;
  }
  if (a.pin_bytes < pin_bytes) {
    if (a.pin) cudaFreeHost(a.pin);
    a.pin = nullptr;
    a.pin_bytes = 0;
    cudaError_t e = cudaMallocHost(&a.pin, pin_bytes);
    if (e != cudaSuccess) {
      set_error("cudaMallocHost of %zu bytes failed: %s", pin_bytes, cudaGetErrorString(e));
      cudaGetLastError();
      return MSB_E_NOMEM;
    }
    a.pin_bytes = pin_bytes;
  }
  return MSB_OK;
}

extern "C" int msb_version(void) { return MSB_VERSION; }
extern "C" const char *msb_last_error(void) { return g_err; }

extern "C" void msb_default_params(msb_params *p) {
  if (!p) return;
  p->time_step = 60;
  p->prec_type = MSB_PREC_UNIFORM;
  p->utc_offset = 0;
  p->lw_type = MSB_LW_PRATA;
  p->lw_cloud = MSB_CLOUD_DEARDORFF;
  p->max_iter = 24;
  p->method = MSB_METHOD_MTCLIM;
  p->sw_daylight = 0;
  p->sw_prec_thresh = 0.0;
  p->rain_scalar = 0.75;
  p->tdew_tol = 1e-6;
  p->tmax_daylength_fraction = 0.67;
  p->tday_coef = 0.45;
  p->lapse_rate = 0.0065;
}

extern "C" int msb_host_ctx_create(int device, msb_host_ctx **out) {
  if (!out) { set_error("msb_host_ctx_create: NULL argument"); return MSB_E_ARG; }
  *out = nullptr;
  MSB_CUDA(cudaSetDevice(device));
  msb_host_ctx *c = new (std::nothrow) msb_host_ctx;
  if (!c) { set_error("msb_host_ctx_create: out of host memory"); return MSB_E_NOMEM; }
  c->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&c->s_main, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking);
  for (auto &x : c->ev) if (e == cudaSuccess) e = cudaEventCreate(&x);
  for (auto &x : c->ev_k) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
  // "tile copied": the calling thread WAITS on these between tiles; a blocking wait leaves its core to
  // the consumer's worker threads instead of spinning on it
  for (auto &x : c->ev_c)
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&x, cudaEventDisableTiming | cudaEventBlockingSync);
  if (e != cudaSuccess) {
    int rc = cuda_fail(e, "msb_host_ctx_create");
    msb_host_ctx_destroy(c);
    return rc;
  }
  *out = c;
  return MSB_OK;
}

extern "C" void msb_host_ctx_destroy(msb_host_ctx *c) {
  if (!c) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  // nothing of this context may still be running when its buffers go away
  if (c->s_main) cudaStreamSynchronize(c->s_main);
  if (c->s_copy) cudaStreamSynchronize(c->s_copy);
  for (auto &x : c->ev) if (x) cudaEventDestroy(x);
  for (auto &x : c->ev_k) if (x) cudaEventDestroy(x);
  for (auto &x : c->ev_c) if (x) cudaEventDestroy(x);
  if (c->s_main) cudaStreamDestroy(c->s_main);
  if (c->s_copy) cudaStreamDestroy(c->s_copy);
  if (c->dev) cudaFree(c->dev);
  if (c->pin) cudaFreeHost(c->pin);
  cudaGetLastError();
  if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
  delete c;
}

extern "C" size_t msb_host_ctx_device_bytes(const msb_host_ctx *c) { return c ? c->dev_bytes : 0; }

namespace {

// what one run carves out of the context's device arena
struct Plan {
  int N, D, ts, tpd, osz, n_lat, n_out, tile_days;
  bool sub, single_pass, lw_rescale;
  size_t dn, sn, small_e, q_e, qw_e, tt_e, prelude, ld_out, tile_elems, ws_bytes, scratch_bytes, need;
};

int check_run(const msb_params *p, const msb_host_run *run) {
  if (!p || !run) { set_error("msb_run_host: NULL argument"); return MSB_E_ARG; }
  if (run->n_cells <= 0 || run->n_days < 2) {
    set_error("msb_run_host: need n_cells > 0 and n_days >= 2");
    return MSB_E_ARG;
  }
  if (!run->lat || !run->lon || !run->elev || !run->doy || !run->month || !run->t_min ||
      !run->t_max || !run->prec || !run->st_t_min || !run->st_t_max || !run->st_prec) {
    set_error("msb_run_host: a required input pointer is NULL");
    return MSB_E_ARG;
  }
  const int ts = p->time_step;
  if (ts <= 0 || kMinPerDay % ts || (ts > 360 && ts != kMinPerDay)) {
    set_error("Time step must be evenly divisible into 1440 minutes (24 hours) and less than 360 "
              "minutes (6 hours). Got %d.", ts);
    return MSB_E_ARG;
  }
  if (run->out_dtype != 4 && run->out_dtype != 8) {
    set_error("msb_run_host: out_dtype must be 4 or 8");
    return MSB_E_ARG;
  }
  if (run->out_is_ring && run->tile_days <= 0) {
    set_error("msb_run_host: out_is_ring needs an explicit tile_days (the ring holds 2 tiles)");
    return MSB_E_ARG;
  }
  return MSB_OK;
}

// Window-major copy of the shortwave table (msb_solar_tables.q_win)?  tune.sw_table 0: when the 32
// consecutive cells of a warp touch more than 8 sectors of the plain table per load on average
// (distinct (latitude row, tiny offset / 4) among them) -- coarse or sparsely masked grids.
bool want_window_major(const msb_params *p, const msb_host_run *run, const std::vector<int32_t> &lat_row) {
  if (p->time_step >= kMinPerDay || run->tune.sw_table == 1) return false;
  if (run->tune.sw_table == 2) return true;
  const int N = run->n_cells;
  if (N < 64) return false;
  long sectors = 0, warps = 0;
  int64_t key[32];
  for (int w0 = 0; w0 + 32 <= N; w0 += 32, ++warps) {
    for (int l = 0; l < 32; ++l) {
      const double lon = std::remainder(run->lon[w0 + l], 360.0);
      key[l] = ((int64_t)lat_row[w0 + l] << 12) + (((int)(lon / 360.0 * kTiny) + kTiny / 2) >> 2);
    }
    std::sort(key, key + 32);
    sectors += std::unique(key, key + 32) - key;
  }
  return sectors > 8 * warps;
}

void make_plan(const msb_params *p, const msb_host_run *run, int n_lat, bool win, Plan &pl) {
  pl.N = run->n_cells; pl.D = run->n_days; pl.ts = p->time_step;
  pl.sub = pl.ts < kMinPerDay;
  pl.tpd = pl.sub ? kMinPerDay / pl.ts : 1;
  pl.osz = run->out_dtype == 8 ? 8 : 4;
  pl.n_lat = n_lat;
  const size_t N = pl.N, D = pl.D;
  pl.n_out = 0;
  for (int v = 0; v < MSB_N_SUBVARS; ++v) pl.n_out += (pl.sub && run->out_host[v]) ? 1 : 0;
  int tile_days = run->tile_days > 0 ? run->tile_days : 0;
  if (pl.sub && tile_days == 0) {
    // two tile buffers of all requested outputs within ~8 GB, at least 1 day
    size_t per_day = (size_t)pl.n_out * pl.tpd * N * pl.osz;
    tile_days = per_day ? (int)((4ull << 30) / per_day) : pl.D;
    if (tile_days < 1) tile_days = 1;
  }
  if (!pl.sub || tile_days > pl.D) tile_days = pl.D;
  pl.tile_days = tile_days;
  pl.dn = D * N; pl.sn = (size_t)kState * N;
  msb_solar_table_sizes(n_lat, &pl.small_e, &pl.q_e, &pl.tt_e);
  pl.qw_e = win ? msb_solar_q_win_elems(n_lat, pl.ts) : 0;
  pl.prelude = msb_solar_prelude_bytes(n_lat);
  // single-pass daily form (daily.cu): no daily column is wanted on the host, the method is mtclim
  // and the six candidate planes are affordable beside the inputs
  bool wants_daily = false;
  for (int k = 0; k < MSB_N_DAILYVARS; ++k) wants_daily |= run->daily_host[k] != nullptr;
  pl.single_pass = pl.sub && !wants_daily && p->method == MSB_METHOD_MTCLIM && !run->tune.daily_two_pass &&
                   2 * MSB_DAILY_CAND * pl.dn * 8 <= ((size_t)16 << 30);
  pl.scratch_bytes = msb_daily_scratch_bytes_chunk(pl.N, pl.D, run->tune.daily_chunk_len);
  // device output tiles with rows on 128-byte boundaries (leading dimension padded to 32 cells):
  // warp-wide stores then never share a sector with the neighbouring row
  pl.ld_out = (N + 31) / 32 * 32;
  pl.tile_elems = (size_t)tile_days * pl.tpd * pl.ld_out;
  pl.lw_rescale = p->method == MSB_METHOD_PASSTHROUGH && run->given_longwave && run->out_host[MSB_V_LONGWAVE];
  pl.ws_bytes = !pl.sub ? 0
                : pl.lw_rescale ? msb_disagg_workspace_bytes_lw((int64_t)pl.ld_out, 0, tile_days, pl.ts)
                                : msb_disagg_workspace_bytes(pl.N, 0, tile_days);
  size_t need = 0;
  auto add = [&](size_t bytes) { need = ((need + 255) & ~(size_t)255) + bytes; };
  add(3 * N * 8); add(N * 4); add(2 * D * 4);                    // lon/elev(+pad), lat_row, doy/month
  for (int k = 0; k < 4; ++k) add(pl.dn * 8);                    // t_min t_max prec wind
  const double *given_h[4] = {run->given_shortwave, run->given_vapor_pressure, run->given_tskc,
                              run->given_longwave};
  for (int k = 0; k < 4; ++k) if (given_h[k]) add(pl.dn * 8);    // passthrough columns
  for (int k = 0; k < 3; ++k) add(pl.sn * 8);
  add(12 * N * 8); add(12 * N * 8); add(2 * N * 8);
  for (int k = 0; k < 4; ++k) add(pl.small_e * 8);
  add(pl.q_e * 8); add(pl.qw_e * 8); add(pl.tt_e * 8); add(pl.prelude);
  if (pl.single_pass) {
    add(pl.dn * 8);                                              // tskc
    add(2 * MSB_DAILY_CAND * pl.dn * 8); add(N * 4); add((N + 1) * 4);
  } else {
    for (int k = 0; k < MSB_N_DAILYVARS; ++k) add(pl.dn * 8);
  }
  add(N * 4); add(pl.scratch_bytes + 8); add(64);
  for (int k = 0; k < 2 * pl.n_out; ++k) add(pl.tile_elems * pl.osz);
  add(pl.ws_bytes);
  pl.need = need + 65536;
}

// unique latitudes (exact bit equality), run-length first: grids are latitude-row-major
void unique_lats(const double *lat, int N, std::vector<int32_t> &lat_row, std::vector<double> &ulat) {
  lat_row.resize(N);
  std::unordered_map<uint64_t, int32_t> seen;
  uint64_t prev_bits = 0;
  int32_t prev_row = -1;
  for (int i = 0; i < N; ++i) {
    uint64_t bits;
    std::memcpy(&bits, &lat[i], 8);
    if (prev_row >= 0 && bits == prev_bits) { lat_row[i] = prev_row; continue; }
    auto it = seen.find(bits);
    if (it == seen.end()) {
      it = seen.emplace(bits, (int32_t)ulat.size()).first;
      ulat.push_back(lat[i]);
    }
    prev_bits = bits;
    prev_row = it->second;
    lat_row[i] = prev_row;
  }
}

}  // namespace

// workspace-size query: what a context will hold on the device / in pinned host memory after this run
extern "C" int msb_run_host_bytes(const msb_params *p, const msb_host_run *run, size_t *device_bytes,
                                  size_t *pinned_bytes) {
  int rc = check_run(p, run);
  if (rc) return rc;
  std::vector<int32_t> lat_row;
  std::vector<double> ulat;
  unique_lats(run->lat, run->n_cells, lat_row, ulat);
  Plan pl;
  make_plan(p, run, (int)ulat.size(), want_window_major(p, run, lat_row), pl);
  if (device_bytes) *device_bytes = pl.need;
  if (pinned_bytes) *pinned_bytes = pl.prelude + 4096;
  return MSB_OK;
}

static int run_locked(msb_host_ctx *ctx, const msb_params *p, msb_host_run *run) {
  const int N = run->n_cells, D = run->n_days;
  std::vector<int32_t> lat_row;
  std::vector<double> ulat;
  unique_lats(run->lat, N, lat_row, ulat);
  const int n_lat = (int)ulat.size();
  Plan pl;
  make_plan(p, run, n_lat, want_window_major(p, run, lat_row), pl);
  const int tpd = pl.tpd, osz = pl.osz, tile_days = pl.tile_days, n_out = pl.n_out;
  const bool sub = pl.sub;
  const size_t dn = pl.dn, sn = pl.sn, ld_out = pl.ld_out, tile_elems = pl.tile_elems, ws_bytes = pl.ws_bytes;

  int rc = ctx_reserve(*ctx, pl.need, pl.prelude + 4096);
  if (rc) return rc;
  Carver cv(ctx->dev);
  double *d_lon = cv.take<double>(N), *d_elev = cv.take<double>(N);
  int32_t *d_row = cv.take<int32_t>(N), *d_doy = cv.take<int32_t>(D), *d_month = cv.take<int32_t>(D);
  double *d_tmin = cv.take<double>(dn), *d_tmax = cv.take<double>(dn), *d_prec = cv.take<double>(dn);
  double *d_wind = run->wind ? cv.take<double>(dn) : nullptr;
  const double *given_h[4] = {run->given_shortwave, run->given_vapor_pressure, run->given_tskc,
                              run->given_longwave};
  double *d_given[4];
  for (int k = 0; k < 4; ++k) d_given[k] = given_h[k] ? cv.take<double>(dn) : nullptr;
  double *d_stmin = cv.take<double>(sn), *d_stmax = cv.take<double>(sn), *d_stprec = cv.take<double>(sn);
  double *d_tpk = run->t_pk12 ? cv.take<double>(12 * (size_t)N) : nullptr;
  double *d_dur = run->dur12 ? cv.take<double>(12 * (size_t)N) : nullptr;
  double *d_tend = run->t_end ? cv.take<double>(2 * (size_t)N) : nullptr;
  msb_solar_tables tb;
  std::memset(&tb, 0, sizeof(tb));
  tb.n_lat = n_lat;
  tb.daylength = cv.take<double>(pl.small_e); tb.potrad = cv.take<double>(pl.small_e);
  tb.rise_min = cv.take<double>(pl.small_e); tb.set_min = cv.take<double>(pl.small_e);
  tb.q = cv.take<double>(pl.q_e);
  if (pl.qw_e) { tb.q_win = cv.take<double>(pl.qw_e); tb.q_win_step = pl.ts; }
  tb.tt_coef = cv.take<double>(pl.tt_e);
  void *d_prelude = cv.take<char>(pl.prelude);
  msb_daily dl;
  std::memset(&dl, 0, sizeof(dl));
  dl.tune = run->tune;
  if (pl.single_pass) {
    dl.var[MSB_D_TSKC] = cv.take<double>(dn);
    dl.cand = cv.take<double>(2 * MSB_DAILY_CAND * dn);
    dl.cand_sel = cv.take<int32_t>(N);
    dl.cand_list = cv.take<int32_t>((size_t)N + 1);
  } else {
    dl.tune.daily_two_pass = 1;
    for (int k = 0; k < MSB_N_DAILYVARS; ++k) {
      bool needed = k == MSB_D_TSKC || k == MSB_D_VAPOR_PRESSURE || k == MSB_D_SHORTWAVE;
      if (needed || run->daily_host[k]) dl.var[k] = cv.take<double>(dn);
    }
  }
  dl.n_iter = cv.take<int32_t>(N);
  dl.partial = cv.take<double>(pl.scratch_bytes / 8 + 1);
  dl.status = cv.take<int32_t>(16);
  void *d_ws = ws_bytes ? (void *)cv.take<char>(ws_bytes) : nullptr;
  void *d_out[2][MSB_N_SUBVARS];
  for (int b = 0; b < 2; ++b)
    for (int v = 0; v < MSB_N_SUBVARS; ++v)
      d_out[b][v] = (sub && run->out_host[v]) ? (void *)cv.take<char>(tile_elems * osz) : nullptr;

  cudaStream_t s_main = ctx->s_main, s_copy = ctx->s_copy;
  cudaEvent_t *ev = ctx->ev, *ev_k = ctx->ev_k, *ev_c = ctx->ev_c;
  // Error exits: kernels and device-to-host copies may still be in flight -- into the caller's
  // (possibly pinned) buffers and into this context's arena.  Nothing returns before they are done.
  auto quiesce = [&]() {
    cudaStreamSynchronize(s_main);
    cudaStreamSynchronize(s_copy);
    cudaGetLastError();
  };
#define MSB_TRY(expr)                               \
  do {                                              \
    int _rc = (expr);                               \
    if (_rc != MSB_OK) { quiesce(); return _rc; }   \
  } while (0)
#define MSB_TRYC(expr)                                                          \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) { int _rc = cuda_fail(_e, #expr); quiesce(); return _rc; } \
  } while (0)

  // ---- uploads (stream-ordered; pinned sources make them true DMA) ----
  MSB_TRYC(cudaEventRecord(ev[0], s_main));
  auto up = [&](void *dst, const void *src, size_t bytes) {
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s_main);
  };
  MSB_TRYC(up(d_lon, run->lon, (size_t)N * 8));
  MSB_TRYC(up(d_elev, run->elev, (size_t)N * 8));
  MSB_TRYC(up(d_row, lat_row.data(), (size_t)N * 4));
  MSB_TRYC(up(d_doy, run->doy, (size_t)D * 4));
  MSB_TRYC(up(d_month, run->month, (size_t)D * 4));
  MSB_TRYC(up(d_tmin, run->t_min, dn * 8));
  MSB_TRYC(up(d_tmax, run->t_max, dn * 8));
  MSB_TRYC(up(d_prec, run->prec, dn * 8));
  if (d_wind) MSB_TRYC(up(d_wind, run->wind, dn * 8));
  for (int k = 0; k < 4; ++k) if (d_given[k]) MSB_TRYC(up(d_given[k], given_h[k], dn * 8));
  MSB_TRYC(up(d_stmin, run->st_t_min, sn * 8));
  MSB_TRYC(up(d_stmax, run->st_t_max, sn * 8));
  MSB_TRYC(up(d_stprec, run->st_prec, sn * 8));
  if (d_tpk) MSB_TRYC(up(d_tpk, run->t_pk12, 12 * (size_t)N * 8));
  if (d_dur) MSB_TRYC(up(d_dur, run->dur12, 12 * (size_t)N * 8));
  if (d_tend) MSB_TRYC(up(d_tend, run->t_end, 2 * (size_t)N * 8));
  MSB_TRYC(cudaEventRecord(ev[1], s_main));

  // ---- solar tables (host prelude overlaps the uploads) ----
  unsigned hw = std::thread::hardware_concurrency();
  MSB_TRY(msb_solar_tables_build(n_lat, ulat.data(), ctx->pin, d_prelude, &tb,
                                 (int)(hw ? hw : 1), s_main));
  MSB_TRYC(cudaEventRecord(ev[2], s_main));

  msb_forcing f;
  std::memset(&f, 0, sizeof(f));
  f.n_cells = N; f.n_days = D; f.ld = N;
  f.lon = d_lon; f.elev = d_elev; f.lat_row = d_row; f.doy = d_doy; f.month = d_month;
  f.t_min = d_tmin; f.t_max = d_tmax; f.prec = d_prec; f.wind = d_wind;
  f.st_t_min = d_stmin; f.st_t_max = d_stmax; f.st_prec = d_stprec;
  f.t_pk12 = d_tpk; f.dur12 = d_dur; f.t_end = d_tend;
  f.given_shortwave = d_given[0]; f.given_vapor_pressure = d_given[1];
  f.given_tskc = d_given[2]; f.given_longwave = d_given[3];

  MSB_TRY(msb_mtclim_daily(p, &f, &tb, &dl, s_main));
  MSB_TRYC(cudaEventRecord(ev[3], s_main));

  // daily outputs back to the host (copy stream, after the daily pass)
  MSB_TRYC(cudaStreamWaitEvent(s_copy, ev[3], 0));
  for (int k = 0; k < MSB_N_DAILYVARS; ++k)
    if (run->daily_host[k])
      MSB_TRYC(cudaMemcpyAsync(run->daily_host[k], dl.var[k], dn * 8, cudaMemcpyDeviceToHost, s_copy));
  if (run->n_iter_host)
    MSB_TRYC(cudaMemcpyAsync(run->n_iter_host, dl.n_iter, (size_t)N * 4, cudaMemcpyDeviceToHost, s_copy));

  // ---- tiled disaggregation with double-buffered D2H ----
  if (sub && n_out > 0) {
    int n_tiles = (D + tile_days - 1) / tile_days;
    // hand a finished tile to the consumer: its D2H is complete, the next tile is in flight
    auto deliver = [&](int k) -> int {
      if (!run->on_tile) return MSB_OK;
      cudaError_t e = cudaEventSynchronize(ev_c[k & 1]);
      if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize(tile copied)");
      const int d0 = k * tile_days, d1 = d0 + tile_days < D ? d0 + tile_days : D;
      if (run->on_tile(run->user, k, d0, d1, run->out_is_ring ? (k & 1) : -1) != 0) {
        set_error("msb_run_host: on_tile callback failed at tile %d", k);
        return MSB_E_ARG;
      }
      return MSB_OK;
    };
    for (int k = 0; k < n_tiles; ++k) {
      const int b = k & 1;
      const int d0 = k * tile_days, d1 = d0 + tile_days < D ? d0 + tile_days : D;
      if (k >= 2) MSB_TRYC(cudaStreamWaitEvent(s_main, ev_c[b], 0));  // buffer b drained
      msb_subdaily so;
      std::memset(&so, 0, sizeof(so));
      for (int v = 0; v < MSB_N_SUBVARS; ++v) {
        so.var[v] = d_out[b][v];
        so.scale[v] = run->scale[v];
        so.offset[v] = run->offset[v];
      }
      so.dtype = osz; so.ld = (int64_t)ld_out; so.day0 = d0; so.day1 = d1;
      so.workspace = d_ws; so.workspace_bytes = ws_bytes;
      so.tune = run->tune;
      MSB_TRY(msb_disaggregate(p, &f, &tb, &dl, &so, s_main));
      MSB_TRYC(cudaEventRecord(ev_k[b], s_main));
      MSB_TRYC(cudaStreamWaitEvent(s_copy, ev_k[b], 0));
      const size_t steps = (size_t)(d1 - d0) * tpd;
      for (int v = 0; v < MSB_N_SUBVARS; ++v) {
        if (!d_out[b][v]) continue;
        size_t host_step0 = run->out_is_ring ? (size_t)b * tile_days * tpd : (size_t)d0 * tpd;
        char *dst = reinterpret_cast<char *>(run->out_host[v]) + host_step0 * (size_t)N * osz;
        if (ld_out == (size_t)N)     // rows already contiguous: one linear copy per variable
          MSB_TRYC(cudaMemcpyAsync(dst, d_out[b][v], steps * (size_t)N * osz, cudaMemcpyDeviceToHost, s_copy));
        else
          MSB_TRYC(cudaMemcpy2DAsync(dst, (size_t)N * osz, d_out[b][v], ld_out * osz, (size_t)N * osz, steps,
                                     cudaMemcpyDeviceToHost, s_copy));
      }
      MSB_TRYC(cudaEventRecord(ev_c[b], s_copy));
      // the previous tile goes to the consumer while this one computes; its host slot is the one
      // tile k + 1 will overwrite, and that copy is only enqueued after the callback has returned
      if (k >= 1) MSB_TRY(deliver(k - 1));
    }
    MSB_TRY(deliver(n_tiles - 1));
  }
  MSB_TRYC(cudaEventRecord(ev[4], s_main));
  MSB_TRYC(cudaStreamSynchronize(s_main));
  MSB_TRYC(cudaStreamSynchronize(s_copy));
  float ms;
  cudaEventElapsedTime(&ms, ev[0], ev[1]); run->ms_h2d = ms;
  cudaEventElapsedTime(&ms, ev[1], ev[2]); run->ms_solar = ms;
  cudaEventElapsedTime(&ms, ev[2], ev[3]); run->ms_daily = ms;
  cudaEventElapsedTime(&ms, ev[3], ev[4]); run->ms_disagg_total = ms;
  rc = msb_daily_check(&f, &dl, s_main);
  return rc;
#undef MSB_TRY
#undef MSB_TRYC
}

extern "C" int msb_run_host_ctx(msb_host_ctx *ctx, const msb_params *p, msb_host_run *run) {
  if (!ctx) { set_error("msb_run_host_ctx: NULL context"); return MSB_E_ARG; }
  int rc = check_run(p, run);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(ctx->mu);
  MSB_CUDA(cudaSetDevice(ctx->device));
  return run_locked(ctx, p, run);
}

// convenience form: a context that lives for this one call on run->device
extern "C" int msb_run_host(const msb_params *p, msb_host_run *run) {
  int rc = check_run(p, run);
  if (rc) return rc;
  msb_host_ctx *ctx = nullptr;
  rc = msb_host_ctx_create(run->device, &ctx);
  if (rc) return rc;
  rc = msb_run_host_ctx(ctx, p, run);
  // keep the message of a failed run: destroying the context resets nothing but the CUDA error state
  msb_host_ctx_destroy(ctx);
  return rc;
}
