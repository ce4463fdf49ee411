// C ABI plumbing: errors, defaults, and the host-buffer entry point msb_run_host().
//
// msb_run_host() replaces the per-cell loop of MetSim.run_slice (metsim/metsim.py:477-493 ->
// wrap_run_cell :693-796) for one shard of cells: it deduplicates latitudes, uploads the
// forcings, runs solar tables -> daily MTCLIM -> tiled disaggregation and streams every output
// tile back to host memory while the next tile is being computed (two device tile buffers, a
// compute stream and a copy stream; nothing larger than two tiles of outputs ever lives in HBM).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
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

// grow-only device / pinned arenas reused across msb_run_host() calls (one per process)
struct Arena {
  void *dev = nullptr;
  size_t dev_bytes = 0;
  int device = -1;
  void *pin = nullptr;
  size_t pin_bytes = 0;
  std::mutex mu;
};
static Arena g_arena;

static int arena_reserve(Arena &a, int device, size_t dev_bytes, size_t pin_bytes) {
  if (a.device != device || a.dev_bytes < dev_bytes) {
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
    a.device = device;
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

using namespace msb;

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

extern "C" int msb_run_host(const msb_params *p, msb_host_run *run) {
  if (!p || !run) { set_error("msb_run_host: NULL argument"); return MSB_E_ARG; }
  const int N = run->n_cells, D = run->n_days;
  if (N <= 0 || D < 2) { set_error("msb_run_host: need n_cells > 0 and n_days >= 2"); return MSB_E_ARG; }
  if (!run->lat || !run->lon || !run->elev || !run->doy || !run->month || !run->t_min ||
      !run->t_max || !run->prec || !run->st_t_min || !run->st_t_max || !run->st_prec) {
    set_error("msb_run_host: a required input pointer is NULL");
    return MSB_E_ARG;
  }
  const int ts = p->time_step;
  const bool sub = ts < kMinPerDay;
  if (ts <= 0 || kMinPerDay % ts || (ts > 360 && ts != kMinPerDay)) {
    set_error("Time step must be evenly divisible into 1440 minutes (24 hours) and less than 360 "
              "minutes (6 hours). Got %d.", ts);
    return MSB_E_ARG;
  }
  const int tpd = sub ? kMinPerDay / ts : 1;
  const int osz = run->out_dtype == 8 ? 8 : 4;
  if (run->out_dtype != 4 && run->out_dtype != 8) { set_error("msb_run_host: out_dtype must be 4 or 8"); return MSB_E_ARG; }
  MSB_CUDA(cudaSetDevice(run->device));

  // ---- unique latitudes (exact bit equality), run-length first: grids are latitude-row-major --
  std::vector<int32_t> lat_row(N);
  std::vector<double> ulat;
  {
    std::unordered_map<uint64_t, int32_t> seen;
    uint64_t prev_bits = 0;
    int32_t prev_row = -1;
    for (int i = 0; i < N; ++i) {
      uint64_t bits;
      std::memcpy(&bits, &run->lat[i], 8);
      if (prev_row >= 0 && bits == prev_bits) { lat_row[i] = prev_row; continue; }
      auto it = seen.find(bits);
      if (it == seen.end()) {
        it = seen.emplace(bits, (int32_t)ulat.size()).first;
        ulat.push_back(run->lat[i]);
      }
      prev_bits = bits;
      prev_row = it->second;
      lat_row[i] = prev_row;
    }
  }
  const int n_lat = (int)ulat.size();

  // ---- carve the device arena ----
  int n_out = 0;
  for (int v = 0; v < MSB_N_SUBVARS; ++v) n_out += (sub && run->out_host[v]) ? 1 : 0;
  int tile_days = run->tile_days > 0 ? run->tile_days : 0;
  if (sub && tile_days == 0) {
    // two tile buffers of all requested outputs within ~8 GB, at least 1 day
    size_t per_day = (size_t)n_out * tpd * (size_t)N * osz;
    tile_days = per_day ? (int)((4ull << 30) / per_day) : D;
    if (tile_days < 1) tile_days = 1;
    if (tile_days > D) tile_days = D;
  }
  if (!sub) tile_days = D;
  const size_t dn = (size_t)D * N, sn = (size_t)kState * N;
  size_t small_e, q_e, tt_e;
  msb_solar_table_sizes(n_lat, &small_e, &q_e, &tt_e);
  const size_t prelude = msb_solar_prelude_bytes(n_lat);
  size_t need = 0;
  auto add = [&](size_t bytes) { need = ((need + 255) & ~(size_t)255) + bytes; };
  add(3 * N * 8); add(N * 4); add(2 * (size_t)D * 4);            // lon/elev(+pad), lat_row, doy/month
  for (int k = 0; k < 4; ++k) add(dn * 8);                       // t_min t_max prec wind
  const double *given_h[4] = {run->given_shortwave, run->given_vapor_pressure, run->given_tskc,
                              run->given_longwave};
  for (int k = 0; k < 4; ++k) if (given_h[k]) add(dn * 8);       // passthrough columns
  for (int k = 0; k < 3; ++k) add(sn * 8);
  add(12 * (size_t)N * 8); add(12 * (size_t)N * 8); add(2 * (size_t)N * 8);
  for (int k = 0; k < 4; ++k) add(small_e * 8);
  add(q_e * 8); add(tt_e * 8); add(prelude);
  for (int k = 0; k < MSB_N_DAILYVARS; ++k) add(dn * 8);
  add((size_t)N * 4); add(msb_daily_scratch_bytes(N, D)); add(64);
  // device output tiles with rows on 128-byte boundaries (leading dimension padded to 32 cells):
  // warp-wide stores then never share a sector with the neighbouring row
  const size_t ld_out = ((size_t)N + 31) / 32 * 32;
  const size_t tile_elems = (size_t)tile_days * tpd * ld_out;
  for (int k = 0; k < 2 * n_out; ++k) add(tile_elems * osz);
  const bool lw_rescale = p->method == MSB_METHOD_PASSTHROUGH && run->given_longwave &&
                          run->out_host[MSB_V_LONGWAVE];
  const size_t ws_bytes = !sub ? 0 : lw_rescale ? msb_disagg_workspace_bytes_lw((int64_t)ld_out, 0, tile_days, ts)
                                                : msb_disagg_workspace_bytes(N, 0, tile_days);
  add(ws_bytes);
  need += 65536;

  std::lock_guard<std::mutex> lock(g_arena.mu);
  int rc = arena_reserve(g_arena, run->device, need, prelude + 4096);
  if (rc) return rc;
  Carver cv(g_arena.dev);
  double *d_lon = cv.take<double>(N), *d_elev = cv.take<double>(N);
  int32_t *d_row = cv.take<int32_t>(N), *d_doy = cv.take<int32_t>(D), *d_month = cv.take<int32_t>(D);
  double *d_tmin = cv.take<double>(dn), *d_tmax = cv.take<double>(dn), *d_prec = cv.take<double>(dn);
  double *d_wind = run->wind ? cv.take<double>(dn) : nullptr;
  double *d_given[4];
  for (int k = 0; k < 4; ++k) d_given[k] = given_h[k] ? cv.take<double>(dn) : nullptr;
  double *d_stmin = cv.take<double>(sn), *d_stmax = cv.take<double>(sn), *d_stprec = cv.take<double>(sn);
  double *d_tpk = run->t_pk12 ? cv.take<double>(12 * (size_t)N) : nullptr;
  double *d_dur = run->dur12 ? cv.take<double>(12 * (size_t)N) : nullptr;
  double *d_tend = run->t_end ? cv.take<double>(2 * (size_t)N) : nullptr;
  msb_solar_tables tb;
  tb.n_lat = n_lat;
  tb.daylength = cv.take<double>(small_e); tb.potrad = cv.take<double>(small_e);
  tb.rise_min = cv.take<double>(small_e); tb.set_min = cv.take<double>(small_e);
  tb.q = cv.take<double>(q_e); tb.tt_coef = cv.take<double>(tt_e);
  void *d_prelude = cv.take<char>(prelude);
  msb_daily dl;
  std::memset(&dl, 0, sizeof(dl));
  for (int k = 0; k < MSB_N_DAILYVARS; ++k) {
    bool needed = k == MSB_D_TSKC || k == MSB_D_VAPOR_PRESSURE || k == MSB_D_SHORTWAVE;
    if (needed || run->daily_host[k]) dl.var[k] = cv.take<double>(dn);
  }
  dl.n_iter = cv.take<int32_t>(N);
  dl.partial = cv.take<double>(msb_daily_scratch_bytes(N, D) / 8 + 1);
  dl.status = cv.take<int32_t>(16);
  void *d_ws = ws_bytes ? (void *)cv.take<char>(ws_bytes) : nullptr;
  void *d_out[2][MSB_N_SUBVARS];
  for (int b = 0; b < 2; ++b)
    for (int v = 0; v < MSB_N_SUBVARS; ++v)
      d_out[b][v] = (sub && run->out_host[v]) ? (void *)cv.take<char>(tile_elems * osz) : nullptr;

  cudaStream_t s_main, s_copy;
  MSB_CUDA(cudaStreamCreateWithFlags(&s_main, cudaStreamNonBlocking));
  MSB_CUDA(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
  cudaEvent_t ev[6], ev_k[2], ev_c[2];
  for (auto &e : ev) MSB_CUDA(cudaEventCreate(&e));
  for (auto &e : ev_k) MSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &e : ev_c) MSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  auto cleanup = [&]() {
    for (auto &e : ev) cudaEventDestroy(e);
    for (auto &e : ev_k) cudaEventDestroy(e);
    for (auto &e : ev_c) cudaEventDestroy(e);
    cudaStreamDestroy(s_main);
    cudaStreamDestroy(s_copy);
  };
#define MSB_TRY(expr)                               \
  do {                                              \
    int _rc = (expr);                               \
    if (_rc != MSB_OK) { cleanup(); return _rc; }   \
  } while (0)
#define MSB_TRYC(expr)                                                          \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #expr); }          \
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
  // lat_row lives in a std::vector: make sure the pageable copy has been staged before it dies
  MSB_TRYC(cudaEventRecord(ev[1], s_main));

  // ---- solar tables (host prelude overlaps the uploads) ----
  unsigned hw = std::thread::hardware_concurrency();
  MSB_TRY(msb_solar_tables_build(n_lat, ulat.data(), g_arena.pin, d_prelude, &tb,
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
      MSB_TRY(msb_disaggregate(p, &f, &tb, &dl, &so, s_main));
      MSB_TRYC(cudaEventRecord(ev_k[b], s_main));
      MSB_TRYC(cudaStreamWaitEvent(s_copy, ev_k[b], 0));
      const size_t steps = (size_t)(d1 - d0) * tpd;
      for (int v = 0; v < MSB_N_SUBVARS; ++v) {
        if (!d_out[b][v]) continue;
        size_t host_step0 = run->out_is_ring ? (size_t)b * tile_days * tpd : (size_t)d0 * tpd;
        char *dst = reinterpret_cast<char *>(run->out_host[v]) + host_step0 * (size_t)N * osz;
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
  cleanup();
  return rc;
#undef MSB_TRY
#undef MSB_TRYC
}
