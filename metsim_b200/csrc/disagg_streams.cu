// Shortwave and precipitation of the interior days [2, D-2) of a record: one lean kernel each.
//
// Reference: metsim/disaggregate.py:618-680 (shortwave) and :265-352 (prec).  Both variables are
// per-day closed forms of a SOURCE stream rolled by the cell's longitude offset (np.roll over the
// whole record, :348-350 / :656-657): on an interior day neither rolled stream wraps around the
// record, so output day d reads exactly two consecutive source days / table rows,
//     A = d + q   up to the roll,    B = A + 1   after it        (q = -1 for negative offsets, else 0)
// and the step at which the roll falls is the same on every day of a cell.  That makes a
// (cell, day) self-contained: no event queue, no state carried from step to step.
//
//   SwCell    a day is at most five segments of windows inside ONE row of the signed prefix table q
//             (solar.cu); a segment is a loop of four INDEPENDENT 8-byte table loads followed by
//             their four windows (the round-1 kernel issued one dependent load per step and stalled
//             3.6 warps per issue on it); tomorrow's entries are prefetched into L1.
//   PrecCell  pure arithmetic: the cumulative closed form of the day's per-minute shape
//             (prec_cum, disagg_common.cuh) differenced at the window ends; NaN days
//             (0/0, :327) are forward-filled in-stream like pandas does (:130).
//   sw_day_kernel / prec_day_kernel run one of them for a (cell, chunk of days) per thread.
//
// Each thread walks one cell through a chunk of days; lanes are consecutive cells, so every store
// is a full 128-byte row.  wind is written by thermo_knot_kernel (it shares tskc's minute roll).
#include "disagg_common.cuh"

namespace msb {

// One segment of a day's windows: all of them inside one table row, so the entries are an arithmetic
// progression of addresses and a window is rd * (next entry - this entry).  A segment of n windows
// goes out as batches of 8, 4, 2, 1 with the loads of a batch issued together.
// B windows at once: B independent 8-byte table loads in flight, then their B windows
template <typename OutT, bool UNITS, int B>
__device__ __forceinline__ void sw_batch(const double *__restrict__ &ptr, int stride, double rd, double &v0,
                                         OutT *&o_sw, size_t ostride, double sc, double of) {
  double e[B];
#pragma unroll
  for (int k = 0; k < B; ++k) e[k] = __ldg(ptr + (long)k * stride);
  ptr += (long)B * stride;
#pragma unroll
  for (int k = 0; k < B; ++k) {
    double s = rd * (e[k] - (k ? e[k - 1] : v0));
    if (UNITS) s = fma(s, sc, of);
    __stcs(o_sw + (size_t)k * ostride, (OutT)s);
  }
  o_sw += (size_t)B * ostride;
  v0 = e[B - 1];
}

template <typename OutT, bool UNITS>
__device__ __forceinline__ void sw_segment(const double *__restrict__ ptr, int stride, int n, double rd,
                                           double &v0, OutT *&o_sw, size_t ostride, double sc, double of) {
  int t = 0;
#pragma unroll 1
  for (; t + 8 <= n; t += 8) sw_batch<OutT, UNITS, 8>(ptr, stride, rd, v0, o_sw, ostride, sc, of);
  if (t + 4 <= n) { sw_batch<OutT, UNITS, 4>(ptr, stride, rd, v0, o_sw, ostride, sc, of); t += 4; }
  if (t + 2 <= n) { sw_batch<OutT, UNITS, 2>(ptr, stride, rd, v0, o_sw, ostride, sc, of); t += 2; }
  if (t < n) sw_batch<OutT, UNITS, 1>(ptr, stride, rd, v0, o_sw, ostride, sc, of);
}

// ---- shortwave of one cell: per-cell constants, then one call per output day --------------------
// WIN: the window-major copy of the table (msb_solar_tables.q_win) instead of the plain one -- a
// compile-time choice, so that the plain path keeps its constant row pitch and the window-major
// one its unit stride.
template <typename OutT, bool UNITS, bool WIN>
struct SwCell {
  const double *qrow, *dayl, *sw_col;   // latitude row of the table / daylength, the cell's daily shortwave column
  const int *s_doy;
  OutT *o;
  size_t ostride, ld;
  double rd_a, inv_den, sc, of;
  int C, c0, q, t_a, t_b, t_n, tpd, win0;
  // where the windows' entries sit in a table row: entry t of row A at row + qa_off + t*S, of row B at
  // row + qb_off + t*S, the row total at row + tot_off.  Plain table: S = C; window-major copy
  // (coarse / masked grids): S = 1 -- a day's entries are consecutive.
  size_t pitch_w;
  int qa_w, qb_w, tot_w;
  bool straddle, noon_in_a, full_day;
  __device__ __forceinline__ size_t pitch() const { return WIN ? pitch_w : (size_t)kQLd; }
  __device__ __forceinline__ int stride() const { return WIN ? 1 : C; }
  __device__ __forceinline__ int qa_off() const { return WIN ? qa_w : c0; }
  __device__ __forceinline__ int qb_off() const { return WIN ? qb_w : c0 - kTiny; }
  __device__ __forceinline__ int tot_off() const { return WIN ? tot_w : kTiny + 1; }
  int pf;

  static __device__ __forceinline__ int row_of(int y) {   // table row y rolled over the table's own 366 rows (:656)
    y = y < 0 ? y + kRows : (y >= kRows ? y - kRows : y);
    return min(y, 364);                                   // row 365 is a copy of row 364 (physics.py:281-284)
  }
  __device__ __forceinline__ double rad_of(int d) const {  // tmp_rad of source day d, disaggregate.py:645
    const double dl = full_day ? kSecPerDay : dayl[s_doy[d - win0] - 1];
    return (sw_col[(size_t)d * ld] * dl) * inv_den;
  }
  __device__ __forceinline__ void init(const DisaggArgs &a, int cell, int d_lo, int win0_, const int *s_doy_) {
    const msb_forcing &f = a.f;
    s_doy = s_doy_; win0 = win0_; tpd = a.tpd; ld = (size_t)f.ld;
    C = 2 * a.p.time_step;                                     // tiny steps per output step (:660)
    const double lon = wrap_lon(f.lon[cell]);                  // disaggregate.py:67
    const int off_tiny = a.p.utc_offset ? (int)((lon / kDegPerRev) * kTiny) : 0;   // :655
    // The day's windows start at tiny step c0 of table row / radiation day A = (.) + q and cross
    // into B = A + 1 at window t_b (the first one that starts at or beyond the end of row A);
    // window t_a = t_b - 1 straddles the two when the crossing is not clean.
    q = off_tiny < 0 ? -1 : 0;
    c0 = off_tiny - q * kTiny;                                 // [0, 2880)
    t_b = (kTiny - c0 + C - 1) / C;
    straddle = (kTiny - c0) % C != 0;
    t_a = straddle ? t_b - 1 : t_b;                            // windows [0, t_a) lie inside row A
    // The window that contains tiny step 1440 of its row: the signed prefix table changes sign
    // there and the row total goes on once (solar.cu).  It splits its row's segment in two.
    noon_in_a = c0 <= kNoon;
    t_n = ((noon_in_a ? kNoon : kNoon + kTiny) - c0) / C;
    const int row0 = f.lat_row[cell];
    if (WIN) {
      const int W = q_win_w(a.p.time_step), ph = c0 % C, j0 = c0 / C;   // c0 = ph + C j0; row B: c0 - 2880 = ph + C (j0 - tpd)
      pitch_w = (size_t)q_win_ld(a.p.time_step);
      qa_w = ph * W + j0; qb_w = ph * W + j0 - tpd; tot_w = C * W;
      qrow = a.t.q_win + (size_t)row0 * 365 * pitch_w;
    } else {
      qrow = a.t.q + (size_t)row0 * 365 * kQLd;
    }
    dayl = a.t.daylength + (size_t)row0 * kRows;
    sw_col = daily_plane(a, a.sw_d, cell) + cell;
    full_day = a.sw_full_day; inv_den = a.inv_sw_den; pf = a.k.pf;
    sc = UNITS ? a.out.scale[MSB_V_SHORTWAVE] : 1.0; of = UNITS ? a.out.offset[MSB_V_SHORTWAVE] : 0.0;
    ostride = (size_t)a.out.ld;
    o = reinterpret_cast<OutT *>(a.out.var[MSB_V_SHORTWAVE]) + cell +
        (size_t)((long)d_lo - (long)a.out.day0) * tpd * ostride;
    rd_a = rad_of(d_lo + q);
  }
  __device__ __forceinline__ void day(int d, bool more) {
    const int y = s_doy[d - win0] - 1;
    const int S = stride();
    const double *const ra = qrow + row_of(y + q) * pitch(), *const rb = qrow + row_of(y + q + 1) * pitch();
    const double *const qa = ra + qa_off();                // entry t of row A at qa[t S]
    const double *const qb = rb + qb_off();                // entry t of row B at qb[t S]
    const double rd_b = rad_of(d + q + 1);
    const double rowtot = (noon_in_a ? ra : rb)[tot_off()];
#ifdef MSB_BOUNDS
    if (qa + (long)t_a * S >= qrow + (size_t)365 * pitch() || qb + (long)tpd * S >= qrow + (size_t)365 * pitch() ||
        qb + (long)t_b * S < qrow) asm volatile("trap;");
#endif
    // The segments below are short loops of dependent batches (four loads, then their four windows),
    // so what hides the table latency is that the loads hit L1: pf = 1 prefetches tomorrow's entries
    // while today computes, pf = 2 the NEXT SEGMENT's entries while this one computes (a third of
    // the L1 footprint, a lead of a few thousand cycles).  One prefetch per entry; neighbouring
    // lanes share the lines.
    auto fetch = [&](const double *ptr, int n) {
      if (WIN) {                          // consecutive entries: one prefetch per 128-byte line
        for (int t = 0; t < n; t += 16) prefetch_l1(ptr + t);
        prefetch_l1(ptr + n - 1);
      } else {
        for (int t = 0; t < n; ++t, ptr += S) prefetch_l1(ptr);
      }
    };
    const double *na = qa, *nb = qb;      // tomorrow's rows (pf = 2: its first segment goes out last)
    if (pf && more) {
      const int y1 = s_doy[d + 1 - win0] - 1;
      na = qrow + row_of(y1 + q) * pitch() + qa_off();
      nb = qrow + row_of(y1 + q + 1) * pitch() + qb_off();
      if (pf == 1) {
        fetch(na, t_a + 1);
        fetch(nb + (long)t_b * S, tpd - t_b + 1);
      }
    }
    const bool seg = pf == 2;
    double v0 = __ldg(qa);
    // ---- windows inside row A: [0, t_a), the noon window (if it is here) splitting them ----
    if (noon_in_a) {
      if (seg) fetch(qa + (long)(t_n + 1) * S, t_a - t_n);
      sw_segment<OutT, UNITS>(qa + S, S, t_n, rd_a, v0, o, ostride, sc, of);
      v0 -= rowtot;
      if (seg) fetch(qb + (long)t_b * S, tpd - t_b + 1);
      sw_segment<OutT, UNITS>(qa + (long)(t_n + 1) * S, S, t_a - t_n, rd_a, v0, o, ostride, sc, of);
    } else {
      if (seg) fetch(qb + (long)t_b * S, t_n - t_b + 1);
      sw_segment<OutT, UNITS>(qa + S, S, t_a, rd_a, v0, o, ostride, sc, of);
    }
    // ---- the straddling window: [.., 2880) of row A (q_A[2880] = 0) + [0, ..) of row B (q_B[0] = 0) ----
    if (straddle) {
      const double v1 = __ldg(qb + (long)t_b * S);
      double sw = fma(rd_b, v1, -(rd_a * v0));
      if (UNITS) sw = fma(sw, sc, of);
      __stcs(o, (OutT)sw);
      o += ostride;
      v0 = v1;
    } else {
      v0 = 0.0;                    // clean crossing: row B starts at its entry 0 (an empty prefix)
    }
    // ---- windows inside row B: [t_b, tpd) ----
    if (noon_in_a) {
      if (seg && more) fetch(na, t_n + 1);
      sw_segment<OutT, UNITS>(qb + (long)(t_b + 1) * S, S, tpd - t_b, rd_b, v0, o, ostride, sc, of);
    } else {
      if (seg) fetch(qb + (long)(t_n + 1) * S, tpd - t_n);
      sw_segment<OutT, UNITS>(qb + (long)(t_b + 1) * S, S, t_n - t_b, rd_b, v0, o, ostride, sc, of);
      v0 -= rowtot;
      if (seg && more) fetch(na, t_a + 1);
      sw_segment<OutT, UNITS>(qb + (long)(t_n + 1) * S, S, tpd - t_n, rd_b, v0, o, ostride, sc, of);
    }
    rd_a = rd_b;
  }
};

// value of the last valid step before step i / first valid step at or after it: the cold paths of
// pandas' ffill / bfill (disaggregate.py:130) around NaN days
__device__ __noinline__ double prec_fill_before(const msb_params &p, const msb_forcing &f, int cell,
                                                int tpd, int off_min, long i) {
  Cell c;
  c.cell = cell; c.D = f.n_days; c.ts = p.time_step; c.tpd = tpd; c.nk = 2 * f.n_days + 4; c.ld = f.ld;
  c.off_min = off_min; c.utc_min = 0.0; c.frac = 0.0; c.rise = c.set = nullptr;
  c.doy = f.doy; c.sdoy = nullptr; c.win0 = 0; c.nwin = 0;
  c.t_min = f.t_min; c.t_max = f.t_max; c.st_t_min = f.st_t_min; c.st_t_max = f.st_t_max;
  c.t_end = f.t_end;
  return prec_lookback(p, f, c, i);
}
__device__ __noinline__ double prec_fill_after(const msb_params &p, const msb_forcing &f, int cell,
                                               int tpd, int off_min, long i, long *first_valid) {
  Cell c;
  c.cell = cell; c.D = f.n_days; c.ts = p.time_step; c.tpd = tpd; c.nk = 2 * f.n_days + 4; c.ld = f.ld;
  c.off_min = off_min; c.utc_min = 0.0; c.frac = 0.0; c.rise = c.set = nullptr;
  c.doy = f.doy; c.sdoy = nullptr; c.win0 = 0; c.nwin = 0;
  c.t_min = f.t_min; c.t_max = f.t_max; c.st_t_min = f.st_t_min; c.st_t_max = f.st_t_max;
  c.t_end = f.t_end;
  const long T = (long)f.n_days * tpd;
  double v = nan("");
  long j = i;
  for (; j < T; ++j) {
    v = prec_step_generic(p, f, c, j);
    if (v == v) break;
  }
  *first_valid = j;
  return v;
}

// ---- precipitation of one cell: per-cell constants and carried state, one call per output day ---
template <typename OutT, bool UNITS>
struct PrecCell {
  const DisaggArgs *a;
  const int *s_month;
  OutT *o, *o_base;
  size_t ostride;
  PrecDay pa;
  double tot_a, s_a, last, sc, of;
  long i_lo;
  int cell, off_min, q, r0, t_a, t_b, n1, ts, tpd, win0;
  bool have_last, need_fixup;

  __device__ __forceinline__ void put(double v) {
    if (UNITS) v = fma(v, sc, of);
    __stcs(o, (OutT)v);
    o += ostride;
  }
  // What pandas' ffill repeats into a NaN step is the last valid value before it: inside the chunk
  // that is simply the value of the previous step (`last`); a chunk that OPENS with a NaN step
  // looks back once (cold), and if nothing valid precedes it the run is back-filled afterwards.
  __device__ __forceinline__ void need_last() {
    if (!have_last) {      // nothing emitted yet: the NaN step is the chunk's first, i_lo
      last = prec_fill_before(a->p, a->f, cell, tpd, off_min, i_lo);
      need_fixup = !(last == last);
      have_last = true;
    }
  }
  __device__ __forceinline__ void setup(int d, PrecDay &pd, double &tot) const {
    prec_day_setup(a->p, a->f, cell, d, s_month[d - win0], pd, &tot);
  }
  __device__ __forceinline__ void init(const DisaggArgs &args, int cell_, int d_lo, int win0_, const int *s_month_) {
    a = &args; cell = cell_; win0 = win0_; s_month = s_month_;
    ts = args.p.time_step; tpd = args.tpd;
    const double lon = wrap_lon(args.f.lon[cell]);                                      // disaggregate.py:67
    off_min = args.p.utc_offset ? (int)((lon / kDegPerRev) * kMinPerDay) : 0;          // :349
    // the day's windows start at minute r0 of source day A = d + q and cross into B = A + 1 at
    // window t_b; n1 minutes of the straddling window t_b - 1 belong to day A (0 = clean)
    q = off_min < 0 ? -1 : 0;
    r0 = off_min - q * kMinPerDay;                   // [0, 1440)
    t_b = (kMinPerDay - r0 + ts - 1) / ts;
    n1 = (kMinPerDay - r0) % ts;
    t_a = n1 ? t_b - 1 : t_b;                        // windows [0, t_a) lie inside day A
    sc = UNITS ? args.out.scale[MSB_V_PREC] : 1.0; of = UNITS ? args.out.offset[MSB_V_PREC] : 0.0;
    ostride = (size_t)args.out.ld;
    i_lo = (long)d_lo * tpd;
    o_base = reinterpret_cast<OutT *>(args.out.var[MSB_V_PREC]) + cell +
             (size_t)(i_lo - (long)args.out.day0 * tpd) * ostride;
    o = o_base;
    setup(d_lo + q, pa, tot_a);
    last = 0.0; have_last = false; need_fixup = false;
    // cumulative shape of day A at the day's first window start; from the second day on it is where
    // the previous day's windows inside (then) day B ended
    s_a = (pa.f == pa.f) ? prec_cum(pa, r0) : 0.0;
  }
  __device__ __forceinline__ void day(int d) {
    PrecDay pb;
    double tot_b;
    setup(d + q + 1, pb, tot_b);
    const bool nan_a = !(pa.f == pa.f), nan_b = !(pb.f == pb.f);
    double s_prev = s_a;
    // windows inside source day A
    if (t_a > 0) {
      if (nan_a) {
        need_last();
        for (int t = 0; t < t_a; ++t) put(last);
      } else {
        int m = r0;
#pragma unroll 2
        for (int t = 0; t < t_a; ++t) {
          m += ts;
          const double s1 = prec_cum(pa, m);
          last = pa.f * (s1 - s_prev);
          s_prev = s1;
          put(last);
        }
        have_last = true;
      }
    }
    // the straddling window: n1 minutes of day A, ts - n1 of day B; a NaN part makes it NaN
    double s_b = 0.0;              // cumulative shape of day B where its first full window starts
    if (n1) {
      if (!nan_b) s_b = prec_cum(pb, ts - n1);
      if (nan_a || nan_b) need_last();
      else last = fma(pb.f, s_b, pa.f * (tot_a - s_prev));
      put(last);
      have_last = true;
    }
    // windows inside source day B
    s_prev = s_b;
    if (t_b < tpd) {
      if (nan_b) {
        need_last();
        for (int t = t_b; t < tpd; ++t) put(last);
      } else {
        int m = t_b * ts + r0 - kMinPerDay;                  // minute of day B at which window t_b starts
#pragma unroll 2
        for (int t = t_b; t < tpd; ++t) {
          m += ts;
          const double s1 = prec_cum(pb, m);
          last = pb.f * (s1 - s_prev);
          s_prev = s1;
          put(last);
        }
        have_last = true;
      }
    }
    pa = pb;
    tot_a = tot_b;
    s_a = s_prev;                  // = S_B(r0): day B is tomorrow's day A
  }
  // a NaN run that reaches back to the start of the record: bfill (disaggregate.py:130)
  __device__ __forceinline__ void finish(int d_hi) {
    if (!need_fixup) return;
    long first_valid;
    const double fillv = prec_fill_after(a->p, a->f, cell, tpd, off_min, i_lo, &first_valid);
    const long i_hi = (long)d_hi * tpd;
    o = o_base;
    for (long j = i_lo; j < (first_valid < i_hi ? first_valid : i_hi); ++j) put(fillv);
  }
};

// which of the two variables a launch produces.  (Both in one kernel -- a thread doing the shortwave
// segments and then the precipitation loops of each day, so that one warp's table loads would be
// covered by another's arithmetic -- was measured and dropped: 96 registers, 20 warps per SM and a
// shared L1 cost more than the overlap gives, 4.42 ms against 1.72 + 1.66 ms per 97-day call.)
enum { kStreamSw = 1, kStreamPrec = 2 };

template <typename OutT, bool UNITS, int WHAT, bool WIN = false>
__device__ __forceinline__ void streams_body(const DisaggArgs &a) {
  __shared__ int s_doy[kDayWin], s_month[kDayWin];
  const msb_forcing &f = a.f;
  const int D = f.n_days;
  const int d_lo = a.k.e0 + blockIdx.y * a.k.tile;
  const int d_hi = min(a.k.e1, d_lo + a.k.tile);
  const int win0 = max(d_lo - 2, 0), nwin = min(min(d_hi + 2, D) - win0, kDayWin);
  for (int j = threadIdx.x; j < nwin; j += blockDim.x) {
    s_doy[j] = f.doy[win0 + j];
    s_month[j] = f.month[win0 + j];
  }
  __syncthreads();
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= f.n_cells || d_lo >= d_hi) return;
#ifdef MSB_BOUNDS   // every day the two cells touch: d_lo - 1 .. d_hi of the staged calendar window
  if (d_lo - 1 < win0 || d_hi >= win0 + nwin || d_hi >= D) asm volatile("trap;");
#endif
  SwCell<OutT, UNITS, WIN> sw;
  PrecCell<OutT, UNITS> pr;
  if (WHAT & kStreamSw) sw.init(a, cell, d_lo, win0, s_doy);
  if (WHAT & kStreamPrec) pr.init(a, cell, d_lo, win0, s_month);
  for (int d = d_lo; d < d_hi; ++d) {
    if (WHAT & kStreamSw) sw.day(d, d + 1 < d_hi);
    if (WHAT & kStreamPrec) pr.day(d);
  }
  if (WHAT & kStreamPrec) pr.finish(d_hi);
}

// latency-bound: 64 registers, 32 warps per SM
template <typename OutT, bool UNITS>
__global__ void __launch_bounds__(kDisaggThreads, 8)
sw_day_kernel(const __grid_constant__ DisaggArgs a) { streams_body<OutT, UNITS, kStreamSw>(a); }
template <typename OutT, bool UNITS>
__global__ void __launch_bounds__(kDisaggThreads, 8)
sw_day_win_kernel(const __grid_constant__ DisaggArgs a) { streams_body<OutT, UNITS, kStreamSw, true>(a); }

// arithmetic: 80 registers, 24 warps per SM (MSB_PREC_MINB: compile-time experiment knob)
#ifndef MSB_PREC_MINB
#define MSB_PREC_MINB 6
#endif
template <typename OutT, bool UNITS>
__global__ void __launch_bounds__(kDisaggThreads, MSB_PREC_MINB)
prec_day_kernel(const __grid_constant__ DisaggArgs a) { streams_body<OutT, UNITS, kStreamPrec>(a); }

constexpr int kStreamTile = 16;   // output days per thread on full-size grids

int launch_interior_streams(DisaggArgs &a, int a0, int b0, int cell_blocks, cudaStream_t st) {
  a.k.i_lo = a.k.i_hi = 0;
  a.k.e0 = a0; a.k.e1 = b0;            // output days, a.k.tile per thread
  a.k.tile = kStreamTile;
  // small shards: shorter chunks so that the grid still fills the 148 SMs several waves deep;
  // msb_tuning.day_tile pins the value (the parity tests run the benchmarked 16 on small shards)
  while (a.k.tile > 2 && (long)cell_blocks * ((b0 - a0 + a.k.tile - 1) / a.k.tile) < 148L * 8 * 6)
    a.k.tile >>= 1;
  if (a.out.tune.day_tile > 0) a.k.tile = a.out.tune.day_tile < kDayWin - 8 ? a.out.tune.day_tile : kDayWin - 8;
  // sw_day_kernel: L1 prefetch of the next segment's (default) or the next day's (1) table entries, -1 = none
  a.k.pf = a.out.tune.day_prefetch < 0 ? 0 : (a.out.tune.day_prefetch == 1 ? 1 : 2);
  const dim3 grid(cell_blocks, (b0 - a0 + a.k.tile - 1) / a.k.tile);
  const bool f4 = a.out.dtype == 4;
  const bool want_sw = a.var_mask & (1u << MSB_V_SHORTWAVE), want_pr = a.var_mask & (1u << MSB_V_PREC);
  const bool units = (want_sw && (a.out.scale[MSB_V_SHORTWAVE] != 1.0 || a.out.offset[MSB_V_SHORTWAVE] != 0.0)) ||
                     (want_pr && (a.out.scale[MSB_V_PREC] != 1.0 || a.out.offset[MSB_V_PREC] != 0.0));
  // (one flag for both launches: a variable without a conversion multiplies by 1 and adds 0)
#define MSB_STREAMS(KERNEL)                                                    \
  do {                                                                        \
    if (f4) {                                                                 \
      if (units) KERNEL<float, true><<<grid, kDisaggThreads, 0, st>>>(a);     \
      else KERNEL<float, false><<<grid, kDisaggThreads, 0, st>>>(a);          \
    } else {                                                                  \
      if (units) KERNEL<double, true><<<grid, kDisaggThreads, 0, st>>>(a);    \
      else KERNEL<double, false><<<grid, kDisaggThreads, 0, st>>>(a);         \
    }                                                                         \
  } while (0)
  const bool win = a.t.q_win && a.t.q_win_step == a.p.time_step;   // the table copy laid out for this time step
  if (want_sw && win) MSB_STREAMS(sw_day_win_kernel);
  if (want_sw && !win) MSB_STREAMS(sw_day_kernel);
  if (want_pr) MSB_STREAMS(prec_day_kernel);
#undef MSB_STREAMS
  MSB_CUDA(cudaGetLastError());
  return MSB_OK;
}

}  // namespace msb
