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
//   sw_day_kernel    the day's tpd + 1 entries of the signed prefix table q (solar.cu) are fetched
//                    in batches of eight INDEPENDENT loads (the round-1 kernel issued one dependent
//                    8-byte load per step and stalled 3.6 warps per issue on it), a window is one
//                    subtraction of neighbouring entries, times the radiation of its source day.
//   prec_day_kernel  pure arithmetic: the cumulative closed form of the day's per-minute shape
//                    (prec_cum, disagg_common.cuh) differenced at the window ends; NaN days
//                    (0/0, :327) are forward-filled in-stream like pandas does (:130).
//
// Each thread walks one cell through a chunk of days; lanes are consecutive cells, so every store
// is a full 128-byte row.  wind is written by thermo_knot_kernel (it shares tskc's minute roll).
#include "disagg_common.cuh"

namespace msb {

// One segment of a day's windows: all of them inside one table row, so the entries are an arithmetic
// progression of addresses and a window is rd * (next entry - this entry).  The loop is unrolled by
// four with the loads first: four independent 8-byte loads in flight per thread.
template <typename OutT, bool UNITS>
__device__ __forceinline__ void sw_segment(const double *__restrict__ ptr, int stride, int n, double rd,
                                           double &v0, OutT *&o_sw, size_t ostride, double sc, double of) {
  int t = 0;
  for (; t + 4 <= n; t += 4) {
    const double a0 = __ldg(ptr), a1 = __ldg(ptr + stride), a2 = __ldg(ptr + 2 * stride),
                 a3 = __ldg(ptr + 3 * stride);
    ptr += 4 * stride;
    double s0 = rd * (a0 - v0), s1 = rd * (a1 - a0), s2 = rd * (a2 - a1), s3 = rd * (a3 - a2);
    if (UNITS) { s0 = fma(s0, sc, of); s1 = fma(s1, sc, of); s2 = fma(s2, sc, of); s3 = fma(s3, sc, of); }
    __stcs(o_sw, (OutT)s0);
    __stcs(o_sw + ostride, (OutT)s1);
    __stcs(o_sw + 2 * ostride, (OutT)s2);
    __stcs(o_sw + 3 * ostride, (OutT)s3);
    o_sw += 4 * ostride;
    v0 = a3;
  }
  for (; t < n; ++t) {
    const double a0 = __ldg(ptr);
    ptr += stride;
    double s0 = rd * (a0 - v0);
    if (UNITS) s0 = fma(s0, sc, of);
    __stcs(o_sw, (OutT)s0);
    o_sw += ostride;
    v0 = a0;
  }
}

template <typename OutT, bool UNITS>
__global__ void __launch_bounds__(kDisaggThreads, 8)
sw_day_kernel(const __grid_constant__ DisaggArgs a) {
  __shared__ int s_doy[kDayWin];
  const msb_params &p = a.p;
  const msb_forcing &f = a.f;
  const int D = f.n_days, ts = p.time_step, tpd = a.tpd;
  const int d_lo = a.k.e0 + blockIdx.y * a.k.tile;
  const int d_hi = min(a.k.e1, d_lo + a.k.tile);
  const int win0 = max(d_lo - 2, 0), nwin = min(min(d_hi + 2, D) - win0, kDayWin);
  for (int j = threadIdx.x; j < nwin; j += blockDim.x) s_doy[j] = f.doy[win0 + j];
  __syncthreads();
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= f.n_cells || d_lo >= d_hi) return;
  const long ld = f.ld;
  const int C = 2 * ts;                                      // tiny steps per output step (:660)
  const double lon = wrap_lon(f.lon[cell]);                  // disaggregate.py:67
  const int off_tiny = p.utc_offset ? (int)((lon / kDegPerRev) * kTiny) : 0;   // :655
  // The day's windows start at tiny step c0 of table row / radiation day A = (.) + q and cross
  // into B = A + 1 at window t_b (the first one that starts at or beyond the end of row A);
  // window t_a = t_b - 1 straddles the two when the crossing is not clean.
  const int q = off_tiny < 0 ? -1 : 0;
  const int c0 = off_tiny - q * kTiny;                       // [0, 2880)
  const int t_b = (kTiny - c0 + C - 1) / C;
  const bool straddle = (kTiny - c0) % C != 0;
  const int t_a = straddle ? t_b - 1 : t_b;                  // windows [0, t_a) lie inside row A
  // The window that contains tiny step 1440 of its row: the signed prefix table changes sign there
  // and the row total goes on once (solar.cu).  It splits its row's segment in two.
  const bool noon_in_a = c0 <= kNoon;
  const int t_n = ((noon_in_a ? kNoon : kNoon + kTiny) - c0) / C;
  const int row0 = f.lat_row[cell];
  const double *const qrow = a.t.q + (size_t)row0 * 365 * kQLd;
  const double *const dayl = a.t.daylength + (size_t)row0 * kRows;
  const double *const sw_d = daily_plane(a, a.sw_d, cell);
  auto row_of = [](int y) {      // table row y rolled over the table's own 366 rows (:656)
    y = y < 0 ? y + kRows : (y >= kRows ? y - kRows : y);
    return min(y, 364);          // row 365 is a copy of row 364 (physics.py:281-284)
  };
  auto rad_of = [&](int d) {     // tmp_rad of source day d, disaggregate.py:645
#ifdef MSB_BOUNDS
    if ((unsigned)(d - win0) >= (unsigned)nwin || (unsigned)d >= (unsigned)D) asm volatile("trap;");
#endif
    const double dl = a.sw_full_day ? kSecPerDay : dayl[s_doy[d - win0] - 1];
    return (sw_d[(size_t)d * ld + cell] * dl) * a.inv_sw_den;
  };
  const double sc = UNITS ? a.out.scale[MSB_V_SHORTWAVE] : 1.0, of = UNITS ? a.out.offset[MSB_V_SHORTWAVE] : 0.0;
  const size_t ostride = (size_t)a.out.ld;
  OutT *o_sw = reinterpret_cast<OutT *>(a.out.var[MSB_V_SHORTWAVE]) + cell +
               (size_t)((long)d_lo - (long)a.out.day0) * tpd * ostride;

  double rd_a = rad_of(d_lo + q);
  for (int d = d_lo; d < d_hi; ++d) {
    const int y = s_doy[d - win0] - 1;
    const double *const qa = qrow + row_of(y + q) * kQLd + c0;                    // entry t of row A at qa[t C]
    const double *const qb = qrow + row_of(y + q + 1) * kQLd + (c0 - kTiny);      // entry t of row B at qb[t C]
    const double rd_b = rad_of(d + q + 1);
    const double rowtot = (noon_in_a ? qa : qb)[kTiny + 1 - (noon_in_a ? c0 : c0 - kTiny)];
#ifdef MSB_BOUNDS
    if (qa + (long)t_a * C >= qrow + (size_t)365 * kQLd || qb + (long)tpd * C >= qrow + (size_t)365 * kQLd ||
        qb + (long)t_b * C < qrow) asm volatile("trap;");
#endif
    // Tomorrow's entries into L1 while today computes: the segments below are short loops of dependent
    // batches (a batch of four loads, then its four windows), so what hides the table latency is
    // that the loads hit L1.  One prefetch per entry; neighbouring lanes share the lines.
    if (a.k.pf && d + 1 < d_hi) {
      const int y1 = s_doy[d + 1 - win0] - 1;
      const double *pa = qrow + row_of(y1 + q) * kQLd + c0;
      const double *pb = qrow + row_of(y1 + q + 1) * kQLd + (c0 - kTiny) + (long)t_b * C;
      for (int t = 0; t <= t_a; ++t, pa += C) prefetch_l1(pa);
      for (int t = t_b; t <= tpd; ++t, pb += C) prefetch_l1(pb);
    }
    double v0 = __ldg(qa);
    // ---- windows inside row A: [0, t_a), the noon window (if it is here) splitting them ----
    if (noon_in_a) {
      sw_segment<OutT, UNITS>(qa + C, C, t_n, rd_a, v0, o_sw, ostride, sc, of);
      v0 -= rowtot;
      sw_segment<OutT, UNITS>(qa + (long)(t_n + 1) * C, C, t_a - t_n, rd_a, v0, o_sw, ostride, sc, of);
    } else {
      sw_segment<OutT, UNITS>(qa + C, C, t_a, rd_a, v0, o_sw, ostride, sc, of);
    }
    // ---- the straddling window: [.., 2880) of row A (q_A[2880] = 0) + [0, ..) of row B (q_B[0] = 0) ----
    if (straddle) {
      const double v1 = __ldg(qb + (long)t_b * C);
      double sw = fma(rd_b, v1, -(rd_a * v0));
      if (UNITS) sw = fma(sw, sc, of);
      __stcs(o_sw, (OutT)sw);
      o_sw += ostride;
      v0 = v1;
    } else {
      v0 = 0.0;                    // clean crossing: row B starts at its entry 0 (an empty prefix)
    }
    // ---- windows inside row B: [t_b, tpd) ----
    if (noon_in_a) {
      sw_segment<OutT, UNITS>(qb + (long)(t_b + 1) * C, C, tpd - t_b, rd_b, v0, o_sw, ostride, sc, of);
    } else {
      sw_segment<OutT, UNITS>(qb + (long)(t_b + 1) * C, C, t_n - t_b, rd_b, v0, o_sw, ostride, sc, of);
      v0 -= rowtot;
      sw_segment<OutT, UNITS>(qb + (long)(t_n + 1) * C, C, tpd - t_n, rd_b, v0, o_sw, ostride, sc, of);
    }
    rd_a = rd_b;
  }
}

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

template <typename OutT, bool UNITS>
__global__ void __launch_bounds__(kDisaggThreads, 6)
prec_day_kernel(const __grid_constant__ DisaggArgs a) {
  __shared__ int s_month[kDayWin];
  const msb_params &p = a.p;
  const msb_forcing &f = a.f;
  const int D = f.n_days, ts = p.time_step, tpd = a.tpd;
  const int d_lo = a.k.e0 + blockIdx.y * a.k.tile;
  const int d_hi = min(a.k.e1, d_lo + a.k.tile);
  const int win0 = max(d_lo - 2, 0), nwin = min(min(d_hi + 2, D) - win0, kDayWin);
  for (int j = threadIdx.x; j < nwin; j += blockDim.x) s_month[j] = f.month[win0 + j];
  __syncthreads();
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= f.n_cells || d_lo >= d_hi) return;
  const double lon = wrap_lon(f.lon[cell]);                                      // disaggregate.py:67
  const int off_min = p.utc_offset ? (int)((lon / kDegPerRev) * kMinPerDay) : 0;  // :349
  // the day's windows start at minute r0 of source day A = d + q and cross into B = A + 1 at
  // window t_b; n1 minutes of the straddling window t_x = t_b - 1 belong to day A (0 = clean)
  const int q = off_min < 0 ? -1 : 0;
  const int r0 = off_min - q * kMinPerDay;                   // [0, 1440)
  const int t_b = (kMinPerDay - r0 + ts - 1) / ts;
  const int n1 = (kMinPerDay - r0) % ts;
  const int t_a = n1 ? t_b - 1 : t_b;                        // windows [0, t_a) lie inside day A
  auto setup = [&](int d, PrecDay &pd, double &tot) {
#ifdef MSB_BOUNDS
    if ((unsigned)(d - win0) >= (unsigned)nwin || (unsigned)d >= (unsigned)D) asm volatile("trap;");
#endif
    prec_day_setup(p, f, cell, d, s_month[d - win0], pd, &tot);
  };
  const double sc = UNITS ? a.out.scale[MSB_V_PREC] : 1.0, of = UNITS ? a.out.offset[MSB_V_PREC] : 0.0;
  const size_t ostride = (size_t)a.out.ld;
  const long i_lo = (long)d_lo * tpd;
  OutT *const o_base = reinterpret_cast<OutT *>(a.out.var[MSB_V_PREC]) + cell +
                       (size_t)(i_lo - (long)a.out.day0 * tpd) * ostride;
  OutT *o_pr = o_base;
  auto put = [&](double v) {
    if (UNITS) v = fma(v, sc, of);
    __stcs(o_pr, (OutT)v);
    o_pr += ostride;
  };

  PrecDay pa, pb;
  double tot_a, tot_b;           // unscaled totals of the two days' shapes, S(1440)
  setup(d_lo + q, pa, tot_a);
  // What pandas' ffill repeats into a NaN step is the last valid value before it: inside the chunk
  // that is simply the value of the previous step (`last`); a chunk that OPENS with a NaN step
  // looks back once (cold), and if nothing valid precedes it the run is back-filled afterwards.
  double last = 0.0;
  bool have_last = false, need_fixup = false;
  auto need_last = [&]() {
    if (!have_last) {      // nothing emitted yet: the NaN step is the chunk's first, i_lo
      last = prec_fill_before(p, f, cell, tpd, off_min, i_lo);
      need_fixup = !(last == last);
      have_last = true;
    }
  };
  // cumulative shape of day A at the day's first window start; from the second day on it is where
  // the previous day's windows inside (then) day B ended
  double s_a = (pa.f == pa.f) ? prec_cum(pa, r0) : 0.0;
  for (int d = d_lo; d < d_hi; ++d) {
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
  if (need_fixup) {
    long first_valid;
    const double fillv = prec_fill_after(p, f, cell, tpd, off_min, i_lo, &first_valid);
    const long i_hi = (long)d_hi * tpd;
    o_pr = o_base;
    for (long j = i_lo; j < (first_valid < i_hi ? first_valid : i_hi); ++j) put(fillv);
  }
}

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
  a.k.pf = a.out.tune.day_prefetch < 0 ? 0 : 1;      // sw_day_kernel: L1 prefetch of the next day's table entries
  const dim3 grid(cell_blocks, (b0 - a0 + a.k.tile - 1) / a.k.tile);
  const bool f4 = a.out.dtype == 4;
  if (a.var_mask & (1u << MSB_V_SHORTWAVE)) {
    const bool units = a.out.scale[MSB_V_SHORTWAVE] != 1.0 || a.out.offset[MSB_V_SHORTWAVE] != 0.0;
    if (f4) {
      if (units) sw_day_kernel<float, true><<<grid, kDisaggThreads, 0, st>>>(a);
      else sw_day_kernel<float, false><<<grid, kDisaggThreads, 0, st>>>(a);
    } else {
      if (units) sw_day_kernel<double, true><<<grid, kDisaggThreads, 0, st>>>(a);
      else sw_day_kernel<double, false><<<grid, kDisaggThreads, 0, st>>>(a);
    }
  }
  if (a.var_mask & (1u << MSB_V_PREC)) {
    const bool units = a.out.scale[MSB_V_PREC] != 1.0 || a.out.offset[MSB_V_PREC] != 0.0;
    if (f4) {
      if (units) prec_day_kernel<float, true><<<grid, kDisaggThreads, 0, st>>>(a);
      else prec_day_kernel<float, false><<<grid, kDisaggThreads, 0, st>>>(a);
    } else {
      if (units) prec_day_kernel<double, true><<<grid, kDisaggThreads, 0, st>>>(a);
      else prec_day_kernel<double, false><<<grid, kDisaggThreads, 0, st>>>(a);
    }
  }
  MSB_CUDA(cudaGetLastError());
  return MSB_OK;
}

}  // namespace msb
