// Solar geometry for the unique latitudes of a shard.
//
// Reference: metsim/physics.py:177-285 (solar_geom) builds, PER CELL, a 366x2880 table of
// 30-second radiation fractions plus daylength / flat_potrad / tt_max0, and
// metsim/disaggregate.py:133-201 (set_min_max_hour) re-derives sunrise/sunset from that table.
// Only tt_max0 depends on the cell's elevation; everything else depends on latitude alone.
//
// B200 design (not a port):
//  * K0, host, glibc libm: per (latitude, yday) the five scalars that decide index placement --
//    cosegeom, sinegeom, hss, the step count of numba's arange and the zenith cosine of the first
//    step.  That first value is mathematically 0 and numerically +-1e-17; its SIGN moves the
//    sunrise index by one step, so it must come from the same libm the reference runs on.
//  * K1, device, one CTA per (yday, latitude): the normalised row is built in shared memory and
//    reduced to what the other kernels need -- daylength, potrad, rise/set minutes, a SIGNED
//    PREFIX TABLE q[] from which any window sum of the row is two loads, and Taylor
//    coefficients of tt_max0 in the pressure ratio p (so the per-cell part is a Horner chain).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "msb_common.cuh"

namespace msb {

struct PreludeView {
  double *cosegeom, *sinegeom, *hss, *cza0;  // [n_lat][365]
  int32_t *nstep;                            // [n_lat][365]
  double *beam;                              // [365] dir_beam_topa of the day (latitude-independent) + [2] cos, sin of dh
};

static PreludeView prelude_view(void *base, int n_lat) {
  PreludeView v;
  size_t n = (size_t)n_lat * 365;
  double *d = reinterpret_cast<double *>(base);
  v.cosegeom = d;
  v.sinegeom = d + n;
  v.hss = d + 2 * n;
  v.cza0 = d + 3 * n;
  v.nstep = reinterpret_cast<int32_t *>(d + 4 * n);
  v.beam = d + 4 * n + (n + 1) / 2;          // after the int32 block, 8-byte aligned
  return v;
}

// The prelude workspace is written by a pool of host threads and then read ONCE by the copy engine.
// Plain stores leave its lines dirty in the caches of cores that go idle right after; the DMA read
// then has to pull every line out of a sleeping core (measured on the 32-vCPU B200 host: 6.8 MB at
// 6.7 GB/s, 1.0 ms of a 5.3 ms table build, against 0.2 ms when one awake thread had written it).
// Streaming stores put the values in memory, where the copy engine reads them at PCIe rate.
static inline void store_stream(double *p, double v) {
#if defined(__x86_64__)
  long long b;
  std::memcpy(&b, &v, 8);
  _mm_stream_si64(reinterpret_cast<long long *>(p), b);
#else
  *p = v;
#endif
}
static inline void store_stream(int32_t *p, int32_t v) {
#if defined(__x86_64__)
  _mm_stream_si32(reinterpret_cast<int *>(p), v);
#else
  *p = v;
#endif
}
static inline void store_fence() {
#if defined(__x86_64__)
  _mm_sfence();
#endif
}

// K0 -- physics.py:212-236, :243 (first iteration of the hour-angle loop).  Compiled by the host
// compiler with -ffp-contract=off: numba/LLVM does not fuse multiply-add.
static void prelude_rows(const double *lat_deg, int r0, int r1, PreludeView v) {
  const double dh = kSwRadDt / kSecPerRad;
  // the declination terms depend on the day alone: once per call, not once per latitude
  double cosdecl[365], sindecl[365];
  for (int i = 0; i < 365; ++i) {
    const double decl = kMinDecl * std::cos((i + kDaysOff) * kRadPerDay);
    cosdecl[i] = std::cos(decl);
    sindecl[i] = std::sin(decl);
  }
  for (int r = r0; r < r1; ++r) {
    double lat = std::fmin(std::fmax(lat_deg[r] * kRadPerDeg, -M_PI / 2.), M_PI / 2.0);
    double coslat = std::cos(lat), sinlat = std::sin(lat);
    for (int i = 0; i < 365; ++i) {
      double cosegeom = coslat * cosdecl[i];
      double sinegeom = sinlat * sindecl[i];
      double coshss = std::fmin(std::fmax(-sinegeom / cosegeom, -1.0), 1.0);
      double hss = std::acos(coshss);
      double start = -hss;
      long n = (long)std::ceil((hss - start) / dh);  // numba arange length
      if (n < 0) n = 0;
      double cosh0 = std::cos(start);
      volatile double prod = cosegeom * cosh0;        // keep the product rounded on its own
      double cza0 = prod + sinegeom;
      size_t o = (size_t)r * 365 + i;
      store_stream(v.cosegeom + o, cosegeom);
      store_stream(v.sinegeom + o, sinegeom);
      store_stream(v.hss + o, hss);
      store_stream(v.cza0 + o, cza0);
      store_stream(v.nstep + o, (int32_t)n);
    }
  }
  store_fence();
}

// Class boundaries of the OPTAM lookup for solar_moments_kernel: for j < 20 the largest double c
// with int(acos(c) / RAD_PER_DEG) >= 70 + j, by bisection over the (ordered) bit patterns of the
// positive doubles with this libm's acos.  acos is decreasing, so the predicate is monotone.
static void optam_thresholds(double *c32) {
  for (int j = 0; j < 32; ++j) c32[j] = -1.0;
  for (int j = 0; j < 20; ++j) {
    auto pred = [j](double c) { return (int)(std::acos(c) / kRadPerDeg) >= 70 + j; };
    double lo_d = 0.0, hi_d = 0.5;                 // pred(0) holds (90 deg), pred(0.5) fails (60 deg)
    uint64_t lo, hi;
    std::memcpy(&lo, &lo_d, 8);
    std::memcpy(&hi, &hi_d, 8);
    while (hi - lo > 1) {
      const uint64_t mid = lo + (hi - lo) / 2;
      double m;
      std::memcpy(&m, &mid, 8);
      if (pred(m)) lo = mid; else hi = mid;
    }
    std::memcpy(&c32[j], &lo, 8);
  }
}

// ---------------------------------------------------------------------------------------------
// 320 threads, each scanning a contiguous chunk of 9 row entries: 320 x 9 = 2880, and an ODD stride in
// 8-byte words keeps the 16 lanes of a half-warp on 16 different shared-memory banks (the earlier
// 256 x 12 layout put them on 4: short-scoreboard stalls of 7 warps per issue in the scan passes)
constexpr int kK1Threads = 320;
constexpr int kChunk = 9;

__constant__ double c_optam[21] = {2.90, 3.05, 3.21, 3.39, 3.69, 3.82, 4.07, 4.37, 4.72, 5.12, 5.60,
                                   6.18, 6.88, 7.77, 8.90, 10.39, 12.44, 15.36, 19.79, 26.96, 30.00};

__device__ __forceinline__ int tiny_index(double h) {
  // physics.py:264-266: int(min(max((43200 + h*SEC_PER_RAD)/30, 0), 2879)) with x = 43200 + h*SEC_PER_RAD
  // evaluated without contraction.  The IEEE quotient fl(x/30) never rounds up across an integer m
  // (x < 30m implies x <= 30m - ulp(30m-) and ulp(30m-) >= 16 ulp(m-), so x/30 <= m - 0.53 ulp(m-)),
  // hence int(fl(x/30)) == floor(x/30) exactly, which needs no division: n = rint(x/30) up to one
  // unit, r = x - 30n in one exactly-signed fma, and r < 0 steps back.
  const double x = __dadd_rn(43200.0, __dmul_rn(h, kSecPerRad));
  const double kMagic = 6755399441055744.0;       // 1.5 * 2^52
  double nd = fma(x, 1.0 / 30.0, kMagic);
  int n = __double2loint(nd);
  nd -= kMagic;
  if (fma(-30.0, nd, x) < 0.0) --n;
  return min(max(n, 0), kTiny - 1);
}

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kWarps = kK1Threads / 32;          // 10
constexpr int kHalfSlots = kK1Threads / 2;       // 160 scan slots per half (five warps), all used
constexpr int kHalfChunks = kTiny / 2 / kChunk;  // 160
static_assert(kHalfChunks * kChunk * 2 == kTiny && kHalfChunks <= kHalfSlots && kHalfSlots % 32 == 0,
              "the two half-row scans must tile the row with whole warps");

__global__ void __launch_bounds__(kK1Threads)
solar_rows_kernel(int n_lat, const double *__restrict__ g_cosegeom,
                  const double *__restrict__ g_sinegeom, const double *__restrict__ g_hss,
                  const double *__restrict__ g_cza0, const int32_t *__restrict__ g_nstep,
                  const double *__restrict__ g_beam, msb_solar_tables tb) {
  __shared__ __align__(16) double row[kTiny];
  __shared__ __align__(16) double qs[kQLd];
  __shared__ double part[kWarps];
  __shared__ double warp_tot[kWarps];
  __shared__ int s_rise, s_set;

  const int i = blockIdx.x % 365;   // yday row 0..364
  const int r = blockIdx.x / 365;   // latitude row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t o = (size_t)r * 365 + i;
  const double cosegeom = g_cosegeom[o], sinegeom = g_sinegeom[o], hss = g_hss[o];
  const double cza_first = g_cza0[o];
  const int n = g_nstep[o];
  const double dh = kSwRadDt / kSecPerRad;

  for (int j = tid; j < kTiny / 2; j += kK1Threads) reinterpret_cast<double2 *>(row)[j] = make_double2(0.0, 0.0);
  if (tid == 0) { s_rise = -1; s_set = -1; }
  __syncthreads();

  const double beam = g_beam[i];               // physics.py:238-239 (host prelude)
  const double start = -hss;

  double tot_loc = 0.0;    // sum of dir_flat_topa over this thread's steps

  // Each thread walks a contiguous chunk of hour-angle steps: cos(h) comes from one library
  // sincos at the chunk start and the rotation by dh afterwards (<= 12 steps, error ~1e-15), and
  // the tiny-step index of step k+1 is carried over as the index of the next step.
  const int per = (n + kK1Threads - 1) / kK1Threads;
  const int k_lo = min(tid * per, n), k_hi = min(k_lo + per, n);
  if (k_lo < k_hi) {
    const double cd = g_beam[365], sd = g_beam[366];
    double ch, sh;
    sincos(__dadd_rn(start, __dmul_rn((double)k_lo, dh)), &sh, &ch);
    int ts = tiny_index(__dadd_rn(start, __dmul_rn((double)k_lo, dh)));
    for (int k = k_lo; k < k_hi; ++k) {
      const double cza = (k == 0) ? cza_first : __dadd_rn(__dmul_rn(cosegeom, ch), sinegeom);
      double dfa = 0.0;
      if (cza > 0.0) {
        dfa = beam * cza;                                  // physics.py:252
        tot_loc += dfa;
      }
      // "assignment, not accumulation" (physics.py:267): the last k mapping to an index wins
      int ts_next = -1;
      if (k + 1 < n) ts_next = tiny_index(__dadd_rn(start, __dmul_rn((double)(k + 1), dh)));
      if (ts_next != ts) row[ts] = dfa;
      ts = ts_next;
      const double cn = fma(ch, cd, -(sh * sd)), sn = fma(sh, cd, ch * sd);   // rotate by dh
      ch = cn; sh = sn;
    }
  }

  // deterministic two-level reduction of the row total (fixed shuffle tree, fixed order)
  {
    const double v = warp_sum(tot_loc);
    if (lane == 0) part[warp] = v;
  }
  __syncthreads();
  double total = 0.0;
  for (int w = 0; w < kWarps; ++w) total += part[w];

  const double daylength = fmin(2.0 * hss * kSecPerRad, kSecPerDay);  // physics.py:236
  const bool has_day = daylength != 0.0;
  // sunrise / sunset: last j with row[j] <= 0 < row[j+1] / row[j] > 0 >= row[j+1]
  // (disaggregate.py:163-176; signs only, so this runs on the unscaled row)
  int rise = -1, set = -1;
  for (int j = tid; j < kTiny - 1; j += kK1Threads) {
    int a = row[j] > 0.0, b = row[j + 1] > 0.0;
    if (b - a > 0) rise = max(rise, j);
    if (b - a < 0) set = max(set, j);
  }
  if (rise >= 0) atomicMax(&s_rise, rise);
  if (set >= 0) atomicMax(&s_set, set);
  __syncthreads();
  // tiny_rad_fract[i] /= sum_flat_potrad (:270-271) as a multiplication by the reciprocal: the
  // entries only enter window sums held to 1e-9, and an IEEE divide per entry was a third of this kernel
  if (has_day && total > 0.0) {
    const double inv_total = div_nr(1.0, total);
    for (int j = tid; j < kTiny; j += kK1Threads) row[j] *= inv_total;
  }
  __syncthreads();

  // signed prefix table: q[j] = sum_{i<j} row[i] for j <= 1440, -sum_{i>=j} row[i] for j > 1440.
  // Threads 0..127 scan the morning half forwards, threads 128..255 the afternoon half backwards,
  // so small sums near sunrise AND sunset are accumulated from the dark side (no cancellation).
  const int half = tid / kHalfSlots, slot = tid % kHalfSlots;
  const bool valid = slot < kHalfChunks;
  double loc[kChunk];
  double csum = 0.0;
  if (valid) {
#pragma unroll
    for (int e = 0; e < kChunk; ++e) {
      const int rr = slot * kChunk + e;
      const int j = half == 0 ? rr : kTiny - 1 - rr;
      loc[e] = csum;              // exclusive
      csum += row[j];
    }
  }
  double incl = csum;             // inclusive scan of chunk totals inside the warp
  for (int o2 = 1; o2 < 32; o2 <<= 1) {
    double up = __shfl_up_sync(0xffffffffu, incl, o2);
    if (lane >= o2) incl += up;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  double base = incl - csum;
  const int w0 = half * (kWarps / 2);
  for (int w = w0; w < warp; ++w) base += warp_tot[w];
  double half_tot[2];
  for (int hh = 0; hh < 2; ++hh) {
    double v = 0.0;
    for (int w = hh * (kWarps / 2); w < (hh + 1) * (kWarps / 2); ++w) v += warp_tot[w];
    half_tot[hh] = v;
  }
  if (valid) {
#pragma unroll
    for (int e = 0; e < kChunk; ++e) {
      const int rr = slot * kChunk + e;
      if (half == 0) {
        qs[rr] = base + loc[e];                                 // prefix, j = rr < 1440
      } else {
        const int j = kTiny - 1 - rr;                           // j in [1440, 2879]
        const double incl_e = base + loc[e] + row[j];           // sum_{i >= j}
        if (j > kNoon) qs[j] = -incl_e;
      }
    }
  }
  if (tid == 0) {
    qs[kNoon] = half_tot[0];                   // prefix up to (not including) tiny step 1440
    qs[kTiny] = 0.0;                           // empty suffix
    qs[kTiny + 1] = half_tot[0] + half_tot[1]; // row total (1 up to rounding, 0 for polar night)
    for (int e = kTiny + 2; e < kQLd; ++e) qs[e] = 0.0;
  }
  __syncthreads();
  // rows are kQLd = 2888 doubles = 23,104 bytes apart: every row starts on a 16-byte boundary
  double2 *q = reinterpret_cast<double2 *>(tb.q + ((size_t)r * 365 + i) * kQLd);
  for (int j = tid; j < kQLd / 2; j += kK1Threads) q[j] = reinterpret_cast<const double2 *>(qs)[j];
  if (tb.q_win) {
    // window-major copy (metsim_b200.h): entry s at [(s % C) * W + s / C].  A warp moves a tile of
    // 8 phases x 4 windows: the 16 lanes of a half-warp read 8 consecutive entries at two multiples
    // of C (for the usual steps C = 8 mod 16 doubles: 16 different bank pairs), and the 4 windows
    // of a phase fill one aligned 32-byte sector of the row (W is a multiple of 4).
    const int ts = tb.q_win_step, C = 2 * ts, W = q_win_w(ts);
    double *qw = tb.q_win + ((size_t)r * 365 + i) * q_win_ld(ts);
    const int n_rt = (C + 7) / 8, n_jt = (W + 3) / 4;
    for (int w = warp; w < n_rt * n_jt; w += kWarps) {
      const int rt = w % n_rt, jt = w / n_rt;
      const int ph = rt * 8 + (lane & 7), j = jt * 4 + (lane >> 3);
      if (ph < C && j < W) {
        const int pos = ph + C * j;
        qw[ph * W + j] = pos <= kTiny ? qs[pos] : 0.0;     // beyond 2880: never read (padding, phase > 0 of the last window)
      }
    }
    if (tid < 4) qw[C * W + tid] = tid == 0 ? qs[kTiny + 1] : 0.0;
  }

  if (tid == 0) {
    const double step_min = kSwRadDt / 60.0;
    double rise_m = (s_rise >= 0) ? s_rise * step_min : 0.0;
    double set_m = (s_set >= 0) ? s_set * step_min : 0.0;
    const double day_tot = (kSecPerDay / kSwRadDt) * (kSwRadDt / 60.0);
    if (rise_m == 0.0) rise_m = day_tot / 6;       // disaggregate.py:179-182
    if (set_m == 0.0) set_m = 4 * day_tot / 6;
    size_t s = (size_t)r * kRows + i;
    tb.rise_min[s] = rise_m;
    tb.set_min[s] = set_m;
    if (i == 364) {  // row 365 copies row 364, physics.py:281-284
      tb.rise_min[s + 1] = rise_m;
      tb.set_min[s + 1] = set_m;
    }
  }
}

// K1b -- one WARP per (latitude, yday): the sums over the hour-angle steps that do not need the
// row in shared memory -- sum of dir_flat_topa (-> potrad) and the Taylor moments of trans**am
// in the pressure ratio (-> tt_coef).  Lanes walk contiguous chunks of steps; cos(h) is re-seeded
// from a library sincos every 16 steps and rotated by dh in between.
constexpr int kMomWarps = 4;
// zenith-angle class boundaries of the OPTAM lookup (physics.py:255-258) as cosines: entry j is the
// largest double c with int(acos(c) / RAD_PER_DEG) >= 70 + j under the HOST's libm -- the one the
// reference runs on -- found by bisection in optam_thresholds(); padded to 32 with -1 (never met)
struct OptamThr { double c[32]; };
__global__ void __launch_bounds__(kMomWarps * 32)
solar_moments_kernel(int n_rows, const double *__restrict__ g_cosegeom,
                     const double *__restrict__ g_sinegeom, const double *__restrict__ g_hss,
                     const double *__restrict__ g_cza0, const int32_t *__restrict__ g_nstep,
                     const double *__restrict__ g_beam, msb_solar_tables tb,
                     const __grid_constant__ OptamThr thr) {
  __shared__ double etab[kExpTab];
  __shared__ double s_thr[32];
  __shared__ double s_optam[21];
  exp_table_init(etab);
  if (threadIdx.x < 32) s_thr[threadIdx.x] = thr.c[threadIdx.x];
  if (threadIdx.x < 21) s_optam[threadIdx.x] = c_optam[threadIdx.x];
  __syncthreads();
  const int w = blockIdx.x * kMomWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= n_rows) return;
  const int i = w % 365, r = w / 365;
  const double cosegeom = g_cosegeom[w], sinegeom = g_sinegeom[w], hss = g_hss[w];
  const double cza_first = g_cza0[w];
  const int n = g_nstep[w];
  const double dh = kSwRadDt / kSecPerRad;
  const double kLnTBase = -0.13926206733350766;  // ln(0.87)
  const double beam = g_beam[i];
  const double start = -hss;
  const double cd = g_beam[365], sd = g_beam[366];

  double m_loc[kTT + 1];   // Taylor moments of trans**am, [kTT] = sum of dir_flat_topa
#pragma unroll
  for (int m = 0; m <= kTT; ++m) m_loc[m] = 0.0;
  const int per = (n + 31) / 32;
  const int k_lo = min(lane * per, n), k_hi = min(k_lo + per, n);
  double ch = 1.0, sh = 0.0;
  for (int k = k_lo; k < k_hi; ++k) {
    if (((k - k_lo) & 15) == 0) sincos(__dadd_rn(start, __dmul_rn((double)k, dh)), &sh, &ch);
    const double cza = (k == 0) ? cza_first : __dadd_rn(__dmul_rn(cosegeom, ch), sinegeom);
    if (cza > 0.0) {
      const double dfa = beam * cza;                       // physics.py:252
      double am = rcp_nr(cza + 0.0000001);
      if (am > 2.9) {
        // ami = min(max(int(acos(cza)/RAD_PER_DEG) - 69, 0), 20) = number of class boundaries at or
        // above cza (they descend): branch-free binary search, no acos
        int ami = 0;
#pragma unroll
        for (int st = 16; st > 0; st >>= 1)
          if (cza <= s_thr[ami + st - 1]) ami += st;
        am = s_optam[ami];
      }
      m_loc[kTT] += dfa;
      // trans**am = exp(am*p*ln(TBASE)); expand in (p - p0): sum_m dfa*e^{a p0} a^m/m! (p-p0)^m
      // four interleaved chains term[m+4] = term[m] * a^4 / ((m+1)(m+2)(m+3)(m+4)): a single chain
      // of kTT dependent multiplies would expose kTT x 8 cycles of DMUL latency per step
      // the power sums sum dfa e^{a p0} a^m are accumulated raw (one multiply and one add per
      // term); the 1/m! go on once per row after the reduction
      const double a = am * kLnTBase;
      const double a2 = a * a, a4 = a2 * a2;
      double t[4];
      t[0] = exp_acc(a * MSB_TT_P0, etab) * dfa;
      t[1] = t[0] * a;
      t[2] = t[0] * a2;
      t[3] = t[1] * a2;
#pragma unroll
      for (int m = 0; m < kTT; ++m) {
        m_loc[m] += t[m & 3];
        t[m & 3] *= a4;
      }
    }
    const double cn = fma(ch, cd, -(sh * sd)), sn = fma(sh, cd, ch * sd);   // rotate by dh
    ch = cn; sh = sn;
  }
  // butterfly reduction: every lane ends with the totals (fixed order -> deterministic)
#pragma unroll
  for (int m = 0; m <= kTT; ++m)
    for (int o2 = 16; o2 > 0; o2 >>= 1) m_loc[m] += __shfl_xor_sync(0xffffffffu, m_loc[m], o2);
  const double total = m_loc[kTT];
  const double daylength = fmin(2.0 * hss * kSecPerRad, kSecPerDay);  // physics.py:236
  const bool has_day = daylength != 0.0;
  // tt_max0(p) = sum_trans/sum_flat_potrad (physics.py:275); 0 when there is no daylight
  double mine = 0.0, inv_fact = 1.0;
#pragma unroll
  for (int m = 0; m < kTT; ++m) {
    if (m > 0) inv_fact *= 1.0 / (double)m;      // compile-time 1/m!
    mine = (lane == m) ? m_loc[m] * inv_fact : mine;
  }
  if (lane < kTT) tb.tt_coef[(size_t)w * kTT + lane] = (has_day && total > 0.0) ? mine / total : 0.0;
  if (lane == 0) {
    const double potrad = has_day ? total / daylength : 0.0;            // physics.py:276-280
    const size_t s = (size_t)r * kRows + i;
    tb.daylength[s] = daylength;
    tb.potrad[s] = potrad;
    if (i == 364) {  // row 365 copies row 364, physics.py:281-284
      tb.daylength[s + 1] = daylength;
      tb.potrad[s + 1] = potrad;
    }
  }
}

}  // namespace msb

using namespace msb;

extern "C" size_t msb_solar_prelude_bytes(int n_lat) {
  size_t n = (size_t)(n_lat > 0 ? n_lat : 0) * 365;
  return (4 * n + (n + 1) / 2 + 365 + 2) * sizeof(double) + 64;
}

extern "C" int msb_solar_table_sizes(int n_lat, size_t *small_elems, size_t *q_elems,
                                     size_t *tt_elems) {
  if (n_lat <= 0) { set_error("msb_solar_table_sizes: n_lat must be > 0"); return MSB_E_ARG; }
  if (small_elems) *small_elems = (size_t)n_lat * kRows;
  if (q_elems) *q_elems = (size_t)n_lat * 365 * kQLd + MSB_Q_PAD;
  if (tt_elems) *tt_elems = (size_t)n_lat * 365 * kTT;
  return MSB_OK;
}

extern "C" size_t msb_solar_q_win_ld(int time_step) {
  if (time_step <= 0 || time_step >= 1440 || 1440 % time_step) return 0;
  return (size_t)q_win_ld(time_step);
}

extern "C" size_t msb_solar_q_win_elems(int n_lat, int time_step) {
  const size_t ld = msb_solar_q_win_ld(time_step);
  return ld && n_lat > 0 ? (size_t)n_lat * 365 * ld + MSB_Q_PAD : 0;
}

extern "C" int msb_solar_tables_build(int n_lat, const double *lat_deg_host, void *host_ws,
                                      void *dev_ws, const msb_solar_tables *tb,
                                      int n_host_threads, void *stream) {
  if (n_lat <= 0 || !lat_deg_host || !host_ws || !dev_ws || !tb) {
    set_error("msb_solar_tables_build: null/empty argument");
    return MSB_E_ARG;
  }
  if (!tb->daylength || !tb->potrad || !tb->rise_min || !tb->set_min || !tb->q || !tb->tt_coef ||
      tb->n_lat != n_lat) {
    set_error("msb_solar_tables_build: table buffers missing or n_lat mismatch");
    return MSB_E_ARG;
  }
  if (tb->q_win && msb_solar_q_win_ld(tb->q_win_step) == 0) {
    set_error("msb_solar_tables_build: q_win_step %d must divide 1440 and be below it", tb->q_win_step);
    return MSB_E_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PreludeView hv = prelude_view(host_ws, n_lat);
  int nt = std::max(1, std::min(n_host_threads, n_lat));
  if (nt == 1) {
    prelude_rows(lat_deg_host, 0, n_lat, hv);
  } else {
    std::vector<std::thread> pool;
    int per = (n_lat + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
      int r0 = t * per, r1 = std::min(n_lat, r0 + per);
      if (r0 >= r1) break;
      pool.emplace_back(prelude_rows, lat_deg_host, r0, r1, hv);
    }
    for (auto &th : pool) th.join();
  }
  // latitude-independent scalars of the hour-angle loop, with this host's libm like the rest of K0:
  // the beam of every day (physics.py:238-239) and the rotation by one 30 s step
  for (int i = 0; i < 365; ++i)
    hv.beam[i] = (1368.0 + 45.5 * std::sin((2.0 * M_PI * i / kDaysPerYear) + 1.7)) * kSwRadDt;
  hv.beam[365] = std::cos(kSwRadDt / kSecPerRad);
  hv.beam[366] = std::sin(kSwRadDt / kSecPerRad);
  size_t bytes = msb_solar_prelude_bytes(n_lat) - 64;
  MSB_CUDA(cudaMemcpyAsync(dev_ws, host_ws, bytes, cudaMemcpyHostToDevice, st));
  PreludeView dv = prelude_view(dev_ws, n_lat);
  unsigned grid = 365u * (unsigned)n_lat;
  solar_rows_kernel<<<grid, kK1Threads, 0, st>>>(n_lat, dv.cosegeom, dv.sinegeom, dv.hss, dv.cza0,
                                                 dv.nstep, dv.beam, *tb);
  OptamThr thr;
  optam_thresholds(thr.c);
  solar_moments_kernel<<<(grid + kMomWarps - 1) / kMomWarps, kMomWarps * 32, 0, st>>>(
      (int)grid, dv.cosegeom, dv.sinegeom, dv.hss, dv.cza0, dv.nstep, dv.beam, *tb, thr);
  MSB_CUDA(cudaGetLastError());
  MSB_CUDA(cudaMemsetAsync(tb->q + (size_t)n_lat * 365 * kQLd, 0, MSB_Q_PAD * sizeof(double), st));
  if (tb->q_win)
    MSB_CUDA(cudaMemsetAsync(tb->q_win + (size_t)n_lat * 365 * q_win_ld(tb->q_win_step), 0,
                             MSB_Q_PAD * sizeof(double), st));
  return MSB_OK;
}
