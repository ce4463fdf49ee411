// Native consumer of output tiles: the host side of the streaming writer.
//
// Reference: MetSim allocates (time, lat, lon) arrays of NaN, scatters every cell's series into
// them (metsim/metsim.py:581-586) and hands them to xarray / netCDF4 (metsim/metsim.py:363-454).
// Here the rows of a finished tile arrive in pinned host memory from msb_run_host_ctx() (on_tile
// hook) at PCIe rate, so the scatter [step][cell] -> [step][lat*lon], the byte order of the classic
// NetCDF format and the pwrite() are done by a pool of host threads, each on its own range of
// steps with its own staging buffer.  Host code only (no kernels); it lives in the same C-ABI
// library so that a reference-side binding has one thing to load.
#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include <unistd.h>

#include "msb_common.cuh"

struct msb_sink {
  msb_sink_desc d;
  std::vector<int64_t> index;        // private copy of cell_index (empty = identity)
  std::vector<std::thread> pool;
  std::vector<std::vector<char>> stage;   // per thread: one dense piece, or chunk_steps rows of the grid, NaN-prefilled
  int64_t chunk_steps = 1;           // masked grids: whole rows per chunk (the scatter needs the NaN background)
  int64_t piece_elems = 0;           // dense grids (> 0): chunks are flat runs of this many values
  int64_t n_total = 0;               // dense: values of the current job
  // one job at a time: rows [0, n_steps) of src to file_offset
  std::mutex call_mu;                // serialises msb_sink_write() callers (mu is released while waiting)
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  const char *src = nullptr;
  int64_t n_steps = 0, file_offset = 0;
  std::atomic<int64_t> next_chunk{0};
  int64_t n_chunks = 0;
  uint64_t generation = 0;
  int running = 0;
  bool stop = false;
  int error = 0;                     // errno of the first failed pwrite of the current job
  std::atomic<uint64_t> checksum{0}; // sum of the stored 32/64-bit patterns of every cell value consumed
  std::atomic<uint64_t> bytes{0};
};

namespace {

inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
inline uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }

// The dense rows (no mask: CONUS, or any shard whose cells are the whole grid in order) are plain
// streaming loops; the AVX2 clone turns the byte swap into one shuffle per 8 values, which is what
// lets a few threads per GPU keep up with the PCIe stream.  The dispatch is the loader's (ifunc).
// A dense job is ONE flat run of values on both sides (row stride = n_cells = n_grid), so it is cut
// into pieces of kPieceBytes regardless of the row length: the staging of a thread then stays in
// its core's L2 and never reaches DRAM, which leaves the host memory system with the DMA writes
// of the tile and one read of it -- with row-sized (MBs) staging the write-allocate + write-back
// of the staging doubled that traffic and capped the pool near 55 GB/s on a 32-core host.
constexpr int64_t kPieceBytes = 256 << 10;
#if defined(__x86_64__) && defined(__GNUC__)
#define MSB_SIMD_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define MSB_SIMD_CLONES
#endif

MSB_SIMD_CLONES uint64_t dense_row32(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int64_t n, bool swap) {
  uint64_t sum = 0;
  if (swap) for (int64_t c = 0; c < n; ++c) { const uint32_t v = in[c]; sum += v; out[c] = bswap32(v); }
  else for (int64_t c = 0; c < n; ++c) { const uint32_t v = in[c]; sum += v; out[c] = v; }
  return sum;
}
MSB_SIMD_CLONES uint64_t dense_row64(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, int64_t n, bool swap) {
  uint64_t sum = 0;
  if (swap) for (int64_t c = 0; c < n; ++c) { const uint64_t v = in[c]; sum += v; out[c] = bswap64(v); }
  else for (int64_t c = 0; c < n; ++c) { const uint64_t v = in[c]; sum += v; out[c] = v; }
  return sum;
}

// masked grid: rows of n_cells values into rows of n_grid, cell c at index[c]
template <typename U>
uint64_t scatter_rows(const msb_sink &s, const U *src, U *dst, int64_t steps) {
  const int64_t n = s.d.n_cells, g = s.d.n_grid;
  const bool swap = s.d.byteswap != 0;
  const int64_t *idx = s.index.data();
  uint64_t sum = 0;
  for (int64_t t = 0; t < steps; ++t) {
    const U *in = src + t * n;
    U *out = dst + t * g;
    for (int64_t c = 0; c < n; ++c) {
      const U v = in[c];
      sum += v;
      out[idx[c]] = swap ? (sizeof(U) == 4 ? (U)bswap32((uint32_t)v) : (U)bswap64((uint64_t)v)) : v;
    }
  }
  return sum;
}

void worker(msb_sink *s, int tid) {
  uint64_t seen = 0;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(s->mu);
      s->cv_work.wait(lk, [&] { return s->stop || s->generation != seen; });
      if (s->stop) return;
      seen = s->generation;
    }
    const int64_t eb = s->d.elem_bytes;
    const int64_t row_in = (int64_t)s->d.n_cells * eb, row_out = s->d.n_grid * eb;
    char *buf = s->stage[tid].data();
    uint64_t sum = 0, written = 0;
    int err = 0;
    for (;;) {
      const int64_t c = s->next_chunk.fetch_add(1);
      if (c >= s->n_chunks) break;
      int64_t left, off;               // bytes of this chunk in the staging, their place in the file
      if (s->piece_elems > 0) {
        const int64_t e0 = c * s->piece_elems, cnt = std::min(s->piece_elems, s->n_total - e0);
        const char *in = s->src + e0 * eb;
        const bool swap = s->d.byteswap != 0;
        sum += eb == 4 ? dense_row32(reinterpret_cast<const uint32_t *>(in), reinterpret_cast<uint32_t *>(buf), cnt, swap)
                       : dense_row64(reinterpret_cast<const uint64_t *>(in), reinterpret_cast<uint64_t *>(buf), cnt, swap);
        left = cnt * eb; off = s->file_offset + e0 * eb;
      } else {
        const int64_t t0 = c * s->chunk_steps;
        const int64_t steps = std::min(s->chunk_steps, s->n_steps - t0);
        const char *in = s->src + t0 * row_in;
        sum += eb == 4 ? scatter_rows<uint32_t>(*s, reinterpret_cast<const uint32_t *>(in),
                                                reinterpret_cast<uint32_t *>(buf), steps)
                       : scatter_rows<uint64_t>(*s, reinterpret_cast<const uint64_t *>(in),
                                                reinterpret_cast<uint64_t *>(buf), steps);
        left = steps * row_out; off = s->file_offset + t0 * row_out;
      }
      if (s->d.fd >= 0) {
        const char *p = buf;
        while (left > 0 && !err) {
          const ssize_t w = pwrite(s->d.fd, p, (size_t)left, (off_t)off);
          if (w < 0) { if (errno == EINTR) continue; err = errno ? errno : EIO; break; }
          left -= w; off += w; p += w; written += (uint64_t)w;
        }
      }
    }
    s->checksum.fetch_add(sum);
    s->bytes.fetch_add(written);
    std::lock_guard<std::mutex> lk(s->mu);
    if (err && !s->error) s->error = err;
    if (--s->running == 0) s->cv_done.notify_all();
  }
}

}  // namespace

using namespace msb;

extern "C" int msb_sink_create(const msb_sink_desc *d, msb_sink **out) {
  if (!d || !out) { set_error("msb_sink_create: NULL argument"); return MSB_E_ARG; }
  *out = nullptr;
  if (d->n_cells <= 0 || d->n_grid < d->n_cells || (d->elem_bytes != 4 && d->elem_bytes != 8)) {
    set_error("msb_sink_create: need 0 < n_cells <= n_grid and elem_bytes 4 or 8");
    return MSB_E_ARG;
  }
  if (!d->cell_index && d->n_grid != d->n_cells) {
    set_error("msb_sink_create: a grid larger than the cell count needs cell_index");
    return MSB_E_ARG;
  }
  msb_sink *s = new (std::nothrow) msb_sink;
  if (!s) { set_error("msb_sink_create: out of memory"); return MSB_E_NOMEM; }
  s->d = *d;
  s->d.cell_index = nullptr;
  if (d->cell_index) {
    bool identity = d->n_grid == d->n_cells;
    std::vector<char> hit((size_t)d->n_grid, 0);
    for (int64_t c = 0; c < d->n_cells; ++c) {
      const int64_t g = d->cell_index[c];
      if (g < 0 || g >= d->n_grid || hit[(size_t)g]) {
        set_error("msb_sink_create: cell_index[%lld] = %lld is outside the grid or repeated", (long long)c, (long long)g);
        delete s;
        return MSB_E_ARG;
      }
      hit[(size_t)g] = 1;
      identity = identity && g == c;
    }
    if (!identity) s->index.assign(d->cell_index, d->cell_index + d->n_cells);
  }
  unsigned hw = std::thread::hardware_concurrency();
  int nt = d->n_threads > 0 ? d->n_threads : (int)(hw ? hw : 1);
  if (nt > 256) nt = 256;
  s->d.n_threads = nt;
  // staging per thread: a dense grid is cut into flat pieces; a masked one needs whole grid rows
  // (as few as fit the same budget, at least one)
  const int64_t row_out = d->n_grid * d->elem_bytes;
  if (s->index.empty()) s->piece_elems = kPieceBytes / d->elem_bytes;
  else s->chunk_steps = std::max<int64_t>(1, (2 * kPieceBytes) / std::max<int64_t>(1, row_out));
  try {
    s->stage.resize(nt);
    for (auto &b : s->stage) {
      if (s->piece_elems > 0) { b.resize((size_t)kPieceBytes); continue; }   // every byte is overwritten per piece
      b.resize((size_t)(s->chunk_steps * row_out));
      // cells outside the mask hold NaN (metsim.py:584); written once, the scatter never touches them
      if (d->elem_bytes == 4) {
        float q = std::numeric_limits<float>::quiet_NaN();
        uint32_t u;
        std::memcpy(&u, &q, 4);
        if (d->byteswap) u = bswap32(u);
        uint32_t *p = reinterpret_cast<uint32_t *>(b.data());
        for (int64_t i = 0; i < s->chunk_steps * d->n_grid; ++i) p[i] = u;
      } else {
        double q = std::numeric_limits<double>::quiet_NaN();
        uint64_t u;
        std::memcpy(&u, &q, 8);
        if (d->byteswap) u = bswap64(u);
        uint64_t *p = reinterpret_cast<uint64_t *>(b.data());
        for (int64_t i = 0; i < s->chunk_steps * d->n_grid; ++i) p[i] = u;
      }
    }
    for (int t = 0; t < nt; ++t) s->pool.emplace_back(worker, s, t);
  } catch (...) {
    set_error("msb_sink_create: could not start %d threads / allocate their staging", nt);
    msb_sink_destroy(s);
    return MSB_E_NOMEM;
  }
  *out = s;
  return MSB_OK;
}

extern "C" int msb_sink_write(msb_sink *s, const void *src, int64_t n_steps, int64_t file_offset) {
  if (!s || !src || n_steps < 0 || file_offset < 0) { set_error("msb_sink_write: bad argument"); return MSB_E_ARG; }
  if (n_steps == 0) return MSB_OK;
  std::lock_guard<std::mutex> one_caller(s->call_mu);
  std::unique_lock<std::mutex> lk(s->mu);
  s->src = reinterpret_cast<const char *>(src);
  s->n_steps = n_steps;
  s->file_offset = file_offset;
  s->n_total = n_steps * (int64_t)s->d.n_cells;
  s->n_chunks = s->piece_elems > 0 ? (s->n_total + s->piece_elems - 1) / s->piece_elems
                                   : (n_steps + s->chunk_steps - 1) / s->chunk_steps;
  s->next_chunk.store(0);
  s->error = 0;
  s->running = (int)s->pool.size();
  ++s->generation;
  s->cv_work.notify_all();
  s->cv_done.wait(lk, [&] { return s->running == 0; });
  if (s->error) {
    set_error("msb_sink_write: pwrite failed: %s", strerror(s->error));
    return MSB_E_ARG;
  }
  return MSB_OK;
}

extern "C" uint64_t msb_sink_checksum(const msb_sink *s) { return s ? s->checksum.load() : 0; }
extern "C" uint64_t msb_sink_bytes_written(const msb_sink *s) { return s ? s->bytes.load() : 0; }

extern "C" void msb_sink_destroy(msb_sink *s) {
  if (!s) return;
  {
    std::lock_guard<std::mutex> lk(s->mu);
    s->stop = true;
  }
  s->cv_work.notify_all();
  for (auto &t : s->pool) if (t.joinable()) t.join();
  delete s;
}
