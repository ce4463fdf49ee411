"""Batched, device-resident driver of the CUDA path: the B200 form of ``MetSim.run_slice``.

Reference call chain replaced here (metsim/metsim.py):
``run_slice`` :456-493 -> per-cell ``wrap_run_cell`` :693-796 -> ``solar_geom`` ->
``method.run`` -> ``disaggregate``.  One :class:`Engine` holds one shard (a contiguous chunk of
active cells) on one GPU; PyTorch supplies device memory and the stream, every kernel is in
``csrc/libmetsim_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import DAILY_VARS, SUB_VARS

KELVIN = 273.15
DEFAULT_OUT_VARS = ("temp", "prec", "shortwave", "longwave", "vapor_pressure", "rel_humid")


def unit_scale_offset(var: str, units: str | None, ts: int):
    """(scale, offset) equivalent of metsim/units.py:3-59 for one variable."""
    if units is None:
        return 1.0, 0.0
    table = {
        "prec": {"mm timestep-1": (1.0, 0.0), "mm h-1": (1.0 / (ts / 60.0), 0.0),
                 "mm s-1": (1.0 / (ts * 60.0), 0.0)},
        "pet": {"mm timestep-1": (1.0, 0.0), "mm h-1": (1.0 / (ts / 60.0), 0.0),
                "mm s-1": (1.0 / (ts * 60.0), 0.0)},
        "shortwave": {"W m-2": (1.0, 0.0)},
        "longwave": {"W m-2": (1.0, 0.0)},
        "temp": {"C": (1.0, 0.0), "K": (1.0, KELVIN)},
        "t_min": {"C": (1.0, 0.0), "K": (1.0, KELVIN)},
        "t_max": {"C": (1.0, 0.0), "K": (1.0, KELVIN)},
        "vapor_pressure": {"Pa": (1.0, 0.0), "hPa": (1.0 / 100.0, 0.0), "kPa": (1.0 / 1000.0, 0.0)},
        "air_pressure": {"kPa": (1.0, 0.0), "hPa": (100.0, 0.0), "Pa": (1000.0, 0.0)},
        "tskc": {"fraction": (1.0, 0.0), "%": (100.0, 0.0)},
        "rel_humid": {"%": (1.0, 0.0), "fraction": (1.0 / 100.0, 0.0)},
        "spec_humid": {"g g-1": (1.0, 0.0)},
        "wind": {"m s-1": (1.0, 0.0)},
        "daylength": {"s": (1.0, 0.0)},
    }
    try:
        return table[var][units]
    except KeyError:
        raise ValueError(f"Could not find unit conversion for {var} to {units}!")


def calendar_vectors(start, n_days: int):
    """1-based day-of-year and month per day (standard calendar), as int32 numpy arrays."""
    d0 = np.datetime64(start, "D")
    days = d0 + np.arange(n_days)
    years = days.astype("datetime64[Y]")
    doy = (days - years.astype("datetime64[D]")).astype(np.int64) + 1
    month = days.astype("datetime64[M]").astype(np.int64) % 12 + 1
    return doy.astype(np.int32), month.astype(np.int32)


class Engine:
    """One shard of cells on one B200.

    Arrays are cell-fastest: daily ``[D][N]``, state ``[90][N]``, monthly ``[12][N]``.  Inputs may
    be numpy arrays (copied through pinned memory) or CUDA tensors (used in place).
    """

    def __init__(self, params: dict | None = None, device: str | int | torch.device = "cuda:0",
                 tuning: dict | None = None):
        """``tuning``: msb_tuning overrides (``_lib.TUNING_KEYS``; tests pin the benchmarked launch
        geometry with it); unset keys fall back to the MSB_* environment variables, then to the
        library's own choice."""
        if not torch.cuda.is_available():
            raise RuntimeError("metsim_b200 needs a CUDA device; there is no CPU fallback")
        self.lib = _lib.lib()
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        self.user_params = dict(params or {})
        self.params = _lib.make_params(self.user_params)
        self.tuning = _lib.make_tuning(tuning)
        self.single_pass_budget = 24 << 30    # bytes of candidate planes the single-pass daily form may take
        self.stream = torch.cuda.Stream(device=self.device)
        self._t = {}
        self._tables = None
        self._daily = None
        self._ws = None
        self._solar_bufs = None
        self._solar_done = None
        self.n_cells = self.n_days = 0
        self.order = None

    # ------------------------------------------------------------------ helpers
    def _dev(self, a, dtype, shape=None):
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device, dtype=dtype).contiguous()
        else:
            h = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
            t = h.pin_memory().to(self.device, non_blocking=True)
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {shape}, got {tuple(t.shape)}")
        return t

    @staticmethod
    def _ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    @property
    def tpd(self) -> int:
        ts = self.params.time_step
        return 1440 // ts if ts < 1440 else 1

    # ------------------------------------------------------------------ inputs
    def load(self, *, lat, lon, elev, doy, month, t_min, t_max, prec, wind=None,
             state_t_min, state_t_max, state_prec, t_pk12=None, dur12=None, t_end=None,
             given_shortwave=None, given_vapor_pressure=None, given_tskc=None,
             given_longwave=None, order=None):
        """``given_*`` are the forcing columns of the 'passthrough' method
        (metsim/methods/passthrough.py:28-46), daily ``[D][N]``; shortwave is required there.

        ``order``: optional engine cell order (a permutation, engine position -> caller position,
        e.g. :func:`metsim_b200.shard.patch_order` for masked grids).  Every per-cell input is
        gathered into that order once, on the device; :meth:`disaggregate` / :meth:`daily` then work
        and return in ENGINE order (``self.order`` maps back), :meth:`run` returns the caller's
        order unless asked otherwise."""
        f64, i32 = torch.float64, torch.int32
        self.order = None
        if order is not None:
            o = np.asarray(order, dtype=np.int64)
            if o.ndim != 1 or not np.array_equal(np.sort(o), np.arange(o.size)):
                raise ValueError("order must be a permutation of the cells")
            with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
                self.order = torch.as_tensor(o, device=self.device)
            pick = lambda a: None if a is None else (
                a.to(self.device).index_select(a.dim() - 1, self.order) if isinstance(a, torch.Tensor)
                else np.ascontiguousarray(np.asarray(a)[..., o]))
            with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
                (lat, lon, elev, t_min, t_max, prec, wind, state_t_min, state_t_max, state_prec, t_pk12, dur12,
                 t_end, given_shortwave, given_vapor_pressure, given_tskc, given_longwave) = (
                    pick(a) for a in (lat, lon, elev, t_min, t_max, prec, wind, state_t_min, state_t_max,
                                      state_prec, t_pk12, dur12, t_end, given_shortwave,
                                      given_vapor_pressure, given_tskc, given_longwave))
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            t = self._t
            t["t_min"] = self._dev(t_min, f64)
            d, n = t["t_min"].shape
            self.n_days, self.n_cells = int(d), int(n)
            t["t_max"] = self._dev(t_max, f64, (d, n))
            t["prec"] = self._dev(prec, f64, (d, n))
            t["wind"] = self._dev(wind, f64, (d, n))
            for k, v in zip(_lib.GIVEN_KEYS, (given_shortwave, given_vapor_pressure, given_tskc,
                                              given_longwave)):
                t[k] = self._dev(v, f64, (d, n))
            for k, v in (("st_t_min", state_t_min), ("st_t_max", state_t_max), ("st_prec", state_prec)):
                t[k] = self._dev(v, f64, (90, n))
            t["t_pk12"] = self._dev(t_pk12, f64, (12, n))
            t["dur12"] = self._dev(dur12, f64, (12, n))
            t["t_end"] = self._dev(t_end, f64, (2, n))
            t["lon"] = self._dev(lon, f64, (n,))
            t["elev"] = self._dev(elev, f64, (n,))
            t["doy"] = self._dev(doy, i32, (d,))
            t["month"] = self._dev(month, i32, (d,))
            # unique latitudes -> rows of the solar tables
            if isinstance(lat, torch.Tensor):
                u, inv = torch.unique(lat.to(self.device, f64), return_inverse=True)
                self._ulat = u.cpu().numpy().astype(np.float64)
                t["lat_row"] = inv.to(i32).contiguous()
            else:
                u, inv = np.unique(np.asarray(lat, dtype=np.float64), return_inverse=True)
                self._ulat = np.ascontiguousarray(u)
                t["lat_row"] = self._dev(inv.astype(np.int32), i32, (n,))
            self.sw_sectors_per_load = self._sw_sectors(t["lon"], t["lat_row"])
        self._tables = None
        self._daily = None
        return self

    def _forcing(self) -> _lib.Forcing:
        t = self._t
        f = _lib.Forcing()
        f.n_cells, f.n_days, f.ld = self.n_cells, self.n_days, self.n_cells
        for name in ("lon", "elev", "lat_row", "doy", "month", "t_min", "t_max", "prec", "wind",
                     "st_t_min", "st_t_max", "st_prec", "t_pk12", "dur12", "t_end") + _lib.GIVEN_KEYS:
            setattr(f, name, self._ptr(t.get(name)))
        return f

    @staticmethod
    def _sw_sectors(lon, lat_row) -> float:
        """Mean number of 32-byte sectors of the plain table q that one warp (32 consecutive cells)
        touches per load of the shortwave kernel: the distinct (latitude row, tiny offset // 4) among
        its lanes.  CONUS 1/16 deg: ~4.5 (neighbours share entries); global 0.5 deg land: ~30."""
        n = lon.numel()
        if n < 64:
            return 0.0
        w = torch.remainder(lon + 180.0, 360.0) - 180.0
        key = (lat_row.to(torch.int64) << 12) + (torch.trunc(w / 360.0 * 2880.0).to(torch.int64) + 1440 >> 2)
        key = key[: n // 32 * 32].view(-1, 32).sort(dim=1).values
        return float(((key[:, 1:] != key[:, :-1]).sum(dim=1) + 1).double().mean().item())

    def use_window_major(self) -> bool:
        """Build / use the window-major copy of the table (msb_solar_tables.q_win)?  tuning.sw_table:
        0 = when a warp's cells touch more than 8 sectors of the plain table per load, 1 = never,
        2 = always (sub-daily time steps only)."""
        if self.params.time_step >= 1440:
            return False
        mode = self.tuning.sw_table
        return mode == 2 or (mode == 0 and self.sw_sectors_per_load > 8.0)

    # ------------------------------------------------------------------ solar tables
    def build_solar(self, n_host_threads: int | None = None):
        n_lat = len(self._ulat)
        win_step = self.params.time_step if self.use_window_major() else 0
        if self._solar_bufs is None or self._solar_bufs[0] != (n_lat, win_step):
            # device tables + pinned / device prelude workspaces, allocated once per latitude count
            small, q, tt = C.c_size_t(), C.c_size_t(), C.c_size_t()
            _lib.check(self.lib.msb_solar_table_sizes(n_lat, C.byref(small), C.byref(q), C.byref(tt)))
            with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
                opts = dict(dtype=torch.float64, device=self.device)
                tb = {k: torch.empty(small.value, **opts)
                      for k in ("daylength", "potrad", "rise_min", "set_min")}
                tb["q"] = torch.empty(q.value, **opts)
                tb["tt_coef"] = torch.empty(tt.value, **opts)
                if win_step:
                    tb["q_win"] = torch.empty(self.lib.msb_solar_q_win_elems(n_lat, win_step), **opts)
                nbytes = self.lib.msb_solar_prelude_bytes(n_lat)
                host_ws = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
                dev_ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            st = _lib.SolarTables()
            st.n_lat = n_lat
            st.q_win_step = win_step
            for k, v in tb.items():
                setattr(st, k, self._ptr(v))
            self._solar_bufs = ((n_lat, win_step), st, tb, host_ws, dev_ws)
        _, st, tb, host_ws, dev_ws = self._solar_bufs
        nthr = n_host_threads or min(os.cpu_count() or 1, 32)
        if self._solar_done is not None:
            self._solar_done.synchronize()   # the previous upload out of the pinned workspace
        with torch.cuda.device(self.device):
            _lib.check(self.lib.msb_solar_tables_build(
                n_lat, self._ulat.ctypes.data_as(C.c_void_p), C.c_void_p(host_ws.data_ptr()),
                C.c_void_p(dev_ws.data_ptr()), C.byref(st), nthr,
                C.c_void_p(self.stream.cuda_stream)))
            self._solar_done = torch.cuda.Event()
            self._solar_done.record(self.stream)
        self._tables = (st, tb, host_ws, dev_ws)
        return tb

    # ------------------------------------------------------------------ daily MTCLIM
    def daily(self, want=DAILY_VARS, single_pass: bool | None = None):
        """Daily MTCLIM pass.  ``want`` names the daily columns to return.  When nothing beyond
        ``tskc`` is wanted (what the disaggregation alone needs) the single-pass form of
        ``csrc/daily.cu`` runs: candidate planes of vapour pressure / shortwave after 4, 5, 6
        evaluations instead of a replay; ``single_pass=False`` forces the two-pass form."""
        if self._tables is None:
            self.build_solar()
        st = self._tables[0]
        n, d = self.n_cells, self.n_days
        want = set(want)
        cand_bytes = 2 * _lib.DAILY_CAND * d * n * 8
        single = (want <= {"tskc"} and self.params.method == 0 and not self.tuning.daily_two_pass
                  and self.params.time_step < 1440 and cand_bytes <= self.single_pass_budget)
        if single_pass is not None:
            if single_pass and not single:
                raise ValueError("the single-pass daily form needs method mtclim, a sub-daily time step, "
                                 "no daily column beyond tskc and candidate planes within the budget")
            single = bool(single_pass)
        need = ({"tskc"} if single else {"tskc", "vapor_pressure", "shortwave"}) | want
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            if (single and self._daily is not None and self._daily[0].cand
                    and self._daily[4] == (d, n)):
                out = dict(self._daily[1])                  # reuse the planes of the previous call
                n_iter, scratch, status = out.pop("n_iter"), self._daily[2], self._daily[3]
                cand, sel, lst = self._daily[5]
            else:
                self._daily = None
                out = {k: torch.empty((d, n), dtype=torch.float64, device=self.device)
                       for k in DAILY_VARS if k in need}
                n_iter = torch.zeros(n, dtype=torch.int32, device=self.device)
                nbytes = self.lib.msb_daily_scratch_bytes_chunk(n, d, self.tuning.daily_chunk_len)
                scratch = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device=self.device)
                status = torch.zeros(4, dtype=torch.int32, device=self.device)
                cand = sel = lst = None
                if single:
                    cand = torch.empty((2 * _lib.DAILY_CAND, d, n), dtype=torch.float64, device=self.device)
                    sel = torch.zeros(n, dtype=torch.int32, device=self.device)
                    lst = torch.zeros(n + 1, dtype=torch.int32, device=self.device)
            dl = _lib.Daily()
            for i, k in enumerate(DAILY_VARS):
                dl.var[i] = out[k].data_ptr() if k in out else None
            dl.n_iter, dl.partial, dl.status = (self._ptr(n_iter), self._ptr(scratch), self._ptr(status))
            dl.cand, dl.cand_sel, dl.cand_list = self._ptr(cand), self._ptr(sel), self._ptr(lst)
            dl.tune = self.tuning
            f = self._forcing()
            assert bool(self.lib.msb_daily_single_pass(C.byref(self.params), C.byref(f), C.byref(dl))) == single
            _lib.check(self.lib.msb_mtclim_daily(C.byref(self.params), C.byref(f), C.byref(st),
                                                 C.byref(dl), C.c_void_p(self.stream.cuda_stream)))
        out["n_iter"] = n_iter
        self._daily = (dl, out, scratch, status, (d, n), (cand, sel, lst))
        self._publish()
        return out

    def _publish(self):
        """The kernels run on the engine's own stream; torch code that follows on the caller's current
        stream (``.cpu()``, arithmetic on the returned tensors) is ordered after them here -- a
        stream-side wait, the host does not block."""
        cur = torch.cuda.current_stream(self.device)
        if cur != self.stream:
            cur.wait_stream(self.stream)

    def daily_for_disagg(self):
        """The daily vapour pressure and shortwave the disaggregation reads, as ``[D][N]`` tensors
        (gathered from the candidate planes when the single-pass form ran)."""
        dl, out, _, _, _, (cand, sel, _) = self._daily
        if cand is None:
            return out["vapor_pressure"], out["shortwave"]
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            idx = (2 * sel.long())[None, None, :].expand(1, cand.shape[1], -1)
            vp = torch.gather(cand, 0, idx)[0]
            sw = torch.gather(cand, 0, idx + 1)[0]
        return vp, sw

    def check(self):
        """Synchronise and raise what the reference would have raised (ValueError for t_max<t_min)."""
        if self._daily is None:
            return
        f = self._forcing()
        _lib.check(self.lib.msb_daily_check(C.byref(f), C.byref(self._daily[0]),
                                            C.c_void_p(self.stream.cuda_stream)))

    # ------------------------------------------------------------------ sub-daily
    def disaggregate(self, out_vars=DEFAULT_OUT_VARS, day0: int = 0, day1: int | None = None,
                     out_dtype=torch.float32, units: dict | None = None, out: dict | None = None):
        """Disaggregate days [day0, day1) into ``{var: tensor[(day1-day0)*tpd][N]}``.

        ``out`` may hold preallocated tensors (a ring buffer); ``units`` maps var -> unit string as
        in the reference's ``out_vars[...]['units']``."""
        if self._daily is None:
            self.daily(want=())
        day1 = self.n_days if day1 is None else day1
        ts = self.params.time_step
        if ts >= 1440:
            raise ValueError("time_step 1440 is daily mode: use daily()")
        steps = (day1 - day0) * self.tpd
        n = self.n_cells
        res = {}
        # Rows of the [step][cell] outputs start on 128-byte boundaries: the leading dimension is
        # padded to a multiple of 32 cells and the caller gets [:, :n] views.  (With ld = n and a
        # cell count that is not a multiple of 8, every warp-wide store is split across sectors
        # that the neighbouring row also writes -- measured 1.5x slower on a 125,684-cell shard.)
        ld_out = (n + 31) // 32 * 32
        sd = _lib.SubDaily()
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            for i, v in enumerate(SUB_VARS):
                sd.scale[i], sd.offset[i] = 1.0, 0.0
                if v not in out_vars or (v == "wind" and self._t.get("wind") is None):
                    sd.var[i] = None
                    continue
                if out is not None and v in out:
                    tns = out[v]
                    assert tns.dtype == out_dtype and tns.shape[0] >= steps and tns.shape[1] == n
                    assert tns.stride(1) == 1, "output rows must be contiguous"
                    if res and tns.stride(0) != ld_out:
                        raise ValueError("all preallocated outputs must share one row stride")
                    ld_out = tns.stride(0)
                else:
                    tns = torch.empty((steps, ld_out), dtype=out_dtype, device=self.device)[:, :n]
                    if tns.stride(0) != ld_out:     # mixing with preallocated outputs of another stride
                        raise ValueError("all outputs must share one row stride")
                res[v] = tns
                sd.var[i] = tns.data_ptr()
                sd.scale[i], sd.offset[i] = unit_scale_offset(v, (units or {}).get(v), ts)
            sd.dtype = 4 if out_dtype == torch.float32 else 8
            sd.ld, sd.day0, sd.day1 = ld_out, day0, day1
            if (self.params.method == 1 and self._t.get("given_longwave") is not None
                    and res.get("longwave") is not None):     # longwave rescale staging plane
                need = self.lib.msb_disagg_workspace_bytes_lw(max(n, ld_out), day0, day1, ts)
            else:
                need = self.lib.msb_disagg_workspace_bytes(n, day0, day1)
            if self._ws is None or self._ws.numel() < need:   # grow-only scratch for the day records
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            sd.workspace, sd.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
            sd.tune = self.tuning
            f = self._forcing()
            _lib.check(self.lib.msb_disaggregate(C.byref(self.params), C.byref(f),
                                                 C.byref(self._tables[0]), C.byref(self._daily[0]),
                                                 C.byref(sd), C.c_void_p(self.stream.cuda_stream)))
        self._publish()
        return res

    def _user_order(self, t):
        """[rows][cells in engine order] -> the caller's cell order (an extra pass; a streaming
        consumer takes ``self.order`` into its ``cell_index`` instead)."""
        if self.order is None or t is None or t.dim() == 0:
            return t
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            out = torch.empty_like(t)
            out.index_copy_(t.dim() - 1, self.order, t)
        return out

    def run(self, out_vars=DEFAULT_OUT_VARS, out_dtype=torch.float32, units=None, daily_vars=(),
            engine_order: bool = False):
        """solar tables -> daily -> disaggregate for the whole record; returns tensors (in the caller's
        cell order unless ``engine_order``).

        ``time_step=1440`` is the reference's daily-output mode (metsim/metsim.py:779-792): no
        disaggregation, and ``res['shortwave']`` is the daily average flux
        (``shortwave * daylength / 86400``); the other daily columns are under ``res['daily']``."""
        self.build_solar()
        daily_mode = self.params.time_step >= 1440
        daily = self.daily(want=tuple(daily_vars) + (("shortwave_daily", "vapor_pressure", "shortwave")
                                                     if daily_mode else ()))
        res = {}
        if daily_mode:
            res["shortwave"] = daily["shortwave_daily"]
        else:
            res = self.disaggregate(out_vars, 0, self.n_days, out_dtype, units)
        if self.order is not None and not engine_order:
            res = {k: self._user_order(v) for k, v in res.items()}
            daily = {k: self._user_order(v) for k, v in daily.items()}
        self.check()
        self.stream.synchronize()
        res["daily"] = daily
        return res


class HostContext:
    """Caller-owned ``msb_host_ctx`` (include/metsim_b200.h): the device arena, pinned staging,
    streams and events of the host-buffer path on one device.  Contexts are independent -- two
    threads with a context each run concurrently; calls on one context are serialised by it."""

    def __init__(self, device: int = 0):
        self.lib = _lib.lib()
        self.device = int(device)
        h = C.c_void_p()
        _lib.check(self.lib.msb_host_ctx_create(self.device, C.byref(h)))
        self.handle = h

    @property
    def device_bytes(self) -> int:
        return int(self.lib.msb_host_ctx_device_bytes(self.handle)) if self.handle else 0

    def close(self):
        if getattr(self, "handle", None):
            self.lib.msb_host_ctx_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_host(params: dict | None, *, lat, lon, elev, doy, month, t_min, t_max, prec, wind=None,
             state_t_min, state_t_max, state_prec, t_pk12=None, dur12=None, t_end=None,
             given_shortwave=None, given_vapor_pressure=None, given_tskc=None, given_longwave=None,
             out_vars=DEFAULT_OUT_VARS, out_dtype=np.float32, units: dict | None = None,
             daily_vars=(), device: int = 0, tile_days: int = 0, out: dict | None = None,
             ring: bool = False, timings: dict | None = None, on_tile=None,
             ctx: HostContext | None = None, tuning: dict | None = None) -> dict:
    """Whole path with HOST (numpy) buffers through ``msb_run_host``: the call a reference-side
    binding makes in place of the cell loop of ``MetSim.run_slice`` (metsim/metsim.py:477-493).
    Host-to-device and device-to-host copies happen inside.

    ``on_tile(tile, day0, day1, slot)`` is the consumer hook of a streaming writer: called once per
    output tile, in order, when the tile's rows are in host memory and while the next tile is
    already computing (``slot`` is the ring slot with ``ring=True``, else -1).  An exception
    raised inside it aborts the run and is re-raised here.

    ``ctx``: a :class:`HostContext` to run in (its device wins over ``device``; the arena is reused
    from call to call); without one a context is created for this call and destroyed afterwards."""
    lib = _lib.lib()
    p = _lib.make_params(params)
    ts = p.time_step
    f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    t_min = f64(t_min)
    d, n = t_min.shape
    tpd = 1440 // ts if ts < 1440 else 1
    arrs = dict(lat=f64(lat), lon=f64(lon), elev=f64(elev), t_min=t_min, t_max=f64(t_max),
                prec=f64(prec), wind=f64(wind), st_t_min=f64(state_t_min),
                st_t_max=f64(state_t_max), st_prec=f64(state_prec), t_pk12=f64(t_pk12),
                dur12=f64(dur12), t_end=f64(t_end), given_shortwave=f64(given_shortwave),
                given_vapor_pressure=f64(given_vapor_pressure), given_tskc=f64(given_tskc),
                given_longwave=f64(given_longwave))
    arrs["doy"] = np.ascontiguousarray(doy, dtype=np.int32)
    arrs["month"] = np.ascontiguousarray(month, dtype=np.int32)
    # the C ABI takes bare pointers: every shape is checked here, before any of them is taken
    want_shape = {"lat": (n,), "lon": (n,), "elev": (n,), "doy": (d,), "month": (d,),
                  "st_t_min": (90, n), "st_t_max": (90, n), "st_prec": (90, n),
                  "t_pk12": (12, n), "dur12": (12, n), "t_end": (2, n)}
    for k, a in arrs.items():
        if a is not None and a.shape != want_shape.get(k, (d, n)):
            raise ValueError(f"run_host: {k} has shape {a.shape}, expected {want_shape.get(k, (d, n))}")
    for k in ("lat", "lon", "elev", "t_max", "prec", "st_t_min", "st_t_max", "st_prec"):
        if arrs[k] is None:
            raise ValueError(f"run_host: {k} is required")
    run = _lib.HostRun()
    run.n_cells, run.n_days, run.device = n, d, (ctx.device if ctx is not None else device)
    run.tune = _lib.make_tuning(tuning)
    for k, a in arrs.items():
        setattr(run, k, None if a is None else a.ctypes.data_as(C.c_void_p))
    res = {}
    odt = np.dtype(out_dtype)
    run.out_dtype = odt.itemsize
    run.tile_days = int(tile_days)
    run.out_is_ring = int(bool(ring))
    for i, v in enumerate(SUB_VARS):
        run.scale[i], run.offset[i] = 1.0, 0.0
        if ts >= 1440 or v not in out_vars or (v == "wind" and wind is None):
            run.out_host[i] = None
            continue
        if out is not None and v in out:
            o = out[v]
            rows = 2 * int(tile_days) * tpd if ring else d * tpd
            if (not isinstance(o, np.ndarray) or o.dtype != odt or o.ndim != 2 or o.shape[1] != n
                    or o.shape[0] < rows or not o.flags.c_contiguous or not o.flags.writeable):
                raise ValueError(f"run_host: out[{v!r}] must be a writeable C-contiguous {odt} array of "
                                 f"shape (>= {rows}, {n})")
            if ring and int(tile_days) <= 0:
                raise ValueError("run_host: ring=True needs an explicit tile_days")
            res[v] = o
        else:
            if ring:
                raise ValueError(f"run_host: ring=True needs a preallocated out[{v!r}]")
            res[v] = np.empty((d * tpd, n), dtype=odt)
        run.out_host[i] = res[v].ctypes.data_as(C.c_void_p)
        run.scale[i], run.offset[i] = unit_scale_offset(v, (units or {}).get(v), ts)
    daily = {}
    for i, v in enumerate(DAILY_VARS):
        if v in daily_vars:
            daily[v] = np.empty((d, n), dtype=np.float64)
            run.daily_host[i] = daily[v].ctypes.data_as(C.c_void_p)
        else:
            run.daily_host[i] = None
    n_iter = np.zeros(n, dtype=np.int32)
    run.n_iter_host = n_iter.ctypes.data_as(C.c_void_p)
    failure = []
    if on_tile is not None:
        def _cb(_user, tile, day0, day1, slot):
            try:
                on_tile(int(tile), int(day0), int(day1), int(slot))
                return 0
            except BaseException as exc:       # must not propagate through the C frame
                failure.append(exc)
                return 1
        cb = _lib.ON_TILE(_cb)                  # keep the thunk alive for the duration of the call
        run.on_tile = C.cast(cb, C.c_void_p)
    if ctx is not None:
        rc = lib.msb_run_host_ctx(ctx.handle, C.byref(p), C.byref(run))
    else:
        rc = lib.msb_run_host(C.byref(p), C.byref(run))
    if failure:
        raise failure[0]
    _lib.check(rc)
    daily["n_iter"] = n_iter
    res["daily"] = daily
    if timings is not None:
        timings.update(ms_h2d=run.ms_h2d, ms_solar=run.ms_solar, ms_daily=run.ms_daily,
                       ms_disagg_total=run.ms_disagg_total)
    return res
