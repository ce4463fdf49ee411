"""The C-ABI shared library loads on a CPU-only box and exports every symbol that
include/metsim_b200.h declares.  No compute calls here (no GPU)."""
import ctypes as C
import os
import re

from metsim_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "metsim_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msb_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = header_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names


def test_version_defaults_and_error_text():
    lib = _lib.lib()
    assert lib.msb_version() == 201
    p = _lib.make_params({})
    # metsim/metsim.py:140-171
    assert (p.time_step, p.prec_type, p.utc_offset) == (60, 0, 0)
    assert (p.lw_type, p.lw_cloud) == (_lib.LW_TYPES["prata"], _lib.LW_CLOUD["cloud_deardorff"])
    assert (p.tdew_tol, p.rain_scalar, p.tday_coef, p.lapse_rate) == (1e-6, 0.75, 0.45, 0.0065)
    assert p.tmax_daylength_fraction == 0.67 and p.sw_prec_thresh == 0.0
    # argument validation happens before any CUDA call, so it is testable without a GPU
    rc = lib.msb_solar_table_sizes(0, None, None, None)
    assert rc == _lib.E_ARG and b"n_lat" in lib.msb_last_error()
    small, q, tt = C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert lib.msb_solar_table_sizes(3, C.byref(small), C.byref(q), C.byref(tt)) == 0
    assert (small.value, q.value, tt.value) == (3 * 366, 3 * 365 * _lib.Q_LD + _lib.Q_PAD, 3 * 365 * _lib.TT_TERMS)
    # window-major copy of the shortwave table: C phases x (tpd + 1 windows padded to a multiple of 4) + the row total
    assert lib.msb_solar_q_win_ld(60) == 120 * 28 + 4 and lib.msb_solar_q_win_ld(15) == 30 * 100 + 4
    assert lib.msb_solar_q_win_ld(1440) == 0 and lib.msb_solar_q_win_ld(7) == 0 and lib.msb_solar_q_win_ld(0) == 0
    assert lib.msb_solar_q_win_elems(3, 180) == 3 * 365 * (360 * 12 + 4) + _lib.Q_PAD
    assert lib.msb_daily_scratch_bytes(1000, 365) % 8 == 0 and lib.msb_daily_scratch_bytes(1000, 365) > 0
    assert lib.msb_mtclim_daily(None, None, None, None, None) == _lib.E_ARG
    assert lib.msb_disaggregate(None, None, None, None, None, None) == _lib.E_ARG


def test_struct_layouts_match_header():
    # sizes implied by the header on LP64 (natural alignment)
    assert C.sizeof(_lib.Params) == 8 * 4 + 6 * 8
    assert C.sizeof(_lib.SolarTables) == 8 + 7 * 8 + 8
    assert C.sizeof(_lib.Forcing) == 16 + 24 * 8
    assert C.sizeof(_lib.Tuning) == 8 * 4
    assert C.sizeof(_lib.Daily) == (13 + 3 + 3) * 8 + 32
    assert C.sizeof(_lib.SubDaily) == 10 * 8 * 3 + 8 + 8 + 8 + 16 + 32


def test_param_translation_errors_match_reference():
    import pytest
    with pytest.raises(ValueError):
        _lib.make_params({"lw_type": "bogus"})      # metsim.py:673-680
    with pytest.raises(ValueError):
        _lib.make_params({"lw_cloud": "bogus"})
    p = _lib.make_params({"prec_type": "MIX", "lw_type": "TVA", "time_step": "180", "utc_offset": True})
    assert (p.prec_type, p.lw_type, p.time_step, p.utc_offset) == (2, 1, 180, 1)
    assert (p.method, p.sw_daylight) == (0, 0)
    p = _lib.make_params({"method": "passthrough", "sw_averaging": "daylight"})   # disaggregate.py:79
    assert (p.method, p.sw_daylight) == (1, 1)
    with pytest.raises(KeyError):
        _lib.make_params({"method": "bogus"})       # metsim.py:462: self.methods[params['method']]


def test_ctypes_structs_match_the_compiled_header(tmp_path):
    """sizeof / offsetof of every struct of include/metsim_b200.h as gcc lays them out, against the
    ctypes mirrors in metsim_b200/_lib.py (catches any drift between the header and the binding)."""
    import subprocess
    pairs = {"msb_tuning": _lib.Tuning, "msb_params": _lib.Params, "msb_solar_tables": _lib.SolarTables, "msb_forcing": _lib.Forcing,
             "msb_daily": _lib.Daily, "msb_subdaily": _lib.SubDaily, "msb_host_run": _lib.HostRun,
             "msb_sink_desc": _lib.SinkDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "metsim_b200.h"', 'int main(void) {']
    for cname, ct in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = os.path.join(tmp_path, "layout.c")
    exe = os.path.join(tmp_path, "layout")
    open(src, "w").write("\n".join(lines))
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
    got = dict(l.split() for l in subprocess.check_output([exe], text=True).splitlines())
    for cname, ct in pairs.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


def test_host_context_argument_checks_need_no_gpu():
    """The context API validates before touching CUDA, and the workspace-size query is pure host
    arithmetic (workspace-size query instead of hidden allocation)."""
    import numpy as np
    lib = _lib.lib()
    assert lib.msb_run_host_ctx(None, None, None) == _lib.E_ARG
    assert b"NULL context" in lib.msb_last_error()
    assert lib.msb_host_ctx_device_bytes(None) == 0
    lib.msb_host_ctx_destroy(None)                       # a no-op, like free(NULL)
    n, d = 64, 10
    z = lambda *s: np.zeros(s)
    arrs = dict(lat=np.linspace(30, 31, n), lon=z(n), elev=z(n), t_min=z(d, n), t_max=z(d, n) + 5,
                prec=z(d, n), st_t_min=z(90, n), st_t_max=z(90, n) + 5, st_prec=z(90, n))
    doy = np.arange(1, d + 1, dtype=np.int32)
    month = np.ones(d, dtype=np.int32)
    run = _lib.HostRun()
    run.n_cells, run.n_days, run.device = n, d, 0
    for k, a in arrs.items():
        setattr(run, k, a.ctypes.data_as(C.c_void_p))
    run.doy, run.month = doy.ctypes.data_as(C.c_void_p), month.ctypes.data_as(C.c_void_p)
    out = np.empty((d * 24, n), dtype=np.float32)
    run.out_host[0] = out.ctypes.data_as(C.c_void_p)
    run.out_dtype = 4
    p = _lib.make_params({})
    dev, pin = C.c_size_t(), C.c_size_t()
    assert lib.msb_run_host_bytes(C.byref(p), C.byref(run), C.byref(dev), C.byref(pin)) == 0
    # at least: 3 daily inputs, the 3 state arrays, the q table of the 64 latitudes, two output tiles
    floor = 3 * d * n * 8 + 3 * 90 * n * 8 + n * 365 * _lib.Q_LD * 8 + 2 * d * 24 * n * 4
    assert dev.value >= floor and pin.value > 0
    run.out_dtype = 2
    assert lib.msb_run_host_bytes(C.byref(p), C.byref(run), C.byref(dev), C.byref(pin)) == _lib.E_ARG
    run.out_dtype, run.out_is_ring = 4, 1                # a ring without tile_days cannot be sized
    assert lib.msb_run_host_bytes(C.byref(p), C.byref(run), C.byref(dev), C.byref(pin)) == _lib.E_ARG


def test_tuning_translation_and_environment(monkeypatch):
    for k in ("MSB_DISAGG_PATH", "MSB_KNOT_TILE", "MSB_DAY_TILE", "MSB_DAILY_CHUNK"):
        monkeypatch.delenv(k, raising=False)
    t = _lib.make_tuning({})
    assert all(getattr(t, k) == 0 for k in _lib.TUNING_KEYS)
    t = _lib.make_tuning({"knot_tile": 16, "day_tile": 16, "daily_chunk_len": 73, "disagg_path": "event"})
    assert (t.knot_tile, t.day_tile, t.daily_chunk_len, t.disagg_path) == (16, 16, 73, 1)
    monkeypatch.setenv("MSB_DISAGG_PATH", "event")
    monkeypatch.setenv("MSB_KNOT_TILE", "8")
    t = _lib.make_tuning({"day_tile": 4})
    assert (t.knot_tile, t.day_tile, t.disagg_path) == (8, 4, 1)
    import pytest
    with pytest.raises(KeyError):
        _lib.make_tuning({"bogus": 1})
