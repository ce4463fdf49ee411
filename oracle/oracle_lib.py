"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the C oracle (oracle/metsim_oracle.c).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this.  The product (metsim_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmetsim_oracle.so")
_lib = None

PREC_TYPES = {"uniform": 0, "triangle": 1, "mix": 2}
LW_TYPES = {"default": 0, "tva": 1, "anderson": 2, "brutsaert": 3, "satterlund": 4, "idso": 5,
            "prata": 6}
LW_CLOUD = {"default": 0, "tva": 0, "cloud_deardorff": 1}

DAILY_VARS = ("t_day", "tfmax", "tskc_daily", "tdew", "vp_daily", "sw_daily", "pet",
              "daylength", "potrad", "tt_max", "seasonal_prec", "smoothed_dtr")
SUB_VARS = ("temp", "prec", "shortwave", "longwave", "vapor_pressure", "rel_humid",
            "air_pressure", "spec_humid", "tskc", "wind")


class Params(C.Structure):
    _fields_ = [("time_step", C.c_int), ("prec_type", C.c_int), ("utc_offset", C.c_int),
                ("lw_type", C.c_int), ("lw_cloud", C.c_int), ("max_iter", C.c_int),
                ("sw_prec_thresh", C.c_double), ("rain_scalar", C.c_double),
                ("tdew_tol", C.c_double), ("tmax_daylength_fraction", C.c_double),
                ("tday_coef", C.c_double), ("lapse_rate", C.c_double)]


class Outputs(C.Structure):
    _fields_ = ([(n, C.POINTER(C.c_double)) for n in DAILY_VARS]
                + [(n, C.POINTER(C.c_double)) for n in SUB_VARS]
                + [("n_iter", C.POINTER(C.c_int))])


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "metsim_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.mso_run.restype = C.c_int
        _lib.mso_default_params.restype = None
        _lib.mso_solar_geom.restype = None
    return _lib


def make_params(user: dict | None = None) -> Params:
    p = Params()
    lib().mso_default_params(C.byref(p))
    user = dict(user or {})
    for k, v in user.items():
        if k == "prec_type":
            p.prec_type = PREC_TYPES[str(v).lower()]
        elif k == "lw_type":
            p.lw_type = LW_TYPES[str(v).lower()]
        elif k == "lw_cloud":
            p.lw_cloud = LW_CLOUD[str(v).lower()]
        elif k == "utc_offset":
            p.utc_offset = int(bool(v))
        elif k in ("time_step", "max_iter"):
            setattr(p, k, int(v))
        elif k in ("sw_prec_thresh", "rain_scalar", "tdew_tol", "tmax_daylength_fraction",
                   "tday_coef", "lapse_rate"):
            setattr(p, k, float(v))
    return p


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == shape, (a.shape, shape)
    return a


def solar_geom(elev: float, lat: float, lr: float = 0.0065):
    trf = np.empty((366, 2880))
    dayl, potrad, ttmax = np.empty(366), np.empty(366), np.empty(366)
    lib().mso_solar_geom(C.c_double(elev), C.c_double(lat), C.c_double(lr), _dp(trf), _dp(dayl),
                         _dp(potrad), _dp(ttmax))
    return trf, dayl, potrad, ttmax


class Passthrough(C.Structure):
    _fields_ = [("sw_daylight", C.c_int), ("shortwave", C.POINTER(C.c_double)),
                ("vapor_pressure", C.POINTER(C.c_double)), ("tskc", C.POINTER(C.c_double)),
                ("longwave", C.POINTER(C.c_double))]


def run(params: dict | None, *, lat, lon, elev, doy, month, t_min, t_max, prec, wind=None,
        state_t_min, state_t_max, state_prec, t_pk12=None, dur12=None, t_end=None,
        given_shortwave=None, given_vapor_pressure=None, given_tskc=None, given_longwave=None,
        want=None, n_threads: int | None = None) -> dict:
    """Run the oracle over N cells x D days; arrays are [D][N] (cell fastest).  Returns a dict of
    numpy arrays for every name in ``want`` (default: everything)."""
    p = make_params(params)
    t_min = _f64(t_min)
    d, n = t_min.shape
    ts = p.time_step
    tpd = 1 if ts >= 1440 else 1440 // ts
    lat, lon, elev = _f64(lat, (n,)), _f64(lon, (n,)), _f64(elev, (n,))
    doy = np.ascontiguousarray(doy, dtype=np.int32)
    month = np.ascontiguousarray(month, dtype=np.int32)
    assert doy.shape == (d,) and month.shape == (d,)
    t_max, prec, wind = _f64(t_max, (d, n)), _f64(prec, (d, n)), _f64(wind, (d, n))
    st = [_f64(a, (90, n)) for a in (state_t_min, state_t_max, state_prec)]
    t_pk12, dur12, t_end = _f64(t_pk12, (12, n)), _f64(dur12, (12, n)), _f64(t_end, (2, n))
    names = list(DAILY_VARS) + (list(SUB_VARS) if ts < 1440 else [])
    if want is not None:
        names = [v for v in names if v in want]
    if wind is None and "wind" in names:
        names.remove("wind")
    res, out = {}, Outputs()
    for v in names:
        res[v] = np.full(((d if v in DAILY_VARS else d * tpd), n), np.nan)
        setattr(out, v, _dp(res[v]))
    res["n_iter"] = np.zeros(n, dtype=np.int32)
    out.n_iter = res["n_iter"].ctypes.data_as(C.POINTER(C.c_int))
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    pt = None
    if str((params or {}).get("method", "mtclim")).lower() == "passthrough":
        if given_shortwave is None:
            raise AssertionError("'shortwave' in df")     # methods/passthrough.py:29
        gs, gv, gt, gl = (_f64(a, (d, n)) for a in (given_shortwave, given_vapor_pressure,
                                                      given_tskc, given_longwave))
        pt = Passthrough(int(str((params or {}).get("sw_averaging", "")).lower() == "daylight"),
                         _dp(gs), _dp(gv), _dp(gt), _dp(gl))
    lib().mso_run_pt.restype = C.c_int
    rc = lib().mso_run_pt(C.byref(p), C.byref(pt) if pt is not None else None, n, d, _dp(lat), _dp(lon), _dp(elev),
                       doy.ctypes.data_as(C.POINTER(C.c_int)),
                       month.ctypes.data_as(C.POINTER(C.c_int)),
                       _dp(t_min), _dp(t_max), _dp(prec), _dp(wind), _dp(st[0]), _dp(st[1]),
                       _dp(st[2]), _dp(t_pk12), _dp(dur12), _dp(t_end), C.byref(out),
                       int(n_threads))
    if rc == -2:
        raise ValueError("Daily maximum temperature lower than daily minimum temperature!")
    if rc != 0:
        raise RuntimeError(f"oracle mso_run failed with code {rc}")
    return res
