"""TEST INFRASTRUCTURE ONLY -- runs the *unmodified* MetSim hot-path modules from a
reference checkout (default /root/reference) so that the C restatement in
``oracle/metsim_oracle.c`` can be pinned against the real thing and golden vectors
can be generated (``oracle/make_golden.py``).

Nothing in the product (``metsim_b200/``), in the ``-m gpu`` tests, in ``smoke()`` or in
``bench.py`` imports this module: the reference tree does not exist on the GPU box.

How the reference is loaded
---------------------------
``import metsim`` fails in this image (xarray/dask/netCDF4 are absent,
metsim/__init__.py:5 -> metsim/metsim.py:44-52).  The hot-path modules themselves only
need numpy/pandas/scipy/numba, so they are imported through a stub parent package
(bypassing ``metsim/__init__.py``) plus stub ``xarray``/``netCDF4`` modules that satisfy the
imports at metsim/datetime.py:5-6.  No reference source is copied: the modules execute from
where they lie.

What is restated here (driver glue only, pandas-3 safe)
-------------------------------------------------------
* ``wrap_run_cell``            metsim/metsim.py:693-796  (solar_geom -> method.run -> disaggregate)
* ``disaggregate()`` wrapper   metsim/disaggregate.py:66-130 (pandas 3 rejects the '{}T' alias
  and ``fillna(method=)``; every kernel function it calls is the reference's own)
* ``_aggregate_state``         metsim/metsim.py:589-618  (xarray rolling means -> plain windowed means)
* ``convert_monthly_param``    metsim/metsim.py:278-284  (including its month off-by-one)
"""
from __future__ import annotations

import importlib
import math
import os
import sys
import types

import numpy as np
import pandas as pd

_REF = None


def find_reference_root():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # baseline/_ref = offline `pip install --target` of the unmodified reference (git-ignored; it
    # travels to the GPU box, /root/reference does not)
    for cand in (os.environ.get("METSIM_REF"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "metsim", "methods")):
            return cand
    return None


def load_reference(root: str | None = None):
    """Return a namespace with the reference's constants/physics/mtclim/disaggregate modules."""
    global _REF
    if _REF is not None:
        return _REF
    root = root or find_reference_root()
    if root is None:
        raise RuntimeError("reference checkout not found (set METSIM_REF)")
    pkg = types.ModuleType("metsim")
    pkg.__path__ = [os.path.join(root, "metsim")]
    sys.modules["metsim"] = pkg
    for name in ("xarray", "netCDF4"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    nc4 = sys.modules["netCDF4"]
    for attr in ("date2num", "num2date"):
        if not hasattr(nc4, attr):
            setattr(nc4, attr, None)
    ns = types.SimpleNamespace()
    ns.cnst = importlib.import_module("metsim.constants")
    ns.physics = importlib.import_module("metsim.physics")
    ns.mtclim = importlib.import_module("metsim.methods.mtclim")
    ns.dg = importlib.import_module("metsim.disaggregate")

    # mtclim.tdew writes into ``seasonal_prec`` in place (mtclim.py:187-188); under pandas 3
    # ``df[...].values`` is read-only, so hand it a private copy.  Numerically identical.
    orig_tdew = ns.mtclim.tdew

    def tdew_copying(pet, t_min, seasonal_prec, dtr):
        return orig_tdew(pet, t_min, np.array(seasonal_prec, dtype=float, copy=True), dtr)

    if not getattr(ns.mtclim.tdew, "_patched", False):
        tdew_copying._patched = True
        ns.mtclim.tdew = tdew_copying
    # methods/passthrough.py binds mtclim's helpers by name at import time: same read-only fix
    ns.passthrough = importlib.import_module("metsim.methods.passthrough")
    ns.passthrough.tdew = tdew_copying
    ns.root = root
    _REF = ns
    return ns


DEFAULT_PARAMS = {
    # metsim/metsim.py:140-171
    "method": "mtclim",
    "time_step": 60,
    "calendar": "standard",
    "prec_type": "uniform",
    "sw_prec_thresh": 0.0,
    "utc_offset": False,
    "lw_cloud": "cloud_deardorff",
    "lw_type": "prata",
    "tdew_tol": 1e-6,
    "tmax_daylength_fraction": 0.67,
    "rain_scalar": 0.75,
    "tday_coef": 0.45,
    "lapse_rate": 0.0065,
    "period_ending": False,
}


def rolling_means(state_tmin, state_tmax, state_prec, t_min, t_max, prec):
    """metsim/metsim.py:589-618 with plain windowed means (xarray is not importable here)."""
    assert len(state_prec) == 90
    dtr = t_max - t_min
    if (dtr < 0).any():
        raise ValueError("Daily maximum temperature lower than daily minimum temperature!")
    tot = np.concatenate([state_prec, prec])
    d = len(prec)
    seasonal = np.array([365.25 * np.mean(tot[i + 1:i + 91]) for i in range(d)])
    full_dtr = np.concatenate([state_tmax - state_tmin, dtr])
    sm = np.array([np.mean(full_dtr[90 + i - 29:90 + i + 1]) for i in range(d)])
    return seasonal, dtr, sm


def monthly_to_daily(monthly12, months, prec):
    """metsim/metsim.py:278-284: entry ``m`` goes to calendar month ``m`` (1..12 vs 0..11), so
    December keeps the copied precipitation values."""
    out = np.array(prec, dtype=float, copy=True)
    for m in range(12):
        out[months == m] = monthly12[m]
    return out


def run_cell(cell: dict, user_params: dict | None = None, root: str | None = None) -> dict:
    """Run one cell through the reference.  ``cell`` keys: lat, lon, elev, start (date str),
    t_min, t_max, prec, [wind], state_t_min, state_t_max, state_prec (90 each),
    [t_pk12, dur12] (12 each), [t_end] (pair)."""
    ref = load_reference(root)
    cnst, dg = ref.cnst, ref.dg
    params = dict(DEFAULT_PARAMS)
    params.update(user_params or {})
    for k in ("sw_prec_thresh", "rain_scalar", "tdew_tol", "tmax_daylength_fraction",
              "tday_coef", "lapse_rate"):
        params[k] = float(params[k])  # metsim.py:467-473
    params["time_step"] = int(params["time_step"])

    t_min = np.asarray(cell["t_min"], dtype=np.float64)
    t_max = np.asarray(cell["t_max"], dtype=np.float64)
    prec = np.asarray(cell["prec"], dtype=np.float64)
    d = len(t_min)
    dates = pd.date_range(cell["start"], periods=d, freq="D")
    df = pd.DataFrame(index=dates)
    df["t_min"], df["t_max"], df["prec"] = t_min, t_max, prec
    if "wind" in cell and cell["wind"] is not None:
        df["wind"] = np.asarray(cell["wind"], dtype=np.float64)
    seasonal, dtr, sm = rolling_means(
        np.asarray(cell["state_t_min"], float), np.asarray(cell["state_t_max"], float),
        np.asarray(cell["state_prec"], float), t_min, t_max, prec)
    df["seasonal_prec"], df["dtr"], df["smoothed_dtr"] = seasonal, dtr, sm
    if params["prec_type"].upper() in ("TRIANGLE", "MIX"):
        months = dates.month.values
        df["dur"] = monthly_to_daily(np.asarray(cell["dur12"], float), months, prec)
        df["t_pk"] = monthly_to_daily(np.asarray(cell["t_pk12"], float), months, prec)

    # wrap_run_cell, metsim.py:723-744
    params["elev"] = float(cell["elev"])
    params["lat"] = float(cell["lat"])
    params["lon"] = float(cell["lon"])
    sg_t = ref.physics.solar_geom(params["elev"], params["lat"], params["lapse_rate"])
    sg = {"tiny_rad_fract": sg_t[0], "daylength": sg_t[1], "potrad": sg_t[2], "tt_max0": sg_t[3]}
    yday = dates.dayofyear.values - 1
    df["daylength"] = sg["daylength"][yday]
    df["potrad"] = sg["potrad"][yday]
    df["tt_max"] = sg["tt_max0"][yday]
    if params["method"] == "passthrough":          # metsim.py:462: methods[params['method']].run
        for k in ("shortwave", "vapor_pressure", "tskc", "longwave"):
            if cell.get(k) is not None:
                df[k] = np.asarray(cell[k], dtype=np.float64)
        df = ref.passthrough.run(df, params)
    else:
        df = ref.mtclim.run(df, params)

    out = {"daily": {k: df[k].values.copy() for k in
                     ("t_day", "tfmax", "tskc", "tdew", "vapor_pressure", "shortwave", "pet",
                      "daylength", "potrad", "tt_max", "seasonal_prec", "smoothed_dtr", "dtr")}}
    ts = params["time_step"]
    if ts >= 1440:
        out["daily"]["shortwave_daily_mean"] = (df["shortwave"] * df["daylength"]
                                                / cnst.SEC_PER_DAY).values
        return out

    # disaggregate() wrapper, disaggregate.py:66-130
    t_begin = [float(cell["state_t_min"][-1]), float(cell["state_t_max"][-1])]  # metsim.py:753-754
    t_end = cell.get("t_end")
    params["lon"] = math.remainder(params["lon"], 360)
    n_days = d
    n_disagg = n_days * (1440 // ts)
    sub = {}
    if params["method"] == "passthrough" and params.get("sw_averaging", "") != "daylight":
        df["daylength"] = 86400.0                  # disaggregate.py:78-80
    sub["shortwave"] = dg.shortwave(df["shortwave"].values, df["daylength"].values,
                                    dates.dayofyear, sg["tiny_rad_fract"], params)
    t_tmin, t_tmax = dg.set_min_max_hour(sg["tiny_rad_fract"], dates.dayofyear - 1,
                                         sg["daylength"], n_days, ts, params)
    sub["temp"] = dg.temp(df["t_min"].values, df["t_max"].values, n_disagg, t_tmin, t_tmax, ts,
                          t_begin, t_end)
    vp = dg.vapor_pressure(df["vapor_pressure"].values, sub["temp"], t_tmin, n_disagg, ts)
    sub["vapor_pressure"] = pd.Series(vp).ffill().bfill().values
    sub["rel_humid"] = dg.relative_humidity(sub["vapor_pressure"], sub["temp"])
    sub["air_pressure"] = dg.pressure(sub["temp"], params["elev"], params["lapse_rate"])
    sub["spec_humid"] = dg.specific_humidity(sub["vapor_pressure"], sub["air_pressure"])
    sub["tskc"] = dg.tskc(df["tskc"].values, ts, params)
    daily_lw = df["longwave"] if "longwave" in df else None     # disaggregate.py:115-118
    sub["longwave"] = dg.longwave(sub["temp"], sub["vapor_pressure"], sub["tskc"], params, daily_lw)
    sub["prec"] = dg.prec(df["prec"], df["t_min"], ts, params, df.get("t_pk"), df.get("dur"))
    if "wind" in df:
        sub["wind"] = dg.wind(df["wind"].values, ts, params)
    frame = pd.DataFrame(sub).ffill().bfill()
    out["sub"] = {k: frame[k].values.copy() for k in frame.columns}
    out["aux"] = {"t_tmin": np.asarray(t_tmin, float), "t_tmax": np.asarray(t_tmax, float)}
    return out
