"""Shared helpers of the test-suite: golden loaders, tolerances, comparisons."""
from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star: fp64 kernels, max relative error <= 1e-9 on every output variable.  The relative
# error needs a floor where a variable crosses zero / carries ~1e-18 noise (SURVEY.md hard part 6):
#   err = |a - b| / max(|b|, FLOOR[var])
TOL = 1e-9
FLOORS = {"temp": 1.0, "prec": 1e-3, "shortwave": 1e-2, "longwave": 1.0, "vapor_pressure": 1e-3,
          "rel_humid": 1e-1, "air_pressure": 1.0, "spec_humid": 1e-6, "tskc": 1e-3, "wind": 1e-2,
          "t_day": 1.0, "tfmax": 1e-3, "tskc_daily": 1e-3, "tdew": 1.0, "vp_daily": 1.0,
          "sw_daily": 1e-1, "pet": 1e-2, "daylength": 1.0, "potrad": 1e-1, "tt_max": 1e-3,
          "seasonal_prec": 1e-1, "smoothed_dtr": 1e-2}
SUB = ("temp", "prec", "shortwave", "longwave", "vapor_pressure", "rel_humid", "air_pressure",
       "spec_humid", "tskc", "wind")
# oracle name -> (engine daily name, golden ref_daily name)
DAILY = {"t_day": "t_day", "tfmax": "tfmax", "tskc_daily": "tskc", "tdew": "tdew",
         "vp_daily": "vapor_pressure", "sw_daily": "shortwave", "pet": "pet",
         "daylength": "daylength", "potrad": "potrad", "tt_max": "tt_max",
         "seasonal_prec": "seasonal_prec", "smoothed_dtr": "smoothed_dtr"}
INPUT_KEYS = ("lat", "lon", "elev", "doy", "month", "t_min", "t_max", "prec", "wind",
              "state_t_min", "state_t_max", "state_prec", "t_pk12", "dur12", "t_end",
              "given_shortwave", "given_vapor_pressure", "given_tskc", "given_longwave")


def relerr(a, b, floor):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    e = np.abs(a - b) / np.maximum(np.abs(b), floor)
    e[both_nan] = 0.0
    e[np.isnan(e)] = np.inf
    return float(e.max()) if e.size else 0.0


def golden_files():
    """Fixtures holding reference OUTPUTS (``inputs_*.npz`` are input-only sets for the driver tests)."""
    return sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not os.path.basename(p).startswith("inputs_"))


def load_golden(path):
    z = np.load(path, allow_pickle=False)
    inputs = {k: z[k] for k in INPUT_KEYS if k in z.files}
    params = {}
    for k in z.files:
        if k.startswith("param_"):
            v = z[k]
            params[k[6:]] = v.item() if v.shape == () else v
    for k, v in list(params.items()):
        if isinstance(v, (np.str_, bytes)):
            params[k] = str(v)
    ref_daily = {k[10:]: z[k] for k in z.files if k.startswith("ref_daily_")}
    ref_sub = {k[8:]: z[k] for k in z.files if k.startswith("ref_sub_")}
    extra = {k: z[k] for k in ("ref_cells", "repo_golden", "repo_golden_columns") if k in z.files}
    return inputs, params, ref_daily, ref_sub, extra


def oracle_kwargs(inputs):
    return {k: inputs.get(k) for k in INPUT_KEYS}


def assert_close(got, want, name, floor_key=None, tol=TOL, context=""):
    e = relerr(got, want, FLOORS[floor_key or name])
    assert e <= tol, f"{context}: {name} relative error {e:.3e} > {tol:g}"
    return e
