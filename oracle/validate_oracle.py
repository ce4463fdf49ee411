"""TEST INFRASTRUCTURE ONLY -- pins the C oracle against the real reference.

Runs the unmodified reference hot-path modules (oracle/reference_harness.py, needs the
reference checkout) and the C restatement (oracle/oracle_lib.py) on the same seeded cells over
a sweep of configurations and prints the worst relative error per variable.  Exit status is
non-zero if any variable exceeds ``TOL``.

    python -m oracle.validate_oracle [--quick]
"""
from __future__ import annotations

import sys

import numpy as np
import pandas as pd

from oracle import oracle_lib, reference_harness as rh

TOL = 1e-9
# relative error floor per variable: err = |a-b| / max(|b|, floor)
FLOORS = {"temp": 1.0, "prec": 1e-3, "shortwave": 1e-2, "longwave": 1.0, "vapor_pressure": 1e-3,
          "rel_humid": 1e-1, "air_pressure": 1.0, "spec_humid": 1e-6, "tskc": 1e-3, "wind": 1e-2,
          "t_day": 1.0, "tfmax": 1e-3, "tskc_daily": 1e-3, "tdew": 1.0, "vp_daily": 1.0,
          "sw_daily": 1e-1, "pet": 1e-2, "daylength": 1.0, "potrad": 1e-1, "tt_max": 1e-3,
          "seasonal_prec": 1e-1, "smoothed_dtr": 1e-2}
REF_DAILY = {"t_day": "t_day", "tfmax": "tfmax", "tskc_daily": "tskc", "tdew": "tdew",
             "vp_daily": "vapor_pressure", "sw_daily": "shortwave", "pet": "pet",
             "daylength": "daylength", "potrad": "potrad", "tt_max": "tt_max",
             "seasonal_prec": "seasonal_prec", "smoothed_dtr": "smoothed_dtr"}


def relerr(a, b, floor):
    a, b = np.asarray(a, float), np.asarray(b, float)
    both_nan = np.isnan(a) & np.isnan(b)
    den = np.maximum(np.abs(b), floor)
    e = np.abs(a - b) / den
    e[both_nan] = 0.0
    e[np.isnan(e)] = np.inf
    return float(e.max()) if e.size else 0.0


def synth_cell(rng, lat, lon, elev, start, d):
    dates = pd.date_range(start, periods=d, freq="D")
    doy = dates.dayofyear.values
    base = 12 - 0.45 * abs(lat) - 8 * np.cos(2 * np.pi * (doy - 15) / 365.25) * np.sign(lat or 1)
    t_min = base + rng.normal(0, 3, d)
    t_max = t_min + rng.uniform(2, 18, d)
    prec = (rng.random(d) < 0.3) * rng.gamma(0.8, 6, d)
    st_tmin = base[0] + rng.normal(0, 3, 90)
    return dict(lat=lat, lon=lon, elev=elev, start=start, t_min=t_min, t_max=t_max, prec=prec,
                wind=rng.uniform(0.5, 8, d), state_t_min=st_tmin,
                state_t_max=st_tmin + rng.uniform(2, 18, 90),
                state_prec=(rng.random(90) < 0.3) * rng.gamma(0.8, 6, 90),
                t_pk12=rng.uniform(300, 1140, 12), dur12=rng.uniform(120, 720, 12))


def oracle_cell(cell, params):
    d = len(cell["t_min"])
    dates = pd.date_range(cell["start"], periods=d, freq="D")
    col = lambda k, n: np.asarray(cell[k], float).reshape(n, 1)
    kw = dict(lat=[cell["lat"]], lon=[cell["lon"]], elev=[cell["elev"]],
              doy=dates.dayofyear.values, month=dates.month.values,
              t_min=col("t_min", d), t_max=col("t_max", d), prec=col("prec", d),
              wind=col("wind", d) if cell.get("wind") is not None else None,
              state_t_min=col("state_t_min", 90), state_t_max=col("state_t_max", 90),
              state_prec=col("state_prec", 90))
    if "t_pk12" in cell:
        kw["t_pk12"], kw["dur12"] = col("t_pk12", 12), col("dur12", 12)
    if cell.get("t_end") is not None:
        kw["t_end"] = np.asarray(cell["t_end"], float).reshape(2, 1)
    for k in ("shortwave", "vapor_pressure", "tskc", "longwave"):   # passthrough columns
        if cell.get(k) is not None:
            kw["given_" + k] = col(k, d)
    return oracle_lib.run(params, n_threads=1, **kw)


def compare(cell, params, worst, label):
    ref = rh.run_cell(cell, params)
    got = oracle_cell(cell, params)
    bad = []
    for ov, rv in REF_DAILY.items():
        e = relerr(got[ov][:, 0], ref["daily"][rv], FLOORS[ov])
        worst[ov] = max(worst.get(ov, 0), e)
        if e > TOL:
            bad.append((ov, e))
    if "sub" in ref:
        for v, arr in ref["sub"].items():
            e = relerr(got[v][:, 0], arr, FLOORS[v])
            worst[v] = max(worst.get(v, 0), e)
            if e > TOL:
                bad.append((v, e))
    if bad:
        print("  MISMATCH", label, bad)
    return not bad


def main(argv):
    quick = "--quick" in argv
    rng = np.random.default_rng(20261019)
    worst, ok, n = {}, True, 0
    cases = []
    lats = [48.3125, 30.03125, -33.9, 0.25, 64.9, 70.3, -54.75, 83.75]
    lons = [-120.5625, -100.03125, 151.2, 0.0, 179.9, -179.75, 25.5, -45.25]
    elevs = [1668.15, 519.0, 30.0, 2200.0, 4100.0, 5.0, 0.0, 800.0]
    for i, (la, lo, el) in enumerate(zip(lats, lons, elevs)):
        for ts, pt, utc in ((60, "triangle", True), (180, "mix", True), (30, "uniform", False),
                            (15, "triangle", True), (360, "mix", False), (1440, "uniform", False)):
            start, d = (("2001-01-01", 365) if i % 3 == 0 else
                        ("2000-02-10", 400) if i % 3 == 1 else ("2003-11-20", 61))
            cases.append((la, lo, el, start, d, dict(time_step=ts, prec_type=pt, utc_offset=utc)))
    # longwave variants and non-default scalars
    for lw in ("default", "tva", "anderson", "brutsaert", "satterlund", "idso", "prata"):
        for cloud in ("default", "cloud_deardorff"):
            cases.append((41.2, -105.1, 1900.0, "2004-02-01", 45,
                          dict(time_step=120, prec_type="uniform", utc_offset=True, lw_type=lw,
                               lw_cloud=cloud, sw_prec_thresh=0.5, rain_scalar=0.7,
                               tday_coef=0.5, tmax_daylength_fraction=0.6, tdew_tol=1e-8)))
    # 'passthrough' method (methods/passthrough.py): supplied shortwave, optionally vapour
    # pressure / tskc / longwave, both sw_averaging modes
    for i, (la, lo, el) in enumerate(zip(lats[:6], lons[:6], elevs[:6])):
        cases.append((la, lo, el, "2002-06-10", 40,
                      dict(method="passthrough", time_step=(60, 180, 30)[i % 3],
                           prec_type=("uniform", "triangle")[i % 2], utc_offset=bool(i % 2),
                           sw_averaging=("daylight" if i % 3 == 0 else ""), _given=i)))
    if quick:
        cases = cases[::5] + cases[-2:]
    for la, lo, el, start, d, prm in cases:
        cell = synth_cell(rng, la, lo, el, start, d)
        if "_given" in prm:
            prm = dict(prm)
            g = prm.pop("_given")
            cell["shortwave"] = rng.uniform(20, 350, d)
            if g & 1:
                cell["vapor_pressure"] = rng.uniform(200, 1800, d)
            if g & 2:
                cell["tskc"] = rng.uniform(0, 1, d)
            if g & 4 or g == 0:
                cell["longwave"] = rng.uniform(180, 420, d)
        if n % 4 == 1:
            cell["t_end"] = [cell["t_min"][-1] - 1.5, cell["t_max"][-1] + 2.0]
        ok &= compare(cell, prm, worst, f"lat={la} lon={lo} {prm}")
        n += 1
    print(f"{n} cell-runs compared; worst relative error per variable (tol {TOL:g}):")
    for k in sorted(worst):
        print(f"  {k:16s} {worst[k]:.3e}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
