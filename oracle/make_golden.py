"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the REAL reference.

Needs the reference checkout (default /root/reference); run in the build container:

    python -m oracle.make_golden

Every fixture holds the inputs (float64) and the outputs of the unmodified reference functions
(through oracle/reference_harness.py).  The fixtures travel to the GPU box, the reference does
not.  Fixture families:

* ``stehekin_*``   the reference's only known-answer test (metsim/tests/test_metsim.py:397-464):
                   VIC binary forcings of cell 48.3125/-120.5625 for 1949, domain stehekin.nc,
                   plus the reference repo's own golden tables (validated_48.3125_-120.5625 and
                   three_hourly_48.3125_-120.5625, stored as float32) and the 90-day state of
                   state_vic.nc, exactly as that test configures it (NetCDF-4 / HDF5 with densely
                   stored links and chunked datasets: read by oracle/mini_hdf5.py).
* ``stehekin_domain`` inputs of the whole 4 x 5 Stehekin domain (16 cells, 1949) for the driver tests that
                   mirror the reference's own (variable rename, unit conversion, passthrough, examples).
* ``testdomain``   BASELINE config #1 as literally configured by examples/example_nc.conf:
                   metsim/data/test.nc + domain.nc + the 90-day state of state_nc.nc (81 cells, 31
                   days, 30 min, triangle, utc_offset), float32 inputs upcast exactly to float64.
                   state_nc.nc is NetCDF-4 / HDF5; it is read by oracle/mini_hdf5.py (contiguous,
                   uncompressed datasets).  Reference outputs are kept for 9 of the 81 cells.
* ``synth_*``      seeded synthetic cells sweeping time step, precipitation type, utc offset,
                   long-wave options, hemispheres, a polar cell and a supplied ``t_end``.
* ``passthrough_*`` the 'passthrough' method (metsim/methods/passthrough.py): supplied daily
                   shortwave (+ vapour pressure / tskc / longwave), both ``sw_averaging`` modes,
                   the longwave rescale of disaggregate.py:583-589.
"""
from __future__ import annotations

import os
import struct
import sys

import numpy as np
import pandas as pd
from scipy.io import netcdf_file

from oracle import reference_harness as rh
from oracle.mini_hdf5 import read_hdf5
from oracle.validate_oracle import synth_cell

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SUB = ("temp", "prec", "shortwave", "longwave", "vapor_pressure", "rel_humid", "air_pressure",
       "spec_humid", "tskc", "wind")
DAILY = ("t_day", "tfmax", "tskc", "tdew", "vapor_pressure", "shortwave", "pet", "daylength",
         "potrad", "tt_max", "seasonal_prec", "smoothed_dtr")


def pack(cells, params, outs):
    """cells: list of cell dicts (same length records) -> dict of stacked arrays [D][N]."""
    st = lambda k: np.stack([np.asarray(c[k], dtype=np.float64) for c in cells], axis=1)
    d = len(cells[0]["t_min"])
    dates = pd.date_range(cells[0]["start"], periods=d, freq="D")
    z = dict(lat=np.array([c["lat"] for c in cells]), lon=np.array([c["lon"] for c in cells]),
             elev=np.array([c["elev"] for c in cells]), start=np.array(cells[0]["start"]),
             doy=dates.dayofyear.values.astype(np.int32), month=dates.month.values.astype(np.int32),
             t_min=st("t_min"), t_max=st("t_max"), prec=st("prec"), wind=st("wind"),
             state_t_min=st("state_t_min"), state_t_max=st("state_t_max"),
             state_prec=st("state_prec"), t_pk12=st("t_pk12"), dur12=st("dur12"))
    if cells[0].get("t_end") is not None:
        z["t_end"] = st("t_end")
    for k in ("shortwave", "vapor_pressure", "tskc", "longwave"):   # passthrough forcing columns
        if cells[0].get(k) is not None:
            z["given_" + k] = st(k)
    for k, v in params.items():
        z["param_" + k] = np.array(v)
    for k in DAILY:
        z["ref_daily_" + k] = np.stack([o["daily"][k] for o in outs], axis=1)
    if "sub" in outs[0]:
        for k in SUB:
            z["ref_sub_" + k] = np.stack([o["sub"][k] for o in outs], axis=1)
    return z


def read_vic_binary(path, n_days):
    """VIC4 binary: u16/40, i16/100 x3 interleaved (metsim/io.py:324-352)."""
    raw = open(path, "rb").read(8 * n_days)
    vals = struct.iter_unpack("<Hhhh", raw)
    a = np.array(list(vals), dtype=np.float64)
    return a[:, 0] / 40.0, a[:, 1] / 100.0, a[:, 2] / 100.0, a[:, 3] / 100.0


def stehekin(root):
    data = os.path.join(root, "metsim", "data")
    dom = netcdf_file(os.path.join(data, "stehekin.nc"), "r", mmap=False)
    la, lo = 1, 4
    lat, lon = float(dom.variables["lat"].data[la]), float(dom.variables["lon"].data[lo])
    elev = float(dom.variables["elev"].data[la, lo])
    assert abs(lat - 48.3125) < 1e-9 and abs(lon + 120.5625) < 1e-9
    prec, tmax, tmin, wind = read_vic_binary(os.path.join(data, "binary", "data_48.3125_-120.5625"), 365)
    state = read_hdf5(os.path.join(data, "state_vic.nc"))      # test_metsim.py:408 'state'
    assert abs(state["lat"][la] - lat) < 1e-9 and abs(state["lon"][lo] - lon) < 1e-9
    assert state["t_min"].shape == (90, 4, 5) and (np.diff(state["time"]) == 1).all()
    st = {k: state[k][:, la, lo].astype(np.float64) for k in ("t_min", "t_max", "prec")}
    assert all(np.isfinite(v).all() and np.abs(v).max() < 1e3 for v in st.values())   # no fill values here
    cell = dict(lat=lat, lon=lon, elev=elev, start="1949-01-01", t_min=tmin, t_max=tmax, prec=prec,
                wind=wind, state_t_min=st["t_min"], state_t_max=st["t_max"], state_prec=st["prec"],
                t_pk12=dom.variables["t_pk"].data[:, la, lo].astype(np.float64),
                dur12=dom.variables["dur"].data[:, la, lo].astype(np.float64))
    for ts, fname, tag in ((60, "validated_48.3125_-120.5625", "hourly"),
                           (180, "three_hourly_48.3125_-120.5625", "3hourly")):
        params = dict(time_step=ts)  # reference defaults otherwise: uniform prec, no utc offset
        out = rh.run_cell(cell, params)
        z = pack([cell], params, [out])
        good = pd.read_table(os.path.join(data, fname), header=None).values.astype(np.float32)
        z["repo_golden"] = good
        z["repo_golden_columns"] = np.array(["prec", "temp", "shortwave", "longwave", "vapor_pressure",
                                             "wind", "rel_humid", "spec_humid", "air_pressure"])
        np.savez_compressed(os.path.join(GOLDEN, f"stehekin_{tag}.npz"), **z)
        print("stehekin", tag, good.shape)


def stehekin_domain(root):
    """The whole Stehekin domain of the reference's driver-level tests (metsim/tests/test_metsim.py:212-395:
    time offset, variable rename, passthrough, unit conversion; examples/example_bin.conf): stehekin.nc
    (4 x 5, 16 active cells), the 16 VIC-4 binary forcing files cut to 1949, the 90-day state of
    state_vic.nc -- inputs only, as arrays; tests/test_gpu_driver.py rebuilds the files from them."""
    data = os.path.join(root, "metsim", "data")
    dom = netcdf_file(os.path.join(data, "stehekin.nc"), "r", mmap=False)
    lat, lon = dom.variables["lat"].data.astype(float), dom.variables["lon"].data.astype(float)
    mask = np.nan_to_num(dom.variables["mask"].data.astype(float), nan=0.0)
    state = read_hdf5(os.path.join(data, "state_vic.nc"))
    assert np.allclose(state["lat"], lat) and np.allclose(state["lon"], lon) and (np.diff(state["time"]) == 1).all()
    z = dict(lat=lat, lon=lon, mask=mask.astype(np.int32), elev=dom.variables["elev"].data.astype(float),
             t_pk=dom.variables["t_pk"].data.astype(np.float32), dur=dom.variables["dur"].data.astype(np.float32),
             start=np.array("1949-01-01"),
             state_t_min=state["t_min"], state_t_max=state["t_max"], state_prec=state["prec"])
    for k in ("prec", "t_max", "t_min", "wind"):
        z[k] = np.full((365, lat.size, lon.size), np.nan)
    n = 0
    for i, la in enumerate(lat):
        for j, lo in enumerate(lon):
            path = os.path.join(data, "binary", f"data_{la}_{lo}")
            if mask[i, j] > 0 and os.path.exists(path):
                z["prec"][:, i, j], z["t_max"][:, i, j], z["t_min"][:, i, j], z["wind"][:, i, j] = read_vic_binary(path, 365)
                n += 1
    assert n == int((mask > 0).sum()) == 16
    np.savez_compressed(os.path.join(GOLDEN, "inputs_stehekin_domain.npz"), **z)
    print("stehekin_domain", n, "cells")


def testdomain(root):
    data = os.path.join(root, "metsim", "data")
    frc = netcdf_file(os.path.join(data, "test.nc"), "r", mmap=False)
    dom = netcdf_file(os.path.join(data, "domain.nc"), "r", mmap=False)
    lat, lon = dom.variables["lat"].data.astype(float), dom.variables["lon"].data.astype(float)
    state = read_hdf5(os.path.join(data, "state_nc.nc"))      # example_nc.conf:17, [state_vars]
    assert np.allclose(state["lat"], lat) and np.allclose(state["lon"], lon)
    assert state["t_min"].shape == (90, 9, 9) and (np.diff(state["time"]) == 1).all()
    cells = []
    for i in range(9):
        for j in range(9):
            g = lambda n: frc.variables[n].data[:, i, j].astype(np.float64)
            cells.append(dict(lat=lat[i], lon=lon[j], elev=float(dom.variables["elev"].data[i, j]),
                              start="1950-01-01", t_min=g("Tmin"), t_max=g("Tmax"), prec=g("Prec"),
                              wind=g("wind"), state_t_min=state["t_min"][:, i, j].copy(),
                              state_t_max=state["t_max"][:, i, j].copy(),
                              state_prec=state["prec"][:, i, j].copy(),
                              t_pk12=dom.variables["t_pk"].data[:, i, j].astype(np.float64),
                              dur12=dom.variables["dur"].data[:, i, j].astype(np.float64)))
    params = dict(time_step=30, prec_type="triangle", utc_offset=True)  # examples/example_nc.conf
    keep = [0, 10, 20, 30, 40, 50, 60, 70, 80]
    outs = [rh.run_cell(cells[k], params) for k in keep]
    z = pack(cells, params, outs)
    z["ref_cells"] = np.array(keep, dtype=np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "testdomain.npz"), **z)
    print("testdomain", len(cells), "cells, reference outputs for", len(keep))


def synth():
    rng = np.random.default_rng(20261019)
    groups = {
        "synth_hourly_triangle_utc": (dict(time_step=60, prec_type="triangle", utc_offset=True), "2001-11-15", 75),
        "synth_3hourly_mix_utc": (dict(time_step=180, prec_type="mix", utc_offset=True), "2003-12-20", 80),
        "synth_30min_uniform_local": (dict(time_step=30, prec_type="uniform", utc_offset=False), "2000-02-10", 40),
        "synth_15min_triangle_utc": (dict(time_step=15, prec_type="triangle", utc_offset=True), "1999-12-01", 35),
        "synth_6hourly_mix_tend": (dict(time_step=360, prec_type="mix", utc_offset=True), "2004-02-20", 45),
        "synth_2hourly_anderson_tva": (dict(time_step=120, prec_type="triangle", utc_offset=True,
                                            lw_type="anderson", lw_cloud="default", sw_prec_thresh=0.5,
                                            rain_scalar=0.7, tday_coef=0.5, tmax_daylength_fraction=0.6,
                                            tdew_tol=1e-8), "2004-06-01", 40),
        "synth_daily": (dict(time_step=1440), "2001-01-01", 120),
    }
    sites = [(48.3125, -120.5625, 1668.15), (30.03125, -100.03125, 519.0), (-33.9, 151.2, 30.0),
             (0.25, 0.0, 2200.0), (64.9, 179.9, 410.0), (70.3, -179.75, 5.0), (-54.75, 25.5, 0.0),
             (83.75, -45.25, 800.0), (25.03125, -67.03125, 3900.0), (41.2, -105.1, 1900.0)]
    for name, (params, start, d) in groups.items():
        cells = [synth_cell(rng, la, lo, el, start, d) for la, lo, el in sites]
        if name.endswith("tend"):
            for c in cells:
                c["t_end"] = [c["t_min"][-1] - 1.5, c["t_max"][-1] + 2.0]
        outs = [rh.run_cell(c, params) for c in cells]
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **pack(cells, params, outs))
        print(name, len(cells), "cells x", d, "days")


def passthrough():
    rng = np.random.default_rng(20261020)
    groups = {
        "passthrough_hourly_sw_lw": (dict(method="passthrough", time_step=60, prec_type="triangle",
                                          utc_offset=True), "2002-06-10", 40,
                                     ("shortwave", "longwave")),
        "passthrough_3hourly_daylight_all": (dict(method="passthrough", sw_averaging="daylight",
                                                  time_step=180, prec_type="mix", utc_offset=True),
                                             "2003-12-15", 35,
                                             ("shortwave", "vapor_pressure", "tskc", "longwave")),
        "passthrough_30min_sw_vp": (dict(method="passthrough", time_step=30, prec_type="uniform",
                                         utc_offset=False, lw_type="idso", lw_cloud="default"),
                                    "2000-02-20", 20, ("shortwave", "vapor_pressure")),
    }
    sites = [(48.3125, -120.5625, 1668.15), (30.03125, -100.03125, 519.0), (-33.9, 151.2, 30.0),
             (0.25, 0.0, 2200.0), (64.9, 179.9, 410.0), (70.3, -179.75, 5.0), (-54.75, 25.5, 0.0),
             (41.2, -105.1, 1900.0)]
    draw = {"shortwave": (20, 350), "vapor_pressure": (200, 1800), "tskc": (0, 1), "longwave": (180, 420)}
    for name, (params, start, d, given) in groups.items():
        cells = [synth_cell(rng, la, lo, el, start, d) for la, lo, el in sites]
        for c in cells:
            for k in given:
                c[k] = rng.uniform(*draw[k], d)
        outs = [rh.run_cell(c, params) for c in cells]
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **pack(cells, params, outs))
        print(name, len(cells), "cells x", d, "days")


def main():
    root = rh.find_reference_root()
    if root is None:
        sys.exit("reference checkout not found")
    os.makedirs(GOLDEN, exist_ok=True)
    import warnings
    warnings.filterwarnings("ignore")
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    steps = dict(stehekin=lambda: stehekin(root), stehekin_domain=lambda: stehekin_domain(root),
                 testdomain=lambda: testdomain(root), synth=synth, passthrough=passthrough)
    for name, fn in steps.items():
        if not only or name in only:
            fn()


if __name__ == "__main__":
    main()
