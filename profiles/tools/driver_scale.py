"""Files in -> files out through the minimal driver at a realistic size (one GPU).

    python profiles/tools/driver_scale.py [--rows 160] [--cols 250] [--days 90] [--dir /tmp/msb_driver]

Writes a dense block of the CONUS 1/16 deg grid (synthetic forcings of SURVEY 8d, float32 like the
reference's own forcing files) as domain / forcing / state NetCDF files plus an ini configuration
with the settings of examples/example_nc.conf (hourly here), then runs
``metsim_b200.driver.main([conf])`` and prints one JSON line: seconds for reading the inputs, for
``run()`` (gather + msb_run_host + the streamed NetCDF write, page cache included) and the output
volume.  A spot check compares a few cells of the file with a direct Engine run.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=160)
    ap.add_argument("--cols", type=int, default=250)
    ap.add_argument("--days", type=int, default=90)
    ap.add_argument("--dir", default="/tmp/msb_driver")
    ap.add_argument("--budget-gb", type=float, default=1.0, help="host ring budget (forces tiling)")
    ap.add_argument("--devices", default=None, help="comma-separated device indices, one cell range each")
    a = ap.parse_args()
    from scipy.io import netcdf_file
    from metsim_b200 import driver, synth
    os.makedirs(a.dir, exist_ok=True)
    rng = np.random.default_rng(synth.SEED)
    lat = 25.03125 + np.arange(a.rows) / 16.0
    lon = -124.96875 + np.arange(a.cols) / 16.0
    la, lo = np.meshgrid(lat, lon, indexing="ij")
    start = "2001-03-01"
    f = synth.make_forcing(la.ravel(), lo.ravel(), a.days, start, rng)
    ny, nx, d = a.rows, a.cols, a.days
    g = lambda x: x.reshape(x.shape[0], ny, nx)

    def write(path, dims, variables):
        with netcdf_file(path, "w", version=2) as nc:
            for k, n in dims.items():
                nc.createDimension(k, n)
            for name, (vd, vals, atts) in variables.items():
                vals = np.asarray(vals)
                v = nc.createVariable(name, vals.dtype.char, vd)
                v[:] = vals
                for k, t in atts.items():
                    setattr(v, k, t)

    d0 = np.datetime64(start, "D")
    coords = {"lat": (("lat",), lat, {}), "lon": (("lon",), lon, {})}
    write(os.path.join(a.dir, "domain.nc"), {"lat": ny, "lon": nx, "month": 12}, dict(
        coords, mask=(("lat", "lon"), np.ones((ny, nx), "i4"), {}),
        elev=(("lat", "lon"), f["elev"].reshape(ny, nx), {}),
        t_pk=(("month", "lat", "lon"), g(f["t_pk12"]).astype("f4"), {}),
        dur=(("month", "lat", "lon"), g(f["dur12"]).astype("f4"), {})))
    t = (d0 - np.datetime64("1900-01-01", "D")).astype(int) + np.arange(d)
    tu = {"units": "days since 1900-01-01 00:00:00", "calendar": "standard"}
    write(os.path.join(a.dir, "forcing.nc"), {"time": d, "lat": ny, "lon": nx}, dict(
        coords, time=(("time",), t.astype("f8"), tu),
        Prec=(("time", "lat", "lon"), g(f["prec"]).astype("f4"), {}),
        Tmax=(("time", "lat", "lon"), g(f["t_max"]).astype("f4"), {}),
        Tmin=(("time", "lat", "lon"), g(f["t_min"]).astype("f4"), {}),
        wind=(("time", "lat", "lon"), g(f["wind"]).astype("f4"), {})))
    write(os.path.join(a.dir, "state.nc"), {"time": 90, "lat": ny, "lon": nx}, dict(
        coords, time=(("time",), (t[0] - 90 + np.arange(90)).astype("f8"), tu),
        t_min=(("time", "lat", "lon"), g(f["state_t_min"]), {}),
        t_max=(("time", "lat", "lon"), g(f["state_t_max"]), {}),
        prec=(("time", "lat", "lon"), g(f["state_prec"]), {})))
    stop = str(d0 + d - 1).replace("-", "/")
    conf = os.path.join(a.dir, "run.conf")
    open(conf, "w").write(f"""[MetSim]
out_vars = ['temp', 'prec', 'shortwave', 'longwave', 'vapor_pressure', 'rel_humid', 'air_pressure', 'wind']
time_step = 60
start = {start.replace('-', '/')}
stop = {stop}
forcing = {a.dir}/forcing.nc
domain = {a.dir}/domain.nc
state = {a.dir}/state.nc
forcing_fmt = netcdf
out_dir = {a.dir}/results
out_prefix = forcing
prec_type = triangle
utc_offset = True
[chunks]
lat = 40
lon = 50
[forcing_vars]
prec = Prec
t_max = Tmax
t_min = Tmin
wind = wind
[state_vars]
prec = prec
t_max = t_max
t_min = t_min
[domain_vars]
lat = lat
lon = lon
mask = mask
elev = elev
t_pk = t_pk
dur = dur
""")
    import torch
    torch.cuda.init()
    t0 = time.perf_counter()
    ms = driver.MetSimB200(driver.read_config(conf))
    ms.state
    t1 = time.perf_counter()
    devs = [int(x) for x in a.devices.split(",")] if a.devices else None
    ms.run(host_budget_bytes=int(a.budget_gb * (1 << 30)), devices=devs)     # first run: arena, pinned ring, page cache
    t2 = time.perf_counter()
    files = ms.run(host_budget_bytes=int(a.budget_gb * (1 << 30)), devices=devs)
    t3 = time.perf_counter()
    out_bytes = sum(os.path.getsize(p) for p in files)
    cts = ny * nx * d * 24
    # spot check against a direct device-resident run of the same float32-rounded inputs
    from metsim_b200.engine import Engine
    gth = ms.gather()
    gth.pop("cell_index")
    eng = Engine(ms.params)
    eng.load(**gth)
    ref = eng.run(out_vars=("temp", "prec", "longwave"))
    ds, = ms.open_output()
    pick = np.linspace(0, ny * nx - 1, 64).astype(int)
    same = all(np.array_equal(ds[v].values.reshape(d * 24, -1)[:, pick], ref[v].cpu().numpy()[:, pick])
               for v in ("temp", "prec", "longwave"))
    tm = {k: (round(v, 3) if not isinstance(v, list) else [{kk: (round(vv, 2) if isinstance(vv, float) else vv)
                                                                for kk, vv in r.items()} for r in v])
          for k, v in ms.timings.items()}
    print(json.dumps({"devices": devs, "cells": ny * nx, "days": d, "cell_timesteps": cts, "read_inputs_s": round(t1 - t0, 3),
                      "first_run_s": round(t2 - t1, 3), "run_s": round(t3 - t2, 3),
                      "cell_timesteps_per_s_files_to_files": cts / (t3 - t2),
                      "output_gb": round(out_bytes / 1e9, 3), "output_gb_per_s": round(out_bytes / 1e9 / (t3 - t2), 2),
                      "timings": tm,
                      "file_equals_direct_engine_run_on_64_cells": bool(same)}))


if __name__ == "__main__":
    main()
