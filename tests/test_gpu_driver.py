"""The minimal driver end to end on the GPU: configuration file + NetCDF inputs on disk ->
``python -m metsim_b200.driver`` -> NetCDF outputs, compared with what the unmodified reference
produced for the same inputs (tests/golden/*.npz).  The input files are rebuilt here from the
fixtures (the reference checkout does not exist on the GPU box)."""
import os
import textwrap

import numpy as np
import pytest

import msb_testutil as util
from test_driver_host import write_nc3

pytestmark = pytest.mark.gpu


def _files_from_fixture(tmp, inputs, ny, nx, start, mask=None, f4=True):
    """domain.nc / forcing.nc / state.nc (NetCDF classic) holding a latitude-row-major fixture."""
    lat, lon = np.unique(inputs["lat"]), np.unique(inputs["lon"])
    assert lat.size == ny and lon.size == nx
    g = lambda a: a.reshape(a.shape[0], ny, nx)
    days = inputs["t_min"].shape[0]
    d0 = np.datetime64(start, "D")
    mask = np.ones((ny, nx), "i4") if mask is None else mask
    fd = "f4" if f4 else "f8"      # the bundled files hold float32 forcings; the fixture upcast them exactly
    dom = {"lat": (("lat",), lat, {}), "lon": (("lon",), lon, {}), "mask": (("lat", "lon"), mask, {}),
           "elev": (("lat", "lon"), inputs["elev"].reshape(ny, nx), {"units": "m"})}
    if "t_pk12" in inputs:
        dom["t_pk"] = (("month", "lat", "lon"), g(inputs["t_pk12"]).astype(fd), {})
        dom["dur"] = (("month", "lat", "lon"), g(inputs["dur12"]).astype(fd), {})
    write_nc3(os.path.join(tmp, "domain.nc"), {"lat": ny, "lon": nx, "month": 12}, dom)
    t = (d0 - np.datetime64("1900-01-01", "D")).astype(int) + np.arange(days)
    write_nc3(os.path.join(tmp, "forcing.nc"), {"time": days, "lat": ny, "lon": nx}, {
        "time": (("time",), t.astype("f8"), {"units": "days since 1900-01-01 00:00:00", "calendar": "standard"}),
        "lat": (("lat",), lat, {}), "lon": (("lon",), lon, {}),
        "Prec": (("time", "lat", "lon"), g(inputs["prec"]).astype(fd), {}),
        "Tmax": (("time", "lat", "lon"), g(inputs["t_max"]).astype(fd), {}),
        "Tmin": (("time", "lat", "lon"), g(inputs["t_min"]).astype(fd), {}),
        "wind": (("time", "lat", "lon"), g(inputs["wind"]).astype(fd), {})})
    ts = (d0 - 90 - np.datetime64("1900-01-01", "D")).astype(int) + np.arange(90)
    write_nc3(os.path.join(tmp, "state.nc"), {"time": 90, "lat": ny, "lon": nx}, {
        "time": (("time",), ts.astype("i4"), {"units": "days since 1900-01-01", "calendar": "proleptic_gregorian"}),
        "lat": (("lat",), lat, {}), "lon": (("lon",), lon, {}),
        "t_min": (("time", "lat", "lon"), g(inputs["state_t_min"]), {}),
        "t_max": (("time", "lat", "lon"), g(inputs["state_t_max"]), {}),
        "prec": (("time", "lat", "lon"), g(inputs["state_prec"]), {})})
    return d0, days


def _conf(tmp, *, ts, start, stop, body="", out_vars=None, sections=""):
    path = os.path.join(tmp, "run.conf")
    open(path, "w").write(textwrap.dedent("""\
        [MetSim]
        {ov}
        time_step = {ts}
        start = {start}
        stop = {stop}
        forcing = {tmp}/forcing.nc
        domain = {tmp}/domain.nc
        state = {tmp}/state.nc
        forcing_fmt = netcdf
        out_dir = {tmp}/results
        out_prefix = forcing
        {body}
        [chunks]
        lat = 3
        lon = 3
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
        {dv}
        {sections}
        """).format(ov=f"out_vars = {out_vars}" if out_vars else "", ts=ts, start=start, stop=stop, tmp=tmp,
                    body=body, sections=sections, dv="t_pk = t_pk\ndur = dur" if "triangle" in body or "mix" in body else ""))
    return path


def test_driver_runs_example_nc_conf_end_to_end(tmp_path):
    """BASELINE config #1 through the command line of the driver: the settings of
    examples/example_nc.conf (30 min, triangle, utc_offset, the 8 out_vars) on the bundled 9x9 test
    domain rebuilt from tests/golden/testdomain.npz, with five cells masked out.  The file holds
    float32 (out_precision f4, the reference default): compared with the reference's float64 outputs
    at float32 rounding."""
    from metsim_b200 import driver
    tmp = str(tmp_path)
    inputs, params, _, ref_sub, extra = util.load_golden(os.path.join(util.GOLDEN, "testdomain.npz"))
    cells = extra["ref_cells"]
    holes = [c for c in (3, 17, 40, 41, 80) if c not in set(cells.tolist())]
    mask = np.ones(81, "i4")
    mask[holes] = 0
    _files_from_fixture(tmp, inputs, 9, 9, "1950-01-01", mask=mask.reshape(9, 9))
    out_vars = "['temp', 'prec', 'shortwave', 'longwave', 'vapor_pressure', 'rel_humid', 'air_pressure', 'wind']"
    conf = _conf(tmp, ts=30, start="1950/1/1", stop="1950/1/31", out_vars=out_vars,
                 body="prec_type = triangle\nutc_offset = True")
    (path,) = driver.main([conf, "-n", "4", "-s", "threading"])
    assert os.path.basename(path) == "forcing_19500101-19500131.nc"
    from metsim_b200 import ncio
    ds = ncio.open_dataset(path)
    assert ds.dims == {"time": 1488, "lat": 9, "lon": 9}
    assert ds["time"].values[0] == np.datetime64("1950-01-01T00:00") and ds["time"].values[-1] == np.datetime64("1950-01-31T23:30")
    assert ds["time"].attrs["standard_name"] == "utc_time"
    for v in ("temp", "prec", "shortwave", "longwave", "vapor_pressure", "rel_humid", "air_pressure", "wind"):
        got = ds[v].values.reshape(1488, 81)
        assert got.dtype == np.float32
        assert np.isnan(got[:, holes]).all() and not np.isnan(np.delete(got, holes, axis=1)).any(), v
        e = util.relerr(got[:, cells], ref_sub[v], util.FLOORS[v])
        assert e <= 2e-7, (v, e)
        # zero / non-zero placement of precipitation is exact even in float32
        if v == "prec":
            assert np.array_equal(got[:, cells] == 0, ref_sub[v] == 0)
    assert ds["vapor_pressure"].attrs["units"] == "Pa" and ds.attrs["prec_type"] == "triangle"


def test_driver_tiles_units_f8_and_monthly_files(tmp_path):
    """The streaming form: a host budget that forces the 2-slot ring, out_precision f8 so that the
    1e-9 bar applies, converted units, out_name from an [out_vars] section, one file per month
    (out_freq), period-ending time labels -- hourly triangle fixture, cells on one latitude row."""
    from metsim_b200 import driver
    tmp = str(tmp_path)
    inputs, params, _, ref_sub, _ = util.load_golden(os.path.join(util.GOLDEN, "synth_hourly_triangle_utc.npz"))
    order = np.lexsort((inputs["lon"], inputs["lat"]))
    lat, lon = np.unique(inputs["lat"]), np.unique(inputs["lon"])
    if lat.size * lon.size != inputs["lat"].size:
        # scattered cells: put each on the diagonal of its own (lat, lon) grid, everything else masked
        ny, nx = lat.size, lon.size
        pos = np.searchsorted(lat, inputs["lat"]) * nx + np.searchsorted(lon, inputs["lon"])
        assert np.unique(pos).size == pos.size
        full = {}
        for k, a in inputs.items():
            if k in ("doy", "month", "lat", "lon"):
                continue
            if a.ndim == 1:
                g = np.zeros(ny * nx); g[pos] = a
            else:
                g = np.zeros((a.shape[0], ny * nx)); g[:, pos] = a
                if k in ("t_max", "state_t_max"):
                    g[:, np.setdiff1d(np.arange(ny * nx), pos)] = 1.0      # keep t_max >= t_min off the mask too
            full[k] = g
        full["lat"], full["lon"] = np.repeat(lat, nx), np.tile(lon, ny)
        mask = np.zeros(ny * nx, "i4"); mask[pos] = 1
    else:
        ny, nx, pos, full, mask = lat.size, lon.size, np.arange(inputs["lat"].size), inputs, np.ones(inputs["lat"].size, "i4")
    days = inputs["t_min"].shape[0]
    start = np.datetime64("2001-01-01") + int(inputs["doy"][0]) - 1
    from metsim_b200.engine import calendar_vectors
    d_chk, m_chk = calendar_vectors(str(start), days)
    if not ((d_chk == inputs["doy"]).all() and (m_chk == inputs["month"]).all()):
        pytest.skip("fixture calendar is not a 2001 record")
    _files_from_fixture(tmp, full, ny, nx, str(start), mask=mask.reshape(ny, nx), f4=False)
    stop = start + days - 1
    extra = "".join(f"\n{k} = {params[k]}" for k in params if k not in ("time_step", "prec_type", "utc_offset"))
    conf = _conf(tmp, ts=int(params["time_step"]), start=str(start).replace("-", "/"), stop=str(stop).replace("-", "/"),
                 body=f"prec_type = {params.get('prec_type', 'uniform')}\nutc_offset = {bool(params.get('utc_offset', False))}"
                      f"\nout_precision = f8\nout_freq = 1M\nperiod_ending = True{extra}",
                 sections="[out_vars]\ntemp = airtemp\nprec = pptrate\nspec_humid = spechum\ntskc = cloud")
    ms = driver.MetSimB200(driver.read_config(conf))
    ms.params["out_vars"]["temp"]["units"] = "K"
    ms.params["out_vars"]["prec"]["units"] = "mm s-1"
    ms._update_unit_attrs(ms.params["out_vars"])
    ts = int(params["time_step"])
    tpd = 1440 // ts
    files = ms.run(host_budget_bytes=2 * 5 * tpd * pos.size * 8 * 4 + 1)       # tiles of 5 days
    assert len(files) == len(ms._times) and len(files) >= 1
    got = {}
    step0 = 0
    for ds in ms.open_output():
        n = ds.dims["time"]
        assert ds["time"].values[0] == start.astype("datetime64[s]") + (step0 + 1) * np.timedelta64(ts, "m")
        for name in ("airtemp", "pptrate", "spechum", "cloud"):
            got.setdefault(name, []).append(ds[name].values.reshape(n, ny * nx))
        step0 += n
    assert step0 == days * tpd
    got = {k: np.concatenate(v, axis=0) for k, v in got.items()}
    off = np.setdiff1d(np.arange(ny * nx), pos)
    for name, v, sc, of in (("airtemp", "temp", 1.0, 273.15), ("pptrate", "prec", 1.0 / (ts * 60.0), 0.0),
                            ("spechum", "spec_humid", 1.0, 0.0), ("cloud", "tskc", 1.0, 0.0)):
        e = util.relerr(got[name][:, pos], ref_sub[v] * sc + of, util.FLOORS[v] * sc)
        assert e <= 2e-9, (name, e)
        assert np.isnan(got[name][:, off]).all()


def test_driver_daily_mode(tmp_path):
    """time_step = 1440: daily out_vars straight from the forcings and the daily pass
    (metsim/metsim.py:779-792), written by the driver."""
    from metsim_b200 import driver
    tmp = str(tmp_path)
    inputs, params, ref_daily, _, _ = util.load_golden(os.path.join(util.GOLDEN, "synth_daily.npz"))
    lat, lon = np.unique(inputs["lat"]), np.unique(inputs["lon"])
    ny, nx = lat.size, lon.size
    pos = np.searchsorted(lat, inputs["lat"]) * nx + np.searchsorted(lon, inputs["lon"])
    if np.unique(pos).size != pos.size:
        pytest.skip("fixture cells share grid positions")
    days = inputs["t_min"].shape[0]
    start = np.datetime64("2001-01-01") + int(inputs["doy"][0]) - 1
    from metsim_b200.engine import calendar_vectors
    d_chk, m_chk = calendar_vectors(str(start), days)
    if not ((d_chk == inputs["doy"]).all() and (m_chk == inputs["month"]).all()):
        pytest.skip("fixture calendar is not a 2001 record")
    full = {}
    for k, a in inputs.items():
        if k in ("doy", "month", "lat", "lon") or a.ndim == 0:
            continue
        if a.ndim == 1:
            g = np.zeros(ny * nx); g[pos] = a
        else:
            g = np.zeros((a.shape[0], ny * nx)); g[:, pos] = a
            if k in ("t_max", "state_t_max"):
                g[:, np.setdiff1d(np.arange(ny * nx), pos)] = 1.0
        full[k] = g
    full["lat"], full["lon"] = np.repeat(lat, nx), np.tile(lon, ny)
    mask = np.zeros(ny * nx, "i4"); mask[pos] = 1
    _files_from_fixture(tmp, full, ny, nx, str(start), mask=mask.reshape(ny, nx), f4=False)
    stop = start + days - 1
    extra = "".join(f"\n{k} = {params[k]}" for k in params if k not in ("time_step", "prec_type", "utc_offset"))
    conf = _conf(tmp, ts=1440, start=str(start).replace("-", "/"), stop=str(stop).replace("-", "/"),
                 out_vars="['t_min', 't_max', 'prec', 'vapor_pressure', 'shortwave', 'tskc', 'pet', 'wind', 'daylength']",
                 body=f"out_precision = f8\nutc_offset = {bool(params.get('utc_offset', False))}{extra}")
    ms = driver.MetSimB200(driver.read_config(conf))
    (path,) = ms.run()
    (ds,) = ms.open_output()
    flat = lambda v: ds[v].values.reshape(days, ny * nx)[:, pos]
    np.testing.assert_array_equal(flat("t_min"), inputs["t_min"])
    np.testing.assert_array_equal(flat("prec"), inputs["prec"])
    np.testing.assert_array_equal(flat("wind"), inputs["wind"])
    util.assert_close(flat("shortwave"), ref_daily["shortwave"] * ref_daily["daylength"] / 86400.0, "sw_daily")
    # (t_day is in the reference's daily list, metsim.py:660, but has no entry in its attrs / converters
    # tables, so no configuration file can name it there either)
    for v, o in (("vapor_pressure", "vp_daily"), ("tskc", "tskc_daily"), ("pet", "pet"), ("daylength", "daylength")):
        util.assert_close(flat(v), ref_daily[v], o, context="driver daily mode")
    assert ds["time"].values.size == days and ds["pet"].attrs["units"] == "mm timestep-1"


@pytest.mark.parametrize("budget", [1 << 30, 60000])
def test_driver_cell_ranges_on_several_contexts_equal_the_single_call(tmp_path, budget):
    """``run(devices=[0, 0])``: two contiguous cell ranges, each on its own host thread and
    msb_host_ctx (here both on device 0 -- contexts are independent, tests/test_gpu_parity.py), gathered
    on the host.  Whole columns when the record fits the budget, the tile barrier otherwise.  The files
    must equal the single-call run bit for bit, and the reference within float32 rounding."""
    from metsim_b200 import driver, ncio
    tmp = str(tmp_path)
    inputs, params, _, ref_sub, extra = util.load_golden(os.path.join(util.GOLDEN, "testdomain.npz"))
    cells = extra["ref_cells"]
    holes = [c for c in (3, 17, 40, 41, 80) if c not in set(cells.tolist())]
    mask = np.ones(81, "i4")
    mask[holes] = 0
    _files_from_fixture(tmp, inputs, 9, 9, "1950-01-01", mask=mask.reshape(9, 9))
    out_vars = "['temp', 'prec', 'shortwave', 'longwave', 'vapor_pressure', 'rel_humid', 'air_pressure', 'wind']"
    conf = _conf(tmp, ts=30, start="1950/1/1", stop="1950/1/31", out_vars=out_vars,
                 body="prec_type = triangle\nutc_offset = True")
    ms = driver.MetSimB200(driver.read_config(conf))
    (path,) = ms.run()
    one = {v: ncio.open_dataset(path)[v].values.copy() for v in ("temp", "prec", "shortwave", "longwave",
                                                                 "vapor_pressure", "rel_humid", "air_pressure", "wind")}
    os.remove(path)
    (path,) = ms.run(host_budget_bytes=budget, devices=[0, 0])
    from metsim_b200.shard import cell_range
    n_act = int(mask.sum())
    assert [r["cells"] for r in ms.timings["ranges"]] == [hi - lo for lo, hi in (cell_range(n_act, k, 2) for k in range(2))]
    two = ncio.open_dataset(path)
    for v, a in one.items():
        np.testing.assert_array_equal(two[v].values, a, err_msg=v)
        e = util.relerr(a.reshape(1488, 81)[:, cells], ref_sub[v], util.FLOORS[v])
        assert e <= 2e-7, (v, e)
