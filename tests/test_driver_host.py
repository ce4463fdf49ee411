"""Host side of the minimal driver (metsim_b200/driver.py, ncio.py, hdf5_min.py) without a GPU:
configuration files, input readers, time bookkeeping, checks, and the tile -> file plumbing of
``MetSimB200.run`` with the CUDA call replaced by a pattern generator (the real call is covered by
the ``-m gpu`` test ``test_driver_runs_example_nc_conf_end_to_end``)."""
import os
import textwrap

import numpy as np
import pytest
from scipy.io import netcdf_file

import msb_testutil as util
from metsim_b200 import driver, ncio

REF = "/root/reference"
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "metsim", "data")),
                                     reason="the reference checkout is not on this machine")


def write_nc3(path, dims, variables, version=2):
    """variables: {name: (dims, values, attrs)}; written with scipy (classic / 64-bit offset)."""
    with netcdf_file(path, "w", version=version) as f:
        for d, n in dims.items():
            f.createDimension(d, n)
        for name, (vdims, values, attrs) in variables.items():
            values = np.asarray(values)
            v = f.createVariable(name, values.dtype.newbyteorder("=").char if values.dtype.kind != "f"
                                 else values.dtype.char, vdims)
            v[:] = values
            for k, a in (attrs or {}).items():
                setattr(v, k, a)


def small_case(tmp, *, ny=3, nx=4, days=10, ts=360, start="2001-01-25", out_freq=None, fmt="ini",
               mask=None, extra=""):
    """A tiny (lat, lon) domain + forcing + state on disk and a configuration naming them."""
    rng = np.random.default_rng(7)
    lat = 40.0 + 0.5 * np.arange(ny)
    lon = -110.0 + 0.5 * np.arange(nx)
    mask = np.ones((ny, nx), np.int32) if mask is None else np.asarray(mask, np.int32)
    d0 = np.datetime64(start, "D")
    write_nc3(os.path.join(tmp, "domain.nc"), {"lat": ny, "lon": nx, "month": 12}, {
        "lat": (("lat",), lat, {}), "lon": (("lon",), lon, {}),
        "mask": (("lat", "lon"), mask, {}),
        "elev": (("lat", "lon"), 500.0 + rng.random((ny, nx)) * 1000, {"units": "m"}),
        "t_pk": (("month", "lat", "lon"), rng.uniform(300, 1140, (12, ny, nx)).astype("f4"), {}),
        "dur": (("month", "lat", "lon"), rng.uniform(120, 720, (12, ny, nx)).astype("f4"), {})})
    # forcing: a longer record than the run, stamped at noon (the reference shifts by 11:59:59 and rounds)
    nrec = days + 6
    t_forc = (d0 - 3 - np.datetime64("1900-01-01", "D")).astype(int) + np.arange(nrec) + 0.5
    tmin = rng.normal(0, 5, (nrec, ny, nx)).astype("f4")
    frc = {"time": (("time",), t_forc, {"units": "days since 1900-01-01 00:00:00", "calendar": "standard"}),
           "lat": (("lat",), lat, {}), "lon": (("lon",), lon, {}),
           "Tmin": (("time", "lat", "lon"), tmin, {}),
           "Tmax": (("time", "lat", "lon"), tmin + rng.uniform(2, 12, tmin.shape).astype("f4"), {}),
           "Prec": (("time", "lat", "lon"), rng.gamma(0.8, 6, tmin.shape).astype("f4"), {"_FillValue": np.float32(1e20)}),
           "wind": (("time", "lat", "lon"), rng.uniform(0.5, 8, tmin.shape).astype("f4"), {})}
    write_nc3(os.path.join(tmp, "forcing.nc"), {"time": nrec, "lat": ny, "lon": nx}, frc)
    t_state = (d0 - 90 - np.datetime64("1990-01-01", "D")).astype(int) + np.arange(90)
    smin = rng.normal(0, 5, (90, ny, nx))
    write_nc3(os.path.join(tmp, "state.nc"), {"time": 90, "lat": ny, "lon": nx}, {
        "time": (("time",), t_state.astype("i4"), {"units": "days since 1990-01-01", "calendar": "proleptic_gregorian"}),
        "lat": (("lat",), lat, {}), "lon": (("lon",), lon, {}),
        "t_min": (("time", "lat", "lon"), smin, {}), "t_max": (("time", "lat", "lon"), smin + 8.0, {}),
        "prec": (("time", "lat", "lon"), rng.gamma(0.8, 6, smin.shape), {})})
    stop = str(d0 + days - 1)
    if fmt == "ini":
        conf = os.path.join(tmp, "run.conf")
        open(conf, "w").write(textwrap.dedent(f"""\
            # a comment first
            [MetSim]
            out_vars = ['temp', 'prec', 'shortwave', 'wind']   # list variant
            time_step = {ts}
            start = {str(d0).replace('-', '/')}
            stop = {stop.replace('-', '/')}
            forcing = {tmp}/forcing.nc
            domain = {tmp}/domain.nc
            state = {tmp}/state.nc
            forcing_fmt = netcdf
            out_dir = {tmp}/results
            out_prefix = small
            prec_type = triangle
            utc_offset = True
            {('out_freq = ' + out_freq) if out_freq else ''}
            {extra}
            [chunks]
            lat = 2
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
            """))
    else:
        conf = os.path.join(tmp, "run.yaml")
        open(conf, "w").write(textwrap.dedent(f"""\
            MetSim:
                time_step: {ts}
                start: {start}
                stop: {stop}
                forcing: '{tmp}/forcing.nc'
                domain: '{tmp}/domain.nc'
                state: '{tmp}/state.nc'
                forcing_fmt: 'netcdf'
                out_dir: '{tmp}/results'
                out_prefix: 'small_yaml'
                utc_offset: True
            out_vars:
                temp:
                    out_name: 'airtemp'
                    units: 'K'
                prec:
                    out_name: 'pptrate'
                    units: 'mm s-1'
                shortwave:
                    out_name: 'SWradAtm'
            chunks:
                lat: 3
            forcing_vars:
                prec: 'Prec'
                t_max: 'Tmax'
                t_min: 'Tmin'
            state_vars:
                prec: 'prec'
                t_max: 't_max'
                t_min: 't_min'
            domain_vars:
                lat: 'lat'
                lon: 'lon'
                mask: 'mask'
                elev: 'elev'
            constant_vars:
                wind: 2.0
            """))
    return conf, dict(lat=lat, lon=lon, mask=mask, frc=frc, d0=d0)


def test_read_config_ini_matches_the_reference_layout(tmp_path):
    conf, _ = small_case(str(tmp_path))
    p = driver.read_config(conf, scheduler="threading", num_workers=3, verbose=True)
    assert p["time_step"] == "360" and p["utc_offset"] is True and p["period_ending"] is False
    assert p["prec_type"] == "triangle" and p["calendar"] == "standard"
    assert p["forcing_vars"] == {"Prec": "prec", "Tmax": "t_max", "Tmin": "t_min", "wind": "wind"}   # file -> metsim
    assert p["domain_vars"]["elev"] == "elev" and p["state_vars"] == {"prec": "prec", "t_max": "t_max", "t_min": "t_min"}
    assert list(p["out_vars"]) == ["temp", "prec", "shortwave", "wind"]
    assert p["out_vars"]["prec"] == {"out_name": "prec", "units": "mm timestep-1"}
    assert p["scheduler"] == "threading" and p["num_workers"] == 3 and p["chunks"] == {"lat": "2"}
    assert os.path.isabs(p["out_dir"])


def test_read_config_yaml_out_names_units_and_constant_vars(tmp_path):
    conf, _ = small_case(str(tmp_path), fmt="yaml")
    assert driver.check_config_type(conf) == "yaml"
    p = driver.read_config(conf)
    assert p["out_vars"]["temp"] == {"out_name": "airtemp", "units": "K"}
    assert p["out_vars"]["shortwave"] == {"out_name": "SWradAtm", "units": "W m-2"}     # default units filled in
    assert p["constant_vars"] == {"wind": 2.0} and p["forcing_vars"]["Tmin"] == "t_min"
    ms = driver.MetSimB200(p)
    g = ms.gather()
    assert (g["wind"] == 2.0).all() and g["wind"].shape == g["t_min"].shape
    assert "t_pk12" not in g                                      # uniform precipitation needs no t_pk / dur


def test_inputs_time_selection_state_window_and_gather(tmp_path):
    mask = [[1, 0, 1, 1], [0, 0, 1, 0], [1, 1, 0, 1]]
    conf, case = small_case(str(tmp_path), mask=mask)
    ms = driver.MetSimB200(driver.read_config(conf))
    md = ms.met_data
    # noon stamps land on their own day; the run is cut to start .. stop out of the longer record
    assert md["time"].values[0] == np.datetime64("2001-01-25") and md["time"].values.size == 10
    assert ms.params["state_start"] == np.datetime64("2000-10-27") and ms.params["state_stop"] == np.datetime64("2001-01-24")
    g = ms.gather()
    sel = np.nonzero(np.asarray(mask).ravel())[0]
    assert g["cell_index"].tolist() == sel.tolist() and g["t_min"].shape == (10, 7)
    want = case["frc"]["Tmin"][1][3:13].reshape(10, -1)[:, sel].astype(np.float64)     # float32 upcast exactly
    np.testing.assert_array_equal(g["t_min"], want)
    assert g["doy"].tolist() == list(range(25, 32)) + [32, 33, 34] and g["month"].tolist() == [1] * 7 + [2] * 3
    np.testing.assert_array_equal(g["lat"], np.repeat(case["lat"], 4)[sel])
    np.testing.assert_array_equal(g["lon"], np.tile(case["lon"], 3)[sel])
    assert g["t_pk12"].shape == (12, 7) and g["state_prec"].shape == (90, 7)
    # output times: period-beginning, 4 per day; period_ending shifts the labels by one step
    assert ms._times[0][0] == np.datetime64("2001-01-25T00:00") and ms._times[0][-1] == np.datetime64("2001-02-03T18:00")
    assert ms._get_output_filename(ms._times[0]).endswith("results/small_20010125-20010203.nc")
    shifted = ms._get_output_times(period_ending=True)[0]
    assert shifted[0] == np.datetime64("2001-01-25T06:00") and shifted[-1] == np.datetime64("2001-02-04T00:00")


def test_forcing_grid_larger_than_the_domain_is_selected_by_coordinate(tmp_path):
    conf, case = small_case(str(tmp_path))
    # a domain that is a reversed subset of the forcing grid: rows 2, 0 and columns 3, 1
    rng = np.random.default_rng(3)
    write_nc3(os.path.join(str(tmp_path), "domain.nc"), {"lat": 2, "lon": 2, "month": 12}, {
        "lat": (("lat",), case["lat"][[2, 0]], {}), "lon": (("lon",), case["lon"][[3, 1]], {}),
        "mask": (("lat", "lon"), np.ones((2, 2), "i4"), {}), "elev": (("lat", "lon"), np.full((2, 2), 100.0), {}),
        "t_pk": (("month", "lat", "lon"), rng.uniform(300, 1140, (12, 2, 2)), {}),
        "dur": (("month", "lat", "lon"), rng.uniform(120, 720, (12, 2, 2)), {})})
    g = driver.MetSimB200(driver.read_config(conf)).gather()
    want = case["frc"]["Tmax"][1][3:13][:, [2, 0]][:, :, [3, 1]].reshape(10, 4).astype(np.float64)
    np.testing.assert_array_equal(g["t_max"], want)
    # a domain coordinate the forcing does not have is a KeyError, as with xarray's .sel
    write_nc3(os.path.join(str(tmp_path), "domain.nc"), {"lat": 1, "lon": 1, "month": 12}, {
        "lat": (("lat",), [12.345], {}), "lon": (("lon",), case["lon"][:1], {}),
        "mask": (("lat", "lon"), np.ones((1, 1), "i4"), {}), "elev": (("lat", "lon"), np.zeros((1, 1)), {}),
        "t_pk": (("month", "lat", "lon"), np.zeros((12, 1, 1)), {}), "dur": (("month", "lat", "lon"), np.zeros((12, 1, 1)), {})})
    with pytest.raises(KeyError, match="12.345"):
        driver.MetSimB200(driver.read_config(conf))


def test_checks_raise_what_the_reference_raises(tmp_path):
    conf, _ = small_case(str(tmp_path))
    p = driver.read_config(conf)
    ms = driver.MetSimB200(dict(p, time_step="50", lw_type="nope", out_vars={"pet": {}}))
    with pytest.raises(Exception) as exc:
        ms._validate_setup()
    msg = str(exc.value)
    assert "Time step must be evenly divisible into 1440 minutes" in msg and "Got 50." in msg
    assert "Cannot output variable pet at timestep 50" in msg and "Invalid option given for lw_type" in msg
    with pytest.raises(AssertionError, match="stop"):               # stop beyond the forcing record
        driver.MetSimB200(dict(p, stop="2001/3/1"))
    with pytest.raises(AssertionError, match="90 days"):            # the state does not cover start-90d
        driver.MetSimB200(dict(p, start="2001/1/27")).state
    # t_max < t_min: the ValueError of metsim.py:612-614
    ms = driver.MetSimB200(p)
    ms.met_data["t_max"].values[2, 1, 1] = -99.0
    with pytest.raises(ValueError, match="Daily maximum temperature lower than daily minimum temperature!"):
        ms.gather()
    # unknown units fall back to the default with a warning (metsim.py:226-237)
    ms = driver.MetSimB200(dict(p, out_vars={"temp": {"units": "furlongs"}}))
    assert ms.params["out_vars"]["temp"] == {"units": "C", "out_name": "temp"}


def test_noleap_calendar_through_the_driver(tmp_path):
    """A forcing file on the noleap calendar across February of a leap year: the records are labelled
    28 February, 1 March (no 29th), the day of year is the label's (59, 61), the state window is the
    90 records before `start`, and the output time axis is written in the run's own calendar."""
    tmp = str(tmp_path)
    conf, case = small_case(tmp, start="2004-02-25", days=10)
    rng = np.random.default_rng(9)
    ny, nx = 3, 4
    coords = {"lat": (("lat",), case["lat"], {}), "lon": (("lon",), case["lon"], {})}
    # noleap day numbers: 2004-02-25 is day 4 * 365 + 55 since 2000-01-01
    d0 = 4 * 365 + 55
    tu = {"units": "days since 2000-01-01 00:00:00", "calendar": "noleap"}
    tmin = rng.normal(0, 5, (16, ny, nx))
    write_nc3(os.path.join(tmp, "forcing.nc"), {"time": 16, "lat": ny, "lon": nx}, dict(
        coords, time=(("time",), (d0 - 3 + np.arange(16)).astype("f8"), tu),
        Tmin=(("time", "lat", "lon"), tmin, {}), Tmax=(("time", "lat", "lon"), tmin + 5.0, {}),
        Prec=(("time", "lat", "lon"), rng.gamma(0.8, 6, tmin.shape), {}),
        wind=(("time", "lat", "lon"), rng.uniform(0.5, 8, tmin.shape), {})))
    smin = rng.normal(0, 5, (90, ny, nx))
    write_nc3(os.path.join(tmp, "state.nc"), {"time": 90, "lat": ny, "lon": nx}, dict(
        coords, time=(("time",), (d0 - 90 + np.arange(90)).astype("f8"), tu),
        t_min=(("time", "lat", "lon"), smin, {}), t_max=(("time", "lat", "lon"), smin + 8.0, {}),
        prec=(("time", "lat", "lon"), rng.gamma(0.8, 6, smin.shape), {})))
    p = driver.read_config(conf)
    p["stop"] = "2004/3/6"                                   # ten noleap days from 25 February
    ms = driver.MetSimB200(p)
    assert ms.params["calendar"] == "noleap"                 # taken from the forcing file (metsim.py:252-254)
    days = ms.met_data["time"].values.astype("datetime64[D]")
    assert days.size == 10 and str(days[3]) == "2004-02-28" and str(days[4]) == "2004-03-01"
    g = ms.gather()
    assert g["doy"].tolist() == [56, 57, 58, 59, 61, 62, 63, 64, 65, 66] and g["month"].tolist() == [2] * 4 + [3] * 6
    assert g["state_t_min"].shape == (90, 12)
    np.testing.assert_array_equal(g["t_min"], tmin[3:13].reshape(10, -1))
    # the time variable of the output: minutes since 2000-01-01 in the noleap calendar
    from metsim_b200 import engine
    import pytest as _pt
    mp = _pt.MonkeyPatch()
    mp.setattr(engine, "run_host", _fake_run_host([]))
    try:
        (path,) = ms.run()
    finally:
        mp.undo()
    from scipy.io import netcdf_file as _nc
    with _nc(path, "r", mmap=False) as f:
        tv = np.array(f.variables["time"].data)
        assert f.variables["time"].calendar == b"noleap"
    assert tv[0] == d0 * 1440 and tv[4 * 4] == (d0 + 4) * 1440 and tv[1] - tv[0] == 360
    (ds,) = ms.open_output()
    assert ds["time"].values[4 * 4] == np.datetime64("2004-03-01T00:00")
    # VIC forcings have no time axis of their own: the run's days are the calendar's
    assert [str(x)[:10] for x in driver._date_range(np.datetime64("2004-02-27"), np.datetime64("2004-03-02"), "noleap")] == \
        ["2004-02-27", "2004-02-28", "2004-03-01", "2004-03-02"]
    assert driver._date_range(np.datetime64("2004-02-27"), np.datetime64("2004-03-02"), "standard").size == 5


def test_vic_ascii_and_binary_cell_files(tmp_path):
    tmp = str(tmp_path)
    conf, case = small_case(tmp, ny=2, nx=2, mask=[[1, 1], [0, 1]])
    rng = np.random.default_rng(11)
    days = 10
    os.makedirs(os.path.join(tmp, "asc")), os.makedirs(os.path.join(tmp, "bin"))
    cols = {}
    for i, la in enumerate(case["lat"]):
        for j, lo in enumerate(case["lon"]):
            if (i, j) == (0, 1):
                continue                                   # an active cell without a file stays NaN
            prec = np.round(rng.gamma(0.8, 6, days + 5) * 40) / 40
            tmax, tmin, wind = (np.round(rng.normal(5, 8, days + 5) * 100) / 100 for _ in range(3))
            cols[(i, j)] = (prec, tmax, tmin, wind)
            np.savetxt(os.path.join(tmp, "asc", f"data_{la}_{lo}"), np.stack([prec, tmax, tmin, wind], 1), fmt="%.6f")
            rec = np.zeros(days + 5, dtype=[("p", "<u2"), ("a", "<i2"), ("b", "<i2"), ("c", "<i2")])
            rec["p"], rec["a"], rec["b"], rec["c"] = np.round(prec * 40), np.round(tmax * 100), np.round(tmin * 100), np.round(wind * 100)
            rec.tofile(os.path.join(tmp, "bin", f"data_{la}_{lo}"))
    open(os.path.join(tmp, "asc", "README"), "w").write("not a forcing file\n")          # skipped
    p = driver.read_config(conf)
    for fmt, fv in (("ascii", {"prec": "prec", "t_max": "t_max", "t_min": "t_min", "wind": "wind"}),
                    ("binary", {"prec": "40.0 unsigned", "t_max": "100.0 signed", "t_min": "100.0 signed",
                                "wind": "100.0 signed"})):
        sub = "asc" if fmt == "ascii" else "bin"
        q = dict(p, forcing_fmt=fmt, forcing_vars=fv, forcing=driver._forcing_files(os.path.join(tmp, sub)))
        md = driver.MetSimB200(q).met_data
        assert md["t_min"].values.shape == (days, 2, 2)
        for (i, j), (prec, tmax, tmin, wind) in cols.items():
            if (i, j) == (1, 0):                           # masked out: never read
                assert np.isnan(md["prec"].values[:, i, j]).all()
                continue
            np.testing.assert_allclose(md["prec"].values[:, i, j], prec[:days], rtol=0, atol=1e-12)
            np.testing.assert_allclose(md["t_max"].values[:, i, j], tmax[:days], rtol=0, atol=1e-12)
            np.testing.assert_allclose(md["wind"].values[:, i, j], wind[:days], rtol=0, atol=1e-12)
        assert np.isnan(md["t_min"].values[:, 0, 1]).all()


def test_decode_time_and_cf_masking():
    t = ncio.decode_time([0, 1.5, 366], "days since 1900-01-01 00:00:00")
    assert t.tolist() == np.array(["1900-01-01T00", "1900-01-02T12", "1901-01-02T00"], "datetime64[s]").tolist()
    assert ncio.decode_time([90], "minutes since 2000-01-01 00:00:00.0")[0] == np.datetime64("2000-01-01T01:30")
    assert ncio.decode_time([2], "hours since 1990-1-1 6:30", "gregorian")[0] == np.datetime64("1990-01-01T08:30")
    with pytest.raises(NotImplementedError, match="julian"):
        ncio.decode_time([0], "days since 2000-01-01", "julian")
    # uniform-year calendars: computed in the calendar, labelled with the Gregorian date of the same
    # year / month / day (CFTimeIndex.to_datetimeindex() in the reference)
    t = ncio.decode_time([0, 58, 59, 365 + 59], "days since 2000-01-01", "noleap")
    assert t.tolist() == np.array(["2000-01-01", "2000-02-28", "2000-03-01", "2001-03-01"], "datetime64[s]").tolist()
    np.testing.assert_array_equal(ncio.encode_time(t, "days since 2000-01-01", "noleap"), [0, 58, 59, 424])
    assert ncio.encode_time(np.datetime64("2001-03-01"), driver.DEFAULT_TIME_UNITS, "noleap") == (365 + 59) * 1440
    assert ncio.decode_time([59], "days since 2004-01-01 06:00", "all_leap")[0] == np.datetime64("2004-02-29T06:00")
    assert ncio.decode_time([30, 360], "days since 2001-01-01", "360_day").tolist() == \
        np.array(["2001-02-01", "2002-01-01"], "datetime64[s]").tolist()
    with pytest.raises(ValueError, match="no Gregorian label"):           # 30 February
        ncio.decode_time([59], "days since 2000-01-01", "360_day")
    with pytest.raises(ValueError, match="does not exist in the noleap calendar"):
        ncio.encode_time(np.datetime64("2000-02-29"), "days since 2000-01-01", "noleap")
    std = ncio.decode_time([1.5], "days since 1900-01-01")
    assert ncio.encode_time(std, "days since 1900-01-01")[0] == 1.5
    with pytest.raises(ValueError):
        ncio.decode_time([0], "fortnights since 2000-01-01")
    a = ncio._decode(np.array([1.0, 1e20, 3.0], "f4"), {"_FillValue": np.float32(1e20)})
    assert np.isnan(a[1]) and a[0] == 1.0 and a.dtype == np.float32
    b = ncio._decode(np.array([1, -9999, 3], "i2"), {"missing_value": np.int16(-9999), "scale_factor": 0.5, "add_offset": 1.0})
    assert np.isnan(b[1]) and b[0] == 1.5 and b[2] == 2.5
    c = np.array([1, 2], "i4")
    assert ncio._decode(c, {}) is c                                      # nothing to decode: untouched


def _fake_run_host(calls):
    """Stands in for metsim_b200.engine.run_host: fills every requested output with
    value = 1000 * variable index + step + cell / 1000 and drives the on_tile hook like the C side."""
    from metsim_b200._lib import SUB_VARS

    def fake(params, *, out_vars=(), out_dtype=np.float32, units=None, daily_vars=(), out=None, ring=False,
             tile_days=0, on_tile=None, timings=None, device=0, **arrays):
        ts = int(params["time_step"])
        d, n = arrays["t_min"].shape
        calls.append(dict(ring=ring, tile_days=tile_days, units=units, out_vars=tuple(out_vars), n=n, d=d))
        cell = np.arange(n)[None, :] / 1000.0
        if ts >= 1440:
            return {"daily": {v: 10.0 * (1 + i) + np.arange(d)[:, None] + cell for i, v in enumerate(daily_vars)}}
        tpd = 1440 // ts
        pat = lambda v, s0, s1: (1000.0 * SUB_VARS.index(v) + np.arange(s0, s1)[:, None] + cell).astype(out_dtype)
        if not ring:
            return {v: pat(v, 0, d * tpd) for v in out_vars}
        for tile, day0 in enumerate(range(0, d, tile_days)):
            day1, slot = min(day0 + tile_days, d), tile % 2
            for v in out_vars:
                out[v][slot * tile_days * tpd:slot * tile_days * tpd + (day1 - day0) * tpd] = pat(v, day0 * tpd, day1 * tpd)
            on_tile(tile, day0, day1, slot)
        return {}
    return fake


@pytest.mark.parametrize("budget", [4 << 30, 1500])
def test_run_writes_reference_shaped_files(tmp_path, monkeypatch, budget):
    """out_freq splits the record into one file per month; a small host budget forces the 2-slot ring
    whose tiles straddle the file boundary; masked cells stay NaN; names, units and the time axis follow
    setup_netcdf_output (metsim/metsim.py:363-436)."""
    from metsim_b200 import engine
    from metsim_b200._lib import SUB_VARS
    mask = [[1, 0, 1, 1], [0, 0, 1, 0], [1, 1, 0, 1]]
    conf, case = small_case(str(tmp_path), mask=mask, out_freq="1M",
                            extra="out_precision = f8" if budget < 4096 else "")
    calls = []
    monkeypatch.setattr(engine, "run_host", _fake_run_host(calls))
    p = driver.read_config(conf)
    p["out_vars"]["prec"]["units"] = "mm h-1"
    p["out_vars"]["temp"]["out_name"] = "airtemp"
    ms = driver.MetSimB200(p)
    files = ms.run(host_budget_bytes=budget)
    assert [os.path.basename(f) for f in files] == ["small_20010125-20010131.nc", "small_20010201-20010203.nc"]
    assert calls[0]["ring"] == (budget < 4096) and calls[0]["units"]["prec"] == "mm h-1" and calls[0]["n"] == 7
    if budget < 4096:
        assert 1 <= calls[0]["tile_days"] < 10
    sel = np.nonzero(np.asarray(mask).ravel())[0]
    step0 = 0
    for ds, nsteps in zip(ms.open_output(), (28, 12)):
        assert ds.dims == {"time": nsteps, "lat": 3, "lon": 4}
        assert set(ds) == {"time", "lat", "lon", "airtemp", "prec", "shortwave", "wind"}
        t = ds["time"]
        assert t.attrs["units"] == driver.DEFAULT_TIME_UNITS and t.attrs["standard_name"] == "utc_time"
        assert t.attrs["calendar"] == "standard"
        assert t.values[0] == np.datetime64("2001-01-25T00:00") + step0 * np.timedelta64(6, "h")
        np.testing.assert_array_equal(ds["lat"].values, case["lat"])
        assert ds["prec"].attrs["units"] == "mm h-1" and ds["airtemp"].attrs["long_name"] == "air temperature"
        for name, v in (("airtemp", "temp"), ("prec", "prec"), ("shortwave", "shortwave"), ("wind", "wind")):
            vals = ds[name].values
            assert vals.dtype == (np.float64 if budget < 4096 else np.float32) and vals.shape == (nsteps, 3, 4)
            flat = vals.reshape(nsteps, -1)
            assert np.isnan(np.delete(flat, sel, axis=1)).all()
            want = (1000.0 * SUB_VARS.index(v) + np.arange(step0, step0 + nsteps)[:, None] + np.arange(7)[None, :] / 1000.0)
            np.testing.assert_array_equal(flat[:, sel], want.astype(vals.dtype))
        step0 += nsteps


class _NoCtx:
    """Stands in for engine.HostContext where there is no CUDA device."""
    def __init__(self, device=0):
        self.device = device

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _fake_run_host_tagged(calls, fail_on=None):
    """Like _fake_run_host, but a value carries its cell's elevation (unique per cell) instead of its
    position in the call, so that ranges of cells computed by separate calls can be told apart:
    value = 1000 * variable index + step + elev / 1e4."""
    import threading
    from metsim_b200._lib import SUB_VARS
    lock = threading.Lock()

    def fake(params, *, out_vars=(), out_dtype=np.float64, units=None, daily_vars=(), out=None, ring=False,
             tile_days=0, on_tile=None, timings=None, device=0, ctx=None, **arrays):
        ts = int(params["time_step"])
        d, n = arrays["t_min"].shape
        with lock:
            calls.append(dict(n=n, device=getattr(ctx, "device", device), ring=ring, tile_days=tile_days,
                              thread=threading.current_thread().name))
        tag = np.asarray(arrays["elev"])[None, :] / 1e4
        if ts >= 1440:
            return {"daily": {v: 10.0 * (1 + i) + np.arange(d)[:, None] + tag for i, v in enumerate(daily_vars)}}
        tpd = 1440 // ts
        pat = lambda v, s0, s1: (1000.0 * SUB_VARS.index(v) + np.arange(s0, s1)[:, None] + tag).astype(out_dtype)
        if not ring:
            return {v: pat(v, 0, d * tpd) for v in out_vars}
        for tile, day0 in enumerate(range(0, d, tile_days)):
            if fail_on is not None and tile == 1 and n == fail_on:
                raise RuntimeError("range failed")
            day1, slot = min(day0 + tile_days, d), tile % 2
            for v in out_vars:
                out[v][slot * tile_days * tpd:slot * tile_days * tpd + (day1 - day0) * tpd] = pat(v, day0 * tpd, day1 * tpd)
            on_tile(tile, day0, day1, slot)
        return {}
    return fake


@pytest.mark.parametrize("budget", [4 << 30, 40000])
def test_run_over_several_devices_gathers_cell_ranges(tmp_path, monkeypatch, budget):
    """``run(devices=[...])``: contiguous cell ranges on their own threads / contexts, gathered on the host
    into full rows -- whole columns when the record fits the budget, tile by tile through the barrier
    otherwise; the files equal what one range would have written."""
    from metsim_b200 import engine
    from metsim_b200._lib import SUB_VARS
    rng = np.random.default_rng(5)
    mask = (rng.random((10, 13)) < 0.9).astype("i4")
    conf, case = small_case(str(tmp_path), ny=10, nx=13, mask=mask, out_freq="1M", extra="out_precision = f8")
    calls = []
    monkeypatch.setattr(engine, "run_host", _fake_run_host_tagged(calls))
    monkeypatch.setattr(engine, "HostContext", _NoCtx)
    ms = driver.MetSimB200(driver.read_config(conf))
    files = ms.run(host_budget_bytes=budget, devices=[0, 1, 1])
    n = int(mask.sum())
    from metsim_b200.shard import cell_range
    widths = [hi - lo for lo, hi in (cell_range(n, k, 3) for k in range(3))]
    assert sorted(c["n"] for c in calls) == sorted(widths) and sum(widths) == n and widths[0] % 32 == 0
    assert sorted(c["device"] for c in calls) == [0, 1, 1] and len({c["thread"] for c in calls}) == 3
    assert all(c["ring"] == (budget < 1 << 20) for c in calls)
    assert [r["cells"] for r in ms.timings["ranges"]] == widths and [r["device"] for r in ms.timings["ranges"]] == [0, 1, 1]
    elev = ms.gather()["elev"]
    sel = np.flatnonzero(mask.ravel())
    step0 = 0
    for ds, nsteps in zip(ms.open_output(), (28, 12)):
        for v in ("temp", "prec", "shortwave", "wind"):
            flat = ds[v].values.reshape(nsteps, -1)
            want = 1000.0 * SUB_VARS.index(v) + np.arange(step0, step0 + nsteps)[:, None] + elev[None, :] / 1e4
            np.testing.assert_array_equal(flat[:, sel], want)
            assert np.isnan(np.delete(flat, sel, axis=1)).all()
        step0 += nsteps
    # too few cells for the devices offered: fewer ranges, same files
    calls.clear()
    conf2, _ = small_case(str(tmp_path), ny=5, nx=8)
    driver.MetSimB200(driver.read_config(conf2)).run(devices=[0, 1, 2, 3])
    assert len(calls) == 1 and calls[0]["n"] == 40


def test_run_over_several_devices_daily_mode_and_a_failing_range(tmp_path, monkeypatch):
    from metsim_b200 import engine
    conf, _ = small_case(str(tmp_path), ny=10, nx=13, ts=1440)
    calls = []
    monkeypatch.setattr(engine, "run_host", _fake_run_host_tagged(calls))
    monkeypatch.setattr(engine, "HostContext", _NoCtx)
    p = driver.read_config(conf)
    av = driver.available_outputs()
    p["out_vars"] = {"prec": dict(av["prec"]), "pet": dict(av["pet"])}
    ms = driver.MetSimB200(dict(p, out_precision="f8"))
    ms.run(devices=[0, 1])
    (ds,) = ms.open_output()
    g = ms.gather()
    np.testing.assert_array_equal(ds["prec"].values.reshape(10, -1), g["prec"])
    np.testing.assert_array_equal(ds["pet"].values.reshape(10, -1), 40.0 + np.arange(10)[:, None] + g["elev"][None, :] / 1e4)
    assert len(calls) == 2
    # a range that fails mid-run: the others are released from the barrier and the first real error is raised
    conf, _ = small_case(str(tmp_path), ny=10, nx=13, ts=360)
    monkeypatch.setattr(engine, "run_host", _fake_run_host_tagged([], fail_on=64))
    ms = driver.MetSimB200(dict(driver.read_config(conf), out_precision="f8"))
    with pytest.raises(RuntimeError, match="range failed"):
        ms.run(host_budget_bytes=40000, devices=[0, 1])


def test_run_daily_mode_writes_forcings_and_daily_columns(tmp_path, monkeypatch):
    """time_step = 1440 (metsim/metsim.py:779-792): t_min / prec come from the forcings, shortwave is the
    daily mean flux column, the rest the daily MTCLIM columns; units converted on the host."""
    from metsim_b200 import engine
    conf, case = small_case(str(tmp_path), ts=1440)
    calls = []
    monkeypatch.setattr(engine, "run_host", _fake_run_host(calls))
    p = driver.read_config(conf)
    av = driver.available_outputs()
    p["out_vars"] = {"t_min": dict(av["t_min"], units="K"), "prec": dict(av["prec"]),
                     "shortwave": dict(av["shortwave"]), "pet": dict(av["pet"])}
    ms = driver.MetSimB200(p)
    (path,) = ms.run()
    (ds,) = ms.open_output()
    g = ms.gather()
    np.testing.assert_array_equal(ds["t_min"].values.reshape(10, -1), (g["t_min"] + 273.15).astype("f4"))
    np.testing.assert_array_equal(ds["prec"].values.reshape(10, -1), g["prec"].astype("f4"))
    order = ("t_day", "vapor_pressure", "tskc", "pet", "daylength", "shortwave_daily")
    cell = np.arange(12)[None, :] / 1000.0
    for name, src in (("shortwave", "shortwave_daily"), ("pet", "pet")):
        want = 10.0 * (1 + order.index(src)) + np.arange(10)[:, None] + cell
        np.testing.assert_array_equal(ds[name].values.reshape(10, -1), want.astype("f4"))
    assert ds["time"].values[1] - ds["time"].values[0] == np.timedelta64(1, "D")
    assert ds["t_min"].attrs["units"] == "K" and ds["time"].attrs["long_name"] == "UTC time"


def test_command_line_needs_an_existing_file(tmp_path, capsys):
    with pytest.raises(SystemExit) as exc:
        driver.main([os.path.join(str(tmp_path), "missing.conf")])
    assert exc.value.code == 1 and "does not exist" in capsys.readouterr().err


# ---------------------------------------------------------------- against the reference's own files
@needs_reference
def test_example_nc_conf_reads_the_golden_inputs_bit_for_bit(monkeypatch):
    """examples/example_nc.conf as the reference ships it (test.nc + domain.nc + the NetCDF-4 state):
    what the driver hands to the GPU equals the inputs stored in tests/golden/testdomain.npz."""
    monkeypatch.chdir(REF)
    ms = driver.MetSimB200(driver.read_config("examples/example_nc.conf"))
    ms._validate_setup()
    g = ms.gather()
    inputs = util.load_golden(os.path.join(util.GOLDEN, "testdomain.npz"))[0]
    for k in ("lat", "lon", "elev", "doy", "month", "t_min", "t_max", "prec", "wind", "state_t_min",
              "state_t_max", "state_prec", "t_pk12", "dur12"):
        np.testing.assert_array_equal(g[k], inputs[k], err_msg=k)
    assert ms.params["time_step"] == "30" and ms._times[0].size == 1488


@needs_reference
def test_example_bin_conf_reads_the_known_answer_cell(monkeypatch):
    """examples/example_bin.conf: VIC-4 binary forcings + stehekin.nc + the densely-linked NetCDF-4
    state; the cell of the reference's known-answer test equals the stehekin fixture's inputs."""
    monkeypatch.chdir(REF)
    ms = driver.MetSimB200(driver.read_config("examples/example_bin.conf"))
    g = ms.gather()
    inputs = util.load_golden(os.path.join(util.GOLDEN, "stehekin_hourly.npz"))[0]
    i = int(np.flatnonzero((g["lat"] == 48.3125) & (g["lon"] == -120.5625))[0])
    assert g["t_min"].shape == (1826, 16)
    for k in ("t_min", "t_max", "prec", "wind"):
        np.testing.assert_array_equal(g[k][:365, i], inputs[k][:, 0], err_msg=k)
    for k in ("state_t_min", "state_t_max", "state_prec"):
        np.testing.assert_array_equal(g[k][:, i], inputs[k][:, 0], err_msg=k)


@needs_reference
@pytest.mark.parametrize("conf,cells,days,nfiles", [
    ("example_ascii.conf", 16, 31, 1), ("example_yaml.yaml", 1, 31, 1), ("example_constant_vars_nc.conf", 81, 31, 1),
    ("example_constant_vars_ascii.conf", 16, 31, 1), ("example_constant_vars_bin.conf", 16, 1826, 1),
    ("example_dimtest.conf", 9, 90, 1)])
def test_every_example_configuration_loads(monkeypatch, conf, cells, days, nfiles):
    monkeypatch.chdir(REF)
    ms = driver.MetSimB200(driver.read_config(os.path.join("examples", conf)))
    ms._validate_setup()
    g = ms.gather()
    assert g["t_min"].shape == (days, cells) and len(ms._times) == nfiles
    assert np.isfinite(g["t_min"]).all() and np.isfinite(g["state_prec"]).all() and np.isfinite(g["elev"]).all()
    if "constant_vars" in conf:
        assert (g["wind"] == 2.0).all()
    if conf == "example_dimtest.conf":                       # a 1-D ('hru',) domain, NetCDF-4 with big-endian scales
        assert ms.domain["mask"].dims == ("hru",) and g["cell_index"].tolist() == list(range(9))
        # test_metsim.py:467-483 (test_coordinate_dimension_matchup): the file has no `hru` coordinate variable,
        # the run gets a sequential one
        assert "hru" not in ms.domain and driver._coordinate(ms.domain, "hru").tolist() == list(range(9))


@needs_reference
def test_hdf5_reader_recovers_dimensions_and_attributes():
    from metsim_b200.hdf5_min import read_hdf5_vars
    v, dims, _ = read_hdf5_vars(os.path.join(REF, "metsim", "data", "state_vic.nc"))     # dense links + dense attributes
    assert dims == {"time": 90, "lat": 4, "lon": 5} and v["t_min"].dims == ("time", "lat", "lon")
    assert v["time"].attrs["units"] == "days since 1948-10-03 00:00:00" and v["lon"].attrs["units"] == "degrees_east"
    v, dims, _ = read_hdf5_vars(os.path.join(REF, "metsim", "data", "dim_test.nc"))
    assert dims == {"time": 180, "hru": 9} and "hru" not in v and v["pptrate"].dims == ("time", "hru")
    v, _, _ = read_hdf5_vars(os.path.join(REF, "metsim", "data", "tiny_domain.nc"))
    assert v["t_pk"].dims == ("month", "lat", "lon") and v["t_pk"].values.dtype == np.float32


def _same_as_scipy(path):
    """ncio's own parser of the classic format against scipy's, variable by variable."""
    ds = ncio.open_dataset(path)
    with netcdf_file(path, "r", mmap=False, maskandscale=False) as f:
        assert set(f.variables) == set(ds)
        for k, v in f.variables.items():
            a, b = np.array(v.data), ds[k].values
            assert tuple(v.dimensions) == ds[k].dims
            for kk, vv in v._attributes.items():
                w = ds[k].attrs[kk]
                if isinstance(vv, bytes):
                    assert vv.decode() == w
                else:
                    assert np.array_equal(np.atleast_1d(vv), np.atleast_1d(w), equal_nan=True)
            if b.dtype.kind == "M":
                continue
            if a.dtype.kind == "S":
                assert a.tobytes() == b.tobytes()
                continue
            keep = ~np.isnan(b) if b.dtype.kind == "f" else np.ones(b.shape, bool)     # decoded fill values
            assert np.array_equal(a.astype(a.dtype.newbyteorder("="))[keep], b[keep]), k
    return ds


@pytest.mark.parametrize("version", [1, 2])
def test_classic_reader_record_variables_against_scipy(tmp_path, version):
    """Fixed-size and record variables (interleaved record by record; a lone 2-byte record variable is
    not padded), text, attributes, classic and 64-bit offset."""
    rng = np.random.default_rng(1)
    p = os.path.join(str(tmp_path), "rec.nc")
    with netcdf_file(p, "w", version=version) as f:
        f.createDimension("time", None), f.createDimension("x", 3), f.createDimension("s", 5)
        t = f.createVariable("time", "i4", ("time",))
        a = f.createVariable("a", "f8", ("time", "x"))
        b = f.createVariable("b", "i2", ("time", "x"))
        c = f.createVariable("c", "f4", ("x",))
        s = f.createVariable("name", "c", ("s",))
        t[:], a[:], b[:], c[:] = np.arange(7), rng.random((7, 3)), np.arange(21).reshape(7, 3), [1.5, 2.5, 3.5]
        s[:] = np.array(list("hello"), "S1")
        t.units, f.title, a.gain = "days since 2000-01-01", "record test", np.float32(2.0)
    ds = _same_as_scipy(p)
    assert ds.dims == {"time": 7, "x": 3, "s": 5} and ds.attrs["title"] == "record test"
    assert ds["time"].values[6] == np.datetime64("2000-01-07")
    p = os.path.join(str(tmp_path), "single.nc")
    with netcdf_file(p, "w", version=version) as f:
        f.createDimension("time", None), f.createDimension("x", 3)
        f.createVariable("b", "i2", ("time", "x"))[:] = np.arange(15).reshape(5, 3)
    assert _same_as_scipy(p)["b"].values.tolist() == np.arange(15).reshape(5, 3).tolist()


def test_classic_reader_reads_what_the_writer_writes_including_cdf5(tmp_path):
    """The driver must be able to read its own outputs back (open_output, the passthrough workflow) at
    any size: ncwriter switches to CDF-5 beyond 4 GiB per variable, which scipy cannot read."""
    from metsim_b200.ncwriter import NC3Writer
    rng = np.random.default_rng(2)
    for version in (2, 5):
        p = os.path.join(str(tmp_path), f"w{version}.nc")
        blk = rng.random((6, 5)).astype("f4")
        with NC3Writer(p, n_steps=6, lat=[1.0, 2.0], lon=[3.0, 4.0, 5.0], cell_index=[0, 1, 2, 4, 5],
                       variables={"temp": {"units": "C", "dtype": "float32"}}, time_values=np.arange(6, dtype=np.int32) * 30,
                       time_units=driver.DEFAULT_TIME_UNITS, version=version, native=False) as w:
            w.write_steps("temp", 0, blk)
        assert open(p, "rb").read(4) == b"CDF" + bytes([version])
        ds = ncio.open_dataset(p)
        g = ds["temp"].values.reshape(6, 6)
        assert np.array_equal(g[:, [0, 1, 2, 4, 5]], blk) and np.isnan(g[:, 3]).all()
        assert ds["temp"].attrs["units"] == "C" and ds["time"].values[1] == np.datetime64("2000-01-01T00:30")
        if version == 2:
            _same_as_scipy(p)


@needs_reference
@pytest.mark.parametrize("name", ["test.nc", "domain.nc", "stehekin.nc"])
def test_classic_reader_on_the_reference_files(name):
    _same_as_scipy(os.path.join(REF, "metsim", "data", name))
