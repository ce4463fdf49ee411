"""The oracle (oracle/metsim_oracle.c) against reference-generated golden vectors and against the
reference repository's own known-answer tables.  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle_lib
import msb_testutil as util


@pytest.mark.parametrize("path", util.golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_reference_outputs(path):
    inputs, params, ref_daily, ref_sub, extra = util.load_golden(path)
    got = oracle_lib.run(params, **util.oracle_kwargs(inputs))
    cells = extra.get("ref_cells")
    pick = (lambda a: a[:, cells]) if cells is not None else (lambda a: a)
    for oname, gname in util.DAILY.items():
        util.assert_close(pick(got[oname]), ref_daily[gname], oname, context=os.path.basename(path))
    for v, want in ref_sub.items():
        util.assert_close(pick(got[v]), want, v, context=os.path.basename(path))


@pytest.mark.parametrize("tag,tol", [("hourly", 0.03), ("3hourly", 0.2)])
def test_oracle_vs_reference_repo_known_answer(tag, tol):
    """metsim/tests/test_metsim.py:397-464 (test_disaggregation_values): NRMSE against the repo's
    validated tables, with the reference's own thresholds (0.03 hourly, 0.2 three-hourly).  The
    90-day state is a declared synthetic one (state_vic.nc is HDF5, unreadable here)."""
    inputs, params, _, _, extra = util.load_golden(os.path.join(util.GOLDEN, f"stehekin_{tag}.npz"))
    got = oracle_lib.run(params, **util.oracle_kwargs(inputs))
    good = extra["repo_golden"].astype(np.float64)
    for j, var in enumerate(extra["repo_golden_columns"]):
        out = got[str(var)][:, 0]
        h = max(good[:, j].max(), out.max())
        l = min(good[:, j].min(), out.min())
        nrmse = np.sqrt(np.mean((good[:, j] - out) ** 2)) / (h - l)
        assert nrmse < tol, (var, nrmse)


def test_oracle_solar_geom_table_properties():
    trf, dayl, potrad, ttmax = oracle_lib.solar_geom(500.0, 45.0)
    assert trf.shape == (366, 2880)
    np.testing.assert_array_equal(trf[365], trf[364])          # physics.py:281-284
    sums = trf[:365].sum(axis=1)
    np.testing.assert_allclose(sums[dayl[:365] > 0], 1.0, rtol=1e-12)
    assert (dayl > 0).all() and (dayl <= 86400).all()
    assert (ttmax > 0.3).all() and (ttmax < 0.9).all()
    # polar night: no daylight rows are all zero, daylength 0
    trf, dayl, potrad, ttmax = oracle_lib.solar_geom(0.0, 85.0)
    assert (dayl == 0).any() and (trf[dayl == 0] == 0).all()


def test_oracle_rejects_negative_dtr():
    inputs, params, *_ = util.load_golden(os.path.join(util.GOLDEN, "synth_daily.npz"))
    bad = dict(inputs)
    bad["t_max"] = inputs["t_max"].copy()
    bad["t_max"][5, 2] = bad["t_min"][5, 2] - 1.0
    with pytest.raises(ValueError):
        oracle_lib.run(params, **util.oracle_kwargs(bad))


def test_testdomain_golden_carries_the_bundled_state():
    """BASELINE config #1 literally: the 90-day state of metsim/data/state_nc.nc (constant 10 / 20
    degC, 3 mm/day) is what the fixture was generated with -- and, where the reference checkout is
    present, what oracle/mini_hdf5.py reads from the file itself."""
    inputs, params, *_ = util.load_golden(os.path.join(util.GOLDEN, "testdomain.npz"))
    assert inputs["state_t_min"].shape == (90, 81)
    assert (inputs["state_t_min"] == 10.0).all() and (inputs["state_t_max"] == 20.0).all()
    assert (inputs["state_prec"] == 3.0).all()
    assert int(params["time_step"]) == 30 and str(params["prec_type"]) == "triangle"
    path = "/root/reference/metsim/data/state_nc.nc"
    if os.path.exists(path):
        from oracle.mini_hdf5 import read_hdf5
        st = read_hdf5(path)
        assert set(st) >= {"time", "lat", "lon", "t_min", "t_max", "prec"}
        np.testing.assert_array_equal(st["t_min"].reshape(90, 81), inputs["state_t_min"])
        np.testing.assert_array_equal(st["prec"].reshape(90, 81), inputs["state_prec"])
        np.testing.assert_allclose(st["lat"], np.unique(inputs["lat"]))


def test_mini_hdf5_reads_the_dense_chunked_state_file_of_the_known_answer_test():
    """state_vic.nc (root-group links in a fractal heap + B-tree, chunked datasets) is the state the
    reference's known-answer test configures (metsim/tests/test_metsim.py:408).  The Stehekin fixtures
    carry what the reader extracted; where the reference checkout is present (the build container)
    the file is read again and must agree; elsewhere the fixture's own consistency is checked."""
    import os
    import numpy as np
    import msb_testutil as util
    inputs, params, _, _, _ = util.load_golden(os.path.join(util.GOLDEN, "stehekin_hourly.npz"))
    st = inputs["state_t_min"][:, 0]
    assert st.shape == (90,) and np.isfinite(st).all() and np.abs(st).max() < 100
    # a state that is NOT the tail of the forcing record (round 1 used that as a declared stand-in)
    assert not np.array_equal(st, inputs["t_min"][-90:, 0])
    path = "/root/reference/metsim/data/state_vic.nc"
    if not os.path.exists(path):
        return
    from oracle.mini_hdf5 import read_hdf5
    f = read_hdf5(path)
    assert sorted(f) == sorted(["t_max", "mask", "t_min", "lon", "frac", "time", "swe", "lat", "prec", "elev", "area"])
    assert f["t_min"].shape == (90, 4, 5) and f["mask"].sum() == 16
    np.testing.assert_array_equal(f["time"], np.arange(90))
    for k in ("t_min", "t_max", "prec"):
        np.testing.assert_array_equal(f[k][:, 1, 4], inputs["state_" + k][:, 0])
    # masked cells hold the netCDF fill value, active ones are finite
    assert (f["t_min"][:, f["mask"] == 0] > 9e36).all() and (np.abs(f["t_min"][:, f["mask"] == 1]) < 100).all()
    p = read_hdf5("/root/reference/metsim/data/passthrough.nc")
    assert p["shortwave"].shape == (365, 4, 5) and p["shortwave"].dtype == np.float32
