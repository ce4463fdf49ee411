"""The streaming NetCDF-3 writer (metsim_b200/ncwriter.py): files are read back with scipy's
independent NetCDF reader.  CPU only -- no compute is involved."""
import os

import numpy as np
from scipy.io import netcdf_file

from metsim_b200.ncwriter import NC3Writer


def test_tiles_roundtrip_with_mask_and_two_dtypes(tmp_path):
    rng = np.random.default_rng(5)
    lat = 30.0 + np.arange(7) / 16.0
    lon = -100.0 + np.arange(9) / 16.0
    mask = rng.random((7, 9)) < 0.6
    idx = np.flatnonzero(mask.ravel())
    n, steps = idx.size, 48
    temp = rng.normal(5, 10, (steps, n)).astype(np.float32)
    prec = rng.gamma(0.8, 2.0, (steps, n))
    path = os.path.join(tmp_path, "out.nc")
    with NC3Writer(path, n_steps=steps, lat=lat, lon=lon, cell_index=idx,
                   variables={"temp": {"units": "C"}, "prec": {"units": "mm timestep-1", "dtype": "float64"}},
                   time_values=np.arange(steps) * 30.0, time_units="minutes since 1950-01-01",
                   global_attrs={"title": "unit test"}) as w:
        for t0, t1 in ((16, 32), (0, 16), (32, 48)):      # tiles in any order
            w.write_steps("temp", t0, temp[t0:t1])
            w.write_steps("prec", t0, prec[t0:t1])
        assert w.steps_written == {"temp": steps, "prec": steps}
    nc = netcdf_file(path, "r", mmap=False)
    assert nc.version_byte == 2
    assert nc.dimensions == {"time": steps, "lat": 7, "lon": 9}
    np.testing.assert_array_equal(nc.variables["lat"].data, lat)
    np.testing.assert_array_equal(nc.variables["lon"].data, lon)
    np.testing.assert_array_equal(nc.variables["time"].data, np.arange(steps) * 30.0)
    assert nc.variables["temp"].units == b"C" and nc.variables["time"].units == b"minutes since 1950-01-01"
    assert nc.title == b"unit test"
    for name, want in (("temp", temp), ("prec", prec)):
        got = nc.variables[name].data.reshape(steps, -1)
        assert got.dtype.itemsize == want.dtype.itemsize
        np.testing.assert_array_equal(got[:, idx], want)                       # bit-exact
        assert np.isnan(np.delete(got, idx, axis=1)).all()                     # NaN outside the mask
    nc.close()


def test_writer_rejects_misfit_blocks(tmp_path):
    import pytest
    w = NC3Writer(os.path.join(tmp_path, "x.nc"), n_steps=4, lat=[0.0, 1.0], lon=[0.0], cell_index=[0, 1],
                  variables={"temp": {}})
    with pytest.raises(ValueError):
        w.write_steps("temp", 2, np.zeros((3, 2), np.float32))   # runs past the end
    with pytest.raises(ValueError):
        w.write_steps("temp", 0, np.zeros((2, 3), np.float32))   # wrong cell count
    w.close()
    with pytest.raises(ValueError):
        NC3Writer(os.path.join(tmp_path, "y.nc"), n_steps=1, lat=[0.0], lon=[0.0], cell_index=[5], variables={})


def test_dense_grid_fast_path(tmp_path):
    lat, lon = np.arange(3.0), np.arange(4.0)
    data = np.arange(5 * 12, dtype=np.float32).reshape(5, 12)
    path = os.path.join(tmp_path, "d.nc")
    with NC3Writer(path, n_steps=5, lat=lat, lon=lon, cell_index=np.arange(12), variables={"wind": {}}) as w:
        assert w._dense
        w.write_steps("wind", 2, data[2:5])
        w.write_steps("wind", 0, data[0:2])
    nc = netcdf_file(path, "r", mmap=False)
    np.testing.assert_array_equal(nc.variables["wind"].data.reshape(5, 12), data)
    nc.close()
