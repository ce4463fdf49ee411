"""The streaming NetCDF-3 writer (metsim_b200/ncwriter.py): files are read back with scipy's
independent NetCDF reader.  CPU only -- no compute is involved."""
import os

import numpy as np
from scipy.io import netcdf_file

from metsim_b200.ncwriter import NC3Writer


import pytest


@pytest.mark.parametrize("native", [True, False], ids=["native-sink", "numpy"])
def test_tiles_roundtrip_with_mask_and_two_dtypes(tmp_path, native):
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
                   global_attrs={"title": "unit test"}, native=native, n_threads=3) as w:
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


def test_native_sink_matches_numpy_scatter_on_many_steps_and_threads(tmp_path):
    """The thread pool of csrc/sink.cu against the plain numpy scatter: thousands of steps (many
    staging chunks per thread), float32 and float64, masked and dense, file bytes identical."""
    from metsim_b200.ncwriter import NativeSink
    rng = np.random.default_rng(9)
    for elem, dense in ((4, False), (8, False), (4, True)):
        n_grid = 5000
        idx = np.arange(n_grid) if dense else np.sort(rng.choice(n_grid, size=3100, replace=False))
        dt = np.dtype("f4" if elem == 4 else "f8")
        steps = 1500
        rows = rng.normal(0, 50, (steps, idx.size)).astype(dt)
        path = os.path.join(tmp_path, f"s{elem}{dense}.bin")
        fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
        head = 123 * 4
        sk = NativeSink(n_cells=idx.size, n_grid=n_grid, elem_bytes=elem,
                        cell_index=None if dense else idx, byteswap=True, n_threads=5, fd=fd)
        sk.write(rows[:700], head)                                     # two calls, like two tiles
        sk.write(rows[700:], head + 700 * n_grid * elem)
        assert sk.bytes_written == steps * n_grid * elem
        assert sk.checksum == int(rows.view(f"u{elem}").astype(np.uint64).sum(dtype=np.uint64))
        sk.close()
        os.close(fd)
        want = np.full((steps, n_grid), np.nan, dtype=dt.newbyteorder(">"))
        want[:, idx] = rows
        got = np.fromfile(path, dtype=dt.newbyteorder(">"), offset=head).reshape(steps, n_grid)
        np.testing.assert_array_equal(got.view(f"u{elem}"), want.view(f"u{elem}"))   # NaN payloads included
    with pytest.raises(Exception):
        NativeSink(n_cells=4, n_grid=3)
    with pytest.raises(Exception):
        NativeSink(n_cells=3, n_grid=5, cell_index=[0, 0, 1])          # repeated grid position
    nofile = NativeSink(n_cells=8, elem_bytes=4, byteswap=False, fd=-1)
    nofile.write(np.ones((10, 8), np.float32))
    assert nofile.bytes_written == 0 and nofile.checksum == 80 * 0x3F800000


def test_large_variables_switch_to_cdf5_and_small_files_stay_cdf2(tmp_path):
    """CDF-2 allows only the last fixed-size variable to exceed 2**32 - 4 bytes.  A file
    that would break that rule is written as CDF-5 (64-bit sizes); the header is parsed back here by
    hand (scipy reads CDF-1/2 only) and a sparse tile written near the end lands where it says."""
    import struct
    lat, lon = np.arange(300.0), np.arange(400.0)
    steps = 9000                                    # 9000 * 120000 * 4 B = 4.32 GB per variable (sparse file)
    path = os.path.join(tmp_path, "big.nc")
    with pytest.raises(ValueError):
        NC3Writer(path, n_steps=steps, lat=lat, lon=lon, cell_index=np.arange(lat.size * lon.size),
                  variables={"temp": {}, "prec": {}}, version=2)
    with NC3Writer(path, n_steps=steps, lat=lat, lon=lon, cell_index=np.arange(lat.size * lon.size),
                   variables={"temp": {"units": "C"}, "prec": {}}, n_threads=2) as w:
        assert w.version == 5
        row = np.arange(lat.size * lon.size, dtype=np.float32)[None, :]
        w.write_steps("prec", steps - 1, row)
        begin_prec = w.vars["prec"][0]
    raw = open(path, "rb").read(4096)
    assert raw[:4] == b"CDF\x05" and struct.unpack(">q", raw[4:12])[0] == 0
    assert struct.unpack(">i", raw[12:16])[0] == 10 and struct.unpack(">q", raw[16:24])[0] == 3   # NC_DIMENSION, 3 dims
    assert raw[24:32] == struct.pack(">q", 4) and raw[32:36] == b"time" and struct.unpack(">q", raw[36:44])[0] == steps
    vsize = steps * lat.size * lon.size * 4
    assert struct.pack(">q", vsize) + struct.pack(">q", begin_prec) in raw       # 64-bit vsize, then begin
    with open(path, "rb") as fh:
        fh.seek(begin_prec + (steps - 1) * lat.size * lon.size * 4)
        got = np.frombuffer(fh.read(lat.size * lon.size * 4), dtype=">f4")
    np.testing.assert_array_equal(got, row[0])
    small = os.path.join(tmp_path, "small.nc")
    with NC3Writer(small, n_steps=4, lat=[0.0, 1.0], lon=[0.0], cell_index=[0, 1], variables={"temp": {}}) as w:
        assert w.version == 2
