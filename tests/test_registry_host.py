"""Host-side logic of the MetSim drop-in (metsim_b200/registry.py) that needs no GPU: the
name-based gather of xarray-like arrays into [time][cell] order, the calendar vectors and the checks
that run before any device work."""
import numpy as np
import pytest

from metsim_b200 import registry


class DA:
    def __init__(self, values, dims):
        self.values = np.asarray(values)
        self.dims = tuple(dims)


def test_cells_gathers_by_dimension_name_in_any_order():
    rng = np.random.default_rng(1)
    a = rng.random((5, 3, 4))                                   # (time, lat, lon)
    mask = np.array([[1, 0, 1, 1], [0, 0, 1, 0], [1, 1, 0, 1]])
    sel = tuple(np.argwhere(mask > 0).T)
    want = a[:, sel[0], sel[1]]
    for perm, dims in (((0, 1, 2), ("time", "lat", "lon")), ((1, 0, 2), ("lat", "time", "lon")),
                       ((2, 1, 0), ("lon", "lat", "time"))):
        got = registry._cells(DA(np.transpose(a, perm), dims), "time", ("lat", "lon"), sel)
        np.testing.assert_array_equal(got, want)
        assert got.flags.c_contiguous and got.dtype == np.float64
    with pytest.raises(ValueError):
        registry._cells(DA(a, ("time", "y", "x")), "time", ("lat", "lon"), sel)


def test_per_cell_handles_1d_coordinates_and_2d_fields():
    mask = np.array([[1, 0, 1], [1, 1, 0]])
    sel = tuple(np.argwhere(mask > 0).T)
    lat, lon = np.array([10.0, 20.0]), np.array([-3.0, -2.0, -1.0])
    ds = {"lat": DA(lat, ("lat",)), "lon": DA(lon, ("lon",)),
          "elev": DA(np.arange(6.0).reshape(3, 2), ("lon", "lat"))}          # transposed on purpose
    np.testing.assert_array_equal(registry._per_cell(ds, "lat", ("lat", "lon"), mask.shape, sel), [10, 10, 20, 20])
    np.testing.assert_array_equal(registry._per_cell(ds, "lon", ("lat", "lon"), mask.shape, sel), [-3, -1, -3, -2])
    np.testing.assert_array_equal(registry._per_cell(ds, "elev", ("lat", "lon"), mask.shape, sel),
                                  np.arange(6.0).reshape(3, 2).T[sel])


def test_calendar_vectors_and_their_checks():
    t = np.datetime64("2000-02-27") + np.arange(5)                           # across the leap day
    doy, month = registry._calendar(t)
    assert doy.tolist() == [58, 59, 60, 61, 62] and month.tolist() == [2, 2, 2, 3, 3]
    assert doy.dtype == np.int32 and month.dtype == np.int32
    doy, _ = registry._calendar(np.array(["2000-12-31", "2001-01-01"], dtype="datetime64[ns]"))
    assert doy.tolist() == [366, 1]
    with pytest.raises(ValueError):
        registry._calendar(np.array(["2001-01-01", "2001-01-03"], dtype="datetime64[D]"))   # a gap
    with pytest.raises(NotImplementedError):
        registry._calendar(np.arange(3))                                      # not a datetime64 axis


def test_out_var_check_happens_before_any_device_work():
    class MS:
        params = {"time_step": 60, "out_vars": {"pet": {}}}                  # pet is a daily-mode variable
        domain = met_data = state = None
    with pytest.raises(ValueError, match="Cannot output variable pet at timestep 60"):
        registry.run_slice_batched(MS())
    MS.params = {"time_step": 1440, "out_vars": {"temp": {}}}
    with pytest.raises(ValueError, match="Cannot output variable temp at timestep 1440"):
        registry.run_slice_batched(MS())


def test_register_adds_both_method_modules():
    methods = {"mtclim": object()}
    out = registry.register(methods)
    assert out is methods and set(methods) == {"mtclim", "mtclim_b200", "passthrough_b200"}
    assert callable(methods["mtclim_b200"].run) and callable(methods["passthrough_b200"].run)


def test_shortwave_table_layout_rule_on_the_benchmark_grids():
    """Engine._sw_sectors: how many 32-byte sectors of the plain shortwave table the 32 cells of a
    warp touch per load decides between the plain table and its window-major copy (above 8).  Pure
    index arithmetic on (lon, lat_row): the dense CONUS grid stays plain, the masked global grids
    and any scattered cell set get the copy, with and without the patch cell order."""
    import torch
    from metsim_b200 import synth
    from metsim_b200.engine import Engine
    from metsim_b200.shard import patch_order

    def sectors(lat, lon, order=None):
        if order is not None:
            lat, lon = lat[order], lon[order]
        _, inv = np.unique(lat, return_inverse=True)
        return Engine._sw_sectors(torch.as_tensor(lon), torch.as_tensor(inv.astype(np.int32)))

    la, lo = synth.grid_cells("conus1/16", np.random.default_rng(synth.SEED))
    dense = sectors(la[:928 * 40], lo[:928 * 40])
    assert 4.0 < dense < 6.0                       # 16 consecutive tiny offsets = 128 bytes per warp load
    la, lo = synth.grid_cells("global0.5", np.random.default_rng(synth.SEED))
    assert sectors(la, lo) > 24.0 and sectors(la, lo, patch_order(la, lo, 4)) > 24.0
    la, lo = synth.grid_cells("global1/8", np.random.default_rng(synth.SEED))
    keep = slice(0, 200000)
    assert sectors(la[keep], lo[keep]) > 8.0
    assert sectors(la[keep], lo[keep], patch_order(la[keep], lo[keep], 4)) > 8.0
    assert sectors(la[:40], lo[:40]) == 0.0        # too few cells to matter
