"""Input side of the minimal driver: NetCDF files as plain numpy.

The reference opens its domain / state / forcing files with ``xr.open_dataset`` (metsim/io.py:264-303).
xarray and netCDF4 are absent from this image, so :func:`open_dataset` reads

* NetCDF classic / 64-bit offset (``CDF\\x01``, ``CDF\\x02``) through ``scipy.io.netcdf_file``;
* NetCDF-4 (HDF5) through :mod:`metsim_b200.hdf5_min`;

and hands back what the driver needs of an ``xarray.Dataset``: per variable ``values`` / ``dims`` /
``attrs``, with the CF decoding xarray applies by default (``_FillValue`` / ``missing_value`` ->
NaN, ``scale_factor`` / ``add_offset``, and the time axis as ``datetime64``).  File format work
only; nothing here is on the compute path.
"""
from __future__ import annotations

import re

import numpy as np

_UNIT_SECONDS = {"day": 86400, "days": 86400, "d": 86400, "hour": 3600, "hours": 3600, "hr": 3600,
                 "hrs": 3600, "h": 3600, "minute": 60, "minutes": 60, "min": 60, "mins": 60,
                 "second": 1, "seconds": 1, "sec": 1, "secs": 1, "s": 1}
STANDARD_CALENDARS = ("standard", "gregorian", "proleptic_gregorian")


class Var:
    """``values`` (numpy), ``dims`` (names) and ``attrs`` of one variable -- the part of
    ``xarray.DataArray`` that metsim_b200.registry / metsim_b200.driver touch."""

    def __init__(self, values, dims, attrs=None):
        self.values = values
        self.dims = tuple(dims)
        self.attrs = dict(attrs or {})

    @property
    def shape(self):
        return self.values.shape

    def __repr__(self):
        return f"Var(dims={self.dims}, shape={self.values.shape}, dtype={self.values.dtype})"


class Dataset(dict):
    """``{name: Var}`` plus ``dims`` (``{name: length}``) and global ``attrs``."""

    def __init__(self, variables=None, dims=None, attrs=None):
        super().__init__(variables or {})
        self.dims = dict(dims or {})
        self.attrs = dict(attrs or {})

    @property
    def variables(self):
        return self


def decode_time(values, units: str, calendar: str = "standard") -> np.ndarray:
    """CF time axis -> ``datetime64[s]`` (standard / gregorian / proleptic_gregorian calendars, all
    treated as numpy's proleptic Gregorian; the record lengths in scope start after 1582)."""
    cal = (calendar or "standard").lower()
    if cal not in STANDARD_CALENDARS:
        raise NotImplementedError(f"calendar {calendar!r}: the minimal driver decodes standard calendars only "
                                  "(the reference needs cftime for the others)")
    m = re.match(r"\s*(\w+)\s+since\s+(\d{1,4})-(\d{1,2})-(\d{1,2})"
                 r"(?:[ T]+(\d{1,2}):(\d{1,2})(?::(\d{1,2})(?:\.\d*)?)?)?", units or "")
    if not m or m.group(1).lower() not in _UNIT_SECONDS:
        raise ValueError(f"cannot decode time units {units!r}")
    y, mo, d = int(m.group(2)), int(m.group(3)), int(m.group(4))
    hh, mi, ss = (int(g) if g else 0 for g in m.group(5, 6, 7))
    ref = (np.datetime64(f"{y:04d}-{mo:02d}-{d:02d}", "s")
           + np.timedelta64(hh * 3600 + mi * 60 + ss, "s"))
    secs = np.round(np.asarray(values, dtype=np.float64) * _UNIT_SECONDS[m.group(1).lower()])
    return ref + secs.astype(np.int64).astype("timedelta64[s]")


def _decode(values: np.ndarray, attrs: dict) -> np.ndarray:
    """xarray's ``mask_and_scale``: fill values -> NaN (floating result), then scale and offset."""
    fills = [attrs[k] for k in ("_FillValue", "missing_value") if k in attrs]
    scale, offset = attrs.get("scale_factor"), attrs.get("add_offset")
    if not fills and scale is None and offset is None:
        return values
    out = values
    bad = None
    for f in fills:
        f = np.asarray(f).ravel()
        for fv in f:
            hit = np.isnan(values) if (values.dtype.kind == "f" and np.isnan(fv)) else values == fv
            bad = hit if bad is None else (bad | hit)
    if (bad is not None and bad.any()) or scale is not None or offset is not None:
        if out.dtype.kind != "f":
            out = out.astype(np.float64)
        elif bad is not None and bad.any():
            out = out.copy()
    if bad is not None and bad.any():
        out[bad] = np.nan
    if scale is not None:
        out = out * float(np.asarray(scale).ravel()[0])
    if offset is not None:
        out = out + float(np.asarray(offset).ravel()[0])
    return out


def _finish(ds: Dataset) -> Dataset:
    for name, v in ds.items():
        units = v.attrs.get("units")
        if isinstance(units, str) and " since " in units and v.values.dtype.kind in "iuf":
            v.values = decode_time(v.values, units, v.attrs.get("calendar", "standard"))
        elif v.values.dtype.kind in "iuf":
            v.values = _decode(v.values, v.attrs)
    return ds


def _open_classic(path: str) -> Dataset:
    from scipy.io import netcdf_file
    txt = lambda a: a.decode("utf-8", "replace") if isinstance(a, bytes) else a
    with netcdf_file(path, "r", mmap=False, maskandscale=False) as f:
        ds = Dataset(dims={k: (0 if n is None else int(n)) for k, n in f.dimensions.items()},
                     attrs={k: txt(v) for k, v in f._attributes.items()})
        for name, var in f.variables.items():
            vals = np.array(var.data)
            if vals.dtype.byteorder == ">" or (vals.dtype.byteorder == "=" and not np.little_endian):
                vals = vals.astype(vals.dtype.newbyteorder("="))
            ds[name] = Var(vals, var.dimensions, {k: txt(v) for k, v in var._attributes.items()})
            for d, n in zip(var.dimensions, vals.shape):
                ds.dims[d] = int(n)                  # the record dimension reads as None above
    return _finish(ds)


def _open_hdf5(path: str) -> Dataset:
    from .hdf5_min import read_hdf5_vars
    variables, dim_len, gatts = read_hdf5_vars(path)
    ds = Dataset(dims=dim_len, attrs=gatts)
    for name, v in variables.items():
        ds[name] = Var(v.values, v.dims, v.attrs)
    return _finish(ds)


def open_dataset(path: str) -> Dataset:
    """Read a whole NetCDF file (classic, 64-bit offset or NetCDF-4) into memory, CF-decoded."""
    with open(path, "rb") as f:
        magic = f.read(8)
    if magic[:3] == b"CDF" and magic[3] in (1, 2):
        return _open_classic(path)
    if magic == b"\x89HDF\r\n\x1a\n":
        return _open_hdf5(path)
    if magic[:3] == b"CDF" and magic[3] == 5:
        raise NotImplementedError(f"{path}: CDF-5 input files are not read by the minimal driver")
    raise ValueError(f"{path}: not a NetCDF file")
