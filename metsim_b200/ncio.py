"""Input side of the minimal driver: NetCDF files as plain numpy.

The reference opens its domain / state / forcing files with ``xr.open_dataset`` (metsim/io.py:264-303).
xarray and netCDF4 are absent from this image, so :func:`open_dataset` reads

* NetCDF classic / 64-bit offset / 64-bit data (``CDF\\x01``, ``CDF\\x02``, ``CDF\\x05``) with its own
  parser of the classic format (scipy reads the first two only; ``ncwriter.py`` writes the third for
  variables beyond 4 GiB);
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


# calendars whose years all have the same length: days per month (CF conventions 4.4.1)
UNIFORM_CALENDARS = {
    "noleap": (31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31),
    "365_day": (31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31),
    "all_leap": (31, 29, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31),
    "366_day": (31, 29, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31),
    "360_day": (30,) * 12,
}


def _parse_units(units: str):
    m = re.match(r"\s*(\w+)\s+since\s+(\d{1,4})-(\d{1,2})-(\d{1,2})"
                 r"(?:[ T]+(\d{1,2}):(\d{1,2})(?::(\d{1,2})(?:\.\d*)?)?)?", units or "")
    if not m or m.group(1).lower() not in _UNIT_SECONDS:
        raise ValueError(f"cannot decode time units {units!r}")
    y, mo, d = int(m.group(2)), int(m.group(3)), int(m.group(4))
    hh, mi, ss = (int(g) if g else 0 for g in m.group(5, 6, 7))
    return _UNIT_SECONDS[m.group(1).lower()], (y, mo, d), hh * 3600 + mi * 60 + ss


def _check_calendar(calendar: str) -> str:
    cal = (calendar or "standard").lower()
    if cal not in STANDARD_CALENDARS and cal not in UNIFORM_CALENDARS:
        raise NotImplementedError(f"calendar {calendar!r} is not decoded here (standard, gregorian, "
                                  "proleptic_gregorian, noleap / 365_day, all_leap / 366_day, 360_day are)")
    return cal


def decode_time(values, units: str, calendar: str = "standard") -> np.ndarray:
    """CF time axis -> ``datetime64[s]``.  Standard / gregorian / proleptic_gregorian are numpy's
    proleptic Gregorian calendar (the records in scope start after 1582).  For noleap / all_leap /
    360_day the date is computed in that calendar and then LABELLED as the Gregorian date with the
    same year, month and day -- what the reference does with ``CFTimeIndex.to_datetimeindex()``
    (metsim/io.py:281-283); a date that does not exist there (30 February) is an error, as it is for
    the reference."""
    cal = _check_calendar(calendar)
    unit, (y, mo, d), ref_secs = _parse_units(units)
    secs = np.round(np.asarray(values, dtype=np.float64) * unit).astype(np.int64)
    if cal in STANDARD_CALENDARS:
        ref = np.datetime64(f"{y:04d}-{mo:02d}-{d:02d}", "s") + np.timedelta64(ref_secs, "s")
        return ref + secs.astype("timedelta64[s]")
    months = np.asarray(UNIFORM_CALENDARS[cal])
    cum = np.concatenate([[0], np.cumsum(months)])
    total = (y * int(cum[-1]) + int(cum[mo - 1]) + (d - 1)) * 86400 + ref_secs + secs
    day, tod = total // 86400, total % 86400
    year, doy0 = day // int(cum[-1]), day % int(cum[-1])
    month0 = np.searchsorted(cum, doy0, side="right") - 1
    dom0 = doy0 - cum[month0]
    first = (year - 1970).astype("datetime64[Y]").astype("datetime64[M]") + month0.astype("timedelta64[M]")
    out = first.astype("datetime64[D]") + dom0.astype("timedelta64[D]")
    if (out.astype("datetime64[M]") != first).any():
        bad = np.atleast_1d(out)[np.atleast_1d(out.astype("datetime64[M]") != first)][0]
        raise ValueError(f"a date of the {cal} calendar has no Gregorian label (rolled over to {bad})")
    return out.astype("datetime64[s]") + tod.astype("timedelta64[s]")


def encode_time(times, units: str, calendar: str = "standard") -> np.ndarray:
    """Inverse of :func:`decode_time` (float64; what ``netCDF4.date2num`` gives the reference's writer,
    metsim/metsim.py:385-388)."""
    cal = _check_calendar(calendar)
    unit, (y, mo, d), ref_secs = _parse_units(units)
    t = np.asarray(times).astype("datetime64[s]")
    if cal in STANDARD_CALENDARS:
        ref = np.datetime64(f"{y:04d}-{mo:02d}-{d:02d}", "s") + np.timedelta64(ref_secs, "s")
        return (t - ref).astype(np.int64) / unit
    months = np.asarray(UNIFORM_CALENDARS[cal])
    cum = np.concatenate([[0], np.cumsum(months)])
    year = t.astype("datetime64[Y]").astype(np.int64) + 1970
    month0 = t.astype("datetime64[M]").astype(np.int64) % 12
    dom0 = (t.astype("datetime64[D]") - t.astype("datetime64[M]").astype("datetime64[D]")).astype(np.int64)
    if (dom0 >= months[month0]).any():
        raise ValueError(f"a date does not exist in the {cal} calendar (29 February?)")
    tod = (t - t.astype("datetime64[D]").astype("datetime64[s]")).astype(np.int64)
    total = (year * int(cum[-1]) + cum[month0] + dom0) * 86400 + tod
    ref_total = (y * int(cum[-1]) + int(cum[mo - 1]) + (d - 1)) * 86400 + ref_secs
    return (total - ref_total) / unit


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


_NC_TYPES = {1: ">i1", 2: "S1", 3: ">i2", 4: ">i4", 5: ">f4", 6: ">f8",
             7: ">u1", 8: ">u2", 9: ">u4", 10: ">i8", 11: ">u8"}        # 7..11: CDF-5 only
_NC_DIMENSION, _NC_VARIABLE, _NC_ATTRIBUTE = 10, 11, 12


def _open_classic(path: str) -> Dataset:
    """NetCDF classic (``CDF\\x01``), 64-bit offset (``\\x02``) and 64-bit data (``\\x05``) files, parsed
    here from the format specification (header = magic numrecs dim_list gatt_list var_list; fixed-size
    variables contiguous at ``begin``, record variables interleaved record by record) -- scipy reads
    the first two only, and the writer of this package (ncwriter.py) emits CDF-5 once a variable
    exceeds 4 GiB, so the driver reads what it writes at any size."""
    import struct
    buf = np.memmap(path, dtype=np.uint8, mode="r")
    version = int(buf[3])
    wide = version == 5
    pos = 4

    def u32():
        nonlocal pos
        v = struct.unpack(">I", buf[pos:pos + 4].tobytes())[0]
        pos += 4
        return v

    def cnt():                                   # NON_NEG: 32-bit, 64-bit in CDF-5
        nonlocal pos
        if not wide:
            return u32()
        v = struct.unpack(">Q", buf[pos:pos + 8].tobytes())[0]
        pos += 8
        return v

    def name():
        nonlocal pos
        n = cnt()
        s = buf[pos:pos + n].tobytes().decode("utf-8", "replace")
        pos += n + (-n % 4)
        return s

    def att_list():
        nonlocal pos
        tag, n = u32(), cnt()
        if tag == 0 and n == 0:
            return {}
        if tag != _NC_ATTRIBUTE:
            raise ValueError(f"{path}: malformed attribute list")
        out = {}
        for _ in range(n):
            k = name()
            t, m = u32(), cnt()
            dt = np.dtype(_NC_TYPES[t])
            nbytes = m * dt.itemsize
            raw = buf[pos:pos + nbytes].tobytes()
            pos += nbytes + (-nbytes % 4)
            if t == 2:
                out[k] = raw.decode("utf-8", "replace").rstrip("\x00")
            else:
                v = np.frombuffer(raw, dtype=dt).astype(dt.newbyteorder("="))
                out[k] = v[0] if m == 1 else v
        return out

    numrecs = cnt()
    streaming = numrecs == (0xFFFFFFFFFFFFFFFF if wide else 0xFFFFFFFF)
    tag, n = u32(), cnt()
    dims = []
    if not (tag == 0 and n == 0):
        if tag != _NC_DIMENSION:
            raise ValueError(f"{path}: malformed dimension list")
        for _ in range(n):
            dims.append((name(), cnt()))
    gatts = att_list()
    tag, n = u32(), cnt()
    variables = []
    if not (tag == 0 and n == 0):
        if tag != _NC_VARIABLE:
            raise ValueError(f"{path}: malformed variable list")
        for _ in range(n):
            vname = name()
            dimids = [cnt() for _ in range(cnt())]
            atts = att_list()
            t = u32()
            vsize = cnt()
            if version == 1:
                begin = u32()
            else:
                begin = struct.unpack(">Q", buf[pos:pos + 8].tobytes())[0]
                pos += 8
            variables.append((vname, dimids, atts, np.dtype(_NC_TYPES[t]), vsize, begin))
    is_rec = lambda dimids: bool(dimids) and dims[dimids[0]][1] == 0
    rec_vars = [v for v in variables if is_rec(v[1])]
    # record size: the sum of the record variables' vsize -- except that a single record variable is
    # not padded to 4 bytes (classic format specification, "Note on padding")
    if len(rec_vars) == 1:
        v = rec_vars[0]
        recsize = int(np.prod([dims[d][1] for d in v[1][1:]], dtype=np.int64)) * v[3].itemsize
    else:
        recsize = sum(v[4] for v in rec_vars)
    if streaming:
        if not rec_vars or recsize == 0:
            numrecs = 0
        else:
            numrecs = (buf.size - min(v[5] for v in rec_vars)) // recsize
    ds = Dataset(dims={k: (numrecs if ln == 0 else int(ln)) for k, ln in dims}, attrs=gatts)
    for vname, dimids, atts, dt, vsize, begin in variables:
        shape = tuple(ds.dims[dims[d][0]] for d in dimids)
        if is_rec(dimids):
            inner = shape[1:]
            if numrecs == 0:
                vals = np.empty(shape, dtype=dt.newbyteorder("="))
            else:
                strides = (recsize,) + tuple(int(np.prod(inner[i + 1:], dtype=np.int64)) * dt.itemsize
                                             for i in range(len(inner)))
                view = np.ndarray(shape=(numrecs,) + inner, dtype=dt, buffer=buf, offset=begin, strides=strides)
                vals = view.astype(dt.newbyteorder("=") if dt.kind != "S" else dt)
        else:
            view = np.ndarray(shape=shape, dtype=dt, buffer=buf, offset=begin)
            vals = view.astype(dt.newbyteorder("=") if dt.kind != "S" else dt)
        ds[vname] = Var(vals, tuple(dims[d][0] for d in dimids), atts)
    del buf
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
    if magic[:3] == b"CDF" and magic[3] in (1, 2, 5):
        return _open_classic(path)
    if magic == b"\x89HDF\r\n\x1a\n":
        return _open_hdf5(path)
    raise ValueError(f"{path}: not a NetCDF file")
