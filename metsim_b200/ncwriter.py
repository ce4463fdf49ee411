"""Streaming NetCDF writer for the sub-daily outputs (the B200 form of ``MetSim.write`` /
``setup_output``, metsim/metsim.py:363-454, :560-587).

The reference allocates ``(time, lat, lon)`` arrays filled with NaN, scatters every cell's series
into them (metsim/metsim.py:584) and hands the dataset to xarray / netCDF4.  Neither library exists
in this image, and a CONUS year of outputs (135 GB) should never sit in host memory at once, so this
writer produces the file directly, tile by tile, as the rows arrive from ``msb_run_host``:

* format: NetCDF classic with 64-bit offsets ("CDF-2"; readable by netCDF4, xarray, scipy), fixed
  dimensions ``(time, lat, lon)``, one float32 / float64 variable per output -- each variable is one
  contiguous big-endian block, so a tile of steps is ONE ``pwrite`` per variable;
* the masked scatter ``[step][cell] -> [step][lat][lon]`` (NaN where ``mask <= 0``), the
  ``out_precision`` cast having happened on the device already;
* it is a consumer of :func:`metsim_b200.engine.run_host`'s ``on_tile`` hook: while tile *k* is being
  scattered and written here, tile *k + 1* is computing and copying.

No compute happens here and nothing falls back to the CPU: the arrays it writes come from the
CUDA path.
"""
from __future__ import annotations

import os
import struct

import numpy as np

NC_BYTE, NC_CHAR, NC_SHORT, NC_INT, NC_FLOAT, NC_DOUBLE = 1, 2, 3, 4, 5, 6
NC_DIMENSION, NC_VARIABLE, NC_ATTRIBUTE = 10, 11, 12
_TYPES = {np.dtype("float32"): (NC_FLOAT, ">f4"), np.dtype("float64"): (NC_DOUBLE, ">f8"),
          np.dtype("int32"): (NC_INT, ">i4")}


def _pad4(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 4)


def _name(s: str) -> bytes:
    raw = s.encode("utf-8")
    return struct.pack(">i", len(raw)) + _pad4(raw)


def _att_list(atts: dict) -> bytes:
    if not atts:
        return struct.pack(">ii", 0, 0)
    out = struct.pack(">ii", NC_ATTRIBUTE, len(atts))
    for k, v in atts.items():
        if isinstance(v, str):
            raw = v.encode("utf-8")
            out += _name(k) + struct.pack(">ii", NC_CHAR, len(raw)) + _pad4(raw)
        else:
            arr = np.atleast_1d(np.asarray(v, dtype=np.float64))
            out += _name(k) + struct.pack(">ii", NC_DOUBLE, arr.size) + arr.astype(">f8").tobytes()
    return out


class NC3Writer:
    """``(time, lat, lon)`` NetCDF-3 (64-bit offset) file written in time tiles.

    ``cell_index`` maps the engine's cell order to flat ``lat * n_lon + lon`` positions of the
    domain grid (``np.flatnonzero(mask > 0)`` for a latitude-row-major shard)."""

    def __init__(self, path: str, *, n_steps: int, lat, lon, cell_index, variables: dict,
                 time_values=None, time_units: str = "minutes since start",
                 global_attrs: dict | None = None):
        self.lat = np.asarray(lat, dtype=np.float64)
        self.lon = np.asarray(lon, dtype=np.float64)
        self.n_steps = int(n_steps)
        self.n_grid = self.lat.size * self.lon.size
        self.cell_index = np.asarray(cell_index, dtype=np.int64)
        if self.cell_index.size and (self.cell_index.min() < 0 or self.cell_index.max() >= self.n_grid):
            raise ValueError("cell_index outside the (lat, lon) grid")
        self.vars = {}
        dims = (("time", self.n_steps), ("lat", self.lat.size), ("lon", self.lon.size))
        coord = {"time": (np.arange(self.n_steps, dtype=np.float64) if time_values is None
                          else np.asarray(time_values, dtype=np.float64)),
                 "lat": self.lat, "lon": self.lon}
        coord_atts = {"time": {"units": time_units}, "lat": {"units": "degrees_north"},
                      "lon": {"units": "degrees_east"}}
        if coord["time"].size != self.n_steps:
            raise ValueError("time_values must have n_steps entries")
        # ---- header (classic format spec: magic, numrecs, dim_list, gatt_list, var_list) ----
        entries = []   # (name, dimids, atts, nc_type, be_dtype, n_elems)
        for i, (dname, _) in enumerate(dims):
            entries.append((dname, (i,), coord_atts[dname], NC_DOUBLE, ">f8", coord[dname].size))
        for vname, spec in variables.items():
            dt = np.dtype(spec.get("dtype", "float32"))
            nct, be = _TYPES[dt]
            atts = {k: v for k, v in spec.items() if k != "dtype"}   # units, long_name, ...
            entries.append((vname, (0, 1, 2), atts, nct, be, self.n_steps * self.n_grid))

        def header(begins):
            h = b"CDF\x02" + struct.pack(">i", 0)
            h += struct.pack(">ii", NC_DIMENSION, len(dims))
            for dname, dlen in dims:
                h += _name(dname) + struct.pack(">i", dlen)
            h += _att_list(dict(global_attrs or {}, source="metsim_b200"))
            h += struct.pack(">ii", NC_VARIABLE, len(entries))
            for (vname, dimids, atts, nct, be, n), begin in zip(entries, begins):
                vsize = n * np.dtype(be).itemsize
                vsize += -vsize % 4
                h += _name(vname) + struct.pack(">i", len(dimids))
                h += b"".join(struct.pack(">i", d) for d in dimids)
                h += _att_list(atts)      # cells outside the mask hold NaN, as in the reference
                h += struct.pack(">i", nct) + struct.pack(">I", min(vsize, 0xFFFFFFFF)) + struct.pack(">q", begin)
            return h

        size0 = len(header([0] * len(entries)))
        begins, pos = [], size0 + (-size0 % 4)
        for (_, _, _, _, be, n) in entries:
            begins.append(pos)
            nbytes = n * np.dtype(be).itemsize
            pos += nbytes + (-nbytes % 4)
        hdr = header(begins)
        assert len(hdr) == size0
        self.path = path
        self.fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
        os.ftruncate(self.fd, pos)
        os.pwrite(self.fd, hdr, 0)
        for (vname, _, _, _, be, n), begin in zip(entries, begins):
            if vname in coord:
                os.pwrite(self.fd, coord[vname].astype(be).tobytes(), begin)
            else:
                self.vars[vname] = (begin, np.dtype(be))
        self.steps_written = {v: 0 for v in self.vars}
        self._grid = {}
        # every grid cell active and in grid order (e.g. CONUS 1/16 deg): no scatter needed
        self._dense = (self.cell_index.size == self.n_grid and
                       bool((self.cell_index == np.arange(self.n_grid)).all()))

    def write_steps(self, name: str, step0: int, block) -> None:
        """Scatter ``block[steps][cells]`` to ``(steps, lat, lon)`` (NaN elsewhere) and write it at
        ``step0``.  Tiles may arrive in any order; each (variable, step) is written once."""
        begin, be = self.vars[name]
        block = np.asarray(block)
        steps = block.shape[0]
        if block.shape[1] != self.cell_index.size or step0 < 0 or step0 + steps > self.n_steps:
            raise ValueError("block does not fit the file")
        if self._dense:                   # byte-swapped copy of the block, written as it is
            os.pwrite(self.fd, np.ascontiguousarray(block, dtype=be).data,
                      begin + step0 * self.n_grid * be.itemsize)
            self.steps_written[name] += steps
            return
        key = (steps, be)
        grid = self._grid.get(key)
        if grid is None:                  # reused scatter buffer, NaN where no cell maps
            grid = np.full((steps, self.n_grid), np.nan, dtype=be)
            self._grid = {key: grid}
        grid[:, self.cell_index] = block  # byte order conversion happens in this assignment
        os.pwrite(self.fd, grid.data, begin + step0 * self.n_grid * be.itemsize)
        self.steps_written[name] += steps

    def tile_consumer(self, out: dict, tile_days: int, tpd: int, ring: bool):
        """``on_tile`` hook for :func:`metsim_b200.engine.run_host`: writes every variable of the tile
        from ``out`` (the host arrays given to ``run_host``; a 2-slot ring if ``ring``)."""
        def on_tile(tile, day0, day1, slot):
            steps = (day1 - day0) * tpd
            src0 = slot * tile_days * tpd if ring else day0 * tpd
            for v in self.vars:
                if v in out:
                    self.write_steps(v, day0 * tpd, out[v][src0:src0 + steps])
        return on_tile

    def close(self) -> None:
        if self.fd is not None:
            os.close(self.fd)
            self.fd = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
