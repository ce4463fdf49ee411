"""Streaming NetCDF writer for the sub-daily outputs (the B200 form of ``MetSim.write`` /
``setup_output``, metsim/metsim.py:363-454, :560-587).

The reference allocates ``(time, lat, lon)`` arrays filled with NaN, scatters every cell's series
into them (metsim/metsim.py:584) and hands the dataset to xarray / netCDF4.  Neither library exists
in this image, and a CONUS year of outputs (135 GB) should never sit in host memory at once, so this
writer produces the file directly, tile by tile, as the rows arrive from ``msb_run_host``:

* format: NetCDF classic, fixed dimensions ``(time, lat, lon)``, one float32 / float64 variable per
  output -- each variable is one contiguous big-endian block, so a tile of steps is a handful of
  ``pwrite`` calls per variable.  Files whose variables all stay below 4 GiB are written with
  64-bit offsets ("CDF-2": readable by netCDF4, xarray AND scipy); larger ones (a CONUS year is 17 GB
  per variable) as "CDF-5" (64-bit data, netCDF-C >= 4.4), because CDF-2 allows only the LAST
  fixed-size variable to exceed 2**32 - 4 bytes;
* the masked scatter ``[step][cell] -> [step][lat][lon]`` (NaN where ``mask <= 0``), the
  ``out_precision`` cast having happened on the device already;
* it is a consumer of :func:`metsim_b200.engine.run_host`'s ``on_tile`` hook: while tile *k* is being
  scattered and written here, tile *k + 1* is computing and copying;
* scatter, byte swap and ``pwrite`` run in the native thread pool of ``csrc/sink.cu``
  (``msb_sink_*``; a single numpy thread cannot keep up with the 50+ GB/s the tiles arrive at);
  ``native=False`` keeps the numpy form, which the tests compare it with.

No compute happens here and nothing falls back to the CPU: the arrays it writes come from the
CUDA path.
"""
from __future__ import annotations

import ctypes as C
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
    """``(time, lat, lon)`` -- or ``(time, *grid_dims)`` -- NetCDF-3 (64-bit offset) file written in
    time tiles.

    ``cell_index`` maps the engine's cell order to flat row-major positions of the domain grid
    (``lat * n_lon + lon``; ``np.flatnonzero(mask > 0)`` for a latitude-row-major shard)."""

    def __init__(self, path: str, *, n_steps: int, lat=None, lon=None, cell_index, variables: dict,
                 time_values=None, time_units: str = "minutes since start",
                 global_attrs: dict | None = None, native: bool = True, n_threads: int = 0,
                 version: int | None = None, grid_dims=None, time_attrs: dict | None = None):
        """``grid_dims``: ``[(name, coordinate values), ...]`` for a domain whose mask is not over
        ``(lat, lon)`` (the reference writes ``('time',) + mask.dims``, metsim/metsim.py:367-368 --
        e.g. ``('time', 'hru')`` for dim_test.nc); ``lat`` / ``lon`` are then ignored.  An integer
        ``time_values`` array is written as NC_INT (the reference's ``i4`` time variable, :384);
        ``time_attrs`` adds to the time variable's attributes (calendar, long_name ...)."""
        if grid_dims is None:
            grid_dims = [("lat", np.asarray(lat, dtype=np.float64)), ("lon", np.asarray(lon, dtype=np.float64))]
            coord_atts = {"lat": {"units": "degrees_north"}, "lon": {"units": "degrees_east"}}
        else:
            grid_dims = [(str(k), np.asarray(v)) for k, v in grid_dims]
            coord_atts = {}
        self.grid_dims = grid_dims
        self.n_steps = int(n_steps)
        self.n_grid = int(np.prod([v.size for _, v in grid_dims]))
        self.cell_index = np.asarray(cell_index, dtype=np.int64)
        if self.cell_index.size and (self.cell_index.min() < 0 or self.cell_index.max() >= self.n_grid):
            raise ValueError("cell_index outside the domain grid")
        self.vars = {}
        dims = (("time", self.n_steps),) + tuple((k, v.size) for k, v in grid_dims)
        tvals = np.arange(self.n_steps, dtype=np.float64) if time_values is None else np.asarray(time_values)
        coord = {"time": tvals, **dict(grid_dims)}
        coord_atts["time"] = dict({"units": time_units}, **(time_attrs or {}))
        if coord["time"].size != self.n_steps:
            raise ValueError("time_values must have n_steps entries")
        # ---- header (classic format spec: magic, numrecs, dim_list, gatt_list, var_list) ----
        entries = []   # (name, dimids, atts, nc_type, be_dtype, n_elems)
        for i, (dname, _) in enumerate(dims):
            if coord[dname].dtype.kind in "iu":       # classic NetCDF has no 64-bit integers
                if coord[dname].size and np.abs(coord[dname]).max() > 0x7FFFFFFF:
                    raise ValueError(f"integer coordinate {dname!r} does not fit NC_INT")
                entries.append((dname, (i,), coord_atts.get(dname, {}), NC_INT, ">i4", coord[dname].size))
            else:
                entries.append((dname, (i,), coord_atts.get(dname, {}), NC_DOUBLE, ">f8", coord[dname].size))
        for vname, spec in variables.items():
            dt = np.dtype(spec.get("dtype", "float32"))
            nct, be = _TYPES[dt]
            atts = {k: v for k, v in spec.items() if k != "dtype"}   # units, long_name, ...
            entries.append((vname, tuple(range(len(dims))), atts, nct, be, self.n_steps * self.n_grid))

        # CDF-2 (64-bit offsets) holds fixed-size variables of up to 2**32 - 4 bytes (only the last one
        # may be larger); beyond that the file is CDF-5: 64-bit counts, dimension lengths and vsize
        biggest = max(n * np.dtype(be).itemsize for (_, _, _, _, be, n) in entries)
        if version is None:
            version = 2 if biggest <= 0xFFFFFFFF - 3 else 5
        if version == 2 and biggest > 0xFFFFFFFF - 3:
            raise ValueError("a variable exceeds 2**32 - 4 bytes: CDF-2 cannot hold it, use version=5")
        if version not in (2, 5):
            raise ValueError("version must be 2 (64-bit offset) or 5 (64-bit data)")
        self.version = version
        cnt = (lambda v: struct.pack(">q", v)) if version == 5 else (lambda v: struct.pack(">i", v))

        def name5(sname):
            raw = sname.encode("utf-8")
            return cnt(len(raw)) + _pad4(raw)

        def atts5(atts):
            if not atts:
                return struct.pack(">i", 0) + cnt(0)
            out = struct.pack(">i", NC_ATTRIBUTE) + cnt(len(atts))
            for k, v in atts.items():
                if isinstance(v, str):
                    raw = v.encode("utf-8")
                    out += name5(k) + struct.pack(">i", NC_CHAR) + cnt(len(raw)) + _pad4(raw)
                else:
                    arr = np.atleast_1d(np.asarray(v, dtype=np.float64))
                    out += name5(k) + struct.pack(">i", NC_DOUBLE) + cnt(arr.size) + arr.astype(">f8").tobytes()
            return out

        def header(begins):
            h = b"CDF" + bytes([version]) + cnt(0)                 # numrecs: no record dimension
            h += struct.pack(">i", NC_DIMENSION) + cnt(len(dims))
            for dname, dlen in dims:
                h += name5(dname) + cnt(dlen)
            h += atts5(dict(global_attrs or {}, source="metsim_b200"))
            h += struct.pack(">i", NC_VARIABLE) + cnt(len(entries))
            for (vname, dimids, atts, nct, be, n), begin in zip(entries, begins):
                vsize = n * np.dtype(be).itemsize
                vsize += -vsize % 4
                h += name5(vname) + cnt(len(dimids))
                h += b"".join(cnt(d) for d in dimids)
                h += atts5(atts)          # cells outside the mask hold NaN, as in the reference
                h += struct.pack(">i", nct) + (struct.pack(">q", vsize) if version == 5
                                                else struct.pack(">I", vsize)) + struct.pack(">q", begin)
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
        self._sinks = {}
        self._native = bool(native)
        self._n_threads = int(n_threads)
        self._cell_index_c = np.ascontiguousarray(self.cell_index, dtype=np.int64)
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
        if self._native and block.dtype.itemsize == be.itemsize and block.dtype.kind == "f":
            self._sink(be.itemsize).write(block, begin + step0 * self.n_grid * be.itemsize)
            self.steps_written[name] += steps
            return
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

    def _sink(self, itemsize: int):
        sk = self._sinks.get(itemsize)
        if sk is None:
            sk = self._sinks[itemsize] = NativeSink(
                n_cells=self.cell_index.size, n_grid=self.n_grid, elem_bytes=itemsize,
                cell_index=None if self._dense else self._cell_index_c, byteswap=True,
                n_threads=self._n_threads, fd=self.fd)
        return sk

    def close(self) -> None:
        for sk in self._sinks.values():
            sk.close()
        self._sinks = {}
        if self.fd is not None:
            os.close(self.fd)
            self.fd = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class NativeSink:
    """``msb_sink`` (include/metsim_b200.h, csrc/sink.cu): masked scatter + byte swap + ``pwrite`` of
    ``[steps][cells]`` rows on a pool of host threads.  ``fd=-1`` consumes without a file (the rows
    are still scattered and swapped into the staging buffers; ``checksum`` proves they were read)."""

    def __init__(self, *, n_cells: int, n_grid: int | None = None, elem_bytes: int = 4, cell_index=None,
                 byteswap: bool = True, n_threads: int = 0, fd: int = -1):
        from . import _lib
        self._lib = _lib.lib()
        d = _lib.SinkDesc()
        d.n_cells, d.elem_bytes = int(n_cells), int(elem_bytes)
        d.n_grid = int(n_grid if n_grid is not None else n_cells)
        self._index = None if cell_index is None else np.ascontiguousarray(cell_index, dtype=np.int64)
        if self._index is not None and self._index.size != n_cells:
            raise ValueError("cell_index must have n_cells entries")
        d.cell_index = None if self._index is None else self._index.ctypes.data_as(C.c_void_p)
        d.byteswap, d.n_threads, d.fd = int(bool(byteswap)), int(n_threads), int(fd)
        h = C.c_void_p()
        _lib.check(self._lib.msb_sink_create(C.byref(d), C.byref(h)))
        self.handle, self.n_cells, self.elem_bytes = h, int(n_cells), int(elem_bytes)

    def write(self, block: np.ndarray, file_offset: int = 0) -> None:
        from . import _lib
        if (block.ndim != 2 or block.shape[1] != self.n_cells or block.dtype.itemsize != self.elem_bytes
                or not block.flags.c_contiguous):
            raise ValueError("sink rows must be a C-contiguous [steps][n_cells] array of the sink's element size")
        _lib.check(self._lib.msb_sink_write(self.handle, block.ctypes.data_as(C.c_void_p),
                                            block.shape[0], int(file_offset)))

    @property
    def checksum(self) -> int:
        return int(self._lib.msb_sink_checksum(self.handle))

    @property
    def bytes_written(self) -> int:
        return int(self._lib.msb_sink_bytes_written(self.handle))

    def close(self) -> None:
        if getattr(self, "handle", None):
            self._lib.msb_sink_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
