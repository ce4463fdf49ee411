"""A minimal pure-Python reader for NetCDF-4 / HDF5 input files (no h5py / netCDF4 / xarray in this
image): what the reference reads through ``xr.open_dataset`` (metsim/io.py:264-303) when a domain,
state or forcing file is NetCDF-4.  ``metsim_b200/driver.py`` opens such files with it; the fixture
generator under ``oracle/`` reads the bundled 90-day states with it.  File format work only -- no
arithmetic of the hot path lives here.

Scope: exactly what the bundled files under metsim/data use -- version-2 object headers ("OHDR",
with "OCHK" continuation chunks), root-group links stored compactly (link messages) or DENSELY
(link info message -> fractal heap + version-2 B-tree name index, what netCDF-4 writes once a group
has more than 8 links: state_vic.nc, passthrough.nc), simple dataspaces, fixed-point / IEEE
floating-point datatypes, CONTIGUOUS layout or CHUNKED layout (version-1 B-tree of raw data chunks --
every variable with the unlimited time dimension), unfiltered or through the deflate / shuffle /
fletcher32 filters netCDF-4 applies (filter pipeline message, versions 1 and 2), little- or big-endian;
attributes stored compactly (attribute messages, versions 1-3) or densely (attribute info message ->
fractal heap + version-2 B-tree), with fixed-length / variable-length strings, numbers and the
``DIMENSION_LIST`` object references of netCDF-4 (global heap), from which every variable's
dimension names are recovered.  Anything else raises ``NotImplementedError`` rather than guessing.  Layout follows the HDF5 File Format Specification 3.0
(sections III.A "Disk Format: Level 1 - Superblock", III.A.2 "Version 2 B-trees", III.G "Fractal
Heap", IV.A.1.b "Version 2 Data Object Header Prefix", IV.A.2 message types 0x01 dataspace, 0x02 link
info, 0x03 datatype, 0x06 link, 0x08 data layout, 0x0C attribute, 0x10 continuation, 0x15 attribute
info; III.E "Global Heap").

    from metsim_b200.hdf5_min import read_hdf5, read_hdf5_vars
    arrays = read_hdf5("state_nc.nc")        # {name: numpy array}
    nc = read_hdf5_vars("state_nc.nc")       # {name: H5Var(values, dims, attrs)}
"""
from __future__ import annotations

import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"


def _messages(b: bytes, off: int):
    """All (type, body) messages of the version-2 object header at ``off``."""
    if b[off:off + 4] != b"OHDR":
        raise NotImplementedError("only version-2 object headers are supported")
    flags = b[off + 5]
    pos = off + 6
    if flags & 0x20:
        pos += 16            # access / modification / change / birth times
    if flags & 0x10:
        pos += 4             # max compact / min dense attribute counts
    szlen = 1 << (flags & 3)
    chunk0 = int.from_bytes(b[pos:pos + szlen], "little")
    pos += szlen
    out, todo = [], [(pos, pos + chunk0)]
    while todo:
        q, end = todo.pop(0)
        while q + 4 <= end:
            mtype = b[q]
            msize = struct.unpack("<H", b[q + 1:q + 3])[0]
            q += 4 + (2 if flags & 0x04 else 0)      # + creation order when tracked
            body = b[q:q + msize]
            q += msize
            if mtype == 0x10:                        # object header continuation
                coff, clen = struct.unpack("<QQ", body[:16])
                if b[coff:coff + 4] != b"OCHK":
                    raise NotImplementedError("bad continuation chunk")
                todo.append((coff + 4, coff + clen - 4))
            else:
                out.append((mtype, body))
    return out


def _link_message(body: bytes):
    """(name, object header address) of one link message body, or None for soft / external links."""
    flags = body[1]
    q = 2
    if flags & 0x08:
        if body[q] != 0:
            return None
        q += 1
    if flags & 0x04:
        q += 8                                       # creation order
    if flags & 0x10:
        q += 1                                       # charset
    n = 1 << (flags & 3)
    ln = int.from_bytes(body[q:q + n], "little")
    q += n
    name = body[q:q + ln].decode("utf-8")
    q += ln
    return name, struct.unpack("<Q", body[q:q + 8])[0]


UNDEF = 0xFFFFFFFFFFFFFFFF


def _fractal_heap(b: bytes, heap_addr: int):
    """(heap ID length, lookup) of the fractal heap at ``heap_addr`` (spec III.G): ``lookup(id)`` returns
    the bytes of a managed object."""
    if b[heap_addr:heap_addr + 4] != b"FRHP" or b[heap_addr + 4] != 0:
        raise NotImplementedError("fractal heap header")
    q = heap_addr + 5
    id_len, filt_len = struct.unpack("<HH", b[q:q + 4]); q += 4
    q += 1                                           # flags
    max_managed = struct.unpack("<I", b[q:q + 4])[0]; q += 4
    q += 8 * 12                                      # huge-object / free-space / statistics fields
    width = struct.unpack("<H", b[q:q + 2])[0]; q += 2
    start_size, max_direct = struct.unpack("<QQ", b[q:q + 16]); q += 16
    heap_bits = struct.unpack("<H", b[q:q + 2])[0]; q += 2
    q += 2                                           # starting rows of the root indirect block
    root_addr = struct.unpack("<Q", b[q:q + 8])[0]; q += 8
    root_rows = struct.unpack("<H", b[q:q + 2])[0]
    if filt_len:
        raise NotImplementedError("filtered fractal heap")
    off_bytes = (heap_bits + 7) // 8
    len_bytes = (min(max_direct, max_managed).bit_length() + 7) // 8
    # ---- direct blocks: (offset in heap space, size, file address) ----
    blocks = []

    def row_size(r):
        return start_size if r < 2 else start_size << (r - 1)

    def walk_indirect(addr, rows):
        if b[addr:addr + 4] != b"FHIB":
            raise NotImplementedError("fractal heap indirect block")
        p = addr + 5 + 8
        block_off = int.from_bytes(b[p:p + off_bytes], "little"); p += off_bytes
        max_direct_rows = (max_direct.bit_length() - 1) - (start_size.bit_length() - 1) + 2
        cur = block_off
        for r in range(rows):
            for _ in range(width):
                child = struct.unpack("<Q", b[p:p + 8])[0]; p += 8
                if r >= max_direct_rows:
                    raise NotImplementedError("nested indirect blocks")
                if child != UNDEF:
                    blocks.append((cur, row_size(r), child))
                cur += row_size(r)

    if root_addr != UNDEF:
        if root_rows == 0:
            blocks.append((0, start_size, root_addr))
        else:
            walk_indirect(root_addr, root_rows)
    for _, _, addr in blocks:
        if b[addr:addr + 4] != b"FHDB":
            raise NotImplementedError("fractal heap direct block")

    def heap_object(hid: bytes) -> bytes:
        if hid[0] & 0xF0:
            raise NotImplementedError("huge / tiny fractal heap objects")
        off = int.from_bytes(hid[1:1 + off_bytes], "little")
        ln = int.from_bytes(hid[1 + off_bytes:1 + off_bytes + len_bytes], "little")
        for block_off, size, addr in blocks:         # offsets count from the block's own header
            if block_off <= off < block_off + size:
                return b[addr + (off - block_off):addr + (off - block_off) + ln]
        raise NotImplementedError("heap ID outside the allocated blocks")

    return id_len, heap_object


def _btree2_leaf_records(b: bytes, btree_addr: int, want_type: int):
    """Records of a version-2 B-tree (spec III.A.2) that is a single leaf -- every group / attribute
    index of the files in scope."""
    if b[btree_addr:btree_addr + 4] != b"BTHD" or b[btree_addr + 5] != want_type:
        raise NotImplementedError("version-2 B-tree header of type %d" % want_type)
    rec_size, depth = struct.unpack("<HH", b[btree_addr + 10:btree_addr + 14])
    root_node = struct.unpack("<Q", b[btree_addr + 16:btree_addr + 24])[0]
    n_root = struct.unpack("<H", b[btree_addr + 24:btree_addr + 26])[0]
    if n_root == 0 or root_node == UNDEF:
        return rec_size, []
    if depth != 0:
        raise NotImplementedError("B-tree deeper than one leaf")
    if b[root_node:root_node + 4] != b"BTLF":
        raise NotImplementedError("version-2 B-tree leaf")
    return rec_size, [b[root_node + 6 + k * rec_size:root_node + 6 + (k + 1) * rec_size]
                      for k in range(n_root)]


def _dense_links(b: bytes, heap_addr: int, btree_addr: int):
    """Links of a group with dense storage: the version-2 B-tree (type 5, link names) lists the
    fractal-heap IDs of the link messages: records = hash (4) + heap ID."""
    id_len, heap_object = _fractal_heap(b, heap_addr)
    rec_size, recs = _btree2_leaf_records(b, btree_addr, 5)
    if recs and rec_size != 4 + id_len:
        raise NotImplementedError("link-name B-tree record size")
    links = {}
    for rec in recs:
        got = _link_message(heap_object(rec[4:4 + id_len]))
        if got is not None:
            links[got[0]] = got[1]
    return links


def _links(msgs, b: bytes = b""):
    """name -> object header address: compact link messages (type 0x06) and, when the group's link
    info message (0x02) points at a fractal heap, the densely stored ones."""
    links = {}
    for mtype, body in msgs:
        if mtype == 0x02:
            q = 2 + (8 if body[1] & 1 else 0)
            heap_addr, name_btree = struct.unpack("<QQ", body[q:q + 16])
            if heap_addr != UNDEF:
                links.update(_dense_links(b, heap_addr, name_btree))
            continue
        if mtype != 0x06:
            continue
        flags = body[1]
        q = 2
        if flags & 0x08:
            if body[q] != 0:
                continue                             # soft / external link
            q += 1
        if flags & 0x04:
            q += 8                                   # creation order
        if flags & 0x10:
            q += 1                                   # charset
        n = 1 << (flags & 3)
        ln = int.from_bytes(body[q:q + n], "little")
        q += n
        name = body[q:q + ln].decode("utf-8")
        q += ln
        links[name] = struct.unpack("<Q", body[q:q + 8])[0]
    return links


def _chunks(b: bytes, addr: int, rank1: int):
    """(offsets, size, address) of every raw data chunk under the version-1 B-tree node at ``addr``
    (spec III.A.1, node type 1)."""
    if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 1:
        raise NotImplementedError("chunk B-tree node")
    level = b[addr + 5]
    used = struct.unpack("<H", b[addr + 6:addr + 8])[0]
    q = addr + 8 + 16                                # past the sibling pointers
    key = 8 + 8 * rank1
    out = []
    for _ in range(used):
        size, mask = struct.unpack("<II", b[q:q + 8])
        offs = struct.unpack("<%dQ" % rank1, b[q + 8:q + key])
        child = struct.unpack("<Q", b[q + key:q + key + 8])[0]
        q += key + 8
        if level > 0:
            out += _chunks(b, child, rank1)
        else:
            out.append((offs[:-1], size, child, mask))
    return out


def _filter_pipeline(body: bytes):
    """[(filter id, client data)] of a filter pipeline message (0x0B; spec IV.A.2.l, versions 1 and 2)."""
    ver, n = body[0], body[1]
    if ver == 1:
        q = 8
    elif ver == 2:
        q = 2
    else:
        raise NotImplementedError("filter pipeline message version %d" % ver)
    out = []
    for _ in range(n):
        fid = struct.unpack("<H", body[q:q + 2])[0]
        q += 2
        nlen = 0
        if ver == 1 or fid >= 256:
            nlen = struct.unpack("<H", body[q:q + 2])[0]
            q += 2
        _flags, ncd = struct.unpack("<HH", body[q:q + 4])
        q += 4
        q += nlen + ((-nlen % 8) if ver == 1 else 0)         # the optional name
        cd = struct.unpack("<%dI" % ncd, body[q:q + 4 * ncd])
        q += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
        out.append((fid, cd))
    return out


def _unfilter(raw: bytes, filters, mask: int, itemsize: int) -> bytes:
    """Undo a chunk's filters, last applied first; bit i of ``mask`` set = filter i was skipped for this
    chunk.  1 = deflate (zlib), 2 = shuffle (bytes of the elements de-interleaved), 3 = fletcher32
    (a 4-byte checksum after the data, not verified here)."""
    import zlib
    for idx in reversed(range(len(filters))):
        if (mask >> idx) & 1:
            continue
        fid, cd = filters[idx]
        if fid == 1:
            raw = zlib.decompress(raw)
        elif fid == 2:
            width = cd[0] if cd else itemsize
            n = len(raw) // width
            body = np.frombuffer(raw, dtype=np.uint8, count=n * width).reshape(width, n).T
            raw = body.tobytes() + raw[n * width:]
        elif fid == 3:
            raw = raw[:-4]
        else:
            raise NotImplementedError("HDF5 filter %d (only deflate, shuffle and fletcher32 are undone here)" % fid)
    return raw


def _datatype(body: bytes):
    """(kind, numpy dtype or None, size, base) of a datatype message body: kind is 'num', 'str'
    (fixed length), 'vlen_str', 'vlen' (sequence of ``base``) or 'ref' (object reference)."""
    cls, bits0 = body[0] & 0x0F, body[1]
    size = struct.unpack("<I", body[4:8])[0]
    order = ">" if bits0 & 1 else "<"
    if cls == 0:
        return "num", np.dtype("%s%s%d" % (order, "i" if bits0 & 0x08 else "u", size)), size, None
    if cls == 1:
        return "num", np.dtype("%sf%d" % (order, size)), size, None
    if cls == 3:
        return "str", None, size, None
    if cls == 7:
        if bits0 & 0x0F != 0 or size != 8:
            raise NotImplementedError("only object references")
        return "ref", np.dtype("<u8"), size, None
    if cls == 9:
        if bits0 & 0x0F == 1:
            return "vlen_str", None, size, None
        return "vlen", None, size, _datatype(body[8:])
    raise NotImplementedError("datatype class %d" % cls)


def _dataspace(body: bytes):
    ver, rank = body[0], body[1]
    if ver == 1:
        start = 8
    elif ver == 2:
        if body[3] == 2:
            return None                              # null dataspace
        start = 4
    else:
        raise NotImplementedError("dataspace version %d" % ver)
    return struct.unpack("<%dQ" % rank, body[start:start + 8 * rank])


def _global_heap_object(b: bytes, addr: int, index: int) -> bytes:
    """Object ``index`` of the global heap collection at ``addr`` (spec III.E)."""
    if b[addr:addr + 4] != b"GCOL" or b[addr + 4] != 1:
        raise NotImplementedError("global heap collection")
    end = addr + struct.unpack("<Q", b[addr + 8:addr + 16])[0]
    q = addr + 16
    while q + 16 <= end:
        idx, _refs, _res, size = struct.unpack("<HHIQ", b[q:q + 16])
        if idx == 0:
            break
        if idx == index:
            return b[q + 16:q + 16 + size]
        q += 16 + size + (-size % 8)
    raise NotImplementedError("global heap object %d not found" % index)


def _attribute(b: bytes, body: bytes):
    """(name, value) of one attribute message body (spec IV.A.2.m, versions 1-3).  Strings come back
    as ``str`` (or a list of them), numbers as numpy scalars / arrays, ``DIMENSION_LIST``-style
    variable-length references as a list of lists of object header addresses."""
    ver = body[0]
    nsz, tsz, ssz = struct.unpack("<HHH", body[2:8])
    q = 8 + (1 if ver == 3 else 0)
    pad = (lambda n: n + (-n % 8)) if ver == 1 else (lambda n: n)
    if ver not in (1, 2, 3):
        raise NotImplementedError("attribute message version %d" % ver)
    name = body[q:q + nsz].split(b"\x00")[0].decode("utf-8"); q += pad(nsz)
    kind, dt, size, base = _datatype(body[q:q + tsz]); q += pad(tsz)
    dims = _dataspace(body[q:q + ssz]); q += pad(ssz)
    if dims is None:
        return name, None
    n = int(np.prod(dims)) if dims else 1
    raw = body[q:q + n * size]
    if kind == "num":
        v = np.frombuffer(raw, dtype=dt, count=n).astype(dt.newbyteorder("="))
        return name, (v[0] if not dims else v.reshape(dims))
    if kind == "str":
        vals = [raw[i * size:(i + 1) * size].split(b"\x00")[0].decode("utf-8", "replace") for i in range(n)]
        return name, (vals[0] if not dims or n == 1 else vals)
    if kind in ("vlen_str", "vlen"):
        vals = []
        for i in range(n):
            ln, gaddr, gidx = struct.unpack("<IQI", raw[16 * i:16 * i + 16])
            if ln == 0 or gaddr == 0:
                vals.append("" if kind == "vlen_str" else [])
                continue
            obj = _global_heap_object(b, gaddr, gidx)
            if kind == "vlen_str":
                vals.append(obj[:ln].decode("utf-8", "replace"))
            elif base[0] == "ref":
                vals.append(list(struct.unpack("<%dQ" % ln, obj[:8 * ln])))
            else:
                raise NotImplementedError("variable-length sequences of non-references")
        return name, (vals[0] if kind == "vlen_str" and (not dims or n == 1) else vals)
    return name, None                                 # bare references (REFERENCE_LIST is compound: skipped above)


def _attributes(b: bytes, msgs) -> dict:
    """All attributes of an object: compact attribute messages (0x0C) plus, through the attribute
    info message (0x15), the densely stored ones (fractal heap + version-2 B-tree of type 8)."""
    out = {}

    def take(body):
        try:
            k, v = _attribute(b, body)
        except NotImplementedError:
            return                                    # e.g. the compound REFERENCE_LIST of dimension scales
        out[k] = v

    for mtype, body in msgs:
        if mtype == 0x0C:
            take(body)
        elif mtype == 0x15:
            flags = body[1]
            q = 2 + (2 if flags & 1 else 0)
            heap_addr, btree_addr = struct.unpack("<QQ", body[q:q + 16])
            if heap_addr == UNDEF or btree_addr == UNDEF:
                continue
            id_len, heap_object = _fractal_heap(b, heap_addr)
            _, recs = _btree2_leaf_records(b, btree_addr, 8)
            for rec in recs:                          # heap ID + message flags (1) + creation order (4) + hash (4)
                take(heap_object(rec[:id_len]))
    return out


def _dataset(b: bytes, off: int, msgs=None):
    dims = dtype = layout = chunked = filters = None
    for mtype, body in (msgs if msgs is not None else _messages(b, off)):
        if mtype == 0x01:                            # dataspace
            dims = _dataspace(body)
        elif mtype == 0x03:                          # datatype
            kind, dtype, _, _ = _datatype(body)
            if kind != "num":
                raise NotImplementedError("only numeric datasets")
        elif mtype == 0x08:                          # data layout
            if body[0] == 3 and body[1] == 1:
                layout = struct.unpack("<QQ", body[2:18])
            elif body[0] == 3 and body[1] == 2:      # chunked: dimensionality, B-tree address, chunk dims
                rank1 = body[2]
                tree = struct.unpack("<Q", body[3:11])[0]
                cdims = struct.unpack("<%dI" % rank1, body[11:11 + 4 * rank1])
                chunked = (rank1, tree, cdims)
            else:
                raise NotImplementedError("only contiguous / chunked layouts (version 3) are supported")
        elif mtype == 0x0B:
            filters = _filter_pipeline(body)
    native = None if dtype is None else dtype.newbyteorder("=")
    if dims is not None and dtype is not None and chunked is not None:
        rank1, tree, cdims = chunked
        if rank1 != len(dims) + 1 or cdims[-1] != dtype.itemsize:
            raise NotImplementedError("chunk shape does not match the dataspace")
        out = np.zeros(dims, dtype=native)
        if tree != UNDEF:
            for offs, size, addr, mask in _chunks(b, tree, rank1):
                cshape = cdims[:-1]
                raw = b[addr:addr + size]
                if filters:
                    raw = _unfilter(raw, filters, mask, dtype.itemsize)
                elif mask:
                    raise NotImplementedError("chunk with a filter mask in a dataset without filters")
                if len(raw) != int(np.prod(cshape)) * dtype.itemsize:
                    raise NotImplementedError("chunk size does not match its shape")
                block = np.frombuffer(raw, dtype=dtype, count=int(np.prod(cshape))).reshape(cshape)
                sl = tuple(slice(o, min(o + c, d)) for o, c, d in zip(offs, cshape, dims))
                out[sl] = block[tuple(slice(0, s_.stop - s_.start) for s_ in sl)]
        return out
    if dims is None or dtype is None or layout is None:
        raise NotImplementedError("not a simple dataset")
    addr, size = layout
    n = int(np.prod(dims)) if dims else 1
    if addr == UNDEF:                                # never written (a netCDF dimension without a variable)
        return np.zeros(dims, dtype=native)
    if size != n * dtype.itemsize:
        raise NotImplementedError("layout size does not match the dataspace")
    return np.frombuffer(b, dtype=dtype, count=n, offset=addr).reshape(dims).astype(native)


def _root(b: bytes) -> int:
    if b[:8] != SIG:
        raise ValueError("not an HDF5 file")
    ver = b[8]
    if ver == 0:          # superblock v0: root group symbol table entry holds the header address
        if b[13] != 8 or b[14] != 8:
            raise NotImplementedError("only 8-byte offsets / lengths")
        return struct.unpack("<Q", b[24 + 4 * 8 + 8:24 + 4 * 8 + 16])[0]
    if ver in (2, 3):
        return struct.unpack("<Q", b[36:44])[0]
    raise NotImplementedError("superblock version %d" % ver)


def read_hdf5(path: str) -> dict:
    """{name: numpy array} of every dataset of the root group."""
    b = open(path, "rb").read()
    root = _root(b)
    return {name: _dataset(b, addr) for name, addr in _links(_messages(b, root), b).items()}


class H5Var:
    """One netCDF-4 variable: ``values`` (numpy), ``dims`` (dimension names) and ``attrs``."""

    def __init__(self, values, dims, attrs):
        self.values, self.dims, self.attrs = values, tuple(dims), attrs

    def __repr__(self):
        return f"H5Var(dims={self.dims}, shape={self.values.shape}, dtype={self.values.dtype})"


DIM_ONLY = "This is a netCDF dimension but not a netCDF variable"


def read_hdf5_vars(path: str):
    """({name: H5Var}, {dimension name: length}, global attributes) of a netCDF-4 file's root group.
    Dimension names come from each variable's ``DIMENSION_LIST`` (object references to the
    dimension scales); a dimension scale names itself.  Dimensions that are not variables
    (``hru`` in dim_test.nc) appear only in the second mapping."""
    b = open(path, "rb").read()
    root = _root(b)
    root_msgs = _messages(b, root)
    links = _links(root_msgs, b)
    by_addr = {addr: name for name, addr in links.items()}
    out, dim_len = {}, {}
    parsed = {}
    for name, addr in links.items():
        msgs = _messages(b, addr)
        if not any(t == 0x01 for t, _ in msgs) or not any(t == 0x03 for t, _ in msgs):
            continue                                  # a sub-group or a committed datatype
        parsed[name] = (msgs, _attributes(b, msgs))
    for name, (msgs, attrs) in parsed.items():
        values = _dataset(b, links[name], msgs)
        is_scale = attrs.get("CLASS") == "DIMENSION_SCALE"
        if is_scale:
            dim_len[name] = values.shape[0] if values.ndim else 1
        if is_scale and str(attrs.get("NAME", "")).startswith(DIM_ONLY):
            continue
        dl = attrs.get("DIMENSION_LIST")
        if dl is not None:
            dims = tuple(by_addr[refs[0]] for refs in dl)
        elif is_scale and values.ndim == 1:
            dims = (name,)
        elif values.ndim == 0:
            dims = ()
        else:
            raise NotImplementedError(f"variable {name!r} has no DIMENSION_LIST")
        clean = {k: v for k, v in attrs.items()
                 if k not in ("CLASS", "NAME", "DIMENSION_LIST", "REFERENCE_LIST", "_Netcdf4Dimid",
                              "_Netcdf4Coordinates", "_nc3_strict")}
        out[name] = H5Var(values, dims, clean)
    try:
        gatts = {k: v for k, v in _attributes(b, root_msgs).items() if not k.startswith("_N")}
    except NotImplementedError:                       # e.g. a global attribute index deeper than one leaf:
        gatts = {}                                    # nothing on the path reads global attributes
    return out, dim_len, gatts
