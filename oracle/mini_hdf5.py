"""TEST INFRASTRUCTURE ONLY -- a minimal pure-Python reader for the reference's NetCDF-4 / HDF5
fixtures (no h5py / netCDF4 in this image).

Scope: exactly what the bundled files under metsim/data use -- version-2 object headers ("OHDR",
with "OCHK" continuation chunks), compact link messages in the root group, simple dataspaces,
fixed-point / IEEE floating-point datatypes, CONTIGUOUS uncompressed layout.  Anything else raises
``NotImplementedError`` rather than guessing.  Layout follows the HDF5 File Format Specification 3.0
(sections III.A "Disk Format: Level 1 - Superblock", IV.A.1.b "Version 2 Data Object Header Prefix",
IV.A.2 message types 0x01 dataspace, 0x03 datatype, 0x06 link, 0x08 data layout, 0x10 continuation).

    from oracle.mini_hdf5 import read_hdf5
    arrays = read_hdf5("/root/reference/metsim/data/state_nc.nc")   # {name: numpy array}
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


def _links(msgs):
    """name -> object header address from compact link messages (type 0x06)."""
    links = {}
    for mtype, body in msgs:
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


def _dataset(b: bytes, off: int):
    dims = dtype = layout = None
    for mtype, body in _messages(b, off):
        if mtype == 0x01:                            # dataspace
            ver, rank = body[0], body[1]
            start = 4 if ver == 2 else 8
            dims = struct.unpack("<%dQ" % rank, body[start:start + 8 * rank])
        elif mtype == 0x03:                          # datatype
            cls, bits0 = body[0] & 0x0F, body[1]
            size = struct.unpack("<I", body[4:8])[0]
            if bits0 & 1:
                raise NotImplementedError("big-endian datatype")
            if cls == 0:
                dtype = np.dtype("<%s%d" % ("i" if bits0 & 0x08 else "u", size))
            elif cls == 1:
                dtype = np.dtype("<f%d" % size)
            else:
                raise NotImplementedError("datatype class %d" % cls)
        elif mtype == 0x08:                          # data layout
            if body[0] != 3 or body[1] != 1:
                raise NotImplementedError("only contiguous layout (version 3, class 1) is supported")
            layout = struct.unpack("<QQ", body[2:18])
        elif mtype == 0x0B:
            raise NotImplementedError("filtered (compressed) dataset")
    if dims is None or dtype is None or layout is None:
        raise NotImplementedError("not a simple dataset")
    addr, size = layout
    n = int(np.prod(dims)) if dims else 1
    if size != n * dtype.itemsize:
        raise NotImplementedError("layout size does not match the dataspace")
    return np.frombuffer(b, dtype=dtype, count=n, offset=addr).reshape(dims).copy()


def read_hdf5(path: str) -> dict:
    b = open(path, "rb").read()
    if b[:8] != SIG:
        raise ValueError("not an HDF5 file")
    ver = b[8]
    if ver == 0:          # superblock v0: root group symbol table entry holds the header address
        if b[13] != 8 or b[14] != 8:
            raise NotImplementedError("only 8-byte offsets / lengths")
        root = struct.unpack("<Q", b[24 + 4 * 8 + 8:24 + 4 * 8 + 16])[0]
    elif ver in (2, 3):
        root = struct.unpack("<Q", b[36:44])[0]
    else:
        raise NotImplementedError("superblock version %d" % ver)
    out = {}
    for name, addr in _links(_messages(b, root)).items():
        try:
            out[name] = _dataset(b, addr)
        except NotImplementedError:
            raise
    return out
