"""TEST INFRASTRUCTURE ONLY -- a minimal pure-Python reader for the reference's NetCDF-4 / HDF5
fixtures (no h5py / netCDF4 in this image).

Scope: exactly what the bundled files under metsim/data use -- version-2 object headers ("OHDR",
with "OCHK" continuation chunks), root-group links stored compactly (link messages) or DENSELY
(link info message -> fractal heap + version-2 B-tree name index, what netCDF-4 writes once a group
has more than 8 links: state_vic.nc, passthrough.nc), simple dataspaces, fixed-point / IEEE
floating-point datatypes, CONTIGUOUS layout or CHUNKED layout without filters (version-1 B-tree
of raw data chunks -- every variable with the unlimited time dimension).  Anything else raises
``NotImplementedError`` rather than guessing.  Layout follows the HDF5 File Format Specification 3.0
(sections III.A "Disk Format: Level 1 - Superblock", III.A.2 "Version 2 B-trees", III.G "Fractal
Heap", IV.A.1.b "Version 2 Data Object Header Prefix", IV.A.2 message types 0x01 dataspace, 0x02 link
info, 0x03 datatype, 0x06 link, 0x08 data layout, 0x10 continuation).

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


def _dense_links(b: bytes, heap_addr: int, btree_addr: int):
    """Links of a group with dense storage: the version-2 B-tree (type 5, link names) lists the
    fractal-heap IDs of the link messages; the IDs are resolved through the heap's block tree."""
    # ---- fractal heap header (spec III.G) ----
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

    if root_addr == UNDEF:
        return {}
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

    # ---- version-2 B-tree, type 5 (link name index): records = hash (4) + heap ID (id_len) ----
    if b[btree_addr:btree_addr + 4] != b"BTHD" or b[btree_addr + 5] != 5:
        raise NotImplementedError("link-name B-tree header")
    rec_size, depth = struct.unpack("<HH", b[btree_addr + 10:btree_addr + 14])
    root_node = struct.unpack("<Q", b[btree_addr + 16:btree_addr + 24])[0]
    n_root = struct.unpack("<H", b[btree_addr + 24:btree_addr + 26])[0]
    if depth != 0:
        raise NotImplementedError("B-tree deeper than one leaf")
    if rec_size != 4 + id_len or b[root_node:root_node + 4] != b"BTLF":
        raise NotImplementedError("link-name B-tree leaf")
    links = {}
    for k in range(n_root):
        rec = b[root_node + 6 + k * rec_size:root_node + 6 + (k + 1) * rec_size]
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
            if mask:
                raise NotImplementedError("filtered chunk")
            out.append((offs[:-1], size, child))
    return out


def _dataset(b: bytes, off: int):
    dims = dtype = layout = chunked = None
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
            raise NotImplementedError("filtered (compressed) dataset")
    if dims is not None and dtype is not None and chunked is not None:
        rank1, tree, cdims = chunked
        if rank1 != len(dims) + 1 or cdims[-1] != dtype.itemsize:
            raise NotImplementedError("chunk shape does not match the dataspace")
        out = np.zeros(dims, dtype=dtype)
        if tree != UNDEF:
            for offs, size, addr in _chunks(b, tree, rank1):
                cshape = cdims[:-1]
                if size != int(np.prod(cshape)) * dtype.itemsize:
                    raise NotImplementedError("chunk size does not match its shape (filtered?)")
                block = np.frombuffer(b, dtype=dtype, count=int(np.prod(cshape)), offset=addr).reshape(cshape)
                sl = tuple(slice(o, min(o + c, d)) for o, c, d in zip(offs, cshape, dims))
                out[sl] = block[tuple(slice(0, s_.stop - s_.start) for s_ in sl)]
        return out
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
    for name, addr in _links(_messages(b, root), b).items():
        try:
            out[name] = _dataset(b, addr)
        except NotImplementedError:
            raise
    return out
