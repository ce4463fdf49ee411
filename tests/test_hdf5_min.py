"""metsim_b200/hdf5_min.py on what cannot be taken from the reference's bundled files: filtered
(deflate / shuffle / fletcher32) chunked datasets, as netCDF-4 writes them when compression is on.
No HDF5 library exists in this image, so the test assembles a minimal HDF5 file by hand from the
format specification (superblock v2, version-2 object headers, a compact link, layout v3 with a
version-1 chunk B-tree, filter pipeline message) and reads it back."""
import os
import struct
import zlib

import numpy as np
import pytest

from metsim_b200 import hdf5_min

UNDEF = 0xFFFFFFFFFFFFFFFF


def _msg(mtype: int, body: bytes) -> bytes:
    return struct.pack("<BHB", mtype, len(body), 0) + body


def _ohdr(messages: bytes) -> bytes:
    """Version-2 object header, 2-byte chunk-0 size, no times / phase-change fields, dummy checksum."""
    return b"OHDR" + bytes([2, 0x01]) + struct.pack("<H", len(messages)) + messages + b"\0\0\0\0"


def _shuffle(raw: bytes, width: int) -> bytes:
    n = len(raw) // width
    return np.frombuffer(raw, dtype=np.uint8, count=n * width).reshape(n, width).T.tobytes() + raw[n * width:]


def build_file(path, data, chunk, pipeline_version, filters, skip_deflate_on=()):
    """One float32 / float64 dataset ``var`` of shape ``data.shape``, chunked, every chunk through
    ``filters`` (a list of 'shuffle', 'deflate', 'fletcher32' in application order)."""
    data = np.ascontiguousarray(data)
    rank, isz = data.ndim, data.dtype.itemsize
    ids = {"deflate": 1, "shuffle": 2, "fletcher32": 3}
    cd = {"deflate": (6,), "shuffle": (isz,), "fletcher32": ()}
    if pipeline_version == 2:
        fp = bytes([2, len(filters)])
        for f in filters:
            fp += struct.pack("<HHH", ids[f], 0, len(cd[f])) + b"".join(struct.pack("<I", c) for c in cd[f])
    else:
        fp = bytes([1, len(filters)]) + b"\0" * 6
        for f in filters:
            name = f.encode() + b"\0"
            name += b"\0" * (-len(name) % 8)
            fp += struct.pack("<HHHH", ids[f], len(name), 0, len(cd[f])) + name
            fp += b"".join(struct.pack("<I", c) for c in cd[f]) + (b"\0" * 4 if len(cd[f]) % 2 else b"")
    # chunks
    blobs, keys = [], []
    grid = [range(0, s, c) for s, c in zip(data.shape, chunk)]
    for k, offs in enumerate(np.array(np.meshgrid(*grid, indexing="ij")).reshape(rank, -1).T):
        block = np.zeros(chunk, dtype=data.dtype)
        sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, data.shape))
        part = data[sl]
        block[tuple(slice(0, n) for n in part.shape)] = part
        raw, mask = block.tobytes(), 0
        for i, f in enumerate(filters):
            if f == "deflate" and k in skip_deflate_on:
                mask |= 1 << i
                continue
            raw = {"shuffle": lambda r: _shuffle(r, isz), "deflate": lambda r: zlib.compress(r, 6),
                   "fletcher32": lambda r: r + b"\xde\xad\xbe\xef"}[f](raw)
        blobs.append(raw)
        keys.append((len(raw), mask, tuple(int(o) for o in offs) + (0,)))
    # layout of the file: superblock (48) | root header | dataset header | B-tree node | chunk data
    dspace = bytes([2, rank, 0, 1]) + b"".join(struct.pack("<Q", s) for s in data.shape)
    if data.dtype == np.float32:
        dtype = bytes([0x11, 0x20, 0x1F, 0x00]) + struct.pack("<I", 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    else:
        dtype = bytes([0x11, 0x20, 0x3F, 0x00]) + struct.pack("<I", 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)

    def dataset_header(btree_addr):
        layout = bytes([3, 2, rank + 1]) + struct.pack("<Q", btree_addr) + b"".join(
            struct.pack("<I", c) for c in tuple(chunk) + (isz,))
        return _ohdr(_msg(0x01, dspace) + _msg(0x03, dtype) + _msg(0x08, layout) + (_msg(0x0B, fp) if filters else b""))

    root_at = 48
    link = bytes([1, 0, 3]) + b"var"
    root = _ohdr(_msg(0x06, link + struct.pack("<Q", 0)))
    dset_at = root_at + len(root)
    btree_at = dset_at + len(dataset_header(0))
    node_len = 8 + 16 + len(keys) * (8 + 8 * (rank + 1) + 8) + 8 + 8 * (rank + 1)
    data_at = btree_at + node_len
    node = b"TREE" + bytes([1, 0]) + struct.pack("<H", len(keys)) + struct.pack("<QQ", UNDEF, UNDEF)
    pos = data_at
    for (size, mask, offs), blob in zip(keys, blobs):
        node += struct.pack("<II", size, mask) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", pos)
        pos += len(blob)
    node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in data.shape) + struct.pack("<Q", 0)
    assert len(node) == node_len
    root = _ohdr(_msg(0x06, link + struct.pack("<Q", dset_at)))
    superblock = (hdf5_min.SIG + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, UNDEF, pos, root_at) + b"\0\0\0\0")
    assert len(superblock) == 48
    with open(path, "wb") as f:
        f.write(superblock + root + dataset_header(btree_at) + node + b"".join(blobs))


@pytest.mark.parametrize("pipeline_version", [1, 2])
@pytest.mark.parametrize("filters", [[], ["deflate"], ["shuffle", "deflate"], ["shuffle", "deflate", "fletcher32"]])
def test_chunked_dataset_through_the_netcdf4_filters(tmp_path, pipeline_version, filters):
    rng = np.random.default_rng(3)
    data = rng.normal(size=(5, 7, 3)).astype(np.float32)
    path = os.path.join(str(tmp_path), "f.h5")
    build_file(path, data, (2, 4, 3), pipeline_version, filters, skip_deflate_on=(1, 4) if "deflate" in filters else ())
    out = hdf5_min.read_hdf5(path)
    assert list(out) == ["var"] and out["var"].dtype == np.float32
    np.testing.assert_array_equal(out["var"], data)


def test_float64_and_an_unknown_filter(tmp_path):
    data = np.arange(24, dtype=np.float64).reshape(4, 6) / 7.0
    path = os.path.join(str(tmp_path), "d.h5")
    build_file(path, data, (3, 4), 2, ["shuffle", "deflate"])
    np.testing.assert_array_equal(hdf5_min.read_hdf5(path)["var"], data)
    # a filter that is not undone here (e.g. 32015 = zstd) is an error, never silently raw bytes
    with pytest.raises(NotImplementedError, match="32015"):
        hdf5_min._unfilter(b"xx", [(32015, ())], 0, 4)


def test_filter_pipeline_message_layouts():
    v2 = bytes([2, 2]) + struct.pack("<HHHI", 2, 0, 1, 4) + struct.pack("<HHHI", 1, 0, 1, 6)
    assert hdf5_min._filter_pipeline(v2) == [(2, (4,)), (1, (6,))]
    name = b"deflate\0"
    v1 = bytes([1, 1]) + b"\0" * 6 + struct.pack("<HHHH", 1, len(name), 1, 1) + name + struct.pack("<I", 9) + b"\0" * 4
    assert hdf5_min._filter_pipeline(v1) == [(1, (9,))]
    # a registered third-party filter carries its name in version 2 as well
    v2n = bytes([2, 1]) + struct.pack("<HHHH", 32015, 4, 0, 0) + b"zstd"
    assert hdf5_min._filter_pipeline(v2n) == [(32015, ())]
