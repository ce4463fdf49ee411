"""Cell-chunk sharding of a domain across GPUs.

Cells never read each other's data (metsim/metsim.py:477-493), so the domain is cut into contiguous
ranges of the latitude-row-major cell enumeration -- latitude bands.  Each rank needs solar tables
only for its own latitudes and every ``[time][cell_chunk]`` array is contiguous on its device.
There is no collective on the data path; :func:`gather_to_root` is the optional host gather for
callers that want the full field on one rank (it uses whatever backend the process group has:
NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

# per-cell arrays ([N]) and cell-fastest 2-D arrays ([rows][N]) of the engine's input dict
CELL_KEYS = ("lat", "lon", "elev")
ROWS_BY_CELL_KEYS = ("t_min", "t_max", "prec", "wind", "state_t_min", "state_t_max", "state_prec",
                     "t_pk12", "dur12", "t_end",
                     # forcing columns of the 'passthrough' method (methods/passthrough.py:28-46)
                     "given_shortwave", "given_vapor_pressure", "given_tskc", "given_longwave")
SHARED_KEYS = ("doy", "month")


def cell_range(n_cells: int, rank: int, world: int, align: int = 32) -> tuple[int, int]:
    """[lo, hi) of rank's cells: equal contiguous chunks, boundaries rounded to ``align`` cells
    (a warp of lanes) so no 128-byte line of a ``[time][cell]`` array is split between ranks."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    if n_cells < 0:
        raise ValueError("n_cells must be >= 0")

    def edge(r):
        e = (n_cells * r) // world
        if align > 1 and 0 < r < world:
            e = min(n_cells, ((e + align // 2) // align) * align)
        return e

    return edge(rank), edge(rank + 1)


def patch_order(lat, lon, rows: int = 4) -> np.ndarray:
    """Engine cell order for MASKED grids: a permutation ``order`` (engine position -> position in
    the caller's latitude-row-major enumeration) that makes every run of 32 consecutive cells -- one
    warp -- compact in longitude.

    The kernels put consecutive cells on the lanes of a warp.  On a dense grid a warp then holds 32
    neighbouring longitudes of one latitude row (2 degrees at 1/16 degree) and the lanes share their
    daily schedule: sunrise / sunset knots, the step at which the longitude roll falls, the table
    rows they read.  With a one-in-three land mask on a 0.5 degree grid the same 32 cells span ~48
    degrees = 3.2 hours of local time, so per-lane state updates run 4-5 times per warp and the
    lanes' step ranges no longer coincide (profiles/README.md).  Taking the cells of ``rows``
    neighbouring latitude rows together, ordered by longitude, divides that span by ``rows`` at the
    price of ``rows`` latitudes (table rows) per warp.  ``rows = 1`` is the identity for
    row-major input.  Results come out in engine order; ``order`` maps them back
    (``out_user[..., order] = out_engine``) or goes into the writer's ``cell_index``."""
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    if lat.shape != lon.shape or lat.ndim != 1:
        raise ValueError("lat and lon must be 1-D arrays of the same length")
    rows = max(1, int(rows))
    ulat, row = np.unique(lat, return_inverse=True)
    band = row // rows
    # bands in latitude order; inside a band by longitude, then latitude (stable: ties keep input order)
    return np.lexsort((row, lon, band)).astype(np.int64)


def shard_inputs(inputs: dict, rank: int, world: int, align: int = 32) -> dict:
    """Slice an engine input dict (numpy arrays or torch tensors) to this rank's cells."""
    n = len(inputs["lat"])
    lo, hi = cell_range(n, rank, world, align)
    out = {}
    for k, v in inputs.items():
        if v is None:
            out[k] = None
        elif k in CELL_KEYS:
            out[k] = v[lo:hi]
        elif k in ROWS_BY_CELL_KEYS:
            out[k] = v[:, lo:hi]
        elif k in SHARED_KEYS:
            out[k] = v
        else:
            # an unknown key must not reach a rank at full domain width by accident
            shape = tuple(getattr(v, "shape", ()))
            if len(shape) >= 1 and shape[-1] == n and n > 1:
                raise KeyError(f"shard_inputs: do not know how to slice {k!r} of shape {shape}; "
                               "add it to CELL_KEYS / ROWS_BY_CELL_KEYS")
            out[k] = v
    return out


def gather_to_root(local: np.ndarray, n_cells: int, align: int = 32, group=None):
    """Host gather of one ``[rows][cells_of_rank]`` output array to rank 0 (returns the full
    ``[rows][n_cells]`` array there, ``None`` elsewhere)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    rows = local.shape[0]
    widths = [cell_range(n_cells, r, world, align) for r in range(world)]
    wmax = max(hi - lo for lo, hi in widths)
    # equal-size buffers (gloo and NCCL gathers want identical shapes): pad the cell axis
    buf = torch.zeros((rows, wmax), dtype=torch.from_numpy(np.empty(0, local.dtype)).dtype, device=dev)
    lo, hi = widths[rank]
    buf[:, :hi - lo] = torch.as_tensor(np.ascontiguousarray(local)).to(dev)
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    if backend == "nccl":   # NCCL has no gather-to-one in older builds: all_gather is equivalent
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
    else:
        dist.gather(buf, parts, dst=0, group=group)
    if rank != 0:
        return None
    full = np.empty((rows, n_cells), dtype=local.dtype)
    for (lo, hi), p in zip(widths, parts):
        full[:, lo:hi] = p[:, :hi - lo].cpu().numpy()
    return full
