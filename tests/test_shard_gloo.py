"""Cell-chunk sharding (metsim_b200/shard.py) with a real world_size-2 process group (gloo, CPU).

The N>1 data path has no collective: every rank runs the hot path on its own contiguous cell range.
Here the per-rank compute is the oracle (no GPU in this container); what is under test is the
partitioning, the input slicing and the optional host gather -- the sharded result must be
bit-identical to the unsharded one."""
import os
import socket

import numpy as np
import pytest

import msb_testutil as util
from metsim_b200 import shard


def test_cell_range_partitions_exactly():
    for n in (0, 1, 31, 32, 33, 81, 1000, 482560, 67421):
        for world in (1, 2, 3, 4, 8):
            edges = [shard.cell_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and a <= b
            if n >= 64 * world:
                assert all(lo % 32 == 0 for lo, _ in edges)
                sizes = [hi - lo for lo, hi in edges]
                assert max(sizes) - min(sizes) <= 64
    with pytest.raises(ValueError):
        shard.cell_range(10, 2, 2)


def test_shard_inputs_slices_passthrough_columns_and_rejects_unknown_wide_arrays():
    """The 'passthrough' forcing columns are [D][N] like every other daily input: a rank must get its
    own cells, never the domain-wide array (silent wrong row stride through run_host)."""
    n, d = 200, 5
    rng = np.random.default_rng(0)
    inputs = dict(lat=np.arange(n, dtype=float), lon=np.zeros(n), elev=np.zeros(n),
                  doy=np.arange(1, d + 1), month=np.ones(d, int),
                  t_min=rng.random((d, n)), t_max=rng.random((d, n)) + 2, prec=rng.random((d, n)),
                  wind=None, state_t_min=rng.random((90, n)), state_t_max=rng.random((90, n)),
                  state_prec=rng.random((90, n)), t_pk12=rng.random((12, n)), dur12=rng.random((12, n)),
                  t_end=rng.random((2, n)))
    for k in ("given_shortwave", "given_vapor_pressure", "given_tskc", "given_longwave"):
        inputs[k] = rng.random((d, n))
    parts = [shard.shard_inputs(inputs, r, 3) for r in range(3)]
    for k, v in inputs.items():
        if v is None or k in ("doy", "month"):
            continue
        np.testing.assert_array_equal(np.concatenate([p[k] for p in parts], axis=-1), v, err_msg=k)
    with pytest.raises(KeyError):
        shard.shard_inputs(dict(inputs, some_new_column=rng.random((d, n))), 0, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from metsim_b200 import synth
    from oracle import oracle_lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        f, params = synth.make_config("conus_16th_hourly", max_cells=150, n_days=12)
        mine = shard.shard_inputs(f, rank, world)
        lo, hi = shard.cell_range(150, rank, world)
        assert mine["t_min"].shape == (12, hi - lo) and mine["lat"].shape == (hi - lo,)
        got = oracle_lib.run(params, n_threads=1, **util.oracle_kwargs(mine))
        full = {v: shard.gather_to_root(got[v], 150) for v in ("temp", "prec", "shortwave")}
        if rank == 0:
            want = oracle_lib.run(params, n_threads=1, **util.oracle_kwargs(f))
            ok = all(np.array_equal(full[v], want[v], equal_nan=True) for v in full)
            q.put(("ok" if ok else "mismatch", {v: full[v].shape for v in full}))
        else:
            assert all(v is None for v in full.values())
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_run_matches_unsharded():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    status, shapes = q.get(timeout=10)
    assert status == "ok"
    assert shapes["temp"] == (12 * 24, 150)


def test_patch_order_is_a_permutation_that_makes_warps_compact():
    """Masked 0.5 degree grid, one cell in three: 32 consecutive cells of the row-major order span
    ~48 degrees of longitude; in patch order (4 latitude rows together) a quarter of that."""
    from metsim_b200 import synth
    lat, lon = synth.grid_cells("global0.5", np.random.default_rng(synth.SEED))
    for rows in (1, 2, 4, 8):
        order = shard.patch_order(lat, lon, rows)
        assert np.array_equal(np.sort(order), np.arange(lat.size))
        la, lo = lat[order], lon[order]
        n_warps = lat.size // 32
        span = np.ptp(lo[:n_warps * 32].reshape(n_warps, 32), axis=1)
        lats = np.array([np.unique(w).size for w in la[:n_warps * 32].reshape(n_warps, 32)])
        if rows == 1:
            assert np.array_equal(order, np.arange(lat.size))          # row-major input stays as it is
            base = np.median(span)
            assert base > 40.0
        else:
            assert np.median(span) < 1.3 * base / rows and np.median(lats) <= rows and lats.max() <= 2 * rows
    with pytest.raises(ValueError):
        shard.patch_order(np.zeros(3), np.zeros(4))
