#!/usr/bin/env python
"""Benchmark of the MTCLIM hot path: cell-timesteps/s (daily sim + sub-daily disaggregation).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the whole hot path over one shard: solar tables -> daily MTCLIM ->
disaggregation of every day of the record into a device ring of output tiles.  Default workload =
BASELINE.json configs[2], the configuration the metric is quoted on: CONUS 1/16 degree (482,560
cells), 1 year daily -> hourly, triangle precipitation, utc_offset, the 8 output variables of
examples/example_nc.conf, float32 outputs (reference default out_precision).  At N > 1 every rank
runs its own shard of that size (cells are independent; no collective on the data path), i.e. the
domain grows with N ("weak").

Prints ONE JSON line (rank 0).  ``value`` is measured with inputs resident in HBM; ``e2e`` is the
same metric through msb_run_host with pinned HOST buffers, H2D/D2H inside the timed region.
``--impl reference`` times the reference's OWN Python functions (the unmodified hot-path modules of
``baseline/_ref``, kind "reference") in a process pool over all host cores on a bounded sample of
the same workload; the C port's figure (oracle/, kind "port") is reported beside it and is the
fall-back when the reference install is absent.

Beyond the contract keys the line carries: ``verify`` (a sample of cells of one extra full pass, same
launch geometry as the timed ones, checked against the oracle after the timed region), ``strong``
(N > 1: ONE domain cut into contiguous cell ranges, the north_star's partition, run after the weak
pass in the same process), ``e2e.d2h_ceiling_gbs_all_ranks`` (bare pinned device-to-host copies on all
ranks at once: what the box can deliver to host memory, the yardstick of ``e2e``) and, on 8 GPUs,
``other_configs`` (BASELINE configs[3] and configs[4] -- the 10-year 3-hourly and the 30-year 15-minute
domains -- cut into cell ranges over the ranks, device-timed like the headline).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OUT_VARS = ("temp", "prec", "shortwave", "longwave", "vapor_pressure", "rel_humid",
            "air_pressure", "wind")            # examples/example_nc.conf:3
WORKLOADS = {
    # name: (synth config, note)
    "conus_16th_hourly": "CONUS 1/16 deg, 482,560 cells, 365 d daily -> hourly, triangle precip",
    "global_half_deg": "global 0.5 deg land (~67k cells), 365 d daily -> hourly, triangle precip",
    "conus_16th_10yr_3h_mix": "CONUS 1/16 deg, 3652 d daily -> 3-hourly, mix precip",
    "global_8th_30yr_15min": "global 1/8 deg land (~1M cells), 10957 d daily -> 15-minute, triangle precip "
                             "(BASELINE configs[4]: run it with --gpus 8 --shard)",
}
METRIC = "cell-timesteps/sec (daily sim + hourly disagg)"
UNIT = "cell-timesteps/s"


def algorithmic_bytes_per_cell_ts(tpd: int, n_out: int, out_bytes: int = 4, n_in: int = 4) -> float:
    """B = V_out*s_out + V_in*8/tpd."""
    return n_out * out_bytes + n_in * 8.0 / tpd


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        top = np.array(sm)[np.array(pw) >= 0.5 * max(pw)] if pw else np.array(sm)
        return {"sm_mhz": float(np.median(top)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def cpu_baseline_run(workload: str, n_days: int | None, cores: int, sample_cells: int | None,
                     steps: int = 1, warmup: int = 0):
    """Oracle (C restatement of the reference, kind='port') on all host cores over a bounded,
    seeded sample of the workload's cells.  Returns (value, seconds per step, sample text)."""
    from metsim_b200 import synth
    from oracle import oracle_lib
    if sample_cells is None:
        sample_cells = int(min(8192, max(64, 64 * cores)))
    f, params = synth.make_config(workload, max_cells=sample_cells, n_days=n_days)
    params = dict(params)
    kw = {k: f.get(k) for k in ("lat", "lon", "elev", "doy", "month", "t_min", "t_max", "prec",
                                "wind", "state_t_min", "state_t_max", "state_prec", "t_pk12", "dur12")}
    d, n = f["t_min"].shape
    tpd = 1440 // params["time_step"]
    for _ in range(warmup):
        oracle_lib.run(params, n_threads=cores, want=OUT_VARS, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_lib.run(params, n_threads=cores, want=OUT_VARS, **kw)
    dt = (time.perf_counter() - t0) / steps
    value = n * d * tpd / dt
    sample = (f"{n} evenly spaced cells of {workload} x {d} d x {tpd}/d, oracle/metsim_oracle.c "
              f"(per-cell solar_geom + mtclim + disaggregate), {cores} threads; driver/xarray/netCDF "
              f"overhead of the real `ms` excluded")
    return value, dt, sample, dict(workload=workload, cells=n, n_days=d, time_step=params["time_step"],
                                   prec_type=params["prec_type"])


_REF_JOB = {}


def _ref_worker_init(root, params, start):
    """Pool initialiser: load the UNMODIFIED reference hot-path modules (baseline/_ref) and pay
    numba's JIT of solar_geom once per process, outside the timed region."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")      # README.md:119-124 of the reference
    os.environ.setdefault("NUMBA_NUM_THREADS", "1")
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import reference_harness as rh
    rh.load_reference(root)
    rh.load_reference().physics.solar_geom(100.0, 40.0, 0.0065)
    _REF_JOB.update(rh=rh, params=params, start=start)


def _ref_worker_cell(cell):
    out = _REF_JOB["rh"].run_cell(dict(cell, start=_REF_JOB["start"]), _REF_JOB["params"])
    return float(np.nansum(out["sub"]["temp"]))        # keep the result alive, return a scalar


def reference_python_run(workload, n_days, cores, sample_cells, steps=1, warmup=0):
    """The reference's OWN functions (physics.solar_geom -> mtclim.run -> disaggregate kernels,
    loaded unmodified from baseline/_ref through oracle/reference_harness.py) per cell in a
    multiprocessing pool over all host cores -- BASELINE.md plan B.  Returns None when the
    reference install is not present."""
    import multiprocessing as mp
    from metsim_b200 import synth
    from oracle import reference_harness as rh
    root = rh.find_reference_root()
    if root is None:
        return None
    if sample_cells is None:
        sample_cells = max(256, 8 * cores)            # BASELINE.md plan B: >= 256 cells; ~0.14 s per cell-year per core
    f, params = synth.make_config(workload, max_cells=sample_cells, n_days=n_days)
    start = synth.CONFIGS[workload][2]
    d, n = f["t_min"].shape
    tpd = 1440 // params["time_step"]
    cells = []
    for c in range(n):
        cell = {k: float(f[k][c]) for k in ("lat", "lon", "elev")}
        for k in ("t_min", "t_max", "prec", "wind", "state_t_min", "state_t_max", "state_prec",
                  "t_pk12", "dur12"):
            cell[k] = np.ascontiguousarray(f[k][:, c])
        cells.append(cell)
    with mp.get_context("spawn").Pool(cores, initializer=_ref_worker_init,
                                      initargs=(root, dict(params), start)) as pool:
        chunk = max(1, n // (4 * cores))
        for _ in range(warmup):
            pool.map(_ref_worker_cell, cells[:cores], chunksize=1)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_ref_worker_cell, cells, chunksize=chunk)
        dt = (time.perf_counter() - t0) / steps
    value = n * d * tpd / dt
    sample = (f"{n} evenly spaced cells of {workload} x {d} d x {tpd}/d through the reference's own "
              f"physics.solar_geom + mtclim.run + disaggregate kernels (unmodified modules from "
              f"baseline/_ref, numba JIT warmed), multiprocessing pool of {cores}; xarray/dask/netCDF "
              f"driver overhead of the real `ms` excluded (not installable offline)")
    return value, dt, sample, dict(workload=workload, cells=n, n_days=d, time_step=params["time_step"],
                                   prec_type=params["prec_type"])


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    steps, warm = max(1, args.steps), max(0, min(args.warmup, 1))
    kind = "reference"
    res = None
    try:
        res = reference_python_run(args.workload, args.days, cores, args.sample_cells, steps, warm)
    except Exception as exc:                           # fall back to the C port, say why
        sys.stderr.write(f"reference (python) arm failed: {exc!r}; using the C port\n")
    port = cpu_baseline_run(args.workload, args.days, cores, None, steps=1, warmup=0)
    if res is None:
        kind = "port"
        res = cpu_baseline_run(args.workload, args.days, cores, args.sample_cells, steps=steps, warmup=warm)
    value, dt, sample, cfg = res
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": sample},
            "c_port": {"value": port[0], "unit": UNIT, "cores": cores, "sample": port[2]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def bind_to_gpu_numa_node(local: int):
    """Pin this rank's threads (and so the first-touch placement of its pinned buffers) to the CPUs
    nearest to its GPU.  Without it eight ranks place their host rings wherever the allocator
    likes and the device-to-host streams cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        cpus = sorted(set(cpus) & os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception as exc:
        sys.stderr.write(f"rank affinity not set: {exc!r}\n")
    return None


def verify_against_oracle(eng, f, params, step_tiles, ring, n_days, tpd, sample=512, seed=1):
    """One extra full pass with the SAME Engine, tile length and launch geometry as the timed ones;
    ``sample`` random cell columns of every variable and every step are pulled off the device ring
    and compared with the oracle run on those cells' inputs (checker only, after the timed region).
    float32 outputs must sit within float32 rounding of the oracle's float64 value plus the 1e-9
    contract; the tdew evaluation counts must be equal."""
    import torch
    from oracle import oracle_lib
    n = eng.n_cells
    rng = np.random.default_rng(seed)
    idx = np.sort(rng.choice(n, size=min(sample, n), replace=False))
    idx_t = torch.as_tensor(idx, device=eng.device)            # columns of the engine's outputs
    idx_in = idx_t if eng.order is None else eng.order.index_select(0, idx_t)   # the same cells in f
    got = {v: [] for v in OUT_VARS}
    eng.build_solar()
    daily = eng.daily(want=())
    with torch.cuda.stream(eng.stream):      # the gathers are ordered after the kernels on the engine's stream
        for k, (d0, d1) in enumerate(step_tiles):
            res = eng.disaggregate(OUT_VARS, d0, d1, torch.float32, out=ring[k & 1])
            for v in OUT_VARS:
                got[v].append(res[v][:(d1 - d0) * tpd].index_select(1, idx_t))
        n_iter = daily["n_iter"].index_select(0, idx_t)
    eng.check()                              # synchronises the stream
    got = {v: [t.cpu() for t in parts] for v, parts in got.items()}
    n_iter = n_iter.cpu().numpy()
    cpu = lambda t: t.index_select(-1, idx_in).cpu().numpy() if t.dim() else t.cpu().numpy()
    kw = {k: cpu(f[k]) for k in ("lat", "lon", "elev", "t_min", "t_max", "prec", "wind", "state_t_min",
                                 "state_t_max", "state_prec", "t_pk12", "dur12")}
    kw["doy"], kw["month"] = f["doy"].cpu().numpy(), f["month"].cpu().numpy()
    want = oracle_lib.run(params, want=OUT_VARS, **kw)
    floors = {"temp": 1.0, "prec": 1e-3, "shortwave": 1e-2, "longwave": 1.0, "vapor_pressure": 1e-3,
              "rel_humid": 1e-1, "air_pressure": 1.0, "wind": 1e-2}
    worst_rel, beyond, values = 0.0, 0, 0
    for v in OUT_VARS:
        g = torch.cat(got[v]).numpy().astype(np.float64)
        w = want[v]
        ulp = np.spacing(np.abs(w).astype(np.float32)).astype(np.float64)
        err = np.abs(g - w)
        scale = np.maximum(np.abs(w), floors[v])
        worst_rel = max(worst_rel, float(np.nanmax(err / scale)))
        beyond += int(np.sum(~(err <= 0.5000001 * ulp + 1e-9 * scale)))
        values += g.size
    zero_equal = bool(np.array_equal(torch.cat(got["prec"]).numpy() == 0, want["prec"].astype(np.float32) == 0))
    return {"cells": int(idx.size), "values": int(values), "worst_rel": worst_rel,
            "beyond_f4_rounding_plus_1e-9": beyond, "n_iter_equal": bool(np.array_equal(n_iter, want["n_iter"])),
            "n_iter_hist": np.bincount(want["n_iter"]).tolist(), "prec_zero_placement_equal": zero_equal,
            "checked": "one extra full pass, same Engine / tile length / launch geometry as the timed steps; "
                       "float32 outputs vs the oracle's float64 (bound: half a float32 ulp + 1e-9)"}


def d2h_ceiling(dev, world, nbytes=2 << 30, reps=3):
    """Bare cudaMemcpyAsync device-to-host of one pinned buffer per rank, all ranks at once."""
    import torch
    import torch.distributed as dist
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del src, dst
    return reps * nbytes * world / float(t.item()) / 1e9


# ------------------------------------------------------------------------------------------
class DevicePass:
    """One shard on this rank's GPU: inputs resident in HBM, the timed step, its launch count."""

    def __init__(self, args, dev, lat, lon, n_days, start, ts, prec_type, seed, tile_days=None, patch_rows=0):
        import torch
        from metsim_b200 import synth
        from metsim_b200.engine import Engine
        self.torch, self.dev = torch, dev
        self.n, self.n_days, self.ts, self.tpd = int(lat.size), n_days, ts, 1440 // ts
        self.params = dict(time_step=ts, prec_type=prec_type, utc_offset=True)
        self.f = synth.make_forcing_torch(lat, lon, n_days, start, dev, seed=seed)
        torch.cuda.empty_cache()
        self.eng = Engine(self.params, dev)
        order = None
        if patch_rows and patch_rows > 1:     # masked grids: engine-defined cell order (shard.patch_order)
            from metsim_b200.shard import patch_order
            order = patch_order(lat, lon, patch_rows)
        self.eng.load(order=order, **self.f)
        self.sw_table = ("window-major copy" if self.eng.use_window_major() else "plain") + \
            " (a warp's cells touch %.1f sectors of the plain table per load)" % self.eng.sw_sectors_per_load
        n, tpd = self.n, self.tpd
        if tile_days is None:    # one ring slot of the 8 float32 outputs within ~36 GB (two slots + inputs,
            # candidate planes and the 4.4 GB table stay below half of the 180 GB of HBM)
            tile_days = max(1, int(36.0e9 // (len(OUT_VARS) * tpd * n * 4)))
            if 7 * n_days * n * 8 > 60e9:       # multi-year records: inputs + daily arrays already fill HBM
                tile_days = min(tile_days, 64)
                self.eng.single_pass_budget = 0  # ... and leave no room for candidate planes
        self.tile = tile = min(tile_days, n_days)
        ld_pad = (n + 31) // 32 * 32      # rows on 128-byte boundaries, as Engine.disaggregate allocates them
        self.ring = [{v: torch.empty((tile * tpd, ld_pad), dtype=torch.float32, device=dev)[:, :n]
                      for v in OUT_VARS} for _ in range(2)]
        self.tiles = [(d0, min(n_days, d0 + tile)) for d0 in range(0, n_days, tile)]
        # kernels of mine per step: 2 solar kernels; daily single pass = fused + reduce, then for the
        # replay list 3 x (pass 1 + reduce) + pass 2 (they exit at once when the list is empty) -- the
        # two-pass form is 4 x (pass 1 + reduce) + pass 2, the same count; per msb_disaggregate call
        # thermo_knot_kernel + sw_day_kernel + prec_day_kernel over the interior days [2, D-2), plus day
        # records + disagg_kernel<thermo> / <streams> over the first / last two days of the record
        def disagg_launches(d0, d1):
            a0, b0 = max(d0, 2), min(d1, n_days - 2)
            if a0 >= b0:
                return 3
            return 3 + (3 if d0 < a0 else 0) + (3 if b0 < d1 else 0)
        self.launches_per_step = 2 + 9 + sum(disagg_launches(*t) for t in self.tiles)

    def step(self, marks=None):
        torch, eng = self.torch, self.eng
        s = eng.stream
        if marks is not None:
            marks[0].record(s)
        eng.build_solar()
        if marks is not None:
            marks[1].record(s)
        eng.daily(want=())
        if marks is not None:
            marks[2].record(s)
        for k, (d0, d1) in enumerate(self.tiles):
            eng.disaggregate(OUT_VARS, d0, d1, torch.float32, out=self.ring[k & 1])
        if marks is not None:
            marks[3].record(s)

    def timed(self, steps, warmup, barrier, sampler=None):
        """(ms per step of this rank, breakdown [solar, daily, disagg] in ms)"""
        torch, eng = self.torch, self.eng
        ev = lambda: torch.cuda.Event(enable_timing=True)
        for _ in range(max(3, warmup)):
            self.step()
        eng.check()
        barrier()
        if sampler is not None:
            sampler.start()
        marks = [[ev() for _ in range(4)] for _ in range(steps)]
        t0, t1 = ev(), ev()
        t0.record(eng.stream)
        for k in range(steps):
            self.step(marks[k])
        t1.record(eng.stream)
        barrier()
        ms = t0.elapsed_time(t1) / steps
        br = np.array([[m[i].elapsed_time(m[i + 1]) for i in range(3)] for m in marks]).mean(0)
        return ms, br

    def release(self):
        self.ring = None
        self.eng = None
        self.f = None
        self.torch.cuda.empty_cache()


def e2e_pass(dp_f, params, n, n_days, tpd, local, steps, barrier, reduce_max, consume=True, world=1):
    """The same step through msb_run_host_ctx (C ABI) with pinned HOST buffers: H2D of the inputs and
    D2H of every output tile inside the timed region, every tile handed to a consumer (the native
    thread pool of csrc/sink.cu: reads every row, scatters / byte-swaps it into its staging --
    what the streaming NetCDF writer does short of the disk) while the next tile computes."""
    import torch
    from metsim_b200.engine import HostContext, run_host
    from metsim_b200.ncwriter import NativeSink
    pin = lambda t: t.cpu().pin_memory().numpy()
    host = {k: pin(v) for k, v in dp_f.items()}
    tile_e2e = max(1, min(n_days, int((2 << 30) // max(1, len(OUT_VARS) * tpd * n * 4))))
    hring = {v: torch.empty((2 * tile_e2e * tpd, n), dtype=torch.float32).pin_memory().numpy()
             for v in OUT_VARS}
    kw = dict(lat=host["lat"], lon=host["lon"], elev=host["elev"], doy=host["doy"],
              month=host["month"], t_min=host["t_min"], t_max=host["t_max"], prec=host["prec"],
              wind=host["wind"], state_t_min=host["state_t_min"], state_t_max=host["state_t_max"],
              state_prec=host["state_prec"], t_pk12=host["t_pk12"], dur12=host["dur12"])
    # the host's cores shared by the ranks (the calling thread sleeps in a blocking event wait between tiles)
    n_thr = max(2, min(32, (os.cpu_count() or 2) // world - (1 if world == 1 else 0)))
    sink = NativeSink(n_cells=n, elem_bytes=4, byteswap=True, n_threads=n_thr, fd=-1) if consume else None
    consumed = [0]

    def on_tile(k, d0, d1, slot):
        rows = (d1 - d0) * tpd
        for v in OUT_VARS:
            sink.write(hring[v][slot * tile_e2e * tpd: slot * tile_e2e * tpd + rows])
            consumed[0] += rows * n * 4
    ctx = HostContext(local)
    call = lambda: run_host(params, out_vars=OUT_VARS, tile_days=tile_e2e, out=hring, ring=True, ctx=ctx,
                            on_tile=on_tile if consume else None, **kw)
    call()  # warm-up: arena allocation, page faults
    barrier()
    consumed[0] = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    barrier()
    dt = reduce_max((time.perf_counter() - t0) / steps)
    # one step without the consumer beside it: what the copies alone take (tiles overwritten unread)
    copy_only = None
    if consume:
        t0 = time.perf_counter()
        run_host(params, out_vars=OUT_VARS, tile_days=tile_e2e, out=hring, ring=True, ctx=ctx, **kw)
        barrier()
        copy_only = reduce_max(time.perf_counter() - t0)
    ctx.close()
    if sink is not None:
        sink.close()
    h2d = sum(host[k].nbytes for k in ("lon", "elev", "doy", "month", "t_min", "t_max", "prec",
                                        "wind", "state_t_min", "state_t_max", "state_prec",
                                        "t_pk12", "dur12")) + n * 4
    d2h = len(OUT_VARS) * n_days * tpd * n * 4
    return dt, int(h2d), int(d2h), int(consumed[0] // max(1, steps)), n_thr, tile_e2e, copy_only


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--workload", default="conus_16th_hourly", choices=sorted(WORKLOADS))
    ap.add_argument("--cells", type=int, default=None, help="truncate the grid (debugging only)")
    ap.add_argument("--days", type=int, default=None, help="truncate the record (debugging only)")
    ap.add_argument("--tile-days", type=int, default=None,
                    help="days per msb_disaggregate call (default: as many as keep one ring slot <= 36 GB, at most 64 "
                         "for the headline grid)")
    ap.add_argument("--sample-cells", type=int, default=None)
    ap.add_argument("--shard", action="store_true",
                    help="make the strong-scaling partition the HEADLINE: the ranks share ONE domain, each takes a "
                         "contiguous cell range (metsim_b200.shard.cell_range).  Default: one full domain per rank "
                         "(weak) as the headline, the partitioned domain reported beside it under 'strong'")
    ap.add_argument("--emulate-shard", default=None, metavar="R/W",
                    help="debugging: on ONE GPU, take the cell range rank R of W would get with --shard")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--other-configs", default="auto", choices=("auto", "on", "off"),
                    help="after the headline, also time BASELINE configs[3] (CONUS 10 years -> 3-hourly, mix) and "
                         "configs[4] (global 1/8 deg, 30 years -> 15-minute) cut into cell ranges over the ranks; "
                         "auto = on 8 GPUs (what BASELINE.json quotes them on)")
    ap.add_argument("--no-consumer", action="store_true", help="e2e without the native tile consumer")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--verify-cells", type=int, default=512)
    ap.add_argument("--patch-rows", type=int, default=None,
                    help="masked grids: latitude rows per patch of the engine-defined cell order "
                         "(metsim_b200.shard.patch_order; default 4 on the masked global grids, 1 = caller's order)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from metsim_b200 import build as msb_build, synth
    from metsim_b200.shard import cell_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(x)]
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        return [float(e.item()) for e in every]

    grid, n_days, start, ts, prec_type = synth.CONFIGS[args.workload]
    n_days = args.days or n_days
    rng = np.random.default_rng(synth.SEED)
    lat_all, lon_all = synth.grid_cells(grid, rng)
    if args.cells:
        idx = np.linspace(0, lat_all.size - 1, args.cells).astype(np.int64)
        lat_all, lon_all = lat_all[idx], lon_all[idx]
    if args.emulate_shard and world == 1:
        r_, w_ = (int(x) for x in args.emulate_shard.split("/"))
        lo, hi = cell_range(lat_all.size, r_, w_)
        lat_all, lon_all = lat_all[lo:hi], lon_all[lo:hi]
    n_domain = lat_all.size
    tpd = 1440 // ts
    headline_strong = args.shard and world > 1
    if headline_strong:      # latitude bands: contiguous ranges of the row-major enumeration
        lo, hi = cell_range(n_domain, rank, world)
        lat, lon = lat_all[lo:hi], lon_all[lo:hi]
    else:
        lat, lon = lat_all, lon_all

    # ---- headline pass: every rank owns a shard (different seed = different part of a larger domain)
    if args.patch_rows is None:
        args.patch_rows = 4 if grid.startswith("global") else 1
    dp = DevicePass(args, dev, lat, lon, n_days, start, ts, prec_type, synth.SEED + rank, args.tile_days,
                    patch_rows=args.patch_rows)
    n, tile = dp.n, dp.tile
    sampler = ClockSampler(local) if rank == 0 else None
    my_ms, br = dp.timed(args.steps, args.warmup, barrier, sampler)
    clocks = sampler.stop() if rank == 0 else None
    rank_ms = gather(my_ms)
    ms_per_step = max(rank_ms)
    cell_ts_rank = n * n_days * tpd
    total_cell_ts = n_domain * n_days * tpd if headline_strong else world * cell_ts_rank
    value = total_cell_ts / (ms_per_step * 1e-3)
    n_tiles = len(dp.tiles)
    disagg_ms_per_launch = br[2] / n_tiles
    bytes_per_cts = algorithmic_bytes_per_cell_ts(tpd, len(OUT_VARS))
    bytes_per_launch = bytes_per_cts * cell_ts_rank / n_tiles
    achieved = bytes_per_launch / (disagg_ms_per_launch * 1e-3) / 1e9
    peak, peak_src = measured_hbm_peak()
    # DRAM traffic of the call from the committed ncu capture -- only if it was taken from THIS build
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "disagg_traffic.json")
    src_hash = msb_build.source_hash()
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("source_hash") == src_hash:
                traffic = float(tj["dram_bytes_per_cell_timestep"]) * cell_ts_rank / n_tiles
            else:
                traffic_note = (f"profiles/disagg_traffic.json was captured from sources {tj.get('source_hash')}, "
                                f"this build is {src_hash}: not reported")
        except Exception as exc:
            traffic_note = f"profiles/disagg_traffic.json unreadable: {exc!r}"

    # ---- parity of the timed configuration (checker only; after the timed region) ----
    verify = None
    if not args.no_verify and rank == 0:
        try:
            verify = verify_against_oracle(dp.eng, dp.f, dp.params, dp.tiles, dp.ring, n_days, tpd,
                                           sample=args.verify_cells)
        except Exception as exc:
            verify = {"error": repr(exc)}
    barrier()

    # ---- end to end through the C ABI with pinned host buffers ----
    e2e, ceiling = None, None
    launches_per_step = dp.launches_per_step
    if not args.no_e2e:
        f_keep, params_keep = dp.f, dp.params
        dp.ring = None
        torch.cuda.empty_cache()
        ceiling = d2h_ceiling(dev, world)
        e_steps = args.e2e_steps or max(1, min(args.steps, 3))
        dt, h2d, d2h, consumed, n_thr, tile_e2e, copy_only = e2e_pass(
            f_keep, params_keep, n, n_days, tpd, local, e_steps, barrier, reduce_max,
            consume=not args.no_consumer, world=world)
        e2e = {"value": total_cell_ts / dt, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt * 1e3, "steps": e_steps,
               "ms_per_step_copies_only": copy_only * 1e3 if copy_only else None,
               "d2h_gbs_all_ranks": world * d2h / dt / 1e9,
               "d2h_ceiling_gbs_all_ranks": ceiling,
               "fraction_of_d2h_ceiling": (world * d2h / dt / 1e9) / ceiling if ceiling else None,
               "fraction_of_d2h_ceiling_copies_only": ((world * d2h / copy_only / 1e9) / ceiling
                                                       if ceiling and copy_only else None),
               "consumer": (f"native sink (csrc/sink.cu), {n_thr} threads per rank: every tile read, byte-swapped "
                            f"and scattered into staging while the next tile computes; {consumed} bytes per step"
                            if not args.no_consumer else "none"),
               "tile_days": tile_e2e, "rank_cpu_affinity": affinity,
               "timer": "host wall clock around the blocking C-ABI call, max over ranks",
               "api": "msb_run_host_ctx (C ABI), pinned host inputs, outputs streamed D2H into a pinned "
                      "2-tile host ring and handed to the consumer tile by tile"}
    dp.release()

    # ---- the north_star's partition: ONE domain cut into contiguous cell ranges (N > 1) ----
    strong = None
    if world > 1 and not headline_strong and not args.no_strong:
        lo, hi = cell_range(n_domain, rank, world)
        sp = DevicePass(args, dev, lat_all[lo:hi], lon_all[lo:hi], n_days, start, ts, prec_type,
                        synth.SEED + 100 + rank, None, patch_rows=args.patch_rows)
        s_ms, s_br = sp.timed(args.steps, args.warmup, barrier)
        s_rank = gather(s_ms)
        s_max = max(s_rank)
        dom_cts = n_domain * n_days * tpd
        strong = {"value": dom_cts / (s_max * 1e-3), "unit": UNIT, "ms_per_step": s_max,
                  "rank_ms_per_step": s_rank, "cells_per_gpu": int(hi - lo), "tile_days": sp.tile,
                  # one GPU running the whole domain = this run's own weak pass (same domain size per rank)
                  "efficiency_vs_n1": (ms_per_step / s_max) / world,
                  "n1_ms_per_step_used": ms_per_step,
                  "sharding": f"one domain of {n_domain} cells cut into {world} contiguous cell ranges, no collective"}
        if not args.no_e2e:
            e_steps = args.e2e_steps or max(1, min(args.steps, 3))
            f_keep, params_keep = sp.f, sp.params
            sp.ring = None
            torch.cuda.empty_cache()
            dt, h2d, d2h, consumed, n_thr, tile_e2e, copy_only = e2e_pass(
                f_keep, params_keep, sp.n, n_days, tpd, local, e_steps, barrier, reduce_max,
                consume=not args.no_consumer, world=world)
            d2h_all = sum(gather(d2h))
            strong["e2e"] = {"value": dom_cts / dt, "unit": UNIT, "ms_per_step": dt * 1e3,
                             "ms_per_step_copies_only": copy_only * 1e3 if copy_only else None,
                             "d2h_gbs_all_ranks": d2h_all / dt / 1e9,
                             "fraction_of_d2h_ceiling": (d2h_all / dt / 1e9) / ceiling if ceiling else None,
                             "efficiency_vs_n1": (e2e["ms_per_step"] / (dt * 1e3)) / world if e2e else None}
        sp.release()

    # ---- BASELINE configs[3] and configs[4]: quoted on 8 B200, cell-chunk sharded ----
    others = None
    if (args.other_configs == "on" or (args.other_configs == "auto" and world == 8)) \
            and args.workload == "conus_16th_hourly" and not args.cells and not args.days:
        others = {}
        for name in ("conus_16th_10yr_3h_mix", "global_8th_30yr_15min"):
            g2, d2, start2, ts2, pt2 = synth.CONFIGS[name]
            la2, lo2 = synth.grid_cells(g2, np.random.default_rng(synth.SEED))
            lo_c, hi_c = cell_range(la2.size, rank, world)
            op, err = None, None
            try:                                 # allocation is the part that can fail (memory): every rank
                op = DevicePass(args, dev, la2[lo_c:hi_c], lo2[lo_c:hi_c], d2, start2, ts2, pt2,   # learns of it
                                synth.SEED + 200 + rank, None, patch_rows=4 if g2.startswith("global") else 1)
            except Exception as exc:             # before anyone enters a collective
                err = repr(exc)
                torch.cuda.empty_cache()
            if reduce_max(1.0 if err else 0.0) > 0:
                others[name] = {"error": err or "another rank could not set the shard up"}
                if op is not None:
                    op.release()
                continue
            o_ms, o_br = op.timed(3, 3, barrier)
            o_rank = gather(o_ms)
            tpd2 = 1440 // ts2
            cts = la2.size * d2 * tpd2
            b2 = algorithmic_bytes_per_cell_ts(tpd2, len(OUT_VARS))
            others[name] = {"value": cts / (max(o_rank) * 1e-3), "unit": UNIT, "ms_per_step": max(o_rank),
                            "rank_ms_per_step": o_rank, "cells": int(la2.size), "cells_per_gpu": int(hi_c - lo_c),
                            "n_days": d2, "time_step_min": ts2, "prec_type": pt2, "tile_days": op.tile,
                            "shortwave_table": op.sw_table,
                            "daily_form": ("single pass" if op.eng._daily and op.eng._daily[5][0] is not None
                                           else "two pass (the candidate planes do not fit the budget beside the record)"),
                            "disaggregate_frac_of_hbm_rank0":
                                b2 * (hi_c - lo_c) * d2 * tpd2 / (float(o_br[2]) * 1e-3) / 1e9 / peak,
                            "breakdown_ms_rank0": {"solar_tables": float(o_br[0]), "daily_mtclim": float(o_br[1]),
                                                   "disaggregate": float(o_br[2])},
                            "sharding": f"one domain of {la2.size} cells cut into {world} contiguous cell ranges, "
                                        "no collective"}
            op.release()
            del op
            barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:      # reported baseline: rank 0 at N = 1 only
        cores = os.cpu_count() or 1
        res, kind = None, "reference"
        try:
            res = reference_python_run(args.workload, args.days, cores, args.sample_cells, 1, 1)
        except Exception as exc:
            sys.stderr.write(f"reference (python) baseline failed: {exc!r}; using the C port\n")
        pv, _, psample, _ = cpu_baseline_run(args.workload, args.days, cores, None)
        if res is None:
            kind, res = "port", (pv, 0.0, psample, None)
        cpu = {"value": res[0], "unit": UNIT, "cores": cores, "kind": kind, "sample": res[2],
               "c_port_value": pv, "c_port_sample": psample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if headline_strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "description": WORKLOADS[args.workload],
                       "cells_per_gpu": int(n), "n_days": int(n_days), "time_step_min": ts,
                       "prec_type": prec_type, "utc_offset": True, "out_vars": list(OUT_VARS),
                       "out_precision": "f4", "tile_days": tile, "shortwave_table": dp.sw_table,
                       "cell_order": ("caller's (latitude-row-major)" if args.patch_rows <= 1 else
                                      f"engine-defined patches of {args.patch_rows} latitude rows (shard.patch_order)"),
                       "launch_geometry": "library defaults: 16 knot-days / 16 output days per thread, daily chunks of "
                                          "ceil(n_days / (2^21 / cells + 1)) days, single-pass daily form",
                       "sharding": (f"one domain of {n_domain} cells cut into {world} contiguous cell ranges, no collective"
                                    if headline_strong else f"cell-chunks x{world}, no collective"),
                       "l2": "each tile writes %.1f GB (>> 126 MB L2) into a 2-tile ring; inputs %.1f GB"
                             % (len(OUT_VARS) * tile * tpd * n * 4 / 1e9, 4 * n_days * n * 8 / 1e9)},
            "breakdown_ms": {"solar_tables": float(br[0]), "daily_mtclim": float(br[1]),
                             "disaggregate": float(br[2])},
            # the metric with and without the (time-independent) solar-geometry set-up
            "rank_ms_per_step": rank_ms,      # device time of every rank (value uses the maximum)
            "value_excl_solar_tables": total_cell_ts / max(1e-9, (ms_per_step - float(br[0])) * 1e-3),
            "roofline": {"bound": "hbm", "kernel": "msb_disaggregate = thermo_knot_kernel + sw_day_kernel + prec_day_kernel (+ "
                                                   "disagg_prepare_kernel and disagg_kernel on the record's first / last "
                                                   "two days), timed together per call (--tile-days days)",
                         "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_note": traffic_note,
                         "algorithmic_bytes_per_cell_timestep": bytes_per_cts,
                         "bytes_per_launch": bytes_per_launch, "ms_per_launch": float(disagg_ms_per_launch),
                         "whole_step_frac": bytes_per_cts * cell_ts_rank / (ms_per_step * 1e-3) / 1e9 / peak},
            "verify": verify, "strong": strong, "other_configs": others,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "build": {"library": "metsim_b200/csrc/libmetsim_b200.so", "source_hash": src_hash,
                      "built_from_these_sources": msb_build.built_hash() == src_hash},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
