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
``--impl reference`` times the CPU restatement of the reference (oracle/, kind "port") on all host
cores on a bounded sample of the same workload.
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
        sample_cells = 8 * cores                      # ~0.14 s per cell-year hourly per core
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


# ------------------------------------------------------------------------------------------
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
                    help="days per msb_disaggregate call (default: as many as keep one ring slot <= 24 GB, at most 64 "
                         "for the headline grid)")
    ap.add_argument("--sample-cells", type=int, default=None)
    ap.add_argument("--shard", action="store_true",
                    help="strong scaling: the ranks share ONE domain, each takes a contiguous cell range "
                         "(metsim_b200.shard.cell_range); default is one full domain per rank (weak)")
    ap.add_argument("--emulate-shard", default=None, metavar="R/W",
                    help="debugging: on ONE GPU, take the cell range rank R of W would get with --shard")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from metsim_b200 import synth
    from metsim_b200.engine import Engine, run_host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    grid, n_days, start, ts, prec_type = synth.CONFIGS[args.workload]
    n_days = args.days or n_days
    rng = np.random.default_rng(synth.SEED)
    lat, lon = synth.grid_cells(grid, rng)
    if args.cells:
        idx = np.linspace(0, lat.size - 1, args.cells).astype(np.int64)
        lat, lon = lat[idx], lon[idx]
    n_domain = lat.size
    if args.emulate_shard and world == 1:
        from metsim_b200.shard import cell_range
        r_, w_ = (int(x) for x in args.emulate_shard.split("/"))
        lo, hi = cell_range(n_domain, r_, w_)
        lat, lon = lat[lo:hi], lon[lo:hi]
        n_domain = lat.size
    if args.shard and world > 1:      # latitude bands: contiguous ranges of the row-major enumeration
        from metsim_b200.shard import cell_range
        lo, hi = cell_range(n_domain, rank, world)
        lat, lon = lat[lo:hi], lon[lo:hi]
    n = lat.size
    tpd = 1440 // ts
    params = dict(time_step=ts, prec_type=prec_type, utc_offset=True)
    # every rank owns a shard of the same shape (different seed = different part of the domain)
    f = synth.make_forcing_torch(lat, lon, n_days, start, dev, seed=synth.SEED + rank)
    torch.cuda.empty_cache()
    eng = Engine(params, dev)
    eng.load(**f)
    if args.tile_days is None:    # one ring slot of the 8 float32 outputs within ~24 GB
        args.tile_days = max(1, int(24.0e9 // (len(OUT_VARS) * tpd * n * 4)))
        if 7 * n_days * n * 8 > 60e9:       # multi-year records: inputs + daily arrays already fill HBM
            args.tile_days = min(args.tile_days, 64)
    tile = min(args.tile_days, n_days)
    ld_pad = (n + 31) // 32 * 32      # rows on 128-byte boundaries, as Engine.disaggregate allocates them
    ring = [{v: torch.empty((tile * tpd, ld_pad), dtype=torch.float32, device=dev)[:, :n] for v in OUT_VARS}
            for _ in range(2)]
    n_tiles = -(-n_days // tile)
    rounds = 4  # ceil(max_iter 24 / 6 evaluations per round), daily.cu
    # 2 solar kernels, daily passes (pass 1 + reduce per round, pass 2), and per msb_disaggregate
    # call: thermo_knot_kernel + streams_day_kernel over the interior days [2, D-2), plus day
    # records + disagg_kernel<thermo> / <streams> over the first / last two days of the record
    def disagg_launches(d0, d1):
        a0, b0 = max(d0, 2), min(d1, n_days - 2)
        if a0 >= b0:
            return 3
        return 2 + (3 if d0 < a0 else 0) + (3 if b0 < d1 else 0)
    launches_per_step = 2 + 2 * rounds + 1 + sum(
        disagg_launches(k * tile, min(n_days, (k + 1) * tile)) for k in range(n_tiles))
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(marks=None):
        s = eng.stream
        if marks is not None:
            marks[0].record(s)
        eng.build_solar()
        if marks is not None:
            marks[1].record(s)
        eng.daily(want=())
        if marks is not None:
            marks[2].record(s)
        for k in range(n_tiles):
            d0, d1 = k * tile, min(n_days, (k + 1) * tile)
            eng.disaggregate(OUT_VARS, d0, d1, torch.float32, out=ring[k & 1])
        if marks is not None:
            marks[3].record(s)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(3, args.warmup)):
        step()
    eng.check()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    marks = [[ev() for _ in range(4)] for _ in range(args.steps)]
    t_start, t_end = ev(), ev()
    t_start.record(eng.stream)
    for k in range(args.steps):
        step(marks[k])
    t_end.record(eng.stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_start.elapsed_time(t_end)
    ms = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    rank_ms = [total_ms / args.steps]
    if world > 1:
        every = [torch.zeros_like(ms) for _ in range(world)]
        dist.all_gather(every, ms)
        rank_ms = [float(t.item()) / args.steps for t in every]
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    cell_ts_rank = n * n_days * tpd
    strong = args.shard and world > 1
    total_cell_ts = n_domain * n_days * tpd if strong else world * cell_ts_rank
    value = total_cell_ts / (ms_per_step * 1e-3)
    br = np.array([[m[i].elapsed_time(m[i + 1]) for i in range(3)] for m in marks]).mean(0)
    disagg_ms_per_launch = br[2] / n_tiles
    bytes_per_cts = algorithmic_bytes_per_cell_ts(tpd, len(OUT_VARS))
    bytes_per_launch = bytes_per_cts * cell_ts_rank / n_tiles
    achieved = bytes_per_launch / (disagg_ms_per_launch * 1e-3) / 1e9
    peak, peak_src = measured_hbm_peak()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "disagg_traffic.json")
    if os.path.exists(tpath):   # measured once per round with ncu --set full; scaled to this launch size
        try:
            per_cts = float(json.load(open(tpath))["dram_bytes_per_cell_timestep"])
            traffic = per_cts * cell_ts_rank / n_tiles
        except Exception:
            traffic = None

    # ---- end to end through the C ABI with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        pin = lambda t: t.cpu().pin_memory().numpy()
        host = {k: pin(v) for k, v in f.items()}
        tile_e2e = max(1, min(n_days, int((2 << 30) // max(1, len(OUT_VARS) * tpd * n * 4))))
        hring = {v: torch.empty((2 * tile_e2e * tpd, n), dtype=torch.float32).pin_memory().numpy()
                 for v in OUT_VARS}
        kw = dict(lat=host["lat"], lon=host["lon"], elev=host["elev"], doy=host["doy"],
                  month=host["month"], t_min=host["t_min"], t_max=host["t_max"], prec=host["prec"],
                  wind=host["wind"], state_t_min=host["state_t_min"], state_t_max=host["state_t_max"],
                  state_prec=host["state_prec"], t_pk12=host["t_pk12"], dur12=host["dur12"])
        e_steps = args.e2e_steps or max(1, min(args.steps, 3))
        del ring
        torch.cuda.empty_cache()
        call = lambda: run_host(params, out_vars=OUT_VARS, device=local, tile_days=tile_e2e,
                                out=hring, ring=True, **kw)
        call()  # warm-up: arena allocation, page faults
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            call()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = sum(host[k].nbytes for k in ("lon", "elev", "doy", "month", "t_min", "t_max", "prec",
                                            "wind", "state_t_min", "state_t_max", "state_prec",
                                            "t_pk12", "dur12")) + n * 4
        d2h = len(OUT_VARS) * n_days * tpd * n * 4
        e2e = {"value": total_cell_ts / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": float(dt.item()) * 1e3, "steps": e_steps,
               "timer": "host wall clock around the blocking C-ABI call, max over ranks",
               "api": "msb_run_host (C ABI), pinned host inputs, outputs streamed D2H into a pinned "
                      "2-tile host ring (what a streaming netCDF writer would drain)"}

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
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": WORKLOADS[args.workload],
                       "cells_per_gpu": int(n), "n_days": int(n_days), "time_step_min": ts,
                       "prec_type": prec_type, "utc_offset": True, "out_vars": list(OUT_VARS),
                       "out_precision": "f4", "tile_days": tile,
                       "sharding": (f"one domain of {n_domain} cells cut into {world} contiguous cell ranges, no collective"
                                    if strong else f"cell-chunks x{world}, no collective"),
                       "l2": "each tile writes %.1f GB (>> 126 MB L2) into a 2-tile ring; inputs %.1f GB"
                             % (len(OUT_VARS) * tile * tpd * n * 4 / 1e9, 4 * n_days * n * 8 / 1e9)},
            "breakdown_ms": {"solar_tables": float(br[0]), "daily_mtclim": float(br[1]),
                             "disaggregate": float(br[2])},
            # the metric with and without the (time-independent) solar-geometry set-up
            "rank_ms_per_step": rank_ms,      # device time of every rank (value uses the maximum)
            "value_excl_solar_tables": total_cell_ts / max(1e-9, (ms_per_step - float(br[0])) * 1e-3),
            "roofline": {"bound": "hbm", "kernel": "msb_disaggregate = disagg_prepare_kernel + thermo_knot_kernel + "
                                                   "streams_day_kernel (+ disagg_kernel on the record's first / last "
                                                   "two days), timed together per call (--tile-days days)",
                         "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_cell_timestep": bytes_per_cts,
                         "bytes_per_launch": bytes_per_launch, "ms_per_launch": float(disagg_ms_per_launch),
                         "whole_step_frac": bytes_per_cts * cell_ts_rank / (ms_per_step * 1e-3) / 1e9 / peak},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
