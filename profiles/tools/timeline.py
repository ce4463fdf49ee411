"""GPU timeline of bench steps (kernels, copies, memsets with start times and the gaps between them).

    python profiles/tools/timeline.py [--workload conus_1_16] [--steps 2] [--threads N] [--out file]

Runs the bench's own DevicePass.step() under torch.profiler (Kineto / CUPTI activity records: no
replay, no serialisation -- unlike ncu the kernels run back to back at the clocks of a real step) and
prints, per step: every GPU activity in start order with its duration and the idle gap before it
(gaps above --gap-us only), then per-name totals.  Also the CUDA-event breakdown of the same loop
with the host prelude on 1 thread and on all threads (is the host on the critical path?)."""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def short(name):
    name = name.replace("msb::", "").replace("void ", "")
    return name.split("(")[0][:44]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="conus_16th_hourly")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--gap-us", type=float, default=20.0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    from metsim_b200 import synth
    grid, n_days, start, ts, prec_type = synth.CONFIGS[a.workload]
    lat, lon = synth.grid_cells(grid, np.random.default_rng(synth.SEED))
    dp = bench.DevicePass(argparse.Namespace(), dev, lat, lon, n_days, start, ts, prec_type, synth.SEED,
                          patch_rows=4 if grid.startswith("global") else 1)
    for _ in range(3):
        dp.step()
    dp.eng.check()

    # 1. CUDA-event breakdown, host prelude threads varied
    ev = lambda: torch.cuda.Event(enable_timing=True)
    orig = dp.eng.build_solar
    for nthr in (None, 1):
        dp.eng.build_solar = (lambda n=nthr: orig(n_host_threads=n)) if nthr else orig
        for _ in range(2):
            dp.step()
        marks = [[ev() for _ in range(4)] for _ in range(6)]
        for m in marks:
            dp.step(m)
        torch.cuda.synchronize()
        br = np.array([[m[i].elapsed_time(m[i + 1]) for i in range(3)] for m in marks]).mean(0)
        print("events, prelude threads %s: solar %.2f  daily %.2f  disagg %.2f ms" % (nthr or "all", *br))
    dp.eng.build_solar = orig

    # 2. activity timeline
    from torch.profiler import ProfilerActivity, profile
    try:
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(a.steps):
                dp.step()
            torch.cuda.synchronize()
    except Exception as exc:  # no CUPTI on this box
        print("torch.profiler unavailable:", exc)
        return
    acts = []
    for e in prof.events():
        if str(getattr(e, "device_type", "")).endswith("CUDA") and e.time_range is not None:
            acts.append((e.time_range.start, e.time_range.end, e.name))
    acts.sort()
    if not acts:
        print("no device activities recorded")
        return
    t0 = acts[0][0]
    lines = []
    tot = collections.OrderedDict()
    gaps = 0.0
    prev_end = t0
    for s, e, name in acts:
        gap = s - prev_end
        if gap > 0:
            gaps += gap
        k = short(name)
        d = tot.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += e - s
        lines.append("%10.1f us  +%8.1f  gap %7.1f  %s" % (s - t0, e - s, gap, k))
        prev_end = max(prev_end, e)
    span = prev_end - t0
    print("timeline of %d steps: span %.2f ms, busy %.2f ms, idle gaps %.2f ms" % (
        a.steps, span / 1e3, (span - gaps) / 1e3, gaps / 1e3))
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("   %8.3f ms/step  %4d launches  %8.1f us each  %s" % (us / 1e3 / a.steps, n, us / n, k))
    print("activities preceded by a gap above %.0f us:" % a.gap_us)
    for ln in lines:
        if float(ln.split("gap")[1].split()[0]) > a.gap_us:
            print("  ", ln)
    if a.out:
        with open(a.out, "w") as fh:
            fh.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
