"""Per-stage time of a mid-latitude and a polar latitude band of the global 1/8 deg grid (same cell count)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metsim_b200 import synth
from metsim_b200.engine import Engine
VARS = ("temp", "prec", "shortwave", "longwave", "vapor_pressure", "rel_humid", "air_pressure", "wind")
rng = np.random.default_rng(synth.SEED)
lat, lon = synth.grid_cells("global1/8", rng)
N = 120000
ts = int(sys.argv[1]) if len(sys.argv) > 1 else 15
days = int(sys.argv[2]) if len(sys.argv) > 2 else 200
start = sys.argv[3] if len(sys.argv) > 3 else "1981-01-01"
bands = [("south band", slice(0, N)), ("middle band", slice(lat.size // 2, lat.size // 2 + N)), ("north band", slice(lat.size - N, lat.size))]
if len(sys.argv) > 5:   # narrow latitude ranges instead: "66.4:67.4,60:61"
    bands = []
    for rng_s in sys.argv[5].split(","):
        a_, b_ = (float(x) for x in rng_s.split(":"))
        bands.append((f"lat {a_}..{b_}", np.flatnonzero((lat >= a_) & (lat < b_))))
for name, sel in bands:
    la, lo = lat[sel], lon[sel]
    f = synth.make_forcing_torch(la, lo, days, start, "cuda")
    eng = Engine(dict(time_step=ts, prec_type="triangle", utc_offset=True))
    eng.load(**f)
    tpd = 1440 // ts
    tile = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    out = {v: torch.empty((tile * tpd, la.size), dtype=torch.float32, device="cuda") for v in VARS}
    def step(marks=None):
        s = eng.stream
        if marks: marks[0].record(s)
        eng.build_solar()
        if marks: marks[1].record(s)
        d = eng.daily(want=())
        if marks: marks[2].record(s)
        for d0 in range(0, days, tile):
            eng.disaggregate(VARS, d0, min(days, d0 + tile), torch.float32, out=out)
        if marks: marks[3].record(s)
        return d
    step(); step(); torch.cuda.synchronize()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    d = step(marks); torch.cuda.synchronize()
    # the two interior kernels separately (masked instantiations), 25-day calls
    sub_ms = {}
    for label, vs in (("thermo", ("temp", "longwave", "vapor_pressure", "rel_humid", "air_pressure", "wind")), ("streams", ("prec", "shortwave"))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per = []
        for d0 in range(0, days, tile):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(eng.stream)
            eng.disaggregate(vs, d0, min(days, d0 + tile), torch.float32, out=out)
            e1.record(eng.stream); torch.cuda.synchronize()
            per.append(round(e0.elapsed_time(e1), 1))
        sub_ms[label] = per
    for label, per in sub_ms.items():
        yr = [round(sum(per[k:k + 15]), 1) for k in range(0, len(per), 15)]    # ~375-day blocks
        print("   ", label, "per ~year:", yr)
    it = d["n_iter"].cpu().numpy()
    print(f"{name:12s} lat {la.min():7.2f}..{la.max():7.2f}  solar {marks[0].elapsed_time(marks[1]):6.1f}  daily {marks[1].elapsed_time(marks[2]):6.1f}  disagg {marks[2].elapsed_time(marks[3]):7.1f} ms   n_iter {np.bincount(it).tolist()}")
    del out, eng, f
    torch.cuda.empty_cache()
