"""Where does the solar-table phase of a bench step go?  (ncu: rows + moments = 5.0 ms, bench: 6.1 ms)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from metsim_b200 import synth
from metsim_b200.engine import Engine
lat, lon = synth.grid_cells("conus1/16", np.random.default_rng(synth.SEED))
dev = torch.device("cuda", 0)
f = synth.make_forcing_torch(lat[:928 * 520:1], lon[:928 * 520:1], 4, "2001-01-01", dev)
eng = Engine(dict(time_step=60, prec_type="triangle", utc_offset=True), dev)
eng.load(**f)
for _ in range(3):
    eng.build_solar()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
t0 = time.perf_counter()
for i in range(20):
    ev[i].record(eng.stream)
    eng.build_solar()
ev[20].record(eng.stream)
host = (time.perf_counter() - t0) / 20
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 20
print("host enqueue per build %.2f ms, wall per build %.2f ms, gpu per build %.2f ms (events)" % (
    host * 1e3, wall * 1e3, ev[0].elapsed_time(ev[20]) / 20))
for nthr in (1, 4, 16, 32):
    t0 = time.perf_counter()
    for i in range(5):
        eng.build_solar(n_host_threads=nthr)
    torch.cuda.synchronize()
    print("host threads", nthr, "wall per build %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
