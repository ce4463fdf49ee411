"""Where does the masked global grid lose time?  Same cell count, different cell placements."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metsim_b200 import synth
from metsim_b200.engine import Engine
VARS = ("temp", "prec", "shortwave", "longwave", "vapor_pressure", "rel_humid", "air_pressure", "wind")
rng = np.random.default_rng(1)
def grid(lats, lons, mask=None):
    la, lo = np.meshgrid(lats, lons, indexing="ij")
    return (la.ravel(), lo.ravel()) if mask is None else (la[mask], lo[mask])
N = 67000
cases = {}
cl, co = grid(25.03125 + np.arange(520) / 16.0, -124.96875 + np.arange(928) / 16.0)
cases["conus dense rows"] = (cl[:N], co[:N])
m = rng.random(cl.size) < 1 / 3
cases["conus random third"] = (cl[m][:N], co[m][:N])
gl, go = grid(np.arange(-55.75, 83.76, 0.5), np.arange(-179.75, 179.76, 0.5))
cases["global dense rows"] = (gl[:N], go[:N])
m = rng.random(gl.size) < 1 / 3
cases["global random third"] = (gl[m][:N], go[m][:N])
k = 3   # every 3rd longitude: regular, not random
la, lo = np.meshgrid(np.arange(-55.75, 83.76, 0.5), np.arange(-179.75, 179.76, 0.5)[::k], indexing="ij")
cases["global every 3rd lon"] = (la.ravel()[:N], lo.ravel()[:N])
for name, (lat, lon) in cases.items():
    f = synth.make_forcing_torch(lat, lon, 365, "2001-01-01", "cuda")
    eng = Engine(dict(time_step=60, prec_type="triangle", utc_offset=True))
    eng.load(**f); eng.build_solar(); eng.daily(want=())
    out = {v: torch.empty((365 * 24, lat.size), dtype=torch.float32, device="cuda") for v in VARS}
    for _ in range(2):
        eng.disaggregate(VARS, 0, 365, torch.float32, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.stream)
    for _ in range(3):
        eng.disaggregate(VARS, 0, 365, torch.float32, out=out)
    e1.record(eng.stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{name:24s} cells {lat.size} lats {np.unique(lat).size:4d}  disagg {ms:7.2f} ms  {lat.size*365*24/ms/1e6:8.1f} M cell-ts/ms")
    del out, eng, f
    torch.cuda.empty_cache()
