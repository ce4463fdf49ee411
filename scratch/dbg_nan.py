import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import msb_testutil as util
from metsim_b200 import synth
from oracle import oracle_lib
from test_gpu_parity import engine_run
rng = np.random.default_rng(5)
n, d = 64, 31
lat = rng.uniform(25, 50, n); lon = rng.uniform(-125, -67, n)
f = synth.make_forcing(lat, lon, d, "2001-12-01", rng)
f["prec"] = np.where(rng.random((d, n)) < 0.5, rng.uniform(0.01, 0.6, (d, n)), f["prec"])
f["prec"][:3, ::2] = 0.3; f["prec"][-4:, 1::2] = 0.2; f["prec"][:, 7] = 0.25
params = dict(time_step=60, prec_type="triangle", utc_offset=True)
want = oracle_lib.run(params, **util.oracle_kwargs(f))
got = engine_run(f, params)
g, w = got["prec"], want["prec"]
bad = np.argwhere((g == 0) != (w == 0))
print(len(bad))
for i, c in bad[:12]:
    dd = i // 24
    print(i, c, 'day', dd, 'hour', i % 24, 'got', g[i, c], 'want', w[i, c], 'P', f["prec"][max(dd-1,0):dd+2, c], 'lon', lon[c])
