import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import msb_testutil as util
from metsim_b200 import synth
from metsim_b200._lib import SUB_VARS
from metsim_b200.engine import Engine
from oracle import oracle_lib
target = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(20261020)
for case in range(target + 1):
    n = int(rng.integers(33, 97)); days = int(rng.integers(2, 14))
    ts = int(rng.choice([15, 20, 30, 60, 90, 120, 180, 240, 360]))
    lat = np.sort(rng.uniform(-89.9, 89.9, n)); lon = rng.uniform(-360.0, 540.0, n)
    start = str(np.datetime64("1999-01-01") + int(rng.integers(0, 3000)))
    f = synth.make_forcing(lat, lon, days, start, rng)
    params = dict(time_step=ts, prec_type=str(rng.choice(["uniform", "triangle", "mix"])), utc_offset=bool(rng.integers(0, 2)))
    call = int(rng.integers(1, days + 1))
print(case, n, days, ts, params, call, start)
want = oracle_lib.run(params, **util.oracle_kwargs(f))
tpd = 1440 // ts
for path in ("auto", "event"):
    if path == "event": os.environ["MSB_DISAGG_PATH"] = "event"
    else: os.environ.pop("MSB_DISAGG_PATH", None)
    for c in (call, days):
        eng = Engine(params); eng.load(**{k: f.get(k) for k in util.INPUT_KEYS}); eng.build_solar(); eng.daily(want=())
        parts = [eng.disaggregate(SUB_VARS, d0, min(days, d0 + c), torch.float64) for d0 in range(0, days, c)]
        torch.cuda.synchronize()
        for v in ("prec", "shortwave", "temp", "wind"):
            got = torch.cat([p[v] for p in parts]).cpu().numpy()
            w = want[v]
            err = np.abs(got - w) / np.maximum(np.abs(w), util.FLOORS[v])
            err[np.isnan(got) & np.isnan(w)] = 0
            bad = np.argwhere(~(err <= 1e-9))
            print(path, "call", c, v, "bad", len(bad), "first", bad[:6].tolist(),
                  [(float(got[i, j]), float(w[i, j])) for i, j in bad[:3]])
            if len(bad):
                cells = sorted(set(bad[:, 1].tolist()))[:5]
                for j in cells[:2]:
                    steps = bad[bad[:, 1] == j][:, 0]
                    print("   cell", j, "lon", lon[j], "lat", lat[j], "steps", steps[:12].tolist(), "days", sorted(set((steps // tpd).tolist())),
                          "month", [int(f["month"][d]) for d in sorted(set((steps // tpd).tolist()))][:4], "prec_d", f["prec"][:, j].round(3).tolist())
