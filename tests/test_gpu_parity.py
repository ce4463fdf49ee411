"""Parity of the CUDA path (through the C ABI) against the oracle and the reference-generated
golden vectors.  Needs a B200; every test is marked gpu.

Bar (BASELINE.json north_star): bit-exact for integer / indexing work, max relative error <= 1e-9
on every fp64 output variable (relative error with the per-variable floor of msb_testutil.FLOORS).
"""
import ctypes as C
import os

import numpy as np
import pytest

import msb_testutil as util

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["auto", "event"])
def disagg_path(request, monkeypatch):
    """Both disaggregation paths: "auto" = the specialised interior-day kernels wherever they apply
    (+ the event-driven kernel on the record's edge days), "event" = the event-driven kernel for
    every day (MSB_DISAGG_PATH, read by msb_disaggregate on every call)."""
    if request.param == "event":
        monkeypatch.setenv("MSB_DISAGG_PATH", "event")
    else:
        monkeypatch.delenv("MSB_DISAGG_PATH", raising=False)
    return request.param


def engine_run(inputs, params, out_dtype=None, units=None, tuning=None):
    import torch
    from metsim_b200.engine import Engine
    from metsim_b200._lib import DAILY_VARS, SUB_VARS
    eng = Engine(params, tuning=tuning)
    eng.load(**{k: inputs.get(k) for k in util.INPUT_KEYS})
    res = eng.run(out_vars=SUB_VARS, out_dtype=out_dtype or torch.float64, units=units,
                  daily_vars=DAILY_VARS)
    out = {k: v.cpu().numpy() for k, v in res.items() if k != "daily"}
    out["daily"] = {k: v.cpu().numpy() for k, v in res["daily"].items()}
    return out


def compare_to_oracle(got, want, ctx, tol=util.TOL):
    worst = {}
    for oname, ename in util.DAILY.items():
        worst[oname] = util.assert_close(got["daily"][ename], want[oname], oname, context=ctx, tol=tol)
    for v in util.SUB:
        if v in want and v in got:
            worst[v] = util.assert_close(got[v], want[v], v, context=ctx, tol=tol)
    np.testing.assert_array_equal(got["daily"]["n_iter"], want["n_iter"], err_msg=ctx + ": tdew evaluations")
    return worst


def test_fp64_building_blocks_accuracy():
    """rcp / exp / sqrt / div sequences of msb_common.cuh: <= 4 ulp-ish against numpy."""
    import torch
    from metsim_b200 import _lib
    lib = _lib.lib()
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(1e-3, 5e3, 200000), rng.uniform(150.0, 400.0, 100000),
                        10.0 ** rng.uniform(-6, 6, 50000)])
    xd = torch.from_numpy(x).cuda()
    out = torch.empty((12, x.size), dtype=torch.float64, device="cuda")
    _lib.check(lib.msb_selftest_math(x.size, C.c_void_p(xd.data_ptr()), C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert np.max(np.abs(o[0] * x - 1.0)) < 1e-15
    print('rcp seed max rel err 2^%.1f' % np.log2(np.max(np.abs(o[5] * x - 1.0))))
    assert np.max(np.abs(o[5] * x - 1.0)) < 2.0 ** -18    # the cubic step needs a >= 18-bit seed
    assert np.max(np.abs(o[2] / np.sqrt(x) - 1.0)) < 2e-15
    assert np.max(np.abs(o[7] / np.sqrt(x) - 1.0)) < 5e-16
    assert np.max(np.abs(o[3] * x / 1.2345678901234567 - 1.0)) < 1e-15
    # second-order sequences of the interior-day kernel: three orders inside the 1e-9 contract
    assert np.max(np.abs(o[9] * x - 1.0)) < 1e-11
    assert np.max(np.abs(o[10] / np.sqrt(x) - 1.0)) < 1e-11
    # exp over the argument range the kernels use
    xe = rng.uniform(-40.0, 40.0, x.size)
    xd.copy_(torch.from_numpy(xe))
    _lib.check(lib.msb_selftest_math(x.size, C.c_void_p(xd.data_ptr()), C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert np.max(np.abs(o[1] / np.exp(xe) - 1.0)) < 1e-12
    assert np.max(np.abs(o[6] / np.exp(xe) - 1.0)) < 1e-15
    assert np.max(np.abs(o[11] / np.exp(xe) - 1.0)) < 2e-11
    # longitude wrap == math.remainder(lon, 360), bit for bit, including the ties at +-180
    import math
    lon = np.concatenate([rng.uniform(-1000.0, 1000.0, x.size - 16),
                          [180.0, -180.0, 540.0, -540.0, 900.0, 179.99999999999997, -179.99999999999997,
                           180.00000000000003, 359.9, -359.75, 0.0, 360.0, -360.0, 720.0, 0.125, -67.03125]])
    xd.copy_(torch.from_numpy(lon))
    _lib.check(lib.msb_selftest_math(x.size, C.c_void_p(xd.data_ptr()), C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy()[8], np.array([math.remainder(v, 360.0) for v in lon]))
    # svp in kPa over meteorological temperatures (physics.py:124-152)
    t = rng.uniform(-80.0, 60.0, x.size)
    xd.copy_(torch.from_numpy(t))
    _lib.check(lib.msb_selftest_math(x.size, C.c_void_p(xd.data_ptr()), C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    svp = 0.61078 * np.exp(17.269 * t / (237.3 + t))
    svp[t < 0] *= 1.0 + .00972 * t[t < 0] + .000042 * t[t < 0] ** 2
    assert np.max(np.abs(out.cpu().numpy()[4] / svp - 1.0)) < 1e-12


def test_solar_tables_against_oracle_table():
    """Per-latitude tables vs the oracle's per-cell 366x2880 table: rise/set indices bit-exact
    (this is the sunrise-sign hazard), window sums of the prefix table <= 1e-12 absolute."""
    import torch
    from metsim_b200.engine import Engine
    from oracle import oracle_lib
    lats = np.array([48.3125, 30.03125, -33.9, 0.25, 64.9, 70.3, -54.75, 83.75, 25.03125, 57.46875])
    n = lats.size
    eng = Engine({})
    dummy = np.zeros((2, n))
    eng.load(lat=lats, lon=np.zeros(n), elev=np.zeros(n), doy=[1, 2], month=[1, 1], t_min=dummy,
             t_max=dummy + 1, prec=dummy, state_t_min=np.zeros((90, n)),
             state_t_max=np.ones((90, n)), state_prec=np.zeros((90, n)))
    tbd = eng.build_solar()
    eng.stream.synchronize()
    tb = {k: v.cpu().numpy() for k, v in tbd.items()}
    ulat = eng._ulat
    q = tb["q"][:n * 365 * 2888].reshape(n, 365, 2888)
    rng = np.random.default_rng(3)
    for r, lat in enumerate(ulat):
        trf, dayl, potrad, _ = oracle_lib.solar_geom(0.0, float(lat))
        np.testing.assert_allclose(tb["daylength"].reshape(n, 366)[r], dayl, rtol=1e-15)
        np.testing.assert_allclose(tb["potrad"].reshape(n, 366)[r], potrad, rtol=1e-13)
        # rise / set (disaggregate.py:163-182), bit-exact
        mask = np.diff((trf > 0).astype(int), axis=1)
        rise, sett = np.zeros(366), np.zeros(366)
        for i, j in zip(*np.where(mask > 0)):
            rise[i] = j * 0.5
        for i, j in zip(*np.where(mask < 0)):
            sett[i] = j * 0.5
        rise[rise == 0] = 240.0
        sett[sett == 0] = 960.0
        np.testing.assert_array_equal(tb["rise_min"].reshape(n, 366)[r], rise)
        np.testing.assert_array_equal(tb["set_min"].reshape(n, 366)[r], sett)
        # window sums
        for y in rng.integers(0, 365, 12):
            c = np.concatenate([[0.0], np.cumsum(trf[y])])
            a = rng.integers(0, 2880, 64)
            b = np.minimum(a + rng.integers(1, 721, 64), 2880)
            w = q[r, y, b] - q[r, y, a] + np.where((a <= 1440) & (b > 1440), q[r, y, 2881], 0.0)
            np.testing.assert_allclose(w, c[b] - c[a], atol=2e-13, rtol=0)


@pytest.mark.parametrize("path", util.golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_matches_reference_golden(path, disagg_path):
    inputs, params, ref_daily, ref_sub, extra = util.load_golden(path)
    got = engine_run(inputs, params)
    cells = extra.get("ref_cells")
    pick = (lambda a: a[:, cells]) if cells is not None else (lambda a: a)
    ctx = os.path.basename(path)
    for oname, ename in util.DAILY.items():
        util.assert_close(pick(got["daily"][ename]), ref_daily[ename], oname, context=ctx)
    for v, want in ref_sub.items():
        util.assert_close(pick(got[v]), want, v, context=ctx)


CASES = [
    ("conus_16th_hourly", 640, 120, {}),
    ("global_half_deg", 700, 90, {}),
    ("conus_16th_10yr_3h_mix", 300, 400, {}),
    ("global_8th_30yr_15min", 256, 45, {}),
    ("test_domain", None, None, {}),
    ("conus_16th_hourly", 200, 60, {"utc_offset": False, "prec_type": "uniform", "time_step": 30}),
    ("global_half_deg", 333, 70, {"lw_type": "satterlund", "lw_cloud": "default", "time_step": 120}),
    ("global_half_deg", 200, 50, {"lw_type": "idso", "time_step": 360, "prec_type": "mix"}),
    ("global_half_deg", 200, 50, {"lw_type": "brutsaert", "time_step": 20}),
]


@pytest.mark.parametrize("name,cells,days,over", CASES,
                         ids=[f"{c[0]}-{c[1]}x{c[2]}-{i}" for i, c in enumerate(CASES)])
def test_cuda_matches_oracle_on_seeded_configs(name, cells, days, over, disagg_path):
    from metsim_b200 import synth
    from oracle import oracle_lib
    f, params = synth.make_config(name, max_cells=cells, n_days=days)
    params.update(over)
    want = oracle_lib.run(params, **util.oracle_kwargs(f))
    got = engine_run(f, params)
    compare_to_oracle(got, want, f"{name} {over}")


EXAMPLE_VARS = ("temp", "prec", "shortwave", "longwave", "vapor_pressure", "rel_humid",
                "air_pressure", "wind")            # examples/example_nc.conf:3


@pytest.mark.parametrize("name,cells,days,over", [
    ("conus_16th_hourly", 1500, 45, {}),
    ("conus_16th_hourly", 700, 400, {"start_dec": True}),
    ("global_half_deg", 900, 30, {}),
    ("conus_16th_10yr_3h_mix", 500, 60, {}),
    ("global_8th_30yr_15min", 400, 12, {}),
    ("conus_16th_hourly", 300, 5, {}),
    ("conus_16th_hourly", 300, 3, {}),
    ("global_half_deg", 400, 16, {"time_step": 360, "prec_type": "mix"}),
    ("global_half_deg", 400, 9, {"time_step": 20, "utc_offset": False}),
    ("conus_16th_hourly", 260, 30, {"time_step": 120, "prec_type": "uniform", "tmax_daylength_fraction": 0.5}),
    ("global_half_deg", 500, 40, {"start": "2000-12-10"}),    # day of year 366 -> 1 inside the interior days
    ("edge", None, 20, {"time_step": 60, "t_end": True}),      # polar rows, |lon| = 180, lon outside [-180, 180]
    ("edge", None, 12, {"time_step": 180, "prec_type": "mix"}),
], ids=lambda v: str(v) if not isinstance(v, dict) else "-".join(f"{k}{x}" for k, x in v.items()) or "std")
def test_fast_path_example_vars_f4_against_oracle(name, cells, days, over):
    """The configuration bench.py times: the 8 variables of examples/example_nc.conf, float32
    outputs, no unit conversion -- the specialised kernels of msb_disaggregate.  Checked against
    the oracle rounded to float32: the fp64 error (<= 1e-9) can only flip a float32 rounding on a
    ~1e-9 / 6e-8 fraction of the values, so all but a handful must be bit-equal and none may
    be off by more than one float32 ulp; precipitation zero / non-zero placement is exact."""
    import torch
    from metsim_b200 import synth
    from metsim_b200.engine import Engine
    from oracle import oracle_lib
    if name == "edge":
        lat = np.array([89.5, -89.5, 66.6, 0.0, 45.0, 45.0, 45.0, 45.0, 78.2, -71.0, 30.0, 30.0] * 3)
        lon = np.array([10.0, -10.0, 180.0, -180.0, 359.9, 540.0, -359.75, 0.125, -179.99, 179.99, 0.0, -105.0] * 3)
        f = synth.make_forcing(lat, lon, days, "2000-12-20", np.random.default_rng(31))
        params = dict(time_step=60, prec_type="triangle", utc_offset=True)
        if over.get("t_end"):
            f["t_end"] = np.stack([f["t_min"][-1] - 1.5, f["t_max"][-1] + 2.0])
    else:
        f, params = synth.make_config(name, max_cells=cells, n_days=days)
    if over.get("start_dec") or over.get("start"):   # e.g. a record that starts in December: NaN days at the head and tail
        doy, month = synth.calendar_vectors(over.get("start", "2001-12-05"), days)
        f["doy"], f["month"] = doy, month
    params.update({k: v for k, v in over.items() if k not in ("start_dec", "t_end", "start")})
    want = oracle_lib.run(params, **util.oracle_kwargs(f))
    eng = Engine(params)
    eng.load(**{k: f.get(k) for k in util.INPUT_KEYS})
    res = eng.run(out_vars=EXAMPLE_VARS, out_dtype=torch.float32)
    # (no daily column requested: this is the single-pass daily form, candidate planes and all)
    np.testing.assert_array_equal(res["daily"]["n_iter"].cpu().numpy(), want["n_iter"])
    for split in (None, 7):          # whole record in one call, then 7-day calls
        if split is not None:
            parts = [eng.disaggregate(EXAMPLE_VARS, d0, min(days, d0 + split), torch.float32)
                     for d0 in range(0, days, split)]
            torch.cuda.synchronize()
            res = {v: torch.cat([p[v] for p in parts]) for v in EXAMPLE_VARS}
        for v in EXAMPLE_VARS:
            got = res[v].cpu().numpy()
            assert got.dtype == np.float32
            w64 = want[v]
            w32 = w64.astype(np.float32)
            same_nan = np.isnan(got) == np.isnan(w32)
            assert same_nan.all(), v
            ok = ~np.isnan(w32)
            neq = got[ok] != w32[ok]
            ulp = np.spacing(np.abs(w32[ok]).astype(np.float32)).astype(np.float64)
            err = np.abs(got[ok].astype(np.float64) - w64[ok])
            floor = util.FLOORS[v]
            # within float32 rounding of the fp64 value (+ the 1e-9 contract)
            assert (err <= 0.5000001 * ulp + 1e-9 * np.maximum(np.abs(w64[ok]), floor)).all(), (v, split, float((err / ulp).max()))
            # (values far below the floor are rounding noise of sums that are mathematically zero --
            #  polar night, the first tiny step after sunrise -- and are left out of the count)
            sig = np.abs(w64[ok]) >= 1e-3 * floor
            assert neq[sig].sum() <= max(3, 2e-4 * sig.sum()), (v, split, int(neq[sig].sum()), int(sig.sum()))
        np.testing.assert_array_equal(res["prec"].cpu().numpy() == 0, want["prec"].astype(np.float32) == 0)
    # the same kernels with float64 outputs: the north-star bar itself, 1e-9 on every variable
    # (this is what pins the second-order exp / rcp / sqrt sequences of the interior-day kernel)
    res64 = eng.run(out_vars=EXAMPLE_VARS, out_dtype=torch.float64)
    worst = {v: util.assert_close(res64[v].cpu().numpy(), want[v], v, context=f"fast path f64 {name}")
             for v in EXAMPLE_VARS}
    print("fast path f64 worst relative errors:", {k: f"{e:.1e}" for k, e in worst.items()})
    for v in EXAMPLE_VARS:       # and the float32 outputs are exactly the rounded float64 ones
        np.testing.assert_array_equal(res[v].cpu().numpy() if split is None else res[v].cpu().numpy(),
                                      res64[v].cpu().numpy().astype(np.float32), err_msg=v)


# ---- the launch geometry bench.py times, and the record lengths of BASELINE configs[3] / [4] ------
# bench.py's shards are large enough that the library keeps its full work per thread: 16 knot-days /
# 16 output days per thread in the interior kernels and day chunks of 73 (CONUS year), 731 (CONUS
# decade) and 645 days (30-year global shard) in the daily pass.  Small shards would shorten all of
# these; msb_tuning pins them, so the oracle comparisons below run what the benchmark runs.
# sw_table=1: the plain shortwave table, what the library picks for the dense CONUS grid (a sample
# of a few hundred cells of it is sparse, and would get the window-major copy on its own)
BENCH_GEOMETRY = dict(knot_tile=16, day_tile=16, daily_chunk_len=73, sw_table=1)
WINDOW_MAJOR = dict(BENCH_GEOMETRY, sw_table=2)      # ... and what it picks for the masked global grids


def run_both_daily_forms(f, params, tuning, want, ctx, sub_vars, day_ranges=None, out_dtype=None):
    """Single-pass daily form (candidate planes, what bench.py times) and the two-pass form, each
    followed by the disaggregation of ``day_ranges`` (default: the whole record) in float64:
    n_iter EQUAL to the oracle's, every value within 1e-9."""
    import torch
    from metsim_b200._lib import DAILY_VARS
    from metsim_b200.engine import Engine
    days = f["t_min"].shape[0]
    tpd = 1440 // params["time_step"]
    day_ranges = day_ranges or [(0, days)]
    worst = {}
    for form in ("single", "two-pass"):
        eng = Engine(params, tuning=tuning)
        eng.load(**{k: f.get(k) for k in util.INPUT_KEYS})
        eng.build_solar()
        if form == "single":
            daily = eng.daily(want=(), single_pass=True)
            vp, sw = eng.daily_for_disagg()
            eng.check()                      # synchronises the engine's stream before the host reads
            util.assert_close(vp.cpu().numpy(), want["vp_daily"], "vp_daily", context=ctx + " single")
            util.assert_close(sw.cpu().numpy(), want["sw_daily"], "sw_daily", context=ctx + " single")
            util.assert_close(daily["tskc"].cpu().numpy(), want["tskc_daily"], "tskc_daily", context=ctx)
        else:
            daily = eng.daily(want=DAILY_VARS)
            eng.check()
            for oname, ename in util.DAILY.items():
                if oname in want:
                    util.assert_close(daily[ename].cpu().numpy(), want[oname], oname, context=ctx + " two-pass")
        np.testing.assert_array_equal(daily["n_iter"].cpu().numpy(), want["n_iter"],
                                      err_msg=f"{ctx} {form}: tdew evaluations")
        for d0, d1 in day_ranges:
            got = eng.disaggregate(sub_vars, d0, d1, out_dtype or torch.float64)
            eng.check()
            torch.cuda.synchronize()
            for v in sub_vars:
                e = util.assert_close(got[v].cpu().numpy(), want[v][d0 * tpd:d1 * tpd], v,
                                      context=f"{ctx} {form} days [{d0},{d1})")
                worst[v] = max(worst.get(v, 0.0), e)
            del got
    return worst


@pytest.mark.parametrize("name,cells,days,over,tuning", [
    ("conus_16th_hourly", 384, 365, {}, BENCH_GEOMETRY),                 # the headline configuration
    ("global_half_deg", 320, 365, {}, WINDOW_MAJOR),                     # n* from 3 to 7: the replay list
    ("conus_16th_hourly", 200, 100, {"tdew_tol": 1e-3}, BENCH_GEOMETRY), # every cell below the candidate window
    ("conus_16th_hourly", 200, 100, {"tdew_tol": 1e-10}, BENCH_GEOMETRY),  # every cell needs a second round
    ("conus_16th_hourly", 160, 80, {}, dict(BENCH_GEOMETRY, disagg_path="event")),
], ids=["conus-year", "global-year", "tol1e-3", "tol1e-10", "event-path"])
def test_benchmarked_launch_geometry_against_oracle(name, cells, days, over, tuning):
    """VERDICT r1 'weak' #1: the oracle-compared tests ran tile = 2 and chunks <= 15 days while
    bench.py runs tile = 16 and 73-day chunks.  Here the benchmarked geometry itself is held to the
    1e-9 / n_iter-equality bar, whole 365-day records, all ten variables, both daily forms."""
    from metsim_b200 import synth
    from metsim_b200._lib import SUB_VARS
    from oracle import oracle_lib
    f, params = synth.make_config(name, max_cells=cells, n_days=days)
    params.update(over)
    want = oracle_lib.run(params, **util.oracle_kwargs(f))
    hist = np.bincount(want["n_iter"])
    worst = run_both_daily_forms(f, params, tuning, want, f"{name} {over} {tuning}", SUB_VARS)
    print("n_iter histogram", hist.tolist(), "worst", {k: f"{e:.1e}" for k, e in worst.items()})


@pytest.mark.parametrize("name,cells,days,chunk,ranges", [
    # BASELINE configs[3]: 10 years -> 3-hourly, mix; bench.py chunks the daily pass by 731 days
    ("conus_16th_10yr_3h_mix", 96, 3652, 731, None),
    # BASELINE configs[4]: 30 years -> 15-minute; 645-day chunks on an 8-way shard.  The whole record
    # goes through the daily pass (the stop rule is an RMS over all 10,957 days, mtclim.py:63); a few
    # day ranges -- the record's ends, year boundaries, a leap day -- are disaggregated and compared
    ("global_8th_30yr_15min", 48, 10957, 645, [(0, 40), (3630, 3700), (7295, 7330), (10900, 10957)]),
], ids=["C4-3652d", "C5-10957d"])
def test_long_records_against_oracle(name, cells, days, chunk, ranges):
    """VERDICT r1 'missing' #3: records longer than 400 days were never oracle-checked."""
    from metsim_b200 import synth
    from oracle import oracle_lib
    f, params = synth.make_config(name, max_cells=cells, n_days=days)
    want = oracle_lib.run(params, **util.oracle_kwargs(f))
    assert want["n_iter"].min() >= 1
    tuning = dict(knot_tile=16, day_tile=16, daily_chunk_len=chunk)
    worst = run_both_daily_forms(f, params, tuning, want, f"{name} {days} d", util.SUB, day_ranges=ranges)
    print("n_iter histogram", np.bincount(want["n_iter"]).tolist(), "worst", {k: f"{e:.1e}" for k, e in worst.items()})


def test_two_host_contexts_run_concurrently_and_agree():
    """no global mutable state.  Two Python threads, a context each, on one GPU, run
    different requests at the same time (ctypes releases the GIL around the C call); each result is
    bit-equal to the same request run alone, and a context can be reused and destroyed freely."""
    import threading
    from metsim_b200 import synth
    from metsim_b200.engine import HostContext, run_host
    fa, pa = synth.make_config("conus_16th_hourly", max_cells=900, n_days=50)
    fb, pb = synth.make_config("global_half_deg", max_cells=700, n_days=35)
    pb = dict(pb, time_step=30, prec_type="mix")
    kwa = {k: fa.get(k) for k in util.INPUT_KEYS}
    kwb = {k: fb.get(k) for k in util.INPUT_KEYS}
    alone_a = run_host(pa, out_vars=EXAMPLE_VARS, **kwa)
    alone_b = run_host(pb, out_vars=EXAMPLE_VARS, tile_days=3, **kwb)
    res, errs = {}, []

    def work(tag, params, kw, tile, reps):
        try:
            with HostContext(0) as ctx:
                for _ in range(reps):
                    res[tag] = run_host(params, out_vars=EXAMPLE_VARS, tile_days=tile, ctx=ctx, **kw)
                assert ctx.device_bytes > 0
        except BaseException as exc:       # surfaced below
            errs.append(exc)
    ta = threading.Thread(target=work, args=("a", pa, kwa, 0, 4))
    tb = threading.Thread(target=work, args=("b", pb, kwb, 3, 4))
    ta.start(); tb.start(); ta.join(); tb.join()
    assert not errs, errs
    for v in EXAMPLE_VARS:
        np.testing.assert_array_equal(res["a"][v], alone_a[v], err_msg=v)
        np.testing.assert_array_equal(res["b"][v], alone_b[v], err_msg=v)
    np.testing.assert_array_equal(res["a"]["daily"]["n_iter"], alone_a["daily"]["n_iter"])


@pytest.mark.parametrize("subset", [
    ("temp", "prec", "shortwave", "longwave", "vapor_pressure", "rel_humid"),   # MetSim.params default out_vars
    ("temp",), ("prec", "shortwave"), ("wind", "longwave"), ("air_pressure", "rel_humid", "prec"),
    ("shortwave",), ("vapor_pressure", "wind"),
], ids=lambda v: "+".join(v))
def test_fast_path_variable_subsets(subset):
    """Any subset of the default variable set takes the specialised kernels (a variable that was
    not requested is simply not stored): same bits as in the full-set run, float32 and float64,
    and the float64 values within 1e-9 of the oracle.  Triangle precipitation without t_pk12 /
    dur12 is fine as long as prec is not requested; without a wind forcing wind is skipped."""
    import torch
    from metsim_b200 import synth
    from metsim_b200.engine import Engine
    from oracle import oracle_lib
    f, params = synth.make_config("conus_16th_hourly", max_cells=333, n_days=21)
    want = oracle_lib.run(params, **util.oracle_kwargs(f))
    eng = Engine(params)
    eng.load(**{k: f.get(k) for k in util.INPUT_KEYS})
    full32 = eng.run(out_vars=EXAMPLE_VARS, out_dtype=torch.float32)
    full64 = eng.run(out_vars=EXAMPLE_VARS, out_dtype=torch.float64)
    g = dict(f)
    if "prec" not in subset:
        g.pop("t_pk12"); g.pop("dur12")
    if "wind" not in subset:
        g.pop("wind")
    sub_eng = Engine(params)
    sub_eng.load(**{k: g.get(k) for k in util.INPUT_KEYS})
    for dtype, full in ((torch.float32, full32), (torch.float64, full64)):
        got = sub_eng.run(out_vars=subset, out_dtype=dtype)
        assert set(got) - {"daily"} == set(subset)
        for v in subset:
            assert torch.equal(got[v], full[v]), (v, dtype)
            if dtype == torch.float64:
                util.assert_close(got[v].cpu().numpy(), want[v], v, context=f"subset {subset}")


def test_fast_path_through_run_host_and_tile_independence():
    """msb_run_host (host buffers, what bench.py's e2e leg and a reference-side binding call) on the
    default variable set: bit-identical to the device-resident fast path for several tile sizes --
    the interior-day kernels and the edge-day kernel must agree wherever a tile boundary falls."""
    import torch
    from metsim_b200 import synth
    from metsim_b200.engine import Engine, run_host
    f, params = synth.make_config("conus_16th_hourly", max_cells=777, n_days=41)
    eng = Engine(params)
    eng.load(**{k: f.get(k) for k in util.INPUT_KEYS})
    ref = eng.run(out_vars=EXAMPLE_VARS, out_dtype=torch.float32)
    ref = {v: ref[v].cpu().numpy() for v in EXAMPLE_VARS}
    kw = {k: f.get(k) for k in util.INPUT_KEYS}
    for tile in (0, 1, 2, 3, 7, 40, 41):
        got = run_host(params, out_vars=EXAMPLE_VARS, tile_days=tile, **kw)
        for v in EXAMPLE_VARS:
            np.testing.assert_array_equal(got[v], ref[v], err_msg=f"{v} tile={tile}")


def test_streaming_writer_through_on_tile_hook(tmp_path):
    """msb_run_host's on_tile consumer hook + the NetCDF-3 writer (metsim/metsim.py:363-454 in
    streaming form): outputs go through a 2-slot pinned host ring, every finished tile is scattered
    into (time, lat, lon) and written while the next one computes; the file, read back with scipy,
    equals the whole-record result bit for bit, NaN outside the mask.  An exception raised by the
    consumer aborts the run and surfaces unchanged."""
    from scipy.io import netcdf_file
    from metsim_b200 import synth
    from metsim_b200.engine import run_host
    from metsim_b200.ncwriter import NC3Writer
    rng = np.random.default_rng(17)
    lats = 40.03125 + np.arange(12) / 16.0
    lons = -110.96875 + np.arange(40) / 16.0
    mask = rng.random((lats.size, lons.size)) < 0.7
    la, lo = np.meshgrid(lats, lons, indexing="ij")
    days, ts, tile = 23, 60, 4
    tpd = 1440 // ts
    f = synth.make_forcing(la[mask], lo[mask], days, "2001-05-01", rng)
    params = dict(time_step=ts, prec_type="triangle", utc_offset=True)
    kw = {k: f.get(k) for k in util.INPUT_KEYS}
    n = int(mask.sum())
    ref = run_host(params, out_vars=EXAMPLE_VARS, **kw)                       # whole record in host memory
    ring = {v: np.empty((2 * tile * tpd, n), dtype=np.float32) for v in EXAMPLE_VARS}
    path = os.path.join(tmp_path, "forcing.nc")
    seen = []
    with NC3Writer(path, n_steps=days * tpd, lat=lats, lon=lons, cell_index=np.flatnonzero(mask.ravel()),
                   variables={v: {"units": "-"} for v in EXAMPLE_VARS},
                   time_values=np.arange(days * tpd) * float(ts), time_units="minutes since 2001-05-01") as w:
        write = w.tile_consumer(ring, tile, tpd, ring=True)

        def on_tile(k, d0, d1, slot):
            seen.append((k, d0, d1, slot))
            write(k, d0, d1, slot)
        run_host(params, out_vars=EXAMPLE_VARS, tile_days=tile, out=ring, ring=True, on_tile=on_tile, **kw)
        assert all(c == days * tpd for c in w.steps_written.values())
    assert seen == [(k, k * tile, min(days, (k + 1) * tile), k & 1) for k in range(-(-days // tile))]
    nc = netcdf_file(path, "r", mmap=False)
    idx = np.flatnonzero(mask.ravel())
    for v in EXAMPLE_VARS:
        got = nc.variables[v].data.reshape(days * tpd, -1)
        np.testing.assert_array_equal(got[:, idx], ref[v], err_msg=v)
        assert np.isnan(np.delete(got, idx, axis=1)).all()
    nc.close()

    class Boom(RuntimeError):
        pass

    def bad(k, d0, d1, slot):
        if k == 2:
            raise Boom("disk full")
    with pytest.raises(Boom):
        run_host(params, out_vars=("temp",), tile_days=tile, on_tile=bad, **kw)
    # the library is usable afterwards
    again = run_host(params, out_vars=("temp",), **kw)
    np.testing.assert_array_equal(again["temp"], ref["temp"])


@pytest.mark.parametrize("geometry", ["auto", "bench", "window"])
def test_seeded_fuzz_small_records_against_oracle(disagg_path, geometry):
    """(geometry "bench": the per-thread work of the benchmarked configuration pinned through
    msb_tuning -- one thread then covers the whole short record; "window": the same with the
    window-major copy of the shortwave table, which only the interior-day kernel reads.)  Forty seeded random small shards: record lengths 2..13 days (so that the interior range is
    empty, one day, or a few days, and every call boundary falls somewhere else), random time step,
    precipitation type, utc offset, start date, call length, latitudes from pole to pole and
    longitudes from -360 to 540.  All ten variables, float64, against the oracle at 1e-9; the
    record is produced in calls of a random length and must not depend on it."""
    import torch
    from metsim_b200 import synth
    from metsim_b200._lib import DAILY_VARS, SUB_VARS
    from metsim_b200.engine import Engine
    from oracle import oracle_lib
    if geometry == "window" and disagg_path == "event":
        pytest.skip("the event-driven kernel reads the plain table only")
    tuning = {"auto": None, "bench": BENCH_GEOMETRY, "window": WINDOW_MAJOR}[geometry]
    rng = np.random.default_rng(20261020)
    for case in range(int(os.environ.get("MSB_FUZZ_CASES", "40"))):   # more cases: a longer one-off hunt
        n = int(rng.integers(33, 97))
        days = int(rng.integers(2, 14))
        ts = int(rng.choice([15, 20, 30, 60, 90, 120, 180, 240, 360]))
        lat = np.sort(rng.uniform(-89.9, 89.9, n))
        lon = rng.uniform(-360.0, 540.0, n)
        start = str(np.datetime64("1999-01-01") + int(rng.integers(0, 3000)))
        f = synth.make_forcing(lat, lon, days, start, rng)
        params = dict(time_step=ts, prec_type=str(rng.choice(["uniform", "triangle", "mix"])),
                      utc_offset=bool(rng.integers(0, 2)))
        want = oracle_lib.run(params, **util.oracle_kwargs(f))
        eng = Engine(params, tuning=tuning)
        eng.load(**{k: f.get(k) for k in util.INPUT_KEYS})
        eng.build_solar()
        daily = eng.daily(want=DAILY_VARS)
        call = int(rng.integers(1, days + 1))
        parts = [eng.disaggregate(SUB_VARS, d0, min(days, d0 + call), torch.float64)
                 for d0 in range(0, days, call)]
        eng.check()
        torch.cuda.synchronize()
        got = {v: torch.cat([p[v] for p in parts]).cpu().numpy() for v in SUB_VARS}
        got["daily"] = {k: v.cpu().numpy() for k, v in daily.items()}
        compare_to_oracle(got, want, f"fuzz case {case}: n={n} days={days} ts={ts} {params} call={call} start={start}")


def test_seeded_fuzz_host_path_subsets_units_and_tiles():
    """Forty seeded random requests through msb_run_host (host buffers, random tile length):
    random variable subsets, float32 / float64, random unit conversions, mtclim or passthrough
    (with random supplied columns), optional t_end.  Every request must equal the oracle (float64:
    1e-9; float32: within rounding) -- whatever mix of specialised and event-driven kernels and
    whatever tile boundaries it ends up with."""
    from metsim_b200 import synth
    from metsim_b200._lib import SUB_VARS
    from metsim_b200.engine import run_host, unit_scale_offset
    from oracle import oracle_lib
    rng = np.random.default_rng(777)
    unit_choices = {"temp": ["C", "K"], "prec": ["mm timestep-1", "mm h-1", "mm s-1"],
                    "vapor_pressure": ["Pa", "hPa", "kPa"], "air_pressure": ["kPa", "hPa", "Pa"],
                    "tskc": ["fraction", "%"], "rel_humid": ["%", "fraction"]}
    for case in range(int(os.environ.get("MSB_FUZZ_CASES", "40"))):
        n = int(rng.integers(33, 80))
        days = int(rng.integers(3, 16))
        ts = int(rng.choice([30, 60, 120, 180, 360]))
        lat = np.sort(rng.uniform(-75.0, 80.0, n))
        lon = rng.uniform(-180.0, 180.0, n)
        start = str(np.datetime64("2000-01-01") + int(rng.integers(0, 2500)))
        f = synth.make_forcing(lat, lon, days, start, rng)
        params = dict(time_step=ts, prec_type=str(rng.choice(["uniform", "triangle", "mix"])),
                      utc_offset=bool(rng.integers(0, 2)))
        if rng.random() < 0.3:
            params["method"] = "passthrough"
            f["given_shortwave"] = rng.uniform(10, 380, (days, n))
            if rng.random() < 0.5:
                f["given_vapor_pressure"] = rng.uniform(150, 2000, (days, n))
            if rng.random() < 0.5:
                f["given_longwave"] = rng.uniform(170, 430, (days, n))
            if rng.random() < 0.5:
                params["sw_averaging"] = "daylight"
        if rng.random() < 0.3:
            f["t_end"] = np.stack([f["t_min"][-1] - 1.5, f["t_max"][-1] + 2.0])
        k = int(rng.integers(1, len(SUB_VARS) + 1))
        subset = tuple(sorted(rng.choice(SUB_VARS, size=k, replace=False).tolist(), key=SUB_VARS.index))
        units = {v: str(rng.choice(unit_choices[v])) for v in subset if v in unit_choices and rng.random() < 0.4}
        dtype = np.float32 if rng.random() < 0.5 else np.float64
        tile = int(rng.integers(0, days + 2))
        ctx = f"host fuzz {case}: n={n} days={days} ts={ts} {params} subset={subset} units={units} {dtype.__name__} tile={tile}"
        want = oracle_lib.run(params, **util.oracle_kwargs(f))
        got = run_host(params, out_vars=subset, out_dtype=dtype, units=units, tile_days=tile,
                       **{k2: f.get(k2) for k2 in util.INPUT_KEYS})
        assert set(got) - {"daily"} == set(subset), ctx
        for v in subset:
            sc, of = unit_scale_offset(v, units.get(v), ts)
            w = want[v] * sc + of
            if dtype == np.float64:
                e = util.relerr(got[v], w, util.FLOORS[v] * abs(sc))
                assert e <= 2e-9, f"{ctx}: {v} relative error {e:.3e}"
            else:
                ok = ~np.isnan(w)
                assert (np.isnan(got[v]) == ~ok).all(), ctx
                ulp = np.spacing(np.abs(w[ok]).astype(np.float32)).astype(np.float64)
                err = np.abs(got[v][ok].astype(np.float64) - w[ok])
                assert (err <= 0.5000001 * ulp + 2e-9 * np.maximum(np.abs(w[ok]), util.FLOORS[v] * abs(sc))).all(), (ctx, v)


def test_edge_longitudes_polar_and_wraparound(disagg_path):
    """lon outside [-180,180], |lon| = 180, polar day/night latitudes, 2-day record."""
    from metsim_b200 import synth
    from oracle import oracle_lib
    rng = np.random.default_rng(11)
    lat = np.array([89.5, -89.5, 66.6, 0.0, 45.0, 45.0, 45.0, 45.0, 78.2, -71.0])
    lon = np.array([10.0, -10.0, 180.0, -180.0, 359.9, 540.0, -359.75, 0.125, -179.99, 179.99])
    for d, start in ((2, "2001-06-20"), (20, "2000-12-20"), (20, "2001-06-10")):
        f = synth.make_forcing(lat, lon, d, start, rng)
        for ts, pt in ((60, "triangle"), (180, "mix"), (360, "uniform")):
            params = dict(time_step=ts, prec_type=pt, utc_offset=True)
            want = oracle_lib.run(params, **util.oracle_kwargs(f))
            got = engine_run(f, params)
            compare_to_oracle(got, want, f"edge d={d} ts={ts} {pt}")


def test_precipitation_nan_days_and_fill(disagg_path):
    """Degenerate triangle days (0/0 -> NaN day, disaggregate.py:327) in runs, at the start and at
    the end of the record, with the ffill/bfill of disaggregate.py:130."""
    from metsim_b200 import synth
    from oracle import oracle_lib
    rng = np.random.default_rng(5)
    n, d = 64, 31
    lat = rng.uniform(25, 50, n)
    lon = rng.uniform(-125, -67, n)
    f = synth.make_forcing(lat, lon, d, "2001-12-01", rng)   # December: t_pk = dur = prec
    f["prec"] = np.where(rng.random((d, n)) < 0.5, rng.uniform(0.01, 0.6, (d, n)), f["prec"])
    f["prec"][:3, ::2] = 0.3      # leading NaN days -> bfill
    f["prec"][-4:, 1::2] = 0.2    # trailing NaN days -> ffill
    f["prec"][:, 7] = 0.25        # an all-NaN column stays NaN
    for ts in (60, 180, 15):
        params = dict(time_step=ts, prec_type="triangle", utc_offset=True)
        want = oracle_lib.run(params, **util.oracle_kwargs(f))
        assert np.isnan(want["prec"][:, 7]).all()
        got = engine_run(f, params)
        compare_to_oracle(got, want, f"nan-days ts={ts}")
        # placement is integer work: zero / non-zero pattern must be identical
        np.testing.assert_array_equal(got["prec"] == 0, want["prec"] == 0)


@pytest.mark.parametrize("ts,prec_type", [(60, "triangle"), (15, "triangle"), (180, "mix"), (360, "uniform"), (20, "uniform")])
def test_window_major_shortwave_table_is_bit_identical(ts, prec_type):
    """msb_solar_tables.q_win holds the SAME entries as q in another order, and the shortwave
    kernel does the same arithmetic on them: float64 outputs must agree bit for bit, on the device
    path (forced through msb_tuning.sw_table) and through msb_run_host, where the library decides
    by the longitude spacing of the cells (scattered cells -> window-major, a dense row -> plain)."""
    import torch
    from metsim_b200 import synth
    from metsim_b200._lib import SUB_VARS
    from metsim_b200.engine import Engine, run_host
    rng = np.random.default_rng(77 + ts)
    n, days = 700, 23
    lat = np.sort(rng.choice(np.linspace(-80.0, 84.0, 41), n))     # 41 table rows (both copies stay small)
    lon = rng.uniform(-200.0, 400.0, n)
    f = synth.make_forcing(lat, lon, days, "2003-11-20", rng)
    params = dict(time_step=ts, prec_type=prec_type, utc_offset=True)
    res = {}
    for mode in (1, 2):
        eng = Engine(params, tuning=dict(sw_table=mode))
        eng.load(**{k: f.get(k) for k in util.INPUT_KEYS})
        assert eng.sw_sectors_per_load > 8.0 and eng.use_window_major() == (mode == 2)
        tb = eng.build_solar()
        assert ("q_win" in tb) == (mode == 2)
        eng.daily(want=())
        got = eng.disaggregate(SUB_VARS, 0, days, torch.float64)
        eng.check()                          # synchronises the engine's stream before the host reads
        res[mode] = {v: t.cpu().numpy() for v, t in got.items()}
    for v in SUB_VARS:
        np.testing.assert_array_equal(res[1][v].view(np.uint64), res[2][v].view(np.uint64), err_msg=v)
    assert np.isfinite(res[2]["shortwave"]).all() and res[2]["shortwave"].max() > 100.0
    # host path: the auto rule picks the window-major copy for these scattered cells, the plain table when told to
    kw = {k: f.get(k) for k in util.INPUT_KEYS}
    host = {mode: run_host(params, out_vars=("shortwave", "temp"), out_dtype=np.float64, tile_days=7,
                           tuning=dict(sw_table=mode), **kw) for mode in (0, 1)}
    for mode in (0, 1):
        np.testing.assert_array_equal(host[mode]["shortwave"].view(np.uint64), res[1]["shortwave"].view(np.uint64))
    # a dense CONUS row keeps the plain table under the auto rule
    la, lo = synth.grid_cells("conus1/16", np.random.default_rng(synth.SEED))
    eng = Engine(params)
    fd = synth.make_forcing(la[:928], lo[:928], 3, "2001-01-01", rng)
    eng.load(**{k: fd.get(k) for k in util.INPUT_KEYS})
    assert eng.sw_sectors_per_load < 8.0 and not eng.use_window_major()


def test_run_host_equals_device_path_and_tiling():
    """msb_run_host (host buffers, tiles, double-buffered D2H) is bit-identical to the
    device-resident path, for several tile sizes and the ring-buffer mode."""
    import torch
    from metsim_b200 import synth
    from metsim_b200.engine import run_host
    from metsim_b200._lib import DAILY_VARS, SUB_VARS
    f, params = synth.make_config("conus_16th_hourly", max_cells=500, n_days=37)
    ref = engine_run(f, params, out_dtype=torch.float32)
    kw = {k: f.get(k) for k in util.INPUT_KEYS}
    for tile in (0, 1, 5, 16, 37):
        got = run_host(params, out_vars=SUB_VARS, daily_vars=DAILY_VARS, tile_days=tile, **kw)
        for v in SUB_VARS:
            np.testing.assert_array_equal(got[v], ref[v], err_msg=f"{v} tile={tile}")
        for v in DAILY_VARS:
            np.testing.assert_array_equal(got["daily"][v], ref["daily"][v], err_msg=v)
    # ring: only the last two tiles survive, at slot (tile index & 1)
    tile, tpd = 5, 24
    ring = {v: np.empty((2 * tile * tpd, 500), dtype=np.float32) for v in ("temp", "prec")}
    run_host(params, out_vars=("temp", "prec"), tile_days=tile, out=ring, ring=True, **kw)
    n_tiles = -(-37 // tile)
    for k in (n_tiles - 2, n_tiles - 1):
        d0, d1 = k * tile, min(37, (k + 1) * tile)
        slot = (k & 1) * tile * tpd
        np.testing.assert_array_equal(ring["temp"][slot:slot + (d1 - d0) * tpd], ref["temp"][d0 * tpd:d1 * tpd])


def test_float32_outputs_and_unit_conversion(disagg_path):
    """out_precision f4 (reference default, metsim.py:154) and the converters of units.py, fused
    into the stores of both disaggregation paths."""
    import torch
    from metsim_b200 import synth
    f, params = synth.make_config("global_half_deg", max_cells=128, n_days=20)
    base = engine_run(f, params)
    f32 = engine_run(f, params, out_dtype=torch.float32)
    for v in util.SUB:
        np.testing.assert_array_equal(f32[v], base[v].astype(np.float32), err_msg=v)
    units = {"temp": "K", "prec": "mm s-1", "vapor_pressure": "hPa", "air_pressure": "Pa",
             "tskc": "%", "rel_humid": "fraction"}
    conv = engine_run(f, params, units=units)
    ts = params["time_step"]
    np.testing.assert_allclose(conv["temp"], base["temp"] + 273.15, rtol=1e-15)
    np.testing.assert_allclose(conv["prec"], base["prec"] / (ts * 60.0), rtol=1e-15)
    np.testing.assert_allclose(conv["vapor_pressure"], base["vapor_pressure"] / 100.0, rtol=1e-15)
    np.testing.assert_allclose(conv["air_pressure"], base["air_pressure"] * 1000.0, rtol=1e-15)
    np.testing.assert_allclose(conv["tskc"], base["tskc"] * 100.0, rtol=1e-15)
    np.testing.assert_allclose(conv["rel_humid"], base["rel_humid"] / 100.0, rtol=1e-15)


def test_errors_mirror_the_reference():
    from metsim_b200 import synth
    from metsim_b200._lib import MetsimB200Error
    f, params = synth.make_config("global_half_deg", max_cells=64, n_days=12)
    bad = dict(f)
    bad["t_max"] = f["t_max"].copy()
    bad["t_max"][3, 5] = bad["t_min"][3, 5] - 0.5
    with pytest.raises(ValueError, match="lower than daily minimum"):   # metsim.py:612-614
        engine_run(bad, params)
    with pytest.raises(MetsimB200Error, match="Time step"):            # metsim.py:641-646
        engine_run(f, dict(params, time_step=7))
    with pytest.raises(MetsimB200Error, match="t_pk12"):
        engine_run({k: v for k, v in f.items() if k not in ("t_pk12", "dur12")},
                   dict(params, prec_type="triangle"))


def test_registry_method_module_run_df():
    """metsim_b200.methods.mtclim_b200.run(df, params): the per-cell callable of the reference
    registry (metsim.py:139, :744) on golden inputs; adds the 7 columns in place, returns df."""
    import pandas as pd
    from metsim_b200 import registry
    from metsim_b200.methods import mtclim_b200
    methods = {}
    registry.register(methods)
    assert methods["mtclim_b200"] is mtclim_b200 and callable(methods["mtclim_b200"].run)
    inputs, params, ref_daily, _, _ = util.load_golden(os.path.join(util.GOLDEN, "synth_hourly_triangle_utc.npz"))
    for c in (0, 3, 5):
        df = pd.DataFrame({k: inputs[k][:, c] for k in ("t_min", "t_max", "prec")})
        for k in ("seasonal_prec", "smoothed_dtr", "daylength", "potrad", "tt_max"):
            df[k] = ref_daily[k][:, c]
        df["dtr"] = df["t_max"] - df["t_min"]
        p = dict(params, elev=float(inputs["elev"][c]), lapse_rate=0.0065, tday_coef=0.45,
                 sw_prec_thresh=0.0, rain_scalar=0.75, lw_cloud="cloud_deardorff", tdew_tol=1e-6)
        out = mtclim_b200.run(df, p)
        assert out is df
        for oname, ename in util.DAILY.items():
            if ename in mtclim_b200.ADDED_COLUMNS:
                util.assert_close(df[ename].values, ref_daily[ename][:, c], oname, context=f"run(df) cell {c}")


def test_passthrough_method_against_oracle_and_errors(disagg_path):
    """'passthrough' (metsim/methods/passthrough.py:28-46) on seeded grids against the oracle:
    supplied shortwave alone and with vapour pressure / tskc / longwave, both sw_averaging modes,
    f4 outputs with units; the longwave rescale makes every day's mean equal the supplied value
    (disaggregate.py:583-589); a missing shortwave column raises like the reference's assert."""
    import torch
    from metsim_b200 import synth
    from metsim_b200._lib import MetsimB200Error
    from oracle import oracle_lib
    rng = np.random.default_rng(23)
    f, params = synth.make_config("conus_16th_hourly", max_cells=300, n_days=45)
    d, n = f["t_min"].shape
    cols = dict(given_shortwave=rng.uniform(10, 380, (d, n)), given_vapor_pressure=rng.uniform(150, 2000, (d, n)),
                given_tskc=rng.uniform(0, 1, (d, n)), given_longwave=rng.uniform(170, 430, (d, n)))
    for given, over in ((("given_shortwave",), {}),
                        (("given_shortwave", "given_longwave"), {"sw_averaging": "daylight", "time_step": 180}),
                        (tuple(cols), {"time_step": 30, "prec_type": "mix"}),
                        (("given_shortwave", "given_tskc", "given_longwave"),
                         {"lw_type": "anderson", "lw_cloud": "default", "utc_offset": False})):
        prm = dict(params, method="passthrough", **over)
        g = dict(f, **{k: cols[k] for k in given})
        want = oracle_lib.run(prm, **util.oracle_kwargs(g))
        got = engine_run(g, prm)
        compare_to_oracle(got, want, f"passthrough {given} {over}")
        if "given_longwave" in given:
            tpd = 1440 // prm["time_step"]
            np.testing.assert_allclose(got["longwave"].reshape(d, tpd, n).mean(1), cols["given_longwave"], rtol=1e-12)
            f32 = engine_run(g, prm, out_dtype=torch.float32, units={"longwave": "W m-2", "temp": "K"})
            np.testing.assert_array_equal(f32["longwave"], got["longwave"].astype(np.float32))
    with pytest.raises(MetsimB200Error, match="shortwave"):          # passthrough.py:29
        engine_run(f, dict(params, method="passthrough"))


def test_passthrough_run_host_and_registry_module():
    """msb_run_host with the passthrough columns equals the device path bit for bit, and the
    registry module passthrough_b200.run(df, params) reproduces the reference's daily columns."""
    import pandas as pd
    from metsim_b200 import registry
    from metsim_b200.engine import run_host
    from metsim_b200._lib import DAILY_VARS, SUB_VARS
    from metsim_b200.methods import passthrough_b200
    path = os.path.join(util.GOLDEN, "passthrough_3hourly_daylight_all.npz")
    inputs, params, ref_daily, _, _ = util.load_golden(path)
    ref = engine_run(inputs, params)
    got = run_host(params, out_vars=SUB_VARS, daily_vars=DAILY_VARS, out_dtype=np.float64, tile_days=4,
                   **{k: inputs.get(k) for k in util.INPUT_KEYS})
    for v in SUB_VARS:
        np.testing.assert_array_equal(got[v], ref[v], err_msg=v)
    methods = {}
    registry.register(methods)
    assert methods["passthrough_b200"] is passthrough_b200
    for c in (0, 4):
        df = pd.DataFrame({k: inputs[k][:, c] for k in ("t_min", "t_max", "prec")})
        for k in ("seasonal_prec", "smoothed_dtr", "daylength", "potrad", "tt_max"):
            df[k] = ref_daily[k][:, c]
        df["dtr"] = df["t_max"] - df["t_min"]
        df["shortwave"] = inputs["given_shortwave"][:, c]
        df["vapor_pressure"] = inputs["given_vapor_pressure"][:, c]
        p = dict(params, elev=float(inputs["elev"][c]), lapse_rate=0.0065, tday_coef=0.45,
                 sw_prec_thresh=0.0, rain_scalar=0.75, lw_cloud="cloud_deardorff", tdew_tol=1e-6)
        out = passthrough_b200.run(df, p)
        assert out is df
        for oname, ename in (("t_day", "t_day"), ("tfmax", "tfmax"), ("pet", "pet"), ("tdew", "tdew")):
            util.assert_close(df[ename].values, ref_daily[ename][:, c], oname, context=f"passthrough run(df) {c}")
        np.testing.assert_array_equal(df["vapor_pressure"].values, inputs["given_vapor_pressure"][:, c])
    with pytest.raises(AssertionError):
        passthrough_b200.run(pd.DataFrame({"t_min": [1.0]}), {})


def test_full_size_properties_conus_hourly_slice():
    """Size-independent properties at the BASELINE grid width (482,560 cells), a few days:
    precipitation mass is conserved per record (roll only moves it), shortwave energy equals the
    daily totals, rel_humid <= 100, vapour pressure <= saturation, results independent of the
    day-tile decomposition."""
    import torch
    from metsim_b200 import synth
    from metsim_b200.engine import Engine
    rng = np.random.default_rng(synth.SEED)
    lat, lon = synth.grid_cells("conus1/16", rng)
    d = 6
    f = synth.make_forcing_torch(lat, lon, d, "2001-01-01", "cuda")
    eng = Engine(dict(time_step=60, prec_type="triangle", utc_offset=True))
    eng.load(**f)
    eng.build_solar()
    daily = eng.daily(want=("shortwave", "daylength"))
    full = eng.disaggregate(("temp", "prec", "shortwave", "vapor_pressure", "rel_humid"), 0, d, torch.float64)
    eng.check()
    torch.cuda.synchronize()
    tot = full["prec"].sum(0)
    ref = f["prec"].sum(0)
    assert torch.allclose(tot, ref, rtol=1e-11, atol=1e-11)
    # without the utc roll every day is self-contained: hourly energy == daily sw * daylength
    loc = Engine(dict(time_step=60, prec_type="triangle", utc_offset=False))
    loc.load(**f)
    loc.build_solar()
    dl = loc.daily(want=("shortwave", "daylength"))
    sub = loc.disaggregate(("shortwave", "prec"), 0, d, torch.float64)
    loc.check()
    torch.cuda.synchronize()
    e_sub = sub["shortwave"].reshape(d, 24, -1).sum(1) * 3600.0
    assert torch.allclose(e_sub, dl["shortwave"] * dl["daylength"], rtol=1e-11, atol=1e-9)
    p_sub = sub["prec"].reshape(d, 24, -1).sum(1)
    assert torch.allclose(p_sub, f["prec"], rtol=1e-11, atol=1e-11)
    assert float(full["rel_humid"].max()) <= 100.0
    assert bool((full["vapor_pressure"] >= 0).all())
    part = eng.disaggregate(("temp", "prec"), 2, 5, torch.float64)
    torch.cuda.synchronize()
    assert torch.equal(part["temp"], full["temp"][2 * 24:5 * 24])
    assert torch.equal(part["prec"], full["prec"][2 * 24:5 * 24])


def test_fast_path_full_width_against_general_path_and_properties():
    """The specialised interior-day kernels at the BASELINE grid width (482,560 cells, 9 days): every
    variable agrees with the event-driven general kernel (float64) to float32 rounding, precipitation
    mass per record is conserved, relative humidity <= 100, and 3-day calls give the same bits."""
    import torch
    from metsim_b200 import synth
    from metsim_b200.engine import Engine
    rng = np.random.default_rng(synth.SEED)
    lat, lon = synth.grid_cells("conus1/16", rng)
    d = 9
    f = synth.make_forcing_torch(lat, lon, d, "2001-03-28", "cuda")
    eng = Engine(dict(time_step=60, prec_type="triangle", utc_offset=True))
    eng.load(**f)
    eng.build_solar()
    eng.daily(want=())
    fast = eng.disaggregate(EXAMPLE_VARS, 0, d, torch.float32)                       # specialised path
    gen = eng.disaggregate(EXAMPLE_VARS + ("tskc",), 0, d, torch.float64)            # general path
    eng.check()
    torch.cuda.synchronize()
    for v in EXAMPLE_VARS:
        a, b = fast[v], gen[v].to(torch.float32)
        neq = (a != b)
        assert float(neq.double().mean()) < 1e-3, v
        scale = b.abs().clamp_min(util.FLOORS[v])
        assert float(((a - b).abs() / scale).max()) < 3e-7, v
    assert torch.allclose(fast["prec"].double().sum(0), f["prec"].sum(0), rtol=2e-6, atol=1e-5)
    assert float(fast["rel_humid"].max()) <= 100.0
    parts = [eng.disaggregate(EXAMPLE_VARS, d0, min(d, d0 + 3), torch.float32) for d0 in range(0, d, 3)]
    torch.cuda.synchronize()
    for v in EXAMPLE_VARS:
        assert torch.equal(torch.cat([p[v] for p in parts]), fast[v]), v


@pytest.mark.parametrize("grid,ts,prec_type,days", [
    ("conus1/16", 180, "mix", 10),        # BASELINE configs[3] shape (3-hourly, mix), full grid width
    ("global1/8", 15, "triangle", 3),     # configs[4] shape (15-minute), ~1M cells
    ("global0.5", 60, "triangle", 334),   # configs[1] grid, January-November (in December the
                                          # reference's t_pk/dur quirk makes 0/0 days, whose fill
                                          # does not conserve mass, metsim.py:278-284)
])
def test_full_size_properties_other_configs(grid, ts, prec_type, days):
    """Size-independent properties at BASELINE.json's full grid sizes: precipitation mass per
    record is conserved by the roll, shortwave energy per record matches the daily totals to a few
    per cent (with the utc roll the table rows shift independently of the record,
    disaggregate.py:656, and the wrapped last day lands in the first), relative humidity <= 100, 0 <= vapour pressure, longwave and
    air pressure inside physical bounds, and the result does not depend on the day-tile split."""
    import torch
    from metsim_b200 import synth
    from metsim_b200.engine import Engine
    rng = np.random.default_rng(synth.SEED)
    lat, lon = synth.grid_cells(grid, rng)
    f = synth.make_forcing_torch(lat, lon, days, "2001-01-01", "cuda")
    tpd = 1440 // ts
    eng = Engine(dict(time_step=ts, prec_type=prec_type, utc_offset=True))
    eng.load(**f)
    eng.build_solar()
    dl = eng.daily(want=("shortwave", "daylength"))
    n = lat.size
    tile = max(1, min(days, int(2e9 // (n * tpd * 8 * 6))))     # keep the f64 outputs under ~12 GB
    names = ("prec", "shortwave", "rel_humid", "vapor_pressure", "longwave", "air_pressure")
    tot_p = torch.zeros(n, dtype=torch.float64, device="cuda")
    tot_e = torch.zeros(n, dtype=torch.float64, device="cuda")
    first = None
    for d0 in range(0, days, tile):
        d1 = min(days, d0 + tile)
        out = eng.disaggregate(names, d0, d1, torch.float64)
        torch.cuda.synchronize()
        tot_p += out["prec"].sum(0)
        tot_e += out["shortwave"].sum(0) * (ts * 60.0)
        assert float(out["rel_humid"].max()) <= 100.0 and float(out["rel_humid"].min()) >= 0.0
        assert bool((out["vapor_pressure"] >= 0).all())
        assert 50.0 < float(out["longwave"].min()) and float(out["longwave"].max()) < 700.0
        assert 40.0 < float(out["air_pressure"].min()) and float(out["air_pressure"].max()) < 110.0
        if first is None:
            first = {k: v[: min(tpd, v.shape[0])].clone() for k, v in out.items()}
        del out
    eng.check()
    assert torch.allclose(tot_p, f["prec"].sum(0), rtol=1e-10, atol=1e-9)
    want_e = (dl["shortwave"] * dl["daylength"]).sum(0)
    big = want_e > 0.05 * want_e.max()     # polar-night cells: the rows rolled in dominate tiny totals
    assert float(((tot_e - want_e).abs() / want_e.clamp_min(1.0))[big].max()) < 0.05
    # a different tile split gives bit-identical values
    again = eng.disaggregate(names, 0, 1, torch.float64)
    torch.cuda.synchronize()
    for k in names:
        assert torch.equal(again[k][: first[k].shape[0]], first[k]), k


def test_daily_output_mode():
    """time_step = 1440 (metsim/metsim.py:779-792): no disaggregation; shortwave is converted from
    the daytime flux to the daily average flux.  Checked against the reference's daily columns."""
    from metsim_b200.engine import Engine
    from metsim_b200._lib import DAILY_VARS
    inputs, params, ref_daily, _, _ = util.load_golden(os.path.join(util.GOLDEN, "synth_daily.npz"))
    assert int(params["time_step"]) == 1440
    eng = Engine(params)
    eng.load(**{k: inputs.get(k) for k in util.INPUT_KEYS})
    res = eng.run(daily_vars=DAILY_VARS)
    assert set(res) == {"shortwave", "daily"}
    want = ref_daily["shortwave"] * ref_daily["daylength"] / 86400.0
    util.assert_close(res["shortwave"].cpu().numpy(), want, "sw_daily", context="daily mode")
    for oname, ename in util.DAILY.items():
        util.assert_close(res["daily"][ename].cpu().numpy(), ref_daily[ename], oname, context="daily mode")


# ---- the MetSim-object drop-in (Level 2 of the boundary): run_slice for a whole chunk -------------
class _DA:
    """Stand-in for the part of xarray.DataArray that MetSim.run_slice touches: values + dims."""
    def __init__(self, values, dims):
        self.values = np.asarray(values)
        self.dims = tuple(dims)
        assert self.values.ndim == len(self.dims)

    @property
    def ndim(self):
        return self.values.ndim


class _MS:
    """Stand-in for a MetSim object after load_inputs(): params, domain, met_data, state, output."""
    def __init__(self, params, domain, met_data, state, n_steps, mask_dims, mask_shape, out_dims=None):
        self.params, self.domain, self.met_data, self.state = params, domain, met_data, state
        self._n_steps, self._mdims, self._mshape, self._odims = n_steps, mask_dims, mask_shape, out_dims
        self.validated = False

    def _validate_setup(self):
        self.validated = True

    def setup_output(self):               # metsim/metsim.py:560-587: NaN-filled (time, *mask.dims)
        dims = self._odims or ("time",) + tuple(self._mdims)
        shape = (self._n_steps,) + tuple(self._mshape[self._mdims.index(d)] for d in dims[1:])
        prec = {"f4": np.float32, "f8": np.float64}[self.params.get("out_precision", "f4")]
        self.output = {v: _DA(np.full(shape, np.nan, dtype=prec), dims) for v in self.params["out_vars"]}


def test_metsim_run_slice_drop_in_on_the_bundled_test_domain():
    """examples/example_nc.conf's settings through run_slice_batched -- the body of the MetSim
    subclass of registry.batched_metsim_class -- on the bundled 9x9 test domain (test.nc +
    domain.nc + state_nc.nc, as stored in tests/golden/testdomain.npz), against what the unmodified
    reference produced for nine of its cells.  The stand-in hands over xarray-like arrays in a
    scrambled dimension order, a mask with holes (masked cells must stay NaN, :477-481 / :584) and
    unit conversions (:489-492)."""
    from metsim_b200.registry import run_slice_batched
    inputs, params, _, ref_sub, extra = util.load_golden(os.path.join(util.GOLDEN, "testdomain.npz"))
    cells = extra["ref_cells"]
    ny = nx = 9
    lat1, lon1 = np.unique(inputs["lat"]), np.unique(inputs["lon"])
    assert lat1.size == ny and lon1.size == nx
    grid = lambda a: a.reshape(a.shape[0], ny, nx)             # cells are latitude-row-major
    mask = np.ones((ny, nx), dtype=np.int32)
    holes = [c for c in (3, 17, 40, 41, 80) if c not in set(cells.tolist())]
    mask.ravel()[holes] = 0
    days, tpd = inputs["t_min"].shape[0], 1440 // int(params["time_step"])
    time = np.datetime64("1950-01-01") + np.arange(days)
    assert (inputs["doy"] == np.arange(1, days + 1)).all()
    domain = {"mask": _DA(mask, ("lat", "lon")), "lat": _DA(lat1, ("lat",)), "lon": _DA(lon1, ("lon",)),
              "elev": _DA(inputs["elev"].reshape(ny, nx).T, ("lon", "lat")),            # transposed on purpose
              "t_pk": _DA(grid(inputs["t_pk12"]), ("month", "lat", "lon")),
              "dur": _DA(np.moveaxis(grid(inputs["dur12"]), 0, 2), ("lat", "lon", "month"))}
    met = {"time": _DA(time, ("time",)),
           "t_min": _DA(grid(inputs["t_min"]), ("time", "lat", "lon")),
           "t_max": _DA(np.moveaxis(grid(inputs["t_max"]), 0, 1), ("lat", "time", "lon")),
           "prec": _DA(grid(inputs["prec"]), ("time", "lat", "lon")),
           "wind": _DA(np.transpose(grid(inputs["wind"]), (2, 1, 0)), ("lon", "lat", "time"))}
    state = {k: _DA(grid(inputs["state_" + k]), ("time", "lat", "lon")) for k in ("t_min", "t_max", "prec")}
    out_units = {"temp": "K", "prec": "mm s-1", "shortwave": "W m-2", "longwave": "W m-2",
                 "vapor_pressure": "hPa", "rel_humid": "fraction", "air_pressure": "Pa", "wind": "m s-1"}
    p = dict(params, method="mtclim_b200", out_precision="f8",
             out_vars={v: {"out_name": v, "units": u} for v, u in out_units.items()})
    ms = _MS(p, domain, met, state, days * tpd, ("lat", "lon"), (ny, nx), out_dims=("time", "lat", "lon"))
    run_slice_batched(ms)
    assert ms.validated
    conv = {"temp": (1.0, 273.15), "prec": (1.0 / (30 * 60.0), 0.0), "vapor_pressure": (0.01, 0.0),
            "rel_humid": (0.01, 0.0), "air_pressure": (1000.0, 0.0)}
    for v in out_units:
        got = ms.output[v].values.reshape(days * tpd, ny * nx)
        sc, of = conv.get(v, (1.0, 0.0))
        e = util.relerr(got[:, cells], ref_sub[v] * sc + of, util.FLOORS[v] * sc)
        assert e <= 2e-9, (v, e)
        assert np.isnan(got[:, holes]).all() and not np.isnan(np.delete(got, holes, axis=1)).any(), v
    # default out_precision (f4), default units, output dimensions in another order
    p4 = dict(params, method="mtclim_b200", out_vars={v: {"out_name": v} for v in ("temp", "shortwave")})
    ms4 = _MS(p4, domain, met, state, days * tpd, ("lat", "lon"), (ny, nx), out_dims=("time", "lon", "lat"))
    run_slice_batched(ms4)
    got = np.transpose(ms4.output["temp"].values, (0, 2, 1)).reshape(days * tpd, ny * nx)
    assert got.dtype == np.float32
    np.testing.assert_allclose(got[:, cells], ref_sub["temp"], rtol=0, atol=2e-6 * 40)
    # an output the time step does not offer, and wind without a wind forcing (KeyError in the reference)
    with pytest.raises(ValueError):
        run_slice_batched(_MS(dict(p4, out_vars={"pet": {}}), domain, met, state, days * tpd, ("lat", "lon"), (ny, nx)))
    met_nowind = {k: v for k, v in met.items() if k != "wind"}
    with pytest.raises(KeyError):
        run_slice_batched(_MS(dict(p4, out_vars={"wind": {}}), domain, met_nowind, state, days * tpd, ("lat", "lon"), (ny, nx)))


def test_metsim_run_slice_drop_in_daily_mode():
    """time_step = 1440 through the drop-in: the daily out_vars of metsim/metsim.py:660-666 come from
    the forcings and the daily pass, shortwave as the daily mean flux (:779-781) -- the round-1 code
    raised KeyError for every one of them but shortwave.  2-D lat / lon arrays (curvilinear-style
    domain) on a 2 x 5 grid."""
    from metsim_b200.registry import run_slice_batched
    inputs, params, ref_daily, _, _ = util.load_golden(os.path.join(util.GOLDEN, "synth_daily.npz"))
    days, n = inputs["t_min"].shape
    ny, nx = 2, n // 2
    g2 = lambda a: a.reshape(ny, nx)
    g3 = lambda a: a.reshape(a.shape[0], ny, nx)
    start = np.datetime64("2001-01-01")
    doy0 = int(inputs["doy"][0])
    time = start + (doy0 - 1) + np.arange(days)
    from metsim_b200.engine import calendar_vectors
    d_chk, m_chk = calendar_vectors(str(time[0]), days)
    if not ((d_chk == inputs["doy"]).all() and (m_chk == inputs["month"]).all()):
        pytest.skip("fixture calendar is not a 2001 record")
    domain = {"mask": _DA(np.ones((ny, nx)), ("y", "x")), "lat": _DA(g2(inputs["lat"]), ("y", "x")),
              "lon": _DA(g2(inputs["lon"]), ("y", "x")), "elev": _DA(g2(inputs["elev"]), ("y", "x"))}
    met = {"time": _DA(time, ("time",))}
    for k in ("t_min", "t_max", "prec", "wind"):
        met[k] = _DA(g3(inputs[k]), ("time", "y", "x"))
    state = {k: _DA(g3(inputs["state_" + k]), ("time", "y", "x")) for k in ("t_min", "t_max", "prec")}
    out_vars = {v: {} for v in ("t_min", "t_max", "t_day", "prec", "vapor_pressure", "shortwave", "tskc",
                                "pet", "wind", "daylength")}
    out_vars["t_max"] = {"units": "K"}
    p = dict(params, out_precision="f8", out_vars=out_vars)
    ms = _MS(p, domain, met, state, days, ("y", "x"), (ny, nx))
    run_slice_batched(ms)
    flat = lambda v: ms.output[v].values.reshape(days, n)
    np.testing.assert_array_equal(flat("t_min"), inputs["t_min"])
    np.testing.assert_array_equal(flat("prec"), inputs["prec"])
    np.testing.assert_array_equal(flat("wind"), inputs["wind"])
    np.testing.assert_allclose(flat("t_max"), inputs["t_max"] + 273.15, rtol=1e-15)
    util.assert_close(flat("shortwave"), ref_daily["shortwave"] * ref_daily["daylength"] / 86400.0, "sw_daily")
    for v, o in (("t_day", "t_day"), ("vapor_pressure", "vp_daily"), ("tskc", "tskc_daily"), ("pet", "pet"),
                 ("daylength", "daylength")):
        util.assert_close(flat(v), ref_daily[v], o, context="drop-in daily mode")


@pytest.mark.parametrize("rows", [2, 4, 8])
def test_patch_cell_order_for_masked_grids_changes_nothing_but_the_order(rows):
    """The engine-defined cell order for masked grids (shard.patch_order): inputs are gathered into
    patch order on the device, the kernels run on it, results equal the oracle's in the caller's order
    (Engine.run) and, in engine order, are the same numbers permuted (what a writer's cell_index absorbs)."""
    import torch
    from metsim_b200 import shard, synth
    from metsim_b200._lib import SUB_VARS
    from metsim_b200.engine import Engine
    from oracle import oracle_lib
    f, params = synth.make_config("global_half_deg", max_cells=1500, n_days=40)
    want = oracle_lib.run(params, **util.oracle_kwargs(f))
    order = shard.patch_order(f["lat"], f["lon"], rows)
    assert not np.array_equal(order, np.arange(order.size))
    eng = Engine(params)
    eng.load(order=order, **{k: f.get(k) for k in util.INPUT_KEYS})
    res = eng.run(out_vars=SUB_VARS, out_dtype=torch.float64)
    for v in util.SUB:
        util.assert_close(res[v].cpu().numpy(), want[v], v, context=f"patch order rows={rows}")
    np.testing.assert_array_equal(res["daily"]["n_iter"].cpu().numpy(), want["n_iter"])
    raw = eng.run(out_vars=("temp", "prec"), out_dtype=torch.float64, engine_order=True)
    for v in ("temp", "prec"):
        np.testing.assert_array_equal(raw[v].cpu().numpy(), res[v].cpu().numpy()[:, order])
