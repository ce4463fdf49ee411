"""Registration of the CUDA path in the reference's method registry.

``MetSim.methods`` is a class-level dict of modules (metsim/metsim.py:139) selected by
``params['method']`` (:143, :462) and never validated (:620-684), so adding a key is all a config
file needs: ``method = mtclim_b200``.  ``disaggregate.py`` only special-cases ``'passthrough'``
(:79, :644-648), so the new name is safe there.

Two levels are offered:

* :func:`register` adds ``'mtclim_b200'`` -> :mod:`metsim_b200.methods.mtclim_b200` and
  ``'passthrough_b200'`` -> :mod:`metsim_b200.methods.passthrough_b200`, whose
  ``run(df, params)`` is called per cell by the unmodified driver (daily kernels on the GPU,
  the reference's own ``disaggregate`` afterwards; note that ``disaggregate.py:79`` keys its
  full-day shortwave averaging on the literal name ``'passthrough'``, so with the unmodified
  driver ``passthrough_b200`` behaves like ``sw_averaging = daylight`` there -- the batched
  engine implements both modes).
* :func:`run_slice_batched` is the body of ``MetSim.run_slice`` (metsim/metsim.py:456-493) with the
  cell loop replaced by ONE call of the batched engine (solar tables, daily pass and
  disaggregation all on the GPU).  It only touches the attributes ``run_slice`` itself touches
  (``params``, ``domain``, ``met_data``, ``state``, ``output``) through the small part of the
  xarray interface it needs (``.values``, ``.dims``), so it runs -- and is tested -- against
  stand-ins where the xarray stack is absent.  :func:`batched_metsim_class` wraps it into a
  ``MetSim`` subclass for installations that have the reference importable.
"""
from __future__ import annotations

import numpy as np

METHOD_NAME = "mtclim_b200"
PASSTHROUGH_NAME = "passthrough_b200"

# metsim/metsim.py:660-666
DAILY_OUT_VARS = ("t_min", "t_max", "t_day", "prec", "vapor_pressure", "shortwave", "tskc", "pet",
                  "wind", "daylength")
SUBDAILY_OUT_VARS = ("temp", "prec", "shortwave", "vapor_pressure", "air_pressure", "rel_humid",
                     "spec_humid", "longwave", "tskc", "wind")


def register(methods: dict | None = None) -> dict:
    """Insert the method module into ``methods`` (default: ``metsim.metsim.MetSim.methods``)."""
    from .methods import mtclim_b200, passthrough_b200
    if methods is None:
        try:
            from metsim.metsim import MetSim  # the reference package, if importable
        except Exception as exc:  # xarray / dask / netCDF4 missing, or metsim not installed
            raise ImportError(
                "the reference `metsim` package is not importable here; pass the registry dict "
                "explicitly (register(MetSim.methods))") from exc
        methods = MetSim.methods
    methods[METHOD_NAME] = mtclim_b200
    methods[PASSTHROUGH_NAME] = passthrough_b200
    return methods


def _cells(da, lead: str, mask_dims: tuple, sel: tuple) -> np.ndarray:
    """``[lead][active cells]`` float64 view of a DataArray-like whose dimensions are ``lead`` plus
    the mask's dimensions in ANY order (the reference selects by name, ``isel(**locs)``)."""
    vals = np.asarray(da.values, dtype=np.float64)
    dims = tuple(da.dims)
    want = (lead,) + tuple(mask_dims)
    if sorted(dims) != sorted(want):
        raise ValueError(f"expected dimensions {want}, got {dims}")
    vals = np.transpose(vals, [dims.index(d) for d in want])
    return np.ascontiguousarray(vals[(slice(None),) + sel])


def _per_cell(ds, name: str, mask_dims: tuple, mask_shape: tuple, sel: tuple) -> np.ndarray:
    """lat / lon / elev of the active cells: 1-D coordinate or an array over the mask's dimensions."""
    da = ds[name]
    vals = np.asarray(da.values, dtype=np.float64)
    dims = tuple(da.dims)
    if vals.ndim == 1:
        axis = mask_dims.index(dims[0])
        shape = [1] * len(mask_dims)
        shape[axis] = -1
        return np.broadcast_to(vals.reshape(shape), mask_shape)[sel].copy()
    vals = np.transpose(vals, [dims.index(d) for d in mask_dims])
    return vals[sel].copy()


def _calendar(time_values, calendar: str = "standard"):
    """1-based day of year and month of a daily time axis (numpy datetime64 or anything ``np.asarray``
    turns into it).  A noleap / all_leap / 360_day axis arrives labelled with Gregorian dates
    (``CFTimeIndex.to_datetimeindex()`` in the reference, metsim/io.py:281-283): consecutive records
    are then up to three calendar days apart (28 February -> 1 March of a leap year) and the day of
    year is the label's, as ``dates.dayofyear`` is for the reference (io.py:291)."""
    t = np.asarray(time_values)
    if not np.issubdtype(t.dtype, np.datetime64):
        raise NotImplementedError("the batched path needs a datetime64 time axis; "
                                  f"got a time axis of dtype {t.dtype}")
    days = t.astype("datetime64[D]")
    step = np.diff(days).astype(np.int64)
    uniform = str(calendar or "standard").lower() in ("noleap", "365_day", "all_leap", "366_day", "360_day")
    ok = ((step >= 1) & (step <= 3)).all() if uniform else (step == 1).all()
    if not ok:
        raise ValueError("the forcing time axis must be daily and contiguous")
    years = days.astype("datetime64[Y]")
    doy = (days - years.astype("datetime64[D]")).astype(np.int64) + 1
    month = days.astype("datetime64[M]").astype(np.int64) % 12 + 1
    return doy.astype(np.int32), month.astype(np.int32)


def run_slice_batched(ms, device="cuda:0") -> None:
    """Body of ``MetSim.run_slice`` (metsim/metsim.py:456-493) for the whole chunk at once.

    ``ms`` needs ``params`` (the MetSim parameter dict), ``domain`` (``mask``, ``lat``, ``lon``,
    ``elev`` and, for triangle / mix precipitation, ``t_pk`` / ``dur`` over ``month``), ``met_data``
    (``time`` plus the daily forcings over ``time``), ``state`` (``t_min``, ``t_max``, ``prec`` over
    the 90 state days) and ``output`` (one ``(time, *mask.dims)`` array per ``out_vars`` entry,
    NaN-initialised by ``setup_output``); ``_validate_setup`` / ``setup_output`` are called when
    present.  Mirrors the reference: masked cells are skipped and stay NaN (:477-481), units are
    converted per ``out_vars[v]['units']`` (:489-492), ``time_step = 1440`` writes the daily columns
    with shortwave as the daily mean flux (:779-792).  The values do not depend on the output-time
    bookkeeping of wrap_run_cell (:764-795): it relabels / keeps every step of the record."""
    from .engine import Engine, unit_scale_offset
    params = ms.params
    if hasattr(ms, "_validate_setup"):
        ms._validate_setup()
    ts = int(params["time_step"])
    daily_mode = ts >= 1440
    if hasattr(ms, "setup_output"):
        ms.disagg = not daily_mode
        ms.setup_output()
    out_vars = list(params["out_vars"])
    allowed = DAILY_OUT_VARS if daily_mode else SUBDAILY_OUT_VARS
    for v in out_vars:
        if v not in allowed:
            raise ValueError(f"Cannot output variable {v} at timestep {ts}")      # metsim.py:667-670
    dom, md, st = ms.domain, ms.met_data, ms.state
    mask_da = dom["mask"]
    mask = np.asarray(mask_da.values)
    mdims = tuple(mask_da.dims)
    active = np.argwhere(mask > 0)            # row-major, like np.ndenumerate (:477)
    if active.size == 0:
        return
    sel = tuple(active[:, i] for i in range(active.shape[1]))
    doy, month = _calendar(md["time"].values)
    method = str(params.get("method", "mtclim")).lower()
    has = lambda ds, k: k in ds
    kw = {}
    if str(params.get("prec_type", "uniform")).upper() in ("TRIANGLE", "MIX"):
        kw["t_pk12"] = _cells(dom["t_pk"], "month", mdims, sel)
        kw["dur12"] = _cells(dom["dur"], "month", mdims, sel)
    if "passthrough" in method:
        for k in ("shortwave", "vapor_pressure", "tskc", "longwave"):
            if has(md, k):
                kw["given_" + k] = _cells(md[k], "time", mdims, sel)
    wind = _cells(md["wind"], "time", mdims, sel) if has(md, "wind") else None
    if "wind" in out_vars and wind is None:
        raise KeyError("wind")                # df['wind'] in the reference (:492)
    forc = {k: _cells(md[k], "time", mdims, sel) for k in ("t_min", "t_max", "prec")}
    eng = Engine(params, device)
    eng.load(lat=_per_cell(dom, "lat", mdims, mask.shape, sel),
             lon=_per_cell(dom, "lon", mdims, mask.shape, sel),
             elev=_per_cell(dom, "elev", mdims, mask.shape, sel),
             doy=doy, month=month, wind=wind, **forc,
             state_t_min=_cells(st["t_min"], "time", mdims, sel),
             state_t_max=_cells(st["t_max"], "time", mdims, sel),
             state_prec=_cells(st["prec"], "time", mdims, sel), **kw)
    units = {v: (params["out_vars"][v] or {}).get("units") for v in out_vars}

    def put(v, block):
        """rows [time][active cells] -> ms.output[v], whose dimensions are (time, *mask.dims)"""
        target = ms.output[v]
        tdims = tuple(target.dims)
        if tdims[0] != "time" or sorted(tdims[1:]) != sorted(mdims):
            raise ValueError(f"output[{v!r}] has dimensions {tdims}, expected ('time', *{mdims})")
        order = [mdims.index(d) for d in tdims[1:]]
        vals = target.values
        if block.shape[0] != vals.shape[0]:
            raise ValueError(f"output[{v!r}] holds {vals.shape[0]} time steps, the run produced {block.shape[0]}")
        vals[(slice(None),) + tuple(sel[i] for i in order)] = block

    if daily_mode:
        res = eng.run(daily_vars=("t_day", "vapor_pressure", "tskc", "pet", "daylength"))
        daily = res["daily"]
        for v in out_vars:
            if v in forc:
                block = forc[v]
            elif v == "wind":
                block = wind
            elif v == "shortwave":
                block = res["shortwave"].cpu().numpy()       # daily mean flux (:779-781)
            else:
                block = daily[v].cpu().numpy()
            sc, of = unit_scale_offset(v, units[v], ts)
            put(v, block * sc + of if (sc, of) != (1.0, 0.0) else block)
        return
    import torch
    prec = str(params.get("out_precision", "f4"))
    res = eng.run(out_vars=out_vars, units=units,
                  out_dtype=torch.float64 if prec == "f8" else torch.float32)
    for v in out_vars:
        put(v, res[v].cpu().numpy())


def batched_metsim_class():
    """``MetSim`` subclass whose ``run_slice`` is :func:`run_slice_batched` (needs the reference's
    xarray stack to be importable)."""
    from metsim.metsim import MetSim

    register(MetSim.methods)

    class MetSimB200(MetSim):
        def run_slice(self):  # replaces metsim/metsim.py:456-493
            run_slice_batched(self)

    return MetSimB200
