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
* :func:`batched_metsim_class` returns a ``MetSim`` subclass whose ``run_slice`` hands the whole
  chunk to the batched engine (solar tables, daily pass and disaggregation all on the GPU).  It
  needs the reference's xarray stack, which is not part of this repo's requirements.
"""
from __future__ import annotations

METHOD_NAME = "mtclim_b200"
PASSTHROUGH_NAME = "passthrough_b200"


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


def batched_metsim_class():
    """``MetSim`` subclass running whole chunks through :class:`metsim_b200.engine.Engine`."""
    import numpy as np
    import metsim.constants as cnst
    from metsim.metsim import MetSim
    from metsim.units import converters  # noqa: F401  (units are applied on the device)

    from .engine import Engine

    register(MetSim.methods)

    class MetSimB200(MetSim):
        def run_slice(self):  # replaces metsim/metsim.py:456-493
            self._validate_setup()
            self.disagg = int(self.params["time_step"]) < cnst.MIN_PER_DAY
            self.setup_output()
            mask = np.asarray(self.domain["mask"].values)
            active = np.argwhere(mask > 0)
            if active.size == 0:
                return
            dims = self.domain["mask"].dims
            sel = tuple(active[:, i] for i in range(active.shape[1]))
            grab = lambda da: np.asarray(da.values, dtype=np.float64)[(slice(None),) + sel]
            md, st, dom = self.met_data, self.state, self.domain
            lat = np.broadcast_to(dom["lat"].values.reshape(
                [-1 if d == "lat" else 1 for d in dims]) if dom["lat"].ndim == 1 else dom["lat"].values,
                mask.shape)[sel]
            lon = np.broadcast_to(dom["lon"].values.reshape(
                [-1 if d == "lon" else 1 for d in dims]) if dom["lon"].ndim == 1 else dom["lon"].values,
                mask.shape)[sel]
            times = md["time"].to_index()
            eng = Engine(self.params)
            kw = {}
            if self.params["prec_type"].upper() in ("TRIANGLE", "MIX"):
                kw = dict(t_pk12=np.asarray(dom["t_pk"].values, float)[(slice(None),) + sel],
                          dur12=np.asarray(dom["dur"].values, float)[(slice(None),) + sel])
            eng.load(lat=lat, lon=lon, elev=np.asarray(dom["elev"].values, float)[sel],
                     doy=times.dayofyear.values, month=times.month.values,
                     t_min=grab(md["t_min"]), t_max=grab(md["t_max"]), prec=grab(md["prec"]),
                     wind=grab(md["wind"]) if "wind" in md else None,
                     **{"given_" + k: grab(md[k]) for k in
                        ("shortwave", "vapor_pressure", "tskc", "longwave")
                        if "passthrough" in str(self.params["method"]) and k in md},
                     state_t_min=grab(st["t_min"]), state_t_max=grab(st["t_max"]),
                     state_prec=grab(st["prec"]), **kw)
            out_vars = list(self.params["out_vars"])
            units = {v: self.params["out_vars"][v]["units"] for v in out_vars}
            res = eng.run(out_vars=out_vars, units=units)
            for v in out_vars:
                block = res[v].cpu().numpy()
                target = self.output[v].values
                target[(slice(None),) + sel] = block

    return MetSimB200
