"""Minimal driver: a reference configuration file in, the reference's output files out, with the
cell loop replaced by ONE batched call of the CUDA path.

SURVEY 8(b): "a ``MetSim`` subclass overrides ``run_slice`` when the xarray stack is importable
(:func:`metsim_b200.registry.batched_metsim_class`); otherwise our own minimal driver reads the same
ini / yaml keys".  This is that driver.  It mirrors the reference's public surface for the path

    ms = MetSimB200(read_config("example_nc.conf"));  ms.run();  ms.open_output()
    python -m metsim_b200.driver example_nc.conf [-n N] [-s SCHED] [-v]          (metsim/cli/ms.py:40-62)

and follows the reference for everything a user can observe:

* configuration: ``[MetSim]``, ``[forcing_vars]``, ``[domain_vars]``, ``[state_vars]``, ``[chunks]``,
  ``[constant_vars]``, ``out_vars`` as a list or a section, ini or yaml (metsim/io.py:41-173) and the
  defaults of ``MetSim.params`` (metsim/metsim.py:140-171);
* inputs: NetCDF domain / state / forcing selected at the domain's coordinates with the
  ``- 11:59:59, round to the day`` time normalisation (io.py:264-303), VIC ascii and VIC-4 binary
  per-cell files named ``<name>_<lat>_<lon>`` (io.py:214-262, :324-369), ``constant_vars``
  (metsim.py:300-305), the 90 state days before ``start`` (metsim.py:268-271, :589-596);
* checks and their messages (``_validate_setup`` metsim.py:620-684, ``_validate_force_times``
  :239-276, the ``t_max < t_min`` ValueError :612-614);
* outputs: one ``<out_prefix>_<YYYYMMDD>-<YYYYMMDD>.nc`` per ``out_freq`` group (:512-557,
  :686-690) holding ``(time, *mask.dims)`` variables named ``out_name`` in ``out_precision``, NaN
  outside the mask, unit-converted (units.py), the ``i4`` time axis in minutes since 2000-01-01 and
  the variable / global attributes of ``setup_netcdf_output`` (:363-436) -- as NetCDF classic
  (CDF-2 / CDF-5, ``ncwriter.py``) because no NetCDF-4 library exists in this image.

What it does NOT do: dask scheduling (``scheduler`` and ``[chunks]`` are accepted and recorded;
``-n N`` spreads the active cells over N GPUs as contiguous ranges, one batched call each, see
:meth:`MetSimB200.run`), calendars other than standard / noleap / all_leap / 360_day (the reference
needs cftime for every non-standard one) and in-memory ``forcing_fmt = data``.  No arithmetic of the path happens
here: every number written comes from ``msb_run_host`` (CUDA); without a GPU :meth:`MetSimB200.run`
raises.
"""
from __future__ import annotations

import argparse
import getpass
import glob
import json
import logging
import os
import sys
import time as _time
from collections import OrderedDict
from configparser import ConfigParser

import numpy as np

from . import ncio
from .ncio import Dataset, Var
from .registry import DAILY_OUT_VARS, SUBDAILY_OUT_VARS, _calendar, _cells, _per_cell

DEFAULT_TIME_UNITS = "minutes since 2000-01-01 00:00:00.0"         # metsim.py:71
_EPOCH = np.datetime64("2000-01-01T00:00:00", "s")

# variable attributes of the output files (metsim.py:76-124): units, long_name, standard_name
_VAR_ATTRS = {
    "pet": ("mm timestep-1", "potential evaporation", "water_potential_evaporation_flux"),
    "prec": ("mm timestep-1", "precipitation", "precipitation_flux"),
    "shortwave": ("W m-2", "shortwave radiation", "surface_downwelling_shortwave_flux"),
    "longwave": ("W m-2", "longwave radiation", "surface_downwelling_longwave_flux"),
    "t_max": ("C", "maximum daily air temperature", "daily_maximum_air_temperature"),
    "t_min": ("C", "minimum daily air temperature", "daily_minimum_air_temperature"),
    "temp": ("C", "air temperature", "air_temperature"),
    "vapor_pressure": ("Pa", "vapor pressure", "vapor_pressure"),
    "air_pressure": ("kPa", "air pressure", "air_pressure"),
    "tskc": ("fraction", "cloud fraction", "cloud_fraction"),
    "rel_humid": ("%", "relative humidity", "relative_humidity"),
    "daylength": ("s", "daylength", "length of day"),
    "spec_humid": ("g g-1", "specific humidity", "specific_humidity"),
    "wind": ("m s-1", "wind speed", "wind_speed"),
}


def available_outputs() -> dict:
    """``metsim.metsim.available_outputs`` (metsim.py:126-127): default out_name and units."""
    return {n: {"out_name": n, "units": a[0]} for n, a in _VAR_ATTRS.items()}


DEFAULT_OUTPUTS = ("temp", "prec", "shortwave", "longwave", "vapor_pressure", "rel_humid")   # metsim.py:128-129


def default_params() -> dict:
    """``MetSim.params`` (metsim.py:140-171)."""
    av = available_outputs()
    return {
        "period_ending": False, "is_worker": False, "method": "mtclim", "domain": "", "state": "",
        "out_dir": "", "out_prefix": "forcing", "start": "forcing", "stop": "forcing",
        "forcing_fmt": "netcdf", "time_step": -1, "calendar": "standard", "prec_type": "uniform",
        "out_precision": "f4", "verbose": 0, "sw_prec_thresh": 0.0, "utc_offset": False,
        "lw_cloud": "cloud_deardorff", "lw_type": "prata", "tdew_tol": 1e-6,
        "tmax_daylength_fraction": 0.67, "rain_scalar": 0.75, "tday_coef": 0.45,
        "lapse_rate": 0.0065, "out_vars": {n: av[n] for n in DEFAULT_OUTPUTS}, "out_freq": None,
        "chunks": {}, "scheduler": "distributed", "num_workers": 1,
    }


# ------------------------------------------------------------------------------ configuration
def _invert(d) -> OrderedDict:
    return OrderedDict((v, k) for k, v in d.items())


def check_config_type(config_file: str) -> str:
    """'ini' or 'yaml' from the first non-comment line (io.py:45-59)."""
    line = "#"
    with open(config_file) as f:
        while line.strip().startswith("#"):
            line = f.readline()
            if line == "":
                break
    line = line.split("#")[0].strip()
    return "yaml" if line == "MetSim:" else "ini"


def _forcing_files(forcing):
    if isinstance(forcing, str) and os.path.isdir(forcing):
        return [os.path.join(forcing, fn) for fn in next(os.walk(forcing))[2]]
    return forcing


def read_config(config_file: str, scheduler: str = "distributed", num_workers: int = 1,
                verbose: bool = False) -> dict:
    """The parameter dict the reference's ``io.read_config`` builds (io.py:62-173) from an ini or
    yaml configuration file; the keyword arguments are the command-line options of ``ms``."""
    av = available_outputs()
    if check_config_type(config_file) == "yaml":
        import yaml
        with open(config_file) as f:
            config = yaml.load(f, Loader=yaml.SafeLoader)
        conf = OrderedDict(config["MetSim"])
        for sec in ("forcing_vars", "domain_vars", "state_vars"):
            conf[sec] = _invert(OrderedDict(config[sec]))
        conf["out_vars"] = OrderedDict((k, dict(v or {})) for k, v in config["out_vars"].items())
        for k, v in conf["out_vars"].items():
            v.setdefault("units", av[k]["units"])
            v.setdefault("out_name", k)
        conf["chunks"] = OrderedDict(config.get("chunks") or {})
        if "constant_vars" in config:
            conf["constant_vars"] = OrderedDict(config["constant_vars"])
    else:
        config = ConfigParser()
        config.optionxform = str
        if not config.read(config_file):
            raise FileNotFoundError(config_file)
        conf = OrderedDict(config["MetSim"])
        conf["forcing_vars"] = OrderedDict(config["forcing_vars"])
        if conf.get("forcing_fmt", "netcdf") != "binary":     # binary: name -> "<scale> <signedness>"
            conf["forcing_vars"] = _invert(conf["forcing_vars"])
        conf["domain_vars"] = _invert(OrderedDict(config["domain_vars"]))
        conf["state_vars"] = _invert(OrderedDict(config["state_vars"]))
        conf["chunks"] = OrderedDict(config["chunks"]) if "chunks" in config else OrderedDict()
        if "constant_vars" in config:
            conf["constant_vars"] = OrderedDict(config["constant_vars"])
        for key in ("utc_offset", "period_ending"):
            conf[key] = str(conf.get(key, "")).strip().lower() == "true"
        conf["prec_type"] = conf.get("prec_type", "uniform")
        out_vars = OrderedDict()
        if "out_vars" in conf:                                  # list variant
            for ov in json.loads(conf["out_vars"].replace("'", '"').split("#")[0]):
                out_vars[ov] = dict(av[ov])
        if "out_vars" in config:                                # section variant: varname = out_name
            for varname, outname in config["out_vars"].items():
                out_vars[varname] = dict(av[varname], out_name=outname)
        conf["out_vars"] = out_vars
    conf.update({"calendar": conf.get("calendar", "standard"), "scheduler": scheduler,
                 "num_workers": num_workers,
                 "verbose": logging.DEBUG if verbose else logging.INFO,
                 "forcing": _forcing_files(conf["forcing"]),
                 "out_dir": os.path.abspath(conf["out_dir"])})
    return {k: v for k, v in conf.items() if v != []}


def _date_range(start, stop, calendar: str = "standard") -> np.ndarray:
    """Daily ``date_range(start, stop, calendar=...)`` (metsim/datetime.py:11-72): every day of the
    calendar between the two dates, labelled with Gregorian dates."""
    units = "days since 1900-01-01"
    a, b = (ncio.encode_time(np.asarray(x, dtype="datetime64[s]"), units, calendar) for x in (start, stop))
    return ncio.decode_time(np.arange(int(round(float(a))), int(round(float(b))) + 1), units, calendar)


def _parse_date(value):
    """``start`` / ``stop``: 'YYYY/M/D', anything pandas parses, or a datetime (metsim.py:241-250)."""
    if isinstance(value, str) and "/" in value:
        y, m, d = (int(x) for x in value.split(":")[0].split("/"))
        return np.datetime64(f"{y:04d}-{m:02d}-{d:02d}", "s")
    import pandas as pd
    return np.datetime64(pd.to_datetime(value).to_datetime64(), "s")


# ------------------------------------------------------------------------------ input readers
def _read_netcdf(handle, var_dict=None) -> Dataset:
    """``io.read_netcdf`` up to the time normalisation (io.py:264-286): open (a glob is concatenated
    along time), keep and rename ``var_dict``'s variables, put the time axis on whole days."""
    paths = sorted(glob.glob(handle)) if isinstance(handle, str) and "*" in handle else [handle]
    if not paths:
        raise FileNotFoundError(handle)
    parts = [ncio.open_dataset(p) for p in paths]
    ds = parts[0]
    if len(parts) > 1:
        for name, v in ds.items():
            if "time" in v.dims:
                ax = v.dims.index("time")
                v.values = np.concatenate([p[name].values for p in parts], axis=ax)
        ds.dims["time"] = ds["time"].values.shape[0]
    if var_dict is not None:
        missing = [k for k in var_dict if k not in ds]
        if missing:
            raise KeyError(f"{handle}: no variable(s) {missing}")
        keep = Dataset(dims=ds.dims, attrs=ds.attrs)
        for file_name, name in var_dict.items():
            keep[name] = ds[file_name]
        for d in ds.dims:                                        # coordinates travel with the selection
            if d in ds and d not in keep and ds[d].dims == (d,):
                keep[d] = ds[d]
        ds = keep
    if "time" in ds and ds["time"].values.dtype.kind == "M":
        t = ds["time"].values.astype("datetime64[s]") - np.timedelta64(11 * 3600 + 59 * 60 + 59, "s")
        day = np.timedelta64(86400, "s")
        ds["time"].values = ((t - _EPOCH + day // 2) // day) * day + _EPOCH   # round to the nearest day
    return ds


def _select_time(ds: Dataset, start, stop) -> Dataset:
    """``ds.sel(time=slice(start, stop))`` (io.py:288-291), both ends included."""
    t = ds["time"].values
    keep = (t >= start) & (t <= stop)
    out = Dataset(dims=dict(ds.dims), attrs=ds.attrs)
    for name, v in ds.items():
        if "time" in v.dims:
            out[name] = Var(np.compress(keep, v.values, axis=v.dims.index("time")), v.dims, v.attrs)
        else:
            out[name] = v
    out.dims["time"] = int(keep.sum())
    return out


def _select_domain(ds: Dataset, domain: Dataset) -> Dataset:
    """``ds.sel({k: domain[k] for k in domain.dims if k in ds.dims})`` (io.py:274-277): the file's
    grid may be larger than / ordered differently from the domain's; labels must match exactly."""
    out = Dataset(dims=dict(ds.dims), attrs=ds.attrs)
    out.update(ds)
    for d in domain.dims:
        if d not in ds.dims or d not in domain or d not in ds or domain[d].dims != (d,):
            continue
        want, have = domain[d].values, ds[d].values
        if want.shape == have.shape and np.array_equal(want, have):
            continue
        pos = {v: i for i, v in enumerate(have.tolist())}
        try:
            idx = np.array([pos[v] for v in want.tolist()], dtype=np.int64)
        except KeyError as exc:
            raise KeyError(f"coordinate {exc} of dimension {d!r} of the domain is not in the file") from None
        for name, v in list(out.items()):
            if d in v.dims:
                out[name] = Var(np.take(v.values, idx, axis=v.dims.index(d)), v.dims, v.attrs)
        out.dims[d] = idx.size
    return out


def _coordinate(domain: Dataset, d: str) -> np.ndarray:
    """Values of dimension ``d``; a dimension without a coordinate variable counts from zero
    (io.py:278-284)."""
    if d in domain and domain[d].dims == (d,):
        return domain[d].values
    return np.arange(domain.dims[d], dtype=np.int32)


def read_domain(params: dict) -> Dataset:
    ds = _read_netcdf(params["domain"], params.get("domain_vars"))
    if "mask" not in ds:
        raise KeyError("the domain needs a `mask` variable ([domain_vars])")
    return ds


def _cell_files(params: dict, domain: Dataset):
    """(path, index into the mask) of every per-cell VIC forcing file that names an active cell of a
    (lat, lon) domain; anything else is skipped like the reference does (io.py:239-249)."""
    mask = domain["mask"]
    lat, lon = domain["lat"].values.tolist(), domain["lon"].values.tolist()
    for job in params["forcing"]:
        try:
            _, la, lo = os.path.basename(job).split("_")[-3:]
            i, j = lat.index(float(la)), lon.index(float(lo))
        except ValueError:
            continue
        loc = tuple({"lat": i, "lon": j}[d] for d in mask.dims)
        if mask.values[loc] > 0:
            yield job, loc


def _read_vic(params: dict, domain: Dataset, times: np.ndarray) -> Dataset:
    """``io.process_vic`` + ``read_ascii`` / ``read_binary`` (io.py:214-262, :324-369): one file per
    cell starting at ``start``; cells without a file stay NaN."""
    fmt = params["forcing_fmt"]
    var_dict = params["forcing_vars"]
    names = list(var_dict.keys())
    mask = domain["mask"]
    n_days = times.size
    shape = (n_days,) + mask.values.shape
    data = {v: np.full(shape, np.nan) for v in names}
    if fmt == "binary":
        kinds = {"signed": "<i2", "unsigned": "<u2"}
        rec = np.dtype([(v, kinds[str(var_dict[v]).split()[-1]]) for v in names])
        scales = [float(str(var_dict[v]).split()[0]) for v in names]
    for job, loc in _cell_files(params, domain):
        if fmt == "binary":
            raw = np.fromfile(job, dtype=rec, count=n_days)
            cols = [raw[v].astype(np.float64) / s for v, s in zip(names, scales)]
        else:
            tab = np.loadtxt(job, ndmin=2, max_rows=n_days)
            cols = [tab[:, k] for k in range(len(names))]
        for v, col in zip(names, cols):
            data[v][(slice(0, col.size),) + loc] = col
    ds = Dataset(dims={"time": n_days, **{d: domain.dims[d] for d in mask.dims}})
    ds["time"] = Var(times, ("time",))
    for v in names:
        ds[v] = Var(data[v], ("time",) + mask.dims)
    return ds


def _pinned(shape, dtype) -> np.ndarray:
    """Page-locked host rows for the output ring (true DMA for the device-to-host copies of
    ``msb_run_host``); plain memory where no CUDA device is visible."""
    try:
        import torch
        if torch.cuda.is_available():
            tdt = {np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64}[np.dtype(dtype)]
            return torch.empty(shape, dtype=tdt, pin_memory=True).numpy()
    except (ImportError, RuntimeError):
        pass
    return np.empty(shape, dtype=dtype)


# ------------------------------------------------------------------------------ the driver
class MetSimB200:
    """``MetSim(params)`` for the CUDA path (metsim/metsim.py:131-493): same parameter dict, same
    lazily-read ``domain`` / ``met_data`` / ``state``, same ``run()`` / ``open_output()``."""

    def __init__(self, params: dict, device: int = 0):
        self.params = default_params()
        self.params.update(params)
        self.device = int(device)
        self._domain = self._met_data = self._state = None
        self._ring = None
        self.logger = logging.getLogger("metsim_b200")
        self.logger.setLevel(self.params["verbose"] or logging.INFO)
        self.var_attrs = {k: {"units": a[0], "long_name": a[1], "standard_name": a[2]}
                          for k, a in _VAR_ATTRS.items()}
        self.time_attrs = {}
        self.timings = {}
        self._times = self._get_output_times(freq=self.params["out_freq"],
                                             period_ending=self.params["period_ending"])
        self._update_unit_attrs(self.params["out_vars"])

    # -------------------------------------------------------------- inputs (lazy, metsim.py:286-317)
    @property
    def domain(self) -> Dataset:
        if self._domain is None:
            self._domain = read_domain(self.params)
        return self._domain

    @property
    def met_data(self) -> Dataset:
        if self._met_data is None:
            p = self.params
            fmt = p["forcing_fmt"]
            if fmt == "netcdf":
                ds = _read_netcdf(p["forcing"], p.get("forcing_vars"))
                ds = _select_domain(ds, self.domain)
                p["calendar"] = ds["time"].attrs.get("calendar", p["calendar"])      # metsim.py:252-254
                self._validate_force_times(ds["time"].values)
                ds = _select_time(ds, p["start"], p["stop"])
            elif fmt in ("ascii", "binary"):
                for key in ("start", "stop"):
                    if p[key] == "forcing":
                        raise ValueError(f"{fmt} forcings carry no time axis: `{key}` must be a date")
                    p[key] = _parse_date(p[key])
                times = _date_range(p["start"], p["stop"], p["calendar"])
                self._validate_force_times(times)
                ds = _read_vic(p, self.domain, times.astype("datetime64[s]"))
            else:
                raise KeyError(fmt)                                     # io.py:185-191
            constant_vars = p.get("constant_vars")
            if constant_vars:                                           # metsim.py:300-305
                template = next(v for k, v in ds.items() if k != "time" and "time" in v.dims)
                for var, val in constant_vars.items():
                    ds[var] = Var(np.full(template.values.shape, float(val)), template.dims)
            self._met_data = ds
        return self._met_data

    @property
    def state(self) -> Dataset:
        if self._state is None:
            self.met_data                                               # fixes state_start / state_stop
            p = self.params
            ds = _read_netcdf(p["state"], p.get("state_vars"))
            ds = _select_domain(ds, self.domain)
            ds = _select_time(ds, p["state_start"], p["state_stop"])
            if ds.dims["time"] != 90:                                   # metsim.py:593
                raise AssertionError(f"the state holds {ds.dims['time']} of the 90 days before `start` "
                                     f"({p['state_start']} .. {p['state_stop']})")
            self._state = ds
        return self._state

    def _validate_force_times(self, force_times: np.ndarray) -> None:
        """``start`` / ``stop`` as dates within the forcing record, the 90 state days before
        ``start`` and the attributes of the output time axis (metsim.py:239-276)."""
        p = self.params
        for key, i in (("start", 0), ("stop", -1)):
            if isinstance(p[key], str):
                p[key] = force_times[i].astype("datetime64[s]") if p[key] == "forcing" else _parse_date(p[key])
            else:
                p[key] = np.datetime64(p[key], "s")
        assert p["start"] >= force_times[0], "`start` lies before the forcing record"
        assert p["stop"] <= force_times[-1], "`stop` lies beyond the forcing record"
        p["state_start"] = p["start"] - np.timedelta64(90, "D")
        p["state_stop"] = p["start"] - np.timedelta64(1, "D")
        self.time_attrs = ({"long_name": "UTC time", "standard_name": "utc_time"} if p["utc_offset"]
                           else {"long_name": "local time at grid location", "standard_name": "local_time"})

    # -------------------------------------------------------------- bookkeeping
    def _update_unit_attrs(self, out_vars: dict) -> None:
        """metsim.py:226-237: unknown units fall back to the variable's default with a warning."""
        from .engine import unit_scale_offset
        av = available_outputs()
        ts = int(self.params["time_step"])
        for k, v in out_vars.items():
            v.setdefault("out_name", k)
            units = v.get("units")
            if units is not None:
                try:
                    unit_scale_offset(k, units, ts if ts > 0 else 60)
                    self.var_attrs[k]["units"] = units
                    continue
                except ValueError:
                    self.logger.warning(f"Could not find unit conversion for {k} to {units}! We will use "
                                        f"the default units of {av[k]['units']} instead.")
            v["units"] = av[k]["units"]

    def _get_output_times(self, freq=None, period_ending: bool = False):
        """List of output time vectors, one per output file (metsim.py:512-557)."""
        ts = int(self.params["time_step"])
        days = self.met_data["time"].values.astype("datetime64[s]")
        self.disagg = ts < 1440
        step = np.timedelta64(ts * 60, "s")
        tpd = 1440 // ts if self.disagg else 1
        times = (days[:, None] + np.arange(tpd)[None, :] * step).ravel()
        if period_ending:
            times = times + step
        if freq is None or freq == "":
            return [times]
        import pandas as pd
        dummy = pd.Series(np.arange(times.size), index=pd.DatetimeIndex(times))
        try:
            grouper = pd.Grouper(freq=freq)
        except ValueError:                          # pandas >= 2.2 renamed the aliases the examples use ('1Y')
            import re
            new = {"Y": "YE", "A": "YE", "M": "ME", "Q": "QE", "H": "h", "T": "min", "S": "s"}
            m = re.fullmatch(r"(\d*)([A-Za-z]+)", str(freq).strip())
            if not m or m.group(2) not in new:
                raise
            grouper = pd.Grouper(freq=m.group(1) + new[m.group(2)])
        return [times[g.values] for _, g in dummy.groupby(grouper) if len(g)]

    def get_nc_output_suffix(self, times) -> str:
        s, e = (str(t.astype("datetime64[D]")).replace("-", "") for t in (times[0], times[-1]))
        return f"{s}-{e}"

    def _get_output_filename(self, times) -> str:
        return os.path.join(os.path.abspath(self.params["out_dir"]),
                            f"{self.params['out_prefix']}_{self.get_nc_output_suffix(times)}.nc")

    def _validate_setup(self) -> None:
        """metsim.py:620-684, same messages."""
        p = self.params
        errs = [""]
        if not len(p.get("forcing", [])):
            errs.append("Requires input forcings to be specified")
        if not len(p.get("forcing_vars", [])):
            errs.append("Requires at least one non-constant forcing")
        for each in ("out_dir", "time_step", "forcing_fmt"):
            if p.get(each, None) is None or p[each] == "":
                errs.append(f"Cannot have empty value for {each}")
        ts = int(p.get("time_step", -1))
        if ts <= 0 or 1440 % ts or (ts > 360 and ts != 1440):
            errs.append("Time step must be evenly divisible into 1440 minutes (24 hours) and less than 360 "
                        f"minutes (6 hours). Got {p['time_step']}.")
        for each in ("t_min", "t_max", "prec"):
            if each not in self.met_data:
                errs.append(f"Input requires {each}")
        if not len(p.get("out_vars", [])):
            errs.append("Output variable list must not be empty")
        allowed = DAILY_OUT_VARS if ts == 1440 else SUBDAILY_OUT_VARS
        for var in p.get("out_vars", []):
            if var not in allowed:
                errs.append(f"Cannot output variable {var} at timestep {p['time_step']}")
        opts = {"out_precision": ["f4", "f8"], "lw_cloud": ["default", "cloud_deardorff"],
                "lw_type": ["default", "tva", "anderson", "brutsaert", "satterlund", "idso", "prata"]}
        for k, v in opts.items():
            if p.get(k, None) not in v:
                errs.append(f"Invalid option given for {k}")
        if len(errs) > 1:
            raise Exception("\n  ".join(errs))

    # -------------------------------------------------------------- the chunk the GPU gets
    def gather(self) -> dict:
        """The active cells of the domain as the ``[D][N]`` / ``[90][N]`` / ``[12][N]`` arrays of the
        C ABI (what ``run_slice`` hands cell by cell to ``wrap_run_cell``, metsim.py:474-488)."""
        p, dom, md, st = self.params, self.domain, self.met_data, self.state
        mask = dom["mask"]
        mdims = mask.dims
        active = np.argwhere(mask.values > 0)                    # row-major, like np.ndenumerate (:477)
        sel = tuple(active[:, i] for i in range(active.shape[1]))
        doy, month = _calendar(md["time"].values, p["calendar"])
        g = dict(lat=_per_cell(dom, "lat", mdims, mask.values.shape, sel),
                 lon=_per_cell(dom, "lon", mdims, mask.values.shape, sel),
                 elev=_per_cell(dom, "elev", mdims, mask.values.shape, sel), doy=doy, month=month)
        for k in ("t_min", "t_max", "prec"):
            g[k] = _cells(md[k], "time", mdims, sel)
            g["state_" + k] = _cells(st[k], "time", mdims, sel)
        g["wind"] = _cells(md["wind"], "time", mdims, sel) if "wind" in md else None
        if str(p.get("prec_type", "uniform")).upper() in ("TRIANGLE", "MIX"):
            g["t_pk12"] = _cells(dom["t_pk"], "month", mdims, sel)
            g["dur12"] = _cells(dom["dur"], "month", mdims, sel)
        if "passthrough" in str(p.get("method", "mtclim")).lower():
            for k in ("shortwave", "vapor_pressure", "tskc", "longwave"):
                if k in md:
                    g["given_" + k] = _cells(md[k], "time", mdims, sel)
        if (g["t_max"] < g["t_min"]).any():                      # metsim.py:612-614
            raise ValueError("Daily maximum temperature lower than daily minimum temperature!")
        g["cell_index"] = np.ravel_multi_index(sel, mask.values.shape).astype(np.int64)
        return g

    def _writers(self):
        """One NetCDF file per output time vector, laid out like ``setup_netcdf_output``
        (metsim.py:363-436)."""
        from .ncwriter import NC3Writer
        p, dom = self.params, self.domain
        mask = dom["mask"]
        dtype = {"f4": "float32", "f8": "float64"}[p["out_precision"]]
        skip = {"elev", "lat", "lon", "is_worker", "out_vars", "forcing_vars", "domain_vars", "state_vars",
                "constant_vars", "references", "verbose", "num_workers", "state_start", "state_stop",
                "out_freq"}
        gatts = {"conventions": "1.6", "title": "Output from MetSim", "institution": "University of Washington",
                 "history": f"Created: {_time.ctime()} by {getpass.getuser()}",
                 "comment": "no comment at this time"}
        for k, v in p.items():
            if k in skip:
                continue
            if isinstance(v, dict):
                v = json.dumps(v)
            elif isinstance(v, (list, tuple)):
                v = ", ".join(str(x) for x in v)
            elif not isinstance(v, (int, float)) or isinstance(v, bool):
                v = str(v)
            gatts[k] = v.replace("'", "").replace('"', "") if isinstance(v, str) else v
        variables = OrderedDict()
        for varname, conf in p["out_vars"].items():
            variables[conf["out_name"]] = dict(self.var_attrs.get(varname, {}), dtype=dtype,
                                               missing_value=np.nan, fill_value=np.nan)
        os.makedirs(os.path.abspath(p["out_dir"]), exist_ok=True)
        grid = [(d, _coordinate(dom, d)) for d in mask.dims]
        cell_index = np.flatnonzero(mask.values.ravel() > 0)
        out = []
        for times in self._times:
            if str(p["calendar"]).lower() in ncio.STANDARD_CALENDARS:
                minutes = ((times - _EPOCH) // np.timedelta64(60, "s")).astype(np.int64)
            else:                                   # date2num in the run's own calendar (metsim.py:385-388)
                minutes = np.round(ncio.encode_time(times, DEFAULT_TIME_UNITS, p["calendar"])).astype(np.int64)
            out.append(NC3Writer(self._get_output_filename(times), n_steps=times.size, grid_dims=grid,
                                 cell_index=cell_index, variables=variables,
                                 time_values=minutes.astype(np.int32), time_units=DEFAULT_TIME_UNITS,
                                 time_attrs=dict(self.time_attrs, calendar=str(p["calendar"])),
                                 global_attrs=gatts))
        return out

    # -------------------------------------------------------------- run
    def run(self, host_budget_bytes: int = 1 << 30, devices=None) -> list:
        """``MetSim.run`` (metsim.py:330-352): validate, create the output files, compute, write.
        The whole domain is one ``msb_run_host`` call; its time tiles stream into the files through
        the ``on_tile`` hook while the next tile computes.  Returns the output file names.

        ``devices``: CUDA device indices to spread the active cells over (default: this object's
        one device).  The cells are cut into contiguous ranges (``shard.cell_range``), one host
        thread + ``msb_host_ctx`` per range, no exchange between them; their tiles are gathered on
        the host into full rows before they go to the file (the north_star's "optional host gather
        of outputs").  The reference's counterpart is one dask task per ``[chunks]`` slice
        (metsim.py:339-343)."""
        from .engine import run_host, unit_scale_offset
        self._validate_setup()
        p = self.params
        ts = int(p["time_step"])
        out_vars = p["out_vars"]
        t0 = _time.perf_counter()
        g = self.gather()
        cell_index = g.pop("cell_index")
        n, d = cell_index.size, g["t_min"].shape[0]
        if "wind" in out_vars and g["wind"] is None:
            raise KeyError("wind")                                # df['wind'] in the reference (:492)
        t1 = _time.perf_counter()
        writers = self._writers()
        t2 = _time.perf_counter()
        names = [w.path for w in writers]
        bounds = np.cumsum([0] + [w.n_steps for w in writers])   # global step range of each file

        def write(varname, step0, block):
            """rows [step0, step0 + len(block)) of the record -> the file(s) they fall into"""
            oname = out_vars[varname]["out_name"]
            lo, hi = step0, step0 + block.shape[0]
            for w, a, b in zip(writers, bounds[:-1], bounds[1:]):
                s0, s1 = max(lo, a), min(hi, b)
                if s0 < s1:
                    w.write_steps(oname, int(s0 - a), block[s0 - lo:s1 - lo])

        devices = [int(x) for x in devices] if devices else [self.device]
        n_shards = max(1, min(len(devices), n // 32))             # at least a warp of cells per range
        daily_cols = ("t_day", "vapor_pressure", "tskc", "pet", "daylength", "shortwave_daily")
        try:
            if n == 0:
                pass
            elif not self.disagg:                                 # daily outputs, metsim.py:779-792
                if n_shards > 1:
                    parts = self._run_shards(g, devices[:n_shards], dict(daily_vars=daily_cols))
                    res = {"daily": {c: np.concatenate([r["daily"][c] for r in parts], axis=1) for c in daily_cols}}
                else:
                    res = run_host(p, **g, device=self.device, timings=self.timings, daily_vars=daily_cols)
                odt = np.dtype({"f4": "float32", "f8": "float64"}[p["out_precision"]])
                for v in out_vars:
                    src = (g[v] if v in ("t_min", "t_max", "prec", "wind")
                           else res["daily"]["shortwave_daily" if v == "shortwave" else v])
                    sc, of = unit_scale_offset(v, out_vars[v]["units"], ts)
                    write(v, 0, np.ascontiguousarray(src * sc + of if (sc, of) != (1.0, 0.0) else src, dtype=odt))
            else:
                tpd = 1440 // ts
                odt = np.dtype({"f4": "float32", "f8": "float64"}[p["out_precision"]])
                row_bytes = n * odt.itemsize * len(out_vars)
                tile_days = int(max(1, min(d, host_budget_bytes // max(1, 2 * tpd * row_bytes))))
                units = {v: out_vars[v]["units"] for v in out_vars}
                if n_shards > 1:
                    kw = dict(out_vars=tuple(out_vars), out_dtype=odt, units=units)
                    if tile_days >= d:                            # the record fits: whole columns per range
                        parts = self._run_shards(g, devices[:n_shards], kw)
                        for v in out_vars:
                            write(v, 0, np.concatenate([r[v] for r in parts], axis=1))
                    else:
                        self._run_shards(g, devices[:n_shards], kw, tile_days=tile_days, tpd=tpd,
                                         consume=lambda v, step0, block: write(v, step0, block))
                elif tile_days >= d:                              # the record fits: one piece, no ring
                    res = run_host(p, **g, out_vars=tuple(out_vars), out_dtype=odt, units=units,
                                   device=self.device, timings=self.timings)
                    for v in out_vars:
                        write(v, 0, res[v])
                else:
                    key = (tuple(out_vars), 2 * tile_days * tpd, n, odt)
                    if self._ring is None or self._ring[0] != key:     # page-locking a ring costs ~0.3 s per GB: keep it
                        self._ring = None
                        self._ring = (key, {v: _pinned((2 * tile_days * tpd, n), odt) for v in out_vars})
                    ring = self._ring[1]

                    def on_tile(tile, day0, day1, slot):
                        rows = slice(slot * tile_days * tpd, slot * tile_days * tpd + (day1 - day0) * tpd)
                        for v in out_vars:
                            write(v, day0 * tpd, ring[v][rows])

                    run_host(p, **g, out_vars=tuple(out_vars), out_dtype=odt, units=units, out=ring,
                             ring=True, tile_days=tile_days, on_tile=on_tile, device=self.device,
                             timings=self.timings)
        finally:
            for w in writers:
                w.close()
        self.timings.update(s_gather=t1 - t0, s_create_files=t2 - t1, s_compute_and_write=_time.perf_counter() - t2)
        return names

    def _run_shards(self, g: dict, devices: list, kw: dict, tile_days: int = 0, tpd: int = 0, consume=None) -> list:
        """One host thread and one ``msb_host_ctx`` per contiguous cell range (``shard.cell_range``;
        the devices may repeat).  Without ``consume``: returns the ranges' ``run_host`` results in
        cell order.  With it (streaming): every range runs a 2-slot ring of ``tile_days``; a tile's
        rows are gathered into one of two full-width host buffers, and when all ranges have
        delivered tile *t* the first thread hands the full rows to ``consume(var, step0, block)``
        while the others already compute tile *t + 1*."""
        import threading
        from . import engine
        from .shard import cell_range, shard_inputs
        p, K = self.params, len(devices)
        n = g["t_min"].shape[1]
        ranges = [cell_range(n, k, K) for k in range(K)]
        results, errors, timings = [None] * K, [], [dict() for _ in range(K)]
        barrier = threading.Barrier(K)
        full = rings = None
        if consume is not None:
            odt = np.dtype(kw["out_dtype"])
            key = ("ranges", tuple(kw["out_vars"]), tile_days * tpd, tuple(ranges), odt)
            if self._ring is None or self._ring[0] != key:        # page-locked rings + gather buffers: kept between runs
                self._ring = None
                self._ring = (key, ([{v: np.empty((tile_days * tpd, n), dtype=odt) for v in kw["out_vars"]} for _ in range(2)],
                                    [{v: _pinned((2 * tile_days * tpd, hi - lo), odt) for v in kw["out_vars"]}
                                     for lo, hi in ranges]))
            full, rings = self._ring[1]

        def work(k):
            lo, hi = ranges[k]
            try:
                gk = shard_inputs(g, k, K)
                with engine.HostContext(devices[k]) as ctx:
                    if consume is None:
                        results[k] = engine.run_host(p, **gk, **kw, ctx=ctx, timings=timings[k])
                        return
                    ring = rings[k]

                    def on_tile(tile, day0, day1, slot):
                        rows, src0 = (day1 - day0) * tpd, slot * tile_days * tpd
                        dst = full[tile % 2]
                        for v in kw["out_vars"]:
                            dst[v][:rows, lo:hi] = ring[v][src0:src0 + rows]
                        barrier.wait()                  # every range has delivered this tile
                        if k == 0:                      # ... and the others are on their way to the next one
                            for v in kw["out_vars"]:
                                consume(v, day0 * tpd, dst[v][:rows])

                    engine.run_host(p, **gk, **kw, out=ring, ring=True, tile_days=tile_days, on_tile=on_tile,
                                    ctx=ctx, timings=timings[k])
            except BaseException as exc:                # noqa: BLE001 -- re-raised by the caller's thread
                errors.append(exc)
                barrier.abort()

        threads = [threading.Thread(target=work, args=(k,), name=f"msb-range-{k}") for k in range(K)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        real = [e for e in errors if not isinstance(e, threading.BrokenBarrierError)]
        if errors:
            raise (real or errors)[0]
        self.timings["ranges"] = [dict(cells=hi - lo, device=dev, **tm) for (lo, hi), dev, tm in zip(ranges, devices, timings)]
        return results

    def open_output(self) -> list:
        """The output files read back (``MetSim.open_output``, metsim.py:326-328), one
        :class:`metsim_b200.ncio.Dataset` per file."""
        return [ncio.open_dataset(self._get_output_filename(t)) for t in self._times]


# ------------------------------------------------------------------------------ command line
def parse(args):
    """The options of ``ms`` (metsim/cli/ms.py:40-54)."""
    parser = argparse.ArgumentParser(prog="python -m metsim_b200.driver")
    parser.add_argument("config", help="Input configuration file")
    parser.add_argument("-n", "--num_workers", default=1, type=int,
                        help="number of GPUs to spread the domain over (contiguous cell ranges); the "
                             "reference's number of worker processes")
    parser.add_argument("-s", "--scheduler", default="distributed", type=str,
                        help="accepted for compatibility (no dask here)")
    parser.add_argument("-v", "--verbose", action="store_true", help="Increase the verbosity")
    parser.add_argument("-d", "--device", default=0, type=int, help="CUDA device index (single-GPU run)")
    from . import __version__
    parser.add_argument("--version", action="version", version=f"metsim_b200 {__version__}",
                        help="Name and version number")
    parser.add_argument("--devices", default=None, type=str,
                        help="comma-separated CUDA device indices, one cell range each (overrides -n / -d)")
    opts = parser.parse_args(args)
    if not os.path.isfile(opts.config):
        parser.print_help()
        parser.exit(status=1, message=f"Configuration file {opts.config} does not exist.\n")
    return opts


def main(argv=None) -> list:
    opts = parse(sys.argv[1:] if argv is None else argv)
    setup = read_config(opts.config, scheduler=opts.scheduler, num_workers=opts.num_workers,
                        verbose=opts.verbose)
    ms = MetSimB200(setup, device=opts.device)
    devices = None
    if opts.devices:
        devices = [int(x) for x in opts.devices.split(",")]
    elif opts.num_workers > 1:
        import torch
        devices = list(range(min(opts.num_workers, max(1, torch.cuda.device_count()))))
    files = ms.run(devices=devices)
    if opts.verbose:
        tm = ms.timings
        ms.logger.info("ran %d cell range(s); gather %.2f s, create files %.2f s, compute + write %.2f s",
                       len(tm.get("ranges", [0])), tm.get("s_gather", 0.0), tm.get("s_create_files", 0.0),
                       tm.get("s_compute_and_write", 0.0))
    for f in files:
        print(f)
    return files


if __name__ == "__main__":
    main()
