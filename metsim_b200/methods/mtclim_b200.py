"""``'mtclim_b200'`` -- drop-in for ``metsim.methods.mtclim`` (metsim/methods/mtclim.py).

The reference registry maps a method name to a *module* exposing ``run(df, params)``
(metsim/metsim.py:139, :462, :744).  This module has that shape.  ``run`` pushes the cell's
daily record through the same CUDA daily kernels the batched engine uses (a 1-cell shard), using
the columns the reference driver has already put on ``df`` (``seasonal_prec``, ``smoothed_dtr``,
``daylength``, ``potrad``, ``tt_max``; metsim.py:729-739) exactly as ``mtclim.run`` does.

Like the reference it adds ``t_day, tfmax, tskc, tdew, vapor_pressure, shortwave, pet`` to ``df``
in place and returns the same object.  There is no CPU fallback: without a GPU / the built
library this raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib

ADDED_COLUMNS = ("t_day", "tfmax", "tskc", "tdew", "vapor_pressure", "shortwave", "pet")
_READ_PARAMS = ("tday_coef", "sw_prec_thresh", "rain_scalar", "lw_cloud", "tdew_tol", "lapse_rate")


def run(df, params: dict):
    """Same contract as metsim/methods/mtclim.py:28-80."""
    return run_daily(df, params, "mtclim")


def run_daily(df, params: dict, method: str, given: dict | None = None):
    """One cell's daily record through ``msb_mtclim_daily`` (shared with ``passthrough_b200``).
    ``given`` maps C-ABI forcing names (``given_shortwave`` ...) to df column names."""
    lib = _lib.lib()
    if not torch.cuda.is_available():
        raise RuntimeError(f"{method}_b200.run needs a CUDA device; there is no CPU fallback")
    dev = torch.device(params.get("device", "cuda:0"))
    p = _lib.make_params(dict({k: params[k] for k in _READ_PARAMS if k in params}, method=method))
    n_days = len(df)
    col = lambda name: torch.from_numpy(
        np.array(df[name].values, dtype=np.float64, copy=True).reshape(n_days, 1)).to(dev)
    cols = {k: col(k) for k in ("t_min", "t_max", "prec", "seasonal_prec", "smoothed_dtr",
                                "daylength", "potrad", "tt_max")}
    if "dtr" in df and (np.asarray(df["dtr"].values) < 0).any():
        raise ValueError("Daily maximum temperature lower than daily minimum temperature!")
    elev = torch.tensor([float(params["elev"])], dtype=torch.float64, device=dev)
    f = _lib.Forcing()
    f.n_cells, f.n_days, f.ld = 1, n_days, 1
    f.elev = elev.data_ptr()
    f.t_min, f.t_max, f.prec = (cols[k].data_ptr() for k in ("t_min", "t_max", "prec"))
    f.given_seasonal_prec = cols["seasonal_prec"].data_ptr()
    f.given_smoothed_dtr = cols["smoothed_dtr"].data_ptr()
    f.given_daylength = cols["daylength"].data_ptr()
    f.given_potrad = cols["potrad"].data_ptr()
    f.given_tt_max = cols["tt_max"].data_ptr()
    for cname, dfname in (given or {}).items():
        cols[cname] = col(dfname)
        setattr(f, cname, cols[cname].data_ptr())
    out = {k: torch.empty((n_days, 1), dtype=torch.float64, device=dev) for k in ADDED_COLUMNS}
    dl = _lib.Daily()
    for i, k in enumerate(_lib.DAILY_VARS):
        dl.var[i] = out[k].data_ptr() if k in out else None
    n_iter = torch.zeros(1, dtype=torch.int32, device=dev)
    scratch = torch.empty(lib.msb_daily_scratch_bytes(1, n_days) // 8 + 1, dtype=torch.float64,
                          device=dev)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    dl.n_iter, dl.partial, dl.status = n_iter.data_ptr(), scratch.data_ptr(), status.data_ptr()
    stream = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.msb_mtclim_daily(C.byref(p), C.byref(f), None, C.byref(dl),
                                        C.c_void_p(stream.cuda_stream)))
        _lib.check(lib.msb_daily_check(C.byref(f), C.byref(dl), C.c_void_p(stream.cuda_stream)))
    for k in ADDED_COLUMNS:
        if method == "mtclim" or k not in df:       # passthrough.py:31-45 keeps supplied columns
            df[k] = out[k].cpu().numpy()[:, 0]
    return df
