"""``'passthrough_b200'`` -- drop-in for ``metsim.methods.passthrough``
(metsim/methods/passthrough.py:28-46).

Same module shape as the reference registry expects (``run(df, params)``, metsim/metsim.py:139,
:462, :744).  Daily shortwave is a forcing column (``assert 'shortwave' in df``); ``t_day``,
``tfmax``, ``tskc``, ``pet``, ``tdew`` and ``vapor_pressure`` are derived from it in ONE evaluation
(no fixed point) by the same CUDA daily kernel as ``mtclim_b200`` unless the column is already on
``df``.  Supplied ``tskc`` / ``vapor_pressure`` columns are honoured; supplying one of the
intermediate columns ``t_day`` / ``tfmax`` / ``pet`` / ``tdew`` is not supported by the kernel and
raises.  The sub-daily consequences of the method (full-day shortwave averaging,
disaggregate.py:78-80; longwave rescaled to a supplied daily longwave, :583-589) live in the
disaggregation kernels (``params['method']`` / ``'sw_averaging'`` through ``msb_params``).
"""
from __future__ import annotations

from .mtclim_b200 import ADDED_COLUMNS, run_daily  # noqa: F401

_UNSUPPORTED = ("t_day", "tfmax", "pet", "tdew")


def run(df, params: dict):
    assert 'shortwave' in df                                   # passthrough.py:29
    bad = [k for k in _UNSUPPORTED if k in df]
    if bad:
        raise ValueError(f"passthrough_b200: supplied intermediate column(s) {bad} are not supported")
    given = {"given_shortwave": "shortwave"}
    if "tskc" in df:
        given["given_tskc"] = "tskc"
    if "vapor_pressure" in df:
        given["given_vapor_pressure"] = "vapor_pressure"
    return run_daily(df, params, "passthrough", given)
