"""Deterministic synthetic domains and forcings.

Cells are enumerated latitude-row-major, so a contiguous chunk of cells is a latitude band and a
shard only needs solar tables for its own latitudes.  ``numpy`` generation is used by the tests
(identical inputs for the oracle and the CUDA path); ``torch`` generation on the device is used by
``bench.py`` for the large grids.
"""
from __future__ import annotations

import numpy as np

from .engine import calendar_vectors

SEED = 20261019

CONFIGS = {
    # name: (grid, n_days, start, time_step, prec_type)
    "test_domain": ("test9x9", 31, "1950-01-01", 30, "triangle"),
    "global_half_deg": ("global0.5", 365, "2001-01-01", 60, "triangle"),
    "conus_16th_hourly": ("conus1/16", 365, "2001-01-01", 60, "triangle"),
    "conus_16th_10yr_3h_mix": ("conus1/16", 3652, "2001-01-01", 180, "mix"),
    "global_8th_30yr_15min": ("global1/8", 10957, "1981-01-01", 15, "triangle"),
}


def grid_cells(grid: str, rng: np.random.Generator):
    """(lat[N], lon[N]) of the active cells, latitude-row-major."""
    if grid == "conus1/16":
        lats = 25.03125 + np.arange(520) / 16.0
        lons = -124.96875 + np.arange(928) / 16.0
        mask = None
    elif grid == "global0.5":
        lats = np.arange(-55.75, 83.75 + 1e-9, 0.5)
        lons = np.arange(-179.75, 179.75 + 1e-9, 0.5)
        mask = rng.random((lats.size, lons.size)) < (1.0 / 3.0)
    elif grid == "global1/8":
        lats = np.arange(-55.9375, 83.9375 + 1e-9, 0.125)
        lons = -179.9375 + np.arange(2880) * 0.125
        mask = rng.random((lats.size, lons.size)) < 0.3125
    elif grid == "test9x9":
        lats = 30.03125 + np.arange(9) / 16.0
        lons = -100.03125 + np.arange(9) / 16.0
        mask = None
    else:
        raise ValueError(grid)
    la, lo = np.meshgrid(lats, lons, indexing="ij")
    if mask is None:
        return la.ravel().copy(), lo.ravel().copy()
    return la[mask], lo[mask]


def make_forcing(lat, lon, n_days: int, start: str, rng: np.random.Generator, wind: bool = True):
    """numpy forcing dict for the given cells (float64, [D][N])."""
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    n = lat.size
    doy, month = calendar_vectors(start, n_days)
    # 90 state days precede the record; run the same process over state + record
    sdoy, _ = calendar_vectors(np.datetime64(start, "D") - 90, 90)
    all_doy = np.concatenate([sdoy, doy]).astype(np.float64)
    sign = np.where(lat >= 0, 1.0, -1.0)
    base = (12.0 - 0.45 * np.abs(lat))[None, :] - 8.0 * np.cos(
        2 * np.pi * (all_doy - 15.0) / 365.25)[:, None] * sign[None, :]
    total = n_days + 90
    noise = np.empty((total, n))
    eps = rng.normal(0.0, 3.0, size=(total, n))
    noise[0] = eps[0]
    innov = np.sqrt(1 - 0.7 ** 2)
    for d in range(1, total):
        noise[d] = 0.7 * noise[d - 1] + innov * eps[d]
    t_min_all = base + noise
    t_max_all = t_min_all + rng.uniform(2.0, 18.0, size=(total, n))
    prec_all = (rng.random((total, n)) < 0.3) * rng.gamma(0.8, 6.0, size=(total, n))
    out = dict(lat=lat, lon=lon, elev=np.clip(rng.gamma(2.0, 400.0, size=n), 0.0, 4500.0),
               doy=doy, month=month,
               t_min=np.ascontiguousarray(t_min_all[90:]), t_max=np.ascontiguousarray(t_max_all[90:]),
               prec=np.ascontiguousarray(prec_all[90:]),
               state_t_min=np.ascontiguousarray(t_min_all[:90]),
               state_t_max=np.ascontiguousarray(t_max_all[:90]),
               state_prec=np.ascontiguousarray(prec_all[:90]),
               t_pk12=rng.uniform(300.0, 1140.0, size=(12, n)),
               dur12=rng.uniform(120.0, 720.0, size=(12, n)))
    if wind:
        out["wind"] = rng.uniform(0.5, 8.0, size=(n_days, n))
    return out


def make_config(name: str, max_cells: int | None = None, n_days: int | None = None, seed: int = SEED):
    """(forcing dict, params dict) for one of the BASELINE.json configurations, optionally
    truncated to the first ``max_cells`` cells / ``n_days`` days (tests use small slices)."""
    grid, d, start, ts, prec_type = CONFIGS[name]
    rng = np.random.default_rng(seed)
    lat, lon = grid_cells(grid, rng)
    if max_cells is not None and lat.size > max_cells:
        # keep whole-domain coverage: evenly spaced cells, still latitude-row-major
        idx = np.linspace(0, lat.size - 1, max_cells).astype(np.int64)
        lat, lon = lat[idx], lon[idx]
    f = make_forcing(lat, lon, n_days or d, start, rng)
    params = dict(time_step=ts, prec_type=prec_type, utc_offset=True)
    return f, params


def make_forcing_torch(lat, lon, n_days: int, start: str, device, seed: int = SEED, wind: bool = True):
    """Same process generated with torch on ``device`` (for the large benchmark grids)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = torch.float64
    lat_t = torch.as_tensor(np.asarray(lat, dtype=np.float64), device=device)
    lon_t = torch.as_tensor(np.asarray(lon, dtype=np.float64), device=device)
    n = lat_t.numel()
    doy, month = calendar_vectors(start, n_days)
    sdoy, _ = calendar_vectors(np.datetime64(start, "D") - 90, 90)
    all_doy = torch.as_tensor(np.concatenate([sdoy, doy]).astype(np.float64), device=device)
    sign = torch.where(lat_t >= 0, 1.0, -1.0).to(f64)
    total = n_days + 90
    rnd = lambda *shape: torch.rand(*shape, generator=g, device=device, dtype=f64)
    t_min_all = torch.empty((total, n), dtype=f64, device=device)
    seas = -8.0 * torch.cos(2 * np.pi * (all_doy - 15.0) / 365.25)
    base_lat = 12.0 - 0.45 * lat_t.abs()
    noise = torch.zeros(n, dtype=f64, device=device)
    innov = float(np.sqrt(1 - 0.7 ** 2))
    for d in range(total):
        eps = torch.randn(n, generator=g, device=device, dtype=f64) * 3.0
        noise = eps if d == 0 else 0.7 * noise + innov * eps
        t_min_all[d] = base_lat + seas[d] * sign + noise
    # gamma(0.8, 6) via torch.distributions is CPU-seeded; use -log(U)*scale*U^(1/k) (Ahrens-Dieter
    # boost of an exponential) which has the right shape and support for a synthetic field
    if total * n <= (1 << 28):
        t_max_all = t_min_all + (2.0 + 16.0 * rnd(total, n))
        wet = rnd(total, n) < 0.3
        prec_all = torch.where(wet, -torch.log(rnd(total, n).clamp_min(1e-300)) * 6.0
                               * rnd(total, n).pow(1.0 / 0.8), torch.zeros((), dtype=f64, device=device))
    else:   # multi-decade records: the same process filled in blocks of days (bounded temporaries)
        t_max_all = torch.empty_like(t_min_all)
        prec_all = torch.empty_like(t_min_all)
        for d0 in range(0, total, 256):
            d1 = min(total, d0 + 256)
            t_max_all[d0:d1] = t_min_all[d0:d1] + (2.0 + 16.0 * rnd(d1 - d0, n))
            wet = rnd(d1 - d0, n) < 0.3
            prec_all[d0:d1] = torch.where(wet, -torch.log(rnd(d1 - d0, n).clamp_min(1e-300)) * 6.0
                                          * rnd(d1 - d0, n).pow(1.0 / 0.8),
                                          torch.zeros((), dtype=f64, device=device))
    elev = (-(torch.log(rnd(n).clamp_min(1e-300)) + torch.log(rnd(n).clamp_min(1e-300))) * 400.0
            ).clamp(0.0, 4500.0)  # gamma(2, 400) = sum of two exponentials
    out = dict(lat=lat_t, lon=lon_t, elev=elev,
               doy=torch.as_tensor(doy, device=device), month=torch.as_tensor(month, device=device),
               t_min=t_min_all[90:].contiguous(), t_max=t_max_all[90:].contiguous(),
               prec=prec_all[90:].contiguous(),
               state_t_min=t_min_all[:90].contiguous(), state_t_max=t_max_all[:90].contiguous(),
               state_prec=prec_all[:90].contiguous(),
               t_pk12=300.0 + 840.0 * rnd(12, n), dur12=120.0 + 600.0 * rnd(12, n))
    if wind:
        if n_days * n <= (1 << 28):
            out["wind"] = 0.5 + 7.5 * rnd(n_days, n)
        else:
            out["wind"] = torch.empty((n_days, n), dtype=f64, device=device)
            for d0 in range(0, n_days, 256):
                d1 = min(n_days, d0 + 256)
                out["wind"][d0:d1] = 0.5 + 7.5 * rnd(d1 - d0, n)
    return out
