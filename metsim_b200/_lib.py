"""ctypes binding of the C ABI in ``include/metsim_b200.h`` (``csrc/libmetsim_b200.so``).

The library is the only compute path.  If it is missing or a symbol cannot be resolved this module
raises -- there is no Python / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MSB_LIBRARY: a differently-compiled build of the same sources (profiles/tools/build_variants.py);
# still the CUDA library, never a fallback
LIB_PATH = os.environ.get("MSB_LIBRARY") or os.path.join(_HERE, "csrc", "libmetsim_b200.so")

N_SUBVARS = 10
N_DAILYVARS = 13
SUB_VARS = ("temp", "prec", "shortwave", "longwave", "vapor_pressure", "rel_humid",
            "air_pressure", "spec_humid", "tskc", "wind")
DAILY_VARS = ("t_day", "tfmax", "tskc", "tdew", "vapor_pressure", "shortwave", "pet",
              "daylength", "potrad", "tt_max", "seasonal_prec", "smoothed_dtr", "shortwave_daily")
PREC_TYPES = {"uniform": 0, "triangle": 1, "mix": 2}
LW_TYPES = {"default": 0, "tva": 1, "anderson": 2, "brutsaert": 3, "satterlund": 4, "idso": 5,
            "prata": 6}
LW_CLOUD = {"default": 0, "cloud_deardorff": 1}
# MetSim.methods keys (metsim/metsim.py:139) and the names this package registers for them
METHODS = {"mtclim": 0, "mtclim_b200": 0, "passthrough": 1, "passthrough_b200": 1}
GIVEN_KEYS = ("given_shortwave", "given_vapor_pressure", "given_tskc", "given_longwave")
Q_LD = 2888
Q_PAD = 8192
TT_TERMS = 28

E_ARG, E_CUDA, E_DTR_NEGATIVE, E_NOT_CONVERGED, E_KNOTS, E_NOMEM = -1, -2, -3, -4, -5, -6

DAILY_CAND = 3          # MSB_DAILY_CAND: candidate planes of the single-pass daily form
DAILY_CAND_K0 = 4       # MSB_DAILY_CAND_K0

EXPORTS = ("msb_version", "msb_last_error", "msb_default_params", "msb_solar_prelude_bytes",
           "msb_solar_table_sizes", "msb_solar_q_win_ld", "msb_solar_q_win_elems",
           "msb_solar_tables_build", "msb_daily_scratch_bytes",
           "msb_daily_scratch_bytes_chunk", "msb_daily_single_pass",
           "msb_mtclim_daily", "msb_daily_check", "msb_disagg_workspace_bytes",
           "msb_disagg_workspace_bytes_lw", "msb_disaggregate",
           "msb_host_ctx_create", "msb_host_ctx_destroy", "msb_host_ctx_device_bytes",
           "msb_run_host_bytes", "msb_run_host_ctx", "msb_run_host",
           "msb_sink_create", "msb_sink_write", "msb_sink_checksum", "msb_sink_bytes_written",
           "msb_sink_destroy",
           "msb_selftest_math")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


class Params(C.Structure):
    _fields_ = [("time_step", C.c_int32), ("prec_type", C.c_int32), ("utc_offset", C.c_int32),
                ("lw_type", C.c_int32), ("lw_cloud", C.c_int32), ("max_iter", C.c_int32),
                ("method", C.c_int32), ("sw_daylight", C.c_int32),
                ("sw_prec_thresh", C.c_double), ("rain_scalar", C.c_double),
                ("tdew_tol", C.c_double), ("tmax_daylength_fraction", C.c_double),
                ("tday_coef", C.c_double), ("lapse_rate", C.c_double)]


class Tuning(C.Structure):
    """msb_tuning: launch-geometry / path overrides, 0 = the library's choice."""
    _fields_ = [("daily_chunk_len", C.c_int32), ("daily_two_pass", C.c_int32),
                ("knot_tile", C.c_int32), ("day_tile", C.c_int32), ("disagg_path", C.c_int32),
                ("day_prefetch", C.c_int32), ("prep_days", C.c_int32), ("sw_table", C.c_int32)]


TUNING_KEYS = tuple(n for n, _ in Tuning._fields_)
# environment spellings of the same knobs (read here, in Python, once per Engine / run_host call --
# the C library itself never looks at the environment)
_TUNING_ENV = {"MSB_DAILY_CHUNK": "daily_chunk_len", "MSB_DAILY_TWO_PASS": "daily_two_pass",
               "MSB_KNOT_TILE": "knot_tile", "MSB_DAY_TILE": "day_tile", "MSB_DAY_PF": "day_prefetch",
               "MSB_PREP_DAYS": "prep_days", "MSB_SW_TABLE": "sw_table"}


def make_tuning(user: dict | None = None) -> Tuning:
    """``user`` keys are the msb_tuning field names (``disagg_path`` also takes 'event' / 'auto').
    Unset fields fall back to the MSB_* environment variables, then to 0 (library default)."""
    t = Tuning()
    for env, field in _TUNING_ENV.items():
        v = os.environ.get(env)
        if v:
            setattr(t, field, int(v))
    if os.environ.get("MSB_DISAGG_PATH", "").startswith("e"):
        t.disagg_path = 1
    for k, v in (user or {}).items():
        if k not in TUNING_KEYS:
            raise KeyError(f"unknown tuning key {k!r}")
        if k == "disagg_path" and isinstance(v, str):
            v = 1 if v.startswith("e") else 0
        setattr(t, k, int(v))
    return t


class SolarTables(C.Structure):
    _fields_ = [("n_lat", C.c_int32), ("daylength", C.c_void_p), ("potrad", C.c_void_p),
                ("rise_min", C.c_void_p), ("set_min", C.c_void_p), ("q", C.c_void_p),
                ("tt_coef", C.c_void_p), ("q_win", C.c_void_p), ("q_win_step", C.c_int32),
                ("reserved_", C.c_int32)]


class Forcing(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("n_days", C.c_int32), ("ld", C.c_int64),
                ("lon", C.c_void_p), ("elev", C.c_void_p), ("lat_row", C.c_void_p),
                ("doy", C.c_void_p), ("month", C.c_void_p),
                ("t_min", C.c_void_p), ("t_max", C.c_void_p), ("prec", C.c_void_p),
                ("wind", C.c_void_p),
                ("st_t_min", C.c_void_p), ("st_t_max", C.c_void_p), ("st_prec", C.c_void_p),
                ("t_pk12", C.c_void_p), ("dur12", C.c_void_p), ("t_end", C.c_void_p),
                ("given_seasonal_prec", C.c_void_p), ("given_smoothed_dtr", C.c_void_p),
                ("given_daylength", C.c_void_p), ("given_potrad", C.c_void_p),
                ("given_tt_max", C.c_void_p),
                ("given_shortwave", C.c_void_p), ("given_vapor_pressure", C.c_void_p),
                ("given_tskc", C.c_void_p), ("given_longwave", C.c_void_p)]


class Daily(C.Structure):
    _fields_ = [("var", C.c_void_p * N_DAILYVARS), ("n_iter", C.c_void_p),
                ("partial", C.c_void_p), ("status", C.c_void_p),
                ("cand", C.c_void_p), ("cand_sel", C.c_void_p), ("cand_list", C.c_void_p),
                ("tune", Tuning)]


class SubDaily(C.Structure):
    _fields_ = [("var", C.c_void_p * N_SUBVARS), ("scale", C.c_double * N_SUBVARS),
                ("offset", C.c_double * N_SUBVARS), ("dtype", C.c_int32), ("ld", C.c_int64),
                ("day0", C.c_int32), ("day1", C.c_int32),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("tune", Tuning)]


class HostRun(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("n_days", C.c_int32), ("device", C.c_int32),
                ("lat", C.c_void_p), ("lon", C.c_void_p), ("elev", C.c_void_p),
                ("doy", C.c_void_p), ("month", C.c_void_p),
                ("t_min", C.c_void_p), ("t_max", C.c_void_p), ("prec", C.c_void_p),
                ("wind", C.c_void_p),
                ("st_t_min", C.c_void_p), ("st_t_max", C.c_void_p), ("st_prec", C.c_void_p),
                ("t_pk12", C.c_void_p), ("dur12", C.c_void_p), ("t_end", C.c_void_p),
                ("given_shortwave", C.c_void_p), ("given_vapor_pressure", C.c_void_p),
                ("given_tskc", C.c_void_p), ("given_longwave", C.c_void_p),
                ("out_host", C.c_void_p * N_SUBVARS), ("scale", C.c_double * N_SUBVARS),
                ("offset", C.c_double * N_SUBVARS), ("out_dtype", C.c_int32),
                ("daily_host", C.c_void_p * N_DAILYVARS), ("n_iter_host", C.c_void_p),
                ("tile_days", C.c_int32), ("out_is_ring", C.c_int32),
                ("on_tile", C.c_void_p), ("user", C.c_void_p),
                ("ms_solar", C.c_double), ("ms_h2d", C.c_double), ("ms_daily", C.c_double),
                ("ms_disagg_total", C.c_double), ("tune", Tuning)]


class SinkDesc(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("elem_bytes", C.c_int32), ("n_grid", C.c_int64),
                ("cell_index", C.c_void_p), ("byteswap", C.c_int32), ("n_threads", C.c_int32),
                ("fd", C.c_int32), ("reserved", C.c_int32)]


# int32_t (*on_tile)(void *user, int32_t tile, int32_t day0, int32_t day1, int32_t slot)
ON_TILE = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32)


class MetsimB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"metsim_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m metsim_b200.build` "
            "(nvcc, sm_100a).  metsim_b200 has no CPU fallback.")
    l = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        getattr(l, name)  # AttributeError if the ABI drifted
    l.msb_version.restype = C.c_int
    l.msb_last_error.restype = C.c_char_p
    l.msb_default_params.restype = None
    l.msb_default_params.argtypes = [C.POINTER(Params)]
    l.msb_solar_prelude_bytes.restype = C.c_size_t
    l.msb_solar_prelude_bytes.argtypes = [C.c_int]
    l.msb_solar_table_sizes.restype = C.c_int
    l.msb_solar_table_sizes.argtypes = [C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_size_t)]
    l.msb_solar_q_win_ld.restype = C.c_size_t
    l.msb_solar_q_win_ld.argtypes = [C.c_int]
    l.msb_solar_q_win_elems.restype = C.c_size_t
    l.msb_solar_q_win_elems.argtypes = [C.c_int, C.c_int]
    l.msb_solar_tables_build.restype = C.c_int
    l.msb_solar_tables_build.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(SolarTables), C.c_int, C.c_void_p]
    l.msb_daily_scratch_bytes.restype = C.c_size_t
    l.msb_daily_scratch_bytes.argtypes = [C.c_int, C.c_int]
    l.msb_mtclim_daily.restype = C.c_int
    l.msb_mtclim_daily.argtypes = [C.POINTER(Params), C.POINTER(Forcing), C.POINTER(SolarTables),
                                   C.POINTER(Daily), C.c_void_p]
    l.msb_daily_check.restype = C.c_int
    l.msb_daily_check.argtypes = [C.POINTER(Forcing), C.POINTER(Daily), C.c_void_p]
    l.msb_disagg_workspace_bytes.restype = C.c_size_t
    l.msb_disagg_workspace_bytes.argtypes = [C.c_int64, C.c_int, C.c_int]
    l.msb_disagg_workspace_bytes_lw.restype = C.c_size_t
    l.msb_disagg_workspace_bytes_lw.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int]
    l.msb_disaggregate.restype = C.c_int
    l.msb_disaggregate.argtypes = [C.POINTER(Params), C.POINTER(Forcing), C.POINTER(SolarTables),
                                   C.POINTER(Daily), C.POINTER(SubDaily), C.c_void_p]
    l.msb_daily_scratch_bytes_chunk.restype = C.c_size_t
    l.msb_daily_scratch_bytes_chunk.argtypes = [C.c_int, C.c_int, C.c_int]
    l.msb_daily_single_pass.restype = C.c_int
    l.msb_daily_single_pass.argtypes = [C.POINTER(Params), C.POINTER(Forcing), C.POINTER(Daily)]
    l.msb_run_host.restype = C.c_int
    l.msb_run_host.argtypes = [C.POINTER(Params), C.POINTER(HostRun)]
    l.msb_host_ctx_create.restype = C.c_int
    l.msb_host_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    l.msb_host_ctx_destroy.restype = None
    l.msb_host_ctx_destroy.argtypes = [C.c_void_p]
    l.msb_host_ctx_device_bytes.restype = C.c_size_t
    l.msb_host_ctx_device_bytes.argtypes = [C.c_void_p]
    l.msb_run_host_bytes.restype = C.c_int
    l.msb_run_host_bytes.argtypes = [C.POINTER(Params), C.POINTER(HostRun), C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_size_t)]
    l.msb_run_host_ctx.restype = C.c_int
    l.msb_run_host_ctx.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(HostRun)]
    l.msb_sink_create.restype = C.c_int
    l.msb_sink_create.argtypes = [C.POINTER(SinkDesc), C.POINTER(C.c_void_p)]
    l.msb_sink_write.restype = C.c_int
    l.msb_sink_write.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    l.msb_sink_checksum.restype = C.c_uint64
    l.msb_sink_checksum.argtypes = [C.c_void_p]
    l.msb_sink_bytes_written.restype = C.c_uint64
    l.msb_sink_bytes_written.argtypes = [C.c_void_p]
    l.msb_sink_destroy.restype = None
    l.msb_sink_destroy.argtypes = [C.c_void_p]
    l.msb_selftest_math.restype = C.c_int
    l.msb_selftest_math.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = l
    return l


def check(rc: int):
    if rc != 0:
        msg = lib().msb_last_error().decode("utf-8", "replace")
        if rc == E_DTR_NEGATIVE:
            raise ValueError(msg)  # same exception type as metsim/metsim.py:612-614
        raise MetsimB200Error(rc, msg)


def make_params(user: dict | None = None) -> Params:
    """Translate the MetSim params dict (metsim/metsim.py:140-171) into the C struct."""
    p = Params()
    lib().msb_default_params(C.byref(p))
    for k, v in (user or {}).items():
        if k == "prec_type":
            p.prec_type = PREC_TYPES[str(v).lower()]
        elif k == "lw_type":
            key = str(v).lower()
            if key not in LW_TYPES:
                raise ValueError("Invalid option given for lw_type")  # metsim.py:673-680
            p.lw_type = LW_TYPES[key]
        elif k == "lw_cloud":
            key = str(v).lower()
            if key not in LW_CLOUD:
                raise ValueError("Invalid option given for lw_cloud")
            p.lw_cloud = LW_CLOUD[key]
        elif k == "utc_offset":
            p.utc_offset = int(bool(v))
        elif k == "method":
            key = str(v).lower()
            if key not in METHODS:
                raise KeyError(key)                        # metsim.py:462: self.methods[...]
            p.method = METHODS[key]
        elif k == "sw_averaging":
            p.sw_daylight = int(str(v).lower() == "daylight")   # disaggregate.py:79
        elif k in ("time_step", "max_iter"):
            setattr(p, k, int(v))
        elif k in ("sw_prec_thresh", "rain_scalar", "tdew_tol", "tmax_daylength_fraction",
                   "tday_coef", "lapse_rate"):
            setattr(p, k, float(v))
    return p
