"""ctypes wrapper of oracle/_build/liboracle.so (the C restatement, parvar_oracle.c).
TEST INFRASTRUCTURE - only tests/, smoke() and bench.py's CPU-baseline legs import this."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liboracle.so"
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)
_lib = None


def build() -> Path:
    subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
    return LIB


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB.exists():
        build()
    lib = C.CDLL(str(LIB))
    lib.pvo_model_index.argtypes = [C.c_char_p]
    lib.pvo_state_name.restype = C.c_char_p
    lib.pvo_param_name.restype = C.c_char_p
    lib.pvo_param_default.restype = C.c_double
    lib.pvo_simulate_batch.argtypes = [C.c_int, C.c_long, C.c_int, _ip, _dp, _dp, _dp, C.c_int, _dp, C.c_int, _ip,
                                       C.c_double, C.c_double, C.c_int, _dp, _dp, C.c_double, _dp, _lp]
    lib.pvo_simulate_batch.restype = C.c_int
    _lib = lib
    return lib


def model_info(model: str):
    lib = load()
    m = lib.pvo_model_index(model.encode())
    if m < 0:
        raise KeyError(model)
    states = [lib.pvo_state_name(m, i).decode() for i in range(lib.pvo_n_states(m))]
    params = [lib.pvo_param_name(m, i).decode() for i in range(lib.pvo_n_params(m))]
    defaults = [lib.pvo_param_default(m, i) for i in range(len(params))]
    return m, states, params, defaults


def max_threads() -> int:
    return int(load().pvo_max_threads())


def simulate_batch(model: str, columns: dict[str, np.ndarray], tgrid: Sequence[float], observables: Sequence[str],
                   overrides: Optional[dict[str, float]] = None, rtol: float = 1e-6, atol: float = 1e-6,
                   threads: int = 0, data: Optional[np.ndarray] = None, sigma: float = 1.0, want_Y: bool = True):
    """-> (Y [T, n_obs, B] or None, logp [B] or None, n_failed, (steps, rejected, rhs evals))."""
    lib = load()
    m, states, params, defaults = model_info(model)
    theta = np.array(defaults, dtype=np.float64)
    y0o = np.full(len(states), np.nan)
    for k, v in (overrides or {}).items():
        k = k.strip("[]")
        if k in params:
            theta[params.index(k)] = v
        else:
            y0o[states.index(k)] = v
    cols = list(columns)
    B = len(next(iter(columns.values()))) if cols else 1
    P = np.ascontiguousarray(np.stack([np.asarray(columns[c], dtype=np.float64) for c in cols])) if cols else np.zeros((1, B))
    par_idx = np.array([params.index(c) for c in cols], dtype=np.int32)
    obs_idx = np.array([states.index(o.strip("[]")) for o in observables], dtype=np.int32)
    tg = np.ascontiguousarray(tgrid, dtype=np.float64)
    Y = np.empty((len(tg), len(obs_idx), B)) if want_Y else None
    lp = np.empty(B) if data is not None else None
    d = np.ascontiguousarray(data, dtype=np.float64) if data is not None else None
    stats = np.zeros(3, dtype=np.int64)
    nf = lib.pvo_simulate_batch(m, B, len(cols), par_idx.ctypes.data_as(_ip), P.ctypes.data_as(_dp),
                                theta.ctypes.data_as(_dp), y0o.ctypes.data_as(_dp), len(tg), tg.ctypes.data_as(_dp),
                                len(obs_idx), obs_idx.ctypes.data_as(_ip), rtol, atol, threads,
                                Y.ctypes.data_as(_dp) if Y is not None else None,
                                d.ctypes.data_as(_dp) if d is not None else None, sigma,
                                lp.ctypes.data_as(_dp) if lp is not None else None, stats.ctypes.data_as(_lp))
    return Y, lp, int(nf), tuple(int(x) for x in stats)
