"""Builds tests/host/host_harness.cpp (the device templates compiled as host C++) and wraps it with ctypes."""
from __future__ import annotations

import ctypes
import hashlib
import json
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
GEN = ROOT / "parameter-variability_b200" / "csrc" / "generated"
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_LIB = None


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    srcs = [HERE / "host_harness.cpp"] + sorted((ROOT / "parameter-variability_b200" / "csrc").glob("*.cuh")) + \
        sorted(GEN.glob("*.cuh"))
    h = hashlib.sha256()
    for s in srcs:
        h.update(s.read_bytes())
    out = Path(tempfile.gettempdir()) / f"pv_host_harness_{h.hexdigest()[:12]}.so"
    if not out.exists():
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(out), str(HERE / "host_harness.cpp")],
                       check=True)
    _LIB = ctypes.CDLL(str(out))
    return _LIB


def meta(model: str) -> dict:
    return json.loads((GEN / f"{model}.json").read_text())


def simulate(model: str, overrides: dict, tgrid, rtol: float, atol: float, method: int = 1, sens: bool = False,
             max_steps: int = 1_000_000, stiff_after: int = 0x7fffffff):
    """-> (rc, Y [T, NX] or [T, NX*(1+NS)], (accepted, rejected)).  ``stiff_after``: DOPRI5 stiffness detection."""
    lib = load()
    lib.hh_set_stiff_after(int(stiff_after))
    m = meta(model)
    th = np.array(m["theta_default"], float)
    y0o = np.full(m["NX"], np.nan)
    for k, v in overrides.items():
        k = k.strip("[]")
        if k in m["theta"]:
            th[m["theta"].index(k)] = v
        else:
            y0o[m["states"].index(k)] = v
    N = m["NX"] * (1 + (m["NS"] if sens else 0))
    tg = np.ascontiguousarray(tgrid, float)
    out = np.zeros((len(tg), N))
    stats = np.zeros(2, np.int32)
    rc = lib.hh_simulate(model.encode(), method, int(sens), th.ctypes.data_as(_dp), y0o.ctypes.data_as(_dp),
                         tg.ctypes.data_as(_dp), len(tg), ctypes.c_double(rtol), ctypes.c_double(atol), max_steps,
                         out.ctypes.data_as(_dp), stats.ctypes.data_as(_ip))
    return rc, out, (int(stats[0]), int(stats[1]))


def rhs_jac(model: str, theta: np.ndarray, y: np.ndarray):
    """-> (f [NX], x = (3 I - J)^-1 f via the generated LU, J values on the sparse pattern)."""
    lib = load()
    m = meta(model)
    f = np.zeros(m["NX"])
    buf = np.zeros(m["NX"] + m["NNZ"])
    th = np.ascontiguousarray(theta, float)
    yy = np.ascontiguousarray(y, float)
    rc = lib.hh_rhs(model.encode(), th.ctypes.data_as(_dp), yy.ctypes.data_as(_dp), f.ctypes.data_as(_dp),
                    buf.ctypes.data_as(_dp))
    assert rc == 0
    return f, buf[:m["NX"]], buf[m["NX"]:]


def rodas4_fixed(model: str, overrides: dict, h: float, n: int) -> np.ndarray:
    """n RODAS4 steps of fixed size h from the model's initial state -> final concentrations [NX]."""
    lib = load()
    m = meta(model)
    th = np.array(m["theta_default"], float)
    y0o = np.full(m["NX"], np.nan)
    for k, v in overrides.items():
        if k in m["theta"]:
            th[m["theta"].index(k)] = v
        else:
            y0o[m["states"].index(k)] = v
    out = np.zeros(m["NX"])
    rc = lib.hh_rodas4_fixed(model.encode(), th.ctypes.data_as(_dp), y0o.ctypes.data_as(_dp), ctypes.c_double(h), n,
                             out.ctypes.data_as(_dp))
    assert rc in (0, -1), rc
    return out
