"""ctypes binding of libparvar_b200.so (include/parvar_b200.h) + a thin ``Problem`` object.

PyTorch is used only for what it is good at here: device memory, streams and (in
``distributed.py``) NCCL.  All arithmetic of the path runs in the CUDA library.  There is no
CPU fallback: if the library or a CUDA device is missing, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(__import__("os").environ.get("PARVAR_B200_LIB", PKG_DIR / "lib" / "libparvar_b200.so"))

METHOD_AUTO, METHOD_DOPRI5, METHOD_RODAS4 = 0, 1, 2
NOISE_NORMAL, NOISE_LOGNORMAL = 0, 1
LAYOUT_AUTO, LAYOUT_LANE, LAYOUT_SPLIT = 0, 1, 2
STATUS_TEXT = {0: "ok", 1: "max_steps", 2: "step_too_small", 3: "nonfinite", 4: "stiff"}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol include/parvar_b200.h declares
ABI = {
    "pv_abi_version": (C.c_int, []),
    "pv_last_error": (C.c_char_p, []),
    "pv_device_count": (C.c_int, []),
    "pv_model_count": (C.c_int, []),
    "pv_model_name": (C.c_char_p, [C.c_int]),
    "pv_problem_create": (C.c_void_p, [C.c_char_p, C.c_int]),
    "pv_problem_destroy": (None, [C.c_void_p]),
    "pv_problem_dims": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pv_problem_state_id": (C.c_char_p, [C.c_void_p, C.c_int]),
    "pv_problem_theta_id": (C.c_char_p, [C.c_void_p, C.c_int]),
    "pv_problem_theta_default": (C.c_double, [C.c_void_p, C.c_int]),
    "pv_problem_sens_theta_index": (C.c_int, [C.c_void_p, C.c_int]),
    "pv_problem_flops": (C.c_int, [C.c_void_p, C.c_char_p]),
    "pv_problem_set_solver": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int]),
    "pv_problem_set_value": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "pv_problem_reset_values": (C.c_int, [C.c_void_p]),
    "pv_problem_set_columns": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p)]),
    "pv_problem_set_timepoints": (C.c_int, [C.c_void_p, C.c_double, C.c_int, _dp]),
    "pv_problem_set_schedule": (C.c_int, [C.c_void_p, C.c_int]),
    "pv_problem_set_layout": (C.c_int, [C.c_void_p, C.c_int]),
    "pv_problem_set_stiff_switch": (C.c_int, [C.c_void_p, C.c_int]),
    "pv_problem_set_tuning": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "pv_problem_set_stiffness_detection": (C.c_int, [C.c_void_p, C.c_int]),
    "pv_problem_set_locality": (C.c_int, [C.c_void_p, C.c_int]),
    "pv_problem_set_lockstep": (C.c_int, [C.c_void_p, C.c_int]),
    "pv_problem_set_group_lanes": (C.c_int, [C.c_void_p, C.c_int]),
    "pv_problem_set_observables": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p)]),
    "pv_problem_set_data": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _dp]),
    "pv_simulate": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pv_simulate_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pv_simulate_logp": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "pv_simulate_sens": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "pv_logp": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                          C.c_void_p, C.c_void_p]),
    "pv_logp_grad": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "pv_logp_grad_fim": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "pv_logp_grad_fim_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_void_p]),
    "pv_hier_logp_grad": (C.c_int, [C.c_void_p, C.c_int, C.c_int64] + [C.c_void_p] * 8),
    "pv_hier_finish": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, _dp, _dp, _dp, C.c_void_p]),
    "pv_nuts_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pv_nuts_start": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pv_nuts_drift": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pv_nuts_leaf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "pv_nuts_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pv_nuts_finish": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_int64, C.c_void_p]),
    "pv_hmc_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pv_hmc_drift": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
    "pv_hmc_kick": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.c_void_p]),
    "pv_hmc_accept": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int64,
                                C.c_void_p]),
    "pv_logp_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pv_logp_grad_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pv_launch_count": (C.c_int64, []),
    "pv_sampler_propose": (C.c_int64, [C.c_int64, C.c_int, C.c_int] + [C.c_void_p] * 8),
    "pv_sampler_step": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64] + [C.c_void_p] * 21),
    "pv_sampler_create": (C.c_void_p, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 7 + [C.c_int64, C.c_int]),
    "pv_sampler_advance": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pv_sampler_advance_mt": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "pv_rng_sampler_block": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pv_sampler_result": (C.c_int, [C.c_void_p] * 9),
    "pv_sampler_destroy": (None, [C.c_void_p]),
    "pv_optimizer_create": (C.c_void_p, [C.c_void_p, C.c_int64, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_double] * 3),
    "pv_optimizer_run": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pv_optimizer_result": (C.c_int, [C.c_void_p] * 7),
    "pv_optimizer_destroy": (None, [C.c_void_p]),
    "pv_last_kernel_ms": (C.c_double, [C.c_void_p]),
    "pv_fma_peak_tflops": (C.c_double, [C.c_int, C.c_int, C.c_int]),
    "pv_host_alloc": (C.c_void_p, [C.c_size_t]),
    "pv_host_free": (None, [C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


class ParvarB200Error(RuntimeError):
    pass


_model_libs: dict[str, C.CDLL] = {}      # run-time compiled models (codegen.compile_model): model name -> its library


def load_model_library(name: str, path: Path) -> C.CDLL:
    """dlopen a library built by ``codegen.compile_model`` (same ABI, one model) and register it under the model's name:
    ``Problem(name)`` then finds it."""
    path = Path(path)
    if not path.exists():
        raise ParvarB200Error(f"{path} is missing")
    lib = C.CDLL(str(path))
    for sym, (res, args) in ABI.items():
        fn = getattr(lib, sym)
        fn.restype = res
        fn.argtypes = args
    if lib.pv_abi_version() != 1:
        raise ParvarB200Error("ABI version mismatch")
    _model_libs[name] = lib
    return lib


def load_library(path: Optional[Path] = None) -> C.CDLL:
    """dlopen the in-tree library and type every exported symbol."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    # PARVAR_B200_LIB: an alternative build of the same library (tuning variants, tools/sweep_variants.sh)
    path = Path(path or os.environ.get("PARVAR_B200_LIB") or LIB_PATH)
    if not path.exists():
        raise ParvarB200Error(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs nvcc).  parvar_b200 has no CPU fallback.")
    lib = C.CDLL(str(path))
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)     # AttributeError = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    if lib.pv_abi_version() != 1:
        raise ParvarB200Error("ABI version mismatch")
    _lib = lib
    return lib


def last_error(lib: Optional[C.CDLL] = None) -> str:
    """Message of the last failed call of this thread in `lib` (every library keeps its own); without `lib` the
    base library's, then the run-time compiled model libraries'."""
    libs = [lib] if lib is not None else [load_library(), *_model_libs.values()]
    msgs = [(l.pv_last_error() or b"").decode() for l in libs]
    return next((m for m in msgs if m), "")


def _check(rc: int, lib: Optional[C.CDLL] = None) -> int:
    if rc < 0:
        raise ParvarB200Error(f"parvar_b200 error {rc}: {last_error(lib)}")
    return rc


def _ptr(t) -> Optional[int]:
    """Device (or host) address of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def _cstr_array(ids: Sequence[str]):
    arr = (C.c_char_p * len(ids))()
    arr[:] = [s.encode() for s in ids]
    return arr


def model_names() -> list[str]:
    lib = load_library()
    return [lib.pv_model_name(i).decode() for i in range(lib.pv_model_count())]


class _PinnedBlock:
    """Owner of one cudaHostAlloc block; numpy arrays made from it keep it alive as their base."""

    def __init__(self, nbytes: int, shape, dtype: np.dtype):
        lib = load_library()
        self._free = lib.pv_host_free
        self.addr = lib.pv_host_alloc(max(int(nbytes), 1))
        if not self.addr:
            raise ParvarB200Error(f"pinned allocation of {nbytes} bytes failed: {last_error()}")
        self.__array_interface__ = {"data": (self.addr, False), "shape": tuple(shape), "typestr": dtype.str,
                                    "version": 3}

    def __del__(self):
        if getattr(self, "addr", None):
            self._free(self.addr)
            self.addr = None


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array on page-locked host memory (cudaHostAlloc) so the *_host calls copy at PCIe speed."""
    dtype = np.dtype(dtype)
    shape = tuple(int(s) for s in (shape if hasattr(shape, "__len__") else (shape,)))
    block = _PinnedBlock(int(np.prod(shape)) * dtype.itemsize, shape, dtype)
    return np.asarray(block)


class Problem:
    """One loaded model + its simulation settings (a ``pv_problem``)."""

    def __init__(self, model: str, device: int = 0):
        # a model compiled at run time lives in its own library (load_model_library); the registry models in the main one
        self._lib = _model_libs.get(model) or load_library()
        # "name@fingerprint" = a run-time compiled model: inside its library it is just "name"
        self._h = self._lib.pv_problem_create(model.split("@")[0].encode(), int(device))
        if not self._h:
            raise ParvarB200Error(f"cannot create problem for model {model!r}: {last_error(self._lib)}")
        self.model = model
        self.device = int(device)
        nx, nth, ns = C.c_int(), C.c_int(), C.c_int()
        _check(self._lib.pv_problem_dims(self._h, C.byref(nx), C.byref(nth), C.byref(ns)), self._lib)
        self.n_states, self.n_theta, self.n_sens = nx.value, nth.value, ns.value
        self.state_ids = [self._lib.pv_problem_state_id(self._h, i).decode() for i in range(self.n_states)]
        self.theta_ids = [self._lib.pv_problem_theta_id(self._h, i).decode() for i in range(self.n_theta)]
        self.theta_default = [self._lib.pv_problem_theta_default(self._h, i) for i in range(self.n_theta)]
        self.sens_ids = [self.theta_ids[self._lib.pv_problem_sens_theta_index(self._h, k)] for k in range(self.n_sens)]
        self.columns: list[str] = []
        self.observables: list[str] = []
        self.timepoints = np.zeros(0)
        self.n_data = 0

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.pv_problem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration -------------------------------------------------------------------
    def set_solver(self, method: int = METHOD_AUTO, rtol: float = 1e-7, atol: float = 1e-10,
                   max_steps: int = 1_000_000) -> None:
        _check(self._lib.pv_problem_set_solver(self._h, method, rtol, atol, max_steps), self._lib)

    def set_value(self, sid: str, value: float) -> None:
        _check(self._lib.pv_problem_set_value(self._h, sid.encode(), float(value)), self._lib)

    def reset_values(self) -> None:
        _check(self._lib.pv_problem_reset_values(self._h), self._lib)

    def set_columns(self, ids: Sequence[str]) -> None:
        ids = list(ids)
        _check(self._lib.pv_problem_set_columns(self._h, len(ids), _cstr_array(ids) if ids else None), self._lib)
        self.columns = ids

    def set_schedule(self, refill_min: int = 0) -> None:
        _check(self._lib.pv_problem_set_schedule(self._h, int(refill_min)), self._lib)

    def set_layout(self, layout: int = 0) -> None:
        """LAYOUT_AUTO / LAYOUT_LANE / LAYOUT_SPLIT of the Rosenbrock sensitivity kernels (``pv_problem_set_layout``)."""
        _check(self._lib.pv_problem_set_layout(self._h, int(layout)), self._lib)

    def set_locality(self, mode: int = 0) -> None:
        """Locality schedule of the batch launches: 0 = auto, 1 = off, 2 = on (``pv_problem_set_locality``)."""
        _check(self._lib.pv_problem_set_locality(self._h, int(mode)), self._lib)

    def set_lockstep(self, mode: int = 0) -> None:
        """Lockstep warps of the DOPRI5 kernels: 0 = auto, 1 = off, 2 = on (``pv_problem_set_lockstep``)."""
        _check(self._lib.pv_problem_set_lockstep(self._h, int(mode)), self._lib)

    def set_group_lanes(self, mode: int = 0) -> None:
        """One trajectory per group of 1 + n_sens lanes for small DOPRI5 logp + gradient launches: 0 = auto, 1 = off, 2 = on."""
        _check(self._lib.pv_problem_set_group_lanes(self._h, int(mode)), self._lib)

    def set_stiffness_detection(self, after_steps: int = 128) -> None:
        _check(self._lib.pv_problem_set_stiffness_detection(self._h, int(after_steps)), self._lib)

    def set_tuning(self, max_ctas_per_sm: int = 0, host_chunk_rows: int = 0) -> None:
        """Launch-configuration knobs (``pv_problem_set_tuning``); 0 = default."""
        _check(self._lib.pv_problem_set_tuning(self._h, int(max_ctas_per_sm), int(host_chunk_rows)), self._lib)

    def set_stiff_switch(self, switch_steps: int = 4000) -> None:
        _check(self._lib.pv_problem_set_stiff_switch(self._h, int(switch_steps)), self._lib)

    def set_timepoints(self, t: Sequence[float], t0: Optional[float] = None) -> None:
        """Output grid; ``t0`` (default: the first grid point) is where the initial values hold."""
        t = np.ascontiguousarray(t, dtype=np.float64)
        t0 = float(t[0]) if t0 is None else float(t0)
        _check(self._lib.pv_problem_set_timepoints(self._h, t0, len(t), t.ctypes.data_as(_dp)), self._lib)
        self.timepoints = t.copy()
        self.n_data = 0

    def set_observables(self, ids: Sequence[str]) -> None:
        ids = list(ids)
        _check(self._lib.pv_problem_set_observables(self._h, len(ids), _cstr_array(ids)), self._lib)
        self.observables = [s.strip("[]") for s in ids]
        self.n_data = 0

    def set_data(self, stats: np.ndarray, sigma: Sequence[float], noise_model: int = NOISE_NORMAL) -> None:
        """stats [3][T][n_obs][n_data] sufficient statistics (see include/parvar_b200.h)."""
        stats = np.ascontiguousarray(stats, dtype=np.float64)
        T, K = len(self.timepoints), len(self.observables)
        if stats.ndim != 4 or stats.shape[:3] != (3, T, K):
            raise ValueError(f"stats must have shape (3, {T}, {K}, n_data), got {stats.shape}")
        sigma = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma, dtype=np.float64), (K,)))
        _check(self._lib.pv_problem_set_data(self._h, noise_model, stats.shape[3], stats.ctypes.data_as(_dp),
                                             sigma.ctypes.data_as(_dp)), self._lib)
        self.n_data = stats.shape[3]

    def flops(self, what: str) -> int:
        return _check(self._lib.pv_problem_flops(self._h, what.encode()), self._lib)

    # ---- device-pointer calls (torch CUDA tensors) --------------------------------------------
    def simulate(self, P, Y=None, status=None, steps=None, stream: int = 0, B: Optional[int] = None) -> None:
        B = int(B if B is not None else (P.shape[-1] if P is not None else 0))
        _check(self._lib.pv_simulate(self._h, B, _ptr(P), _ptr(Y), _ptr(status), _ptr(steps), stream), self._lib)

    def simulate_logp(self, P, Y, logp, status=None, steps=None, stream: int = 0, B: Optional[int] = None) -> None:
        B = int(B if B is not None else P.shape[-1])
        _check(self._lib.pv_simulate_logp(self._h, B, _ptr(P), _ptr(Y), _ptr(logp), _ptr(status), _ptr(steps), stream), self._lib)

    def simulate_sens(self, P, Y, S, status=None, steps=None, stream: int = 0, B: Optional[int] = None) -> None:
        B = int(B if B is not None else P.shape[-1])
        _check(self._lib.pv_simulate_sens(self._h, B, _ptr(P), _ptr(Y), _ptr(S), _ptr(status), _ptr(steps), stream), self._lib)

    def logp(self, P, logp, seg_len: int = 0, logp_seg=None, status=None, steps=None, stream: int = 0,
             B: Optional[int] = None) -> None:
        B = int(B if B is not None else P.shape[-1])
        _check(self._lib.pv_logp(self._h, B, _ptr(P), _ptr(logp), seg_len, _ptr(logp_seg), _ptr(status),
                                 _ptr(steps), stream), self._lib)

    def logp_grad(self, P, logp, grad, seg_len: int = 0, logp_seg=None, status=None, steps=None,
                  stream: int = 0, B: Optional[int] = None) -> None:
        B = int(B if B is not None else P.shape[-1])
        _check(self._lib.pv_logp_grad(self._h, B, _ptr(P), _ptr(logp), _ptr(grad), seg_len, _ptr(logp_seg),
                                      _ptr(status), _ptr(steps), stream), self._lib)

    def hier_logp_grad(self, n_chains: int, n_local: int, theta, mu, omega, dtheta, red, status=None, steps=None,
                       stream: int = 0) -> None:
        """Device-resident hierarchical logp + gradient (``pv_hier_logp_grad``): CUDA tensors in, CUDA tensors out,
        two launches on ``stream``, no synchronisation."""
        _check(self._lib.pv_hier_logp_grad(self._h, int(n_chains), int(n_local), _ptr(theta), _ptr(mu), _ptr(omega),
                                           _ptr(dtheta), _ptr(red), _ptr(status), _ptr(steps), stream), self._lib)

    # ---- host-buffer calls (numpy) -----------------------------------------------------------
    def simulate_host(self, P: Optional[np.ndarray], B: int, Y: Optional[np.ndarray] = None,
                      status: Optional[np.ndarray] = None, steps: Optional[np.ndarray] = None,
                      logp: Optional[np.ndarray] = None):
        """P [n_cols, B] float64 -> Y [B, T, n_obs] (and logp [B] when given).  Returns (Y, n_failed)."""
        T, K = len(self.timepoints), len(self.observables)
        if P is not None:
            P = np.ascontiguousarray(P, dtype=np.float64)
            if P.shape != (len(self.columns), B):
                raise ValueError(f"P must have shape ({len(self.columns)}, {B}), got {P.shape}")
        if Y is None:
            Y = np.empty((B, T, K), dtype=np.float64)
        elif Y.shape != (B, T, K) or not Y.flags.c_contiguous:
            raise ValueError(f"Y must be a C-contiguous array of shape ({B}, {T}, {K})")
        n_failed = _check(self._lib.pv_simulate_host(self._h, B, _ptr(P), _ptr(Y), _ptr(logp), _ptr(status), _ptr(steps)), self._lib)
        return Y, n_failed

    def logp_fim_host(self, P: np.ndarray, seg_len: int = 0):
        """-> (logp [B], grad [n_sens, B], fim [n_sens(n_sens+1)/2, B], logp_seg or None, n_failed)."""
        P = np.ascontiguousarray(P, dtype=np.float64)
        B = P.shape[-1]
        lp, g = np.empty(B), np.empty((self.n_sens, B))
        f = np.empty((self.n_sens * (self.n_sens + 1) // 2, B))
        seg = np.empty(B // seg_len) if seg_len else None
        nf = _check(self._lib.pv_logp_grad_fim_host(self._h, B, _ptr(P), _ptr(lp), _ptr(g), _ptr(f), seg_len, _ptr(seg)), self._lib)
        return lp, g, f, seg, nf

    def logp_host(self, P: np.ndarray, seg_len: int = 0, grad: bool = False, out: Optional[np.ndarray] = None):
        """-> (logp [B], grad [n_sens, B] or None, logp_seg [B/seg_len] or None, n_failed).  ``out``: caller's logp buffer
        (e.g. ``pinned_empty(B)`` so the copy back runs at PCIe speed)."""
        P = np.ascontiguousarray(P, dtype=np.float64)
        B = P.shape[-1]
        if out is not None and (out.shape != (B,) or out.dtype != np.float64 or not out.flags.c_contiguous):
            raise ValueError(f"out must be a contiguous float64 array of shape ({B},)")
        lp = np.empty(B) if out is None else out
        seg = np.empty(B // seg_len) if seg_len else None
        if grad:
            g = np.empty((self.n_sens, B))
            nf = _check(self._lib.pv_logp_grad_host(self._h, B, _ptr(P), _ptr(lp), _ptr(g), seg_len, _ptr(seg)), self._lib)
            return lp, g, seg, nf
        nf = _check(self._lib.pv_logp_host(self._h, B, _ptr(P), _ptr(lp), seg_len, _ptr(seg)), self._lib)
        return lp, None, seg, nf

    @property
    def last_kernel_ms(self) -> float:
        return self._lib.pv_last_kernel_ms(self._h)


class SamplerOptions(C.Structure):
    """``pv_sampler_options`` (include/parvar_b200.h)."""
    _fields_ = [("decay_constant", C.c_double), ("target_acceptance_rate", C.c_double), ("reg_factor", C.c_double),
                ("nu", C.c_double), ("eta", C.c_double), ("threshold_sample", C.c_int64), ("adapt_betas", C.c_int32)]


def hier_finish(n_chains: int, n_cols: int, red, mu, omega, mu_prior_mean, mu_prior_sd, omega_prior, stream: int = 0) -> None:
    """Hyper-priors added in place to the (all-reduced) buffer ``red`` [n_chains, 1 + 2 n_cols] (``pv_hier_finish``)."""
    m, s, h = (np.ascontiguousarray(a, dtype=np.float64) for a in (mu_prior_mean, mu_prior_sd, omega_prior))
    _check(load_library().pv_hier_finish(int(n_chains), int(n_cols), _ptr(red), _ptr(mu), _ptr(omega), m.ctypes.data_as(_dp),
                                         s.ctypes.data_as(_dp), h.ctypes.data_as(_dp), stream))


class HMCState(C.Structure):
    """``pv_hmc_state`` (include/parvar_b200.h): sizes and the device addresses of every buffer of the device-resident HMC."""
    POINTERS = ["lth", "mu", "lom", "g_lth", "g_hyp", "lp", "lth_q", "mu_q", "lom_q", "g_lth_q", "g_hyp_q", "p_lth", "p_hyp",
                "im_lth", "im_hyp", "eps", "theta", "omega", "dtheta", "red", "energy", "ok", "da", "w_mean_lth", "w_m2_lth",
                "w_mean_hyp", "w_m2_hyp", "samples_lth", "samples_hyp", "samples_lp", "acc_sum", "n_div"]
    _fields_ = [("n_chains", C.c_int32), ("n_par", C.c_int32), ("own_hyper", C.c_int32), ("reserved", C.c_int32),
                ("n_local", C.c_int64), ("n_total", C.c_int64), ("first", C.c_int64)] + [(n, C.c_void_p) for n in POINTERS]


NUTS_MAX_CHAINS = 64
# slots of pv_nuts_state.vec / .sc / .ic (the enums of include/parvar_b200.h, in order)
NUTS_VEC = ["X", "G", "XL", "RL", "GL", "XR", "RR", "GR", "XP", "GP", "RSUM", "XE", "RE", "GE", "SRSUM", "SX", "SG", "XN", "RN", "GN",
            "IM", "WMEAN", "WM2"]
NUTS_SC = ["LP", "H0", "LPP", "LOGW", "SLOGW", "SLP", "SUMACC", "EPS", "DA_MU", "DA_LEB", "DA_HBAR", "ACCSUM"]
NUTS_IC = ["DEPTH", "DONE", "DIVERGED", "NLEAF", "RIGHT", "LEAF", "STURN", "SDIV", "GROWING", "NDIV"]


class NUTSState(C.Structure):
    """``pv_nuts_state`` (include/parvar_b200.h)."""
    POINTERS = ["vec", "ck", "sc", "ic", "theta", "mu_q", "omega", "dtheta", "red", "samples", "samples_lp", "samples_depth",
                "host_flags"]
    _fields_ = [("n_chains", C.c_int32), ("n_par", C.c_int32), ("max_depth", C.c_int32), ("reserved", C.c_int32),
                ("n_ind", C.c_int64), ("n_samples", C.c_int64)] + [(n, C.c_void_p) for n in POINTERS]


class NUTSUniforms(C.Structure):
    """``pv_nuts_uniforms``: one U(0,1) per chain, passed by value inside the kernel parameters."""
    _fields_ = [("u", C.c_double * NUTS_MAX_CHAINS)]


class MT19937State(C.Structure):
    """``pv_mt19937_state``: numpy's legacy stream as ``RandomState.get_state()`` exports it."""
    _fields_ = [("key", C.c_uint32 * 624), ("pos", C.c_int32), ("has_gauss", C.c_int32), ("gauss", C.c_double)]

    @classmethod
    def from_random_state(cls, rs: "np.random.RandomState") -> "MT19937State":
        name, key, pos, has_gauss, gauss = rs.get_state()
        if name != "MT19937":
            raise ValueError("a legacy MT19937 RandomState is required")
        st = cls()
        C.memmove(st.key, np.ascontiguousarray(key, dtype=np.uint32).ctypes.data, 624 * 4)
        st.pos, st.has_gauss, st.gauss = int(pos), int(has_gauss), float(gauss)
        return st

    def to_random_state(self, rs: "np.random.RandomState") -> None:
        rs.set_state(("MT19937", np.frombuffer(self.key, dtype=np.uint32).copy(), int(self.pos), int(self.has_gauss), float(self.gauss)))


def launch_count() -> int:
    return int(load_library().pv_launch_count())


def fma_peak_tflops(device: int = 0, fp64: bool = True, iters: int = 4096) -> float:
    return float(load_library().pv_fma_peak_tflops(device, 1 if fp64 else 0, iters))
