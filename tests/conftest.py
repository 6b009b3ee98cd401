"""Shared test plumbing.  `-m "not gpu"` tests run in the GPU-less build container; `-m gpu` tests are the
parity tests proper and call the CUDA library through its C ABI on a B200."""
import importlib
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

PKG = "parameter-variability_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pkg(sub: str = ""):
    return importlib.import_module(PKG + (("." + sub) if sub else ""))


@pytest.fixture(scope="session")
def oracle():
    import parvar_oracle
    return parvar_oracle


@pytest.fixture(scope="session")
def rt():
    return pkg("runtime")


def has_cuda() -> bool:
    try:
        return pkg("runtime").load_library().pv_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def cuda_lib(rt):
    lib = rt.load_library()
    if lib.pv_device_count() <= 0:
        pytest.fail("gpu test selected but no CUDA device is visible (there is no CPU fallback)")
    return lib


def lognormal_draws(seed: int, n: int, mean: float = 1.0, sd: float = 1.0) -> np.ndarray:
    """np.random.seed(seed); lognormal draws with the reference's transform (petab_factory.py:94-98)."""
    sigma_ln = np.sqrt(np.log(1 + (sd / mean) ** 2))
    mu_ln = np.log(mean) - sigma_ln ** 2 / 2
    rs = np.random.RandomState(seed)
    return rs.lognormal(mu_ln, sigma_ln, size=n)


def parity_violation(got: np.ndarray, ref: np.ndarray, rtol: float = 1e-6, atol: float = 1e-9) -> float:
    """max |got-ref| / (atol + rtol*|ref|); parity holds when <= 1 (north_star: rtol 1e-6 / atol 1e-9)."""
    return float(np.max(np.abs(got - ref) / (atol + rtol * np.abs(ref))))
