"""parvar_b200 - B200-native (sm_100a) implementation of the data-parallel hot path of
matthiaskoenig/parameter-variability (``parvar`` 0.2.1): batched forward simulation of the SBML
ODE models over many parameter draws and the log-likelihood + forward-sensitivity gradient.

The directory name carries a hyphen, so import it with
``importlib.import_module("parameter-variability_b200")``.

Layout (mirrors the reference modules that own the path):
  experiments/petab_factory.py   ODESampleSimulator, create_samples_parameters, ...  (reference
                                 src/parvar/experiments/petab_factory.py:94-121, 180-272)
  optimization/petab_optimization.py  objective ``fun/grad`` callables + batched logp_and_grad
                                 (reference src/parvar/optimization/petab_optimization.py:41-48, 144-173)
  codegen/                       SBML -> CUDA code generator
  csrc/                          hand-written CUDA kernels + the C ABI (include/parvar_b200.h)
  runtime.py                     ctypes binding
"""
from pathlib import Path

__version__ = "0.1.0"

PKG_DIR = Path(__file__).resolve().parent
MODELS_DIR = PKG_DIR / "models" / "sbml"

# same keys as reference src/parvar/__init__.py:26-30 (no results/ mkdirs at import)
MODEL_SIMPLE_CHAIN = MODELS_DIR / "simple_chain.xml"
MODEL_SIMPLE_PK = MODELS_DIR / "simple_pk.xml"
MODEL_ICG = MODELS_DIR / "icg_body_flat.xml"
MODELS: dict[str, Path] = {
    "simple_chain": MODEL_SIMPLE_CHAIN,
    "simple_pk": MODEL_SIMPLE_PK,
    "icg_body_flat": MODEL_ICG,
}
