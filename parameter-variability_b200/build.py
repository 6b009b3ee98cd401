"""Build libparvar_b200.so in-tree: SBML -> generated CUDA headers -> nvcc (sm_100a).

The reference needs no build of its own for this path because the native code arrives as
wheels (libroadrunner) or is generated and compiled at run time by AMICI
(``run_optimization.py:43-46`` pre-compiles the model once).  Here the three registry models
are generated and compiled ahead of time; the resulting ``lib/libparvar_b200.so`` travels to
the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libparvar_b200.so"
STAMP = LIB_DIR / "libparvar_b200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: libparvar_b200.so cannot be built (there is no CPU fallback)")
    return exe


def _source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(CSRC.glob("*.cu*")) + sorted((CSRC / "generated").glob("*.cuh")) + \
        sorted((CSRC / "generated").glob("*.inc")) + [PKG_DIR.parent / "include" / "parvar_b200.h"]
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def generate_sources() -> None:
    """Regenerate csrc/generated/* from models/sbml/*.xml (needs sympy)."""
    from . import codegen
    codegen.generate_builtin()


def build_library(force: bool = False, regenerate: bool = True, verbose: bool = False) -> Path:
    """Compile the library if sources changed.  Returns the path of the .so."""
    if regenerate:
        generate_sources()
    LIB_DIR.mkdir(exist_ok=True)
    digest = _source_hash()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH), str(CSRC / "pv_abi.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    proc = subprocess.run(cmd, capture_output=True, text=True, cwd=str(CSRC))
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({' '.join(cmd)}):\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    STAMP.write_text(digest + "\n")
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
