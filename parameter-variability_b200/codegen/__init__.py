"""SBML -> CUDA code generator (north_star subsystem 1).

``generate_model(sbml_path, sens)`` parses an SBML file, derives the hoisted ODE system and
returns the generated header + metadata.  ``generate_builtin()`` (re)writes
``csrc/generated/<model>.cuh`` and ``<model>.json`` for the three models the reference ships
(registry ``BUILTIN``; reference ``src/parvar/__init__.py:26-30``).
"""
from __future__ import annotations

import hashlib
import json
from pathlib import Path
from typing import Sequence

from .emit import STIFF_SPECTRAL_RADIUS, EmittedModel, emit_model, emit_tables, spectral_radius_at_defaults
from .sbml import SBMLModel, read_sbml
from .symbolic import HoistedSystem, OdeSystem, build_ode_system, count_flops, hoist_system

PKG_DIR = Path(__file__).resolve().parent.parent
SBML_DIR = PKG_DIR / "models" / "sbml"
GENERATED_DIR = PKG_DIR / "csrc" / "generated"

# model -> sensitivity parameters compiled into the gradient kernels: the parameters the
# reference samples/estimates (factories/simple_chain.py, simple_pk.py:19-44, icg_body_flat.py:23-42)
BUILTIN: dict[str, list[str]] = {
    "simple_chain": ["k1"],
    "simple_pk": ["k_abs", "CL"],
    "icg_body_flat": ["BW", "LI__ICGIM_Vmax"],
}


def model_fingerprint(ode: OdeSystem, sens: Sequence[str]) -> str:
    """Hash of the mathematical content (not of the file bytes): the reference's annotated SBML
    and the math-only copy give the same fingerprint, hence the same compiled library."""
    h = hashlib.sha256()
    h.update(repr((ode.states, ode.theta, [str(e) for e in ode.f], [str(e) for e in ode.y0],
                   [str(e) for e in ode.obs_scale], [(str(s), str(e)) for s, e in ode.static_rules],
                   list(sens))).encode())
    return h.hexdigest()[:16]


def generate_model(sbml_path: str | Path, sens: Sequence[str] = (), struct_name: str | None = None,
                   name: str | None = None) -> EmittedModel:
    model = read_sbml(sbml_path)
    ode = build_ode_system(model)
    # the registry key is the file stem (reference src/parvar/__init__.py:26-30), not the SBML model id
    # (icg_body_flat.xml carries id "icg_body")
    ode.name = name or Path(sbml_path).stem
    hs = hoist_system(ode, sens)
    em = emit_model(hs, struct_name=struct_name)
    em.meta["fingerprint"] = model_fingerprint(ode, sens)
    em.meta["spectral_radius"] = spectral_radius_at_defaults(hs)
    em.meta["stiff"] = em.meta["spectral_radius"] > STIFF_SPECTRAL_RADIUS
    return em


def generate_builtin(outdir: Path | None = None, sbml_dir: Path | None = None) -> dict[str, EmittedModel]:
    outdir = Path(outdir or GENERATED_DIR)
    sbml_dir = Path(sbml_dir or SBML_DIR)
    outdir.mkdir(parents=True, exist_ok=True)
    out: dict[str, EmittedModel] = {}
    for name, sens in BUILTIN.items():
        em = generate_model(sbml_dir / f"{name}.xml", sens)
        _write_if_changed(outdir / f"{name}.cuh", em.header)
        _write_if_changed(outdir / f"{name}.json", json.dumps(em.meta, indent=1) + "\n")
        out[name] = em
    _write_if_changed(outdir / "model_tables.inc", emit_tables(list(out.values())))
    return out


def _write_if_changed(path: Path, text: str) -> None:
    if path.exists() and path.read_text() == text:
        return
    path.write_text(text)


__all__ = ["BUILTIN", "EmittedModel", "HoistedSystem", "OdeSystem", "SBMLModel", "build_ode_system",
           "count_flops", "emit_model", "generate_builtin", "generate_model", "hoist_system",
           "model_fingerprint", "read_sbml"]
