#!/usr/bin/env python
"""Write minimal, math-only SBML files for the three parvar models.

The reference keeps its ODE models as SBML L3V2 files produced by sbmlutils
(/root/reference/src/parvar/models/sbml/{simple_chain,simple_pk,icg_body_flat}.xml).
About 85 % of those files is RDF annotation, unit definitions and XHTML notes that do
not enter the arithmetic.  The model files are the *input data* of the hot path (SURVEY.md
§2 row 2), and /root/reference does not exist on the GPU box, so this script re-emits each
model keeping only what defines the ODE system:

  compartments (id, size, constant), species (id, compartment, initial value, flags),
  parameters (id, value, constant), assignment/rate rules (variable + MathML),
  reactions (id, stoichiometry, kinetic-law MathML).

Usage:  python tools/strip_sbml.py [/root/reference/src/parvar/models/sbml] [outdir]
"""
from __future__ import annotations

import sys
import xml.etree.ElementTree as ET
from pathlib import Path

CORE = "http://www.sbml.org/sbml/level3/version2/core"
MATHML = "http://www.w3.org/1998/Math/MathML"

KEEP_ATTRS = {
    "compartment": ("id", "size", "constant"),
    "species": ("id", "compartment", "initialAmount", "initialConcentration",
                "hasOnlySubstanceUnits", "boundaryCondition", "constant"),
    "parameter": ("id", "value", "constant"),
    "assignmentRule": ("variable",),
    "rateRule": ("variable",),
    "reaction": ("id", "reversible"),
    "speciesReference": ("species", "stoichiometry", "constant"),
}


def _local(tag: str) -> str:
    return tag.split("}", 1)[1] if "}" in tag else tag


def _copy_math(src: ET.Element) -> ET.Element:
    """Deep-copy a MathML subtree dropping unit attributes (sbml:units on <cn>)."""
    dst = ET.Element(src.tag)
    if _local(src.tag) == "cn" and "type" in src.attrib:
        dst.set("type", src.attrib["type"])
    dst.text = src.text.strip() if src.text and src.text.strip() else None
    for ch in src:
        c = _copy_math(ch)
        c.tail = ch.tail.strip() if ch.tail and ch.tail.strip() else None
        dst.append(c)
    return dst


def _strip(src: ET.Element) -> ET.Element | None:
    name = _local(src.tag)
    if name in ("notes", "annotation", "listOfUnitDefinitions"):
        return None
    if src.tag == f"{{{MATHML}}}math":
        return _copy_math(src)
    dst = ET.Element(src.tag)
    for k in KEEP_ATTRS.get(name, ("id",) if name == "model" else ()):
        if k in src.attrib:
            dst.set(k, src.attrib[k])
    for ch in src:
        c = _strip(ch)
        if c is not None:
            dst.append(c)
    return dst


def strip_file(src: Path, dst: Path) -> None:
    ET.register_namespace("", CORE)
    ET.register_namespace("math", MATHML)
    root = ET.parse(src).getroot()
    out = ET.Element(root.tag, {"level": "3", "version": "2"})
    for ch in root:
        c = _strip(ch)
        if c is not None:
            out.append(c)
    ET.indent(out, space=" ")
    header = ("<?xml version='1.0' encoding='UTF-8'?>\n"
              f"<!-- math-only SBML of parvar model '{src.stem}', written by tools/strip_sbml.py;\n"
              "     annotations, units and notes removed; ODE content unchanged -->\n")
    dst.write_text(header + ET.tostring(out, encoding="unicode") + "\n")


def main() -> None:
    srcdir = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/parvar/models/sbml")
    here = Path(__file__).resolve().parent.parent
    outdir = Path(sys.argv[2]) if len(sys.argv) > 2 else here / "parameter-variability_b200" / "models" / "sbml"
    outdir.mkdir(parents=True, exist_ok=True)
    for name in ("simple_chain", "simple_pk", "icg_body_flat"):
        strip_file(srcdir / f"{name}.xml", outdir / f"{name}.xml")
        print("wrote", outdir / f"{name}.xml")


if __name__ == "__main__":
    main()
