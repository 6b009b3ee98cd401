"""Minimal SBML L3 reader (stdlib ``xml.etree`` only) -> plain-Python model description.

Replaces, for the hot path, what libSBML/libroadrunner/AMICI do when the reference loads a
model (``roadrunner.RoadRunner(str(model_path))``, reference
``src/parvar/experiments/petab_factory.py:185``; ``PetabImporter.create_problem``,
``src/parvar/optimization/petab_optimization.py:43-45``).  Only the SBML subset the parvar
models use is supported: compartments, species, parameters, assignment
rules, rate rules, reactions with kinetic laws; MathML ``ci cn plus minus times divide
power`` (+ ``exp ln log root abs`` and ``csymbol time`` for generality).  Events, function
definitions, initial assignments, algebraic rules, piecewise and delays raise
``NotImplementedError`` instead of being ignored.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional

import sympy as sp

MATHML_NS = "http://www.w3.org/1998/Math/MathML"
TIME_SYMBOL = sp.Symbol("time")


def _local(tag: str) -> str:
    return tag.split("}", 1)[1] if "}" in tag else tag


@dataclass
class Compartment:
    id: str
    size: float
    constant: bool


@dataclass
class Species:
    id: str
    compartment: str
    initial_amount: Optional[float]
    initial_concentration: Optional[float]
    has_only_substance_units: bool
    boundary_condition: bool
    constant: bool


@dataclass
class Parameter:
    id: str
    value: Optional[float]
    constant: bool


@dataclass
class Reaction:
    id: str
    reactants: list[tuple[str, float]]
    products: list[tuple[str, float]]
    kinetic_law: sp.Expr


@dataclass
class SBMLModel:
    id: str
    compartments: dict[str, Compartment] = field(default_factory=dict)
    species: dict[str, Species] = field(default_factory=dict)
    parameters: dict[str, Parameter] = field(default_factory=dict)
    assignment_rules: dict[str, sp.Expr] = field(default_factory=dict)
    rate_rules: dict[str, sp.Expr] = field(default_factory=dict)
    reactions: list[Reaction] = field(default_factory=list)


# ----------------------------------------------------------------------------------------
# MathML -> sympy
# ----------------------------------------------------------------------------------------
def _cn(node: ET.Element) -> sp.Expr:
    kind = node.attrib.get("type", "real")
    text = (node.text or "").strip()
    if kind == "integer":
        return sp.Integer(int(text))
    if kind == "e-notation":
        # <cn type="e-notation"> mantissa <sep/> exponent </cn>
        sep = node.find(f"{{{MATHML_NS}}}sep")
        mant = text
        expo = (sep.tail or "").strip() if sep is not None else "0"
        return sp.Float(float(f"{mant}e{expo}"))
    if kind == "rational":
        sep = node.find(f"{{{MATHML_NS}}}sep")
        return sp.Rational(int(text), int((sep.tail or "1").strip()))
    val = float(text)
    if val == int(val) and abs(val) < 2**53 and ("." not in text and "e" not in text.lower()):
        return sp.Integer(int(val))
    return sp.Float(val)


def mathml_to_sympy(node: ET.Element) -> sp.Expr:
    """Translate one MathML element (child of <math>) into a sympy expression."""
    name = _local(node.tag)
    if name == "math":
        children = list(node)
        if len(children) != 1:
            raise ValueError("<math> must have exactly one child")
        return mathml_to_sympy(children[0])
    if name == "ci":
        return sp.Symbol((node.text or "").strip())
    if name == "cn":
        return _cn(node)
    if name == "csymbol":
        url = node.attrib.get("definitionURL", "")
        if url.endswith("/time"):
            return TIME_SYMBOL
        raise NotImplementedError(f"csymbol {url!r}")
    if name == "exponentiale":
        return sp.E
    if name == "pi":
        return sp.pi
    if name == "apply":
        children = list(node)
        op = _local(children[0].tag)
        rest = children[1:]
        if op == "root":
            degree = sp.Integer(2)
            operands = []
            for ch in rest:
                if _local(ch.tag) == "degree":
                    degree = mathml_to_sympy(list(ch)[0])
                else:
                    operands.append(mathml_to_sympy(ch))
            return sp.Pow(operands[0], 1 / degree)
        if op == "log":
            base = sp.Integer(10)
            operands = []
            for ch in rest:
                if _local(ch.tag) == "logbase":
                    base = mathml_to_sympy(list(ch)[0])
                else:
                    operands.append(mathml_to_sympy(ch))
            return sp.log(operands[0]) / sp.log(base)
        args = [mathml_to_sympy(ch) for ch in rest]
        if op == "plus":
            return sp.Add(*args)
        if op == "times":
            return sp.Mul(*args)
        if op == "minus":
            if len(args) == 1:
                return -args[0]
            return args[0] - args[1]
        if op == "divide":
            return args[0] / args[1]
        if op == "power":
            return sp.Pow(args[0], args[1])
        if op == "exp":
            return sp.exp(args[0])
        if op == "ln":
            return sp.log(args[0])
        if op == "abs":
            return sp.Abs(args[0])
        raise NotImplementedError(f"MathML operator <{op}> is outside the supported subset")
    raise NotImplementedError(f"MathML element <{name}> is outside the supported subset")


# ----------------------------------------------------------------------------------------
# SBML -> SBMLModel
# ----------------------------------------------------------------------------------------
def _bool(attr: Optional[str], default: bool) -> bool:
    if attr is None:
        return default
    return attr.strip().lower() in ("true", "1")


def _float(attr: Optional[str]) -> Optional[float]:
    return None if attr is None else float(attr)


def read_sbml(path: str | Path) -> SBMLModel:
    """Parse an SBML file.  Works on the full sbmlutils output and on the math-only copies
    in ``models/sbml`` (tools/strip_sbml.py) - both give the same :class:`SBMLModel`."""
    root = ET.parse(str(path)).getroot()
    if _local(root.tag) != "sbml":
        raise ValueError(f"{path}: not an SBML document")
    ns = root.tag.split("}")[0][1:] if "}" in root.tag else ""
    q = (lambda t: f"{{{ns}}}{t}") if ns else (lambda t: t)
    mnode = root.find(q("model"))
    if mnode is None:
        raise ValueError(f"{path}: no <model>")
    model = SBMLModel(id=mnode.attrib.get("id", Path(path).stem))

    for unsupported in ("listOfEvents", "listOfFunctionDefinitions", "listOfInitialAssignments",
                        "listOfConstraints"):
        lst = mnode.find(q(unsupported))
        if lst is not None and len(lst):
            raise NotImplementedError(f"{path}: <{unsupported}> is not supported")

    for c in mnode.findall(f"{q('listOfCompartments')}/{q('compartment')}"):
        model.compartments[c.attrib["id"]] = Compartment(
            id=c.attrib["id"], size=float(c.attrib.get("size", "1")),
            constant=_bool(c.attrib.get("constant"), True))
    for s in mnode.findall(f"{q('listOfSpecies')}/{q('species')}"):
        model.species[s.attrib["id"]] = Species(
            id=s.attrib["id"], compartment=s.attrib["compartment"],
            initial_amount=_float(s.attrib.get("initialAmount")),
            initial_concentration=_float(s.attrib.get("initialConcentration")),
            has_only_substance_units=_bool(s.attrib.get("hasOnlySubstanceUnits"), False),
            boundary_condition=_bool(s.attrib.get("boundaryCondition"), False),
            constant=_bool(s.attrib.get("constant"), False))
    for p in mnode.findall(f"{q('listOfParameters')}/{q('parameter')}"):
        model.parameters[p.attrib["id"]] = Parameter(
            id=p.attrib["id"], value=_float(p.attrib.get("value")),
            constant=_bool(p.attrib.get("constant"), True))
    rules = mnode.find(q("listOfRules"))
    if rules is not None:
        for r in rules:
            kind = _local(r.tag)
            math = r.find(f"{{{MATHML_NS}}}math")
            if kind == "assignmentRule":
                model.assignment_rules[r.attrib["variable"]] = mathml_to_sympy(math)
            elif kind == "rateRule":
                model.rate_rules[r.attrib["variable"]] = mathml_to_sympy(math)
            else:
                raise NotImplementedError(f"{path}: <{kind}> is not supported")
    for r in mnode.findall(f"{q('listOfReactions')}/{q('reaction')}"):
        def refs(tag):
            out = []
            for sr in r.findall(f"{q(tag)}/{q('speciesReference')}"):
                out.append((sr.attrib["species"], float(sr.attrib.get("stoichiometry", "1"))))
            return out
        kl = r.find(q("kineticLaw"))
        if kl is None:
            raise ValueError(f"reaction {r.attrib['id']} has no kineticLaw")
        if kl.find(q("listOfLocalParameters")) is not None and len(kl.find(q("listOfLocalParameters"))):
            raise NotImplementedError("local parameters are not supported")
        math = kl.find(f"{{{MATHML_NS}}}math")
        model.reactions.append(Reaction(id=r.attrib["id"], reactants=refs("listOfReactants"),
                                        products=refs("listOfProducts"),
                                        kinetic_law=mathml_to_sympy(math)))
    return model
