"""SBML model -> symbolic ODE system split into a per-trajectory *hoisted* block and a
per-stage *dynamic* block.

What the reference gets from libroadrunner's LLVM JIT (forward RHS) and from AMICI's sympy
pipeline (Jacobian, forward sensitivities), this module derives for the CUDA code generator:

* ``theta``   - every quantity ``RoadRunner.setValue(pid, v)`` may change before a run and
  that is constant during it (constant parameters, fixed compartment sizes, boundary
  species), reference ``petab_factory.py:231-239``.  They are per-trajectory kernel inputs.
* ``states``  - species with reactions / rate rules (concentration form: the symbol in the
  kinetic laws is the concentration, dX/dt = sum(nu*R)/V; amount form for
  ``hasOnlySubstanceUnits`` species), plus rate-rule parameters.
* ``c``       - hoisted coefficients: every maximal sub-expression of the RHS that does not
  depend on a state or on time.  They are evaluated once per trajectory, so the RHS that
  runs in every Runge-Kutta/Rosenbrock stage is ``f_i = sum_j c_ij * g_j(y)``.
* ``dc``      - d c / d theta_s for the sensitivity parameters (chain rule through all
  parameter-only assignment rules), also evaluated once per trajectory.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, Optional, Sequence

import sympy as sp

from .sbml import SBMLModel, TIME_SYMBOL


@dataclass
class OdeSystem:
    name: str
    states: list[str]
    theta: list[str]
    theta_default: list[float]
    y0: list[sp.Expr]                    # initial state as expression of theta symbols
    obs_scale: list[sp.Expr]             # "[sid]" = y_i * obs_scale_i  (1, or 1/V for amounts)
    static_rules: list[tuple[sp.Symbol, sp.Expr]]   # parameter-only assignment rules, sorted
    f: list[sp.Expr]                     # RHS in states, theta, static-rule symbols, time
    uses_time: bool
    dead_rules: list[str]                # state-dependent rules that feed no ODE (dead outputs)
    rule_values: dict[str, sp.Expr] = field(default_factory=dict)  # every rule, fully substituted

    @property
    def state_syms(self) -> list[sp.Symbol]:
        return [sp.Symbol(s) for s in self.states]

    @property
    def theta_syms(self) -> list[sp.Symbol]:
        return [sp.Symbol(s) for s in self.theta]


def _toposort(rules: dict[str, sp.Expr]) -> list[str]:
    order: list[str] = []
    state: dict[str, int] = {}

    def visit(v: str) -> None:
        if state.get(v) == 2:
            return
        if state.get(v) == 1:
            raise ValueError(f"algebraic loop through assignment rule {v!r}")
        state[v] = 1
        for name in sorted(s.name for s in rules[v].free_symbols):
            if name in rules:
                visit(name)
        state[v] = 2
        order.append(v)

    for v in rules:
        visit(v)
    return order


def build_ode_system(model: SBMLModel) -> OdeSystem:
    """Derive dX/dt for every dynamic quantity of ``model``."""
    assign: dict[str, sp.Expr] = dict(model.assignment_rules)
    for r in model.reactions:
        if r.id in assign:
            raise ValueError(f"id clash: reaction {r.id}")
        assign[r.id] = r.kinetic_law

    # --- which species are integrated ---------------------------------------------------
    net: dict[str, sp.Expr] = {}
    for r in model.reactions:
        rate = sp.Symbol(r.id)
        for sid, nu in r.reactants:
            net[sid] = net.get(sid, sp.Integer(0)) - _num(nu) * rate
        for sid, nu in r.products:
            net[sid] = net.get(sid, sp.Integer(0)) + _num(nu) * rate

    states: list[str] = []
    rhs_raw: dict[str, sp.Expr] = {}
    for sid, s in model.species.items():
        if sid in assign:
            continue                      # algebraic species (assignment rule)
        if sid in model.rate_rules:
            states.append(sid)
            rhs_raw[sid] = model.rate_rules[sid]
            continue
        if s.constant or s.boundary_condition or sid not in net:
            continue                      # constant -> goes to theta
        states.append(sid)
        if s.has_only_substance_units:
            rhs_raw[sid] = net[sid]
        else:
            rhs_raw[sid] = net[sid] / sp.Symbol(s.compartment)
    for vid, expr in model.rate_rules.items():
        if vid in model.species:
            continue
        if vid in model.compartments:
            raise NotImplementedError("time-dependent compartment sizes are not supported")
        states.append(vid)
        rhs_raw[vid] = expr

    # --- theta: everything settable and constant during a run --------------------------------
    theta: list[str] = []
    theta_default: list[float] = []
    for pid, p in model.parameters.items():
        if pid in assign or pid in model.rate_rules:
            continue
        if p.value is None:
            raise ValueError(f"parameter {pid} has no value and no rule")
        theta.append(pid)
        theta_default.append(float(p.value))
    for cid, c in model.compartments.items():
        if cid in assign:
            continue
        theta.append(cid)
        theta_default.append(float(c.size))
    for sid, s in model.species.items():
        if sid in states or sid in assign:
            continue
        theta.append(sid)
        theta_default.append(_species_initial_conc_value(model, sid))

    dyn = {sp.Symbol(s) for s in states} | {TIME_SYMBOL}

    # --- classify rules: static (parameter-only) vs dynamic ---------------------------------
    order = _toposort(assign)
    static_names: set[str] = set()
    full: dict[str, sp.Expr] = {}         # every rule with *dynamic* rules substituted away
    for v in order:
        e = assign[v]
        # substitute dynamic rules (keep static rule symbols as atoms)
        e = e.xreplace({sp.Symbol(k): full[k] for k in full if k not in static_names
                        and sp.Symbol(k) in e.free_symbols})
        full[v] = e
        if not (e.free_symbols & dyn):
            static_names.add(v)
    static_rules = [(sp.Symbol(v), assign[v]) for v in order if v in static_names]

    for cid in model.compartments:
        if cid in assign and cid not in static_names:
            raise NotImplementedError(f"compartment {cid} depends on state/time")

    f: list[sp.Expr] = []
    used_dynamic_rules: set[str] = set()
    for sid in states:
        e = rhs_raw[sid]
        e = e.xreplace({sp.Symbol(k): full[k] for k in full if k not in static_names
                        and sp.Symbol(k) in e.free_symbols})
        f.append(e)
    # dead dynamic rules = those whose symbol never appears in a (pre-substitution) RHS closure
    live: set[str] = set()
    stack = [s.name for sid in states for s in rhs_raw[sid].free_symbols if s.name in assign]
    while stack:
        v = stack.pop()
        if v in live:
            continue
        live.add(v)
        stack.extend(s.name for s in assign[v].free_symbols if s.name in assign)
    dead_rules = [v for v in model.assignment_rules if v not in live]
    static_rules = [(s, e) for (s, e) in static_rules if s.name in live or _feeds_y0_or_obs(model, s.name)]

    # --- initial values / observable scaling ------------------------------------------------
    y0: list[sp.Expr] = []
    obs_scale: list[sp.Expr] = []
    for sid in states:
        if sid in model.species:
            s = model.species[sid]
            V = sp.Symbol(s.compartment)
            if s.has_only_substance_units:
                if s.initial_amount is not None:
                    y0.append(_num(s.initial_amount))
                else:
                    y0.append(_num(s.initial_concentration or 0.0) * V)
                obs_scale.append(1 / V)
            else:
                if s.initial_concentration is not None:
                    y0.append(_num(s.initial_concentration))
                else:
                    y0.append(_num(s.initial_amount or 0.0) / V)
                obs_scale.append(sp.Integer(1))
        else:
            y0.append(_num(model.parameters[sid].value or 0.0))
            obs_scale.append(sp.Integer(1))

    # all rules fully substituted down to theta/states (for the oracle cross-check & tests)
    rule_values: dict[str, sp.Expr] = {}
    for v in order:
        e = assign[v]
        rule_values[v] = e.xreplace({sp.Symbol(k): rule_values[k] for k in rule_values
                                     if sp.Symbol(k) in e.free_symbols})

    uses_time = any(TIME_SYMBOL in e.free_symbols for e in f)
    return OdeSystem(name=model.id, states=states, theta=theta, theta_default=theta_default,
                     y0=y0, obs_scale=obs_scale, static_rules=static_rules, f=f,
                     uses_time=uses_time, dead_rules=dead_rules, rule_values=rule_values)


def _feeds_y0_or_obs(model: SBMLModel, rule: str) -> bool:
    return rule in model.compartments


def _num(x: float) -> sp.Expr:
    x = float(x)
    return sp.Integer(int(x)) if x == int(x) and abs(x) < 2**53 else sp.Float(x)


def _species_initial_conc_value(model: SBMLModel, sid: str) -> float:
    s = model.species[sid]
    if s.initial_concentration is not None:
        return float(s.initial_concentration)
    amount = float(s.initial_amount or 0.0)
    if s.has_only_substance_units:
        return amount
    return amount / model.compartments[s.compartment].size


# ----------------------------------------------------------------------------------------
# hoisting
# ----------------------------------------------------------------------------------------
class Hoister:
    """Collects state-free sub-expressions and names them c0, c1, ..."""

    def __init__(self, dyn: Iterable[sp.Symbol], prefix: str = "c"):
        self.dyn = set(dyn)
        self.prefix = prefix
        self.table: dict[sp.Expr, sp.Symbol] = {}
        self.exprs: list[sp.Expr] = []

    def is_static(self, e: sp.Expr) -> bool:
        return not (e.free_symbols & self.dyn)

    def hoist(self, e: sp.Expr) -> sp.Expr:
        if e.is_Number:
            return e
        if e in self.table:
            return self.table[e]
        neg = -e
        if neg in self.table:
            return -self.table[neg]
        # prefer storing the form without a leading minus sign
        if e.could_extract_minus_sign():
            return -self.hoist(neg)
        sym = sp.Symbol(f"{self.prefix}{len(self.exprs)}")
        self.table[e] = sym
        self.exprs.append(e)
        return sym

    def visit(self, e: sp.Expr) -> sp.Expr:
        if self.is_static(e):
            return self.hoist(e)
        if e.is_Symbol:
            return e
        if e.is_Add or e.is_Mul:
            stat = [a for a in e.args if self.is_static(a)]
            dyn = [self.visit(a) for a in e.args if not self.is_static(a)]
            if stat:
                return e.func(self.hoist(e.func(*stat)), *dyn)
            return e.func(*dyn)
        return e.func(*[self.visit(a) for a in e.args])

    def canonical_rhs(self, e: sp.Expr) -> sp.Expr:
        """f_i = sum_j coeff_j(theta) * g_j(y): expand, group by state-dependent factor."""
        dyn = sorted(self.dyn, key=lambda s: s.name)
        groups: dict[sp.Expr, sp.Expr] = {}
        for term in sp.Add.make_args(sp.expand(e)):
            coeff, dep = term.as_independent(*dyn, as_Add=False)
            groups[dep] = groups.get(dep, sp.Integer(0)) + coeff
        out = sp.Integer(0)
        for dep, coeff in groups.items():
            if coeff == 0:
                continue
            coeff = sp.factor_terms(coeff)
            if dep == 1:
                out += self.hoist(coeff)
            else:
                out += self.hoist(coeff) * self.visit(dep)
        return out


@dataclass
class HoistedSystem:
    """Everything the emitters need, in symbols y*, c*, dc*_*, s*_*."""
    ode: OdeSystem
    sens: list[str]                        # sensitivity parameters (subset of theta)
    c_syms: list[sp.Symbol]
    c_exprs: list[sp.Expr]                 # in theta + static-rule symbols
    c_exprs_theta: list[sp.Expr]           # static rules substituted: theta only
    dc_exprs: list[list[sp.Expr]]          # [sens][c]  d c / d theta_s, theta only
    y0_exprs: list[sp.Expr]                # theta only
    dy0_exprs: list[list[sp.Expr]]         # [sens][state]
    obs_scale_exprs: list[sp.Expr]         # theta only
    f: list[sp.Expr]                       # in y symbols, c symbols (, time)
    jac: dict[tuple[int, int], sp.Expr]    # sparse d f_i / d y_j
    dfdc: dict[tuple[int, int], sp.Expr]   # sparse d f_i / d c_m
    jac_const: dict[tuple[int, int], bool]  # True when the entry is state independent


def hoist_system(ode: OdeSystem, sens: Sequence[str] = ()) -> HoistedSystem:
    for s in sens:
        if s not in ode.theta:
            raise ValueError(f"sensitivity parameter {s!r} is not a settable constant of "
                             f"model {ode.name} (theta = {ode.theta})")
    dyn = set(ode.state_syms) | {TIME_SYMBOL}
    h = Hoister(dyn)
    f = [h.canonical_rhs(e) for e in ode.f]

    static_map: dict[sp.Symbol, sp.Expr] = {}
    for sym, expr in ode.static_rules:
        static_map[sym] = expr.xreplace(static_map)

    def to_theta(e: sp.Expr) -> sp.Expr:
        return e.xreplace(static_map)

    c_syms = [sp.Symbol(f"c{i}") for i in range(len(h.exprs))]
    c_theta = [to_theta(e) for e in h.exprs]
    sens_syms = [sp.Symbol(s) for s in sens]
    dc = [[sp.diff(e, s) for e in c_theta] for s in sens_syms]
    y0 = [to_theta(e) for e in ode.y0]
    dy0 = [[sp.diff(e, s) for e in y0] for s in sens_syms]
    obs = [to_theta(e) for e in ode.obs_scale]

    ysyms = ode.state_syms
    jac: dict[tuple[int, int], sp.Expr] = {}
    jac_const: dict[tuple[int, int], bool] = {}
    for i, fi in enumerate(f):
        for j, yj in enumerate(ysyms):
            if yj in fi.free_symbols:
                d = sp.diff(fi, yj)
                if d != 0:
                    jac[(i, j)] = d
                    jac_const[(i, j)] = not (d.free_symbols & dyn)
    dfdc: dict[tuple[int, int], sp.Expr] = {}
    for i, fi in enumerate(f):
        for m, cm in enumerate(c_syms):
            if cm in fi.free_symbols:
                d = sp.diff(fi, cm)
                if d != 0:
                    dfdc[(i, m)] = d
    return HoistedSystem(ode=ode, sens=list(sens), c_syms=c_syms, c_exprs=list(h.exprs),
                         c_exprs_theta=c_theta, dc_exprs=dc, y0_exprs=y0, dy0_exprs=dy0,
                         obs_scale_exprs=obs, f=f, jac=jac, dfdc=dfdc, jac_const=jac_const)


# ----------------------------------------------------------------------------------------
# flop counting on expression DAGs (mul/add = 1, div = 1, pow/exp/log = 1)
# ----------------------------------------------------------------------------------------
def count_flops(exprs: Iterable[sp.Expr]) -> int:
    """Flops of evaluating ``exprs`` after common-subexpression elimination."""
    exprs = [sp.sympify(e) for e in exprs]
    repl, reduced = sp.cse(exprs, optimizations="basic") if exprs else ([], [])
    total = 0
    for e in [r for _, r in repl] + list(reduced):
        total += _flops(e)
    return total


def _flops(e: sp.Expr) -> int:
    if e.is_Atom:
        return 0
    n = sum(_flops(a) for a in e.args)
    if e.is_Add:
        return n + len(e.args) - 1
    if e.is_Mul:
        k = len(e.args) - 1
        # a leading -1 is a sign flip folded into the neighbouring add/fma
        if e.args[0] == -1:
            k -= 1
        return n + max(k, 0)
    if e.is_Pow:
        b, p = e.args
        if p.is_Integer:
            ip = abs(int(p))
            return n + max(ip - 1, 0) + (1 if int(p) < 0 else 0)
        return n + 1
    return n + 1
