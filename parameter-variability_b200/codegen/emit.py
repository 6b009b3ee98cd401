"""Emit a model's hoisted ODE system as straight-line C++/CUDA device functions.

Plays the role of libroadrunner's LLVM code generation and AMICI's generated C++ model
(reference call sites ``petab_factory.py:185`` and ``petab_optimization.py:43-45``) for the
sm_100a integrators in ``csrc/``: one header per model with

  hoist()        theta -> c[], y0[], obs_scale[]          once per trajectory
  hoist_sens()   theta -> dc[] (compact), dy0[][]         once per trajectory (gradient mode)
  rhs()          f(t, y; c)                                every stage
  rhs_sens()     f and J*s_k + (df/dc) dc_k                every stage (gradient mode)
  jac()          sparse J values (Rosenbrock)              every step
  lu_factor()/lu_solve()   symbolic sparse LU of W = d*I - J, fully unrolled, no pivoting
  sens_stage_term()        (d(J s + f_theta)/dy) k  directional term for Rosenbrock sensitivities

All arrays are indexed with literals only, so after inlining they live in registers.
The same text compiles as host C++ (``PV_HD`` empty) for the CPU-side checks in tests/.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import sympy as sp
from sympy.printing.c import C99CodePrinter

from .sbml import TIME_SYMBOL
from .symbolic import HoistedSystem, count_flops


class _Printer(C99CodePrinter):
    def __init__(self, symmap: dict[sp.Symbol, str]):
        super().__init__({"precision": 17, "user_functions": {"pv_rcp_newton": "pv_rcp_newton"}})
        self._symmap = symmap

    def _print_Symbol(self, expr):
        return self._symmap.get(expr, expr.name)

    def _print_Float(self, expr):
        return repr(float(expr))

    def _print_Integer(self, expr):
        return f"{int(expr)}.0"

    def _print_Rational(self, expr):
        return f"({int(expr.p)}.0/{int(expr.q)}.0)"

    def _print_Pow(self, expr):
        b, e = expr.args
        if e.is_Integer and 1 < abs(int(e)) <= 4:
            s = "*".join([self.parenthesize(b, 50)] * abs(int(e)))
            return f"1.0/({s})" if int(e) < 0 else f"({s})"
        if e == -1:
            return f"1.0/{self.parenthesize(b, 50)}"
        if e == sp.Rational(1, 2):
            return f"sqrt({self._print(b)})"
        return f"pow({self._print(b)}, {self._print(e)})"


_RCP = sp.Function("pv_rcp_newton")


def _fast_rcp(e: sp.Expr) -> sp.Expr:
    """b**-n -> pv_rcp_newton(b)**n (before CSE, so 1/b and 1/b^2 share one reciprocal): the per-stage functions divide by Michaelis-Menten denominators, which are normal
    finite numbers; the branch-free reciprocal (MUFU.RCP64H + two Newton steps, pv_common.cuh) replaces the IEEE
    division, whose slow-path branch is a scheduling barrier in the middle of straight-line code."""
    return e.replace(lambda x: x.is_Pow and x.exp.is_Integer and x.exp < 0 and not x.base.is_Number,
                     lambda x: _RCP(x.base) ** (-x.exp))


def _emit_block(assignments: Sequence[tuple[str, sp.Expr]], symmap: dict[sp.Symbol, str],
                tmp_prefix: str, indent: str = "    ", fast_rcp: bool = False) -> tuple[str, int]:
    """CSE + print ``lhs = expr;`` lines.  Returns (code, flops); a division counts one flop either way."""
    exprs = [sp.sympify(e) for _, e in assignments]
    if not exprs:
        return "", 0
    flops = count_flops(exprs)
    if fast_rcp:
        exprs = [_fast_rcp(e) for e in exprs]
    repl, reduced = sp.cse(exprs, symbols=sp.numbered_symbols(tmp_prefix), optimizations="basic")
    pr = _Printer(symmap)
    lines = []
    for sym, e in repl:
        lines.append(f"{indent}const double {sym.name} = {pr.doprint(e)};")
    for (lhs, _), e in zip(assignments, reduced):
        lines.append(f"{indent}{lhs} = {pr.doprint(e)};")
    return "\n".join(lines), flops


# ----------------------------------------------------------------------------------------
# symbolic sparse LU (no pivoting) of  W = d*I - J  on the structural pattern of J
# ----------------------------------------------------------------------------------------
@dataclass
class SparseLU:
    perm: list[int]                       # elimination order (symmetric permutation)
    pattern: list[tuple[int, int]]        # filled pattern in permuted indices, row-major
    ops_factor: list[tuple]               # ("div", (i,k), (k,k)) / ("fnma", (i,j), (i,k), (k,j))
    n: int

    @property
    def nlu(self) -> int:
        return len(self.pattern)


def symbolic_lu(n: int, nz: set[tuple[int, int]]) -> SparseLU:
    """Greedy minimum-degree (Markowitz on the symmetrised pattern) ordering + fill-in."""
    pat = {(i, j) for (i, j) in nz} | {(i, i) for i in range(n)}
    remaining = set(range(n))
    rows = {i: {j for (a, j) in pat if a == i} for i in range(n)}
    cols = {j: {i for (i, b) in pat if b == j} for j in range(n)}
    perm: list[int] = []
    work_rows = {i: set(s) for i, s in rows.items()}
    work_cols = {j: set(s) for j, s in cols.items()}
    while remaining:
        # Markowitz count (r-1)(c-1) restricted to the remaining sub-matrix
        best = min(remaining, key=lambda k: ((len(work_rows[k] & remaining) - 1) *
                                             (len(work_cols[k] & remaining) - 1), k))
        perm.append(best)
        remaining.discard(best)
        rr = work_cols[best] & remaining          # rows that have an entry in column `best`
        cc = work_rows[best] & remaining          # columns that have an entry in row `best`
        for i in rr:
            for j in cc:
                work_rows[i].add(j)
                work_cols[j].add(i)
    inv = {old: new for new, old in enumerate(perm)}
    P = {(inv[i], inv[j]) for (i, j) in pat}
    ops: list[tuple] = []
    for k in range(n):
        below = sorted(i for i in range(k + 1, n) if (i, k) in P)
        right = sorted(j for j in range(k + 1, n) if (k, j) in P)
        for i in below:
            ops.append(("div", (i, k), (k, k)))
            for j in right:
                P.add((i, j))
                ops.append(("fnma", (i, j), (i, k), (k, j)))
    return SparseLU(perm=perm, pattern=sorted(P), ops_factor=ops, n=n)


# ----------------------------------------------------------------------------------------
@dataclass
class EmittedModel:
    name: str
    struct_name: str
    header: str
    meta: dict


def emit_model(hs: HoistedSystem, struct_name: str | None = None) -> EmittedModel:
    ode = hs.ode
    name = ode.name
    struct_name = struct_name or f"pv_model_{name}"
    NX, NTH, NC, NS = len(ode.states), len(ode.theta), len(hs.c_syms), len(hs.sens)
    ysyms = ode.state_syms
    symmap: dict[sp.Symbol, str] = {}
    for i, s in enumerate(ysyms):
        symmap[s] = f"y[{i}]"
    for k, s in enumerate(ode.theta_syms):
        symmap[s] = f"th[{k}]"
    for m, s in enumerate(hs.c_syms):
        symmap[s] = f"c[{m}]"
    symmap[TIME_SYMBOL] = "t"

    # ---- hoist ---------------------------------------------------------------------------
    hoist_assign = [(f"c[{m}]", e) for m, e in enumerate(hs.c_exprs_theta)]
    hoist_assign += [(f"y0[{i}]", e) for i, e in enumerate(hs.y0_exprs)]
    hoist_assign += [(f"obs[{i}]", e) for i, e in enumerate(hs.obs_scale_exprs)]
    hoist_code, hoist_flops = _emit_block(hoist_assign, symmap, "h")

    # ---- hoist_sens: compact dc ----------------------------------------------------------
    dc_index: dict[tuple[int, int], int] = {}
    dc_assign = []
    for s in range(NS):
        for m in range(NC):
            e = hs.dc_exprs[s][m]
            if e != 0:
                dc_index[(s, m)] = len(dc_index)
                dc_assign.append((f"dc[{dc_index[(s, m)]}]", e))
    NDC = max(len(dc_index), 1)
    dy0_assign = [(f"dy0[{s}][{i}]", hs.dy0_exprs[s][i]) for s in range(NS) for i in range(NX)]
    hs_code, hs_flops = _emit_block(dc_assign + dy0_assign, symmap, "g")
    dc_syms = {k: sp.Symbol(f"dc_{k[0]}_{k[1]}") for k in dc_index}
    for k, sym in dc_syms.items():
        symmap[sym] = f"dc[{dc_index[k]}]"

    # ---- rhs -----------------------------------------------------------------------------
    rhs_assign = [(f"f[{i}]", e) for i, e in enumerate(hs.f)]
    rhs_code, rhs_flops = _emit_block(rhs_assign, symmap, "r", fast_rcp=True)

    # ---- rhs_sens ------------------------------------------------------------------------
    ssyms = [[sp.Symbol(f"s_{k}_{j}") for j in range(NX)] for k in range(NS)]
    for k in range(NS):
        for j in range(NX):
            symmap[ssyms[k][j]] = f"s[{k}][{j}]"
    fs_exprs = []
    for k in range(NS):
        for i in range(NX):
            e = sp.Integer(0)
            for j in range(NX):
                if (i, j) in hs.jac:
                    e += hs.jac[(i, j)] * ssyms[k][j]
            for m in range(NC):
                if (i, m) in hs.dfdc and (k, m) in dc_syms:
                    e += hs.dfdc[(i, m)] * dc_syms[(k, m)]
            fs_exprs.append((f"fs[{k}][{i}]", e))
    rs_code, rs_flops = _emit_block(rhs_assign + fs_exprs, symmap, "q", fast_rcp=True)

    # ---- jacobian (sparse, row-major pattern) ------------------------------------------------
    pattern = sorted(hs.jac.keys())
    NNZ = max(len(pattern), 1)
    jac_assign = [(f"J[{n}]", hs.jac[ij]) for n, ij in enumerate(pattern)]
    jac_code, jac_flops = _emit_block(jac_assign, symmap, "j", fast_rcp=True)
    n_const = sum(1 for ij in pattern if hs.jac_const[ij])

    # ---- LU of W = d*I - J --------------------------------------------------------------------
    lu = symbolic_lu(NX, set(pattern))
    inv = {old: new for new, old in enumerate(lu.perm)}
    lu_pos = {ij: n for n, ij in enumerate(lu.pattern)}
    NLU = lu.nlu
    asm = [f"    W[{n}] = 0.0;" for n in range(NLU)]
    for n, (i, j) in enumerate(pattern):
        asm[lu_pos[(inv[i], inv[j])]] = f"    W[{lu_pos[(inv[i], inv[j])]}] = -J[{n}];"
    for i in range(NX):
        p = lu_pos[(i, i)]
        old = lu.perm[i]
        if (old, old) in hs.jac:
            asm[p] = f"    W[{p}] = d - J[{pattern.index((old, old))}];"
        else:
            asm[p] = f"    W[{p}] = d;"
    fac = []
    lu_flops = 0
    # diagonal reciprocals are stored in place of U_kk so the solves multiply instead of divide
    done_diag: set[int] = set()
    for op in lu.ops_factor:
        if op[0] == "div":
            _, ik, kk = op
            k = kk[0]
            if k not in done_diag:
                fac.append(f"    W[{lu_pos[kk]}] = pv_rcp_newton(W[{lu_pos[kk]}]);")
                done_diag.add(k)
                lu_flops += 1
            fac.append(f"    W[{lu_pos[ik]}] *= W[{lu_pos[kk]}];")
            lu_flops += 1
        else:
            _, ij, ik, kj = op
            fac.append(f"    W[{lu_pos[ij]}] = fma(-W[{lu_pos[ik]}], W[{lu_pos[kj]}], W[{lu_pos[ij]}]);")
            lu_flops += 2
    for k in range(NX):
        if k not in done_diag:
            fac.append(f"    W[{lu_pos[(k, k)]}] = pv_rcp_newton(W[{lu_pos[(k, k)]}]);")
            lu_flops += 1
    sol = []
    sol_flops = 0
    sol.append("    double x[NX];")
    for new in range(NX):
        sol.append(f"    x[{new}] = b[{lu.perm[new]}];")
    for i in range(NX):          # forward (unit lower)
        for j in range(i):
            if (i, j) in lu_pos:
                sol.append(f"    x[{i}] = fma(-W[{lu_pos[(i, j)]}], x[{j}], x[{i}]);")
                sol_flops += 2
    for i in reversed(range(NX)):  # backward (U with reciprocal diagonal)
        for j in range(i + 1, NX):
            if (i, j) in lu_pos:
                sol.append(f"    x[{i}] = fma(-W[{lu_pos[(i, j)]}], x[{j}], x[{i}]);")
                sol_flops += 2
        sol.append(f"    x[{i}] *= W[{lu_pos[(i, i)]}];")
        sol_flops += 1
    for new in range(NX):
        sol.append(f"    b[{lu.perm[new]}] = x[{new}];")

    # ---- Rosenbrock sensitivity stage term: (d/dy [J s_k + f_theta_k]) . v ----------------------
    vsyms = [sp.Symbol(f"v_{j}") for j in range(NX)]
    for j in range(NX):
        symmap[vsyms[j]] = f"v[{j}]"
    st_exprs = []
    dyn_jac = [ij for ij in pattern if not hs.jac_const[ij]]
    for k in range(NS):
        for i in range(NX):
            base = sp.Integer(0)
            for j in range(NX):
                if (i, j) in hs.jac and not hs.jac_const[(i, j)]:
                    base += hs.jac[(i, j)] * ssyms[k][j]
            for m in range(NC):
                if (i, m) in hs.dfdc and (k, m) in dc_syms:
                    base += hs.dfdc[(i, m)] * dc_syms[(k, m)]
            e = sp.Integer(0)
            for l, yl in enumerate(ysyms):
                if yl in base.free_symbols:
                    e += sp.diff(base, yl) * vsyms[l]
            st_exprs.append((f"out[{k}][{i}]", e))
    st_code, st_flops = _emit_block(st_exprs, symmap, "w", fast_rcp=True)

    # ---- J*v product (used by Rosenbrock sensitivities: J s) ------------------------------------
    jv_lines = []
    for i in range(NX):
        terms = [f"J[{n}]*v[{j}]" for n, (a, j) in enumerate(pattern) if a == i]
        jv_lines.append(f"    out[{i}] = {' + '.join(terms) if terms else '0.0'};")
    jv_flops = 2 * len(pattern) - NX

    # ---- df/dtheta_k (explicit part; used by Rosenbrock sensitivities) --------------------------
    fp_exprs = []
    for k in range(NS):
        for i in range(NX):
            e = sp.Integer(0)
            for m in range(NC):
                if (i, m) in hs.dfdc and (k, m) in dc_syms:
                    e += hs.dfdc[(i, m)] * dc_syms[(k, m)]
            fp_exprs.append((f"fp[{k}][{i}]", e))
    fp_code, fp_flops = _emit_block(fp_exprs, symmap, "p", fast_rcp=True)

    # ---- per-block forms (one sensitivity parameter at a time) for the block-per-warp Rosenbrock kernel -----
    # s[j] / v[j] are the block's own vectors; c, dc, W may be any indexable container (registers or shared memory)
    bsyms = [sp.Symbol(f"sb_{j}") for j in range(NX)]
    symmap_b = dict(symmap)
    for j in range(NX):
        symmap_b[bsyms[j]] = f"s[{j}]"
    blk_rhs, blk_term, blk_prep, blk_apply = [], [], [], []
    blk_rhs_flops, blk_term_flops, blk_prep_flops, blk_apply_flops, blk_nhc = [], [], [], [], []

    def kw_of(k):
        return "if" if k == 0 else "else if"
    for k in range(NS):
        exprs_r, exprs_t = [], []
        for i in range(NX):
            e = sp.Integer(0)
            base = sp.Integer(0)
            for j in range(NX):
                if (i, j) in hs.jac:
                    e += hs.jac[(i, j)] * bsyms[j]
                    if not hs.jac_const[(i, j)]:
                        base += hs.jac[(i, j)] * bsyms[j]
            for m in range(NC):
                if (i, m) in hs.dfdc and (k, m) in dc_syms:
                    e += hs.dfdc[(i, m)] * dc_syms[(k, m)]
                    base += hs.dfdc[(i, m)] * dc_syms[(k, m)]
            exprs_r.append((f"out[{i}]", e))
            d = sp.Integer(0)
            for l, yl in enumerate(ysyms):
                if yl in base.free_symbols:
                    d += sp.diff(base, yl) * vsyms[l]
            exprs_t.append((f"out[{i}]", d))
        code_r, fl_r = _emit_block(exprs_r, symmap_b, f"a{k}_", indent="            ", fast_rcp=True)
        # the stage term is linear in v: out_i += sum_l H_il v_l.  Entries that are more than a (signed, scaled) symbol are
        # evaluated once per step (sens_term_prep -> hc[]); sens_term_apply does the sparse product at every stage.
        prep_assign, apply_lines, n_mul = [], [], 0
        pr_b = _Printer(symmap_b)
        for i in range(NX):
            base = sp.Integer(0)
            for j in range(NX):
                if (i, j) in hs.jac and not hs.jac_const[(i, j)]:
                    base += hs.jac[(i, j)] * bsyms[j]
            for m in range(NC):
                if (i, m) in hs.dfdc and (k, m) in dc_syms:
                    base += hs.dfdc[(i, m)] * dc_syms[(k, m)]
            terms = []
            for l, yl in enumerate(ysyms):
                if yl not in base.free_symbols:
                    continue
                ent = sp.diff(base, yl)
                if ent == 0:
                    continue
                coeff, rest = ent.as_coeff_Mul()
                if rest.is_Symbol or rest == 1:
                    terms.append(f"({pr_b.doprint(ent)})*v[{l}]")
                else:
                    terms.append(f"hc[{len(prep_assign)}]*v[{l}]")
                    prep_assign.append((f"hc[{len(prep_assign)}]", ent))
                n_mul += 1
            if terms:
                apply_lines.append(f"            out[{i}] += {' + '.join(terms)};")
        code_p, fl_p = _emit_block(prep_assign, symmap_b, f"e{k}_", indent="            ", fast_rcp=True)
        blk_prep.append(f"        {kw_of(k)} constexpr (K == {k}) {{\n{code_p if code_p else '            (void)hc;'}\n        }}")
        blk_apply.append(f"        {kw_of(k)} constexpr (K == {k}) {{\n" + ("\n".join(apply_lines) if apply_lines else "            (void)out;") + "\n        }")
        blk_prep_flops.append(fl_p)
        blk_apply_flops.append(2 * n_mul)
        blk_nhc.append(len(prep_assign))
        # the stage term ACCUMULATES into out (it is added to the stage right-hand side); zero rows are skipped
        exprs_t = [(lhs.replace("out[", "acc["), e) for lhs, e in exprs_t if e != 0]
        code_t, fl_t = _emit_block(exprs_t, symmap_b, f"b{k}_", indent="            ", fast_rcp=True)
        code_t = code_t.replace("acc[", "out[").replace("] = ", "] += ") if code_t else "            (void)out;"
        fl_t += len(exprs_t)
        kw = "if" if k == 0 else "else if"
        blk_rhs.append(f"        {kw} constexpr (K == {k}) {{\n{code_r}\n        }}")
        blk_term.append(f"        {kw} constexpr (K == {k}) {{\n{code_t}\n        }}")
        blk_rhs_flops.append(fl_r)
        blk_term_flops.append(fl_t)
    blk_rhs_code = "\n".join(blk_rhs) if blk_rhs else "        (void)out;"
    blk_term_code = "\n".join(blk_term) if blk_term else "        (void)out;"
    blk_prep_code = "\n".join(blk_prep) if blk_prep else "        (void)hc;"
    blk_apply_code = "\n".join(blk_apply) if blk_apply else "        (void)out;"
    NHC = max(max(blk_nhc, default=0), 1)
    blk_prep_max = max(blk_prep_flops, default=0)
    blk_apply_max = max(blk_apply_flops, default=0)
    blk_rhs_max = max(blk_rhs_flops, default=0)
    blk_term_max = max(blk_term_flops, default=0)

    # ---- role-uniform block function for the block-per-lane DOPRI5 kernel: out = J(y) v + (df/dc)(y) w ---------------
    # every lane of a trajectory's group runs the SAME code on its own block vector v and its own dense coefficient
    # direction w (sensitivity block k: v = s_k, w = d c / d theta_k); dc_dense builds w for a run-time k
    wsyms = [sp.Symbol(f"w_{m}") for m in range(NC)]
    symmap_u = dict(symmap)
    for j in range(NX):
        symmap_u[vsyms[j]] = f"v[{j}]"
    for m in range(NC):
        symmap_u[wsyms[m]] = f"w[{m}]"
    uni_exprs = []
    for i in range(NX):
        e = sp.Integer(0)
        for j in range(NX):
            if (i, j) in hs.jac:
                e += hs.jac[(i, j)] * vsyms[j]
        for m in range(NC):
            if (i, m) in hs.dfdc:
                e += hs.dfdc[(i, m)] * wsyms[m]
        uni_exprs.append((f"out[{i}]", e))
    uni_code, uni_flops = _emit_block(uni_exprs, symmap_u, "u", fast_rcp=True)
    # f(y) = (df/dc)(y) c exactly when every f_i is linear and homogeneous in the hoisted coefficients
    zero_c = {cm: 0 for cm in hs.c_syms}
    lin_in_c = all(sp.simplify(fi.xreplace(zero_c)) == 0 and
                   all(sp.diff(fi, cm, cn) == 0 for cm in hs.c_syms for cn in hs.c_syms if cm in fi.free_symbols and cn in fi.free_symbols)
                   for fi in hs.f)
    dense_lines = [f"    w[{m}] = 0.0;" for m in range(NC)]
    for (k_, m_), n_ in sorted(dc_index.items(), key=lambda kv: kv[1]):
        dense_lines.append(f"    w[{m_}] = (k == {k_}) ? dc[{n_}] : w[{m_}];")

    unused = "(void)t; (void)y; (void)c;"
    NSd = max(NS, 1)
    header = f"""// GENERATED by parameter-variability_b200/codegen (emit.py) from SBML model '{name}'. Do not edit.
// states  : {', '.join(ode.states)}
// theta   : {', '.join(ode.theta)}
// sens    : {', '.join(hs.sens) if hs.sens else '(none)'}
// flops   : rhs {rhs_flops}, rhs+sens {rs_flops}, jac {jac_flops}, lu {lu_flops}, solve {sol_flops}, hoist {hoist_flops}
#pragma once
#include "../pv_common.cuh"

struct {struct_name} {{
    static constexpr int NX = {NX};     // integrated states
    static constexpr int NTH = {NTH};   // per-trajectory constants (theta)
    static constexpr int NC = {NC};     // hoisted coefficients
    static constexpr int NS = {NS};     // sensitivity parameters
    static constexpr int NSD = {NSd};   // max(NS,1) for array extents
    static constexpr int NDC = {NDC};   // non-zero d c / d theta entries
    static constexpr int NNZ = {NNZ};   // structural non-zeros of the Jacobian
    static constexpr int NNZ_CONST = {n_const};  // of which state independent
    static constexpr int NLU = {NLU};   // non-zeros of the filled LU
    static constexpr bool USES_TIME = {'true' if ode.uses_time else 'false'};
    static constexpr bool OBS_UNIT = {'true' if all(e == 1 for e in hs.obs_scale_exprs) else 'false'};  // every "[sid]" is the state itself
    static constexpr int FLOPS_RHS = {rhs_flops};
    static constexpr int FLOPS_RHS_SENS = {rs_flops};
    static constexpr int FLOPS_JAC = {jac_flops};
    static constexpr int FLOPS_LU = {lu_flops};
    static constexpr int FLOPS_SOLVE = {sol_flops};
    static constexpr int FLOPS_HOIST = {hoist_flops};
    static constexpr int FLOPS_SENS_TERM = {st_flops};
    static constexpr int FLOPS_DFDP = {fp_flops};
    static constexpr int FLOPS_JV = {jv_flops};
    static constexpr int FLOPS_JV_DFDC = {uni_flops};
    static constexpr bool RHS_LINEAR_IN_C = {'true' if lin_in_c else 'false'};  // rhs(y) == (df/dc)(y) c

    PV_HD static void hoist(const double (&th)[NTH], double (&c)[NC], double (&y0)[NX], double (&obs)[NX]) {{
{hoist_code}
    }}

    PV_HD static void hoist_sens(const double (&th)[NTH], double (&dc)[NDC], double (&dy0)[NSD][NX]) {{
        (void)th; (void)dc; (void)dy0;
{hs_code}
    }}

    template <class CV>
    PV_HD static void rhs(double t, const double (&y)[NX], const CV& c, double (&f)[NX]) {{
        {unused}
{rhs_code}
    }}

    PV_HD static void rhs_sens(double t, const double (&y)[NX], const double (&s)[NSD][NX],
                               const double (&c)[NC], const double (&dc)[NDC],
                               double (&f)[NX], double (&fs)[NSD][NX]) {{
        {unused} (void)s; (void)dc; (void)fs;
{rs_code}
    }}

    // sparse Jacobian values, pattern (row,col): {' '.join(f'({i},{j})' for i, j in pattern)}
    template <class CV>
    PV_HD static void jac(double t, const double (&y)[NX], const CV& c, double (&J)[NNZ]) {{
        {unused} (void)J;
{jac_code}
    }}

    // out = J v on the sparse pattern
    PV_HD static void jac_vec(const double (&J)[NNZ], const double (&v)[NX], double (&out)[NX]) {{
        (void)J; (void)v;
{chr(10).join(jv_lines)}
    }}

    // W = d*I - J in the filled, permuted LU layout (elimination order: {lu.perm})
    PV_HD static void lu_assemble(double d, const double (&J)[NNZ], double (&W)[NLU]) {{
        (void)J;
{chr(10).join(asm)}
    }}

    // in-place LU without pivoting; U's diagonal is stored inverted
    PV_HD static void lu_factor(double (&W)[NLU]) {{
{chr(10).join(fac)}
    }}

    // b <- W^-1 b   (W: any indexable container of the NLU factor entries)
    template <class WV>
    PV_HD static void lu_solve(const WV& W, double (&b)[NX]) {{
{chr(10).join(sol)}
    }}

    // explicit parameter derivative  fp[k] = (df/dc) dc_k
    PV_HD static void dfdp(double t, const double (&y)[NX], const double (&c)[NC], const double (&dc)[NDC],
                           double (&fp)[NSD][NX]) {{
        {unused} (void)dc; (void)fp;
{fp_code}
    }}

    // out[k] = (d/dy [J s_k + f_theta_k]) v   (second-order term of the sensitivity Jacobian)
    PV_HD static void sens_stage_term(double t, const double (&y)[NX], const double (&s)[NSD][NX],
                                      const double (&v)[NX], const double (&c)[NC], const double (&dc)[NDC],
                                      double (&out)[NSD][NX]) {{
        {unused} (void)s; (void)v; (void)dc; (void)out;
{st_code}
    }}

    // ---- role-uniform block right-hand side (block-per-lane DOPRI5 kernel): every lane of a trajectory's group runs this
    // one function on its own block vector v and its own coefficient direction w: out = J(y) v + (df/dc)(y) w.
    // Sensitivity block k: v = s_k, w = d c / d theta_k (dc_dense); the state lane selects rhs(y) instead. ----
    PV_HD static void jv_plus_dfdc(double t, const double (&y)[NX], const double (&v)[NX], const double (&c)[NC],
                                   const double (&w)[NC], double (&out)[NX]) {{
        {unused} (void)v; (void)w; (void)out;
{uni_code}
    }}
    // w = d c / d theta_k as a dense [NC] vector for a run-time k (zeros where c_m does not depend on theta_k)
    PV_HD static void dc_dense(int k, const double (&dc)[NDC], double (&w)[NC]) {{
        (void)k; (void)dc;
{chr(10).join(dense_lines)}
    }}

    // ---- one sensitivity block at a time (block-per-warp Rosenbrock kernel): s, v, out are the block's own vectors;
    // c and dc may live in registers or in shared memory (any indexable container) ----
    static constexpr int FLOPS_SENS_RHS_BLOCK = {blk_rhs_max};    // max over K
    static constexpr int FLOPS_SENS_TERM_BLOCK = {blk_term_max};  // max over K
    // out = J(y) s + (df/dc) dc_K
    template <int K, class YV, class CV, class DV>
    PV_HD static void sens_rhs_block(double t, const YV& y, const double (&s)[NX], const CV& c, const DV& dc,
                                     double (&out)[NX]) {{
        {unused} (void)s; (void)dc; (void)out;
{blk_rhs_code}
    }}
    // out += (d/dy [J s + f_theta_K]) v
    template <int K, class YV, class VV, class CV, class DV>
    PV_HD static void sens_stage_term_block(double t, const YV& y, const double (&s)[NX], const VV& v,
                                            const CV& c, const DV& dc, double (&out)[NX]) {{
        {unused} (void)s; (void)v; (void)dc; (void)out;
{blk_term_code}
    }}
    // the same term in two parts: the entries of H_K = d/dy [J s + f_theta_K] that cost more than a coefficient read,
    // evaluated once per step (y, s at the start of the step) ...
    static constexpr int NHC = {NHC};
    static constexpr int FLOPS_SENS_TERM_PREP = {blk_prep_max};    // max over K
    static constexpr int FLOPS_SENS_TERM_APPLY = {blk_apply_max};  // max over K
    template <int K, class YV, class CV, class DV>
    PV_HD static void sens_term_prep(double t, const YV& y, const double (&s)[NX], const CV& c, const DV& dc,
                                     double (&hc)[NHC]) {{
        {unused} (void)s; (void)dc; (void)hc;
{blk_prep_code}
    }}
    // ... and the sparse product out += H_K v at every stage
    template <int K, class VV, class CV, class DV>
    PV_HD static void sens_term_apply(const double (&hc)[NHC], const VV& v, const CV& c, const DV& dc, double (&out)[NX]) {{
        (void)hc; (void)v; (void)c; (void)dc; (void)out;
{blk_apply_code}
    }}
}};
"""
    meta = {
        "name": name,
        "struct": struct_name,
        "states": ode.states,
        "theta": ode.theta,
        "theta_default": ode.theta_default,
        "sens": hs.sens,
        "sens_index": [ode.theta.index(s) for s in hs.sens],
        "NX": NX, "NTH": NTH, "NC": NC, "NS": NS, "NDC": NDC, "NNZ": NNZ, "NLU": NLU,
        "jac_pattern": [list(ij) for ij in pattern],
        "lu_perm": lu.perm,
        "uses_time": ode.uses_time,
        "dead_rules": ode.dead_rules,
        "flops": {"rhs": rhs_flops, "rhs_sens": rs_flops, "jac": jac_flops, "lu": lu_flops,
                  "solve": sol_flops, "hoist": hoist_flops, "hoist_sens": hs_flops,
                  "sens_stage_term": st_flops, "dfdp": fp_flops, "jac_vec": jv_flops, "jv_plus_dfdc": uni_flops},
    }
    return EmittedModel(name=name, struct_name=struct_name, header=header, meta=meta)


# ----------------------------------------------------------------------------------------
# registry tables for the C ABI (generated/model_tables.inc)
# ----------------------------------------------------------------------------------------
def spectral_radius_at_defaults(hs: HoistedSystem) -> float:
    """max |eig J| at theta defaults, y = 1e-3 (a scale typical of the models' concentrations).
    Used only to choose the default integrator (explicit vs Rosenbrock) per model."""
    import numpy as np
    ode = hs.ode
    subs = {sp.Symbol(k): v for k, v in zip(ode.theta, ode.theta_default)}
    cvals = {c: sp.Float(float(e.xreplace(subs))) for c, e in zip(hs.c_syms, hs.c_exprs_theta)}
    yvals = {y: sp.Float(1e-3) for y in ode.state_syms}
    yvals[TIME_SYMBOL] = sp.Float(0.0)
    n = len(ode.states)
    J = np.zeros((n, n))
    for (i, j), e in hs.jac.items():
        J[i, j] = float(e.xreplace(cvals).xreplace(yvals))
    return float(np.max(np.abs(np.linalg.eigvals(J)))) if n else 0.0


STIFF_SPECTRAL_RADIUS = 50.0


def emit_tables(models: Sequence[EmittedModel]) -> str:
    def carr(name, ctype, items):
        body = ", ".join(items) if items else "0"
        return f"static const {ctype} {name}[] = {{{body}}};"

    out = ["// GENERATED by parameter-variability_b200/codegen (emit.py). Do not edit.",
           "#pragma once",
           "struct pv_model_table {",
           "    const char* name; int n_states, n_theta, n_sens;",
           "    const char* const* state_ids; const char* const* theta_ids; const double* theta_default;",
           "    const int* sens_index; int stiff; int n_flops; const char* const* flop_names; const int* flop_values;",
           "};"]
    flop_names = ["rhs", "rhs_sens", "jac", "lu", "solve", "hoist", "step_dopri5", "step_dopri5_sens",
                  "step_rodas4", "step_rodas4_sens", "dense_dopri5", "dense_dopri5_sens", "dense_rodas4",
                  "dense_rodas4_sens"]
    out.append(carr("pv_flop_names", "char* const", [f'"{n}"' for n in flop_names]))
    for em in models:
        m = em.meta
        s = em.struct_name
        out.append(carr(f"{s}_states", "char* const", [f'"{x}"' for x in m["states"]]))
        out.append(carr(f"{s}_theta", "char* const", [f'"{x}"' for x in m["theta"]]))
        out.append(carr(f"{s}_theta_default", "double", [repr(float(x)) for x in m["theta_default"]]))
        out.append(carr(f"{s}_sens_index", "int", [str(x) for x in m["sens_index"]]))
        fl = [f"{s}::FLOPS_RHS", f"{s}::FLOPS_RHS_SENS", f"{s}::FLOPS_JAC", f"{s}::FLOPS_LU", f"{s}::FLOPS_SOLVE",
              f"{s}::FLOPS_HOIST",
              f"pv_dopri5_cost<pv_system<{s}, false>>::FLOPS_STEP", f"pv_dopri5_cost<pv_system<{s}, true>>::FLOPS_STEP",
              f"pv_rodas4_cost<{s}, false>::FLOPS_STEP", f"pv_rodas4_cost<{s}, true>::FLOPS_STEP",
              f"pv_dopri5_cost<pv_system<{s}, false>>::FLOPS_DENSE", f"pv_dopri5_cost<pv_system<{s}, true>>::FLOPS_DENSE",
              f"pv_rodas4_cost<{s}, false>::FLOPS_DENSE", f"pv_rodas4_cost<{s}, true>::FLOPS_DENSE"]
        out.append(carr(f"{s}_flops", "int", fl))
    out.append(f"#define PV_N_MODELS {len(models)}")
    out.append("static const pv_model_table pv_model_tables[PV_N_MODELS] = {")
    for em in models:
        m = em.meta
        s = em.struct_name
        out.append(f'    {{"{m["name"]}", {m["NX"]}, {m["NTH"]}, {m["NS"]}, {s}_states, {s}_theta, {s}_theta_default, '
                   f'{s}_sens_index, {1 if m["stiff"] else 0}, {len(flop_names)}, pv_flop_names, {s}_flops}},')
    out.append("};")
    out.append("// switch body: `case i: CALL<Model_i> ARGS; break;`")
    cases = " ".join(f"case {i}: CALL<{em.struct_name}> ARGS; break;" for i, em in enumerate(models))
    out.append(f"#define PV_MODEL_DISPATCH(CALL, ARGS) {cases}")
    return "\n".join(out) + "\n"
