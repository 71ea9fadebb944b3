#!/usr/bin/env python3
"""Lower the reference model's expression graph to CUDA device functions.

Runs `export_robot_ode_model()` of /root/reference/src/models/robot_model.py:8-67 UNMODIFIED
against the SymPy-backed shims in codegen/shims/ (CasADi / acados_template are not installable
offline) and emits `csrc/model_generated.cuh`:

  model_gl4_step   one IRK Gauss-Legendre-4 step x+ = F(x,u;h) plus the state-dependent entries
                   of the sensitivities dF/dx, dF/du.  The collocation system K_i = f(x + h sum_j
                   a_ij K_j, u) is solved SYMBOLICALLY by forward substitution over the model's
                   dependency DAG (v, omega -> psi -> x, y), which is what acados' Newton iteration
                   converges to (proven in tests/test_oracle_model.py).
  model_rk4_step   classical RK4 with sensitivities (north_star's named alternative integrator).
  model_con_h      obstacle constraint h_j and its gradient, robot_model.py:60-65.

The generated header is committed (the GPU box has no /root/reference); re-run this script
whenever the reference model changes:   python codegen/gen_model.py [--reference /root/reference]
"""
import argparse
import os
import sys

import mpmath
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_model(ref_root):
    sys.path.insert(0, os.path.join(HERE, "shims"))
    src = os.path.join(ref_root, "src")
    sys.path.insert(0, src)
    sys.path.insert(0, os.path.join(src, "models"))
    from models.robot_model import export_robot_ode_model   # noqa: E402  (reference code, unmodified)
    from models import world_specification as ws            # noqa: E402
    return export_robot_ode_model(), ws


def gauss_legendre(s, dps=40):
    """(A, b, c) of the s-stage Gauss-Legendre collocation method, to `dps` digits."""
    mpmath.mp.dps = dps
    t = sp.Symbol("t")
    P = sp.legendre(s, 2 * t - 1)
    c = sorted(mpmath.mpf(sp.N(r, dps)) for r in sp.Poly(P, t).nroots(n=dps))
    A = [[None] * s for _ in range(s)]
    b = [None] * s
    for j in range(s):
        num = [mpmath.mpf(1)]
        den = mpmath.mpf(1)
        for m in range(s):
            if m != j:
                num = [(num[i - 1] if i > 0 else 0) - c[m] * (num[i] if i < len(num) else 0) for i in range(len(num) + 1)]
                den *= (c[j] - c[m])
        # num holds coefficients of prod (t - c_m), lowest degree first
        integ = lambda x: sum(co / (i + 1) * x ** (i + 1) for i, co in enumerate(num)) / den  # noqa: E731
        for i in range(s):
            A[i][j] = integ(c[i])
        b[j] = integ(mpmath.mpf(1))
    return A, b, c


def to_rat(v):
    """mpmath -> sympy Float with enough digits that printing 17 sig. digits is exact"""
    return sp.Float(mpmath.nstr(v, 30), 30)


def topo_levels(f, x):
    n = len(x)
    deps = [{q for q in range(n) if f[r].has(x[q])} for r in range(n)]
    done, order = set(), []
    while len(done) < n:
        ready = [r for r in range(n) if r not in done and deps[r] <= done]
        if not ready:
            raise SystemExit("model ODE is not triangular: symbolic collocation solve impossible "
                             "(a generic on-device IRK-Newton would be needed)")
        order.extend(ready)
        done.update(ready)
    return order


def symbolic_gl4(f, x, u, h, stages=4):
    A, b, c = gauss_legendre(stages)
    n = len(x)
    order = topo_levels(f, x)
    K = [[None] * n for _ in range(stages)]
    for r in order:
        for i in range(stages):
            sub = {}
            for q in range(n):
                if f[r].has(x[q]):
                    sub[x[q]] = x[q] + h * sum(to_rat(A[i][j]) * K[j][q] for j in range(stages))
            K[i][r] = sp.expand(f[r].xreplace(sub)) if not any(isinstance(a, (sp.sin, sp.cos)) for a in sp.preorder_traversal(f[r])) \
                else f[r].xreplace({k: sp.collect(sp.expand(v), [h]) for k, v in sub.items()})
    xn = [x[r] + h * sum(to_rat(b[i]) * K[i][r] for i in range(stages)) for r in range(n)]
    verify_collocation(f, x, u, h, A, K, stages)
    return xn


def verify_collocation(f, x, u, h, A, K, stages):
    """Generator-time assertion: the stage derivatives obtained by forward substitution must satisfy the
    collocation system K_i = f(x + h sum_j a_ij K_j, u) that acados' IRK solves with Newton -- checked numerically at
    random points with 30 digits.  A model whose ODE is not triangular never gets here (topo_levels exits); a wrong
    substitution would show up as a residual."""
    import random
    rnd = random.Random(0)
    n = len(x)
    for _ in range(3):
        val = {s_: sp.Float(rnd.uniform(-3, 3), 30) for s_ in list(x) + list(u)}
        val[h] = sp.Float(rnd.choice([0.05, 0.1, 0.2]), 30)
        Kn = [[sp.N(K[i][r].xreplace(val), 30) for r in range(n)] for i in range(stages)]
        for i in range(stages):
            xs = {x[q]: val[x[q]] + val[h] * sum(to_rat(A[i][j]) * Kn[j][q] for j in range(stages)) for q in range(n)}
            xs.update({uu: val[uu] for uu in u})
            for r in range(n):
                res = abs(sp.N(f[r].xreplace(xs), 30) - Kn[i][r])
                if res > 1e-20:
                    raise SystemExit(f"closed-form collocation solve does not satisfy the IRK system (stage {i}, state {r}: residual {res})")


def symbolic_rk4(f, x, u, h):
    n = len(x)
    def F(xs):
        return [fi.xreplace(dict(zip(x, xs))) for fi in f]
    k1 = F(list(x))
    k2 = F([x[i] + h / 2 * k1[i] for i in range(n)])
    k3 = F([x[i] + h / 2 * k2[i] for i in range(n)])
    k4 = F([x[i] + h * k3[i] for i in range(n)])
    return [x[i] + h / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]) for i in range(n)]


class Emitter:
    def __init__(self):
        self.lines = []

    def emit_function(self, name, doc, in_decl, xn, S_entries, x, u, h, extra_outputs=()):
        """xn: list of exprs; S_entries: list of (label, expr) for the variable sensitivities."""
        exprs = list(xn) + [e for _, e in S_entries]
        # pull every distinct trig argument into one sincos() call
        args = []
        for e in exprs:
            for node in sp.preorder_traversal(e):
                if isinstance(node, (sp.sin, sp.cos)) and node.args[0] not in args:
                    args.append(node.args[0])
        trig_sub = {}
        pre = []
        for k, a in enumerate(args):
            sn, cs, ang = sp.Symbol(f"sn{k}"), sp.Symbol(f"cs{k}"), sp.Symbol(f"ang{k}")
            trig_sub[sp.sin(a)] = sn
            trig_sub[sp.cos(a)] = cs
            pre.append((ang, a, sn, cs))
        exprs = [e.xreplace(trig_sub) for e in exprs]
        ang_exprs = [a for _, a, _, _ in pre]
        repl, red = sp.cse(ang_exprs + exprs, symbols=sp.numbered_symbols("t"), optimizations="basic")
        red_ang, red_out = red[:len(ang_exprs)], red[len(ang_exprs):]
        L = self.lines
        L.append(f"/* {doc} */")
        L.append(f"__host__ __device__ __forceinline__ void {name}({in_decl})")
        L.append("{")
        # temporaries that the angles need must come first; emit all temporaries that do not
        # depend on sn/cs, then sincos, then the rest
        trig_syms = {s for _, _, sn, cs in pre for s in (sn, cs)}
        early, late = [], []
        tainted = set(trig_syms)
        for sym, ex in repl:
            if ex.free_symbols & tainted:
                tainted.add(sym)
                late.append((sym, ex))
            else:
                early.append((sym, ex))
        for sym, ex in early:
            L.append(f"    const double {sym} = {ccode(ex)};")
        for (ang, _, sn, cs), ex in zip(pre, red_ang):
            L.append(f"    double {sn}, {cs};")
            L.append(f"    MODEL_SINCOS({ccode(ex)}, &{sn}, &{cs});")
        for sym, ex in late:
            L.append(f"    const double {sym} = {ccode(ex)};")
        nx = len(xn)
        for i in range(nx):
            L.append(f"    xn[{i}] = {ccode(red_out[i])};")
        for k, (label, _) in enumerate(S_entries):
            L.append(f"    sv[{k}] = {ccode(red_out[nx + k])};   /* {label} */")
        L.append("}")
        L.append("")


from sympy.printing.c import C99CodePrinter


class _Printer(C99CodePrinter):
    """integer powers become products (no pow() calls in device code)"""

    def _print_Pow(self, expr):
        if expr.exp.is_Integer and 2 <= int(expr.exp) <= 4:
            return "model_ipow%d(%s)" % (int(expr.exp), self._print(expr.base))
        return super()._print_Pow(expr)


def ccode(e):
    s = _Printer().doprint(e).replace("\n", " ")
    assert "pow(" not in s, s
    return s


def classify(S, xu, h):
    """pattern chars: '0' zero, '1' one, 'c' constant in (x,u) (depends on h only), 'v' variable"""
    pat = []
    for r in range(S.rows):
        row = ""
        for cidx in range(S.cols):
            e = sp.simplify(S[r, cidx])
            if e == 0:
                row += "0"
            elif e == 1:
                row += "1"
            elif not (e.free_symbols & set(xu)):
                row += "c"
            else:
                row += "v"
        pat.append(row)
    return pat


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "..", "csrc", "model_generated.cuh"))
    a = ap.parse_args()
    model, ws = load_reference_model(a.reference)
    x = model.x.elems
    u = model.u.elems
    p = model.p.elems
    f = model.f_expl_expr.elems
    hsym = sp.Symbol("h", real=True, positive=True)
    nx, nu = len(x), len(u)
    assert (nx, nu) == (5, 2)
    # f_impl must be xdot - f_expl (robot_model.py:45)
    assert all(sp.simplify(fi - (xd - fe)) == 0 for fi, xd, fe in zip(model.f_impl_expr.elems, model.xdot.elems, f))

    xs = [sp.Symbol(f"x[{i}]", real=True) for i in range(nx)]
    us = [sp.Symbol(f"u[{i}]", real=True) for i in range(nu)]
    ren = dict(zip(x + u, xs + us))
    f_r = [fi.xreplace(ren) for fi in f]

    em = Emitter()
    L = em.lines
    L.append("// AUTO-GENERATED by codegen/gen_model.py from the expression graph of")
    L.append("// src/models/robot_model.py:8-67 (reference model, executed unmodified against a SymPy shim).")
    L.append("// DO NOT EDIT BY HAND -- regenerate with `python codegen/gen_model.py`.")
    L.append("#pragma once")
    L.append("#include <math.h>")
    L.append("/* sin/cos of the heading: csrc/b2_math.cuh defines the same operation sequence for the GPU and the CPU mirror */")
    L.append('#include "b2_math.cuh"')
    L.append("#ifndef MODEL_SINCOS")
    L.append("#define MODEL_SINCOS(x, s, c) b2_sincos((x), (s), (c))")
    L.append("#endif")
    L.append(f"#define MODEL_NX {nx}")
    L.append(f"#define MODEL_NU {nu}")
    L.append(f"#define MODEL_N_OBST_DEFAULT {len(p) // 2}   /* world_specification.py:25 */")
    L.append("__host__ __device__ __forceinline__ double model_ipow2(double a) { return a * a; }")
    L.append("__host__ __device__ __forceinline__ double model_ipow3(double a) { return a * a * a; }")
    L.append("__host__ __device__ __forceinline__ double model_ipow4(double a) { double b = a * a; return b * b; }")
    L.append("")

    patterns = {}
    for tag, xn in (("gl4", symbolic_gl4(f_r, xs, us, hsym)), ("rk4", symbolic_rk4(f_r, xs, us, hsym))):
        S = sp.Matrix(xn).jacobian(sp.Matrix(xs + us))
        pat = classify(S, xs + us, hsym)
        patterns[tag] = pat
        entries = []
        consts = []
        for r in range(nx):
            for cidx in range(nx + nu):
                name = ("A" if cidx < nx else "B") + f"[{r}][{cidx if cidx < nx else cidx - nx}]"
                if pat[r][cidx] == "v":
                    entries.append((name, S[r, cidx]))
                elif pat[r][cidx] == "c":
                    consts.append((name, sp.simplify(S[r, cidx])))
        L.append(f"#define MODEL_{tag.upper()}_NSV {len(entries)}")
        L.append(f"/* sensitivity pattern of {tag} (rows x+, cols [x;u]); 0 zero, 1 one, c const in h, v variable:")
        for row in pat:
            L.append(f"     {row}")
        L.append("   const entries: " + ", ".join(f"{n} = {ccode(e)}" for n, e in consts) + " */")
        L.append(f"#define MODEL_{tag.upper()}_PATTERN \"" + "/".join(pat) + "\"")
        doc = {"gl4": "one IRK Gauss-Legendre-4 step (closed form of the collocation system) + variable sensitivities",
               "rk4": "one classical RK4 step + variable sensitivities"}[tag]
        em.emit_function(f"model_{tag}_step", doc,
                         "const double* __restrict__ x, const double* __restrict__ u, const double h, "
                         "double* __restrict__ xn, double* __restrict__ sv", xn, entries, xs, us, hsym)

    # ---- obstacle constraint, robot_model.py:60-65 ------------------------------------
    hexpr = model.con_h_expr.elems
    assert [sp.simplify(a - b) for a, b in zip(hexpr, model.con_h_expr_e.elems)] == [0] * len(hexpr)
    px, py = sp.Symbol("px", real=True), sp.Symbol("py", real=True)
    h0 = hexpr[0].xreplace({p[0]: px, p[1]: py}).xreplace(ren)
    for j in range(len(hexpr)):   # every obstacle uses the same expression on its own (p_2j, p_2j+1)
        assert sp.simplify(hexpr[j].xreplace({p[2 * j]: px, p[2 * j + 1]: py}).xreplace(ren) - h0) == 0
    grad = [sp.diff(h0, s) for s in xs]
    assert all(g == 0 for g in grad[2:]), "constraint must depend on (x,y) only"
    r2 = -h0.xreplace({xs[0]: 0, xs[1]: 0, px: 0, py: 0})
    L.append(f"#define MODEL_R2 {ccode(sp.Float(r2, 17))}   /* (R_OBST+R_ROBOT+MARGIN)^2, robot_model.py:62 */")
    L.append("/* obstacle constraint h(x,p_j) and its gradient w.r.t. (x,y), robot_model.py:60-65 */")
    L.append("__host__ __device__ __forceinline__ void model_con_h(const double* __restrict__ x, const double px, const double py, "
             "const double r2, double* hval, double* cx, double* cy)")
    L.append("{")
    L.append(f"    *hval = {ccode(h0 + r2)} - r2;")
    L.append(f"    *cx = {ccode(grad[0])};")
    L.append(f"    *cy = {ccode(grad[1])};")
    L.append("}")
    L.append("")
    # world constants the host wrapper cross-checks
    L.append(f"#define MODEL_WORLD_N_OBST {ws.N_OBST}")
    L.append(f"#define MODEL_WORLD_C_MAX {float(ws.C_MAX)!r}")
    L.append(f"#define MODEL_WORLD_V_MAX_ROBOT {float(ws.V_MAX_ROBOT)!r}")
    with open(a.out, "w") as fh:
        fh.write("\n".join(L) + "\n")
    print("wrote", os.path.abspath(a.out))
    for tag, pat in patterns.items():
        print(tag, pat)


if __name__ == "__main__":
    main()
