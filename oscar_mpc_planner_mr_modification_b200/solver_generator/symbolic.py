"""sympy helpers of the CUDA emitter: C++ expression printer, CSE'd statement blocks, and access to
the sympy-backed `casadi` stand-in the module definitions are evaluated with."""
import importlib
import os
import re
import sys

import sympy as sp
from sympy.printing.c import C99CodePrinter

_COMPAT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "casadi_compat")


def symbolic_namespace():
    """The `casadi` module the problem definition scripts import (sympy-backed stand-in)."""
    if _COMPAT not in sys.path:
        sys.path.insert(0, _COMPAT)
    mod = importlib.import_module("casadi")
    if not hasattr(mod, "to_sympy"):
        raise RuntimeError("a real casadi is on the path; the CUDA emitter needs the sympy-backed stand-in "
                           "(put %s first on sys.path)" % _COMPAT)
    return mod


def to_exprs(v):
    return symbolic_namespace().to_sympy(v)


class _Printer(C99CodePrinter):
    def _print_Pow(self, expr):
        b, e = expr.args
        if e.is_Integer and 2 <= int(e) <= 4:
            s = self.parenthesize(b, 100)
            return "(" + "*".join([s] * int(e)) + ")"
        if e.is_Integer and -4 <= int(e) <= -1:
            s = self.parenthesize(b, 100)
            return "MPCGEN_RCP(" + "*".join([s] * (-int(e))) + ")"
        if e.is_Rational and e.q == 2:
            # half-integer power: x^(n/2) = x^(n//2) * sqrt(x)
            n = abs(int(e.p)) // 2
            s = self._print(b)
            body = "*".join(["(%s)" % s] * n + ["sqrt(%s)" % s])
            return "(%s)" % body if e.p > 0 else "MPCGEN_RCP(%s)" % body
        return super()._print_Pow(expr)

    def _print_Integer(self, expr):
        return "%d.0" % int(expr)

    def _print_Rational(self, expr):
        return "(%d.0/%d.0)" % (expr.p, expr.q)

    def _print_Float(self, expr):
        return repr(float(expr))

    def _print__fmod(self, expr):
        return "fmod(%s, %s)" % (self._print(expr.args[0]), self._print(expr.args[1]))


_printer = _Printer()


def ccode(expr):
    return _printer.doprint(sp.sympify(expr))


_TRIG = re.compile(r"^(\s*)const double (\w+) = (sin|cos)\((.*)\);$")


def _pair_sincos(lines, indent):
    """`const double a = cos(X);` and `const double b = sin(X);` of the same argument share one range reduction:
    emitted as `double a, b; sincos(X, &b, &a);` at the first of the two."""
    first = {}
    for i, ln in enumerate(lines):
        m_ = _TRIG.match(ln)
        if m_:
            first.setdefault(m_.group(4), {})[m_.group(3)] = (i, m_.group(2))
    drop = set()
    for arg, d in first.items():
        if "sin" in d and "cos" in d:
            (i_s, n_s), (i_c, n_c) = d["sin"], d["cos"]
            at, other = min(i_s, i_c), max(i_s, i_c)
            lines[at] = "%sdouble %s, %s; sincos(%s, &%s, &%s);" % (indent, n_s, n_c, arg, n_s, n_c)
            drop.add(other)
    return [ln for i, ln in enumerate(lines) if i not in drop]


def emit_block(outputs, subs_map, tmp_prefix="t", indent="    "):
    """Statements for outputs = [(lhs, expr)] ; an lhs ending in '+' accumulates (lhs += expr)."""
    exprs = [sp.sympify(e) for _, e in outputs]
    ren = {s: sp.Symbol(c) for s, c in subs_map.items()}
    # sin(X) and cos(X) of the same argument share one range reduction: both are replaced by symbols filled by one
    # sincos(X, &s, &c) emitted ahead of the block (arguments are printed in full: they are short input expressions)
    sins = {a.args[0] for e in exprs for a in e.atoms(sp.sin)}
    coss = {a.args[0] for e in exprs for a in e.atoms(sp.cos)}
    lines, trig = [], {}
    for n_, arg in enumerate(sorted(sins & coss, key=sp.default_sort_key)):
        if arg.atoms(sp.sin, sp.cos):
            continue                                     # nested trigonometry: leave to the CSE
        sn, cn = sp.Symbol("%ssn%d" % (tmp_prefix, n_)), sp.Symbol("%scs%d" % (tmp_prefix, n_))
        trig[sp.sin(arg)], trig[sp.cos(arg)] = sn, cn
        lines.append("%sdouble %s, %s; sincos(%s, &%s, &%s);" % (indent, sn, cn, ccode(arg.xreplace(ren)), sn, cn))
    if trig:
        exprs = [e.xreplace(trig) for e in exprs]
    rep, red = sp.cse(exprs, symbols=sp.numbered_symbols(tmp_prefix), order="none")
    for s, e in rep:
        lines.append("%sconst double %s = %s;" % (indent, s, ccode(e.xreplace(ren))))
    lines = _pair_sincos(lines, indent)
    for (lhs, _), e in zip(outputs, red):
        op = "="
        if lhs.endswith("+"):
            lhs, op = lhs[:-1].strip(), "+="
        lines.append("%s%s %s %s;" % (indent, lhs, op, ccode(e.xreplace(ren))))
    return "\n".join(lines)
