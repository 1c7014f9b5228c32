"""sympy helpers of the CUDA emitter: C++ expression printer, CSE'd statement blocks, and access to
the sympy-backed `casadi` stand-in the module definitions are evaluated with."""
import importlib
import os
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
            return "(1.0/(" + "*".join([s] * (-int(e))) + "))"
        if e.is_Rational and e.q == 2:
            # half-integer power: x^(n/2) = x^(n//2) * sqrt(x)
            n = abs(int(e.p)) // 2
            s = self._print(b)
            body = "*".join(["(%s)" % s] * n + ["sqrt(%s)" % s])
            return "(%s)" % body if e.p > 0 else "(1.0/(%s))" % body
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


def emit_block(outputs, subs_map, tmp_prefix="t", indent="    "):
    """Statements for outputs = [(lhs, expr)] ; an lhs ending in '+' accumulates (lhs += expr)."""
    exprs = [sp.sympify(e) for _, e in outputs]
    rep, red = sp.cse(exprs, symbols=sp.numbered_symbols(tmp_prefix), order="none")
    ren = {s: sp.Symbol(c) for s, c in subs_map.items()}
    lines = []
    for s, e in rep:
        lines.append("%sconst double %s = %s;" % (indent, s, ccode(e.xreplace(ren))))
    for (lhs, _), e in zip(outputs, red):
        op = "="
        if lhs.endswith("+"):
            lhs, op = lhs[:-1].strip(), "+="
        lines.append("%s%s %s %s;" % (indent, lhs, op, ccode(e.xreplace(ren))))
    return "\n".join(lines)
