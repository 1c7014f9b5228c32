"""sympy -> C statement helpers shared by the generators (generation-time tooling only)."""
import sympy as sp
from sympy.printing.c import C99CodePrinter


class _Printer(C99CodePrinter):
    """C99 printer: small integer powers as products, every literal a double."""

    def _print_Pow(self, expr):
        b, e = expr.args
        if e.is_Integer and 2 <= int(e) <= 4:
            s = self.parenthesize(b, 100)
            return "(" + "*".join([s] * int(e)) + ")"
        if e.is_Integer and -4 <= int(e) <= -1:
            s = self.parenthesize(b, 100)
            return "(1.0/(" + "*".join([s] * (-int(e))) + "))"
        if e == sp.Rational(1, 2):
            return "sqrt(%s)" % self._print(b)
        if e == sp.Rational(-1, 2):
            return "(1.0/sqrt(%s))" % self._print(b)
        if e == sp.Rational(3, 2):
            s = self._print(b)
            return "((%s)*sqrt(%s))" % (s, s)
        if e == sp.Rational(-3, 2):
            s = self._print(b)
            return "(1.0/((%s)*sqrt(%s)))" % (s, s)
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


def emit_block(outputs, subs_map, tmp_prefix="t", indent="    ", decl="const double"):
    """C statements assigning `outputs` = [(lhs_string, sympy_expr), ...].

    subs_map maps model symbols to C lvalue strings (e.g. x_0 -> "z[2]").
    Common subexpressions are hoisted into `const double tN` temporaries.
    """
    exprs = [sp.sympify(e) for _, e in outputs]
    tmp_syms = sp.numbered_symbols(tmp_prefix)
    rep, red = sp.cse(exprs, symbols=tmp_syms, order="none")
    ren = {s: sp.Symbol(c) for s, c in subs_map.items()}
    lines = []
    for s, e in rep:
        lines.append("%s%s %s = %s;" % (indent, decl, s, ccode(e.xreplace(ren))))
    for (lhs, _), e in zip(outputs, red):
        lines.append("%s%s = %s;" % (indent, lhs, ccode(e.xreplace(ren))))
    return "\n".join(lines)


def count_ops(outputs):
    exprs = [sp.sympify(e) for _, e in outputs]
    rep, red = sp.cse(exprs, order="none")
    return sum(sp.count_ops(e) for _, e in rep) + sum(sp.count_ops(e) for e in red)
