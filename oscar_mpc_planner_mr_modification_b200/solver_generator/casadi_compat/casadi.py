"""Minimal `casadi` stand-in backed by sympy.

casadi is not installed in this image (and there is no network), but the reference's problem
definition scripts (`solver_generator/*.py`, `mpc_planner_modules/scripts/*.py`) do
`import casadi as cd`.  Putting this directory on `sys.path` lets those scripts be imported
UNCHANGED so that their symbolic expressions (dynamics, stage cost, constraints) can be used as the
source of truth by the generators in this repository.  Only what the reference touches is provided:
`SX.sym`, `SX(r, c)`, `SX(ndarray)`, `vertcat`, elementary functions, `.T`, `@`, indexing.

This is generation-time tooling only; nothing here runs on the solve path.
"""
import numpy as _np
import sympy as _sp

pi = _np.pi
inf = _np.inf


def _raw(v):
    """sympy expression of a scalar-like value."""
    if isinstance(v, SXElem):
        return v.e
    if isinstance(v, SX):
        if v.m.shape != (1, 1):
            raise ValueError("expected a scalar SX, got shape %s" % (v.m.shape,))
        return v.m[0, 0]
    if isinstance(v, _np.ndarray) and v.shape == ():
        return _raw(v.item())
    if isinstance(v, (int, _np.integer)):
        return _sp.Integer(int(v))
    if isinstance(v, (float, _np.floating)):
        return _sp.Float(float(v))
    if isinstance(v, _sp.Basic):
        return v
    raise TypeError("cannot convert %r to a symbolic scalar" % (type(v),))


class SXElem:
    """Scalar symbolic value. Deliberately NOT a sequence so numpy stores it as one object."""

    __slots__ = ("e",)

    def __init__(self, e):
        self.e = _raw(e)

    # arithmetic
    def __add__(self, o): return _bin(self, o, lambda a, b: a + b)
    def __radd__(self, o): return _bin(o, self, lambda a, b: a + b)
    def __sub__(self, o): return _bin(self, o, lambda a, b: a - b)
    def __rsub__(self, o): return _bin(o, self, lambda a, b: a - b)
    def __mul__(self, o): return _bin(self, o, lambda a, b: a * b)
    def __rmul__(self, o): return _bin(o, self, lambda a, b: a * b)
    def __truediv__(self, o): return _bin(self, o, lambda a, b: a / b)
    def __rtruediv__(self, o): return _bin(o, self, lambda a, b: a / b)
    def __pow__(self, o): return _bin(self, o, lambda a, b: a ** b)
    def __rpow__(self, o): return _bin(o, self, lambda a, b: a ** b)
    def __neg__(self): return SXElem(-self.e)
    def __pos__(self): return self
    # casadi SX is 1x1-matrix like: `.T`, `@` and shape on scalars are legal
    @property
    def T(self): return self
    @property
    def shape(self): return (1, 1)
    def __matmul__(self, o): return self * o
    def __rmatmul__(self, o): return o * self

    # comparisons against plain numbers (used by `assert obj > 0` style checks after substitution)
    def __float__(self): return float(self.e)
    def __gt__(self, o): return float(self) > float(o)
    def __lt__(self, o): return float(self) < float(o)
    def __ge__(self, o): return float(self) >= float(o)
    def __le__(self, o): return float(self) <= float(o)
    def __eq__(self, o):
        try:
            return bool(_sp.simplify(self.e - _raw(o)) == 0)
        except TypeError:
            return False
    def __hash__(self): return hash(self.e)

    # numpy's object-dtype ufunc loops call these methods (np.exp(x) -> x.exp())
    def exp(self): return SXElem(_sp.exp(self.e))
    def sqrt(self): return SXElem(_sp.sqrt(self.e))
    def cos(self): return SXElem(_sp.cos(self.e))
    def sin(self): return SXElem(_sp.sin(self.e))
    def tan(self): return SXElem(_sp.tan(self.e))
    def log(self): return SXElem(_sp.log(self.e))
    def arctan(self): return SXElem(_sp.atan(self.e))

    def __repr__(self): return "SXElem(%s)" % (self.e,)


def _bin(a, b, f):
    if isinstance(a, (SX, _np.ndarray)) or isinstance(b, (SX, _np.ndarray)):
        return NotImplemented
    return SXElem(f(_raw(a), _raw(b)))


class SX:
    """Dense symbolic matrix."""

    __array_ufunc__ = None  # numpy binary operators defer to the reflected methods below

    def __init__(self, *args):
        if len(args) == 0:
            self.m = _sp.zeros(0, 1)
        elif len(args) == 2 and all(isinstance(a, (int, _np.integer)) for a in args):
            self.m = _sp.zeros(int(args[0]), int(args[1]))
        elif len(args) == 1:
            a = args[0]
            if isinstance(a, SX):
                self.m = a.m.copy()
            elif isinstance(a, _sp.MatrixBase):
                self.m = _sp.Matrix(a)
            elif isinstance(a, _np.ndarray):
                if a.ndim == 1:
                    self.m = _sp.Matrix([[_raw(v)] for v in a])
                elif a.ndim == 2:
                    self.m = _sp.Matrix([[_raw(v) for v in row] for row in a])
                else:
                    raise ValueError("SX from ndarray: ndim must be 1 or 2")
            elif isinstance(a, (list, tuple)):
                self.m = _sp.Matrix([[_raw(v)] for v in a])
            else:
                self.m = _sp.Matrix([[_raw(a)]])
        else:
            raise TypeError("unsupported SX constructor arguments")

    @staticmethod
    def sym(name, *dims):
        n = int(dims[0]) if len(dims) >= 1 else 1
        c = int(dims[1]) if len(dims) >= 2 else 1
        if n == 1 and c == 1:
            return SXElem(_sp.Symbol(name, real=True))
        if c == 1:
            return SX(_sp.Matrix([[_sp.Symbol("%s_%d" % (name, i), real=True)] for i in range(n)]))
        return SX(_sp.Matrix([[_sp.Symbol("%s_%d_%d" % (name, i, j), real=True) for j in range(c)] for i in range(n)]))

    @property
    def shape(self): return self.m.shape
    def size(self): return self.m.shape
    def size1(self): return self.m.shape[0]
    def size2(self): return self.m.shape[1]
    @property
    def T(self): return SX(self.m.T)
    def __len__(self): return self.m.shape[0]

    def _vec_index(self, key):
        r, c = self.m.shape
        if c == 1:
            return key, 0
        if r == 1:
            return 0, key
        raise IndexError("single index on a matrix")

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            if isinstance(key, slice):
                r, c = self.m.shape
                if c == 1:
                    rows = range(r)[key]
                    return SX(_sp.Matrix([[self.m[i, 0]] for i in rows])) if len(rows) else SX(0, 1)
                key = (0, key)
            else:
                if key < 0:
                    key += max(self.m.shape)
                key = self._vec_index(key)
        out = self.m[key]
        if isinstance(out, _sp.MatrixBase):
            return SX(out)
        return SXElem(out)

    def __setitem__(self, key, value):
        if not isinstance(key, tuple):
            key = self._vec_index(key)
        self.m[key] = _raw(value)

    def __iter__(self):
        if self.m.shape[1] != 1 and self.m.shape[0] != 1:
            raise TypeError("iteration over a matrix")
        return (SXElem(v) for v in self.m)

    @staticmethod
    def _coerce(o):
        if isinstance(o, SX):
            return o
        if isinstance(o, (_np.ndarray, list, tuple)):
            return SX(_np.asarray(o, dtype=object))
        return None

    def _ew(self, o, f, swap=False):
        om = SX._coerce(o)
        if om is None:
            s = _raw(o)
            return SX(self.m.applyfunc(lambda a: f(s, a) if swap else f(a, s)))
        if om.m.shape == (1, 1):
            return self._ew(om.m[0, 0], f, swap)
        if self.m.shape == (1, 1):
            return om._ew(self.m[0, 0], f, not swap)
        if om.m.shape != self.m.shape:
            raise ValueError("shape mismatch %s vs %s" % (self.m.shape, om.m.shape))
        r, c = self.m.shape
        return SX(_sp.Matrix(r, c, lambda i, j: f(om.m[i, j], self.m[i, j]) if swap else f(self.m[i, j], om.m[i, j])))

    def __add__(self, o): return self._ew(o, lambda a, b: a + b)
    def __radd__(self, o): return self._ew(o, lambda a, b: a + b, True)
    def __sub__(self, o): return self._ew(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._ew(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._ew(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._ew(o, lambda a, b: a * b, True)
    def __truediv__(self, o): return self._ew(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._ew(o, lambda a, b: a / b, True)
    def __pow__(self, o): return self._ew(o, lambda a, b: a ** b)
    def __neg__(self): return SX(-self.m)

    def __matmul__(self, o):
        om = SX._coerce(o)
        if om is None:
            return self * o
        return SX(self.m * om.m)

    def __rmatmul__(self, o):
        om = SX._coerce(o)
        if om is None:
            return self * o
        if om.m.shape[1] != self.m.shape[0] and om.m.shape[0] == self.m.shape[0] and om.m.shape[1] == 1:
            om = om.T  # numpy 1-D vector on the left acts as a row
        return SX(om.m * self.m)

    def __float__(self): return float(_raw(self))
    def __gt__(self, o): return float(self) > float(o)
    def __lt__(self, o): return float(self) < float(o)

    def __repr__(self): return "SX(%s)" % (self.m,)


def _flatten(args):
    rows = []
    for a in args:
        if isinstance(a, SX):
            if a.m.shape[0] == 0:
                continue
            if a.m.shape[1] != 1:
                raise ValueError("vertcat of non-column SX")
            rows.extend(list(a.m))
        elif isinstance(a, (_np.ndarray, list, tuple)):
            rows.extend(_raw(v) for v in _np.asarray(a, dtype=object).ravel())
        else:
            rows.append(_raw(a))
    return rows


def vertcat(*args):
    rows = _flatten(args)
    if not rows:
        return SX(0, 1)
    return SX(_sp.Matrix([[r] for r in rows]))


def _unary(f):
    def g(x):
        if isinstance(x, SX):
            return SX(x.m.applyfunc(f))
        if isinstance(x, _np.ndarray) and x.shape != ():
            return _np.vectorize(lambda v: SXElem(f(_raw(v))), otypes=[object])(x)
        return SXElem(f(_raw(x)))
    return g


cos = _unary(_sp.cos)
sin = _unary(_sp.sin)
tan = _unary(_sp.tan)
sqrt = _unary(_sp.sqrt)
exp = _unary(_sp.exp)
log = _unary(_sp.log)
erf = _unary(_sp.erf)
fabs = _unary(_sp.Abs)
arctan = _unary(_sp.atan)
atan = arctan


class _fmod(_sp.Function):
    """C `fmod` (result has the sign of the dividend); piecewise-constant offset => d/dx = 1."""
    nargs = 2

    def fdiff(self, argindex=1):
        if argindex == 1:
            return _sp.Integer(1)
        return -_sp.floor(self.args[0] / self.args[1])

    @classmethod
    def eval(cls, a, b):
        if a.is_Number and b.is_Number:
            import math
            return _sp.Float(math.fmod(float(a), float(b)))


def fmod(a, b):
    return SXElem(_fmod(_raw(a), _raw(b)))


def atan2(y, x):
    return SXElem(_sp.atan2(_raw(y), _raw(x)))


def fmax(a, b):
    return SXElem(_sp.Max(_raw(a), _raw(b)))


def fmin(a, b):
    return SXElem(_sp.Min(_raw(a), _raw(b)))


def to_sympy(v):
    """Column list of sympy expressions for an SX/SXElem/number (generator-side helper)."""
    if isinstance(v, SX):
        return list(v.m)
    return [_raw(v)]
