"""Minimal `casadi` stand-in backed by SymPy.

Implements exactly the surface src/models/robot_model.py:5,14-64 and
src/simulation/robot_ocp_problem.py:21-22 use: `SX.sym`, `vertcat`, `sin`, `cos`, arithmetic,
`**`, integer indexing and `.size()`.  CasADi itself is not installable in the build image
(no network); the model file is executed UNMODIFIED against this shim and its expression
graph is then lowered to CUDA device functions by codegen/gen_model.py.
"""
import sympy as sp


class SX:
    """column vector of sympy expressions"""

    def __init__(self, elems):
        self.elems = list(elems)

    @staticmethod
    def sym(name, n=1, m=1):
        assert m == 1
        if n == 1:
            return SX([sp.Symbol(name, real=True)])
        return SX([sp.Symbol(f"{name}_{i}", real=True) for i in range(n)])

    def size(self):
        return (len(self.elems), 1)

    @property
    def shape(self):
        return self.size()

    def __len__(self):
        return len(self.elems)

    def __getitem__(self, i):
        r = self.elems[i]
        return SX(r) if isinstance(r, list) else SX([r])

    def _bin(self, other, op):
        if isinstance(other, SX):
            a, b = self.elems, other.elems
            if len(a) == 1 and len(b) > 1:
                a = a * len(b)
            if len(b) == 1 and len(a) > 1:
                b = b * len(a)
            assert len(a) == len(b)
            return SX([op(x, y) for x, y in zip(a, b)])
        return SX([op(x, sp.sympify(other)) for x in self.elems])

    def __add__(self, o): return self._bin(o, lambda x, y: x + y)
    def __radd__(self, o): return self._bin(o, lambda x, y: y + x)
    def __sub__(self, o): return self._bin(o, lambda x, y: x - y)
    def __rsub__(self, o): return self._bin(o, lambda x, y: y - x)
    def __mul__(self, o): return self._bin(o, lambda x, y: x * y)
    def __rmul__(self, o): return self._bin(o, lambda x, y: y * x)
    def __truediv__(self, o): return self._bin(o, lambda x, y: x / y)
    def __pow__(self, o): return self._bin(o, lambda x, y: x ** y)
    def __neg__(self): return SX([-x for x in self.elems])


def vertcat(*args):
    out = []
    for a in args:
        if isinstance(a, SX):
            out.extend(a.elems)
        else:
            out.append(sp.sympify(a))
    return SX(out)


def sin(a):
    return SX([sp.sin(x) for x in a.elems]) if isinstance(a, SX) else sp.sin(a)


def cos(a):
    return SX([sp.cos(x) for x in a.elems]) if isinstance(a, SX) else sp.cos(a)
