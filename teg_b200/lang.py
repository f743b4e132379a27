"""Minimal base-language node classes, attribute-compatible with Teg's own.

The CUDA backend consumes *reduced* Teg programs (base language only).  When the
real ``teg`` package is importable its node classes are consumed directly (the
frontend in :mod:`teg_b200.tegp` duck-types on class names).  On machines where
``teg`` is not installed (e.g. the GPU test box) these classes let users and
tests build base-language programs with the same spelling as the reference:

    x, t = TegVar('x'), Var('t', 0.5)
    expr = Teg(0, 1, IfElse(x < t, 1, 0), x)

Mirrors (names, constructor arguments, attributes):
  reference teg/lang/base.py:12-189  (ITeg, Var, Const, SmoothFunc, Add, Mul, Invert,
                                      IfElse, Tup, LetIn, Bool, And, Or)
  reference teg/lang/teg.py:7-36     (TegVar, Teg)
  reference teg/lang/operator_overloads.py:25-120 (arithmetic / comparison sugar)
  reference teg/math/smooth.py:14-146 (Sqrt, Sqr, Sin, Cos, ASin, ATan2)

Derivative transforms and delta reduction are *not* here: they are the program
producer (out of scope); their output reaches this backend either as live
``teg`` objects or as serialized ``.tegp`` programs.
"""
from __future__ import annotations

from typing import List, Optional


def _const(x):
    if isinstance(x, (int, float)) and not isinstance(x, bool):
        return Const(x)
    try:
        import numpy as np
        if isinstance(x, (np.integer, np.floating)):
            return Const(x.item())
    except ImportError:  # pragma: no cover
        pass
    return x


class ITeg:
    def __init__(self, children: List["ITeg"]):
        self.children = children
        for child in children:
            assert isinstance(child, ITeg), f"Non-ITeg expression {child} cannot be used in graph."
        self.value = None

    # -- operator sugar (reference teg/lang/operator_overloads.py:25-99) --
    def __add__(self, other):
        return Add([self, _const(other)])

    def __radd__(self, other):
        return _const(other) + self

    def __neg__(self):
        return Const(-1) * self

    def __sub__(self, other):
        return self + (-_const(other))

    def __rsub__(self, other):
        return _const(other) - self

    def __mul__(self, other):
        return Mul([self, _const(other)])

    def __rmul__(self, other):
        return _const(other) * self

    def __truediv__(self, other):
        return Mul([self, Invert(_const(other))])

    def __rtruediv__(self, other):
        return _const(other) / self

    def __pow__(self, exp):
        exp = _const(exp)
        if exp.value == 0:
            return Const(1)
        assert isinstance(exp.value, int) and exp.value > 0, "We only support positive integer powers."
        return self * self ** (exp.value - 1)

    def __lt__(self, other):
        return Bool(self, _const(other))

    def __le__(self, other):
        return Bool(self, _const(other), allow_eq=True)

    def __gt__(self, other):
        return Bool(_const(other), self)

    def __ge__(self, other):
        return Bool(_const(other), self, allow_eq=True)

    __hash__ = object.__hash__


class Var(ITeg):
    global_uid = 0

    def __init__(self, name: str, value: Optional[float] = None, uid: Optional[int] = None):
        super().__init__(children=[])
        self.name = name
        self.value = value
        if uid is None:
            self.uid = Var.global_uid
            Var.global_uid += 1
        else:
            self.uid = uid

    def __eq__(self, other):
        return isinstance(other, Var) and self.name == other.name and self.uid == other.uid

    def __hash__(self):
        return hash(f'{self.name}_{self.uid}')

    def __repr__(self):
        return f'{type(self).__name__}({self.name}_{self.uid})'


class Const(Var):
    def __init__(self, value, name: str = ""):
        super().__init__(name=name, value=value)


class TegVar(Var):
    def __init__(self, name: str = '', uid: Optional[int] = None):
        super().__init__(name=name, uid=uid)


class Add(ITeg):
    name = "add"


class Mul(ITeg):
    name = "mul"


class Invert(ITeg):
    def __init__(self, child):
        super().__init__(children=[_const(child)])
        self.child = self.children[0]


class SmoothFunc(ITeg):
    def __init__(self, expr, name: str = "SmoothFunc"):
        super().__init__(children=[_const(expr)])
        self.expr = self.children[0]
        self.name = name


class Sqrt(SmoothFunc):
    pass


class Sqr(SmoothFunc):
    pass


class Sin(SmoothFunc):
    pass


class Cos(SmoothFunc):
    pass


class ASin(SmoothFunc):
    pass


class ATan2(SmoothFunc):
    """theta = atan2(Tup(a, b)) -- reference teg/math/smooth.py:108-129."""
    pass


class IfElse(ITeg):
    def __init__(self, cond, if_body, else_body):
        super().__init__(children=[_const(e) for e in (cond, if_body, else_body)])
        self.cond, self.if_body, self.else_body = self.children


class Tup(ITeg):
    def __init__(self, *args):
        super().__init__(children=[_const(e) for e in args])

    def __iter__(self):
        return iter(self.children)

    def __len__(self):
        return len(self.children)


class LetIn(ITeg):
    def __init__(self, new_vars, new_exprs, expr):
        new_exprs = list(new_exprs)
        super().__init__(children=[_const(e) for e in (expr, *new_exprs)])
        self.new_vars = list(new_vars)
        self.new_exprs = self.children[1:]
        self.expr = self.children[0]


class Teg(ITeg):
    def __init__(self, lower, upper, body, dvar: TegVar):
        super().__init__(children=[_const(e) for e in (lower, upper, body)])
        self.lower, self.upper, self.body = self.children
        assert isinstance(dvar, TegVar), f'Can only integrate over TegVar variables. {dvar} is not a TegVar'
        self.dvar = dvar


class ITegBool(ITeg):
    # boolean combinators (reference operator_overloads.py:325-396 attach & and | to ITegBool)
    def __and__(self, other):
        return And(self, other)

    def __or__(self, other):
        return Or(self, other)


class Bool(ITegBool):
    def __init__(self, left_expr, right_expr, allow_eq: bool = False):
        super().__init__(children=[_const(e) for e in (left_expr, right_expr)])
        self.left_expr, self.right_expr = self.children
        self.allow_eq = allow_eq


class And(ITegBool):
    def __init__(self, left_expr, right_expr):
        super().__init__(children=[left_expr, right_expr])
        self.left_expr, self.right_expr = self.children


class Or(ITegBool):
    def __init__(self, left_expr, right_expr):
        super().__init__(children=[left_expr, right_expr])
        self.left_expr, self.right_expr = self.children


true = Bool(Const(0), Const(1))
false = Bool(Const(1), Const(0))

__all__ = [
    'ITeg', 'Var', 'Const', 'TegVar', 'Add', 'Mul', 'Invert', 'SmoothFunc', 'Sqrt', 'Sqr', 'Sin', 'Cos',
    'ASin', 'ATan2', 'IfElse', 'Tup', 'LetIn', 'Teg', 'ITegBool', 'Bool', 'And', 'Or', 'true', 'false',
]
