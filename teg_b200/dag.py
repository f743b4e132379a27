"""Semantic DAG of a reduced Teg program (the backend's own IR).

This replaces the reference's tree -> ``IR_Function`` lowering (teg/passes/compile.py:51-216,
teg/ir/instr/*.py) with a representation designed for batched GPU evaluation:

  * scalar SSA values, structurally hash-consed (the reference has no CSE: teg/__main__.py:69-70
    leaves it commented out; a triangle gradient is a 22.5k-node tree but < 1k unique values);
  * ``LetIn`` is eliminated by substitution (compile.py:154-180 binds by label; values are pure, so
    substituting the bound value's node is the same computation, shared);
  * ``Tup`` vectors are scalarised component-wise -- the same arithmetic the reference performs with
    ``generic_array`` loops (to_c.py:263-283, teg_generic_array.h:23-69);
  * integration variables are canonical: two integrals with the same bounds at the same nesting depth
    share one bound variable, so equal integrands and equal integrals (and equal guards around them)
    collapse, and integrals over the same domain can later be fused into one sample loop;
  * ``Invert`` -> ``1.0 / x`` (compile.py:69-74), ``Sqr`` -> ``x * x`` (c_math.py:34-37),
    ``Bool`` -> strict ``<`` whatever ``allow_eq`` says (compile.py:210-211).

Only rewrites that are exact in IEEE arithmetic are applied while building:
``x + 0 -> x`` **only where x is provably not -0** (the reference computes ``-0 + 0 = +0``; sums that
start from an integral's accumulator, squares and non-zero literals can never be -0, see ``never_neg_zero``),
``1 * x -> x``, ``sel(c, a, a) -> a``.  Nothing is re-associated, distributed or contracted.
"""
from __future__ import annotations

import sys
from dataclasses import dataclass
from typing import Dict, FrozenSet, List, Optional, Sequence, Tuple, Union

from .tegp import Program

Value = Union[int, Tuple[int, ...]]          # scalar node id, or a vector of scalar node ids

FLOAT_OPS = ('const', 'in', 'bvar', 'add', 'mul', 'div', 'sel', 'sqrt', 'sin', 'cos', 'asin', 'atan2', 'quad',
             'domtri')                       # domtri / domis: cull.py
BOOL_OPS = ('lt', 'and', 'or', 'domis')

_UNARY = {'Sqrt': 'sqrt', 'Sin': 'sin', 'Cos': 'cos', 'ASin': 'asin'}


@dataclass(frozen=True)
class DNode:
    op: str
    args: Tuple[int, ...] = ()
    attr: object = None      # const: decimal text; in: argument index; bvar: nesting depth; quad: bvar id

    @property
    def is_bool(self) -> bool:
        return self.op in BOOL_OPS


class Dag:
    def __init__(self, name: str, arg_labels: Sequence[str], defaults: Dict[str, Optional[float]]):
        self.name = name
        self.nodes: List[DNode] = []
        self._table: Dict[tuple, int] = {}
        self.arg_labels = list(arg_labels)
        self.defaults = dict(defaults)
        self.outputs: Tuple[int, ...] = ()
        self.out_is_tuple = False
        self.out_groups: List[int] = []      # bundles: number of outputs of each member program
        self._fv: Dict[int, FrozenSet[int]] = {}
        self._has_quad: Dict[int, bool] = {}
        self._nnz: Dict[int, bool] = {}
        self._qb: Dict[int, FrozenSet[int]] = {}

    def clone(self) -> "Dag":
        """Independent copy (node ids are preserved) for transforms that add nodes and replace the outputs."""
        d = Dag(self.name, self.arg_labels, self.defaults)
        d.nodes = list(self.nodes)
        d._table = dict(self._table)
        d.outputs = self.outputs
        d.out_is_tuple = self.out_is_tuple
        d.out_groups = list(self.out_groups)
        d._fv = dict(self._fv)
        d._has_quad = dict(self._has_quad)
        d._nnz = dict(self._nnz)
        d._qb = dict(self._qb)
        return d

    # ------------------------------------------------------------ building
    def intern(self, op: str, args: Tuple[int, ...] = (), attr=None) -> int:
        key = (op, args, attr)
        idx = self._table.get(key)
        if idx is None:
            idx = len(self.nodes)
            self.nodes.append(DNode(op, args, attr))
            self._table[key] = idx
        return idx

    def const(self, text: str) -> int:
        return self.intern('const', (), text)

    def _is_const(self, i: int, value: float) -> bool:
        n = self.nodes[i]
        return n.op == 'const' and float(n.attr) == value and not (value == 0 and n.attr.lstrip().startswith('-'))

    def never_neg_zero(self, i: int) -> bool:
        """Is the value of node i provably never -0?  (Round to nearest: a sum is -0 only when both operands are -0,
        so an integral -- a left fold that starts at +0 -- never is; nor is x * x, a literal other than -0, or a
        select between two such values.)  Conservative: False when unknown."""
        nnz = self._nnz
        stack = [i]
        while stack:
            j = stack[-1]
            if j in nnz:
                stack.pop()
                continue
            n = self.nodes[j]
            if n.op == 'quad':
                nnz[j] = True
            elif n.op == 'const':
                nnz[j] = not (float(n.attr) == 0.0 and n.attr.lstrip().startswith('-'))
            elif n.op == 'add':
                pend = [a for a in n.args if a not in nnz]
                if pend:
                    stack.extend(pend)
                    continue
                nnz[j] = nnz[n.args[0]] or nnz[n.args[1]]
            elif n.op == 'sel':
                pend = [a for a in n.args[1:] if a not in nnz]
                if pend:
                    stack.extend(pend)
                    continue
                nnz[j] = nnz[n.args[1]] and nnz[n.args[2]]
            elif n.op == 'mul':
                nnz[j] = n.args[0] == n.args[1]
            else:
                nnz[j] = False
            stack.pop()
        return nnz[i]

    def add(self, a: int, b: int) -> int:
        # x + 0 == x bit for bit unless x is -0 (-0 + 0 = +0): fold only where x cannot be -0
        if self._is_const(b, 0.0) and self.never_neg_zero(a):
            return a
        if self._is_const(a, 0.0) and self.never_neg_zero(b):
            return b
        return self.intern('add', (a, b))

    def mul(self, a: int, b: int) -> int:
        if self._is_const(b, 1.0):
            return a
        if self._is_const(a, 1.0):
            return b
        return self.intern('mul', (a, b))

    def sel(self, c: int, a: int, b: int) -> int:
        if a == b:
            return a
        return self.intern('sel', (c, a, b))

    # ------------------------------------------------------------ analyses
    def free_bvars(self, i: int) -> FrozenSet[int]:
        """Bound (integration) variables occurring free in node i."""
        fv = self._fv
        if i in fv:
            return fv[i]
        # iterative post-order to avoid deep recursion
        stack = [i]
        while stack:
            j = stack[-1]
            if j in fv:
                stack.pop()
                continue
            n = self.nodes[j]
            pending = [a for a in n.args if a not in fv]
            if pending:
                stack.extend(pending)
                continue
            if n.op == 'bvar':
                r = frozenset((j,)) | fv[n.args[0]] | fv[n.args[1]]
            elif n.op == 'quad':
                r = fv[n.args[0]] | fv[n.args[1]] | (fv[n.args[2]] - {n.attr})
            else:
                r = frozenset().union(*(fv[a] for a in n.args)) if n.args else frozenset()
            fv[j] = r
            stack.pop()
        return fv[i]

    def has_quad(self, i: int) -> bool:
        hq = self._has_quad
        if i in hq:
            return hq[i]
        stack = [i]
        while stack:
            j = stack[-1]
            if j in hq:
                stack.pop()
                continue
            n = self.nodes[j]
            pending = [a for a in n.args if a not in hq]
            if pending:
                stack.extend(pending)
                continue
            hq[j] = n.op == 'quad' or any(hq[a] for a in n.args)
            stack.pop()
        return hq[i]

    def quad_bvars(self, i: int) -> FrozenSet[int]:
        """Bound variables of every integral contained in node i (closed or not)."""
        qb = self._qb
        stack = [i]
        while stack:
            j = stack[-1]
            if j in qb:
                stack.pop()
                continue
            n = self.nodes[j]
            pending = [a for a in n.args if a not in qb]
            if pending:
                stack.extend(pending)
                continue
            r = frozenset().union(*(qb[a] for a in n.args)) if n.args else frozenset()
            if n.op == 'quad':
                r = r | {n.attr}
            qb[j] = r
            stack.pop()
        return qb[i]

    def reachable(self, roots: Sequence[int]) -> List[int]:
        seen = set()
        stack = list(roots)
        while stack:
            j = stack.pop()
            if j in seen:
                continue
            seen.add(j)
            stack.extend(self.nodes[j].args)
        return sorted(seen)

    def stats(self) -> Dict[str, int]:
        live = self.reachable(self.outputs)
        ops: Dict[str, int] = {}
        for j in live:
            ops[self.nodes[j].op] = ops.get(self.nodes[j].op, 0) + 1
        return {'live_nodes': len(live), **ops}


def build_dag(prog: Program) -> Dag:
    """TEGP program -> Dag.  Raises the same exception types as the reference lowering for the same
    causes (ValueError for inexpressible nodes / unbound structure, AssertionError for typing violations
    that teg/ir/passes/typing.py asserts on, NotImplementedError-like AssertionError for ``Exp``)."""
    dag = Dag(prog.name, prog.args, prog.defaults)
    nodes = prog.nodes
    arg_index = {lab: i for i, lab in enumerate(prog.args)}

    # free labels per TEGP node (for memoising conversion under an environment)
    free: List[Tuple[str, ...]] = [()] * len(nodes)
    for i, n in enumerate(nodes):
        if n.op == 'V':
            free[i] = (n.attr,)
        elif n.op == 'C':
            free[i] = ()
        elif n.op == 'T':
            s = set(free[n.kids[0]]) | set(free[n.kids[1]]) | (set(free[n.kids[2]]) - {n.attr})
            free[i] = tuple(sorted(s))
        elif n.op == 'L':
            k = len(n.attr)
            s = set()
            for e in n.kids[:k]:
                s |= set(free[e])
            s |= set(free[n.kids[k]]) - set(n.attr)
            free[i] = tuple(sorted(s))
        else:
            s = set()
            for kid in n.kids:
                s |= set(free[kid])
            free[i] = tuple(sorted(s))

    memo: Dict[tuple, Value] = {}

    def scalar(v: Value, what: str) -> int:
        assert isinstance(v, int), f'{what} must be a scalar'
        return v

    def broadcast2(a: Value, b: Value, fn) -> Value:
        # typing.py:125-137: sizes must match or one side is scalar
        if isinstance(a, int) and isinstance(b, int):
            return fn(a, b)
        if isinstance(a, int):
            return tuple(fn(a, y) for y in b)
        if isinstance(b, int):
            return tuple(fn(x, b) for x in a)
        assert len(a) == len(b), 'Binary operands have incompatible sizes'
        return tuple(fn(x, y) for x, y in zip(a, b))

    def conv(i: int, env: Dict[str, Value], depth: int) -> Value:
        n = nodes[i]
        key = (i, depth if _needs_depth[i] else -1, tuple(env.get(l) for l in free[i]))
        hit = memo.get(key)
        if hit is not None:
            return hit
        r = conv_(i, n, env, depth)
        memo[key] = r
        return r

    # a node's conversion depends on the binder depth only if it contains an integral
    _needs_depth = [False] * len(nodes)
    for i, n in enumerate(nodes):
        _needs_depth[i] = n.op == 'T' or any(_needs_depth[k] for k in n.kids)

    def conv_(i: int, n, env: Dict[str, Value], depth: int) -> Value:
        op = n.op
        if op == 'C':
            return dag.const(n.attr)
        if op == 'V':
            if n.attr in env:
                return env[n.attr]
            if n.attr not in arg_index:
                raise ValueError(f'unbound variable {n.attr}')
            return dag.intern('in', (), arg_index[n.attr])
        if op == 'A':
            return broadcast2(conv(n.kids[0], env, depth), conv(n.kids[1], env, depth), dag.add)
        if op == 'M':
            return broadcast2(conv(n.kids[0], env, depth), conv(n.kids[1], env, depth), dag.mul)
        if op == 'I':
            one = dag.const('1.0')
            x = conv(n.kids[0], env, depth)
            return broadcast2(one, x, lambda a, b: dag.intern('div', (a, b)))
        if op == 'F':
            x = conv(n.kids[0], env, depth)
            if n.attr == 'ATan2':
                assert not isinstance(x, int) and len(x) == 2, 'atan2() takes exactly 2 arguments'
                return dag.intern('atan2', (x[0], x[1]))
            if n.attr == 'Sqr':
                f = lambda a: dag.intern('mul', (a, a))           # noqa: E731  (c_math.py:34-37: x * x)
            elif n.attr in _UNARY:
                f = lambda a, o=_UNARY[n.attr]: dag.intern(o, (a,))   # noqa: E731
            else:
                # to_c.py:338-341: "Math operation ... has no defined C transformation"
                raise AssertionError(f'Math operation {n.attr} has no defined C transformation.')
            return f(x) if isinstance(x, int) else tuple(f(e) for e in x)
        if op == 'B':
            a = scalar(conv(n.kids[0], env, depth), 'comparison operand')
            b = scalar(conv(n.kids[1], env, depth), 'comparison operand')
            return dag.intern('lt', (a, b))
        if op in ('&', '|'):
            a = scalar(conv(n.kids[0], env, depth), 'boolean operand')
            b = scalar(conv(n.kids[1], env, depth), 'boolean operand')
            assert dag.nodes[a].is_bool and dag.nodes[b].is_bool, 'Binary operands have incompatible types'
            return dag.intern('and' if op == '&' else 'or', (a, b))
        if op == 'S':
            c = scalar(conv(n.kids[0], env, depth), 'condition')
            assert dag.nodes[c].is_bool, 'Condition type must be a scalar boolean'
            a = conv(n.kids[1], env, depth)
            b = conv(n.kids[2], env, depth)
            if isinstance(a, int) and isinstance(b, int):
                return dag.sel(c, a, b)
            # typing.py:146-148 asserts equal branch types; the reference's scalar/vector mix is
            # undefined behaviour and is rejected here instead of reproduced.
            assert not isinstance(a, int) and not isinstance(b, int) and len(a) == len(b), \
                'if and else branches have incompatible types'
            return tuple(dag.sel(c, x, y) for x, y in zip(a, b))
        if op == 'U':
            elems = []
            for k in n.kids:
                e = conv(k, env, depth)
                assert isinstance(e, int), 'Teg IR currently does not support multidimensional packing'
                elems.append(e)
            return tuple(elems)
        if op == 'J':               # bundle root: concatenate the members' results
            flat: List[int] = []
            sizes = []
            for kdx in n.kids:
                v = conv(kdx, env, depth)
                v = (v,) if isinstance(v, int) else tuple(v)
                sizes.append(len(v))
                flat.extend(v)
            dag.out_groups = sizes
            return tuple(flat)
        if op == 'L':
            k = len(n.attr)
            vals = [conv(e, env, depth) for e in n.kids[:k]]       # parallel let: outer scope
            env2 = dict(env)
            for lab, v in zip(n.attr, vals):
                env2[lab] = v
            return conv(n.kids[k], env2, depth)
        if op == 'T':
            lo = scalar(conv(n.kids[0], env, depth), 'integration bound')
            hi = scalar(conv(n.kids[1], env, depth), 'integration bound')
            # Canonical binder = (bounds, nesting depth) -- unless a value substituted into the body (a LetIn-bound
            # expression, converted in an outer scope) already contains an integral over that very binder: nesting
            # the two would make the inner integrand read the outer loop's variable.  Such a binder moves to the
            # next free depth (it only loses sharing with integrals it must not be confused with).
            taken = set()
            for lab in free[n.kids[2]]:
                v = env.get(lab) if lab != n.attr else None
                for e in (() if v is None else ((v,) if isinstance(v, int) else v)):
                    taken |= dag.quad_bvars(e)
            d = depth
            while dag._table.get(('bvar', (lo, hi), d)) in taken and taken:
                d += 1
            bvar = dag.intern('bvar', (lo, hi), d)
            env2 = dict(env)
            env2[n.attr] = bvar
            body = conv(n.kids[2], env2, d + 1)
            mk = lambda b: dag.intern('quad', (lo, hi, b), bvar)   # noqa: E731
            return mk(body) if isinstance(body, int) else tuple(mk(b) for b in body)
        raise ValueError(f'The type of the expr "{op}" cannot be expressed in IR.')

    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 200000))
    try:
        out = conv(prog.root, {}, 0)
    finally:
        sys.setrecursionlimit(old)
    if isinstance(out, int):
        assert not dag.nodes[out].is_bool, 'program result must be a float expression'
        dag.outputs = (out,)
        dag.out_is_tuple = False
    else:
        dag.outputs = tuple(out)
        dag.out_is_tuple = True
    if not dag.out_groups:
        dag.out_groups = [len(dag.outputs)]
    return dag
