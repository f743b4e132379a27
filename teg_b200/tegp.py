"""TEGP: the exchange format for *reduced* (base-language) Teg programs.

A TEGP program is a structurally hash-consed copy of a base-language expression
tree: the exact node kinds of reference teg/lang/base.py + teg/lang/teg.py, with
variables identified by their C label ``f'{name}_{uid}'`` exactly as the
reference lowering does (teg/passes/compile.py:62-67).  Nothing is evaluated,
re-associated or simplified here: the format is a lossless picture of what the
reference's ``to_ir`` would be handed, so that

  * the CUDA backend (teg_b200.dag / schedule / codegen),
  * the CPU oracle (oracle/teg_oracle.c), and
  * machines without the ``teg`` frontend (the GPU box)

all consume the same program.  ``from_teg`` accepts live objects of either the
reference ``teg`` package or ``teg_b200.lang`` (duck-typed on class names, the
same way compile.py:182 detects FwdDeriv/RevDeriv by MRO name).

Text form (one node per line, children always precede parents):

    TEGP 1
    name <identifier>
    nodes <count>
    C <decimal-literal>              constant (decimal repr kept verbatim: each consumer
                                     parses it in its own precision, like the C literal
                                     emitted at to_c.py:231-247)
    V <label> <default|none>         Var / TegVar reference by label
    A <a> <b>     M <a> <b>          Add / Mul           (compile.py:185-198)
    I <a>                            Invert -> 1.0 / a   (compile.py:69-74)
    F <Name> <a>                     SmoothFunc subclass (compile.py:76-81, c_math.py)
    S <c> <a> <b>                    IfElse              (compile.py:83-103)
    T <lo> <hi> <body> <label>       Teg over TegVar     (compile.py:105-135)
    U <n> <a0> ... <an-1>            Tup                 (compile.py:137-152)
    L <k> <label0> <e0> ... <body>   LetIn (parallel)    (compile.py:154-180)
    B <a> <b>                        Bool: strict a < b, allow_eq ignored (compile.py:210-211)
    & <a> <b>     | <a> <b>          And / Or            (compile.py:205-208)
    J <n> <r0> ... <rn-1>            (backend-only) root of a *bundle*: the results of several programs that
                                     share arguments, evaluated together in one kernel (see ``bundle``)
    root <id>
    args <n> <label0> ... <labeln-1>   input order (teg/__main__.py:86-88 semantics)
"""
from __future__ import annotations

import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

OPS = ('C', 'V', 'A', 'M', 'I', 'F', 'S', 'T', 'U', 'L', 'B', '&', '|', 'J')

SMOOTH_FUNCS = ('Sqr', 'Sqrt', 'Sin', 'Cos', 'ASin', 'ATan2', 'Exp')


def label_of(var) -> str:
    return f'{var.name}_{var.uid}'


def _num_repr(v) -> str:
    """Decimal text of a constant exactly as Python's ``str`` prints it.

    The reference emits ``f'{value}'`` (+ ``f`` suffix in fp32) into the C source
    (to_c.py:231-247), so the decimal string *is* the constant's definition.
    """
    if isinstance(v, bool):
        raise ValueError('boolean constants are not part of the base language')
    try:
        import numpy as np
        if isinstance(v, (np.integer, np.floating)):
            v = v.item()
    except ImportError:  # pragma: no cover
        pass
    if isinstance(v, int):
        return str(v)
    if isinstance(v, float):
        return repr(v)
    return repr(float(v))


@dataclass
class Node:
    op: str
    kids: Tuple[int, ...] = ()
    # op-specific payload: 'C' -> decimal string; 'V' -> label; 'F' -> func name;
    # 'T' -> dvar label; 'L' -> tuple of bound labels
    attr: object = None


@dataclass
class Program:
    name: str
    nodes: List[Node] = field(default_factory=list)
    root: int = -1
    args: List[str] = field(default_factory=list)            # ordered input labels
    defaults: Dict[str, Optional[float]] = field(default_factory=dict)
    # label -> live Var object when the program was converted from live expressions (not serialised): the
    # reference reads ``teg_var.value`` at *eval* time (c_eval.py:136), so later bind_variable() calls count
    live: Dict[str, object] = field(default_factory=dict, repr=False, compare=False)

    # ------------------------------------------------------------------ io
    def dumps(self) -> str:
        out = ['TEGP 1', f'name {self.name}', f'nodes {len(self.nodes)}']
        for n in self.nodes:
            if n.op == 'C':
                out.append(f'C {n.attr}')
            elif n.op == 'V':
                d = self.defaults.get(n.attr)
                out.append(f'V {n.attr} {"none" if d is None else _num_repr(d)}')
            elif n.op == 'F':
                out.append(f'F {n.attr} {n.kids[0]}')
            elif n.op == 'T':
                out.append(f'T {n.kids[0]} {n.kids[1]} {n.kids[2]} {n.attr}')
            elif n.op in ('U', 'J'):
                out.append(n.op + ' ' + ' '.join(map(str, (len(n.kids), *n.kids))))
            elif n.op == 'L':
                k = len(n.attr)
                parts = [str(k)]
                for lab, e in zip(n.attr, n.kids[:k]):
                    parts += [lab, str(e)]
                parts.append(str(n.kids[k]))
                out.append('L ' + ' '.join(parts))
            else:
                out.append(n.op + ' ' + ' '.join(map(str, n.kids)))
        out.append(f'root {self.root}')
        out.append('args ' + ' '.join([str(len(self.args)), *self.args]))
        return '\n'.join(out) + '\n'

    def dump(self, path: str) -> None:
        with open(path, 'w') as f:
            f.write(self.dumps())

    @staticmethod
    def loads(text: str) -> "Program":
        lines = [ln for ln in text.splitlines() if ln.strip()]
        if not lines or lines[0].split() != ['TEGP', '1']:
            raise ValueError('not a TEGP 1 program')
        name = lines[1].split(None, 1)[1]
        count = int(lines[2].split()[1])
        prog = Program(name=name)
        for ln in lines[3:3 + count]:
            t = ln.split()
            op = t[0]
            if op == 'C':
                prog.nodes.append(Node('C', (), t[1]))
            elif op == 'V':
                prog.nodes.append(Node('V', (), t[1]))
                prog.defaults[t[1]] = None if t[2] == 'none' else float(t[2])
            elif op == 'F':
                if t[1] not in SMOOTH_FUNCS:
                    raise ValueError(f'unknown smooth function {t[1]}')
                prog.nodes.append(Node('F', (int(t[2]),), t[1]))
            elif op == 'T':
                prog.nodes.append(Node('T', (int(t[1]), int(t[2]), int(t[3])), t[4]))
            elif op == 'U':
                k = int(t[1])
                prog.nodes.append(Node('U', tuple(int(v) for v in t[2:2 + k])))
            elif op == 'L':
                k = int(t[1])
                labels = tuple(t[2 + 2 * i] for i in range(k))
                exprs = tuple(int(t[3 + 2 * i]) for i in range(k))
                prog.nodes.append(Node('L', (*exprs, int(t[2 + 2 * k])), labels))
            elif op == 'J':
                k = int(t[1])
                prog.nodes.append(Node('J', tuple(int(v) for v in t[2:2 + k])))
            elif op in ('A', 'M', 'B', '&', '|'):
                prog.nodes.append(Node(op, (int(t[1]), int(t[2]))))
            elif op == 'I':
                prog.nodes.append(Node('I', (int(t[1]),)))
            elif op == 'S':
                prog.nodes.append(Node('S', (int(t[1]), int(t[2]), int(t[3]))))
            else:
                raise ValueError(f'unknown TEGP op {op!r}')
        for i, n in enumerate(prog.nodes):
            if any(k >= i or k < 0 for k in n.kids):
                raise ValueError(f'node {i}: children must precede parents')
        tail = lines[3 + count:]
        prog.root = int(tail[0].split()[1])
        t = tail[1].split()
        prog.args = t[2:2 + int(t[1])]
        return prog

    @staticmethod
    def load(path: str) -> "Program":
        with open(path) as f:
            return Program.loads(f.read())

    # ------------------------------------------------------------- queries
    def free_labels(self) -> List[str]:
        """Free variables in first-use order (the order compile.py's free_symbols dict yields)."""
        order: List[str] = []
        seen = set()

        memo: Dict[int, Tuple[str, ...]] = {}

        def free(i: int) -> Tuple[str, ...]:
            if i in memo:
                return memo[i]
            n = self.nodes[i]
            if n.op == 'V':
                r = (n.attr,)
            elif n.op == 'C':
                r = ()
            elif n.op == 'T':
                lo, hi, body = n.kids
                r = _uniq((*free(lo), *free(hi), *(v for v in free(body) if v != n.attr)))
            elif n.op == 'L':
                k = len(n.attr)
                bound = set(n.attr)
                acc: List[str] = []
                for e in n.kids[:k]:
                    acc += free(e)
                acc += [v for v in free(n.kids[k]) if v not in bound]
                r = _uniq(acc)
            else:
                acc = []
                for kid in n.kids:
                    acc += free(kid)
                r = _uniq(acc)
            memo[i] = r
            return r

        old = sys.getrecursionlimit()
        sys.setrecursionlimit(max(old, 200000))
        try:
            for v in free(self.root):
                if v not in seen:
                    seen.add(v)
                    order.append(v)
        finally:
            sys.setrecursionlimit(old)
        return order

    def tree_size(self) -> int:
        """Number of nodes of the un-shared tree this DAG stands for."""
        size = [0] * len(self.nodes)
        for i, n in enumerate(self.nodes):
            size[i] = 1 + sum(size[k] for k in n.kids)
        return size[self.root]


def _uniq(seq: Sequence[str]) -> Tuple[str, ...]:
    seen = set()
    out = []
    for s in seq:
        if s not in seen:
            seen.add(s)
            out.append(s)
    return tuple(out)


# ---------------------------------------------------------------- frontend
def _mro_names(obj) -> set:
    return {t.__name__ for t in type(obj).__mro__}


def from_teg(expr, arglist: Sequence = (), name: str = 'teg_method') -> Program:
    """Convert a live base-language expression (reference ``teg`` or ``teg_b200.lang``).

    Raises ``ValueError`` for nodes that are not expressible, mirroring
    compile.py:215-216 (``The type of the expr ... cannot be expressed in IR``).
    """
    prog = Program(name=name)
    table: Dict[tuple, int] = {}

    def intern(op: str, kids: Tuple[int, ...] = (), attr=None) -> int:
        key = (op, kids, attr)
        idx = table.get(key)
        if idx is None:
            idx = len(prog.nodes)
            prog.nodes.append(Node(op, kids, attr))
            table[key] = idx
        return idx

    def conv(e) -> int:
        names = _mro_names(e)
        if 'Const' in names:
            return intern('C', (), _num_repr(e.value))
        if 'Var' in names:      # Var, TegVar (Placeholder is a Var too but never survives reduction)
            if 'Placeholder' in names:
                raise ValueError(f'The type of the expr "{type(e)}" cannot be expressed in IR.')
            lab = label_of(e)
            if lab not in prog.defaults or prog.defaults[lab] is None:
                prog.defaults[lab] = None if e.value is None else float(e.value)
            prog.live.setdefault(lab, e)
            return intern('V', (), lab)
        if 'FwdDeriv' in names or 'RevDeriv' in names:
            return conv(e.deriv_expr)
        if 'Invert' in names:
            return intern('I', (conv(e.child),))
        if 'SmoothFunc' in names:
            fn = type(e).__name__
            if fn not in SMOOTH_FUNCS:
                raise ValueError(f'SmoothFunc "{fn}" has no lowering')
            return intern('F', (conv(e.expr),), fn)
        if 'IfElse' in names:
            return intern('S', (conv(e.cond), conv(e.if_body), conv(e.else_body)))
        if 'Teg' in names:
            return intern('T', (conv(e.lower), conv(e.upper), conv(e.body)), label_of(e.dvar))
        if 'Tup' in names:
            return intern('U', tuple(conv(c) for c in e.children))
        if 'LetIn' in names:
            labels = tuple(label_of(v) for v in e.new_vars)
            exprs = tuple(conv(c) for c in e.new_exprs)
            return intern('L', (*exprs, conv(e.expr)), labels)
        if 'Add' in names or 'Mul' in names:
            assert len(e.children) == 2, 'Binary expression has an invalid number of operands'
            return intern('A' if 'Add' in names else 'M', (conv(e.children[0]), conv(e.children[1])))
        if 'And' in names:
            return intern('&', (conv(e.left_expr), conv(e.right_expr)))
        if 'Or' in names:
            return intern('|', (conv(e.left_expr), conv(e.right_expr)))
        if 'Bool' in names:
            return intern('B', (conv(e.left_expr), conv(e.right_expr)))
        raise ValueError(f'The type of the expr "{type(e)}" cannot be expressed in IR.')

    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 200000))
    try:
        prog.root = conv(expr)
    finally:
        sys.setrecursionlimit(old)

    free = prog.free_labels()
    ordered = [label_of(v) for v in arglist]
    for v in arglist:
        prog.live.setdefault(label_of(v), v)
        if prog.defaults.get(label_of(v)) is None and v.value is not None:
            prog.defaults[label_of(v)] = float(v.value)
    for lab in ordered:
        if lab not in free:
            # An argument the program does not use is still a (dead) input slot, as in
            # teg/__main__.py:86-88 it would raise KeyError; we keep the slot so callers
            # can pass a fixed signature (e.g. pixel boxes) to every program of a scene.
            prog.defaults.setdefault(lab, None)
    prog.args = list(ordered) + [lab for lab in free if lab not in ordered]
    return prog


def bundle(programs: Sequence[Program], name: Optional[str] = None) -> Program:
    """Several programs over the same variables (matched by label) as ONE program whose result is the
    concatenation of theirs -- e.g. a value and its reverse derivative, evaluated by one kernel that shares
    the argument loads and every common sub-expression.  The bundle root is the backend-only node ``J``."""
    out = Program(name=name or '_'.join(p.name for p in programs))
    table: Dict[tuple, int] = {}
    roots = []
    for p in programs:
        remap: List[int] = []
        for n in p.nodes:
            key = (n.op, tuple(remap[k] for k in n.kids), n.attr)
            idx = table.get(key)
            if idx is None:
                idx = len(out.nodes)
                out.nodes.append(Node(n.op, key[1], n.attr))
                table[key] = idx
            remap.append(idx)
        roots.append(remap[p.root])
        for lab in p.args:
            if lab not in out.args:
                out.args.append(lab)
        for lab, d in p.defaults.items():
            if out.defaults.get(lab) is None:
                out.defaults[lab] = d
        for lab, v in p.live.items():
            out.live.setdefault(lab, v)
    out.nodes.append(Node('J', tuple(roots)))
    out.root = len(out.nodes) - 1
    return out
