"""Domain culling: skip the per-sample comparisons of instances whose discontinuities miss the integration box.

A rasterisation integrand is a step function: ``IfElse(inside(x, y), shade, bg)``.  For all but the pixels an
edge actually crosses, ``inside`` has ONE value over the whole pixel box, and the reference C backend
(teg/ir/passes/to_c.py:396-487) still evaluates every comparison at every sample.  Here each top-level loop
nest gets a per-instance *interval evaluation* of its select conditions over the sampled box:

    tri = domtri(cond)          # 1: true at every sample, 0: false at every sample, 2: unresolved
    out = tri == 2 ? quad                                     # the reference's nest, untouched
                   : (tri == 1 ? quad[cond := true] : quad[cond := false])

The specialised nests contain no comparison; what does not depend on the integration variables any more is
hoisted by the scheduler (an inside pixel of the triangle costs 32 additions instead of 256 samples), and a nest
whose integrand became the literal zero collapses to ``0 + 0 * step`` (the exact value of its left fold).

**Exactness.**  The kernels compute every value with individually rounded IEEE operations (``--fmad=false``)
and round-to-nearest is monotone: for a <= a', RN(a + b) <= RN(a' + b), likewise for products with a fixed
sign, quotients and sqrt.  So the corner values of an operation over a box of operands, computed with the
*same* rounded operations, bound the rounded value at every point inside.  The sample abscissae lie between
the first and the last midpoint (``lb + step * (i + 0.5)`` is monotone in i).  A comparison is declared uniform
only when the bounds separate (``a.hi < b.lo`` / ``a.lo >= b.hi``); anything that could hide a NaN or an
infinity poisons its interval to NaN, which no decision accepts.  Whenever a condition is declared uniform the
specialised nest therefore performs the same additions of the same terms in the same order as the reference's:
results are bit-identical, not approximately equal (tests/test_gpu_cull.py, tests/test_host_logic.py).

With several distinct conditions in one nest the resolved values are passed as per-instance booleans
(``quad[cond_j := tri_j == 1]``) instead of compile-time constants.

New DAG ops: ``domtri`` (float-typed tri-state; attr = (condition node, bound variables of the nest); args =
the per-instance values the interval evaluation reads) and ``domis`` (bool; attr 'T' / 'U').
"""
from __future__ import annotations

from typing import Dict, FrozenSet, List, Optional, Sequence, Set, Tuple

from .dag import Dag, DNode

IV_FLOAT_OPS = ('add', 'mul', 'div', 'sqrt', 'sel', 'bvar')      # interval-evaluable ops that depend on the box
IV_BOOL_OPS = ('lt', 'and', 'or')


class _Nest:
    """The cone of one top-level integral: its bound variables, nesting depth and select conditions."""

    def __init__(self, dag: Dag, q: int):
        self.dag = dag
        self.q = q
        nodes = dag.nodes
        self.cone: Set[int] = set(dag.reachable([q]))
        self.bvars: Set[int] = {nodes[i].attr for i in self.cone if nodes[i].op == 'quad'}
        depth: Dict[int, int] = {}
        for i in sorted(self.cone):              # node ids are topological (arguments are interned first)
            n = nodes[i]
            depth[i] = max([depth[a] for a in n.args], default=0) + (1 if n.op == 'quad' else 0)
        self.depth = depth[q]
        self.B: FrozenSet[int] = frozenset(self.bvars)

    def interior(self, i: int) -> bool:
        return bool(self.dag.free_bvars(i) & self.B)

    def roots(self) -> List[int]:
        """Conditions of the selects that are evaluated per sample."""
        nodes = self.dag.nodes
        out = []
        for i in sorted(self.cone):
            n = nodes[i]
            if n.op == 'sel' and self.interior(n.args[0]) and n.args[0] not in out:
                out.append(n.args[0])
        return out

    def interval_cone(self, root: int) -> Optional[Tuple[List[int], List[int]]]:
        """(interior nodes in dependency order, frontier nodes) of a condition, or None when some interior
        operation has no interval rule (transcendentals, integrals) or a loop's bounds move with the box."""
        nodes = self.dag.nodes
        order: List[int] = []
        frontier: List[int] = []
        seen: Set[int] = set()
        stack = [(root, False)]
        while stack:
            i, done = stack.pop()
            if done:
                order.append(i)
                continue
            if i in seen:
                continue
            seen.add(i)
            n = nodes[i]
            if not self.interior(i):
                if n.op != 'const':
                    frontier.append(i)
                continue
            if n.op not in IV_FLOAT_OPS and n.op not in IV_BOOL_OPS:
                return None
            stack.append((i, True))
            if n.op == 'bvar':
                if self.interior(n.args[0]) or self.interior(n.args[1]):
                    return None
            for a in n.args:
                stack.append((a, False))
        return order, sorted(frontier)


class _Rebuilder:
    """Copy of a cone with some conditions replaced by constants (``known``) or by other bool nodes (``repl``)."""

    def __init__(self, dag: Dag, num_samples: int, known: Dict[int, bool], repl: Dict[int, int],
                 finite_pairs: Sequence[Tuple[int, int]] = ()):
        # finite_pairs: (lo, hi) bound pairs whose width is known to be finite where the copy runs (tile-resolved
        # code: the classification kernel poisons on every edge and width of the tile), so ``0 + 0 * (hi - lo)``
        # is the literal zero there
        self.finite_pairs = set(finite_pairs)
        self.dag = dag
        self.N = int(num_samples)
        self.known = known
        self.repl = repl
        self.memo: Dict[int, object] = {}
        self.zeroish: Set[int] = set()          # nodes whose value is +0 (or NaN when a step is not finite)

    def _zero(self) -> int:
        return self.dag.const('0')

    def is_zeroish(self, i: int) -> bool:
        n = self.dag.nodes[i]
        return i in self.zeroish or (n.op == 'const' and float(n.attr) == 0.0 and not n.attr.strip().startswith('-'))

    def bool_of(self, i: int):
        """-> True / False / node id"""
        if i in self.known:
            return self.known[i]
        if i in self.repl:
            return self.repl[i]
        n = self.dag.nodes[i]
        if n.op in ('and', 'or'):
            a, b = self.bool_of(n.args[0]), self.bool_of(n.args[1])
            if n.op == 'and':
                if a is False or b is False:
                    return False
                if a is True:
                    return b
                if b is True:
                    return a
            else:
                if a is True or b is True:
                    return True
                if a is False:
                    return b
                if b is False:
                    return a
            return self.dag.intern(n.op, (a, b))
        if n.op == 'lt':
            return self.dag.intern('lt', (self.value(n.args[0]), self.value(n.args[1])))
        if n.op == 'domis':
            return self.dag.intern('domis', (self.value(n.args[0]),), n.attr)
        return i

    def value(self, root: int) -> int:
        dag, memo = self.dag, self.memo
        nodes = dag.nodes
        stack = [root]
        while stack:
            i = stack[-1]
            if i in memo:
                stack.pop()
                continue
            n = nodes[i]
            if n.op in ('const', 'in', 'domtri'):
                memo[i] = i                      # (a preset memo entry replaces a domtri by a tile input)
                stack.pop()
                continue
            if n.op == 'sel':
                c = self.bool_of(n.args[0])
                need = [n.args[1]] if c is True else [n.args[2]] if c is False else [n.args[1], n.args[2]]
            else:
                need = list(n.args)
            pend = [a for a in need if a not in memo and not nodes[a].is_bool]
            if pend:
                stack.extend(pend)
                continue
            stack.pop()
            if n.op == 'sel':
                c = self.bool_of(n.args[0])
                if c is True:
                    memo[i] = memo[n.args[1]]
                elif c is False:
                    memo[i] = memo[n.args[2]]
                else:
                    memo[i] = dag.sel(c, memo[n.args[1]], memo[n.args[2]])
            elif n.op == 'bvar':
                memo[i] = dag.intern('bvar', (memo[n.args[0]], memo[n.args[1]]), n.attr)
            elif n.op == 'quad':
                lo, hi, body = (memo[a] for a in n.args)
                bv = memo[n.attr] if n.attr in memo else n.attr
                if self.is_zeroish(body) and (lo, hi) in self.finite_pairs:
                    memo[i] = self._zero()
                elif self.is_zeroish(body):
                    # left fold of N terms (zero * step) from +0: exactly 0 + zero * step.  zero * ((hi - lo) / N) and
                    # zero * (hi - lo) are the same value (a zero of the sign of hi - lo, or NaN when hi - lo is not
                    # finite), so the division is not emitted
                    width = dag.intern('add', (hi, dag.intern('mul', (dag.const('-1'), lo))))
                    r = dag.intern('add', (self._zero(), dag.intern('mul', (body, width))))
                    self.zeroish.add(r)
                    memo[i] = r
                else:
                    memo[i] = dag.intern('quad', (lo, hi, body), bv)
            elif n.op == 'add':
                memo[i] = dag.add(memo[n.args[0]], memo[n.args[1]])
            elif n.op == 'mul':
                memo[i] = dag.mul(memo[n.args[0]], memo[n.args[1]])
            else:
                memo[i] = dag.intern(n.op, tuple(memo[a] for a in n.args), n.attr)
        return memo[root]


def _extend_known(nodes: Sequence[DNode], cond: int, value: bool) -> Dict[int, bool]:
    k: Dict[int, bool] = {}
    stack = [(cond, value)]
    while stack:
        c, v = stack.pop()
        k[c] = v
        n = nodes[c]
        if n.op == 'and' and v:
            stack += [(n.args[0], True), (n.args[1], True)]
        elif n.op == 'or' and not v:
            stack += [(n.args[0], False), (n.args[1], False)]
    return k


def cull(dag: Dag, num_samples: int, min_samples: int = 64, max_roots: int = 4, max_cone: int = 96,
         min_depth: int = 2) -> Dag:
    """Rewrite every closed top-level integral of ``dag`` whose select conditions admit interval evaluation.
    Returns a copy of ``dag`` (outputs replaced); ``.culled`` lists (integral, conditions) for the record."""
    dag = dag.clone()
    nodes = dag.nodes
    N = int(num_samples)
    memo: Dict[int, int] = {}
    record: List[Tuple[int, Tuple[int, ...]]] = []

    def wrap(q: int) -> int:
        nest = _Nest(dag, q)
        # one-dimensional nests (the delta integrals of reverse-mode programs) are short and usually straddle their
        # discontinuity: the test would cost more than it saves
        if nest.depth < min_depth or float(N) ** nest.depth < min_samples:
            return q
        roots = nest.roots()
        if not roots or len(roots) > max_roots:
            return q
        tris = []
        for r in roots:
            cone = nest.interval_cone(r)
            if cone is None or len(cone[0]) > max_cone:
                return q
            order, frontier = cone
            tris.append(dag.intern('domtri', tuple(frontier), (r, nest.B)))
        record.append((q, tuple(roots)))
        unresolved = None
        for t in tris:
            u = dag.intern('domis', (t,), 'U')
            unresolved = u if unresolved is None else dag.intern('or', (unresolved, u))
        if len(roots) == 1:
            is_true = dag.intern('domis', (tris[0],), 'T')
            q_t = _Rebuilder(dag, N, _extend_known(nodes, roots[0], True), {}).value(q)
            q_f = _Rebuilder(dag, N, _extend_known(nodes, roots[0], False), {}).value(q)
            spec = dag.sel(is_true, q_t, q_f)
        else:
            repl = {r: dag.intern('domis', (t,), 'T') for r, t in zip(roots, tris)}
            spec = _Rebuilder(dag, N, {}, repl).value(q)
        return dag.sel(unresolved, q, spec)

    def rewrite(root: int) -> int:
        stack = [root]
        while stack:
            i = stack[-1]
            if i in memo:
                stack.pop()
                continue
            n = nodes[i]
            if n.op == 'quad' and not dag.free_bvars(i):
                memo[i] = wrap(i)
                stack.pop()
                continue
            if dag.free_bvars(i) or not dag.has_quad(i):
                memo[i] = i                      # inside a nest, or nothing to rewrite below
                stack.pop()
                continue
            pend = [a for a in n.args if a not in memo]
            if pend:
                stack.extend(pend)
                continue
            stack.pop()
            args = tuple(memo[a] for a in n.args)
            if args == n.args:
                memo[i] = i
            elif n.op == 'sel':
                memo[i] = dag.sel(*args)
            elif n.op == 'add':
                memo[i] = dag.add(*args)
            elif n.op == 'mul':
                memo[i] = dag.mul(*args)
            else:
                memo[i] = dag.intern(n.op, args, n.attr)
        return memo[root]

    dag.outputs = tuple(rewrite(o) for o in dag.outputs)
    dag.culled = record
    return dag


# ------------------------------------------------------------------------------------------ tile level
def interval_evaluable(dag: Dag, root: int, interior, forbidden=None) -> bool:
    """Can the cone of ``root`` above the nodes for which ``interior`` is false be bounded by intervals?
    ``forbidden(i)``: node i must not occur in a point (per-pixel inputs that are not box edges)."""
    nodes = dag.nodes
    seen: Set[int] = set()
    stack = [root]
    while stack:
        i = stack.pop()
        if i in seen:
            continue
        if not interior(i):
            # a point of the classification kernel: must be plain launch-invariant arithmetic
            if dag.has_quad(i) or any(nodes[j].op in ('domtri', 'domis') or (forbidden is not None and forbidden(j))
                                      for j in dag.reachable([i])):
                return False
            continue
        seen.add(i)
        n = nodes[i]
        if n.op == 'in':
            continue
        if n.op not in IV_FLOAT_OPS and n.op not in IV_BOOL_OPS:
            return False
        stack.extend(n.args)
    return len(seen) <= 160


def tile_cull(dag: Dag, grid_args: Sequence[int], num_samples: int, max_conds: int = 12,
              other_array_args: Sequence[int] = ()) -> Dag:
    """Second level of culling for render launches (pixel boxes generated from the pixel index): the same interval
    evaluation, but over a whole TILE of pixels, done once per tile by a small classification kernel.

    Tile conditions: (a) every ``domtri`` of the per-instance culling (now: "uniform over every sample of every
    pixel of the tile"), (b) every per-instance guard of an integral (the bound checks around the delta terms of
    reverse-mode programs: "true / false for every pixel of the tile").  Condition k becomes the pseudo argument
    ``n_args + k`` (a tri-state the main kernel loads per block), and the program gets several *bodies*, chosen per
    tile by the kernel (``.bodies`` = [(label, {condition: admitted states} | None, outputs)], first match wins):

        quiet / quiet0 / quiet1   every guard false over the tile (a single domain condition fixed to 0 / 1): static
        resolved                  every condition resolved: conditions substituted by the tile's values
        general                   some condition unresolved: the per-instance program, untouched

    In a resolved tile no pixel evaluates an interval or a guard: interior pixels of the triangle run the 32
    additions, exterior ones a handful of instructions, and a reverse-mode program whose guards are all false
    reduces to its constant zeros.  Exactness as for ``cull``: the tile box contains every pixel box of the tile,
    bounds are monotone.  Returns a copy with ``.tile_conds`` = [(kind, node)] ('dom' | 'guard'); unchanged
    (``tile_conds`` empty) when there is nothing to classify."""
    dag = dag.clone()
    dag.tile_conds = []
    nodes = dag.nodes
    n_args = len(dag.arg_labels)
    grid = set(grid_args)
    other = set(other_array_args) - grid          # per-pixel inputs that are not box edges (an upstream dL image)
    live = dag.reachable(dag.outputs)

    def forbidden(j: int) -> bool:
        return nodes[j].op == 'in' and nodes[j].attr in other

    dep_cache: Dict[int, bool] = {}

    def on_grid(i: int) -> bool:
        """does node i depend on a pixel-box argument?"""
        if i in dep_cache:
            return dep_cache[i]
        stack = [i]
        while stack:
            j = stack[-1]
            if j in dep_cache:
                stack.pop()
                continue
            n = nodes[j]
            pend = [a for a in n.args if a not in dep_cache]
            if pend:
                stack.extend(pend)
                continue
            dep_cache[j] = (n.op == 'in' and n.attr in grid) or any(dep_cache[a] for a in n.args)
            stack.pop()
        return dep_cache[i]

    conds: List[Tuple[str, int]] = []
    for i in live:
        n = nodes[i]
        if n.op == 'domtri':
            root, B = n.attr
            if interval_evaluable(dag, root, lambda j, B=B: bool(dag.free_bvars(j) & B) or on_grid(j), forbidden):
                conds.append(('dom', i))
    for i in live:
        n = nodes[i]
        if n.op != 'sel' or dag.free_bvars(i):
            continue
        c = n.args[0]
        if not (dag.has_quad(n.args[1]) or dag.has_quad(n.args[2])):
            continue
        if dag.free_bvars(c) or dag.has_quad(c) or not on_grid(c) or ('guard', c) in conds:
            continue
        if any(nodes[j].op in ('domis', 'domtri') for j in dag.reachable([c])):
            continue
        if interval_evaluable(dag, c, on_grid, forbidden):
            conds.append(('guard', c))
    if not conds or len(conds) > max_conds:
        return dag
    tins = [dag.intern('in', (), n_args + k) for k in range(len(conds))]
    in_node = {a: dag.intern('in', (), a) for a in grid_args}
    gl = list(grid_args)
    pairs = [(in_node[gl[0]], in_node[gl[1]]), (in_node[gl[2]], in_node[gl[3]])] if len(gl) == 4 else []

    def build(states: Dict[int, Optional[int]]) -> Tuple[int, ...]:
        """outputs with condition k fixed to states[k] (0 / 1), or read from its tile input (None)"""
        rb = _Rebuilder(dag, num_samples, {}, {}, finite_pairs=pairs)
        for k, (kind, node) in enumerate(conds):
            st = states[k]
            if kind == 'dom':
                if st is None:
                    rb.known[dag.intern('domis', (node,), 'U')] = False
                    rb.memo[node] = tins[k]
                else:
                    # specialise the original nest again (not cull's copy): zero nests become the literal zero here
                    rb.known[dag.intern('domis', (node,), 'U')] = True
                    rb.known.update(_extend_known(nodes, nodes[node].attr[0], bool(st)))
            elif st is None:
                rb.repl[node] = dag.intern('domis', (tins[k],), 'T')
            else:
                rb.known.update(_extend_known(nodes, node, bool(st)))
        return tuple(rb.value(o) for o in dag.outputs)

    doms = [k for k, (kind, _) in enumerate(conds) if kind == 'dom']
    guards = [k for k, (kind, _) in enumerate(conds) if kind == 'guard']
    every = list(range(len(conds)))
    bodies: List[Tuple[str, Optional[Dict[int, Tuple[int, ...]]], Tuple[int, ...]]] = []
    # most tiles: no guard holds anywhere in the tile (no discontinuity line crosses it)
    quiet = {k: 0 for k in guards}
    if len(doms) == 1:
        for st in (0, 1):
            bodies.append((f'quiet{st}', {**{k: (0,) for k in guards}, doms[0]: (st,)}, build({**quiet, doms[0]: st})))
    elif guards:
        bodies.append(('quiet', {**{k: (0,) for k in guards}, **{k: (0, 1) for k in doms}},
                       build({**quiet, **{k: None for k in doms}})))
    if guards or len(doms) != 1:
        bodies.append(('resolved', {k: (0, 1) for k in every}, build({k: None for k in every})))
    bodies.append(('general', None, tuple(dag.outputs)))
    dag.bodies = bodies
    dag.tile_conds = conds
    return dag
