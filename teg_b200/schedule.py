"""Region scheduling: Dag -> structured per-instance program (blocks of evals, guards, sample loops).

The reference emits one C function per ``IfElse`` branch / integral body / let body, nested exactly
like the source tree, and leaves hoisting to the C compiler (teg/ir/passes/to_c.py:396-619).  Here
the structure is decided explicitly, because it is what determines GPU cost:

  * **rates** -- every value is evaluated at the slowest rate it can be: *uniform* values (functions
    of scalar-bound arguments only) once per launch by a 1-thread prologue kernel, published through
    ``__constant__`` memory; *per-instance* values once per thread; *per-sample* values inside the
    sample loops.  A value whose free integration variables do not include a loop's variable is
    hoisted out of that loop.
  * **guards** -- ``sel`` nodes whose branches hold real work (in particular the bound checks that
    ``eliminate_deltas`` wraps around every delta integral, teg/passes/delta.py:426-439) become real
    ``if`` statements so inactive instances skip the integral, as in to_c.py:459-487.  All selects on
    the same condition share one ``if``; inside it the condition is *known*, so nested selects on it
    fold away.  Cheap branches are evaluated speculatively and become ``cond ? a : b`` (values are
    pure, so speculation cannot change a result).
  * **loop fusion** -- integrals over the same domain (same bounds, same canonical variable) that do
    not depend on one another run as one sample loop with several accumulators.  A reverse-mode
    triangle gradient is 12 delta integrals in the reference; they become 3 loops here.

Sample placement and the per-accumulator left fold are untouched: a fused loop visits exactly the
reference's midpoints ``lb + step * (i + 0.5)`` in the same order for every accumulator.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, FrozenSet, List, Optional, Sequence, Set, Tuple

from .dag import Dag


# ----------------------------------------------------------------------------- statements
@dataclass
class Eval:
    node: int                       # plain op, or a speculated select (ternary)


@dataclass
class IfGroup:
    cond: int
    sels: List[int]                 # select nodes defined by this guard
    then_block: "Block" = None
    else_block: "Block" = None


@dataclass
class LoopGroup:
    bvar: int
    lo: int
    hi: int
    quads: List[int]                # quad nodes accumulated by this loop
    body: "Block" = None


@dataclass
class Block:
    stmts: List[object] = field(default_factory=list)
    known: Dict[int, bool] = field(default_factory=dict)     # conditions whose value is implied here
    spec: Set[int] = field(default_factory=set)              # selects evaluated speculatively (ternaries)

    def walk(self):
        for s in self.stmts:
            yield s
            if isinstance(s, IfGroup):
                yield from s.then_block.walk()
                yield from s.else_block.walk()
            elif isinstance(s, LoopGroup):
                yield from s.body.walk()


@dataclass
class Schedule:
    dag: Dag
    main: Block
    uniform_nodes: List[int]        # topological list of nodes evaluated by the prologue kernel
    uniform_slots: Dict[int, int]   # node -> index into the published uniform table
    array_args: List[int]           # argument indices that are per-instance arrays
    scalar_args: List[int]          # argument indices bound to one scalar for the whole launch
    outputs: Tuple[int, ...]
    # tile-culled render programs (cull.tile_cull): alternative bodies chosen per tile, [(label, match, block, outputs)];
    # the last one is the general body (== main)
    bodies: List[Tuple[str, Optional[Dict[int, Tuple[int, ...]]], Block, Tuple[int, ...]]] = field(default_factory=list)

    def loops(self) -> List[LoopGroup]:
        return [s for s in self.main.walk() if isinstance(s, LoopGroup)]

    def guards(self) -> List[IfGroup]:
        return [s for s in self.main.walk() if isinstance(s, IfGroup)]

    def loop_depth(self) -> int:
        """Deepest nesting of sample loops (0 = the program has no integral left)."""
        def depth(b: Block) -> int:
            d = 0
            for st in b.stmts:
                if isinstance(st, LoopGroup):
                    d = max(d, 1 + depth(st.body))
                elif isinstance(st, IfGroup):
                    d = max(d, depth(st.then_block), depth(st.else_block))
            return d
        return depth(self.main)


class Scheduler:
    def __init__(self, dag: Dag, array_args: Sequence[int], spec_limit: int = 6,
                 fuse_loops: bool = True, uniform_hoist: bool = True):
        self.dag = dag
        self.nodes = dag.nodes
        self.array_args = sorted(set(array_args))
        self.spec_limit = spec_limit
        self.fuse_loops = fuse_loops
        self.uniform_hoist = uniform_hoist
        self._varying: Dict[int, bool] = {}
        self.uniform_used: Set[int] = set()      # uniform nodes referenced from per-instance code

    # ------------------------------------------------------------------ rates
    def varying(self, i: int) -> bool:
        """Does node i depend on a per-instance array argument?"""
        var = self._varying
        if i in var:
            return var[i]
        arr = set(self.array_args)
        stack = [i]
        while stack:
            j = stack[-1]
            if j in var:
                stack.pop()
                continue
            n = self.nodes[j]
            pending = [a for a in n.args if a not in var]
            if pending:
                stack.extend(pending)
                continue
            var[j] = (n.op == 'in' and n.attr in arr) or any(var[a] for a in n.args)
            stack.pop()
        return var[i]

    def is_uniform(self, i: int) -> bool:
        """Launch-invariant, loop-free, integral-free: evaluated once by the prologue kernel."""
        if not self.uniform_hoist:
            return False
        n = self.nodes[i]
        if n.op == 'const':
            return False            # literals are immediates, not table entries
        return (not self.varying(i)) and (not self.dag.free_bvars(i)) and (not self.dag.has_quad(i))

    def is_free(self, i: int) -> bool:
        """No code needed at the use site (literal or published uniform value)."""
        return self.nodes[i].op == 'const' or self.is_uniform(i)

    # ------------------------------------------------------------- known conds
    @staticmethod
    def _extend_known(nodes, known: Dict[int, bool], cond: int, value: bool) -> Dict[int, bool]:
        k = dict(known)
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

    def bool_value(self, c: int, known: Dict[int, bool]) -> Optional[bool]:
        if c in known:
            return known[c]
        n = self.nodes[c]
        if n.op in ('and', 'or'):
            a = self.bool_value(n.args[0], known)
            b = self.bool_value(n.args[1], known)
            if n.op == 'and':
                if a is False or b is False:
                    return False
                if a is True and b is True:
                    return True
            else:
                if a is True or b is True:
                    return True
                if a is False and b is False:
                    return False
        return None

    def resolve(self, i: int, known: Dict[int, bool]) -> int:
        """Follow selects whose condition is implied by the enclosing guards."""
        while True:
            n = self.nodes[i]
            if n.op != 'sel':
                return i
            v = self.bool_value(n.args[0], known)
            if v is None:
                return i
            i = n.args[1] if v else n.args[2]

    # --------------------------------------------------------- strict closure
    def strict_closure(self, roots: Sequence[int], avail: Set[int], known: Dict[int, bool],
                       spec: Set[int]) -> List[int]:
        """Nodes that must be evaluated whenever this region runs, in dependency (post) order.

        Branches of (non-speculated) selects are lazy; integrand bodies belong to their loop, except
        for their loop-invariant strict part, which is needed as soon as the loop runs at all."""
        order: List[int] = []
        seen: Set[int] = set()
        dag = self.dag

        def visit(root: int):
            stack = [(root, False)]
            while stack:
                i, done = stack.pop()
                if done:
                    order.append(i)
                    continue
                i = self.resolve(i, known)
                if i in seen or i in avail or self.is_free(i):
                    if self.is_uniform(i):
                        self.uniform_used.add(i)
                    continue
                n = self.nodes[i]
                if n.is_bool and self.bool_value(i, known) is not None:
                    continue
                seen.add(i)
                stack.append((i, True))
                if n.op == 'sel':
                    stack.append((n.args[0], False))
                    if i in spec:
                        stack.append((n.args[1], False))
                        stack.append((n.args[2], False))
                elif n.op == 'quad':
                    lo, hi, body = n.args
                    bvar = n.attr
                    inner, _ = self.region_strict([body], avail | seen | {bvar}, known)
                    for m in reversed(inner):
                        if bvar not in dag.free_bvars(m):
                            stack.append((m, False))
                    stack.append((hi, False))
                    stack.append((lo, False))
                elif n.op == 'bvar':
                    pass            # defined by the enclosing loop (always in avail there)
                else:
                    for a in reversed(n.args):
                        stack.append((a, False))

        for r in roots:
            visit(r)
        return order

    def _lazy_sets(self, strict: List[int], avail: Set[int], known: Dict[int, bool], spec: Set[int]):
        """For every guarded select in ``strict``: the extra nodes each branch would need."""
        base = avail | set(strict)
        out = []
        for s in strict:
            n = self.nodes[s]
            if n.op != 'sel' or s in spec:
                continue
            c = n.args[0]
            kt = self._extend_known(self.nodes, known, c, True)
            kf = self._extend_known(self.nodes, known, c, False)
            then_set = self.strict_closure([n.args[1]], base, kt, spec)
            else_set = self.strict_closure([n.args[2]], base, kf, spec)
            out.append((s, c, then_set, else_set))
        return out

    def region_strict(self, roots: Sequence[int], avail: Set[int], known: Dict[int, bool]):
        """Fixpoint: the strict node list of a region, after deciding which selects are cheap enough to
        evaluate speculatively and hoisting (integral-free) work that several guards would repeat."""
        roots = [self.resolve(r, known) for r in roots]
        spec: Set[int] = set()
        extra_roots: List[int] = []
        strict: List[int] = []
        for _ in range(64):
            strict = self.strict_closure(list(roots) + extra_roots, avail, known, spec)
            lazy = self._lazy_sets(strict, avail, known, spec)
            changed = False
            owners: Dict[int, Set[Tuple[int, bool]]] = {}
            for s, c, then_set, else_set in lazy:
                n = self.nodes[s]
                quad_free = not self.dag.has_quad(n.args[1]) and not self.dag.has_quad(n.args[2])
                if quad_free and len(then_set) + len(else_set) <= self.spec_limit:
                    spec.add(s)
                    changed = True
                    continue
                for m in then_set:
                    owners.setdefault(m, set()).add((c, True))
                for m in else_set:
                    owners.setdefault(m, set()).add((c, False))
            if not changed:
                for m, own in owners.items():
                    if len(own) >= 2 and not self.dag.has_quad(m) and m not in extra_roots:
                        extra_roots.append(m)
                        changed = True
            if not changed:
                break
        return strict, spec

    # ----------------------------------------------------------------- region
    def region(self, roots: Sequence[int], avail: Set[int], known: Dict[int, bool]) -> Block:
        strict, spec = self.region_strict(roots, avail, known)
        strict_set = set(strict)

        # dependencies between strict nodes (through lazy branches and integrand bodies as well)
        deps_cache: Dict[int, Set[int]] = {}

        def deps_of(i: int) -> Set[int]:
            if i in deps_cache:
                return deps_cache[i]
            n = self.nodes[i]
            found: Set[int] = set()

            def scan(start: Sequence[int], kn: Dict[int, bool], bound: FrozenSet[int]):
                seen_local: Set[int] = set()
                stack = list(start)
                while stack:
                    j = self.resolve(stack.pop(), kn)
                    if j in seen_local or j in avail or self.is_free(j) or j in bound:
                        continue
                    seen_local.add(j)
                    if j in strict_set and j != i:
                        found.add(j)
                        continue
                    m = self.nodes[j]
                    if m.is_bool and self.bool_value(j, kn) is not None:
                        continue
                    if m.op == 'sel' and j not in spec:
                        stack.append(m.args[0])
                        scan([m.args[1]], self._extend_known(self.nodes, kn, m.args[0], True), bound)
                        scan([m.args[2]], self._extend_known(self.nodes, kn, m.args[0], False), bound)
                    elif m.op == 'quad':
                        stack += [m.args[0], m.args[1]]
                        scan([m.args[2]], kn, bound | {m.attr})
                    else:
                        stack.extend(m.args)

            if n.op == 'sel' and i not in spec:
                c = n.args[0]
                scan([c], known, frozenset())
                scan([n.args[1]], self._extend_known(self.nodes, known, c, True), frozenset())
                scan([n.args[2]], self._extend_known(self.nodes, known, c, False), frozenset())
            elif n.op == 'quad':
                scan([n.args[0], n.args[1]], known, frozenset())
                scan([n.args[2]], known, frozenset({n.attr}))
            else:
                scan(list(n.args), known, frozenset())
            deps_cache[i] = found
            return found

        def is_group_member(i: int) -> bool:
            n = self.nodes[i]
            return (n.op == 'sel' and i not in spec) or n.op == 'quad'

        # Kahn topological order, plain evaluations first so guards / loops can merge
        indeg = {i: 0 for i in strict}
        users: Dict[int, List[int]] = {i: [] for i in strict}
        for i in strict:
            for d in deps_of(i):
                indeg[i] += 1
                users[d].append(i)
        rank = {i: k for k, i in enumerate(strict)}
        ready_plain = sorted([i for i in strict if indeg[i] == 0 and not is_group_member(i)], key=rank.get)
        ready_group = sorted([i for i in strict if indeg[i] == 0 and is_group_member(i)], key=rank.get)
        block = Block(known=dict(known))
        pos: Dict[int, int] = {}
        emitted = 0
        while ready_plain or ready_group:
            if ready_plain:
                i = ready_plain.pop(0)
            else:
                i = ready_group.pop(0)
            n = self.nodes[i]
            placed = False
            dmax = max([pos[d] for d in deps_of(i)], default=-1)
            if n.op == 'sel' and i not in spec:
                for idx, st in enumerate(block.stmts):
                    if isinstance(st, IfGroup) and st.cond == n.args[0] and idx > dmax:
                        st.sels.append(i)
                        pos[i] = idx
                        placed = True
                        break
                if not placed:
                    block.stmts.append(IfGroup(cond=n.args[0], sels=[i]))
            elif n.op == 'quad':
                if self.fuse_loops:
                    for idx, st in enumerate(block.stmts):
                        if (isinstance(st, LoopGroup) and st.bvar == n.attr and st.lo == n.args[0]
                                and st.hi == n.args[1] and idx > dmax):
                            st.quads.append(i)
                            pos[i] = idx
                            placed = True
                            break
                if not placed:
                    block.stmts.append(LoopGroup(bvar=n.attr, lo=n.args[0], hi=n.args[1], quads=[i]))
            else:
                block.stmts.append(Eval(i))
            if not placed:
                pos[i] = len(block.stmts) - 1
            emitted += 1
            newly = []
            for u in users[i]:
                indeg[u] -= 1
                if indeg[u] == 0:
                    newly.append(u)
            for u in sorted(newly, key=rank.get):
                (ready_group if is_group_member(u) else ready_plain).append(u)
            ready_plain.sort(key=rank.get)
            ready_group.sort(key=rank.get)
        assert emitted == len(strict), 'cyclic dependency while scheduling (internal error)'

        # sub-regions
        defined_before: List[Set[int]] = []
        acc = set(avail)
        for st in block.stmts:
            defined_before.append(set(acc))
            if isinstance(st, Eval):
                acc.add(st.node)
            elif isinstance(st, IfGroup):
                acc.update(st.sels)
            else:
                acc.update(st.quads)
        for st, av in zip(block.stmts, defined_before):
            if isinstance(st, IfGroup):
                kt = self._extend_known(self.nodes, known, st.cond, True)
                kf = self._extend_known(self.nodes, known, st.cond, False)
                st.then_block = self.region([self.nodes[s].args[1] for s in st.sels], av, kt)
                st.else_block = self.region([self.nodes[s].args[2] for s in st.sels], av, kf)
            elif isinstance(st, LoopGroup):
                st.body = self.region([self.nodes[q].args[2] for q in st.quads], av | {st.bvar}, known)
        block.spec = set(spec)
        return block


def make_schedule(dag: Dag, array_args: Sequence[int], **opts) -> Schedule:
    sch = Scheduler(dag, array_args, **opts)
    main = sch.region(list(dag.outputs), set(), {})
    for o in dag.outputs:
        if sch.is_uniform(o):
            sch.uniform_used.add(o)
    bodies = []
    for label, match, outs in getattr(dag, 'bodies', []):
        if match is None:
            bodies.append((label, None, main, tuple(outs)))
            continue
        blk = sch.region(list(outs), set(), {})
        for o in outs:
            if sch.is_uniform(o):
                sch.uniform_used.add(o)
        bodies.append((label, dict(match), blk, tuple(outs)))
    # prologue: every uniform value referenced from per-instance code, with its (uniform) dependencies
    need: List[int] = []
    seen: Set[int] = set()

    def visit(i: int):
        stack = [(i, False)]
        while stack:
            j, done = stack.pop()
            if done:
                need.append(j)
                continue
            if j in seen or dag.nodes[j].op == 'const':
                continue
            seen.add(j)
            stack.append((j, True))
            for a in reversed(dag.nodes[j].args):
                stack.append((a, False))

    for u in sorted(sch.uniform_used):
        visit(u)
    slots = {u: k for k, u in enumerate(sorted(sch.uniform_used))}
    n_args = len(dag.arg_labels)
    arr = set(sch.array_args)
    return Schedule(dag=dag, main=main, uniform_nodes=need, uniform_slots=slots,
                    array_args=sorted(arr), scalar_args=[a for a in range(n_args) if a not in arr],
                    outputs=tuple(dag.outputs), bodies=bodies)
