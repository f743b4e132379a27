"""CUDA emitter: Schedule -> one sm_100a translation unit (the new ``CUDA`` target).

Counterpart of the reference's text emitter (teg/ir/passes/to_c.py:260-651 + c_math.py +
data/templates/rectangular_rule.template.c), but what comes out is a *batched kernel*, not a header
of scalar ``__device__`` functions (which is all ``--target CUDA_C`` gives, teg/__main__.py:100-109):

  <name>_uniform   1 thread: evaluates every launch-invariant value once and writes the table that
                   the runtime copies into ``tegU`` (``__constant__``), so per-instance code reads
                   them as constant-bank operands: no registers, no loads.
  <name>_main      thread per instance: coalesced SoA loads of the per-instance arguments, the
                   scheduled guards / fused midpoint loops, coalesced SoA stores of the outputs.

Arithmetic is emitted one IEEE operation per reference operation, in the reference's order, and the
unit is compiled with ``--fmad=false``: sample placement ``lb + step * (i + 0.5)``, every comparison
operand and every left-fold ``out = out + f * step`` round exactly as in the C backend, so
discontinuity classification -- and therefore every result -- is reproducible bit for bit on the
same inputs (libm transcendentals excepted).

The two instance bodies are emitted as ``TEGB_FN`` functions; the runtime header makes that
``__device__ __forceinline__``.  (tests/ re-define it to compile the very same text for the host and
compare it with the oracle on machines without a GPU -- a checker, never a product path.)
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

from .dag import Dag
from .schedule import Block, Eval, IfGroup, LoopGroup, Schedule, Scheduler

# Philox4x32-10 (Salmon et al., SC'11) and the map to [0, 1): the same text as oracle/teg_oracle.c
_MC_PREAMBLE = '''#define TEGB_MC_SEED @SEED@
TEGB_FN T tegb_uniform(const unsigned long long seed, const unsigned long long inst, const unsigned int depth,
                       const unsigned int i) {
    unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
    unsigned int c0 = (unsigned int)inst, c1 = (unsigned int)(inst >> 32), c2 = depth, c3 = i;
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const unsigned int n0 = (unsigned int)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned int)p1;
        const unsigned int n2 = (unsigned int)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned int)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    if (sizeof(T) == 4) return (T)(c0 >> 8) * (T)(1.0 / 16777216.0);
    return (T)(((unsigned long long)(c0 >> 5) << 26) | (unsigned long long)(c1 >> 6)) * (T)(1.0 / 9007199254740992.0);
}'''

_MATH = {
    'float': {'sqrt': 'sqrtf', 'sin': 'sinf', 'cos': 'cosf', 'asin': 'asinf', 'atan2': 'atan2f'},
    'double': {'sqrt': 'sqrt', 'sin': 'sin', 'cos': 'cos', 'asin': 'asin', 'atan2': 'atan2'},
}


INTEGRATION_MODES = ('rectangular_quadrature', 'trapezoidal_quadrature', 'monte_carlo_uniform')


def literal(text: str, ctype: str) -> str:
    """Decimal constant in working precision, the way to_c.py:231-247 prints it (``1.0f`` / ``1``)."""
    t = text.strip()
    if not any(ch in t for ch in '.eEn'):      # integer spelling ('n' covers inf/nan)
        t += '.0'
    if 'inf' in t or 'nan' in t:
        raise ValueError('non-finite literal')
    return f'{t}f' if ctype == 'float' else t


@dataclass
class KernelSource:
    name: str
    ctype: str
    num_samples: int
    source: str
    uniform_kernel: str
    main_kernel: str
    n_uniform: int
    scalar_args: List[int]          # argument indices passed (by value) to the uniform kernel, in order
    array_args: List[int]           # argument indices passed (as pointers) to the main kernel, in order
    grid_args: List[int]            # (x_lb, x_ub, y_lb, y_ub) argument indices generated from the pixel index
    n_outputs: int
    block_size: int
    reduce_outputs: int             # number of trailing outputs summed over the instances (per-block partial sums)
    team: int                       # threads cooperating on one instance (1 = thread per instance)
    grouped: bool = False           # grouped render launch: per-group scalar arguments + pixel windows
    pixel_args: List[int] = None    # arguments read from pixel-indexed images (grouped launches)
    pixel_diff: List[int] = None    # the subset of pixel_args bound to a DIFFERENCE of two images (a - b, formed on load)
    accumulate: bool = False        # outputs are added into images
    vec: int = 1                    # instances per thread (16-byte vector loads/stores), streaming programs
    tile_conds: int = 0             # tile conditions classified by <name>_tiles before a render launch (cull.tile_cull)
    tile_rows: int = 0              # pixel rows per tile
    tile_cols: int = 0              # pixel columns per tile (whole warps; divides block_size)
    tile_kernel: str = ''
    gather_kernel: str = ''         # grouped accumulating programs: <name>_gather, one block per IMAGE block walking its bin of groups
    takes_instance_index: bool = False     # instance functions take the instance index (monte_carlo_uniform)
    tile_bodies: List = None        # [(label, {condition: admitted states})]: bodies <name>_tile_<label> for resolved tiles
    stats: Dict[str, float] = None

    @property
    def key(self) -> str:
        return hashlib.sha256(self.source.encode()).hexdigest()[:24]


class Emitter:
    def __init__(self, sch: Schedule, ctype: str, num_samples: int, name: str,
                 block_size: int = 128, unroll_inner: int = 32, unroll_outer: int = 1, pred_accum: bool = True, use_asm: bool = True, team: int = 1, coop: bool = True,
                 pred_mix: float = 0.25, jam: int = 4, coop_max_lanes: int = 4,
                 integration_mode: str = 'rectangular_quadrature', tile_rows: int = 16, tile_cols: int = 32,
                 mc_seed: int = 0, min_blocks: Optional[int] = None, coop_groups: int = 4):
        assert ctype in ('float', 'double')
        self.tile_rows = int(tile_rows)
        self.mc_seed = int(mc_seed) & 0xFFFFFFFFFFFFFFFF
        # resident blocks per SM the render kernels are compiled for (0: the compiler's choice; default: 8 for tiled
        # kernels, whose common path is a short streaming body that needs warps in flight -- the rare general path may
        # spill -- and the compiler's choice otherwise).  --maxrregcount is ignored for kernels with launch bounds, so
        # the register cap has to be expressed here.  Measured on the 4096^2 value kernel: 128 regs 41 us, 64 regs 35 us.
        self.min_blocks = None if min_blocks is None else int(min_blocks)
        # tiles are tile_cols columns wide: whole warps, so that the body choice is uniform within a warp
        assert tile_cols % 32 == 0 and block_size % tile_cols == 0, 'tile_cols: a multiple of 32 dividing the block size'
        self.tile_cols = int(tile_cols)
        # the reference names three modes (to_c.py:389-393) and ships a template for the first only
        assert integration_mode in INTEGRATION_MODES, f'unknown integration mode {integration_mode}'
        self.mode = integration_mode
        if self.mode != 'rectangular_quadrature':
            # the sample-parallel / blocked / predicated loop forms are derived for the midpoint left fold
            assert team == 1, 'teams are defined for rectangular_quadrature only'
            coop, jam, pred_accum = False, 0, False
        self.sch = sch
        self.dag: Dag = sch.dag
        self.nodes = sch.dag.nodes
        self.ctype = ctype
        self.N = int(num_samples)
        self.name = name
        self.block_size = block_size
        self.unroll_inner = unroll_inner
        self.unroll_outer = unroll_outer
        self.pred_accum = pred_accum
        self.use_asm = use_asm
        self.pred_mix = float(pred_mix)
        assert team >= 1 and (team & (team - 1)) == 0 and team <= 1024, 'team must be a power of two <= 1024'
        self.team = team
        self.helper = Scheduler(sch.dag, sch.array_args)      # for resolve / bool_value
        # warp-cooperative guarded loops (see _coop_if / _coop_loop); needs warp intrinsics, thread per instance
        self.coop = bool(coop) and team == 1
        self.guard_ctx: List[str] = []        # per-lane guard conditions enclosing the current (converged) code
        self.divergent = 0                    # > 0 while emitting inside an ordinary (divergent) if
        self.coop_map: Dict[int, str] = {}    # node -> name of its warp-broadcast copy inside a coop loop
        self.ref_log = None                   # dry-run bookkeeping: nodes referenced / defined
        self.def_log = None
        self.uses_coop = False
        self.jam_map: Dict[int, str] = {}     # node -> per-copy name while emitting an unroll-and-jam copy
        self.jam = int(jam)
        self.coop_max_lanes = int(coop_max_lanes)
        self.coop_groups = max(1, int(coop_groups))      # owners served per cooperative pass (short loops), at most
        self.lines: List[str] = []
        self.loop_counter = 0
        self.active_bvars: List[int] = []     # variables of the sample loops enclosing the code being emitted

    # --------------------------------------------------------------- references
    def ref(self, i: int, known: Dict[int, bool]) -> str:
        i = self.helper.resolve(i, known)
        n = self.nodes[i]
        if n.is_bool:
            v = self.helper.bool_value(i, known)
            if v is not None:
                return 'true' if v else 'false'
        if n.op == 'const':
            return literal(n.attr, self.ctype)
        if i in self.sch.uniform_slots and self.in_main:
            slot = self.sch.uniform_slots[i]
            return f'(tegU[{slot}] != 0)' if n.is_bool else f'tegU[{slot}]'
        if self.ref_log is not None:
            self.ref_log.add(i)
        if i in self.coop_map:
            return self.coop_map[i]
        if i in self.jam_map:
            return self.jam_map[i]
        if n.op == 'bvar':
            return f'x{i}'
        return f'v{i}'

    def _define(self, i: int):
        if self.def_log is not None:
            self.def_log.add(i)

    def expr(self, i: int, known: Dict[int, bool]) -> str:
        n = self.nodes[i]
        r = lambda a: self.ref(a, known)     # noqa: E731
        op = n.op
        m = _MATH[self.ctype]
        if op == 'add':
            return f'{r(n.args[0])} + {r(n.args[1])}'
        if op == 'mul':
            return f'{r(n.args[0])} * {r(n.args[1])}'
        if op == 'div':
            return f'{r(n.args[0])} / {r(n.args[1])}'
        if op == 'lt':
            return f'{r(n.args[0])} < {r(n.args[1])}'
        if op == 'and':
            return f'{r(n.args[0])} && {r(n.args[1])}'
        if op == 'or':
            return f'{r(n.args[0])} || {r(n.args[1])}'
        if op == 'sel':
            return f'{r(n.args[0])} ? {r(n.args[1])} : {r(n.args[2])}'
        if op in ('sqrt', 'sin', 'cos', 'asin'):
            return f'{m[op]}({r(n.args[0])})'
        if op == 'atan2':
            return f'{m[op]}({r(n.args[0])}, {r(n.args[1])})'
        if op == 'in':
            return f'a{n.attr}'
        if op == 'domis':
            code = {'T': 1, 'F': 0, 'U': 2}[n.attr]
            return f'{r(n.args[0])} == {literal(str(code), self.ctype)}'
        raise NotImplementedError(f'CUDA backend does not support instruction class "{op}"')

    def decl(self, i: int) -> str:
        return 'bool' if self.nodes[i].is_bool else 'T'

    # ------------------------------------------------------------------- blocks
    def emit(self, line: str, ind: int):
        self.lines.append('    ' * ind + line)

    def block(self, b: Block, ind: int, depth: int):
        known = b.known
        for st in b.stmts:
            if isinstance(st, Eval):
                i = st.node
                self._define(i)
                if self.nodes[i].op == 'domtri':
                    self._emit_domtri(i, known, ind)
                else:
                    self.emit(f'const {self.decl(i)} v{i} = {self.expr(i, known)};', ind)
            elif isinstance(st, IfGroup) and self._coop_ok(st, depth):
                self._coop_if(st, known, ind, depth)
            elif isinstance(st, IfGroup):
                for s in st.sels:
                    self._define(s)
                    self.emit(f'T v{s};', ind)
                self.emit(f'if ({self.ref(st.cond, known)}) {{', ind)
                self.divergent += 1
                self.block(st.then_block, ind + 1, depth)
                for s in st.sels:
                    self.emit(f'v{s} = {self.ref(self.nodes[s].args[1], st.then_block.known)};', ind + 1)
                self.emit('} else {', ind)
                self.block(st.else_block, ind + 1, depth)
                for s in st.sels:
                    self.emit(f'v{s} = {self.ref(self.nodes[s].args[2], st.else_block.known)};', ind + 1)
                self.divergent -= 1
                self.emit('}', ind)
            elif isinstance(st, LoopGroup):
                # two nested loops over ONE canonical variable would make the inner integrand read the outer sample
                # (dag.build_dag keeps such binders apart; this is the backstop)
                assert st.bvar not in self.active_bvars, 'nested integrals share a bound variable (internal error)'
                self.active_bvars.append(st.bvar)
                if self.coop and depth == 0 and not self.divergent and len(self.guard_ctx) > 1:
                    self._coop_loop(st, known, ind)
                else:
                    self.loop(st, known, ind, depth)
                self.active_bvars.pop()
            else:  # pragma: no cover
                raise TypeError(st)

    def loop(self, st: LoopGroup, known: Dict[int, bool], ind: int, depth: int, declare: bool = True):
        """data/templates/rectangular_rule.template.c:1-11, several accumulators per pass."""
        if self.mode == 'trapezoidal_quadrature':
            return self._trapezoid_loop(st, known, ind, depth, declare)
        if self.mode == 'monte_carlo_uniform':
            return self._mc_loop(st, known, ind, depth, declare)
        if self._jam_ok(st, depth):
            return self._jam_loop(st, known, ind, depth, declare)
        k = self.loop_counter
        self.loop_counter += 1
        N = self.N
        innermost = not any(isinstance(s, LoopGroup) for s in st.body.walk())
        for q in st.quads:
            self._define(q)
            if declare:
                self.emit(f'T v{q} = {literal("0", self.ctype)};', ind)
        self._define(st.bvar)
        self.emit('{', ind)
        self.emit(f'const T lo{k} = {self.ref(st.lo, known)};', ind + 1)
        self.emit(f'const T step{k} = ((T)({self.ref(st.hi, known)} - lo{k})) / (T){N};', ind + 1)
        if innermost:
            unroll = N if N <= self.unroll_inner else 4
        else:
            unroll = self.unroll_outer
        teamed = self.team > 1 and depth == 0
        # pipe balancing: fully unrolled innermost loops alternate two exact forms of the predicated accumulate
        # (see _emit_pred_add) so that neither the ALU pipe (compares) nor the issue slots are the lone limit
        mixed = (innermost and not teamed and self.pred_mix > 0 and self.use_asm and N <= self.unroll_inner
                 and self.ctype == 'float')
        zero = literal('0', self.ctype)
        if teamed:
            # the instance's team splits the outermost sample loop: chunks of TEAM consecutive samples, member r
            # evaluates sample chunk + r, then every member adds the chunk's terms in sample order (tegb_team_fold):
            # the reference's left fold, bit for bit
            self.emit('#pragma unroll 1', ind + 1)
            self.emit(f'for (unsigned int base{k} = 0; base{k} < {N}u; base{k} += {self.team}u) {{', ind + 1)
            self.emit(f'const unsigned int s{k} = base{k} + team_rank;', ind + 2)
            for q in st.quads:
                self.emit(f'T t{k}_{q}[1] = {{{zero}}};', ind + 2)
            self.emit(f'if (s{k} < {N}u) {{', ind + 2)
        elif mixed:
            self.emit('{', ind + 1)         # unrolled by hand below: one scope per sample, in sample order
        else:
            self.emit(f'#pragma unroll {unroll}', ind + 1)
            self.emit(f'for (unsigned int s{k} = 0; s{k} < {N}u; ++s{k}) {{', ind + 1)
        fixups = []

        def sample(sidx: str, form: str, ind2: int):
            self.emit(f'const T x{st.bvar} = lo{k} + step{k} * ((T){sidx} + {literal("0.5", self.ctype)});', ind2)
            self.block(st.body, ind2, depth + 1)
            for q in st.quads:
                pf = self._pred_form(self.nodes[q].args[2], st.body) if self.pred_accum else None
                if pf is not None and pf[0]:
                    conds, val = pf
                    term = f'step{k}' if val is None else f'{val} * step{k}'
                    if teamed:
                        cond = ' && '.join(f'({self.ref(c, st.body.known)})' for c in conds)
                        self.emit(f't{k}_{q}[0] = ({cond}) ? {term} : {zero};', ind2)
                    else:
                        self._emit_pred_add(f'v{q}', conds, term, st.body.known, ind2, form=form)
                    if q not in fixups:
                        fixups.append(q)
                else:
                    body = self.ref(self.nodes[q].args[2], st.body.known)
                    if teamed:
                        self.emit(f't{k}_{q}[0] = {body} * step{k};', ind2)
                    else:
                        self.emit(f'v{q} = v{q} + {body} * step{k};', ind2)

        if mixed:
            # a fraction pred_mix of the samples (evenly spread) uses the min3 form, the rest the chained form
            taken = 0
            for j in range(N):
                want = int((j + 1) * self.pred_mix + 1e-9)
                form = 'min3' if want > taken else 'chain'
                taken = want
                self.emit('{', ind + 2)
                sample(f'{j}u', form, ind + 3)
                self.emit('}', ind + 2)
        elif teamed:
            self.divergent += 1              # (members beyond the last sample sit this out: no collectives inside)
            sample(f's{k}', 'chain', ind + 3)
            self.divergent -= 1
            self.emit('}', ind + 2)
            self.emit(f'const unsigned int cnt{k} = ({N}u - base{k}) < {self.team}u ? ({N}u - base{k}) : {self.team}u;', ind + 2)
            for q in st.quads:
                self.emit(f'v{q} = tegb_team_fold<T, {self.team}, 1>(v{q}, t{k}_{q}, cnt{k});', ind + 2)
        else:
            sample(f's{k}', 'chain', ind + 2)
        self.emit('}', ind + 1)
        for q in fixups:
            # skipped terms were `0 * step`: identical for every finite step; keeps NaN/inf steps poisonous
            self.emit(f'v{q} = v{q} + {zero} * step{k};', ind + 1)
        self.emit('}', ind)

    # ------------------------------------------------------------ domain culling (cull.py)
    def _emit_domtri(self, i: int, known: Dict[int, bool], ind: int):
        """Interval evaluation of one select condition over the sampled box of its loop nest -> ``v{i}`` =
        1 (true at every sample), 0 (false at every sample) or 2 (unresolved); see ``_iv_cone``."""
        root, B = self.nodes[i].attr
        fv = self.dag.free_bvars
        zero = literal('0', self.ctype)
        pre = f'd{i}_'
        self.emit(f'T v{i};', ind)
        self.emit('{', ind)
        self.emit(f'T {pre}pz = {zero};', ind + 1)
        t, f = self._iv_cone(root, lambda j: bool(fv(j) & B), {}, known, ind + 1, pre)
        self.emit(f'v{i} = {self._iv_result(t, f, pre)};', ind + 1)
        self.emit('}', ind)

    def _iv_result(self, t: str, f: str, pre: str) -> str:
        T = self.ctype
        zero, one, two = literal('0', T), literal('1', T), literal('2', T)
        return f'({pre}pz * {zero} == {zero}) ? ({t} ? {one} : ({f} ? {zero} : {two})) : {two}'

    def _iv_cone(self, root: int, interior, leaf: Dict[int, tuple], known: Dict[int, bool], ind: int, pre: str):
        """Emit the interval evaluation of the bool node ``root``; returns the names of its (always-true,
        always-false) flags.  ``interior(j)``: node j varies over the box (everything else is a point, read through
        ``ref``); ``leaf``: interior nodes whose (lo, hi) names are already defined.

        Every bound is computed with the same individually rounded operations as the per-sample code
        (round-to-nearest is monotone).  Corner values of products / quotients / square roots and the box abscissae
        are also summed into the poison accumulator ``{pre}pz``: one non-finite term anywhere in the cone leaves
        the condition unresolved, so a NaN that ``fmin`` / ``fmax`` drop (or an infinity that could hide one
        downstream) can never support a decision.  Sums need no poison: an overflowed sum bound is +-inf on the
        side no decision can use."""
        assert self.mode == 'rectangular_quadrature', 'culling is derived for the midpoint sample placement'
        T, N = self.ctype, self.N
        half, zero = literal('0.5', T), literal('0', T)
        fmin, fmax = ('fminf', 'fmaxf') if T == 'float' else ('fmin', 'fmax')
        pz = f'{pre}pz'

        order: List[int] = []
        seen = set()
        stack = [(root, False)]
        while stack:
            j, done = stack.pop()
            if done:
                order.append(j)
                continue
            if j in seen or j in leaf or not interior(j):
                continue
            seen.add(j)
            stack.append((j, True))
            for a in self.nodes[j].args:
                stack.append((a, False))

        def lo(j: int) -> str:
            if j in leaf:
                return leaf[j][0]
            return f'{pre}l{j}' if interior(j) else self.ref(j, known)

        def hi(j: int) -> str:
            if j in leaf:
                return leaf[j][1]
            return f'{pre}h{j}' if interior(j) else self.ref(j, known)

        def is_t(j: int) -> str:
            return f'{pre}t{j}' if interior(j) else f'({self.ref(j, known)})'

        def is_f(j: int) -> str:
            return f'{pre}f{j}' if interior(j) else f'(!({self.ref(j, known)}))'

        def is_point(j: int) -> bool:
            return j not in leaf and not interior(j)

        def const_value(j: int):
            m = self.nodes[self.helper.resolve(j, known)]
            return float(m.attr) if m.op == 'const' else None

        e = lambda line: self.emit(line, ind)     # noqa: E731
        for j in order:
            n = self.nodes[j]
            op = n.op
            L, H = f'{pre}l{j}', f'{pre}h{j}'
            if op == 'bvar':
                # samples lo + step * (i + 0.5), i = 0 .. N-1, step = (hi - lo) / N: every factor is monotone
                a, b = n.args
                e(f'const T {pre}sl{j} = ((T)({lo(b)} - {hi(a)})) / (T){N}, {pre}sh{j} = ((T)({hi(b)} - {lo(a)})) / (T){N};')
                c0, c1 = f'((T)0u + {half})', f'((T){N - 1}u + {half})'
                e(f'const T {L} = {lo(a)} + {fmin}({pre}sl{j} * {c0}, {pre}sl{j} * {c1});')
                e(f'const T {H} = {hi(a)} + {fmax}({pre}sh{j} * {c0}, {pre}sh{j} * {c1});')
                e(f'{pz} = {pz} + {L}; {pz} = {pz} + {H};')
            elif op == 'add':
                x, y = n.args
                e(f'const T {L} = {lo(x)} + {lo(y)}, {H} = {hi(x)} + {hi(y)};')
            elif op in ('mul', 'div'):
                x, y = n.args
                o = '*' if op == 'mul' else '/'
                kx, ky = const_value(x), const_value(y)
                if op == 'mul' and ((kx == -1.0 and not is_point(y)) or (ky == -1.0 and not is_point(x))):
                    first = kx == -1.0 and not is_point(y)     # negation is exact: swap the bounds
                    w = y if first else x
                    neg = (lambda t: f'{self.ref(x, known)} * {t}') if first else (lambda t: f'{t} * {self.ref(y, known)}')
                    e(f'const T {L} = {neg(hi(w))}, {H} = {neg(lo(w))};')
                    continue
                if not is_point(x) and not is_point(y):
                    corners = [f'{lo(x)} {o} {lo(y)}', f'{lo(x)} {o} {hi(y)}', f'{hi(x)} {o} {lo(y)}', f'{hi(x)} {o} {hi(y)}']
                elif not is_point(x):
                    corners = [f'{lo(x)} {o} {lo(y)}', f'{hi(x)} {o} {lo(y)}']
                else:
                    corners = [f'{lo(x)} {o} {lo(y)}', f'{lo(x)} {o} {hi(y)}']
                ps = [f'{pre}p{j}_{c}' for c in range(len(corners))]
                for nm, txt in zip(ps, corners):
                    e(f'const T {nm} = {txt};')
                if len(ps) == 2:
                    e(f'const T {L} = {fmin}({ps[0]}, {ps[1]}), {H} = {fmax}({ps[0]}, {ps[1]});')
                else:
                    e(f'const T {L} = {fmin}({fmin}({ps[0]}, {ps[1]}), {fmin}({ps[2]}, {ps[3]}));')
                    e(f'const T {H} = {fmax}({fmax}({ps[0]}, {ps[1]}), {fmax}({ps[2]}, {ps[3]}));')
                e(' '.join(f'{pz} = {pz} + {nm};' for nm in ps))
                if op == 'div' and not is_point(y):
                    # a divisor interval that touches zero bounds nothing
                    e(f'if (!({lo(y)} > {zero} || {hi(y)} < {zero})) {pz} = TEGB_IV_NAN;')
            elif op == 'sqrt':
                x = n.args[0]
                fn = _MATH[T]['sqrt']
                e(f'const T {L} = {fn}({lo(x)}), {H} = {fn}({hi(x)});')
                e(f'{pz} = {pz} + {L}; {pz} = {pz} + {H};')
            elif op == 'sel':
                c, x, y = n.args
                e(f'const T {L} = {is_t(c)} ? {lo(x)} : ({is_f(c)} ? {lo(y)} : {fmin}({lo(x)}, {lo(y)}));')
                e(f'const T {H} = {is_t(c)} ? {hi(x)} : ({is_f(c)} ? {hi(y)} : {fmax}({hi(x)}, {hi(y)}));')
                e(f'if (!({is_t(c)} || {is_f(c)})) {pz} = {pz} + (({lo(x)} + {lo(y)}) + ({hi(x)} + {hi(y)}));')
            elif op == 'lt':
                x, y = n.args
                e(f'const bool {pre}t{j} = {hi(x)} < {lo(y)}, {pre}f{j} = {lo(x)} >= {hi(y)};')
            elif op == 'and':
                x, y = n.args
                e(f'const bool {pre}t{j} = {is_t(x)} && {is_t(y)}, {pre}f{j} = {is_f(x)} || {is_f(y)};')
            elif op == 'or':
                x, y = n.args
                e(f'const bool {pre}t{j} = {is_t(x)} || {is_t(y)}, {pre}f{j} = {is_f(x)} && {is_f(y)};')
            else:  # pragma: no cover  (cull.py admits interval-evaluable cones only)
                raise NotImplementedError(f'no interval rule for "{op}"')
        return is_t(root), is_f(root)

    # ---------------------------------------------------------------------- tile classification kernel
    def tiles_body(self, grid: Sequence[int], walk: bool = False, sparse_bodies=None) -> List[str]:
        """Body of ``<name>_tiles`` (cull.tile_cull): one thread per tile of BLOCK columns x ``tile_rows`` rows
        bounds the pixel-box arguments and the sample abscissae over its pixels (by walking its rows and columns:
        exact, no reasoning about rounding) and evaluates every tile condition with ``_iv_cone``."""
        T, N = self.ctype, self.N
        dag, nodes = self.dag, self.nodes
        conds = dag.tile_conds
        half, zero = literal('0.5', T), literal('0', T)
        fmin, fmax = ('fminf', 'fmaxf') if T == 'float' else ('fmin', 'fmax')
        gset = set(grid)
        saved_main = getattr(self, 'in_main', True)
        self.in_main = False
        self.lines = []
        dep: Dict[int, bool] = {}
        for j in sorted(dag.reachable([c for _, c in conds] + [nodes[c].attr[0] for k, c in conds if k == 'dom'])):
            n = nodes[j]
            dep[j] = (n.op == 'in' and n.attr in gset) or any(dep.get(a, False) for a in n.args)
        fv = dag.free_bvars
        # grid leaves: [min, max] of each pixel-box argument over the rows / columns of the tile
        in_node = {a: dag.intern('in', (), a) for a in grid}
        xl, xu, yl, yu = (in_node[a] for a in grid)
        leaf: Dict[int, tuple] = {xl: ('g_xl_lo', 'g_xl_hi'), xu: ('g_xu_lo', 'g_xu_hi'),
                                  yl: ('g_yl_lo', 'g_yl_hi'), yu: ('g_yu_lo', 'g_yu_hi')}
        # integration variables whose bounds are exactly a pixel-box pair: exact range of their abscissae
        roots = [(k, c, nodes[c].attr[0] if k == 'dom' else c, nodes[c].attr[1] if k == 'dom' else frozenset())
                 for k, c in conds]
        bx, by = [], []
        for _, _, r, B in roots:
            for j in dag.reachable([r]):
                if nodes[j].op == 'bvar' and j in B:
                    pair = tuple(self.helper.resolve(a, {}) for a in nodes[j].args)
                    if pair == (xl, xu) and j not in bx:
                        bx.append(j)
                    elif pair == (yl, yu) and j not in by:
                        by.append(j)
        for j in bx + by:
            leaf[j] = (f'g_b{j}_lo', f'g_b{j}_hi')
        e = lambda line, ind=1: self.emit(line, ind)     # noqa: E731

        def points():
            # launch-invariant values the cones read (points), from the scalar arguments
            need: List[int] = []
            seen = set()
            for _, _, r, B in roots:
                interior = (lambda j, B=B: bool(fv(j) & B) or dep.get(j, False))
                stack = [r]
                visited = set()
                while stack:
                    j = stack.pop()
                    if j in visited:
                        continue
                    visited.add(j)
                    if j in leaf:
                        continue
                    if interior(j):
                        stack.extend(nodes[j].args)
                        continue
                    sub = [(j, False)]          # a point: emit its whole (launch-invariant) cone
                    while sub:
                        m, done = sub.pop()
                        if done:
                            need.append(m)
                            continue
                        if m in seen or nodes[m].op == 'const':
                            continue
                        seen.add(m)
                        sub.append((m, True))
                        for a2 in reversed(nodes[m].args):
                            sub.append((a2, False))
            for m in need:
                e(f'const {self.decl(m)} v{m} = {self.expr(m, {})};')

        def cones(ind: int, index: str, guard: str):
            e('bool unresolved = false;', ind)
            for k, (kind, c, r, B) in enumerate(roots):
                pre = f'k{k}_'
                e(f'T {pre}tri;', ind)
                e('{', ind)
                e(f'T {pre}pz = g_pz;', ind + 1)
                t, f = self._iv_cone(r, (lambda j, B=B: bool(fv(j) & B) or dep.get(j, False)), leaf, {}, ind + 1, pre)
                e(f'{pre}tri = {self._iv_result(t, f, pre)};', ind + 1)
                e(f'{guard}tiles[(size_t){k} * n_tiles + {index}] = {pre}tri;', ind + 1)
                e(f'unresolved = unresolved || {pre}tri == {literal("2", T)};', ind + 1)
                e('}', ind)

        if walk:
            # grouped launches (tiles are relative to each group's window): one THREAD per tile walks the tile's rows
            # [r0, r1) and columns [c0, c1) itself -- exact ranges of the box edges and the midpoint abscissae, no
            # reasoning about rounding -- and evaluates the cones once (a warp per tile spent 32 lanes on one answer)
            e(f'T g_pz = {zero};')
            for tab, first, last, lo_n, hi_n, bvs in (('xe', 'r0', 'r1', 'xl', 'xu', bx), ('ye', 'c0', 'c1', 'yl', 'yu', by)):
                e(f'T g_{lo_n}_lo = {tab}[{first}], g_{lo_n}_hi = {tab}[{first}], g_{hi_n}_lo = {tab}[{first} + 1], '
                  f'g_{hi_n}_hi = {tab}[{first} + 1];')
                e(f'T g_s{lo_n}_lo = {tab}[{first}], g_s{lo_n}_hi = {tab}[{first}];')
                e(f'{{ T lb = {tab}[{first}];')
                e(f'for (int q = {first}; q < {last}; ++q) {{')
                e(f'const T ub = {tab}[q + 1];', 2)
                e(f'g_{lo_n}_lo = {fmin}(g_{lo_n}_lo, lb); g_{lo_n}_hi = {fmax}(g_{lo_n}_hi, lb);', 2)
                e(f'g_{hi_n}_lo = {fmin}(g_{hi_n}_lo, ub); g_{hi_n}_hi = {fmax}(g_{hi_n}_hi, ub);', 2)
                e(f'const T st = ((T)(ub - lb)) / (T){N};', 2)
                e(f'const T sa = lb + st * ((T)0u + {half}), sb = lb + st * ((T){N - 1}u + {half});', 2)
                e(f'g_s{lo_n}_lo = {fmin}(g_s{lo_n}_lo, {fmin}(sa, sb)); g_s{lo_n}_hi = {fmax}(g_s{lo_n}_hi, {fmax}(sa, sb));', 2)
                e('g_pz = g_pz + lb; g_pz = g_pz + ub; g_pz = g_pz + (ub - lb); g_pz = g_pz + sa; g_pz = g_pz + sb;', 2)
                e('lb = ub;', 2)
                e('} }')
                for j in bvs:
                    e(f'const T g_b{j}_lo = g_s{lo_n}_lo, g_b{j}_hi = g_s{lo_n}_hi;')
            points()
            cones(1, 't', '')
            e('(void)unresolved;')
        else:
            # [min, max] of the box edges and of the sample abscissae over a tile row / tile column come from the
            # runtime's tile_ranges kernel (exact walk over the rows / columns, cached per grid): 7 values per entry.
            # One thread per BLOCK of the main kernel: its TEGB_TILE_SUB tiles (tile_cols columns each, one or more
            # warps) are classified in turn, and the block goes to the front of the launch order if any is unresolved.
            e('const T* __restrict__ rx = ranges + (size_t)ty * 7;')
            e('const T g_xl_lo = rx[0], g_xl_hi = rx[1], g_xu_lo = rx[2], g_xu_hi = rx[3];')
            for j in bx:
                e(f'const T g_b{j}_lo = rx[4], g_b{j}_hi = rx[5];')
            points()
            # one thread per tile; the TEGB_TILE_SUB tiles of a block sit in adjacent lanes
            e('const int txs_ = tx * TEGB_TILE_SUB + sub;')
            e('const bool dead = !active || txs_ >= n_txs || txs_ * TEGB_TILE_COLS >= res_y;       // no pixel in this tile')
            e('const int txs = dead ? 0 : txs_;')
            e('const T* __restrict__ ry = ranges + ((size_t)n_ty + txs) * 7;')
            e('const T g_yl_lo = ry[0], g_yl_hi = ry[1], g_yu_lo = ry[2], g_yu_hi = ry[3];')
            for j in by:
                e(f'const T g_b{j}_lo = ry[4], g_b{j}_hi = ry[5];')
            e('const T g_pz = rx[6] + ry[6];')
            cones(1, '(size_t)ty * n_txs + txs', 'if (!dead) ')
            if sparse_bodies is not None:
                # Programs whose outputs are all summed, one partial sum per tile: a tile whose body is "every output
                # is +0" (or that has no pixel) gets its zero partials HERE, unresolved tiles go to the queue, and only
                # blocks with a tile that is resolved to something else are listed for the main kernel (front of the
                # launch order, counters[0] of them): the main kernel launches no block for tiles without work.
                T_ = self.ctype
                e(f'int body_ = {len(sparse_bodies)};')
                for bi, (match, zeros) in reversed(list(enumerate(sparse_bodies))):
                    terms = []
                    for k2 in sorted(match):
                        alts = ' || '.join(f'k{k2}_tri == {literal(str(v), T_)}' for v in match[k2])
                        terms.append(f'({alts})')
                    e(f'if ({" && ".join(terms)}) body_ = {bi};')
                zero_ids = [bi for bi, (_, zeros) in enumerate(sparse_bodies) if zeros]
                zexpr = ' || '.join(f'body_ == {bi}' for bi in zero_ids) if zero_ids else 'false'
                e(f'const bool no_work = dead || ({zexpr});')
                e('if (active && txs_ < n_txs && no_work) {')
                for k in range(len(self.sch.outputs)):
                    e(f'partials[(size_t){k} * n_tiles + (size_t)ty * n_txs + txs_] = {zero};', 2)
                e('}')
                e('int any_unresolved = (!no_work && !unresolved) ? 1 : 0;        // (here: "a tile the main kernel must visit")')
            else:
                # launch order of the main kernel: blocks with an unresolved tile from the front, the others from the back
                e('int any_unresolved = (!dead && unresolved) ? 1 : 0;')
            e('for (int off = 1; off < TEGB_TILE_SUB; off <<= 1) any_unresolved |= __shfl_xor_sync(0xffffffffu, any_unresolved, off);')
            # the block's slot in the launch order, and its entry: (tile row << 16) | tile column
            e('int pos = 0;')
            e('if (active && sub == 0) {')
            e('pos = any_unresolved ? atomicAdd(&counters[0], 1) : n_tx * n_ty - 1 - atomicAdd(&counters[1], 1);', 2)
            e('order[pos] = (int)(((unsigned int)ty << 16) | (unsigned int)tx);', 2)
            e('}')
            e('pos = __shfl_sync(0xffffffffu, pos, (int)((threadIdx.x & 31u) & ~(unsigned int)(TEGB_TILE_SUB - 1)));')
            # the same tri-states once more, in LAUNCH order: a block of the main kernel reads its order entry and its
            # tiles' states with independent loads (tiles without a pixel read as resolved-false)
            for k in range(len(roots)):
                e(f'if (active) tiles_ord[(size_t){k} * n_tiles + (size_t)pos * TEGB_TILE_SUB + sub] = dead ? {zero} : k{k}_tri;')
            # unresolved tiles of image programs are not walked by the block that owns them: their rows go on a queue
            # that dedicated blocks of the main kernel serve (ulist: the tiles, in any order)
            e('const bool mine_unresolved = active && !dead && unresolved;')
            e('if (mine_unresolved) ulist[atomicAdd(&counters[2], 1)] = (int)(((unsigned int)ty << 16) | (unsigned int)txs);')
        self.in_main = saved_main
        return self.lines

    def _trapezoid_loop(self, st: LoopGroup, known: Dict[int, bool], ind: int, depth: int, declare: bool = True):
        """The numpy backend's rule (teg/eval/numpy_eval.py:57-68), the semantics of the reference's
        ``trapezoidal_quadrature`` mode (to_c.py:389-393 names a template that is not shipped):
        ``x, step = np.linspace(lo, hi, N, retstep=True)``, i.e. ``x_i = i * step + lo`` with ``step = (hi - lo) /
        (N - 1)`` and the last point set to ``hi`` exactly; ``(1 if lo < hi else -1) * step * sum(y[:-1] + y[1:]) / 2``
        with the pair sums folded left to right."""
        k = self.loop_counter
        self.loop_counter += 1
        N, T = self.N, self.ctype
        zero = literal('0', T)
        innermost = not any(isinstance(s, LoopGroup) for s in st.body.walk())
        for q in st.quads:
            self._define(q)
            if declare:
                self.emit(f'T v{q} = {zero};', ind)
        self._define(st.bvar)
        self.emit('{', ind)
        self.emit(f'const T lo{k} = {self.ref(st.lo, known)};', ind + 1)
        self.emit(f'const T hi{k} = {self.ref(st.hi, known)};', ind + 1)
        self.emit(f'const T step{k} = (hi{k} - lo{k}) / (T){N - 1};', ind + 1)
        for q in st.quads:
            self.emit(f'T prev{k}_{q} = {zero};', ind + 1)
            if not declare:
                self.emit(f'v{q} = {zero};', ind + 1)
        unroll = N if (innermost and N <= self.unroll_inner) else 1
        self.emit(f'#pragma unroll {unroll}', ind + 1)
        self.emit(f'for (unsigned int s{k} = 0; s{k} < {N}u; ++s{k}) {{', ind + 1)
        self.emit(f'const T x{st.bvar} = (s{k} == {N - 1}u) ? hi{k} : (T)s{k} * step{k} + lo{k};', ind + 2)
        self.block(st.body, ind + 2, depth + 1)
        for q in st.quads:
            body = self.ref(self.nodes[q].args[2], st.body.known)
            self.emit(f'const T f{k}_{q} = {body};', ind + 2)
            self.emit(f'if (s{k} > 0u) v{q} = v{q} + (prev{k}_{q} + f{k}_{q});', ind + 2)
            self.emit(f'prev{k}_{q} = f{k}_{q};', ind + 2)
        self.emit('}', ind + 1)
        one = literal('1', T)
        for q in st.quads:
            self.emit(f'v{q} = ((((lo{k} < hi{k}) ? {one} : -{one}) * step{k}) * v{q}) / {literal("2", T)};', ind + 1)
        self.emit('}', ind)

    def _mc_loop(self, st: LoopGroup, known: Dict[int, bool], ind: int, depth: int, declare: bool = True):
        """``monte_carlo_uniform``: the reference names the mode (to_c.py:389-393) but ships neither a template nor a
        definition; the semantics are defined here and restated by the oracle (oracle/teg_oracle_impl.h, case 'T'):
        ``u_i = Philox4x32-10(key = seed; counter = (instance lo, instance hi, depth, i))`` mapped to [0, 1),
        ``x_i = lb + (ub - lb) * u_i``, ``step = (ub - lb) / N``, ``out = out + f(x_i) * step`` folded left to right.
        The points depend on the instance, the integral's nesting depth in the program and the sample index only, so
        integrals of the same depth share their points (common random numbers) and loop fusion / hoisting stay exact.
        Reproducible: the same (seed, instance, program) gives the same bits on any launch geometry."""
        k = self.loop_counter
        self.loop_counter += 1
        N, T = self.N, self.ctype
        tree_depth = int(self.nodes[st.bvar].attr)
        for q in st.quads:
            self._define(q)
            if declare:
                self.emit(f'T v{q} = {literal("0", T)};', ind)
        self._define(st.bvar)
        self.emit('{', ind)
        self.emit(f'const T lo{k} = {self.ref(st.lo, known)};', ind + 1)
        self.emit(f'const T width{k} = {self.ref(st.hi, known)} - lo{k};', ind + 1)
        self.emit(f'const T step{k} = ((T)width{k}) / (T){N};', ind + 1)
        self.emit('#pragma unroll 1', ind + 1)
        self.emit(f'for (unsigned int s{k} = 0; s{k} < {N}u; ++s{k}) {{', ind + 1)
        self.emit(f'const T x{st.bvar} = lo{k} + width{k} * tegb_uniform(TEGB_MC_SEED, tegb_inst, {tree_depth}u, s{k});', ind + 2)
        self.block(st.body, ind + 2, depth + 1)
        for q in st.quads:
            body = self.ref(self.nodes[q].args[2], st.body.known)
            self.emit(f'v{q} = v{q} + {body} * step{k};', ind + 2)
        self.emit('}', ind + 1)
        self.emit('}', ind)

    # ------------------------------------------------------------ outer-loop blocking (unroll and jam)
    # Nest  for x { pre(x); for y { body(x, y) }; post }  with a long inner loop: U consecutive x-samples are
    # carried through ONE pass of the inner loop, so everything that depends on y only (sample placement, the
    # y-terms of every affine form) is computed once per U samples instead of once per sample.  Each of the U
    # inner accumulators still receives its own samples in order and the outer accumulator receives the U
    # inner results in sample order: every sum is the same left fold, bit for bit.  (For short inner loops
    # the full unroll lets ptxas hoist the same terms by itself.)
    def _jam_ok(self, st: LoopGroup, depth: int) -> bool:
        U, N = self.jam, self.N
        if U < 2 or N <= self.unroll_inner:
            return False
        teamed = self.team > 1 and depth == 0
        if N % (U * (self.team if teamed else 1)) != 0:
            return False
        inner = [s for s in st.body.stmts if isinstance(s, LoopGroup)]
        if len(inner) != 1 or any(isinstance(s, IfGroup) for s in st.body.stmts):
            return False
        inner = inner[0]
        if any(not isinstance(s, Eval) for s in inner.body.stmts):
            return False
        fv = self.dag.free_bvars
        if st.bvar in fv(inner.lo) or st.bvar in fv(inner.hi):
            return False
        # the inner body must have something to share
        return any(st.bvar not in fv(s.node) for s in inner.body.stmts)

    def _jam_loop(self, st: LoopGroup, known: Dict[int, bool], ind: int, depth: int, declare: bool = True):
        U, N, T = self.jam, self.N, self.ctype
        zero, half = literal('0', T), literal('0.5', T)
        k = self.loop_counter
        self.loop_counter += 1
        teamed = self.team > 1 and depth == 0
        inner = [s for s in st.body.stmts if isinstance(s, LoopGroup)][0]
        assert inner.bvar != st.bvar and inner.bvar not in self.active_bvars, 'nested integrals share a bound variable'
        pos = st.body.stmts.index(inner)
        pre, post = st.body.stmts[:pos], st.body.stmts[pos + 1:]
        fv = self.dag.free_bvars
        shared = [s for s in inner.body.stmts if st.bvar not in fv(s.node)]
        percopy = [s for s in inner.body.stmts if st.bvar in fv(s.node)]
        for q in st.quads:
            self._define(q)
            if declare:
                self.emit(f'T v{q} = {zero};', ind)
        self._define(st.bvar)
        self.emit('{', ind)
        self.emit(f'const T lo{k} = {self.ref(st.lo, known)};', ind + 1)
        self.emit(f'const T step{k} = ((T)({self.ref(st.hi, known)} - lo{k})) / (T){N};', ind + 1)
        self.emit('#pragma unroll 1', ind + 1)
        if teamed:
            # chunks of TEAM * U consecutive samples: member r carries the U samples chunk + r * U + u through one pass
            # of the inner loop; the chunk's TEAM * U terms are then added in sample order by every member
            self.emit(f'for (unsigned int base{k} = 0; base{k} < {N}u; base{k} += {self.team * U}u) {{', ind + 1)
            self.emit(f'const unsigned int s{k} = base{k} + team_rank * {U}u;', ind + 2)
            for q in st.quads:
                self.emit(f'T t{k}_{q}[{U}];', ind + 2)
        else:
            self.emit(f'for (unsigned int s{k} = 0; s{k} < {N}u; s{k} += {U}u) {{', ind + 1)
        copy_nodes = [st.bvar] + [s.node for s in pre] + list(inner.quads) + [s.node for s in post]
        maps = [{i: (f'x{i}_{u}' if self.nodes[i].op == 'bvar' else f'v{i}_{u}') for i in copy_nodes} for u in range(U)]
        saved = dict(self.jam_map)
        kk = self.loop_counter
        self.loop_counter += 1
        # per-copy outer prologue
        for u in range(U):
            self.jam_map = {**saved, **maps[u]}
            self.emit(f'const T x{st.bvar}_{u} = lo{k} + step{k} * ((T)(s{k} + {u}u) + {half});', ind + 2)
            for s in pre:
                self._define(s.node)
                self.emit(f'const {self.decl(s.node)} v{s.node}_{u} = {self.expr(s.node, st.body.known)};', ind + 2)
            for q in inner.quads:
                self._define(q)
                self.emit(f'T v{q}_{u} = {zero};', ind + 2)
        # the one shared pass over the inner loop
        self.jam_map = dict(saved)
        self._define(inner.bvar)
        self.emit('{', ind + 2)
        self.emit(f'const T lo{kk} = {self.ref(inner.lo, st.body.known)};', ind + 3)
        self.emit(f'const T step{kk} = ((T)({self.ref(inner.hi, st.body.known)} - lo{kk})) / (T){N};', ind + 3)
        self.emit('#pragma unroll 2', ind + 3)
        self.emit(f'for (unsigned int s{kk} = 0; s{kk} < {N}u; ++s{kk}) {{', ind + 3)
        self.emit(f'const T x{inner.bvar} = lo{kk} + step{kk} * ((T)s{kk} + {half});', ind + 4)
        for s in shared:
            self._define(s.node)
            self.emit(f'const {self.decl(s.node)} v{s.node} = {self.expr(s.node, inner.body.known)};', ind + 4)
        for u in range(U):
            self.jam_map = {**saved, **maps[u]}
            self.emit('{', ind + 4)
            for s in percopy:
                self._define(s.node)
                self.emit(f'const {self.decl(s.node)} v{s.node} = {self.expr(s.node, inner.body.known)};', ind + 5)
            for q in inner.quads:
                body = self.ref(self.nodes[q].args[2], inner.body.known)
                self.emit(f'v{q}_{u} = v{q}_{u} + {body} * step{kk};', ind + 5)
            self.emit('}', ind + 4)
        self.emit('}', ind + 3)
        self.emit('}', ind + 2)
        # per-copy epilogue and the outer accumulate, in sample order
        for u in range(U):
            self.jam_map = {**saved, **maps[u]}
            for s in post:
                self._define(s.node)
                self.emit(f'const {self.decl(s.node)} v{s.node}_{u} = {self.expr(s.node, st.body.known)};', ind + 2)
            for q in st.quads:
                body = self.ref(self.nodes[q].args[2], st.body.known)
                if teamed:
                    self.emit(f't{k}_{q}[{u}] = {body} * step{k};', ind + 2)
                else:
                    self.emit(f'v{q} = v{q} + {body} * step{k};', ind + 2)
        self.jam_map = saved
        if teamed:
            for q in st.quads:
                self.emit(f'v{q} = tegb_team_fold<T, {self.team}, {U}>(v{q}, t{k}_{q}, {self.team * U}u);', ind + 2)
        self.emit('}', ind + 1)
        self.emit('}', ind)

    # ------------------------------------------------------ warp-cooperative guarded loops
    # Reverse-mode programs guard every delta integral with a bound check (teg/passes/delta.py:426-439): a
    # fraction of a percent of the instances run a loop, scattered along the discontinuities, so a warp that
    # meets one spends N iterations with one or two useful lanes.  Instead: the whole warp enters the guard when
    # ANY lane passes it (everything but the loops is evaluated speculatively -- values are pure), and each
    # pending loop is executed by the warp together: the owner lane's loop inputs are broadcast, lane s
    # evaluates sample s, and every lane folds the N terms in sample order (the reference's left fold, bit for
    # bit); the owner keeps the result.
    @staticmethod
    def _has_loop(b: Block) -> bool:
        return any(isinstance(s, LoopGroup) for s in b.walk())

    def _coop_ok(self, st: IfGroup, depth: int) -> bool:
        if not (self.coop and depth == 0 and not self.divergent and self.in_main):
            return False
        if not self._has_loop(st.then_block):
            return False
        # a straight-line else side is evaluated unconditionally (speculation); one with structure of its own is
        # accepted only for the culling guards (cull.py), whose else side is the cheap specialised nest
        if not any(isinstance(s, (LoopGroup, IfGroup)) for s in st.else_block.walk()):
            return True
        return self._is_cull_guard(st.cond)

    def _is_cull_guard(self, c: int) -> bool:
        n = self.nodes[c]
        return n.op == 'domis' or (n.op == 'or' and all(self._is_cull_guard(a) for a in n.args))

    def _coop_if(self, st: IfGroup, known: Dict[int, bool], ind: int, depth: int):
        self.uses_coop = True
        cond = self.ref(st.cond, known)
        if any(isinstance(s, (LoopGroup, IfGroup)) for s in st.else_block.walk()):
            for s in st.sels:
                self._define(s)
                self.emit(f'T v{s} = {literal("0", self.ctype)};', ind)
            self.emit(f'if (!({cond})) {{', ind)
            self.divergent += 1
            self.block(st.else_block, ind + 1, depth)
            self.divergent -= 1
            for s in st.sels:
                self.emit(f'v{s} = {self.ref(self.nodes[s].args[2], st.else_block.known)};', ind + 1)
            self.emit('}', ind)
        else:
            self.block(st.else_block, ind, depth)
            for s in st.sels:
                self._define(s)
                self.emit(f'T v{s} = {self.ref(self.nodes[s].args[2], st.else_block.known)};', ind)
        ctx = ' && '.join(self.guard_ctx + [f'({cond})'])
        self.emit(f'if (__any_sync(0xffffffffu, {ctx})) {{', ind)
        self.guard_ctx.append(f'({cond})')
        self.block(st.then_block, ind + 1, depth)
        self.guard_ctx.pop()
        for s in st.sels:
            self.emit(f'if ({cond}) v{s} = {self.ref(self.nodes[s].args[1], st.then_block.known)};', ind + 1)
        self.emit('}', ind)

    def _coop_loop(self, st: LoopGroup, known: Dict[int, bool], ind: int):
        self.uses_coop = True
        N = self.N
        T = self.ctype
        zero = literal('0', T)
        # dry run: which per-lane values does the loop (bounds + body + integrands) read from outside?
        saved = (self.lines, self.loop_counter, self.ref_log, self.def_log)
        self.lines, self.ref_log, self.def_log = [], set(), set()
        k = self.loop_counter
        self.loop_counter += 1
        self._coop_loop_body(st, known, 0, dry=True, k=k)
        inputs = sorted(self.ref_log - self.def_log)
        self.lines, _, self.ref_log, self.def_log = saved
        self.loop_counter = k + 1
        if self.ref_log is not None:
            self.ref_log.update(inputs)
        for q in st.quads:
            self._define(q)
            self.emit(f'T v{q} = {zero};', ind)
        self.emit('{', ind)
        guard = ' && '.join(self.guard_ctx)
        self.emit(f'unsigned int coop_m{k} = __ballot_sync(0xffffffffu, {guard});', ind + 1)
        # Short loops leave most of the warp idle when one owner is served at a time (N = 16: half of the lanes), so a
        # pass serves G owners at once: the warp splits into G groups of W lanes (W = N rounded up to a power of two),
        # group j takes the j-th pending owner, lane % W is the sample.  Every group folds its own N terms in sample
        # order: the same left fold, bit for bit.
        W = 32
        if N <= 16:
            W = 1
            while W < N:
                W *= 2
        G = min(32 // W, self.coop_groups)
        if G == 1:
            W = 32
        # Many pending lanes: the ordinary per-lane loop (all of them busy at once) is cheaper than serving
        # them pass by pass.  Few (the usual case along a discontinuity): serve them with the whole warp.
        self.emit(f'if (__popc(coop_m{k}) > {self.coop_max_lanes * G}) {{', ind + 1)
        self.emit(f'if ({guard}) {{', ind + 2)
        saved_div, saved_k = self.divergent, self.loop_counter
        self.divergent += 1
        self.loop(st, known, ind + 3, 0, declare=False)
        self.divergent = saved_div
        self.emit('}', ind + 2)
        self.emit(f'}} else while (coop_m{k}) {{', ind + 1)
        if G == 1:
            self.emit(f'const int coop_src{k} = __ffs(coop_m{k}) - 1;', ind + 2)
            self.emit(f'coop_m{k} &= coop_m{k} - 1;', ind + 2)
        else:
            for j in range(G):
                self.emit(f'const int coop_src{k}_{j} = coop_m{k} ? __ffs(coop_m{k}) - 1 : -1;', ind + 2)
                self.emit(f'coop_m{k} &= coop_m{k} - 1;', ind + 2)         # (0 & anything stays 0)
            pick = f'coop_src{k}_{G - 1}'
            for j in reversed(range(G - 1)):
                pick = f'(coop_grp{k} == {j} ? coop_src{k}_{j} : {pick})'
            self.emit(f'const int coop_grp{k} = (int)((threadIdx.x & 31u) / {W}u);', ind + 2)
            self.emit(f'const int coop_own{k} = {pick};', ind + 2)
            self.emit(f'const bool coop_has{k} = coop_own{k} >= 0;                  // this group has an owner to serve', ind + 2)
            self.emit(f'const int coop_src{k} = coop_has{k} ? coop_own{k} : coop_src{k}_0;', ind + 2)
        old_map = dict(self.coop_map)
        for i in inputs:
            n = self.nodes[i]
            cur = self.ref(i, {})           # current per-lane name (possibly an outer broadcast copy)
            name = f'c{k}_{i}'
            if n.is_bool:
                self.emit(f'const bool {name} = __shfl_sync(0xffffffffu, (int){cur}, coop_src{k}) != 0;', ind + 2)
            else:
                self.emit(f'const T {name} = __shfl_sync(0xffffffffu, {cur}, coop_src{k});', ind + 2)
        for i in inputs:
            self.coop_map[i] = f'c{k}_{i}'
        self._coop_loop_body(st, known, ind + 2, dry=False, k=k, width=W, groups=G)
        self.coop_map = old_map
        for q in st.quads:
            if G == 1:
                self.emit(f'if ((int)(threadIdx.x & 31u) == coop_src{k}) v{q} = acc{k}_{q};', ind + 2)
            else:
                for j in range(G):
                    self.emit(f'{{ const T r_ = __shfl_sync(0xffffffffu, acc{k}_{q}, {j * W}); '
                              f'if ((int)(threadIdx.x & 31u) == coop_src{k}_{j}) v{q} = r_; }}', ind + 2)
        self.emit('}', ind + 1)
        self.emit('}', ind)

    def _coop_loop_body(self, st: LoopGroup, known: Dict[int, bool], ind: int, dry: bool, k: int, width: int = 32,
                        groups: int = 1):
        """One pending loop per group of ``width`` lanes, executed for the owner lane's (broadcast) inputs."""
        N = self.N
        lane = '(threadIdx.x & 31u)' if width == 32 else f'((threadIdx.x & 31u) % {width}u)'
        first = '0' if width == 32 else f'((threadIdx.x & 31u) & ~{width - 1}u)'
        has = f'coop_has{k} && ' if groups > 1 else ''
        T = self.ctype
        zero = literal('0', T)
        self._define(st.bvar)
        self.emit(f'const T lo{k} = {self.ref(st.lo, known)};', ind)
        self.emit(f'const T step{k} = ((T)({self.ref(st.hi, known)} - lo{k})) / (T){N};', ind)
        for q in st.quads:
            self.emit(f'T acc{k}_{q} = {zero};', ind)
        self.emit(f'for (unsigned int base{k} = 0; base{k} < {N}u; base{k} += {width}u) {{', ind)
        self.emit(f'const unsigned int s{k} = base{k} + {lane};', ind + 1)
        self.emit(f'const T x{st.bvar} = lo{k} + step{k} * ((T)s{k} + {literal("0.5", T)});', ind + 1)
        saved_div = self.divergent
        self.divergent += 1                     # nothing inside the per-sample body may use warp collectives
        self.block(st.body, ind + 1, 1)
        self.divergent = saved_div
        fix = []
        for q in st.quads:
            pf = self._pred_form(self.nodes[q].args[2], st.body) if self.pred_accum else None
            if pf is not None and pf[0]:
                conds, val = pf
                term = f'step{k}' if val is None else f'{val} * step{k}'
                cond = ' && '.join([f'{has}s{k} < {N}u'] + [f'({self.ref(c, st.body.known)})' for c in conds])
                self.emit(f'const T t{k}_{q} = ({cond}) ? {term} : {zero};', ind + 1)
                fix.append(q)
            else:
                body = self.ref(self.nodes[q].args[2], st.body.known)
                self.emit(f'const T t{k}_{q} = ({has}s{k} < {N}u) ? {body} * step{k} : {zero};', ind + 1)
        # left fold in sample order, computed identically by every lane (adding the +0 of a missing /
        # skipped sample never changes the accumulator)
        if N <= width:
            # one pass covers the loop: the fold has a compile-time length (unrolled: independent shuffles, one add chain)
            self.emit(f'#pragma unroll', ind + 1)
            self.emit(f'for (unsigned int l{k} = 0; l{k} < {N}u; ++l{k}) {{', ind + 1)
        else:
            self.emit(f'const unsigned int cnt{k} = ({N}u - base{k}) < {width}u ? ({N}u - base{k}) : {width}u;', ind + 1)
            self.emit(f'for (unsigned int l{k} = 0; l{k} < cnt{k}; ++l{k}) {{', ind + 1)
        for q in st.quads:
            self.emit(f'acc{k}_{q} = acc{k}_{q} + __shfl_sync(0xffffffffu, t{k}_{q}, {first} + l{k});', ind + 2)
        self.emit('}', ind + 1)
        self.emit('}', ind)
        for q in fix:
            self.emit(f'acc{k}_{q} = acc{k}_{q} + {zero} * step{k};', ind)

    # ------------------------------------------------------- predicated accumulate (ALU-pipe relief)
    def _fold_compare(self, c: int, known: Dict[int, bool]):
        """lt node -> (ptx cmp, lhs text, rhs text).  ``0 < a + b`` is emitted as ``a > -b`` (exact in IEEE
        arithmetic without flush-to-zero: a float sum is positive iff the exact sum is), which drops the add
        from the per-sample chain; ``-1 * w`` operands are negated by swapping sides instead."""
        n = self.nodes[c]
        x = self.helper.resolve(n.args[0], known)
        y = self.helper.resolve(n.args[1], known)

        def is_zero(i):
            m = self.nodes[i]
            return m.op == 'const' and float(m.attr) == 0.0

        def neg_of(i):
            """text of -(node i) when that is free, else None"""
            m = self.nodes[i]
            if m.op == 'mul':
                for ka, wa in ((m.args[0], m.args[1]), (m.args[1], m.args[0])):
                    kk = self.nodes[self.helper.resolve(ka, known)]
                    if kk.op == 'const' and float(kk.attr) == -1.0:
                        return self.ref(wa, known)
            return None

        for zero, summ, cmp in ((x, y, 'gt'), (y, x, 'lt')):
            if is_zero(zero) and self.nodes[summ].op == 'add' and summ not in self.sch.uniform_slots:
                a, b = (self.helper.resolve(t, known) for t in self.nodes[summ].args)
                nb = neg_of(b)
                if nb is not None:
                    return cmp, self.ref(a, known), nb
                na = neg_of(a)
                if na is not None:
                    return cmp, self.ref(b, known), na
                return cmp, self.ref(a, known), f'(-{self.ref(b, known)})'
        return 'lt', self.ref(x, known), self.ref(y, known)

    def _emit_pred_add(self, acc: str, conds: List[int], term: str, known: Dict[int, bool], ind: int,
                       form: str = 'chain'):
        """acc = acc + term  iff all conds: one chained FSETP per comparison, then one predicated add."""
        leaves: List[int] = []

        def flatten(c: int):
            c_res = c
            v = self.helper.bool_value(c_res, known)
            if v is True:
                return
            n = self.nodes[c_res]
            if n.op == 'and' and c_res not in self.sch.uniform_slots:
                flatten(n.args[0])
                flatten(n.args[1])
            else:
                leaves.append(c_res)

        for c in conds:
            flatten(c)
        if not self.use_asm or not leaves or len(leaves) > 12:
            cond = ' && '.join(self.ref(c, known) for c in conds)
            self.emit(f'if ({cond}) {acc} = {acc} + {term};', ind)
            return
        sfx, cons = ('f32', 'f') if self.ctype == 'float' else ('f64', 'd')
        all_lt = all(self.nodes[c].op == 'lt' and c not in self.sch.uniform_slots and
                     self.helper.bool_value(c, known) is None for c in leaves)
        if form == 'min3' and len(leaves) == 3 and all_lt and self.ctype == 'float':
            # the same predicate through the FMA pipe: d_k = lhs_k - rhs_k (positive iff lhs_k > rhs_k, exactly),
            # one 3-input NaN-propagating minimum (FMNMX3, new on sm_100), one compare.  2 ALU-pipe instructions
            # instead of 3; alternated with the chained form the two pipes are evenly loaded.
            ops2: List[str] = []
            subs = []
            for j, c in enumerate(leaves):
                cmp, lhs, rhs = self._fold_compare(c, known)
                a, b = (lhs, rhs) if cmp == 'gt' else (rhs, lhs)
                i0 = len(ops2) + 2
                ops2 += [f'"f"({a})', f'"f"({b})']
                subs.append(f'sub.rn.f32 e{j}, %{i0}, %{i0 + 1};')
            body = ('.reg .pred p; .reg .f32 e0, e1, e2, m; ' + ' '.join(subs) +
                    ' min.NaN.f32 m, e0, e1, e2; setp.gt.f32 p, m, 0f00000000; @p add.rn.f32 %0, %0, %1;')
            self.emit('#ifdef TEGB_HAVE_MIN3', 0)
            self.emit(f'asm("{{ {body} }}" : "+f"({acc}) : "f"({term}), {", ".join(ops2)});', ind)
            self.emit('#else', 0)
            self._emit_pred_add(acc, conds, term, known, ind, form='chain')
            self.emit('#endif', 0)
            return
        ops: List[str] = []
        lines: List[str] = ['.reg .pred p;']
        first = True
        for c in leaves:
            n = self.nodes[c]
            if n.op == 'lt' and c not in self.sch.uniform_slots and self.helper.bool_value(c, known) is None:
                cmp, lhs, rhs = self._fold_compare(c, known)
                i0 = len(ops) + 2
                ops += [f'"{cons}"({lhs})', f'"{cons}"({rhs})']
                lines.append(f'setp.{cmp}{"" if first else ".and"}.{sfx} p, %{i0}, %{i0 + 1}{"" if first else ", p"};')
            else:
                i0 = len(ops) + 2
                ops.append(f'"r"((int)({self.ref(c, known)}))')
                lines.append(f'setp.ne{"" if first else ".and"}.b32 p, %{i0}, 0{"" if first else ", p"};')
            first = False
        lines.append(f'@p add.rn.{sfx} %0, %0, %1;')
        body = ' '.join(lines)
        self.emit(f'asm("{{ {body} }}" : "+{cons}"({acc}) : "{cons}"({term}), {", ".join(ops)});', ind)

    def _pred_form(self, i: int, blk: Block):
        """Integrand of the shape  [K *] (c ? A : 0)  ->  ([conditions], value text | None for 1).

        ``out + ((c ? A : 0) * K) * step`` equals ``out`` when c is false (adding a signed zero to an
        accumulator that starts at +0 never changes it), so the accumulate can be predicated on c.
        Only constants K are accepted: their product with zero is exactly zero."""
        known = blk.known
        i = self.helper.resolve(i, known)
        n = self.nodes[i]
        if n.op == 'sel' and i in blk.spec:
            z = self.helper.resolve(n.args[2], known)
            if self.nodes[z].op == 'const' and float(self.nodes[z].attr) == 0.0:
                cond = n.args[0]
                a = self.helper.resolve(n.args[1], known)
                inner = self._pred_form(a, blk)
                if inner is not None and inner[0]:
                    return [cond] + inner[0], inner[1]
                na = self.nodes[a]
                if na.op == 'const' and float(na.attr) == 1.0:
                    return [cond], None
                return [cond], self.ref(a, known)
            return None
        if n.op == 'mul':
            for ka, sa in ((n.args[0], n.args[1]), (n.args[1], n.args[0])):
                kk = self.helper.resolve(ka, known)
                if self.nodes[kk].op == 'const':
                    inner = self._pred_form(sa, blk)
                    if inner is not None and inner[0]:
                        kt = self.ref(kk, known)
                        vt = literal('1', self.ctype) if inner[1] is None else inner[1]
                        first = ka == n.args[0]
                        return inner[0], (f'({kt} * {vt})' if first else f'({vt} * {kt})')
        return None

    # ------------------------------------------------------------------- kernels
    def uniform_body(self) -> List[str]:
        self.in_main = False
        self.lines = []
        for i in self.sch.uniform_nodes:
            n = self.nodes[i]
            if n.op == 'domtri':
                self._emit_domtri(i, {}, 1)
            else:
                self.emit(f'const {self.decl(i)} v{i} = {self.expr(i, {})};', 1)
        for i, slot in sorted(self.sch.uniform_slots.items(), key=lambda kv: kv[1]):
            n = self.nodes[i]
            val = f'(v{i} ? {literal("1", self.ctype)} : {literal("0", self.ctype)})' if n.is_bool else f'v{i}'
            self.emit(f'U[{slot}] = {val};', 1)
        return self.lines

    def main_body(self) -> List[str]:
        return self.body_lines(self.sch.main, self.sch.outputs)

    def body_lines(self, block: Block, outputs: Sequence[int]) -> List[str]:
        self.in_main = True
        self.lines = []
        self.guard_ctx = ['live'] if self.coop else []
        self.block(block, 1, 0)
        for k, o in enumerate(outputs):
            self.emit(f'res[{k}] = {self.ref(o, block.known)};', 1)
        return self.lines

    def generate(self, grid_args: Sequence[int] = (), reduce_outputs: bool = False, grouped: bool = False,
                 pixel_args: Sequence[int] = (), accumulate: bool = False, vec: int = 1,
                 pixel_diff: Sequence[int] = ()) -> KernelSource:
        """grid_args: argument indices of (x_lb, x_ub, y_lb, y_ub) when they are generated from the pixel
        index (render programs, tests/plot.py:8-29) instead of being streamed from per-instance arrays.
        reduce_outputs: instead of one output stream per component, every block sums its instances'
        results in a fixed order and writes one partial per component; the runtime adds the partials
        (reverse-mode gradients w.r.t. shared parameters).
        grouped (render programs only): the launch covers G *groups* (e.g. triangles), each with its own
        values of the scalar arguments and its own window of pixels; launch-invariant values become one
        table row per group (global memory, staged through shared memory per block) instead of the
        constant bank.  pixel_args: arguments read from images indexed by pixel (an upstream dL); those listed in
        pixel_diff are bound to TWO images and read as ``a[p] - b[p]`` (one IEEE subtraction, what an elementwise
        ``a - b`` kernel would have stored): the upstream gradient ``image - target`` of an L2 loss is never written.
        accumulate: outputs are added into images (groups of one launch must not overlap).
        vec: instances per thread for streaming (loop-free) programs: 16-byte vector loads/stores of ``vec``
        consecutive instances put 4x the bytes in flight per thread, which is what an HBM-bound kernel
        needs (a 4-byte load per thread saturates at ~2 TB/s on B200); all pointers must be 16-byte aligned."""
        sch = self.sch
        T = self.ctype
        name = self.name
        K = len(sch.outputs)
        # number of TRAILING outputs that are summed over the instances (True = all of them)
        nr = K if reduce_outputs is True else int(reduce_outputs)
        assert 0 <= nr <= K
        ns_out = K - nr
        reduce_outputs = nr > 0
        nu = max(1, len(sch.uniform_slots))
        scal = list(sch.scalar_args)
        grid = list(grid_args)
        pix = list(pixel_args)
        pdiff = [a for a in pix if a in set(pixel_diff)]
        assert all(g in sch.array_args for g in grid + pix), 'grid / pixel arguments must be scheduled as per-instance'
        assert not grid or len(grid) == 4, 'a render program takes exactly x_lb, x_ub, y_lb, y_ub on the grid'
        assert not grouped or grid, 'grouped launches are render launches'
        assert not pix or grouped, 'pixel-indexed arguments need a grouped render launch'
        n_prog_args = len(self.dag.arg_labels)
        tile = [a for a in sch.array_args if a >= n_prog_args]      # tile conditions (pseudo arguments, cull.tile_cull)
        assert not tile or (grid and self.team == 1), 'tile conditions belong to render launches'
        assert len(tile) == len(getattr(self.dag, 'tile_conds', []) if tile else [])
        arrs = [a for a in sch.array_args if a not in grid and a not in pix and a not in tile]
        assert not (grid and arrs), 'render programs cannot also stream per-instance arrays'
        used_in = {self.nodes[i].attr for i in sch.uniform_nodes if self.nodes[i].op == 'in'}
        uni = self.uniform_body()
        main = self.main_body()
        varying = list(sch.array_args)       # instance-function value parameters (arrays, grid, pixel alike)
        team = self.team
        bs = self.block_size

        ubody_params = ', '.join([f'const T a{a}' for a in scal] + ['T* __restrict__ U'])
        outs_decl = [f'T* __restrict__ out{k}' for k in range(ns_out)] + (['T* __restrict__ partials'] if nr else [])
        coop = self.uses_coop
        mc = self.mode == 'monte_carlo_uniform'          # instance functions then take the instance index
        assert not (mc and (grouped or team != 1)), 'monte_carlo_uniform: thread-per-instance array and render launches'

        def inst(expr: str) -> List[str]:
            return [f'(unsigned long long)({expr})'] if mc else []

        iparams = ', '.join((['const T* __restrict__ tegU'] if grouped else []) +
                            (['const bool live'] if coop else []) +
                            ([] if team == 1 else ['const unsigned int team_rank']) +
                            [f'const T a{a}' for a in varying] +
                            (['const unsigned long long tegb_inst'] if mc else []) + [f'T (&res)[{K}]'])
        src: List[str] = []
        src.append(f'// generated by teg_b200.codegen for program "{self.dag.name}"  (precision {T}, {self.N} samples)')
        src.append('#include "teg_b200_runtime.cuh"')
        src.append(f'typedef {T} T;')
        src.append(f'#define TEGB_NUM_UNIFORM {nu}')
        if tile or any(self.nodes[i].op == 'domtri' for i in self.dag.reachable(sch.outputs)):
            src.append(f'#define TEGB_TILE_COLS {self.tile_cols}')
            src.append(f'#define TEGB_TILE_SUB {self.block_size // self.tile_cols}')
            # the poison value of the domain-culling prologue (cull.py)
            src.append(f'#define TEGB_IV_NAN {"nanf" if T == "float" else "nan"}("")')
        if not grouped:
            src.append('TEGB_UNIFORM_TABLE(T, tegU, TEGB_NUM_UNIFORM);')
        if mc:
            src += _MC_PREAMBLE.replace('@SEED@', f'{self.mc_seed}ull').split('\n')
        src.append('')
        src.append(f'TEGB_FN void {name}_uniform_body({ubody_params}) {{')
        for a in scal:
            if a not in used_in:
                src.append(f'    (void)a{a};')
        src += uni
        src.append('}')
        src.append('')
        src.append(f'TEGB_FN void {name}_instance({iparams}) {{')
        for a in varying:
            src.append(f'    (void)a{a};')
        src += main
        src.append('}')
        src.append('')
        # tile-culled render programs: the bodies for resolved tiles (cull.tile_cull); the kernel picks one per tile
        tile_bodies = []
        if tile:
            general_coop = self.uses_coop
            for label, match, blk, outs in sch.bodies:
                if match is None:
                    continue
                self.uses_coop = False
                lines = self.body_lines(blk, outs)
                b_coop = self.uses_coop
                bparams = ', '.join((['const T* __restrict__ tegU'] if grouped else []) +
                                    (['const bool live'] if b_coop else []) +
                                    [f'const T a{a}' for a in varying] + [f'T (&res)[{K}]'])
                src.append(f'TEGB_FN void {name}_tile_{label}({bparams}) {{')
                for a in varying:
                    src.append(f'    (void)a{a};')
                src += lines
                src.append('}')
                src.append('')
                consts = []
                for o in outs:
                    m = self.nodes[self.helper.resolve(o, blk.known)]
                    consts.append(literal(m.attr, T) if m.op == 'const' else None)
                tile_bodies.append((label, match, b_coop, consts if all(c is not None for c in consts) else None))
            self.uses_coop = general_coop
            coop = general_coop
        src.append('#ifdef __CUDACC__')
        zero = literal('0', T)
        if grouped:
            assert team == 1, 'render programs are thread-per-pixel'
            assert nr in (0, K), 'grouped launches reduce all outputs or none'
            uparams = ', '.join([f'const T* __restrict__ g{a}' for a in scal] +
                                ['const int group_begin', 'const int group_count', 'T* __restrict__ U'])
            src.append(f'extern "C" __global__ void {name}_uniform({uparams}) {{')
            src.append('    const int j = blockIdx.x * blockDim.x + threadIdx.x;')
            src.append('    if (j >= group_count) return;')
            src.append('    const int g = group_begin + j;')
            src.append(f'    {name}_uniform_body(' + ', '.join([f'g{a}[g]' for a in scal] +
                                                             ['U + (size_t)g * TEGB_NUM_UNIFORM']) + ');')
            src.append('}')
            src.append('')
            if tile:
                # classification of every (group, window tile): one warp per tile, launch-wide values of the group
                tparams = ', '.join([f'const T* __restrict__ g{a}' for a in scal] +
                                    ['const T* __restrict__ xe', 'const T* __restrict__ ye', 'const int res_y',
                                     'const int row0', 'const int row1', 'const int* __restrict__ win',
                                     'const int group_begin', 'const int group_count', 'const int tile_rows',
                                     'const int n_tx', 'const int n_ty', 'const int tab_groups', 'const int aligned',
                                     'T* __restrict__ tiles'])
                src.append(f'extern "C" __global__ void {name}_tiles({tparams}) {{')
                src.append('    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per tile')
                src.append('    const int per_group = n_tx * n_ty;')
                src.append('    if (w >= (long long)group_count * per_group) return;')
                src.append('    const int j = (int)(w / per_group), tt = (int)(w - (long long)j * per_group);')
                src.append('    const int ty = tt / n_tx, tx = tt - ty * n_tx;')
                src.append('    const int g = group_begin + j;')
                src.append('    const int w_r0 = max(win[4 * g + 0], row0), w_r1 = min(win[4 * g + 1], row1);')
                src.append('    const int w_c0 = win[4 * g + 2], w_c1 = min(win[4 * g + 3], res_y);')
                src.append('    // aligned (gather launches): tiles of the absolute grid -- rows from row0 in steps of tile_rows, columns in')
                src.append('    // steps of TEGB_TILE_COLS -- numbered from the first one the window reaches into; else tiles of the window')
                src.append('    const int o_r0 = aligned ? row0 + ((w_r0 - row0) / tile_rows) * tile_rows : w_r0;')
                src.append('    const int o_c0 = aligned ? (w_c0 / TEGB_TILE_COLS) * TEGB_TILE_COLS : w_c0;')
                src.append('    const int r0 = max(o_r0 + ty * tile_rows, w_r0), r1 = min(o_r0 + (ty + 1) * tile_rows, w_r1);')
                src.append('    const int c0 = max(o_c0 + tx * TEGB_TILE_COLS, w_c0), c1 = min(o_c0 + (tx + 1) * TEGB_TILE_COLS, w_c1);')
                src.append('    // the table is indexed by ABSOLUTE group number: one classification pass may serve several launches')
                src.append('    const size_t n_tiles = (size_t)tab_groups * per_group;')
                src.append('    const size_t t = (size_t)g * per_group + tt;')
                src.append('    if (r0 >= r1 || c0 >= c1) {                    // no pixel: resolved-false, nobody walks it')
                for k in range(len(tile)):
                    src.append(f'        tiles[(size_t){k} * n_tiles + t] = {zero};')
                src.append('        return;')
                src.append('    }')
                for a2 in scal:
                    src.append(f'    const T a{a2} = g{a2}[g];')
                    src.append(f'    (void)a{a2};')
                src += self.tiles_body(grid, walk=True)
                src.append('}')
                src.append('')
            def emit_grouped_main(gather: bool):
                """``<name>_main``: block (bx, by, bz) = tile (bx, by) of the window of group ``group_begin + bz``.
                ``<name>_gather`` (accumulating image programs): block (bx, by) = block (bx, by) of the IMAGE; it walks
                the groups whose windows reach into it, in the order of its bin (ascending group number), so every
                pixel has one owner for the whole launch: overlapping windows need no launch of their own, and a pixel
                sums its contributions in group order.  Tiles are then those of the absolute grid (classification with
                ``aligned`` = 1): rows from row0 in steps of tile_rows, columns in steps of TEGB_TILE_COLS."""
                I = '        ' if gather else '    '
                mparams = ', '.join(['const T* __restrict__ xe', 'const T* __restrict__ ye', 'const int res_y',
                                     'const int row0', 'const int row1', 'const int* __restrict__ win'] +
                                    (['const int* __restrict__ bin_begin', 'const int* __restrict__ bin_groups', 'const int clear']
                                     if gather else ['const int group_begin']) +
                                    ['const T* __restrict__ Utab', 'const int tile_rows'] +
                                    (['const T* __restrict__ tiles', 'const int tab_groups'] if tile else []) +
                                    (['const int n_tx', 'const int n_ty'] if tile and gather else []) +
                                    [p_ for a in pix for p_ in ([f'const T* __restrict__ pix{a}'] +
                                                                ([f'const T* __restrict__ pixb{a}'] if a in pdiff else []))] +
                                    outs_decl)
                gmb = (0 if self.min_blocks is None else self.min_blocks)
                glb = f'{bs}, {gmb}' if gmb else f'{bs}'
                kname = f'{name}_gather' if gather else f'{name}_main'
                src.append(f'extern "C" __global__ void __launch_bounds__({glb}) {kname}({mparams}) {{')
                src.append('    __shared__ T sU[TEGB_NUM_UNIFORM];')
                if gather:
                    src.append('    // block (bx, by) of the image: B columns x tile_rows rows; its bin lists the groups to walk')
                    src.append('    const int nx0 = row0 + blockIdx.y * tile_rows;')
                    src.append('    const int col0 = (int)(blockIdx.x * blockDim.x);')
                    src.append('    const int ny = col0 + threadIdx.x;')
                    src.append('    const int bin = blockIdx.y * gridDim.x + blockIdx.x;')
                    src.append('    if (clear && ny < res_y) {              // the launch starts from a zero image: no fill pass before it')
                    src.append('        const int rows_blk = min(tile_rows, row1 - nx0);')
                    src.append('        for (int r = 0; r < rows_blk; ++r) {')
                    for k in range(K):
                        src.append(f'            out{k}[(long long)(nx0 + r - row0) * res_y + ny] = {zero};')
                    src.append('        }')
                    src.append('    }')
                    src.append('    for (int e_ = bin_begin[bin]; e_ < bin_begin[bin + 1]; ++e_) {')
                    src.append('        const int g = bin_groups[e_];')
                    src.append('        const int w_r0 = max(win[4 * g + 0], row0), w_r1 = min(win[4 * g + 1], row1);')
                    src.append('        const int w_c0 = win[4 * g + 2], w_c1 = min(win[4 * g + 3], res_y);')
                    src.append('        if (nx0 + tile_rows <= w_r0 || nx0 >= w_r1 || col0 + (int)blockDim.x <= w_c0 || col0 >= w_c1) '
                               'continue;     // (block-uniform)')
                    src.append('        __syncthreads();                    // everybody is done with the previous group')
                    src.append('        for (int j = threadIdx.x; j < TEGB_NUM_UNIFORM; j += blockDim.x) '
                               'sU[j] = Utab[(size_t)g * TEGB_NUM_UNIFORM + j];')
                    src.append('        __syncthreads();')
                    # rows of this block inside the window: r_first <= r < rows_here (relative to nx0)
                    src.append('        const int r_first = max(w_r0 - nx0, 0);')
                    src.append('        const int rows_here = min(tile_rows, w_r1 - nx0);')
                    src.append('        const bool live = ny >= w_c0 && ny < w_c1;')
                    src.append('        const int nyc = live ? ny : w_c0;')
                else:
                    src.append('    // block (bx, by, bz): group group_begin + bz, tile (bx, by) of its window: B columns x tile_rows rows')
                    src.append('    const int g = group_begin + blockIdx.z;')
                    src.append('    const int w_r0 = max(win[4 * g + 0], row0), w_r1 = min(win[4 * g + 1], row1);')
                    src.append('    const int w_c0 = win[4 * g + 2], w_c1 = win[4 * g + 3];')
                    src.append('    const int nx0 = w_r0 + blockIdx.y * tile_rows;')
                    src.append('    const int ny = w_c0 + blockIdx.x * blockDim.x + threadIdx.x;')
                    src.append('    const bool block_live = nx0 < w_r1 && (w_c0 + (int)(blockIdx.x * blockDim.x)) < w_c1;')
                    if reduce_outputs:
                        src.append('    const unsigned long long tiles_n = (unsigned long long)gridDim.x * gridDim.y;')
                        src.append('    const unsigned long long tile_i = (unsigned long long)blockIdx.y * gridDim.x + blockIdx.x;')
                        src.append(f'    T* gp = partials + (unsigned long long)blockIdx.z * {K} * tiles_n;')
                        src.append('    if (!block_live) {')
                        src.append(f'        if (threadIdx.x < {K}) gp[threadIdx.x * tiles_n + tile_i] = {zero};')
                        src.append('        return;')
                        src.append('    }')
                    else:
                        src.append('    if (!block_live) return;')
                    src.append('    for (int j = threadIdx.x; j < TEGB_NUM_UNIFORM; j += blockDim.x) '
                               'sU[j] = Utab[(size_t)g * TEGB_NUM_UNIFORM + j];')
                    src.append('    __syncthreads();')
                    src.append('    const bool live = ny < w_c1;')
                    # every lane of a warp runs coop bodies (dead lanes on a clamped pixel, results dropped)
                    src.append('    const int nyc = live ? ny : w_c0 + (int)(blockIdx.x * blockDim.x);')
                    src.append('    const int rows_here = min(tile_rows, w_r1 - nx0);')
                r_first = 'r_first' if gather else '0'
                src.append(f'{I}const T ye_lo = ye[nyc], ye_hi = ye[nyc + 1];')
                loads = {grid[0]: 'x_lo', grid[1]: 'x_hi', grid[2]: 'ye_lo', grid[3]: 'ye_hi'}
                if tile:
                    if gather:
                        # tiles of the absolute grid, numbered within the group's (aligned) window
                        src.append('        const size_t n_tiles = (size_t)n_tx * n_ty * tab_groups;')
                        src.append('        const int o_c0 = (w_c0 / TEGB_TILE_COLS) * TEGB_TILE_COLS;')
                        src.append('        const int t_row = (nx0 - row0) / tile_rows - (w_r0 - row0) / tile_rows;')
                        src.append('        // tile columns of this block, relative to the window (negative / past the end: outside)')
                        src.append('        const int t_col0 = (col0 - o_c0) / TEGB_TILE_COLS;          // (both multiples of TEGB_TILE_COLS)')
                        src.append('        const size_t tile_row_base = ((size_t)g * n_ty + t_row) * n_tx;')
                        src.append('        const int t_col = t_col0 + (int)(threadIdx.x / TEGB_TILE_COLS);')
                        src.append('        const bool in_win = t_col >= 0 && o_c0 + t_col * TEGB_TILE_COLS < w_c1;      // warp-uniform')
                        for k, a in enumerate(tile):
                            src.append(f'        const T tile{k} = in_win ? tiles[{k} * n_tiles + tile_row_base + t_col] : {zero};')
                            loads[a] = f'tile{k}'
                    else:
                        src.append('    const size_t n_tiles = (size_t)gridDim.x * TEGB_TILE_SUB * gridDim.y * tab_groups;')
                        src.append('    const size_t tile_base = ((size_t)g * gridDim.y + blockIdx.y) * (gridDim.x * TEGB_TILE_SUB) + '
                                   'blockIdx.x * TEGB_TILE_SUB;         // (absolute group number)')
                        src.append('    const size_t tile_id = tile_base + threadIdx.x / TEGB_TILE_COLS;')
                        for k, a in enumerate(tile):
                            src.append(f'    const T tile{k} = tiles[{k} * n_tiles + tile_id];')
                            loads[a] = f'tile{k}'
                if reduce_outputs:
                    src.append(f'    T rr[{K}];')
                    src.append(f'    for (int k = 0; k < {K}; k++) rr[k] = {zero};')

                def g_row_loop(fn: str, with_live: bool, ind: str, r_begin: str = '0', r_step: str = '1', batch: int = 4):
                    """Rows of the tile in batches of ``batch``: every global read of the batch (the upstream-gradient
                    pixels, the old image values of accumulating launches) is issued before the first use, so one memory
                    latency is paid per batch instead of per row."""
                    if gather:
                        r_begin = f'r_first + {r_begin}' if r_begin != '0' else 'r_first'
                    src.append(f'{ind}#pragma unroll 1')
                    src.append(f'{ind}for (int r0 = {r_begin}; r0 < rows_here; r0 += {batch} * {r_step}) {{')
                    for a in pix:
                        src.append(f'{ind}    T p{a}_[{batch}];')
                    if accumulate and not reduce_outputs:
                        src.append(f'{ind}    T old_[{batch}][{K}];')
                    src.append(f'{ind}    #pragma unroll')
                    src.append(f'{ind}    for (int u = 0; u < {batch}; u++) {{')
                    src.append(f'{ind}        const int r = r0 + u * {r_step};')
                    src.append(f'{ind}        const bool on = live && r < rows_here;')
                    src.append(f'{ind}        const long long ic = (long long)(nx0 + (r < rows_here ? r : {r_first}) - row0) * res_y + nyc;')
                    for a in pix:
                        src.append(f'{ind}        p{a}_[u] = pix{a}[ic]' + (f' - pixb{a}[ic];' if a in pdiff else ';'))
                    if accumulate and not reduce_outputs:
                        for k in range(K):
                            src.append(f'{ind}        old_[u][{k}] = on ? out{k}[ic] : {zero};')
                    src.append(f'{ind}        (void)on; (void)ic;')
                    src.append(f'{ind}    }}')
                    src.append(f'{ind}    #pragma unroll')
                    src.append(f'{ind}    for (int u = 0; u < {batch}; u++) {{')
                    src.append(f'{ind}        const int r = r0 + u * {r_step};')
                    src.append(f'{ind}        if (r >= rows_here) break;                          // warp-uniform')
                    src.append(f'{ind}        const int nx = nx0 + r;')
                    src.append(f'{ind}        const T x_lo = xe[nx], x_hi = xe[nx + 1];')
                    src.append(f'{ind}        const long long i = (long long)(nx - row0) * res_y + ny;')
                    src.append(f'{ind}        T res[{K}];')
                    src.append(f'{ind}        for (int k = 0; k < {K}; k++) res[k] = {zero};')
                    lu = dict(loads)
                    for a in pix:
                        lu[a] = f'p{a}_[u]'
                    call = ', '.join(['sU'] + (['live'] if with_live else []) + [lu[a] for a in varying] + ['res'])
                    if with_live:
                        src.append(f'{ind}        {fn}({call});')
                    else:
                        src.append(f'{ind}        if (live) {fn}({call});')
                    if reduce_outputs:
                        for k in range(K):
                            src.append(f'{ind}        if (live) rr[{k}] = rr[{k}] + res[{k}];')
                    elif accumulate:
                        for k in range(K):
                            src.append(f'{ind}        if (live) out{k}[i] = old_[u][{k}] + res[{k}];')
                    else:
                        for k in range(K):
                            src.append(f'{ind}        if (live) out{k}[i] = res[{k}];')
                    src.append(f'{ind}    }}')
                    src.append(f'{ind}}}')

                if tile_bodies:
                    first = True
                    for label, match, b_coop, consts in tile_bodies:
                        terms = ['in_win'] if gather else []
                        for k2 in sorted(match):
                            alts = ' || '.join(f'tile{k2} == {literal(str(v), T)}' for v in match[k2])
                            terms.append(f'({alts})')
                        src.append(f'{I}{"if" if first else "else if"} ({" && ".join(terms)}) {{')
                        zeros = consts is not None and all(float(c.rstrip('f')) == 0.0 and not c.startswith('-') for c in consts)
                        if zeros and (reduce_outputs or accumulate):
                            src.append(f'{I}    // every output is +0 over this tile: nothing to add')
                        else:
                            g_row_loop(f'{name}_tile_{label}', b_coop, I + '    ')
                        src.append(f'{I}}}')
                        first = False
                    # Unresolved tiles of this block: the per-instance program, shared by ALL warps of the block (warp w
                    # takes rows w, w + W, ... of each unresolved tile): one warp walking its tile alone would keep the
                    # other warps of the block waiting at the block sum / the end of the block.
                    mine_u = ' || '.join(f'tile{k} == {literal("2", T)}' for k in range(len(tile)))
                    src.append(f'{I}if (__syncthreads_or({mine_u})) {{')
                    src.append(f'{I}    const int warp_ = threadIdx.x >> 5, lane_ = threadIdx.x & 31;')
                    src.append(f'{I}    for (int sub = 0; sub < TEGB_TILE_SUB; ++sub) {{')
                    if gather:
                        src.append(f'{I}        const int s_col = t_col0 + sub;')
                        src.append(f'{I}        if (s_col < 0 || o_c0 + s_col * TEGB_TILE_COLS >= w_c1) continue;      // outside the window')
                        conds_u = ' || '.join(f'tiles[{k} * n_tiles + tile_row_base + s_col] == {literal("2", T)}'
                                              for k in range(len(tile)))
                    else:
                        conds_u = ' || '.join(f'tiles[{k} * n_tiles + tile_base + sub] == {literal("2", T)}'
                                              for k in range(len(tile)))
                    src.append(f'{I}        if (!({conds_u})) continue;                 // block-uniform')
                    src.append(f'{I}        for (int cw = 0; cw < TEGB_TILE_COLS / 32; ++cw) {{')
                    if gather:
                        src.append(f'{I}            const int ny = col0 + sub * TEGB_TILE_COLS + cw * 32 + lane_;')
                        src.append(f'{I}            const bool live = ny >= w_c0 && ny < w_c1;')
                        src.append(f'{I}            const int nyc = live ? ny : w_c0;')
                    else:
                        src.append(f'{I}            const int ny = w_c0 + (int)(blockIdx.x * {bs}) + sub * TEGB_TILE_COLS + cw * 32 + lane_;')
                        src.append(f'{I}            const bool live = ny < w_c1;')
                        src.append(f'{I}            const int nyc = live ? ny : w_c0 + (int)(blockIdx.x * {bs});')
                    src.append(f'{I}            const T ye_lo = ye[nyc], ye_hi = ye[nyc + 1];')
                    for k in range(len(tile)):
                        src.append(f'{I}            const T tile{k} = {literal("2", T)};')
                        src.append(f'{I}            (void)tile{k};')
                    g_row_loop(f'{name}_instance', coop, I + '            ', 'warp_', f'{bs // 32}', batch=1)
                    src.append(f'{I}        }}')
                    src.append(f'{I}    }}')
                    src.append(f'{I}}}')
                else:
                    g_row_loop(f'{name}_instance', coop, I, batch=1)
                if gather:
                    src.append('    }')
                if reduce_outputs:
                    src.append(f'    tegb_block_sum_store<T, {K}, {bs}>(rr, gp, tiles_n, tile_i);')
                src.append('}')

            emit_grouped_main(False)
            self.has_gather = bool(accumulate and not reduce_outputs)
            if self.has_gather:
                # (an image block walks SEVERAL groups: rows outside a group's window are skipped by r_first / rows_here)
                src.append('')
                emit_grouped_main(True)
        else:
            src.append(f'extern "C" __global__ void {name}_uniform({ubody_params}) {{')
            src.append(f'    if (blockIdx.x == 0 && threadIdx.x == 0) {name}_uniform_body('
                       + ', '.join([f'a{a}' for a in scal] + ['U']) + ');')
            src.append('}')
            src.append('')
            if grid:
                assert team == 1, 'render programs are thread-per-pixel'
                # tiled programs whose outputs are ALL summed (one partial per 32-column tile): the classification kernel
                # writes the zero partials of tiles without work and the main kernel visits only the others
                sparse = bool(tile) and bool(tile_bodies) and nr == K and self.tile_cols == 32
                sparse_bodies = [(match, consts is not None and all(float(c.rstrip('f')) == 0.0 and not c.startswith('-')
                                                                    for c in consts))
                                 for _, match, _, consts in tile_bodies]
                # pixel (nx, ny) -> instance nx * res_y + ny; box bounds from the two np.linspace edge tables
                if tile:
                    tparams = ', '.join([f'const T a{a}' for a in scal] +
                                        ['const T* __restrict__ ranges', 'const int res_y', 'const int n_tx', 'const int n_ty', 'const int n_txs',
                                         'T* __restrict__ tiles', 'T* __restrict__ tiles_ord', 'int* __restrict__ order',
                                         'int* __restrict__ counters', 'int* __restrict__ ulist'] +
                                        (['T* __restrict__ partials'] if sparse else []))
                    src.append(f'extern "C" __global__ void {name}_tiles({tparams}) {{')
                    src.append('    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;    // one thread per tile')
                    src.append('    const int t_ = gtid / TEGB_TILE_SUB, sub = gtid % TEGB_TILE_SUB;')
                    src.append('    const bool active = t_ < n_tx * n_ty;                      // (whole warps stay: shuffles below)')
                    src.append('    const int t = active ? t_ : 0;')
                    src.append('    const size_t n_tiles = (size_t)n_txs * n_ty;')
                    src.append('    const int ty = t / n_tx, tx = t - ty * n_tx;')
                    for a in scal:
                        src.append(f'    (void)a{a};')
                    src += self.tiles_body(grid, sparse_bodies=sparse_bodies if sparse else None)
                    src.append('}')
                    src.append('')
                # one block = one tile: block_size columns x tile_rows rows; every thread walks its column's rows
                # (the per-block set-up, the tile tri-states and the block sum are paid once per tile)
                mparams = ', '.join(['const T* __restrict__ xe', 'const T* __restrict__ ye', 'const int res_y',
                                     'const int row0', 'const int n_rows', 'const int tile_rows', 'const int row_stride',
                                     'const int row_phase'] +
                                    (['const T* __restrict__ tiles', 'const int* __restrict__ order',
                                      'const int* __restrict__ ulist', 'int* queue', 'const int n_tx', 'const int n_ty',
                                      'const int n_workers'] if tile else []) +
                                    outs_decl +
                                    [])
                # (sparse programs run their per-pixel program only -- the quiet tiles cost no block --: 128 registers, no
                # spills, measured 22.6 -> 17.5 us on the 4096^2 gradient kernel; the others keep 8 blocks per SM in flight)
                mb = ((4 if sparse else 8) if tile and bs == 128 else 0) if self.min_blocks is None else self.min_blocks
                lb = f'{bs}, {mb}' if mb else f'{bs}'
                src.append(f'extern "C" __global__ void __launch_bounds__({lb}) {name}_main({mparams}) {{')
                if tile:
                    # 1-D grid: n_workers blocks that serve the queue of unresolved rows (below), then one block per
                    # n_tx x n_ty tile block, taken in the order the classification listed them
                    src.append('    const int n_txs = n_tx * TEGB_TILE_SUB;')
                    src.append('    const size_t n_tiles = (size_t)n_txs * n_ty;')
                    QUEUE_MARK = len(src)
                    if sparse:
                        # (a fixed number of blocks walks the counters[0] listed tile blocks: usually none or a few)
                        src.append('    const unsigned int n_listed = (unsigned int)queue[0];')
                        src.append('    for (unsigned int slot = blockIdx.x - (unsigned int)n_workers; slot < n_listed; '
                                   'slot += gridDim.x - (unsigned int)n_workers) {')
                    else:
                        src.append('    const unsigned int slot = blockIdx.x - (unsigned int)n_workers;')
                    src.append('    const unsigned int ord = (unsigned int)order[slot];')
                    src.append('    const int tby = (int)(ord >> 16), tbx = (int)(ord & 0xffffu);')
                else:
                    src.append('    const int tby = blockIdx.y, tbx = blockIdx.x;')
                src.append('    const int ny = tbx * blockDim.x + threadIdx.x;')
                src.append('    const bool live = ny < res_y;')
                ycol = 'nyc' if coop else 'ny'
                if coop:
                    src.append('    const int nyc = live ? ny : 0;')
                src.append(f'    const T ye_lo = ye[{ycol}], ye_hi = ye[{ycol} + 1];' if coop else
                           '    const T ye_lo = live ? ye[ny] : ye[0], ye_hi = live ? ye[ny + 1] : ye[1];')
                loads = {grid[0]: 'x_lo', grid[1]: 'x_hi', grid[2]: 'ye_lo', grid[3]: 'ye_hi'}
                if tile:
                    src.append('    const size_t tile_base = (size_t)slot * TEGB_TILE_SUB;          // states are stored in launch order')
                    src.append('    const size_t tile_id = tile_base + threadIdx.x / TEGB_TILE_COLS;')
                    for k, a in enumerate(tile):
                        src.append(f'    const T tile{k} = tiles[{k} * n_tiles + tile_id];')
                        loads[a] = f'tile{k}'
                # rows of this block: local tile row tby is tile row tby * row_stride + row_phase of the row range
                # (interleaved sharding: rank r renders tile rows r, r + world, ...); the local image is compact
                src.append('    const int rel0 = (tby * row_stride + row_phase) * tile_rows;')
                src.append('    const int rows_here = min(tile_rows, n_rows - rel0);')
                src.append('    const int nx0 = row0 + rel0;')
                src.append('    const long long i0 = (long long)(tby * tile_rows) * res_y + ny;')
                if nr:
                    src.append(f'    T rr[{nr}];')
                    src.append(f'    for (int k = 0; k < {nr}; k++) rr[k] = {zero};')

                def row_loop(fn: str, with_live: bool, ind: str, r_begin: str = '0', r_step: str = '1', unroll: int = 1):
                    # (Monte Carlo: the instance index of a pixel is its index in the whole image, whatever the band)
                    call = ', '.join((['live'] if with_live else []) + [loads[a] for a in varying] +
                                     inst('(long long)(nx0 + r) * res_y + ny') + ['res'])
                    src.append(f'{ind}#pragma unroll {unroll}')
                    src.append(f'{ind}for (int r = {r_begin}; r < rows_here; r += {r_step}) {{')
                    src.append(f'{ind}    const T x_lo = xe[nx0 + r], x_hi = xe[nx0 + r + 1];')
                    src.append(f'{ind}    const long long i = i0 + (long long)r * res_y;')
                    src.append(f'{ind}    T res[{K}];')
                    if nr:
                        src.append(f'{ind}    for (int k = 0; k < {K}; k++) res[k] = {zero};')
                    if with_live:
                        src.append(f'{ind}    {fn}({call});')
                    else:
                        src.append(f'{ind}    if (live) {fn}({call});')
                    for k in range(ns_out):
                        src.append(f'{ind}    if (live) out{k}[i] = res[{k}];')
                    for j in range(nr):
                        # rows are added in order: a fixed summation order per thread, then the fixed block sum
                        src.append(f'{ind}    if (live) rr[{j}] = rr[{j}] + res[{ns_out + j}];')
                    src.append(f'{ind}}}')

                if tile_bodies:
                    # Rows of unresolved tiles.  One row of an unresolved tile (per-pixel interval test, then cooperative
                    # sample loops for the pixels an edge really crosses) is a long dependent chain; a block walking 16
                    # such rows sets the kernel's duration (19 unresolved tiles of 32768 cost 9 us).
                    # Image programs: the first n_workers blocks of the grid serve a queue of (unresolved tile, row) items,
                    # one row per warp at a time; surplus workers leave at once; pixels are written in place, so results
                    # do not depend on who serves an item.  Programs with summed outputs keep the walk inside the owning
                    # block (all its warps share its unresolved tiles, below): their partial sums stay tied to the block.
                    per_tile_sums = bool(nr) and self.tile_cols == 32
                    if per_tile_sums:
                        # Programs with summed outputs: one partial sum per TILE (slot = the tile's index), so the sums
                        # do not depend on who computes a tile either.  Resolved tiles: their own warp (warp-level
                        # fixed-order sum, no block barrier).  Unresolved tiles: a queue of tiles served by the worker
                        # blocks, all warps of a worker share the rows of its tile and the block adds up in a fixed order.
                        q = []
                        q.append('    if (blockIdx.x < (unsigned int)n_workers) {')
                        q.append('        __shared__ int q_next;')
                        q.append('        const int warp_ = threadIdx.x >> 5, lane_ = threadIdx.x & 31;')
                        q.append('        const int q_nu = queue[2];                          // unresolved tiles (classification kernel)')
                        q.append('        if ((int)blockIdx.x >= q_nu) return;                  // more workers than tiles')
                        q.append('        for (;;) {')
                        q.append('            __syncthreads();')
                        q.append('            if (threadIdx.x == 0) q_next = atomicAdd(&queue[3], 1);')
                        q.append('            __syncthreads();')
                        q.append('            const int q_u = q_next;')
                        q.append('            if (q_u >= q_nu) break;                            // block-uniform')
                        q.append('            const unsigned int q_e = (unsigned int)ulist[q_u];')
                        q.append('            const int q_ty = (int)(q_e >> 16), q_txs = (int)(q_e & 0xffffu);')
                        q.append('            const int q_rel0 = (q_ty * row_stride + row_phase) * tile_rows;')
                        q.append('            const int q_rows = min(tile_rows, n_rows - q_rel0);')
                        q.append(f'            T q_rr[{nr}];')
                        q.append(f'            for (int k = 0; k < {nr}; k++) q_rr[k] = {zero};')
                        q.append('            const int ny = q_txs * 32 + lane_;')
                        q.append('            const bool live = ny < res_y;')
                        q.append('            const int nyc = live ? ny : 0;')
                        q.append('            const T ye_lo = ye[nyc], ye_hi = ye[nyc + 1];')
                        qloads = dict(loads)
                        for k, a in enumerate(tile):
                            q.append(f'            const T tile{k} = {literal("2", T)};')
                            q.append(f'            (void)tile{k};')
                            qloads[a] = f'tile{k}'
                        q.append('            #pragma unroll 1')
                        q.append(f'            for (int q_r = warp_; q_r < q_rows; q_r += {bs // 32}) {{')
                        q.append('                const int q_nx = row0 + q_rel0 + q_r;')
                        q.append('                const T x_lo = xe[q_nx], x_hi = xe[q_nx + 1];')
                        q.append('                const long long i = (long long)(q_ty * tile_rows + q_r) * res_y + ny;')
                        q.append('                (void)i;')
                        q.append(f'                T res[{K}];')
                        q.append(f'                for (int k = 0; k < {K}; k++) res[k] = {zero};')
                        qcall = ', '.join((['live'] if coop else []) + [qloads[a] for a in varying] +
                                          inst('(long long)q_nx * res_y + ny') + ['res'])
                        if coop:
                            q.append(f'                {name}_instance({qcall});')
                        else:
                            q.append(f'                if (live) {name}_instance({qcall});')
                        for k in range(ns_out):
                            q.append(f'                if (live) out{k}[i] = res[{k}];')
                        for j in range(nr):
                            q.append(f'                if (live) q_rr[{j}] = q_rr[{j}] + res[{ns_out + j}];')
                        q.append('            }')
                        q.append(f'            tegb_block_sum_store<T, {nr}, {bs}>(q_rr, partials, (unsigned long long)n_tiles, '
                                 '(unsigned long long)q_ty * n_txs + q_txs);')
                        q.append('        }')
                        q.append('        return;')
                        q.append('    }')
                        src[QUEUE_MARK:QUEUE_MARK] = q
                    if nr == 0:
                        q = []
                        q.append('    if (blockIdx.x < (unsigned int)n_workers) {')
                        q.append('        const int lane_ = threadIdx.x & 31;')
                        q.append('        const int q_nu = queue[2];                          // unresolved tiles (classification kernel)')
                        q.append('        const int q_items = q_nu * tile_rows;')
                        q.append(f'        if ((long long)blockIdx.x * {bs // 32} >= q_items) return;          // more workers than rows')
                        q.append('        for (;;) {')
                        q.append('            int it = 0;')
                        q.append('            if (lane_ == 0) it = atomicAdd(&queue[3], 1);')
                        q.append('            it = __shfl_sync(0xffffffffu, it, 0);')
                        q.append('            if (it >= q_items) break;')
                        q.append('            const int q_r = it / q_nu, q_u = it - q_r * q_nu;      // neighbouring items: different tiles')
                        q.append('            const unsigned int q_e = (unsigned int)ulist[q_u];')
                        q.append('            const int q_ty = (int)(q_e >> 16), q_txs = (int)(q_e & 0xffffu);')
                        q.append('            const int q_rel0 = (q_ty * row_stride + row_phase) * tile_rows;')
                        q.append('            if (q_r >= min(tile_rows, n_rows - q_rel0)) continue;          // (the last tile row may be short)')
                        q.append('            const int q_nx = row0 + q_rel0 + q_r;')
                        q.append('            const T x_lo = xe[q_nx], x_hi = xe[q_nx + 1];')
                        q.append('            for (int cw = 0; cw < TEGB_TILE_COLS / 32; ++cw) {')
                        q.append('                const int ny = q_txs * TEGB_TILE_COLS + cw * 32 + lane_;')
                        q.append('                const bool live = ny < res_y;')
                        q.append('                const int nyc = live ? ny : 0;')
                        q.append('                const T ye_lo = ye[nyc], ye_hi = ye[nyc + 1];')
                        q.append('                const long long i = (long long)(q_ty * tile_rows + q_r) * res_y + ny;')
                        qloads = dict(loads)
                        for k, a in enumerate(tile):
                            q.append(f'                const T tile{k} = {literal("2", T)};')
                            q.append(f'                (void)tile{k};')
                            qloads[a] = f'tile{k}'
                        q.append(f'                T res[{K}];')
                        qcall = ', '.join((['live'] if coop else []) + [qloads[a] for a in varying] +
                                          inst('(long long)q_nx * res_y + ny') + ['res'])
                        if coop:
                            q.append(f'                {name}_instance({qcall});')
                        else:
                            q.append(f'                if (live) {name}_instance({qcall});')
                        for k in range(ns_out):
                            q.append(f'                if (live) out{k}[i] = res[{k}];')
                        q.append('            }')
                        q.append('        }')
                        q.append('        return;')
                        q.append('    }')
                        src[QUEUE_MARK:QUEUE_MARK] = q
                    per = 4 if T == 'float' else 2                  # elements of a 16-byte vector
                    V = 'float4' if T == 'float' else 'double2'
                    if ns_out:
                        aligned = ' && '.join([f'(((size_t)out{k}) & 15) == 0' for k in range(ns_out)] + [f'(res_y % {per}) == 0'])
                        src.append(f'    const bool vec_ok = {aligned};                 // uniform')
                    # the tile's conditions are the same for every thread of the block: a uniform branch picks the body
                    first = True
                    for label, match, b_coop, consts in tile_bodies:
                        terms = []
                        for k2 in sorted(match):
                            alts = ' || '.join(f'tile{k2} == {literal(str(v), T)}' for v in match[k2])
                            terms.append(f'({alts})')
                        src.append(f'    {"if" if first else "else if"} ({" && ".join(terms)}) {{')
                        zeros = consts is not None and all(float(c.rstrip('f')) == 0.0 and not c.startswith('-')
                                                           for c in consts[ns_out:])
                        if consts is not None and (nr == 0 or zeros):
                            # every output of this body is a constant: fill the tile's rows (summed outputs that are
                            # +0 leave the partial sums as they are).  A warp owns 32 columns x rows_here rows: with
                            # 16-byte stores every instruction covers {per} rows of 128 / 256 bytes each.
                            if ns_out:
                                src.append('        if (vec_ok) {')
                                src.append('            const int lane_ = threadIdx.x & 31;')
                                src.append(f'            const int vcol = (ny - lane_) + (lane_ % {32 // per}) * {per};')
                                src.append(f'            if (vcol < res_y) {{')
                                src.append(f'                const long long v0 = (long long)(tby * tile_rows) * res_y + vcol;')
                                src.append('                #pragma unroll 4')
                                src.append(f'                for (int r = lane_ / {32 // per}; r < rows_here; r += {per}) {{')
                                for k in range(ns_out):
                                    vals = ', '.join([consts[k]] * per)
                                    src.append(f'                    *reinterpret_cast<{V}*>(out{k} + v0 + (long long)r * res_y) = make_{V}({vals});')
                                src.append('                }')
                                src.append('            }')
                                src.append('        } else if (live) {')
                                src.append('            #pragma unroll 4')
                                src.append('            for (int r = 0; r < rows_here; ++r) {')
                                for k in range(ns_out):
                                    src.append(f'                out{k}[i0 + (long long)r * res_y] = {consts[k]};')
                                src.append('            }')
                                src.append('        }')
                        else:
                            # bodies without a sample loop left (every condition fixed) are a short dependent chain per
                            # pixel: several rows in flight hide it
                            row_loop(f'{name}_tile_{label}', b_coop, '        ', unroll=1 if b_coop else 4)
                        src.append('    }')
                        first = False
                    if nr and not per_tile_sums:
                        # unresolved tiles of this block, shared by ALL its warps (warp w takes rows w, w + W, ... of each)
                        mine_u = ' || '.join(f'tile{k} == {literal("2", T)}' for k in range(len(tile)))
                        src.append(f'    if (__syncthreads_or({mine_u})) {{')
                        src.append('        const int warp_ = threadIdx.x >> 5, lane_ = threadIdx.x & 31;')
                        src.append('        for (int sub = 0; sub < TEGB_TILE_SUB; ++sub) {')
                        conds_u = ' || '.join(f'tiles[{k} * n_tiles + tile_base + sub] == {literal("2", T)}' for k in range(len(tile)))
                        src.append(f'            if (!({conds_u})) continue;                 // block-uniform')
                        src.append('            for (int cw = 0; cw < TEGB_TILE_COLS / 32; ++cw) {')
                        src.append(f'                const int ny = tbx * {bs} + sub * TEGB_TILE_COLS + cw * 32 + lane_;')
                        src.append('                const bool live = ny < res_y;')
                        src.append('                const int nyc = live ? ny : 0;')
                        src.append('                const T ye_lo = ye[nyc], ye_hi = ye[nyc + 1];')
                        src.append('                const long long i0 = (long long)(tby * tile_rows) * res_y + ny;')
                        for k in range(len(tile)):
                            src.append(f'                const T tile{k} = {literal("2", T)};')
                        src.append(f'                (void)nyc; (void)ye_lo; (void)ye_hi; (void)i0;')
                        row_loop(f'{name}_instance', coop, '                ', 'warp_', f'{bs // 32}')
                        src.append('            }')
                        src.append('        }')
                        src.append('    }')
                else:
                    row_loop(f'{name}_instance', coop, '    ')
                if nr:
                    # the partial of a tile lives at the tile's own index: the sums do not depend on the block order
                    if tile and self.tile_cols == 32:
                        # one partial per tile; unresolved tiles are summed by the worker that served them
                        unres = ' || '.join(f'tile{k} == {literal("2", T)}' for k in range(len(tile)))
                        src.append(f'    if (!({unres}))                              // warp-uniform')
                        src.append(f'        tegb_warp_sum_store<T, {nr}>(rr, partials + ((unsigned long long)tby * n_txs + tbx * TEGB_TILE_SUB '
                                   '+ (threadIdx.x >> 5)), (unsigned long long)n_tiles);')
                    elif tile:
                        src.append(f'    tegb_block_sum_store<T, {nr}, {bs}>(rr, partials, (unsigned long long)n_tx * n_ty, '
                                   '(unsigned long long)tby * n_tx + tbx);')
                    else:
                        src.append(f'    tegb_block_sum_store<T, {nr}, {bs}>(rr, partials, (unsigned long long)gridDim.x * gridDim.y, '
                                   '(unsigned long long)tby * gridDim.x + tbx);')
                if sparse:
                    src.append('    }')
                src.append('}')
                src.append('#endif')
                text = '\n'.join(src) + '\n'
                return KernelSource(name=name, ctype=T, num_samples=self.N, source=text, takes_instance_index=mc,
                                    uniform_kernel=f'{name}_uniform', main_kernel=f'{name}_main',
                                    n_uniform=nu, scalar_args=scal, array_args=arrs, grid_args=grid,
                                    n_outputs=K, block_size=bs, reduce_outputs=nr, team=team,
                                    grouped=False, pixel_args=pix, accumulate=False,
                                    tile_conds=len(tile), tile_rows=self.tile_rows, tile_cols=self.tile_cols,
                                    tile_kernel=f'{name}_tiles' if tile else '',
                                    tile_bodies=[(lb, mt) for lb, mt, _, _ in tile_bodies],
                                    stats={'uses_coop': self.uses_coop, 'loops': self.loop_counter})
            else:
                mparams = ', '.join([f'const T* __restrict__ in{a}' for a in arrs] + outs_decl + ['const long long n'])
                src.append(f'extern "C" __global__ void __launch_bounds__({bs}) {name}_main({mparams}) {{')
                if vec > 1:
                    assert team == 1 and not coop and nr == 0 and vec == 4
                    V = 'float4' if T == 'float' else 'double2'
                    per = 4 if T == 'float' else 2          # elements per 16-byte vector
                    nvec = vec // per
                    comps = 'xyzw' if T == 'float' else 'xy'
                    src.append('    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;')
                    src.append(f'    const long long i0 = t * {vec};')
                    src.append('    if (i0 >= n) return;')
                    src.append(f'    if (i0 + {vec} <= n) {{')
                    for a in varying:
                        for h in range(nvec):
                            src.append(f'        const {V} A{a}_{h} = *reinterpret_cast<const {V}*>(in{a} + i0 + {h * per});')
                    for j in range(vec):
                        h, c = divmod(j, per)
                        src.append(f'        T r{j}[{K}];')
                        src.append(f'        {name}_instance(' + ', '.join([f'A{a}_{h}.{comps[c]}' for a in varying] +
                                                                       inst(f'i0 + {j}') + [f'r{j}']) + ');')
                    for k in range(K):
                        for h in range(nvec):
                            vals = ', '.join(f'r{h * per + c}[{k}]' for c in range(per))
                            src.append(f'        *reinterpret_cast<{V}*>(out{k} + i0 + {h * per}) = make_{V}({vals});')
                    src.append('    } else {')
                    src.append('        for (long long i = i0; i < n; ++i) {')
                    src.append(f'            T res[{K}];')
                    src.append(f'            {name}_instance(' + ', '.join([f'in{a}[i]' for a in varying] + inst('i') + ['res']) + ');')
                    for k in range(K):
                        src.append(f'            out{k}[i] = res[{k}];')
                    src.append('        }')
                    src.append('    }')
                    src.append('}')
                    src.append('#endif')
                    text = '\n'.join(src) + '\n'
                    return KernelSource(name=name, ctype=T, num_samples=self.N, source=text, takes_instance_index=mc,
                                        uniform_kernel=f'{name}_uniform', main_kernel=f'{name}_main',
                                        n_uniform=nu, scalar_args=scal, array_args=arrs, grid_args=grid,
                                        n_outputs=K, block_size=bs, reduce_outputs=0, team=1, vec=vec,
                                        stats={'uses_coop': False, 'loops': self.loop_counter})
                if team == 1:
                    src.append('    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;')
                    if coop:
                        src.append('    const long long ic = i < n ? i : 0;')
                        call = ', '.join(['i < n'] + [f'in{a}[ic]' for a in varying] + inst('ic') + ['res'])
                    else:
                        call = ', '.join([f'in{a}[i]' for a in varying] + inst('i') + ['res'])
                else:
                    # TEAM threads per instance (a slice of a warp, a warp, or a whole block)
                    src.append('    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;')
                    src.append(f'    const long long i = gtid / {team};')
                    src.append(f'    const unsigned int team_rank = (unsigned int)(gtid % {team});')
                    call = ', '.join(['team_rank'] + [f'in{a}[i]' for a in varying] + ['res'])
                src.append('    const bool live = i < n;')
            src.append(f'    T res[{K}];')
            if nr:
                # every thread of the block reaches the block sum: dead lanes contribute zeros
                src.append(f'    for (int k = 0; k < {K}; k++) res[k] = {zero};')
                if coop:
                    src.append(f'    {name}_instance({call});')
                else:
                    src.append(f'    if (live) {name}_instance({call});')
                owner = 'live' if team == 1 else '(live && team_rank == 0)'
                for k in range(ns_out):
                    src.append(f'    if ({owner}) out{k}[i] = res[{k}];')
                src.append(f'    T rr[{nr}];')
                for j in range(nr):
                    src.append(f'    rr[{j}] = {owner} ? res[{ns_out + j}] : {zero};')
                src.append(f'    tegb_block_sum_store<T, {nr}, {bs}>(rr, partials);')
            else:
                if coop:
                    src.append(f'    {name}_instance({call});')
                    guard = 'if (live) '
                else:
                    src.append('    if (!live) return;')
                    src.append(f'    {name}_instance({call});')
                    guard = '' if team == 1 else 'if (team_rank == 0) '
                for k in range(K):
                    src.append(f'    {guard}out{k}[i] = res[{k}];')
            src.append('}')
        src.append('#endif')
        text = '\n'.join(src) + '\n'
        return KernelSource(name=name, ctype=T, num_samples=self.N, source=text, takes_instance_index=mc,
                            uniform_kernel=f'{name}_uniform', main_kernel=f'{name}_main',
                            n_uniform=nu, scalar_args=scal, array_args=arrs, grid_args=grid,
                            n_outputs=K, block_size=bs, reduce_outputs=nr, team=team,
                            grouped=bool(grouped), pixel_args=pix, pixel_diff=pdiff, accumulate=bool(accumulate),
                            tile_conds=len(tile), tile_rows=self.tile_rows if (tile or grouped) else 0, tile_cols=self.tile_cols,
                            tile_kernel=f'{name}_tiles' if tile else '',
                            gather_kernel=f'{name}_gather' if (grouped and getattr(self, 'has_gather', False)) else '',
                            tile_bodies=[(lb, mt) for lb, mt, _, _ in tile_bodies],
                            stats={'uses_coop': self.uses_coop, 'loops': self.loop_counter})


def generate(sch: Schedule, ctype: str = 'float', num_samples: int = 50, name: Optional[str] = None,
             grid_args: Sequence[int] = (), reduce_outputs: bool = False, grouped: bool = False,
             pixel_args: Sequence[int] = (), accumulate: bool = False, vec: int = 1,
             pixel_diff: Sequence[int] = (), **opts) -> KernelSource:
    name = name or _c_ident(sch.dag.name)
    return Emitter(sch, ctype, num_samples, name, **opts).generate(
        grid_args=grid_args, reduce_outputs=reduce_outputs, grouped=grouped, pixel_args=pixel_args,
        accumulate=accumulate, vec=vec, pixel_diff=pixel_diff)


def _c_ident(s: str) -> str:
    out = ''.join(ch if (ch.isalnum() or ch == '_') else '_' for ch in s)
    if not out or out[0].isdigit():
        out = 'p_' + out
    return out
