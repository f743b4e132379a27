"""Randomised check of domain culling (teg_b200/cull.py): random 2-D and 3-D integrands whose conditions mix sums,
products, quotients, square roots, nested selects, && and ||.  The generated code (culling on) is compiled for the
host (tests/host_harness.py) and must equal the oracle bit for bit on boxes that straddle, touch and miss the
discontinuities -- including boxes where a divisor interval crosses zero or a radicand turns negative (NaN)."""
import numpy as np
import pytest

from host_harness import run_on_host
from teg_b200.codegen import generate
from teg_b200.cull import cull
from teg_b200.dag import build_dag
from teg_b200.lang import Const, IfElse, Sqr, Sqrt, Teg, TegVar, Var
from teg_b200.schedule import make_schedule
from teg_b200.tegp import from_teg


def _rand_expr(rng, leaves, depth):
    if depth == 0 or rng.random() < 0.25:
        return leaves[rng.integers(len(leaves))]
    k = rng.integers(7)
    a = _rand_expr(rng, leaves, depth - 1)
    b = _rand_expr(rng, leaves, depth - 1)
    if k == 0:
        return a + b
    if k == 1:
        return a * b
    if k == 2:
        return a - b
    if k == 3:
        return a / (b + Const(float(rng.choice([0.0, 0.7, 3.0]))))         # divisor intervals may cross zero
    if k == 4:
        return Sqrt(Sqr(a) + b)                                              # radicand may turn negative
    if k == 5:
        return Sqr(a)
    return IfElse(a < b, a, b * Const(0.5))                                  # a select inside the condition's cone


def _rand_cond(rng, leaves, depth):
    if depth == 0 or rng.random() < 0.4:
        return _rand_expr(rng, leaves, 2) < _rand_expr(rng, leaves, 2)
    a = _rand_cond(rng, leaves, depth - 1)
    b = _rand_cond(rng, leaves, depth - 1)
    return (a & b) if rng.random() < 0.6 else (a | b)


@pytest.mark.parametrize('seed', range(24))
def test_random_programs_culled_equals_oracle(seed):
    from oracle import Oracle
    rng = np.random.default_rng(1000 + seed)
    dims = 2 if seed % 3 else 3
    tv = [TegVar(f'x{i}') for i in range(dims)]
    lo = [Var(f'lo{i}', 0.0) for i in range(dims)]
    hi = [Var(f'hi{i}', 1.0) for i in range(dims)]
    params = [Var(f'p{i}', 0.5) for i in range(3)]
    leaves = tv + params + [Const(1.0), Const(-0.5)]
    n_conds = 1 + seed % 3
    body = Const(0.0)
    for _ in range(n_conds):
        body = body + IfElse(_rand_cond(rng, leaves, 2), _rand_expr(rng, leaves, 2), _rand_expr(rng, params + [Const(0.0)], 1))
    expr = body
    for i in reversed(range(dims)):
        expr = Teg(lo[i], hi[i], expr, tv[i])
    args = lo + hi + params
    prog = from_teg(expr, args, name=f'fuzz{seed}')
    ns = 8 if dims == 2 else 4
    dag0 = build_dag(prog)
    dag = cull(dag0, ns)
    n = 400
    centre = rng.uniform(-1.0, 1.0, size=(dims, n))
    width = 10.0 ** rng.uniform(-3, 0.3, size=(dims, n))
    ins = [centre[i] - width[i] / 2 for i in range(dims)] + [centre[i] + width[i] / 2 for i in range(dims)]
    ins += [rng.uniform(-1, 1, size=n) for _ in params]
    ins[0][:5] = [np.nan, np.inf, -np.inf, 0.3, 0.3]                         # non-finite and empty boxes
    ins[dims][3] = 0.3
    ins[dims][4] = 0.1                                                         # reversed
    arr = list(range(len(args)))
    for dtype, ctype in ((np.float32, 'float'), (np.float64, 'double')):
        vals = [np.asarray(v, dtype=dtype) for v in ins]
        want = Oracle(prog.dumps()).eval(vals, ns, dtype)
        ks = generate(make_schedule(dag, arr), ctype, ns, use_asm=False, coop=False)
        got = run_on_host(ks, vals, n, dtype)
        assert np.array_equal(got, want, equal_nan=True), (seed, ctype, np.count_nonzero(~((got == want) | (np.isnan(got) & np.isnan(want)))))
        if not dag.culled:
            continue
        dense = generate(make_schedule(dag0, arr), ctype, ns, use_asm=False, coop=False)
        assert np.array_equal(run_on_host(dense, vals, n, dtype), want, equal_nan=True)


@pytest.mark.parametrize('seed', range(12))
def test_random_single_condition_programs(seed):
    """One condition per nest: the static true / false specialisations (zero nests collapse to 0 + 0 * width)."""
    from oracle import Oracle
    rng = np.random.default_rng(5000 + seed)
    x, y = TegVar('x'), TegVar('y')
    lo = [Var('lo0', 0.0), Var('lo1', 0.0)]
    hi = [Var('hi0', 1.0), Var('hi1', 1.0)]
    params = [Var(f'p{i}', 0.5) for i in range(3)]
    aff = lambda: params[rng.integers(3)] * x + params[rng.integers(3)] * y + Const(float(rng.uniform(-1, 1)))   # noqa: E731
    cond = (aff() < aff()) if seed % 2 else ((aff() < Const(0.0)) & (Const(0.0) < aff()))
    then = [Const(1.0), x * y, Sqr(x) + params[0], params[1]][seed % 4]
    other = [Const(0.0), Const(0.0), params[2], Const(0.0)][seed % 4]
    expr = Teg(lo[0], hi[0], Teg(lo[1], hi[1], IfElse(cond, then, other), y), x)
    args = lo + hi + params
    prog = from_teg(expr, args, name=f'single{seed}')
    dag = cull(build_dag(prog), 8)
    assert dag.culled and all(len(r) == 1 for _, r in dag.culled)
    n = 500
    c = rng.uniform(-1.0, 1.0, size=(2, n))
    w = 10.0 ** rng.uniform(-3, 0.3, size=(2, n))
    ins = [c[0] - w[0] / 2, c[1] - w[1] / 2, c[0] + w[0] / 2, c[1] + w[1] / 2] + [rng.uniform(-1, 1, size=n) for _ in params]
    ins[0][:3] = [np.nan, np.inf, 0.9]
    for dtype, ctype in ((np.float32, 'float'), (np.float64, 'double')):
        vals = [np.asarray(v, dtype=dtype) for v in ins]
        want = Oracle(prog.dumps()).eval(vals, 8, dtype)
        ks = generate(make_schedule(dag, list(range(len(args)))), ctype, 8, use_asm=False, coop=False)
        assert np.array_equal(run_on_host(ks, vals, n, dtype), want, equal_nan=True), (seed, ctype)
        # with the box as launch-wide scalars the whole decision moves to the uniform prologue
        one = [float(v[7]) for v in vals]
        ks1 = generate(make_schedule(dag, []), ctype, 8, use_asm=False, coop=False)
        assert np.array_equal(run_on_host(ks1, one, 1, dtype)[:, 0], want[:, 7], equal_nan=True)
