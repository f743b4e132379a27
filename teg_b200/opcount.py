"""Algorithmic operation count of a scheduled program (the roofline numerator).

Rule (fixed before measuring): every unique DAG value is charged once per execution of the region it
lives in after hash-consing + loop-invariant hoisting only (no algebraic rewriting):

    add, mul, div, compare, select, sqrt/sin/cos/asin/atan2      1 op each
    && / ||, loads, constants                                     0
    per sample-loop iteration: placement ``lb + step*(i+0.5)``    2 (shared by fused integrals)
                               accumulate ``out + f*step``         2 per integral

A region's count is multiplied by its trip count: per instance for the body of the kernel, N per loop
level for integrand bodies.  Guarded regions (``IfGroup`` branches) are weighted by how often the guard
is taken; ``count(schedule, N)`` returns the all-false / all-true bounds and a per-guard breakdown so
a caller that knows the dynamic guard frequencies (bench.py measures them) can form the exact figure.
Values published by the uniform prologue are charged once per launch, i.e. not at all per instance.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

from .schedule import Block, Eval, IfGroup, LoopGroup, Schedule

_CHARGE = {'add': 1, 'mul': 1, 'div': 1, 'lt': 1, 'sel': 1, 'sqrt': 1, 'sin': 1, 'cos': 1, 'asin': 1,
           'atan2': 1, 'and': 0, 'or': 0, 'in': 0, 'const': 0, 'bvar': 0}


@dataclass
class GuardCost:
    cond: int
    depth_trips: float          # how many times per instance the guard itself is evaluated
    then_ops: float             # ops executed per taken guard (nested guards counted as not taken)
    else_ops: float
    nested: List["GuardCost"] = field(default_factory=list)


@dataclass
class OpCount:
    unguarded: float            # ops per instance with every guard false (else branches)
    all_true: float             # ops per instance with every guard true
    guards: List[GuardCost]
    per_sample_inner: float     # ops charged per innermost-loop iteration of the deepest unguarded loop nest
    samples_unguarded: float    # integrand evaluations per instance outside guards


def _block(sch: Schedule, b: Block, N: int, trips: float, take: bool, guards: List[GuardCost], info: Dict) -> float:
    nodes = sch.dag.nodes
    total = 0.0
    for st in b.stmts:
        if isinstance(st, Eval):
            total += _CHARGE[nodes[st.node].op]
        elif isinstance(st, IfGroup):
            total += len(st.sels)                       # the selects themselves
            nested_t: List[GuardCost] = []
            nested_f: List[GuardCost] = []
            t_ops = _block(sch, st.then_block, N, trips, False, nested_t, {})
            f_ops = _block(sch, st.else_block, N, trips, False, nested_f, {})
            guards.append(GuardCost(st.cond, trips, t_ops, f_ops, nested_t + nested_f))
            if take:
                total += _block(sch, st.then_block, N, trips, True, [], {})
            else:
                total += f_ops
        elif isinstance(st, LoopGroup):
            sub: Dict = {}
            body = _block(sch, st.body, N, trips * N, take, guards, sub)
            per_iter = 2 + 2 * len(st.quads) + body
            total += N * per_iter
            if 'inner' not in sub:
                info['inner'] = 2 + 2 * len(st.quads) + body
                info['samples'] = N
            else:
                info['inner'] = sub['inner']
                info['samples'] = N * sub['samples']
    return total


def count(sch: Schedule, num_samples: int) -> OpCount:
    guards: List[GuardCost] = []
    info: Dict = {}
    low = _block(sch, sch.main, num_samples, 1.0, False, guards, info)
    high = _block(sch, sch.main, num_samples, 1.0, True, [], {})
    return OpCount(unguarded=low, all_true=high, guards=guards,
                   per_sample_inner=info.get('inner', 0.0), samples_unguarded=info.get('samples', 0.0))
