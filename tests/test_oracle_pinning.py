"""Pin the CPU oracle (oracle/teg_oracle.c) to the reference before anything trusts it.

1. golden vectors: outputs of the reference's OWN generated C (tests/golden/make_golden.py), committed;
   the interpreter must reproduce them bit for bit, fp32 and fp64, all 20 scenes.
2. where oracle/_ref/ exists (this container; it also travels to the GPU box): live comparison with the
   reference's compiled C on fresh seeded inputs.
3. closed-form known answers of the reference's test-suite for the base language
   (/root/reference/tests/test.py:112-350), rebuilt with teg_b200.lang.
"""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT, program_text

from oracle import Oracle, RefLib, ref_lib_path
from teg_b200.workloads import SCENE_SAMPLES, bind, scene_values

VECTORS = sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'vectors', '*.npz')))


def test_fixtures_present():
    assert len(VECTORS) >= 37


@pytest.mark.parametrize('path', VECTORS, ids=[os.path.basename(p)[:-4] for p in VECTORS])
def test_oracle_reproduces_reference_vectors(path):
    z = np.load(path)
    name = os.path.basename(path)[:-4]
    scene, prec, ns = name.rsplit('_', 2)
    dtype = np.float32 if prec == 'f32' else np.float64
    got = Oracle(program_text(scene)).eval(list(z['inputs']), int(z['num_samples']), dtype, threads=2)
    want = z['outputs']
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want, equal_nan=True), f'{name}: {np.count_nonzero(got != want)} values differ'


HIGH_N = sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'vectors_high_n', '*.npz')))


@pytest.mark.parametrize('path', HIGH_N, ids=[os.path.basename(p)[:-4] for p in HIGH_N])
def test_oracle_reproduces_reference_vectors_at_high_sample_counts(path):
    """config 4 at the top of its sweep (N = 1024 / 4096 in 2-D, 128 / 256 in 3-D): the interpreter against the
    reference's compiled C.  The largest value programs (1.7e7 samples per instance) are checked on 16 of the 64
    instances to keep the CPU suite short; the GPU tests use all 64."""
    z = np.load(path)
    name = os.path.basename(path)[:-4]
    scene, prec, ns = name.rsplit('_', 2)
    dtype = np.float32 if prec == 'f32' else np.float64
    ns = int(z['num_samples'])
    m = 16 if (scene.endswith('value') and ns in (4096, 256)) else 64
    got = Oracle(program_text(scene)).eval([v[:m] for v in z['inputs']], ns, dtype, threads=os.cpu_count() or 2)
    want = z['outputs'][:, :m]
    assert np.array_equal(got, want, equal_nan=True), f'{name}: {np.count_nonzero(got != want)} values differ'


def test_high_n_fixtures_present():
    assert len(HIGH_N) >= 10


LIVE = [(s, p) for s in SCENE_SAMPLES for p in ('f32', 'f64')
        if os.path.exists(ref_lib_path(s, np.float32 if p == 'f32' else np.float64, SCENE_SAMPLES[s]))]


@pytest.mark.parametrize('scene,prec', LIVE)
def test_oracle_matches_compiled_reference_live(scene, prec):
    from teg_b200.tegp import Program
    dtype = np.float32 if prec == 'f32' else np.float64
    ns = SCENE_SAMPLES[scene]
    prog = Program.loads(program_text(scene))
    inputs = bind(prog.args, scene_values(scene, res=(37, 29), n=777))
    a = Oracle(program_text(scene)).eval(inputs, ns, dtype, threads=4)
    b = RefLib(scene, dtype, ns).eval(inputs, threads=4)
    assert np.array_equal(a, b, equal_nan=True)


# ---- closed forms from the reference's own tests (tests/test.py), base language only -------------
def _ev(expr, bindings=None, num_samples=50, dtype=np.float32):
    from teg_b200.tegp import from_teg, label_of
    prog = from_teg(expr)
    b = {label_of(k): v for k, v in (bindings or {}).items()}
    inputs = [b.get(lab, prog.defaults.get(lab)) for lab in prog.args]
    out = Oracle(prog.dumps()).eval(inputs, num_samples, dtype)
    return out[:, 0] if out.shape[0] > 1 else float(out[0, 0])


def test_reference_known_answers_base_language():
    from teg_b200.lang import Const, IfElse, LetIn, Sqr, Sqrt, Teg, TegVar, Tup, Var
    x, y = TegVar('x'), TegVar('y')
    a, b = Var('a', 0), Var('b', 1)
    # test.py:112-140 (TestArithmetic / TestIntegrations): linear integrals are exact for the midpoint rule
    assert abs(_ev(Teg(a, b, x, x), num_samples=100) - 0.5) < 1e-6
    assert abs(_ev(Teg(a, b, x * x, x), num_samples=1000) - 1 / 3) < 1e-5
    assert abs(_ev(Const(3) + Const(4) * Const(2)) - 11) < 1e-7
    assert abs(_ev(Const(1) / Const(4)) - 0.25) < 1e-7
    # nested integral, test.py:~190
    assert abs(_ev(Teg(a, b, Teg(a, b, x * y, x), y), num_samples=200) - 0.25) < 1e-6
    # conditionals: README primal = clamp(t, 0, 1)
    t = Var('t', 0.5)
    assert abs(_ev(Teg(0, 1, IfElse(x < t, 1, 0), x)) - 0.5) < 1e-6
    # allow_eq is ignored by the C backend: IfElse(y <= 1, 10, 20) at y = 1 gives 20
    yv = Var('yv', 1)
    assert _ev(IfElse(yv <= 1, 10, 20)) == 20.0
    # reversed bounds integrate negatively (Appendix A)
    tt = TegVar('tt')
    assert abs(_ev(Teg(1, 0, tt, tt), num_samples=100) + 0.5) < 1e-6
    # tuples + let, test.py:300-350
    v = _ev(Tup(Const(1), Const(2)) + Tup(Const(3), Const(4)))
    assert list(v) == [4.0, 6.0]
    z = Var('z')
    assert abs(_ev(LetIn([z], [Const(2) * t], z + z)) - 2.0) < 1e-7
    assert abs(_ev(Sqrt(Sqr(Const(3)) + Sqr(Const(4)))) - 5.0) < 1e-6


def test_oracle_rejects_exp_like_the_reference():
    """`Exp` has no C lowering in the reference (to_c.py:338-341): the oracle refuses it too."""
    text = 'TEGP 1\nname e\nnodes 2\nC 1\nF Exp 0\nroot 1\nargs 0\n'
    with pytest.raises((RuntimeError, ValueError)):
        Oracle(text).eval([], 1)


NUMPY_VECTORS = sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'vectors_numpy', '*.npz')))


@pytest.mark.parametrize('path', NUMPY_VECTORS, ids=[os.path.basename(p)[:-4] for p in NUMPY_VECTORS])
def test_trapezoid_oracle_reproduces_numpy_backend_vectors(path):
    """The oracle's trapezoid mode against outputs of the reference's numpy backend (numpy_eval.py:57-68,
    recorded by tests/golden/make_numpy_golden.py).  float64; np.sum adds pairwise, the oracle left to right,
    so agreement is to a few ulp of the largest term rather than bit for bit."""
    z = np.load(path)
    scene, ns = os.path.basename(path)[:-4].rsplit('_N', 1)
    got = Oracle(program_text(scene)).eval(list(z['inputs']), int(ns), np.float64, quadrature='trapezoid')
    want = z['outputs']
    assert got.shape == want.shape and np.count_nonzero(want) > 0
    assert np.all(np.abs(got - want) <= 1e-15 * np.maximum(1.0, np.abs(want).max()))


def test_numpy_fixtures_present():
    assert len(NUMPY_VECTORS) >= 10


# ---- monte_carlo_uniform: the generator is the published Philox4x32-10 ---------------------------------------------
def test_philox_known_answers():
    """Known-answer vectors of Random123 (kat_vectors, philox4x32 with 10 rounds): the oracle's generator is the
    published one (the mode itself has no reference implementation: 'parity unpinned' beyond the generator)."""
    import ctypes
    import oracle
    oracle.build()
    lib = ctypes.CDLL(oracle.LIB_PATH)
    out = (ctypes.c_uint * 4)()
    cases = [((0, 0), (0, 0, 0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
             ((0xffffffff, 0xffffffff), (0xffffffff,) * 4, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
             ((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
              (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for key, ctr, want in cases:
        lib.tego_philox4x32_10(ctypes.c_uint(key[0]), ctypes.c_uint(key[1]), *[ctypes.c_uint(c) for c in ctr], out)
        assert tuple(out) == want


def test_monte_carlo_oracle_converges_and_is_reproducible():
    from teg_b200.lang import IfElse, Teg, TegVar, Var
    from teg_b200.tegp import from_teg
    x, y, t = TegVar('x'), TegVar('y'), Var('t', 0.5)
    prog = from_teg(Teg(0, 1, Teg(0, 1, IfElse(x + y < t, x * y, 0), y), x), [t])
    o = Oracle(prog.dumps())
    tv = np.full(256, 0.9)
    exact = 0.9 ** 4 / 24                                    # integral of xy over the triangle x + y < t
    errs = []
    for ns in (16, 64, 256):
        v = o.eval([tv], ns, np.float64, quadrature='monte_carlo', seed=3)[0]
        assert np.array_equal(v, o.eval([tv], ns, np.float64, quadrature='monte_carlo', seed=3, threads=3)[0])
        assert len(set(v.tolist())) == v.size               # every instance draws its own points
        errs.append(np.sqrt(np.mean((v - exact) ** 2)))
    assert errs[0] > errs[1] > errs[2] and errs[2] < 0.25 * exact
    a = o.eval([tv[:8]], 64, np.float64, quadrature='monte_carlo', seed=3, inst_base=100)[0]
    b = o.eval([tv[:108]], 64, np.float64, quadrature='monte_carlo', seed=3)[0]
    assert np.array_equal(a, b[100:108])                    # the stream belongs to the instance index
    assert not np.array_equal(a, o.eval([tv[:8]], 64, np.float64, quadrature='monte_carlo', seed=4, inst_base=100)[0])
