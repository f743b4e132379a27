"""CPU tests of the backend's host logic: exchange format, DAG, scheduling, code generation, NVRTC.

None of these run a kernel.  The lowering is checked for *value* parity by compiling the generated
instance functions for the host (tests/host_harness.py, a checker) and comparing with the oracle.
"""
import glob
import os

import numpy as np
import pytest

from conftest import load_program, program_text

from host_harness import run_on_host
from teg_b200.codegen import generate
from teg_b200.dag import build_dag
from teg_b200.schedule import make_schedule
from teg_b200.tegp import Program, from_teg
from teg_b200.workloads import SCENE_SAMPLES, bind, scene_values


def test_tegp_round_trip_and_sharing():
    for scene in SCENE_SAMPLES:
        text = program_text(scene)
        prog = Program.loads(text)
        assert prog.dumps() == text
    grad = load_program('tri_grad')
    assert grad.tree_size() > 20000 and len(grad.nodes) < 1200        # 22.5k-node tree, < 1k unique


def test_lang_mirror_builds_the_readme_program():
    from teg_b200.lang import IfElse, Teg, TegVar, Var
    x, t = TegVar('x'), Var('t', 0.5)
    prog = from_teg(Teg(0, 1, IfElse(x < t, 1, 0), x), [t])
    assert [n.op for n in prog.nodes].count('T') == 1
    assert prog.args == [f't_{t.uid}'] and prog.defaults[prog.args[0]] == 0.5
    with pytest.raises(ValueError):
        from_teg(object())


def test_schedule_structure_of_the_triangle_gradient():
    """12 delta integrals in the reference -> 3 fused loops (one per edge) under 6 guards."""
    dag = build_dag(load_program('tri_grad'))
    assert len(dag.outputs) == 6 and dag.out_is_tuple
    assert dag.stats()['live_nodes'] < 500
    sch = make_schedule(dag, [0, 1, 2, 3])
    loops = sch.loops()
    assert len(loops) == 3 and sorted(len(lp.quads) for lp in loops) == [4, 4, 4]
    assert len(sch.uniform_slots) > 30          # per-edge normalisation etc. leaves the per-pixel code
    # with every argument per-instance nothing is uniform
    sch_all = make_schedule(dag, list(range(10)))
    assert len(sch_all.uniform_slots) == 0


@pytest.mark.parametrize('scene', list(SCENE_SAMPLES))
def test_generated_code_matches_oracle_on_host(scene):
    from oracle import Oracle
    prog = load_program(scene)
    ns = SCENE_SAMPLES[scene]
    inputs = bind(prog.args, scene_values(scene, res=(20, 20), n=300))
    n = max(np.size(v) for v in inputs)
    arr = [i for i, v in enumerate(inputs) if np.ndim(v) > 0]
    sch = make_schedule(build_dag(prog), arr)
    for dtype, ctype in ((np.float32, 'float'), (np.float64, 'double')):
        ks = generate(sch, ctype, ns, use_asm=False, coop=False)       # inline PTX cannot be compiled for the host
        got = run_on_host(ks, inputs, n, dtype)
        want = Oracle(program_text(scene)).eval(inputs, ns, dtype)
        assert np.array_equal(got, want, equal_nan=True), f'{scene}/{ctype}'


def test_all_arguments_per_instance_and_all_scalar_variants():
    """The uniform / per-instance split must not change values: every argument streamed, and none."""
    from oracle import Oracle
    prog = load_program('tri_grad')
    inputs = bind(prog.args, scene_values('tri_grad', res=(16, 16)))
    n = 256
    dag = build_dag(prog)
    want = Oracle(program_text('tri_grad')).eval(inputs, 16, np.float32)
    ks = generate(make_schedule(dag, list(range(len(prog.args)))), 'float', 16, use_asm=False, coop=False)
    assert np.array_equal(run_on_host(ks, inputs, n, np.float32), want)
    one = [float(np.ravel(v)[137]) if np.ndim(v) else v for v in inputs]
    ks = generate(make_schedule(dag, []), 'float', 16, use_asm=False, coop=False)
    assert np.array_equal(run_on_host(ks, one, 1, np.float32)[:, 0], want[:, 137])


@pytest.mark.parametrize('scene', ['tri_primal', 'tri_grad', 'shaded_grad', 'nested3d_grad', 'circle_dt'])
def test_nvrtc_compiles_generated_kernels_for_sm100a(scene, tmp_path, monkeypatch):
    """NVRTC needs no GPU: every generated unit must compile for sm_100a (cubin comes back)."""
    from teg_b200 import runtime
    monkeypatch.setenv('TEGB200_CACHE_DIR', str(tmp_path))
    prog = load_program(scene)
    sch = make_schedule(build_dag(prog), [0, 1, 2, 3])
    ks = generate(sch, 'float', SCENE_SAMPLES[scene])
    mod = runtime.Module(ks.source)
    assert not mod.from_cache and len(mod.cubin()) > 1000
    again = runtime.Module(ks.source)
    assert again.from_cache and again.cubin() == mod.cubin()
    if scene.startswith('tri'):
        ksr = generate(sch, 'float', 16, grid_args=[0, 1, 2, 3])
        assert 'xe[nx0 + r + 1]' in ksr.source
        runtime.Module(ksr.source)


def test_nvrtc_error_surfaces_as_exception():
    from teg_b200 import runtime
    with pytest.raises(runtime.TegbError) as e:
        runtime.Module('this is not CUDA', use_cache=False)
    assert 'NVRTC' in str(e.value)


def test_opcount_rule_on_the_triangle():
    from teg_b200.opcount import count
    sch = make_schedule(build_dag(load_program('tri_primal')), [0, 1, 2, 3])
    oc = count(sch, 16)
    # per inner sample: 3 edges x (add, mul, add) + 3 compares + 1 select + 4 template = 17
    assert oc.per_sample_inner == 17 and oc.samples_unguarded == 256
    assert oc.unguarded == oc.all_true == 16 * (16 * 17 + 9 + 4)
    g = count(make_schedule(build_dag(load_program('tri_grad')), [0, 1, 2, 3]), 16)
    assert g.all_true > g.unguarded and len(g.guards) >= 3


def test_exp_is_rejected_like_the_reference():
    text = 'TEGP 1\nname e\nnodes 2\nC 1\nF Exp 0\nroot 1\nargs 0\n'
    with pytest.raises(AssertionError):
        build_dag(Program.loads(text))


def test_backend_registers_with_reference_registry_when_teg_is_importable():
    """drop-in: `import teg_b200` adds 'CUDA' to teg.eval.backend.BACKENDS (backend.py:6-10) without editing it"""
    import os
    import sys
    ref = os.environ.get('TEG_REFERENCE_PATH', '/root/reference')
    if not os.path.isdir(os.path.join(ref, 'teg')):
        pytest.skip('reference frontend not available on this machine')
    sys.path.insert(0, ref)
    sys.dont_write_bytecode = True
    import teg.eval.backend as ref_backend
    import teg_b200
    assert teg_b200.register()
    assert ref_backend.BACKENDS['CUDA'][0] is teg_b200.CUDA_EvalMode
    # a live reference expression lowers to the same TEGP text as the committed fixture
    from teg import IfElse, Teg, TegVar, Var
    from teg.lang.base import Var as RefVar
    saved = RefVar.global_uid
    RefVar.global_uid = 1000
    try:
        x, t = TegVar('x'), Var('t', 0.5)
        prog = from_teg(Teg(0, 1, IfElse(x < t, 1, 0), x), [t], name='readme_primal')
    finally:
        RefVar.global_uid = max(saved, RefVar.global_uid)
    assert prog.dumps() == program_text('readme_primal')


def test_unbound_arguments_follow_the_live_var_value_like_c_eval():
    """c_eval.py:136 reads ``teg_var.value`` at eval time: bind_variable() after the mode was built must count
    (tests/test.py:74-97 relies on it)."""
    from teg_b200.eval import CUDA_EvalMode
    from teg_b200.lang import IfElse, Teg, TegVar, Var
    x, t = TegVar('x'), Var('t', 0.5)
    mode = CUDA_EvalMode.__new__(CUDA_EvalMode)          # host-side logic only: no device needed for _resolve
    prog = from_teg(Teg(0, 1, IfElse(x < t, 1, 0), x))
    mode.program = prog
    mode.arglist = [(lab, prog.defaults.get(lab)) for lab in prog.args]
    assert mode._resolve({}) == [0.5]
    t.value = 0.75
    assert mode._resolve({}) == [0.75]
    assert mode._resolve({t: 0.1}) == [0.1]
    t.value = None
    with pytest.raises(AssertionError):
        mode._resolve({})


@pytest.mark.parametrize('scene', ['nested2d_value', 'nested3d_value', 'tri_primal', 'shaded_primal', 'circle'])
def test_unroll_and_jam_keeps_every_fold_exact(scene):
    """Long inner loops (N > 32) carry 4 outer samples through one inner pass; results must stay bit-identical."""
    from oracle import Oracle
    prog = load_program(scene)
    inputs = bind(prog.args, scene_values(scene, res=(6, 6), n=40))
    n = max(np.size(v) for v in inputs)
    arr = [i for i, v in enumerate(inputs) if np.ndim(v) > 0]
    sch = make_schedule(build_dag(prog), arr)
    for ns in (36, 40):
        ks = generate(sch, 'float', ns, use_asm=False, coop=False)
        assert '_3 = ' in ks.source                          # four copies were emitted
        want = Oracle(program_text(scene)).eval(inputs, ns, np.float32)
        assert np.array_equal(run_on_host(ks, inputs, n, np.float32), want, equal_nan=True)
        off = generate(sch, 'float', ns, use_asm=False, coop=False, jam=1)
        assert '_3 = ' not in off.source
        assert np.array_equal(run_on_host(off, inputs, n, np.float32), want, equal_nan=True)


NUMPY_VECTORS = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'vectors_numpy',
                                              '*.npz')))


@pytest.mark.parametrize('path', NUMPY_VECTORS, ids=[os.path.basename(p)[:-4] for p in NUMPY_VECTORS])
def test_trapezoid_lowering_matches_oracle_and_numpy_backend_on_host(path):
    """``integration_mode='trapezoidal_quadrature'`` (to_c.py:389-393 names it, numpy_eval.py:57-68 defines it):
    the generated code is bit-identical to the oracle's trapezoid restatement and reproduces the committed
    outputs of the reference's numpy backend (np.sum is pairwise, the fold here is left to right: <= 4 ulp)."""
    from oracle import Oracle
    z = np.load(path)
    scene, ns = os.path.basename(path)[:-4].rsplit('_N', 1)
    ns = int(ns)
    prog = load_program(scene)
    inputs = list(z['inputs'])
    n = inputs[0].shape[0]
    sch = make_schedule(build_dag(prog), list(range(len(inputs))))
    ks = generate(sch, 'double', ns, use_asm=False, integration_mode='trapezoidal_quadrature')
    got = run_on_host(ks, inputs, n, np.float64)
    want = Oracle(program_text(scene)).eval(inputs, ns, np.float64, quadrature='trapezoid')
    assert np.array_equal(got, want, equal_nan=True)
    ref = z['outputs']
    assert np.all(np.abs(got - ref) <= 1e-15 * np.maximum(1.0, np.abs(ref).max()))
    # fp32 lowering against the fp32 oracle
    ks32 = generate(sch, 'float', ns, use_asm=False, integration_mode='trapezoidal_quadrature')
    in32 = [np.asarray(v, np.float32) for v in inputs]
    assert np.array_equal(run_on_host(ks32, in32, n, np.float32),
                          Oracle(program_text(scene)).eval(in32, ns, np.float32, quadrature='trapezoid'), equal_nan=True)


# ---- domain culling (teg_b200/cull.py): interval-resolved instances run comparison-free nests, same bits ----------
CULLED_SCENES = ['tri_primal', 'raster_theta_primal', 'shaded_primal', 'shaded_grad', 'tri_color_grad',
                 'nested2d_value', 'nested3d_value', 'nested3d_grad', 'simple_line', 'circle', 'rect_hyperbola',
                 'bilinear_lerp']


@pytest.mark.parametrize('scene', CULLED_SCENES)
def test_culled_lowering_is_bit_identical_to_the_oracle_on_host(scene):
    from oracle import Oracle
    from teg_b200.cull import cull
    prog = load_program(scene)
    ns = SCENE_SAMPLES[scene]
    inputs = bind(prog.args, scene_values(scene, res=(40, 40), n=600))
    n = max(np.size(v) for v in inputs)
    arr = [i for i, v in enumerate(inputs) if np.ndim(v) > 0]
    dag = cull(build_dag(prog), ns)
    assert dag.culled, 'nothing was culled'
    assert any(nd.op == 'domtri' for nd in dag.nodes)
    sch = make_schedule(dag, arr)
    for dtype, ctype in ((np.float32, 'float'), (np.float64, 'double')):
        ks = generate(sch, ctype, ns, use_asm=False, coop=False)
        assert 'pz' in ks.source
        got = run_on_host(ks, inputs, n, dtype)
        want = Oracle(program_text(scene)).eval(inputs, ns, dtype)
        assert np.array_equal(got, want, equal_nan=True), f'{scene}/{ctype}'


def test_culling_decision_boundary_on_host():
    """Edges through pixel corners / sample abscissae, slivers, reversed, empty and non-finite boxes, huge
    coordinates: the interval decisions must never disagree with the per-sample comparisons."""
    from oracle import Oracle
    from teg_b200.cull import cull
    res = 32
    e = np.linspace(0, 1, res + 1, dtype=np.float32)
    xl, yl = np.meshgrid(e[:-1], e[:-1], indexing='ij')
    xu, yu = np.meshgrid(e[1:], e[1:], indexing='ij')
    big = np.float32(3e38)
    odd_xl = np.array([0.2, 0.3, 0.3, np.nan, 0.1, np.inf, 0.5, -big, 0.3], dtype=np.float32)
    odd_xu = np.array([0.3, 0.3, 0.2, 0.4, np.inf, 0.2, 0.5, big, 0.2], dtype=np.float32)
    odd_yl = np.array([0.2, 0.3, 0.4, 0.3, 0.2, 0.3, -np.inf, -big, 0.5], dtype=np.float32)
    odd_yu = np.array([0.3, 0.4, 0.3, 0.4, 0.3, np.nan, 0.1, big, 0.4], dtype=np.float32)
    boxes = [np.concatenate([a.ravel(), b]) for a, b in ((xl, odd_xl), (xu, odd_xu), (yl, odd_yl), (yu, odd_yu))]
    n = boxes[0].size
    cases = [(0.25, 0.25, 0.75, 0.25, 0.25, 0.75), (0.0, 0.0, 1.0, 0.0, 0.0, 1.0),
             (0.5, 0.5, 0.5 + 2 ** -12, 0.5, 0.5, 0.5 + 2 ** -12), (0.1, 0.1, 0.9, 0.1000001, 0.5, 0.10000005),
             (0.15, 0.2, 0.4, 0.9, 0.85, 0.3), (0.3, 0.3, 0.3, 0.3, 0.3, 0.3),
             (0.515625, 0.515625, 0.546875, 0.515625, 0.515625, 0.546875), (0.0, 0.0, big, 0.0, 0.0, big),
             (np.nan, 0.2, 0.8, 0.3, 0.4, 0.9), (np.inf, 0.2, 0.8, 0.3, 0.4, 0.9)]
    prog = load_program('tri_primal')
    dag = cull(build_dag(prog), 16)
    for arr in ([0, 1, 2, 3], list(range(10))):
        ks = generate(make_schedule(dag, arr), 'float', 16, use_asm=False, coop=False)
        for verts in cases:
            inputs = boxes + [np.float32(v) for v in verts]
            got = run_on_host(ks, inputs, n, np.float32)
            want = Oracle(program_text('tri_primal')).eval(inputs, 16, np.float32)
            assert np.array_equal(got, want, equal_nan=True), (arr, verts, np.count_nonzero(got != want))


def test_cull_transform_structure():
    from teg_b200.cull import cull
    dag0 = build_dag(load_program('tri_primal'))
    n0, out0 = len(dag0.nodes), dag0.outputs
    dag = cull(dag0, 16)
    assert len(dag0.nodes) == n0 and dag0.outputs == out0            # the input DAG is left alone
    (q, roots), = dag.culled
    assert len(roots) == 1
    top = dag.nodes[dag.outputs[0]]
    assert top.op == 'sel' and dag.nodes[top.args[0]].op == 'domis' and top.args[1] == q
    spec = dag.nodes[top.args[2]]
    assert spec.op == 'sel' and dag.nodes[spec.args[0]].attr == 'T'
    # the specialised "true" nest has no comparison left; the "false" nest collapsed to 0 + 0 * step
    ops_true = {dag.nodes[i].op for i in dag.reachable([spec.args[1]])}
    ops_false = {dag.nodes[i].op for i in dag.reachable([spec.args[2]])}
    assert 'lt' not in ops_true and 'quad' in ops_true
    assert 'quad' not in ops_false and 'lt' not in ops_false
    # short nests are left alone (N^depth below the threshold), and so are gradient programs made of 1-D delta loops
    assert not cull(build_dag(load_program('tri_primal')), 4).culled
    assert not cull(build_dag(load_program('tri_grad')), 16).culled
    # transcendental conditions have no interval rule: left alone
    assert all(len(r) for _, r in cull(build_dag(load_program('circle')), 20).culled)


def test_tile_cull_structure_and_spec_path_on_host(tmp_path, monkeypatch):
    """cull.tile_cull: render programs read one tri-state per tile condition.  With every tile 'unresolved' the
    program is the per-instance one; with a (brute-force) valid classification the substituted program must give
    the same bits.  The classification kernel itself is checked on the GPU (tests/test_gpu_cull.py)."""
    from oracle import Oracle
    from teg_b200 import runtime
    from teg_b200.cull import cull, tile_cull
    monkeypatch.setenv('TEGB200_CACHE_DIR', str(tmp_path))
    res = 48
    for scene, kinds in (('tri_primal', ['dom']), ('tri_grad', ['guard'] * 3), ('shaded_grad', ['dom'] + ['guard'] * 3)):
        prog = load_program(scene)
        names = [a.rsplit('_', 1)[0] for a in prog.args]
        box = [names.index(b) for b in ('x_lb', 'x_ub', 'y_lb', 'y_ub')]
        dag = tile_cull(cull(build_dag(prog), 16), box, 16)
        assert [k for k, _ in dag.tile_conds] == kinds
        n_args = len(prog.args)
        arr = box + [n_args + k for k in range(len(kinds))]
        sch = make_schedule(dag, arr)
        inputs = bind(prog.args, scene_values(scene, res=(res, res)))
        n = res * res
        want = Oracle(program_text(scene)).eval(inputs, 16, np.float32)
        ks = generate(sch, 'float', 16, grid_args=box, use_asm=False, coop=False)
        assert ks.tile_conds == len(kinds) and ks.tile_kernel == f'{scene}_tiles' and f'{scene}_tiles(' in ks.source
        # (1) everything unresolved -> the per-instance program
        got = run_on_host(ks, inputs, n, np.float32, tile_inputs=[np.full(n, 2.0)] * len(kinds))
        assert np.array_equal(got, want, equal_nan=True), scene
        # (2) the device unit compiles for sm_100a (tile kernel included)
        dev = generate(sch, 'float', 16, grid_args=box, reduce_outputs=scene.endswith('grad'))
        assert len(runtime.Module(dev.source).cubin()) > 1000
    # (3) a valid classification for the triangle value: pixels that are exactly full / exactly empty
    prog = load_program('tri_primal')
    dag = tile_cull(cull(build_dag(prog), 16), [0, 1, 2, 3], 16)
    ks = generate(make_schedule(dag, [0, 1, 2, 3, 10]), 'float', 16, grid_args=[0, 1, 2, 3], use_asm=False, coop=False)
    inputs = bind(prog.args, scene_values('tri_primal', res=(res, res)))
    want = Oracle(program_text('tri_primal')).eval(inputs, 16, np.float32)[0]
    tri = np.where(want == want.max(), 1.0, np.where(want == 0, 0.0, 2.0))
    # (a pixel with value 0 has every sample outside, one with the maximal value every sample inside)
    got = run_on_host(ks, inputs, res * res, np.float32, tile_inputs=[tri])
    assert np.array_equal(got[0], want)
    assert 0 < np.count_nonzero(tri == 1) and 0 < np.count_nonzero(tri == 0) and 0 < np.count_nonzero(tri == 2)


@pytest.mark.parametrize('scene,ns', [('tri_primal', 16), ('tri_grad', 16), ('shaded_grad', 8), ('nested3d_grad', 6),
                                      ('readme_primal', 50), ('rect_hyperbola_grad', 12)])
def test_monte_carlo_lowering_matches_oracle_on_host(scene, ns, tmp_path, monkeypatch):
    """``integration_mode='monte_carlo_uniform'``: generated code == oracle restatement, bit for bit (fused loops,
    hoisted integrals and Tup programs included: the points depend on instance, depth and sample index only)."""
    from oracle import Oracle
    from teg_b200 import runtime
    monkeypatch.setenv('TEGB200_CACHE_DIR', str(tmp_path))
    prog = load_program(scene)
    inputs = bind(prog.args, scene_values(scene, res=(12, 12), n=150))
    n = max(np.size(v) for v in inputs)
    arr = [i for i, v in enumerate(inputs) if np.ndim(v) > 0]
    sch = make_schedule(build_dag(prog), arr)
    for dtype, ctype in ((np.float32, 'float'), (np.float64, 'double')):
        ks = generate(sch, ctype, ns, use_asm=False, integration_mode='monte_carlo_uniform', mc_seed=2024)
        assert ks.takes_instance_index
        got = run_on_host(ks, inputs, n, dtype, inst_base=77)
        want = Oracle(program_text(scene)).eval(inputs, ns, dtype, quadrature='monte_carlo', seed=2024, inst_base=77)
        assert np.array_equal(got, want, equal_nan=True), (scene, ctype)
    dev = generate(sch, 'float', ns, integration_mode='monte_carlo_uniform', mc_seed=2024)
    assert len(runtime.Module(dev.source).cubin()) > 1000          # the device unit compiles for sm_100a
    with pytest.raises(AssertionError):
        generate(sch, 'float', ns, integration_mode='monte_carlo_uniform', team=4)


@pytest.mark.parametrize('scene', ['tri_grad', 'shaded_grad'])
def test_tile_bodies_with_exact_per_pixel_states_on_host(scene):
    """Every body of a tile-culled program (quiet / resolved / general), driven with a VALID classification: each
    pixel is its own tile, and its states are the per-pixel truth of the tile conditions (evaluated by a probe program
    lowered from the same DAG).  Whatever body the states select, the result is the oracle's, bit for bit."""
    from oracle import Oracle
    from teg_b200.cull import cull, tile_cull
    prog = load_program(scene)
    names = [a.rsplit('_', 1)[0] for a in prog.args]
    box = [names.index(b) for b in ('x_lb', 'x_ub', 'y_lb', 'y_ub')]
    ns = SCENE_SAMPLES[scene]
    res = 40
    inputs = bind(prog.args, scene_values(scene, res=(res, res)))
    n = res * res
    base = cull(build_dag(prog), ns)
    tiled = tile_cull(base, box, ns)
    kinds = [k for k, _ in tiled.tile_conds]
    assert 'guard' in kinds
    # probe: per-pixel value of every tile condition (guard -> 0 / 1, domain tri-state -> 0 / 1 / 2)
    probe = base.clone()
    probe.outputs = tuple(node if kind == 'dom' else probe.sel(node, probe.const('1'), probe.const('0'))
                          for kind, node in tiled.tile_conds)
    pks = generate(make_schedule(probe, box), 'float', ns, use_asm=False, coop=False, name='probe')
    states = run_on_host(pks, inputs, n, np.float32)
    assert set(np.unique(states)) <= {0.0, 1.0, 2.0} and (states == 1).any() and (states == 0).any()
    n_args = len(prog.args)
    sch = make_schedule(tiled, box + [n_args + k for k in range(len(kinds))])
    ks = generate(sch, 'float', ns, grid_args=box, use_asm=False, coop=False)
    labels = [lb for lb, _ in ks.tile_bodies]
    assert any(lb.startswith('quiet') for lb in labels) and 'resolved' in labels
    want = Oracle(program_text(scene)).eval(inputs, ns, np.float32)
    got = run_on_host(ks, inputs, n, np.float32, tile_inputs=list(states))
    assert np.array_equal(got, want, equal_nan=True), scene
    # how the pixels were routed: all three kinds of body must have been exercised
    guards = [i for i, k in enumerate(kinds) if k == 'guard']
    known = np.all(states != 2, axis=0)
    quiet = known & np.all(states[guards] == 0, axis=0)
    assert quiet.any() and (known & ~quiet).any() and (~known).any() == ('dom' in kinds and (states == 2).any())


def _host_vs_oracle(prog, ns, vals, array_args, culled=(False, True)):
    """generated code (host-compiled) == oracle, bit for bit, with and without culling, fp32 and fp64"""
    from oracle import Oracle
    from teg_b200.cull import cull
    n = max(int(np.size(v)) for v in vals)
    for dtype, ctype in ((np.float32, 'float'), (np.float64, 'double')):
        v = [np.asarray(x, dtype=dtype) if np.ndim(x) else dtype(x) for x in vals]
        want = Oracle(prog.dumps()).eval(v, ns, dtype)
        for c in culled:
            dag = build_dag(prog)
            dag = cull(dag, ns, min_samples=1) if c else dag
            ks = generate(make_schedule(dag, array_args), ctype, ns, use_asm=False, coop=False)
            got = run_on_host(ks, v, n, dtype)
            assert np.array_equal(got, want, equal_nan=True), (ctype, c, got, want)


def test_let_bound_integral_used_inside_an_integral_over_the_same_domain():
    """``let w = int_0^1 x*x*t dx in int_0^1 (y*y*t < 0.3 ? w * (y*y*t) : 0) dy``.  Both integrals
    have the bounds (0, 1) at nesting depth 0; with binders canonicalised by (bounds, depth) alone the inner
    integrand ``x*x*t`` was hash-consed with the outer ``y*y*t`` and read the outer sample."""
    from teg_b200.lang import Const, IfElse, LetIn, Teg, TegVar, Var
    x, y, t = TegVar('x'), TegVar('y'), Var('t', 0.1)
    w = Var('w')
    inner = Teg(Const(0), Const(1), x * x * t, x)
    body = Teg(Const(0), Const(1), IfElse(y * y * t < Const(0.3), w * (y * y * t), Const(0)), y)
    prog = from_teg(LetIn([w], [inner], body), [t], name='let_nested')
    dag = build_dag(prog)
    bvars = {n.attr for n in dag.nodes if n.op == 'quad'}
    assert len(bvars) == 2, 'the two integrals must not share a bound variable'
    ts = np.array([0.1, 0.5, 1.0, -0.3, 2.0])
    _host_vs_oracle(prog, 8, [ts], [0])
    _host_vs_oracle(prog, 8, [0.1], [])
    # one level down: the let sits inside an outer integral, its value and its body integrate over (0, 1) at depth 1
    z = TegVar('z')
    inner2 = Teg(Const(0), Const(1), x * z * t, x)
    body2 = Teg(Const(0), Const(1), IfElse(y * t < Const(0.3), w * (y * z * t), Const(0)), y)
    prog2 = from_teg(Teg(Const(0), Const(2), LetIn([w], [inner2], body2), z), [t], name='let_nested2')
    _host_vs_oracle(prog2, 6, [ts], [0])
    # integrals that do NOT nest still share their variable (fusion / sharing is kept)
    both = from_teg(Teg(Const(0), Const(1), x * t, x) + Teg(Const(0), Const(1), y * y * t, y), [t], name='siblings')
    assert len({n.attr for n in build_dag(both).nodes if n.op == 'quad'}) == 1


def test_adding_zero_to_a_negative_zero_is_not_folded_away():
    """the reference computes ``-0 + 0 = +0``; ``1 / (((-1 * t) * 0) + 0)`` is +inf, not -inf."""
    from teg_b200.lang import Const, Var
    t = Var('t', 1.0)
    prog = from_teg(Const(1) / (((Const(-1) * t) * Const(0)) + Const(0)), [t], name='negzero')
    _host_vs_oracle(prog, 4, [np.array([1.0, 2.0, -3.0, 0.0])], [0], culled=(False,))
    # ... while the folds that ARE exact stay: 0 + integral, x*x + 0, literal + 0
    from teg_b200.lang import Teg, TegVar
    x = TegVar('x')
    q = from_teg(Const(0) + Teg(Const(0), Const(1), x * t, x) + Const(0), [t], name='q0')
    d = build_dag(q)
    assert d.nodes[d.outputs[0]].op == 'quad'
    d2 = build_dag(from_teg((t * t) + Const(0), [t], name='sq0'))
    assert d2.nodes[d2.outputs[0]].op == 'mul'
    d3 = build_dag(from_teg(t + Const(0), [t], name='t0'))
    assert d3.nodes[d3.outputs[0]].op == 'add'


def test_multi_triangle_scene_classes_and_group_selection():
    """Config 5 host logic: nine window-disjoint classes whatever the number of bands (the class comes from the GLOBAL
    grid row), and a rank keeps exactly the triangles whose windows reach into its rows, class by class."""
    from teg_b200.raster import select_groups
    from teg_b200.workloads import multi_triangle_scene, windows_disjoint
    for bands in (1, 2, 3):
        sc = multi_triangle_scene(bands=bands, res=512)
        K = sc['verts'].shape[0]
        assert K == 256 * bands and len(sc['classes']) == 9
        assert sc['classes'][0][0] == 0 and sc['classes'][-1][1] == K
        assert all(a1 == b0 for (_, a1), (b0, _) in zip(sc['classes'], sc['classes'][1:]))
        assert all(windows_disjoint(sc['windows'][a:b]) for a, b in sc['classes'])
        w = sc['windows']
        seen = np.zeros(K, dtype=int)
        for rank in range(bands):
            rows = (512 * rank, 512 * (rank + 1))
            ids, local_classes, local_of = select_groups(w, sc['classes'], rows)
            assert sorted(ids) == [k for k in range(K) if w[k][0] < rows[1] and w[k][1] > rows[0]]
            assert all(local_of[g] == j for j, g in enumerate(ids)) and (local_of >= 0).sum() == len(ids)
            # local classes are contiguous, cover the held groups, and keep the class order
            assert local_classes[0][0] == 0 and local_classes[-1][1] == len(ids)
            cls_of = {g: c for c, (a, b) in enumerate(sc['classes']) for g in range(a, b)}
            order = [cls_of[g] for g in ids]
            assert order == sorted(order)
            seen[ids] += 1
        assert seen.min() >= 1                                   # every triangle is held by some rank
        assert bands == 1 or seen.max() == 2                     # windows across a band boundary: held by both
    # strong scaling: one band split into row slices
    sc = multi_triangle_scene(bands=1, res=512)
    held = [set(select_groups(sc['windows'], sc['classes'], (128 * r, 128 * (r + 1)))[0]) for r in range(4)]
    assert set().union(*held) == set(range(256)) and all(len(h) < 256 for h in held)
    ids, lc, lo = select_groups(sc['windows'], sc['classes'], (0, 0))
    assert ids == [] and lc == [] and (lo == -1).all()


def test_group_bins_list_every_overlapping_group_in_ascending_order():
    """raster.group_bins (the bins of tegb_render_groups_gather) against a brute-force overlap test per block."""
    from teg_b200.raster import group_bins
    rng = np.random.default_rng(3)
    res_y, rows, bc, tr = 300, (37, 251), 128, 16
    w = np.zeros((40, 4), dtype=np.int32)
    w[:, 0] = rng.integers(0, 280, 40)
    w[:, 1] = w[:, 0] + rng.integers(0, 90, 40)          # (some windows are empty, some leave the band)
    w[:, 2] = rng.integers(0, 290, 40)
    w[:, 3] = np.minimum(w[:, 2] + rng.integers(0, 200, 40), 320)
    begin, groups = group_bins(w, rows, res_y, bc, tr)
    nbx, nby = -(-res_y // bc), -(-(rows[1] - rows[0]) // tr)
    assert begin.dtype == np.int32 and groups.dtype == np.int32 and len(begin) == nbx * nby + 1
    assert begin[0] == 0 and begin[-1] == len(groups)
    for by in range(nby):
        for bx in range(nbx):
            a, b = rows[0] + by * tr, min(rows[0] + (by + 1) * tr, rows[1])
            c, d = bx * bc, min((bx + 1) * bc, res_y)
            want = [k for k in range(40) if max(w[k, 0], a) < min(w[k, 1], b) and max(w[k, 2], c) < min(w[k, 3], d)]
            got = list(groups[begin[by * nbx + bx]:begin[by * nbx + bx + 1]])
            assert got == want, (by, bx)
    e_begin, e_groups = group_bins(np.zeros((0, 4), dtype=np.int32), rows, res_y, bc, tr)
    assert not e_begin.any() and len(e_groups) == 1


def test_bins_of_row_bands_partition_the_pairs_of_the_whole_image():
    """N ranks, one band of rows each (MultiTriangleRasterizer per rank: select_groups, then group_bins over the held
    groups): every (pixel row, triangle) pair of the whole image is served by exactly one rank, and a rank's bins list its
    LOCAL group numbers in ascending global order (class order: the order the image adds them in)."""
    from teg_b200.raster import group_bins, select_groups
    from teg_b200.workloads import multi_triangle_scene
    res, bands, tr, bc = 128, 1, 16, 128
    sc = multi_triangle_scene(bands=bands, per_band_grid=4, res=res)
    w = np.asarray(sc['windows']).astype(np.int64)
    K = len(w)
    total_rows = res * bands
    served = np.zeros((total_rows, K), dtype=np.int32)
    for world in (1, 2, 4):
        served[:] = 0
        for rank in range(world):
            rows = (total_rows * rank // world, total_rows * (rank + 1) // world)
            ids, classes, local_of = select_groups(w, sc['classes'], rows)
            assert ids == sorted(ids) or all(ids[a:b] == sorted(ids[a:b]) for a, b in classes)
            begin, groups = group_bins(w[ids] if ids else np.zeros((0, 4), dtype=np.int64), rows, res, bc, tr)
            nbx = -(-res // bc)
            for b in range(len(begin) - 1):
                mine = groups[begin[b]:begin[b + 1]]
                assert list(mine) == sorted(mine)                     # ascending local = ascending (class, band) order
                by = b // nbx
                r_lo, r_hi = rows[0] + by * tr, min(rows[0] + (by + 1) * tr, rows[1])
                for g in mine:
                    k = ids[g]
                    a, c = max(w[k, 0], r_lo), min(w[k, 1], r_hi)
                    assert a < c                                      # the bin's groups do reach into the block
                    served[a:c, k] += 1
        want = np.zeros_like(served)
        for k in range(K):
            want[w[k, 0]:w[k, 1], k] = 1
        assert np.array_equal(served, want), world
