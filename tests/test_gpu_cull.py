"""Domain culling (teg_b200/cull.py) on the GPU: results must be bit-identical to the dense kernels and to the oracle.

The culled kernels evaluate the select conditions of a loop nest over the pixel box with interval arithmetic and
run comparison-free nests where the conditions are uniform.  These tests (1) hold the culled path to the oracle on
every scene that has something to cull, (2) compare culled and dense kernels bit for bit at the BASELINE size,
(3) aim at the decision boundary: edges through pixel corners and sample points, sliver triangles, degenerate,
reversed and non-finite boxes, and (4) check that culling actually resolves the instances it should.
"""
import numpy as np
import pytest

from conftest import load_program, program_text

from teg_b200.workloads import SCENE_SAMPLES, TRIANGLE, bind, scene_values

pytestmark = pytest.mark.gpu

CULLED = ['tri_primal', 'raster_theta_primal', 'shaded_primal', 'shaded_grad', 'tri_color_primal', 'tri_color_grad',
          'nested2d_value', 'nested3d_value', 'nested3d_grad', 'simple_line', 'rect_hyperbola', 'bilinear_lerp']


def _oracle(scene, inputs, ns, dtype=np.float32):
    from oracle import Oracle
    return Oracle(program_text(scene)).eval(inputs, ns, dtype, threads=8)


@pytest.mark.parametrize('dtype', [np.float32, np.float64], ids=['f32', 'f64'])
@pytest.mark.parametrize('scene', CULLED)
def test_culled_and_dense_kernels_match_the_oracle(cuda_device, scene, dtype):
    from teg_b200 import CUDA_EvalMode
    prog = load_program(scene)
    ns = SCENE_SAMPLES[scene]
    inputs = bind(prog.args, scene_values(scene, res=(96, 96), n=6000))
    want = _oracle(scene, inputs, ns, dtype)
    fw = 'float' if dtype == np.float32 else 'double'
    for cull in (True, False):
        mode = CUDA_EvalMode(prog, num_samples=ns, float_width=fw, cull=cull)
        assert bool(getattr(mode.sched_dag, 'culled', [])) == cull
        got = np.asarray(mode.eval(dict(zip(prog.args, inputs)))).reshape(want.shape)
        assert np.array_equal(got, want, equal_nan=True), (scene, cull, np.count_nonzero(got != want))


def test_full_size_culled_equals_dense_bit_for_bit(cuda_device):
    """4096 x 4096 (BASELINE size): the culled and the dense render kernels produce the same image, bit for bit,
    for the triangle value, the shaded value and (summed in-kernel) the shaded gradient."""
    import torch
    from teg_b200 import CUDA_EvalMode
    res = 4096
    for scene in ('tri_primal', 'shaded_primal'):
        prog = load_program(scene)
        vals = dict(zip(prog.args, bind(prog.args, scene_values(scene, res=(2, 2)))))
        params = {lab: float(vals[lab]) for lab in prog.args[4:]}
        imgs = []
        for cull in (True, False):
            mode = CUDA_EvalMode(prog, num_samples=16, cull=cull)
            imgs.append(mode.render(prog.args[:4], (res, res), ((0, 1), (0, 1)), params)[0])
        torch.cuda.synchronize()
        assert torch.equal(imgs[0], imgs[1]), scene
        assert abs(float(imgs[0].double().sum()) - (0.2325 if scene == 'tri_primal' else float(imgs[1].double().sum()))) < 2e-6
    prog = load_program('shaded_grad')
    names = [a.rsplit('_', 1)[0] for a in prog.args]
    box = [prog.args[names.index(b)] for b in ('x_lb', 'x_ub', 'y_lb', 'y_ub')]
    vals = dict(zip(prog.args, bind(prog.args, scene_values('shaded_grad', res=(2, 2)))))
    params = {lab: float(vals[lab]) for lab in prog.args if lab not in box}
    outs = []
    for cull in (True, False):
        mode = CUDA_EvalMode(prog, num_samples=16, cull=cull)
        outs.append(mode.render(box, (1024, 1024), ((0, 1), (0, 1)), params))
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])


def _tri_inputs(x_lb, x_ub, y_lb, y_ub, verts):
    return [np.asarray(x_lb, np.float32), np.asarray(x_ub, np.float32), np.asarray(y_lb, np.float32),
            np.asarray(y_ub, np.float32)] + [np.float32(v) for v in verts]


@pytest.mark.parametrize('verts', [
    (0.25, 0.25, 0.75, 0.25, 0.25, 0.75),            # edges along pixel boundaries and through pixel corners
    (0.0, 0.0, 1.0, 0.0, 0.0, 1.0),                  # the diagonal passes through sample points' neighbourhood
    (0.5, 0.5, 0.5 + 2 ** -12, 0.5, 0.5, 0.5 + 2 ** -12),    # smaller than a pixel
    (0.1, 0.1, 0.9, 0.1000001, 0.5, 0.10000005),     # sliver
    (0.15, 0.2, 0.4, 0.9, 0.85, 0.3),                # clockwise: empty everywhere
    (0.3, 0.3, 0.3, 0.3, 0.3, 0.3),                  # degenerate
    (0.53125, 0.53125, 0.53125 + 1 / 64, 0.53125, 0.53125, 0.53125 + 1 / 64),   # vertices on sample abscissae (N = 16, 64 px)
], ids=['axis', 'diag', 'subpixel', 'sliver', 'cw', 'point', 'on_samples'])
def test_decision_boundary_cases(cuda_device, verts):
    from teg_b200 import CUDA_EvalMode
    res = 64
    e = np.linspace(0, 1, res + 1, dtype=np.float32)
    xl, yl = np.meshgrid(e[:-1], e[:-1], indexing='ij')
    xu, yu = np.meshgrid(e[1:], e[1:], indexing='ij')
    for scene in ('tri_primal', 'tri_grad'):
        prog = load_program(scene)
        inputs = _tri_inputs(xl.ravel(), xu.ravel(), yl.ravel(), yu.ravel(), verts)
        want = _oracle(scene, inputs, 16)
        got = np.asarray(CUDA_EvalMode(prog, num_samples=16).eval(dict(zip(prog.args, inputs)))).reshape(want.shape)
        assert np.array_equal(got, want, equal_nan=True), (scene, np.count_nonzero(got != want))


def test_non_finite_reversed_and_degenerate_boxes(cuda_device):
    from teg_b200 import CUDA_EvalMode
    big = np.float32(3e38)
    x_lb = np.array([0.2, 0.3, 0.3, np.nan, 0.1, np.inf, 0.5, -big, 0.3, 0.0], dtype=np.float32)
    x_ub = np.array([0.3, 0.3, 0.2, 0.4, np.inf, 0.2, 0.5, big, 0.2, 1.0], dtype=np.float32)
    y_lb = np.array([0.2, 0.3, 0.4, 0.3, 0.2, 0.3, -np.inf, -big, 0.5, 0.0], dtype=np.float32)
    y_ub = np.array([0.3, 0.4, 0.3, 0.4, 0.3, np.nan, 0.1, big, 0.4, 1.0], dtype=np.float32)
    tri = [TRIANGLE[k] for k in ('px0', 'py0', 'px1', 'py1', 'px2', 'py2')]
    for scene in ('tri_primal', 'shaded_primal'):
        prog = load_program(scene)
        extra = [0.2, 0.5, 0.3, 0.05] if scene == 'shaded_primal' else []
        for verts in (tri, [0.0, 0.0, big, 0.0, 0.0, big], [np.nan, 0.2, 0.8, 0.3, 0.4, 0.9],
                      [np.inf, 0.2, 0.8, 0.3, 0.4, 0.9], [0.2, 0.2, 0.2, 0.2, 0.2, 0.2]):
            inputs = [x_lb, x_ub, y_lb, y_ub] + [np.float32(v) for v in list(verts) + extra]
            assert len(inputs) == len(prog.args)
            want = _oracle(scene, inputs, 16)
            for ns_mode in (dict(), dict(coop=False)):
                got = np.asarray(CUDA_EvalMode(prog, num_samples=16, **ns_mode).eval(dict(zip(prog.args, inputs))))
                assert np.array_equal(got.reshape(want.shape), want, equal_nan=True), (scene, verts)


def test_random_triangles_against_oracle(cuda_device):
    """2000 random (pixel box, triangle) pairs with every argument per instance: boxes of very different sizes,
    triangles of both orientations, many of them grazing the box."""
    from teg_b200 import CUDA_EvalMode
    rng = np.random.default_rng(7)
    n = 2000
    c = rng.uniform(0.2, 0.8, size=(2, n))
    w = 10.0 ** rng.uniform(-4, -0.5, size=(2, n))
    x_lb, y_lb = (c - w / 2).astype(np.float32)
    x_ub, y_ub = (c + w / 2).astype(np.float32)
    # vertices near the box so that edges cross, touch or just miss it
    v = (c[:, None, :] + rng.normal(scale=1.5, size=(2, 3, n)) * w[:, None, :]).astype(np.float32)
    inputs = [x_lb, x_ub, y_lb, y_ub, v[0, 0], v[1, 0], v[0, 1], v[1, 1], v[0, 2], v[1, 2]]
    for scene in ('tri_primal', 'tri_grad'):
        prog = load_program(scene)
        want = _oracle(scene, inputs, 16)
        got = np.asarray(CUDA_EvalMode(prog, num_samples=16).eval(dict(zip(prog.args, inputs)))).reshape(want.shape)
        assert np.array_equal(got, want, equal_nan=True), (scene, np.count_nonzero(got != want))
    assert 0.05 < np.mean(want.sum(axis=0) != 0) < 0.9         # the batch really is a mix


def test_culling_resolves_what_it_should(cuda_device):
    """Every pixel of the 256 x 256 triangle grid that no edge comes near must be resolved by the interval test
    (checked through the exported tri-state), and no resolved pixel is a partial one."""
    import torch
    from teg_b200 import CUDA_EvalMode
    from teg_b200.cull import cull
    from teg_b200.dag import build_dag
    from teg_b200.tegp import Program
    prog = load_program('tri_primal')
    # a probe program: same DAG, output replaced by the tri-state of the (single) condition
    dag = cull(build_dag(prog), 16)
    tri_nodes = [i for i, n in enumerate(dag.nodes) if n.op == 'domtri']
    assert len(tri_nodes) == 1
    mode = CUDA_EvalMode(prog, num_samples=16)
    mode.sched_dag = dag
    dag.outputs = (tri_nodes[0],)
    res = 256
    params = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in prog.args[4:]}
    tri = mode.render(prog.args[:4], (res, res), ((0, 1), (0, 1)), params)[0]
    img = CUDA_EvalMode(prog, num_samples=16).render(prog.args[:4], (res, res), ((0, 1), (0, 1)), params)[0]
    torch.cuda.synchronize()
    px = 1.0 / res
    full, empty = img == img.max(), img == 0
    assert torch.all(full[tri == 1]) and torch.all(empty[tri == 0])
    partial = ~(full | empty)
    assert torch.all(tri[partial] == 2)
    unresolved = int((tri == 2).sum())
    assert int(partial.sum()) <= unresolved <= 3 * int(partial.sum()) + 64, (unresolved, int(partial.sum()))
    assert px > 0
