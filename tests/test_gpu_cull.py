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


# ---- tile-level classification (cull.tile_cull + <name>_tiles) ------------------------------------------------------
def _render_all_ways(scene, res, bounds, params, rows=None, reduce=False):
    """render with tiles, with per-pixel culling only, and dense: three tensors that must be identical"""
    import torch
    from teg_b200 import CUDA_EvalMode
    prog = load_program(scene)
    names = [a.rsplit('_', 1)[0] for a in prog.args]
    box = [prog.args[names.index(b)] for b in ('x_lb', 'x_ub', 'y_lb', 'y_ub')]
    outs = []
    for opts in (dict(), dict(tiles=False), dict(cull=False)):
        mode = CUDA_EvalMode(prog, num_samples=16, **opts)
        ks, _ = mode.compiled([prog.args.index(b) for b in box], [prog.args.index(b) for b in box], reduce=reduce)
        assert (ks.tile_conds > 0) == (not opts), (scene, opts, ks.tile_conds)
        outs.append(mode.render(box, res, bounds, params, rows=rows, reduce=reduce))
    torch.cuda.synchronize()
    return outs


def test_tiled_render_equals_untiled_and_dense_at_full_size(cuda_device):
    import torch
    for scene, res, reduce in (('tri_primal', (4096, 4096), False), ('tri_grad', (4096, 4096), True),
                               ('tri_grad', (2048, 2048), False), ('shaded_grad', (1024, 1024), False),
                               ('shaded_grad', (1024, 1024), True)):
        prog = load_program(scene)
        vals = dict(zip(prog.args, bind(prog.args, scene_values(scene, res=(2, 2)))))
        params = {lab: float(v) for lab, v in vals.items() if np.ndim(v) == 0}
        a, b, c = _render_all_ways(scene, res, ((0, 1), (0, 1)), params, reduce=reduce)
        if reduce:
            # sums: unresolved tiles are walked by all warps of their block in the tiled kernels, i.e. a different
            # (still fixed) order of the same terms: equal up to fp32 summation rounding, and reproducible
            assert torch.equal(b, c) and torch.allclose(a, b, rtol=2e-5, atol=1e-7), (scene, res, a, b)
            a2, _, _ = _render_all_ways(scene, res, ((0, 1), (0, 1)), params, reduce=reduce)
            assert torch.equal(a, a2)
        else:
            assert torch.equal(a, b) and torch.equal(b, c), (scene, res, reduce)


def test_tiled_render_random_triangles_odd_grids_and_bands(cuda_device):
    """Random triangles (slivers, clockwise, sub-pixel, far outside), grids that are not multiples of the block or the
    tile, shifted / reversed bounds, and row bands that start inside a tile: tiles + per-pixel culling + dense
    renders agree bit for bit, and rows agree with the array path that the oracle tests cover."""
    import torch
    rng = np.random.default_rng(11)
    geoms = [((300, 200), ((0, 1), (0, 1)), None), ((257, 391), ((-0.5, 1.5), (0.25, 0.75)), (37, 200)),
             ((64, 1000), ((1, 0), (0, 1)), (5, 64)), ((1000, 130), ((0, 1), (2, -1)), (990, 1000))]
    for t in range(24):
        kind = t % 6
        if kind == 0:
            v = rng.uniform(0, 1, 6)
        elif kind == 1:
            c = rng.uniform(0.2, 0.8, 2)
            v = np.concatenate([c + rng.normal(scale=0.002, size=2) for _ in range(3)])
        elif kind == 2:
            p, q = rng.uniform(0, 1, 2), rng.uniform(0, 1, 2)
            v = np.concatenate([p, q, (p + q) / 2 + rng.normal(scale=1e-4, size=2)])
        elif kind == 3:
            v = rng.uniform(-3, 4, 6)
        elif kind == 4:
            v = np.array([0.25, 0.25, 0.75, 0.25, 0.25, 0.75]) + (rng.integers(0, 8, 6) / 64.0)
        else:
            v = rng.uniform(0.4, 0.6, 6)
        res, bounds, rows = geoms[t % len(geoms)]
        for scene in ('tri_primal', 'tri_grad'):
            prog = load_program(scene)
            params = {lab: float(np.float32(x)) for lab, x in zip(prog.args[4:], v)}
            a, b, c = _render_all_ways(scene, res, bounds, params, rows=rows)
            assert torch.equal(a, b) and torch.equal(b, c), (scene, t, res, bounds, rows)
            if scene == 'tri_grad':
                a, b, c = _render_all_ways(scene, res, bounds, params, rows=rows, reduce=True)
                scale = float(b.abs().max()) + 1e-12
                assert torch.equal(b, c) and float((a - b).abs().max()) <= 2e-5 * scale, (scene, t, 'reduce')
    # one geometry against the oracle through the array path
    from teg_b200.workloads import pixel_boxes
    res, bounds, rows = (257, 391), ((-0.5, 1.5), (0.25, 0.75)), (37, 200)
    prog = load_program('tri_grad')
    v = np.float32([0.15, 0.3, 0.95, 0.35, 0.4, 0.7])
    params = {lab: float(x) for lab, x in zip(prog.args[4:], v)}
    a, _, _ = _render_all_ways('tri_grad', res, bounds, params, rows=rows)
    box = pixel_boxes(res, bounds=bounds, rows=rows)
    inputs = list(box) + [np.float32(x) for x in v]
    want = _oracle('tri_grad', inputs, 16)
    assert np.array_equal(a.cpu().numpy().reshape(want.shape), want)


def test_tile_classification_resolves_most_tiles(cuda_device):
    """4096 x 4096 triangle: the classification leaves only the tiles an edge passes through (or near) unresolved."""
    import torch
    from teg_b200 import CUDA_EvalMode
    prog = load_program('tri_primal')
    mode = CUDA_EvalMode(prog, num_samples=16)
    params = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in prog.args[4:]}
    mode.render(prog.args[:4], (4096, 4096), ((0, 1), (0, 1)), params)
    torch.cuda.synchronize()
    ks, p = mode.compiled([0, 1, 2, 3], [0, 1, 2, 3])
    n_tiles = (4096 // ks.tile_cols) * (4096 // ks.tile_rows)
    tiles = p.tile_table(n_tiles * ks.tile_conds)
    frac = float(np.mean(tiles == 2))
    assert set(np.unique(tiles)) <= {0.0, 1.0, 2.0} and 0.01 < frac < 0.2, frac
    assert 0.1 < float(np.mean(tiles == 1)) < 0.3           # the triangle covers 23 % of the square


def test_interleaved_tile_rows_tile_the_image(cuda_device):
    """interleave=(stride, phase) renders tile rows phase, phase + stride, ...: the pieces of all phases reassemble
    to the contiguous render bit for bit (value image, per-pixel gradients), and the summed gradients add up."""
    import torch
    from teg_b200 import CUDA_EvalMode
    from teg_b200.eval import interleaved_rows
    for scene, res, bounds, rows in (('tri_primal', (300, 200), ((0, 1), (0, 1)), None),
                                     ('tri_grad', (257, 391), ((-0.5, 1.5), (0.25, 0.75)), (37, 200)),
                                     ('tri_primal', (1024, 1024), ((0, 1), (0, 1)), None)):
        prog = load_program(scene)
        params = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in prog.args[4:]}
        for cull in (True, False):
            mode = CUDA_EvalMode(prog, num_samples=16, cull=cull)
            full = mode.render(prog.args[:4], res, bounds, params, rows=rows).clone()
            n_rows = full.shape[1]
            ks, _ = mode.compiled([0, 1, 2, 3], [0, 1, 2, 3])
            for stride in (2, 3, 8):
                got = torch.full_like(full, float('nan'))
                for phase in range(stride):
                    idx = interleaved_rows(n_rows, ks.tile_rows, stride, phase)
                    part = mode.render(prog.args[:4], res, bounds, params, rows=rows, interleave=(stride, phase))
                    assert part.shape[1] == len(idx)
                    if idx:
                        got[:, torch.as_tensor(idx, device=got.device)] = part
                torch.cuda.synchronize()
                assert torch.equal(got, full), (scene, cull, stride)
            if scene == 'tri_grad':
                whole = mode.render(prog.args[:4], res, bounds, params, rows=rows, reduce=True).double()
                parts = sum(mode.render(prog.args[:4], res, bounds, params, rows=rows, reduce=True,
                                        interleave=(4, ph)).double().clone() for ph in range(4))
                assert torch.allclose(parts, whole, rtol=1e-5, atol=1e-8)


def test_render_with_nothing_to_do_and_trapezoid_render(cuda_device):
    """A rank whose phase has no tile row (or whose row range is empty) gets zero sums, not stale memory; and the
    render path of the trapezoid mode (no culling there) equals its array path."""
    import torch
    from teg_b200 import CUDA_EvalMode
    from teg_b200.workloads import pixel_boxes
    prog = load_program('tri_grad')
    params = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in prog.args[4:]}
    mode = CUDA_EvalMode(prog, num_samples=16)
    out = torch.full((6,), 7.0, device='cuda')
    mode.render(prog.args[:4], (40, 64), ((0, 1), (0, 1)), params, reduce=True, interleave=(8, 5), out=out)   # 3 tile rows
    torch.cuda.synchronize()
    assert torch.equal(out, torch.zeros(6, device='cuda'))
    out.fill_(7.0)
    mode.render(prog.args[:4], (40, 64), ((0, 1), (0, 1)), params, reduce=True, rows=(9, 9), out=out)
    torch.cuda.synchronize()
    assert torch.equal(out, torch.zeros(6, device='cuda'))
    trap = CUDA_EvalMode(prog, num_samples=12, integration_mode='trapezoidal_quadrature')
    res, rows = (70, 150), (3, 66)
    img = trap.render(prog.args[:4], res, ((0, 1), (0, 1)), params, rows=rows)
    box = pixel_boxes(res, rows=rows)
    arr = trap.eval(dict(zip(prog.args, list(box) + [params[lab] for lab in prog.args[4:]])))
    torch.cuda.synchronize()
    assert np.array_equal(img.cpu().numpy().reshape(6, -1), np.asarray(arr))


def test_prepared_render_rebinds_scalars(cuda_device):
    """A bound render re-run with other vertices must re-classify its tiles and re-publish its uniform table:
    results equal those of a fresh render, for images and for summed gradients."""
    import torch
    from teg_b200 import CUDA_EvalMode
    rng = np.random.default_rng(3)
    for scene, reduce in (('tri_primal', False), ('tri_grad', True), ('tri_grad', False)):
        prog = load_program(scene)
        mode = CUDA_EvalMode(prog, num_samples=16)
        first = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in prog.args[4:]}
        call = mode.prepare_render(prog.args[:4], (512, 384), ((0, 1), (0, 1)), first, reduce=reduce)
        stream = torch.cuda.current_stream().cuda_stream
        for _ in range(4):
            v = rng.uniform(0.05, 0.95, 6)
            params = {lab: float(np.float32(x)) for lab, x in zip(prog.args[4:], v)}
            call.set_scalars([params[prog.args[a]] for a in call.ks.scalar_args])
            got = call.run(stream).clone()
            want = CUDA_EvalMode(prog, num_samples=16).render(prog.args[:4], (512, 384), ((0, 1), (0, 1)),
                                                              params, reduce=reduce)
            torch.cuda.synchronize()
            assert torch.equal(got, want), (scene, reduce)
