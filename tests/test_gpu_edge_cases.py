"""Edge cases and full-size properties of the CUDA path (empty / ragged inputs, maximum sizes,
non-finite values, reversed bounds; BASELINE sizes through size-independent properties)."""
import numpy as np
import pytest

from conftest import load_program, program_text

from teg_b200.workloads import TRIANGLE, bind, scene_values

pytestmark = pytest.mark.gpu


def _oracle(scene, inputs, ns, dtype=np.float32):
    from oracle import Oracle
    return Oracle(program_text(scene)).eval(inputs, ns, dtype, threads=8)


@pytest.mark.parametrize('n', [1, 2, 31, 32, 33, 127, 128, 129, 1000, 4097])
def test_ragged_batch_sizes(cuda_device, n):
    from teg_b200 import CUDA_EvalMode
    prog = load_program('tri_grad')
    inputs = bind(prog.args, scene_values('tri_grad', res=(80, 80)))
    sub = [np.asarray(v)[1234:1234 + n] if np.ndim(v) else v for v in inputs]
    got = np.asarray(CUDA_EvalMode(prog, num_samples=16).eval(dict(zip(prog.args, sub)))).reshape(6, n)
    assert np.array_equal(got, _oracle('tri_grad', sub, 16))


def test_empty_batch(cuda_device):
    import torch
    from teg_b200 import CUDA_EvalMode
    prog = load_program('readme_primal')
    mode = CUDA_EvalMode(prog, num_samples=50)
    out = mode.eval({prog.args[0]: np.zeros(0, dtype=np.float32)})
    assert np.asarray(out).shape == (0,)
    out = mode.eval({prog.args[0]: torch.zeros(0, device='cuda')})
    assert tuple(out.shape) == (0,)


def test_non_finite_and_degenerate_inputs_match_oracle(cuda_device):
    """NaN / inf / zero-width boxes / reversed bounds / a degenerate (zero-area) triangle: the IEEE results of the
    reference's operations, NaNs included, must come out identically."""
    from teg_b200 import CUDA_EvalMode
    for scene in ('tri_primal', 'tri_grad'):
        prog = load_program(scene)
        x_lb = np.array([0.2, 0.3, 0.3, np.nan, 0.1, np.inf, 0.5, 0.25], dtype=np.float32)
        x_ub = np.array([0.3, 0.3, 0.2, 0.4, np.inf, 0.2, 0.5, 0.26], dtype=np.float32)
        y_lb = np.array([0.2, 0.3, 0.4, 0.3, 0.2, 0.3, -np.inf, 0.25], dtype=np.float32)
        y_ub = np.array([0.3, 0.4, 0.3, 0.4, 0.3, np.nan, 0.1, 0.26], dtype=np.float32)
        tri = [TRIANGLE[k] for k in ('px0', 'py0', 'px1', 'py1', 'px2', 'py2')]
        for verts in (tri, [0.2, 0.2, 0.2, 0.2, 0.2, 0.2], [0.1, 0.1, 0.5, 0.5, 0.9, 0.9]):
            inputs = [x_lb, x_ub, y_lb, y_ub] + list(verts)
            got = np.asarray(CUDA_EvalMode(prog, num_samples=16).eval(dict(zip(prog.args, inputs))))
            want = _oracle(scene, inputs, 16)
            assert np.array_equal(got.reshape(want.shape), want, equal_nan=True), (scene, verts)


@pytest.mark.parametrize('ns', [1, 2, 3, 7, 33, 100])
def test_sample_counts(cuda_device, ns):
    from teg_b200 import CUDA_EvalMode
    for scene in ('nested2d_value', 'nested2d_grad', 'readme_primal'):
        prog = load_program(scene)
        inputs = bind(prog.args, scene_values(scene, n=257))
        got = np.asarray(CUDA_EvalMode(prog, num_samples=ns).eval(dict(zip(prog.args, inputs))))
        want = _oracle(scene, inputs, ns)
        assert np.array_equal(got.reshape(want.shape), want, equal_nan=True), (scene, ns)


def test_all_arguments_as_arrays_and_none(cuda_device):
    """Which arguments are per-instance is a launch-time choice; values must not depend on it."""
    from teg_b200 import CUDA_EvalMode
    prog = load_program('shaded_grad')
    inputs = bind(prog.args, scene_values('shaded_grad', res=(40, 40)))
    n = 1600
    full = [np.broadcast_to(np.asarray(v, dtype=np.float32), (n,)).copy() for v in inputs]
    mode = CUDA_EvalMode(prog, num_samples=16)
    a = np.asarray(mode.eval(dict(zip(prog.args, inputs))))
    b = np.asarray(mode.eval(dict(zip(prog.args, full))))
    assert np.array_equal(a, b, equal_nan=True)
    one = {lab: float(v[777]) for lab, v in zip(prog.args, full)}
    c = mode.eval(one)
    assert isinstance(c, list) and np.array_equal(np.float32(c), a[:, 777])


def test_full_size_properties_4096(cuda_device):
    """BASELINE size (4096 x 4096): properties that do not need the oracle on 16.7 M instances."""
    import torch
    from teg_b200 import CUDA_EvalMode
    res = 4096
    primal, grad = load_program('tri_primal'), load_program('tri_grad')
    mp, mg = CUDA_EvalMode(primal, num_samples=16), CUDA_EvalMode(grad, num_samples=16)
    params = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in primal.args[4:]}
    gparams = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in grad.args[4:]}
    img = mp.render(primal.args[:4], (res, res), ((0, 1), (0, 1)), params)[0]
    g_img = mg.render(grad.args[:4], (res, res), ((0, 1), (0, 1)), gparams)
    g_sum = mg.render(grad.args[:4], (res, res), ((0, 1), (0, 1)), gparams, reduce=True)
    torch.cuda.synchronize()
    # (1) every pixel value is a multiple of the sample weight 2^-32 in [0, pixel area]; total = triangle area
    px = 1.0 / res
    assert float(img.min()) >= 0 and float(img.max()) <= px * px * (1 + 1e-6)
    assert abs(float(img.double().sum()) - 0.2325) < 2e-6
    k = img.double() * (res * 16) ** 2
    assert float((k - k.round()).abs().max()) == 0.0
    # (2) interior pixels are exactly full, exterior exactly empty: only pixels crossed by an edge are partial,
    #     and exactly those can carry a gradient
    partial = (img > 0) & (img < px * px)
    has_grad = (g_img.abs().sum(dim=0) > 0)
    assert int(partial.sum()) < 3 * 2 * res and int((has_grad & ~partial & (img > 0) & (img < px * px)).sum()) == 0
    # (3) gradient of the area w.r.t. the vertices, analytically
    x0, y0, x1, y1, x2, y2 = (TRIANGLE[k] for k in ('px0', 'py0', 'px1', 'py1', 'px2', 'py2'))
    want = np.array([(y1 - y2) / 2, (x2 - x1) / 2, (y2 - y0) / 2, (x0 - x2) / 2, (y0 - y1) / 2, (x1 - x0) / 2])
    # (the delta integrals sample a conservative interval with 16 midpoints: ~1 % quadrature error per edge pixel,
    #  independent of the resolution -- the oracle / reference produce the same values, see the parity tests)
    assert np.allclose(g_sum.double().cpu().numpy(), want, atol=6e-3)
    assert np.allclose(g_img.double().sum(dim=(1, 2)).cpu().numpy(), g_sum.double().cpu().numpy(), rtol=1e-5, atol=1e-7)
    # (4) translation invariance of the sum: moving every vertex by (dx, dx) leaves the summed x/y gradients unchanged
    #     (sum of the three x-gradients = 0: the area does not change under translation)
    gs = g_sum.double().cpu().numpy()
    assert abs(gs[0] + gs[2] + gs[4]) < 1.2e-2 and abs(gs[1] + gs[3] + gs[5]) < 1.2e-2
    # (5) rows of the render path == the same rows through the array path (bit for bit), on a sample of rows
    rows = [0, 1237, 2048, 4095]
    for r in rows:
        vals = scene_values('tri_primal', res=(res, res), rows=(r, r + 1))
        a = np.asarray(mp.eval(dict(zip(primal.args, bind(primal.args, vals)))))
        assert np.array_equal(a, img[r].cpu().numpy())
    # (5b) ... and against the ORACLE itself, value image and per-pixel gradient images: 24 rows spread over the image
    #      (interior, exterior and edge tiles; the rows of unresolved tiles come from the worker queue of the kernel)
    rows = [int(r) for r in np.linspace(600, 3700, 24)]
    vals = [scene_values('tri_primal', res=(res, res), rows=(r, r + 1)) for r in rows]
    cat = {k: (np.concatenate([v[k] for v in vals]) if np.ndim(vals[0][k]) else vals[0][k]) for k in vals[0]}
    want = _oracle('tri_primal', bind(primal.args, cat), 16).reshape(len(rows), res)
    assert np.array_equal(img[rows].cpu().numpy(), want)
    want_g = _oracle('tri_grad', bind(grad.args, cat), 16).reshape(6, len(rows), res)
    assert np.array_equal(g_img[:, rows].cpu().numpy(), want_g)
    assert np.count_nonzero(want_g) > 100 and 0 < np.count_nonzero(want) < want.size
    # (6) sharding: two bands rendered separately tile the full image
    top = mp.render(primal.args[:4], (res, res), ((0, 1), (0, 1)), params, rows=(0, 1500))[0]
    bot = mp.render(primal.args[:4], (res, res), ((0, 1), (0, 1)), params, rows=(1500, res))[0]
    torch.cuda.synchronize()
    assert torch.equal(torch.cat([top, bot], dim=0), img)


def test_large_batch_readme_2_26(cuda_device):
    """maximum-size style batch: 2^26 instances of the README derivative, checked in closed form"""
    import torch
    from teg_b200 import CUDA_EvalMode
    prog = load_program('readme_deriv')
    n = 1 << 26
    t = torch.rand(n, device='cuda') * 1.5 - 0.25
    b = {prog.args[0]: t}
    if len(prog.args) > 1:
        b[prog.args[1]] = 1.0
    out = CUDA_EvalMode(prog, num_samples=50).eval(b)
    want = ((t > 0) & (t < 1)).float()
    assert torch.equal(out, want)


@pytest.mark.parametrize('n', [1, 2, 3, 4, 5, 7, 8, 9, 127, 1001, 4099])
def test_vectorised_streaming_kernels_handle_tails_and_misalignment(cuda_device, n):
    """loop-free programs run 4 instances per thread with 16-byte loads/stores; tails, odd sizes, multi-output
    rows that are not 16-byte aligned and offset tensors must all give the oracle's bits"""
    import torch
    from teg_b200 import CUDA_EvalMode
    for scene in ('readme_deriv', 'raster_theta'):
        prog = load_program(scene)
        inputs = bind(prog.args, scene_values(scene, res=(70, 70), n=4900))
        sub = [np.asarray(v)[11:11 + n] if np.ndim(v) else v for v in inputs]
        want = _oracle(scene, sub, 50 if scene.startswith('readme') else 10)
        mode = CUDA_EvalMode(prog, num_samples=50 if scene.startswith('readme') else 10)
        got = np.asarray(mode.eval(dict(zip(prog.args, sub)))).reshape(want.shape)          # host path
        assert np.array_equal(got, want, equal_nan=True)
        # device tensors at an odd element offset (4-byte aligned only): must fall back to scalar accesses
        big = {lab: (torch.as_tensor(np.asarray(v, dtype=np.float32), device='cuda') if np.ndim(v) else v)
               for lab, v in zip(prog.args, inputs)}
        off = {lab: (v[11:11 + n] if torch.is_tensor(v) else v) for lab, v in big.items()}
        got = mode.eval(off).cpu().numpy().reshape(want.shape)
        assert np.array_equal(got, want, equal_nan=True)
