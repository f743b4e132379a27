"""Config 5: grouped launches over (pixel, triangle) pairs -- forward image, per-triangle gradients, bands."""
import numpy as np
import pytest

from conftest import load_program, program_text

from teg_b200.workloads import multi_triangle_scene, pixel_boxes, windows_disjoint

pytestmark = pytest.mark.gpu


def _oracle_scene(scene, res, bands, dL=None):
    """image [res*bands, res] accumulated class by class (fp32, same order as the device), and the float64
    per-triangle gradient sums [K, 7], from the CPU oracle evaluated on every triangle's window."""
    from oracle import Oracle
    prim = Oracle(program_text('tri_color_primal'))
    grad = Oracle(program_text('tri_color_grad'))
    rows_total = res * bands
    img = np.zeros((rows_total, res), dtype=np.float32)
    K = scene['verts'].shape[0]
    g = np.zeros((K, 7))
    gabs = np.zeros((K, 7))
    xs = np.linspace(0.0, float(bands), rows_total + 1)
    ys = np.linspace(0.0, 1.0, res + 1)
    for g0, g1 in scene['classes']:
        for k in range(g0, g1):
            r0, r1, c0, c1 = scene['windows'][k]
            nx = np.repeat(np.arange(r0, r1), c1 - c0)
            ny = np.tile(np.arange(c0, c1), r1 - r0)
            box = [xs[nx], xs[nx + 1], ys[ny], ys[ny + 1]]
            par = list(scene['verts'][k]) + [scene['col'][k]]
            v = prim.eval(box + par, 16, np.float32, threads=8)[0]
            img[nx, ny] = img[nx, ny] + v
            if dL is not None:
                gk = grad.eval([dL[nx, ny]] + box + par, 16, np.float32, threads=8)
                g[k] = gk.astype(np.float64).sum(axis=1)
                gabs[k] = np.abs(gk.astype(np.float64)).sum(axis=1)
    return img, g, gabs


@pytest.mark.parametrize('bands', [1, 2])
def test_multi_triangle_forward_backward_match_oracle(cuda_device, bands):
    import torch
    from teg_b200.raster import MultiTriangleRasterizer
    res = 96
    scene = multi_triangle_scene(bands=bands, per_band_grid=4, res=res)
    assert all(windows_disjoint(scene['windows'][a:b]) for a, b in scene['classes'])
    rng = np.random.default_rng(5)
    dL = rng.uniform(-1, 1, size=(res * bands, res)).astype(np.float32)
    want_img, want_g, want_gabs = _oracle_scene(scene, res, bands, dL)
    imgs, grads = [], []
    for rank in range(bands):
        r = MultiTriangleRasterizer(load_program('tri_color_primal'), load_program('tri_color_grad'), res,
                                    bands=bands, rank=rank)
        br = scene['band_ranges']
        active = (br[max(rank - 1, 0)][0], br[min(rank + 1, bands - 1)][1]) if bands > 1 else None
        r.set_scene(scene['verts'], scene['col'], scene['windows'], scene['classes'], scene['max_rows'],
                    scene['max_cols'], active=active)
        img = r.forward()
        band = torch.as_tensor(dL[res * rank: res * (rank + 1)]).cuda().contiguous()
        g = r.backward(band)
        torch.cuda.synchronize()
        imgs.append(img[0].cpu().numpy())
        grads.append(g.double().cpu().numpy())
    got_img = np.concatenate(imgs, axis=0)
    assert np.array_equal(got_img, want_img), f'{np.count_nonzero(got_img != want_img)} pixels differ'
    got_g = sum(grads)                       # what the all-reduce computes
    # dL has random signs, so the sums cancel: the fp32 summation error scales with sum |terms|, not |sum|
    assert np.all(np.abs(got_g - want_g) <= 2e-6 * want_gabs + 1e-12)
    assert np.abs(want_g).max() > 1e-3 and np.count_nonzero(want_img) > 100


def test_captured_step_follows_parameter_updates(cuda_device):
    """A CUDA graph of forward + dL + backward (MultiTriangleRasterizer.capture) reads the triangles from device
    arrays: after update_params the replay equals a fresh launch-by-launch step, bit for bit."""
    import torch
    from teg_b200.raster import MultiTriangleRasterizer
    res = 128
    scene = multi_triangle_scene(bands=1, per_band_grid=4, res=res)
    r = MultiTriangleRasterizer(load_program('tri_color_primal'), load_program('tri_color_grad'), res)
    r.set_scene(scene['verts'], scene['col'], scene['windows'], scene['classes'], scene['max_rows'], scene['max_cols'])
    img = torch.empty((1, res, res), dtype=torch.float32, device='cuda')
    dL = torch.empty((res, res), dtype=torch.float32, device='cuda')
    g = torch.empty((scene['verts'].shape[0], 7), dtype=torch.float32, device='cuda')
    target = torch.full((res, res), 0.25, device='cuda')

    def body(sp):
        r.forward(out=img, stream=sp)
        torch.sub(img[0], target, out=dL)
        r.backward(dL, out=g, stream=sp)

    graph = r.capture(body)
    rng = np.random.default_rng(9)
    for _ in range(3):
        verts = torch.as_tensor(scene['verts'] + rng.normal(scale=2e-3, size=scene['verts'].shape), dtype=torch.float32).cuda()
        col = torch.as_tensor(rng.uniform(0.2, 1.0, size=scene['col'].shape), dtype=torch.float32).cuda()
        r.update_params(verts, col)
        graph.replay()
        torch.cuda.synchronize()
        got_img, got_g = img.clone(), g.clone()
        body(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert torch.equal(got_img, img) and torch.equal(got_g, g)
        assert float(img.sum()) > 0 and float(g.abs().sum()) > 0


# ---- BASELINE size: 4096 x 4096 per band, 256 triangles per band -------------------------------
def _oracle_rows(scene, res, bands, r_lo, r_hi):
    """image rows [r_lo, r_hi) of the full scene, accumulated class by class like the device (fp32): every triangle
    whose window reaches into the rows contributes; only the pixels inside the rows are evaluated."""
    from oracle import Oracle
    prim = Oracle(program_text('tri_color_primal'))
    rows_total = res * bands
    xs = np.linspace(0.0, float(bands), rows_total + 1)
    ys = np.linspace(0.0, 1.0, res + 1)
    img = np.zeros((r_hi - r_lo, res), dtype=np.float32)
    used = 0
    for g0, g1 in scene['classes']:
        for k in range(g0, g1):
            r0, r1, c0, c1 = scene['windows'][k]
            a0, a1 = max(r0, r_lo), min(r1, r_hi)
            if a0 >= a1:
                continue
            used += 1
            nx = np.repeat(np.arange(a0, a1), c1 - c0)
            ny = np.tile(np.arange(c0, c1), a1 - a0)
            par = list(scene['verts'][k]) + [scene['col'][k]]
            v = prim.eval([xs[nx], xs[nx + 1], ys[ny], ys[ny + 1]] + par, 16, np.float32, threads=16)[0]
            img[nx - r_lo, ny] = img[nx - r_lo, ny] + v
    return img, used


def _oracle_grad(scene, res, bands, dL, ks):
    """float64 sums (and sums of magnitudes) of the per-pixel reverse derivatives over the FULL windows of triangles ks"""
    from oracle import Oracle
    grad = Oracle(program_text('tri_color_grad'))
    rows_total = res * bands
    xs = np.linspace(0.0, float(bands), rows_total + 1)
    ys = np.linspace(0.0, 1.0, res + 1)
    g, gabs = {}, {}
    for k in ks:
        r0, r1, c0, c1 = scene['windows'][k]
        nx = np.repeat(np.arange(r0, r1), c1 - c0)
        ny = np.tile(np.arange(c0, c1), r1 - r0)
        par = list(scene['verts'][k]) + [scene['col'][k]]
        gk = grad.eval([dL[nx, ny], xs[nx], xs[nx + 1], ys[ny], ys[ny + 1]] + par, 16, np.float32, threads=16)
        g[k] = gk.astype(np.float64).sum(axis=1)
        gabs[k] = np.abs(gk.astype(np.float64)).sum(axis=1)
    return g, gabs


@pytest.mark.parametrize('bands', [1, 2])
def test_config5_at_baseline_size_against_the_oracle(cuda_device, bands):
    """4096 x 4096 pixels and 256 triangles per band, the grouped + tile-culled path exactly as bench.py launches it.
    Image: a strip of 384 rows (for two bands: the strip across the band boundary, where windows are clipped to the
    ranks' bands) equals the oracle's class-ordered accumulation bit for bit.  Gradients: the [K, 7] sums of >= 16
    triangles over their full windows (for two bands: the triangles whose windows straddle the boundary, summed over
    the ranks like the all-reduce does) within fp32 summation error of the oracle's float64 sums."""
    import torch
    from teg_b200.raster import MultiTriangleRasterizer
    res = 4096
    scene = multi_triangle_scene(bands=bands, res=res)
    K = scene['verts'].shape[0]
    assert K == 256 * bands
    r_lo, r_hi = (1800, 2184) if bands == 1 else (res - 192, res + 192)
    want_img, used = _oracle_rows(scene, res, bands, r_lo, r_hi)
    assert used >= 32 and np.count_nonzero(want_img) > 10000
    w = scene['windows']
    if bands == 1:
        ks = [int(k) for k in np.linspace(0, K - 1, 16).astype(int)]
    else:
        ks = [k for k in range(K) if w[k][0] < res < w[k][1]][:16]          # windows crossing the band boundary
        ks += [k for k in (0, 100, 300, K - 1) if k not in ks]
    assert len(ks) >= 16
    rng = np.random.default_rng(11)
    dL = rng.uniform(-1, 1, size=(res * bands, res)).astype(np.float32)
    want_g, want_gabs = _oracle_grad(scene, res, bands, dL, ks)
    got_rows = np.zeros_like(want_img)
    got_g = np.zeros((K, 7))
    for rank in range(bands):
        r = MultiTriangleRasterizer(load_program('tri_color_primal'), load_program('tri_color_grad'), res,
                                    bands=bands, rank=rank)
        br = scene['band_ranges']
        active = (br[max(rank - 1, 0)][0], br[min(rank + 1, bands - 1)][1]) if bands > 1 else None
        r.set_scene(scene['verts'], scene['col'], scene['windows'], scene['classes'], scene['max_rows'],
                    scene['max_cols'], active=active)
        img = r.forward()
        band = torch.as_tensor(dL[res * rank: res * (rank + 1)]).cuda().contiguous()
        g = r.backward(band)
        torch.cuda.synchronize()
        a0, a1 = max(r_lo, res * rank), min(r_hi, res * (rank + 1))
        got_rows[a0 - r_lo:a1 - r_lo] = img[0, a0 - res * rank:a1 - res * rank].cpu().numpy()
        got_g += g.double().cpu().numpy()
        if rank == 0:
            # the same step as ONE captured CUDA graph (how bench.py replays it): identical bits at this size
            img2 = torch.empty_like(img)
            dl2 = torch.empty((res, res), dtype=torch.float32, device='cuda')
            g2 = torch.empty_like(g)
            target = torch.full((res, res), 0.25, device='cuda')

            def body(sp):
                r.forward(out=img2, stream=sp)
                torch.sub(img2[0], target, out=dl2)
                r.backward(dl2, out=g2, stream=sp)

            graph = r.capture(body)
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(img2, img), 'captured forward differs from the launch-by-launch forward'
            g3 = r.backward((img[0] - target).contiguous())
            torch.cuda.synchronize()
            assert torch.equal(g2, g3), 'captured backward differs from the launch-by-launch backward'
            del graph
    assert np.array_equal(got_rows, want_img), f'{np.count_nonzero(got_rows != want_img)} pixels of the strip differ'
    for k in ks:
        assert np.all(np.abs(got_g[k] - want_g[k]) <= 2e-6 * want_gabs[k] + 1e-12), (k, got_g[k], want_g[k])
    assert max(np.abs(want_g[k]).max() for k in ks) > 1e-3


def test_upstream_gradient_as_a_difference_of_two_images(cuda_device):
    """``backward((image, target))`` forms dL = image - target inside the kernel (one IEEE subtraction on load): the same
    bits as ``backward(image - target)`` with the difference written out by an elementwise kernel first."""
    import torch
    from teg_b200.raster import MultiTriangleRasterizer
    res = 192
    scene = multi_triangle_scene(bands=1, per_band_grid=4, res=res)
    r = MultiTriangleRasterizer(load_program('tri_color_primal'), load_program('tri_color_grad'), res)
    r.set_scene(scene['verts'], scene['col'], scene['windows'], scene['classes'], scene['max_rows'], scene['max_cols'])
    img = r.forward()
    target = torch.rand((res, res), device='cuda')
    dL = (img[0] - target).contiguous()
    a = r.backward(dL).clone()
    b = r.backward((img[0], target)).clone()
    torch.cuda.synchronize()
    assert torch.equal(a, b) and float(a.abs().sum()) > 0


@pytest.mark.parametrize('rows,cull', [(None, True), ((40, 171), True), ((40, 171), False)])
def test_gather_launch_equals_the_launches_per_class(cuda_device, rows, cull):
    """tegb_render_groups_gather (one launch over the image, every block walking the triangles of its bin in ascending
    order) against one accumulating launch per class of window-disjoint triangles: the same image, bit for bit -- for
    windows that start anywhere (not on tile boundaries) and a band of rows that starts and ends anywhere."""
    import torch
    from teg_b200.raster import MultiTriangleRasterizer
    res = 192
    scene = multi_triangle_scene(bands=1, per_band_grid=5, res=res)
    w = scene['windows']
    assert np.any(w[:, 0] % 16 != 0) and np.any(w[:, 2] % 32 != 0), 'the scene should have unaligned windows'
    imgs = {}
    for gather in (True, False):
        r = MultiTriangleRasterizer(load_program('tri_color_primal'), load_program('tri_color_grad'), res, rows=rows,
                                    gather=gather, cull=cull)           # (cull=False: no tile table, the dense program)
        r.set_scene(scene['verts'], scene['col'], scene['windows'], scene['classes'], scene['max_rows'],
                    scene['max_cols'])
        assert (r.bins is not None) == gather
        out = torch.full((1, r.n_rows, res), float('nan'), dtype=torch.float32, device='cuda')     # (forward clears it)
        imgs[gather] = r.forward(out=out)[0].cpu().numpy()
    assert np.count_nonzero(imgs[False]) > 1000
    assert np.array_equal(imgs[True].view(np.uint32), imgs[False].view(np.uint32)), \
        f'{np.count_nonzero(imgs[True] != imgs[False])} pixels differ'
    if not cull:
        # (and the dense program equals the culled one: the culled image of the same scene, bit for bit)
        r = MultiTriangleRasterizer(load_program('tri_color_primal'), load_program('tri_color_grad'), res, rows=rows)
        r.set_scene(scene['verts'], scene['col'], scene['windows'], scene['classes'], scene['max_rows'],
                    scene['max_cols'])
        assert np.array_equal(r.forward()[0].cpu().numpy().view(np.uint32), imgs[True].view(np.uint32))


def test_gather_launch_with_heavily_overlapping_windows_against_the_oracle(cuda_device):
    """Twelve triangles whose windows all overlap (no class structure at all; some windows cut their triangle): the
    gather launch adds every pixel's contributions in ascending triangle order, like the oracle's loop."""
    import torch
    from oracle import Oracle
    import teg_b200
    from teg_b200.raster import BOX, PARAMS, group_bins
    from teg_b200.workloads import base_name
    res = 160
    rng = np.random.default_rng(11)
    K = 12
    verts = rng.uniform(0.15, 0.85, size=(K, 6))
    col = rng.uniform(0.2, 1.0, size=K)
    win = np.zeros((K, 4), dtype=np.int32)
    for k in range(K):
        win[k] = (max(int(verts[k, 0::2].min() * res) - 2, 0), min(int(verts[k, 0::2].max() * res) + 3, res),
                  max(int(verts[k, 1::2].min() * res) - 2, 0), min(int(verts[k, 1::2].max() * res) + 3, res))
    win[3, 1] = (win[3, 0] + win[3, 1]) // 2            # a window that cuts its triangle
    win[7, 2] = (win[7, 2] + win[7, 3]) // 2
    prog = load_program('tri_color_primal')
    mode = teg_b200.CUDA_EvalMode(prog, num_samples=16, tile_rows=16)
    lab = {base_name(l): l for l in prog.args}
    dev = torch.device('cuda', 0)
    gb = {lab[nm]: torch.as_tensor(np.ascontiguousarray(verts[:, i]), dtype=torch.float32, device=dev)
          for i, nm in enumerate(PARAMS[:6])}
    gb[lab['col']] = torch.as_tensor(col, dtype=torch.float32, device=dev)
    bb, bg = group_bins(win, (0, res), res, mode.block_size, 16)
    out = torch.zeros((1, res, res), dtype=torch.float32, device=dev)
    mode.render_groups([lab[b] for b in BOX], (res, res), ((0, 1), (0, 1)), gb,
                       torch.as_tensor(win, device=dev), int((win[:, 1] - win[:, 0]).max()),
                       int((win[:, 3] - win[:, 2]).max()), out=out, accumulate=True,
                       gather=(torch.as_tensor(bb, device=dev), torch.as_tensor(bg, device=dev)))
    got = out[0].cpu().numpy()
    orc = Oracle(program_text('tri_color_primal'))
    want = np.zeros((res, res), dtype=np.float32)
    xs = np.linspace(0.0, 1.0, res + 1)
    v32 = verts.astype(np.float32).astype(np.float64)
    c32 = col.astype(np.float32).astype(np.float64)
    for k in range(K):
        r0, r1, c0, c1 = win[k]
        nx = np.repeat(np.arange(r0, r1), c1 - c0)
        ny = np.tile(np.arange(c0, c1), r1 - r0)
        v = orc.eval([xs[nx], xs[nx + 1], xs[ny], xs[ny + 1]] + list(v32[k]) + [c32[k]], 16, np.float32, threads=8)[0]
        want[nx, ny] = want[nx, ny] + v
    assert np.count_nonzero(want) > 2000
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f'{np.count_nonzero(got != want)} pixels differ'
