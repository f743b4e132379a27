#!/usr/bin/env python
"""Small run of every kernel variant (odd grids, bands that start inside a tile, every launch path).  Where
compute-sanitizer may be used, wrap it: ``compute-sanitizer --tool memcheck python tools/sanitize_smoke.py``; on this
round's GPU pool the sanitizer is closed, so the parity suite (every variant against the oracle) is the bounds check."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from teg_b200 import CUDA_EvalMode  # noqa: E402
from teg_b200.raster import MultiTriangleRasterizer  # noqa: E402
from teg_b200.tegp import Program  # noqa: E402
from teg_b200.workloads import bind, multi_triangle_scene, scene_values  # noqa: E402

P = os.path.join(ROOT, 'tests', 'golden', 'programs')


def load(s):
    return Program.load(os.path.join(P, f'{s}.tegp'))


def main():
    n_launch = 0
    for scene, ns in (('tri_primal', 16), ('tri_grad', 16), ('shaded_grad', 16), ('nested3d_grad', 8), ('circle_dt', 20)):
        prog = load(scene)
        inputs = bind(prog.args, scene_values(scene, res=(37, 29), n=700))
        b = dict(zip(prog.args, inputs))
        for fw in ('float', 'double'):
            m = CUDA_EvalMode(prog, num_samples=ns, float_width=fw)
            m.eval(b)                                   # thread per instance (+ coop guards)
            m.eval(b, reduce=True)                      # in-kernel reduction
            for team in (4, 32, 256):
                CUDA_EvalMode(prog, num_samples=ns, float_width=fw, team=team).eval(b)
            n_launch += 5
        if scene.startswith('tri'):
            m = CUDA_EvalMode(prog, num_samples=ns)
            scal = {lab: v for lab, v in b.items() if np.ndim(v) == 0}
            m.render(prog.args[:4], (45, 70), ((0, 1), (0, 1)), scal)
            m.render(prog.args[:4], (45, 70), ((0, 1), (0, 1)), scal, reduce=True, rows=(3, 40))
            # tile classification on grids that are not multiples of the tile, interleaved tile rows, dense kernels
            m.render(prog.args[:4], (300, 200), ((0, 1), (0, 1)), scal, interleave=(3, 1))
            m.render(prog.args[:4], (300, 200), ((0, 1), (0, 1)), scal, reduce=True, interleave=(2, 1), rows=(5, 290))
            CUDA_EvalMode(prog, num_samples=ns, cull=False).render(prog.args[:4], (100, 130), ((0, 1), (0, 1)), scal)
            CUDA_EvalMode(prog, num_samples=ns, tiles=False).render(prog.args[:4], (100, 130), ((0, 1), (0, 1)), scal)
    # sparse gradient programs whose interior is NOT zero (listed tile blocks next to the queue workers)
    prog = load('shaded_grad')
    b = dict(zip(prog.args, bind(prog.args, scene_values('shaded_grad', res=(2, 2)))))
    scal = {lab: float(v) for lab, v in b.items() if np.ndim(v) == 0}
    box = [next(a for a in prog.args if a.rsplit('_', 1)[0] == nm) for nm in ('x_lb', 'x_ub', 'y_lb', 'y_ub')]
    CUDA_EvalMode(prog, num_samples=16).render(box, (150, 333), ((0, 1), (0, 1)), scal, reduce=True, rows=(7, 149))
    # grouped: gather forward (image blocks walk their bins; unaligned windows, a band that starts inside a tile), the
    # launches per class, the dense programs, the backward pass with dL and with dL as a difference of two images
    for res, rows, kw in ((64, None, {}), (192, (40, 171), {}), (192, (40, 171), {'gather': False}), (96, (5, 90), {'cull': False})):
        scene = multi_triangle_scene(bands=1, per_band_grid=4, res=res)
        r = MultiTriangleRasterizer(load('tri_color_primal'), load('tri_color_grad'), res, rows=rows, **kw)
        r.set_scene(scene['verts'], scene['col'], scene['windows'], scene['classes'], scene['max_rows'], scene['max_cols'])
        img = r.forward()
        g = r.backward(img[0].contiguous())
        g = g + r.backward((img[0].contiguous(), torch.zeros_like(img[0])))
    torch.cuda.synchronize()
    print('sanitize_smoke ok', float(img.sum()), float(g.abs().sum()))


if __name__ == '__main__':
    main()
