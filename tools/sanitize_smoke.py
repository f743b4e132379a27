#!/usr/bin/env python
"""Small run of every kernel variant, meant to be wrapped in compute-sanitizer:

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
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
    res = 64
    scene = multi_triangle_scene(bands=1, per_band_grid=4, res=res)
    r = MultiTriangleRasterizer(load('tri_color_primal'), load('tri_color_grad'), res)
    r.set_scene(scene['verts'], scene['col'], scene['windows'], scene['classes'], scene['max_rows'], scene['max_cols'])
    img = r.forward()
    g = r.backward(img[0].contiguous())
    torch.cuda.synchronize()
    print('sanitize_smoke ok', float(img.sum()), float(g.abs().sum()))


if __name__ == '__main__':
    main()
