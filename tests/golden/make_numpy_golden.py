"""Golden vectors of the reference's NUMPY backend (trapezoid rule on linspace, teg/eval/numpy_eval.py:57-68).

    python tests/golden/make_numpy_golden.py        (needs /root/reference; a few seconds)

The numpy backend is a recursive Python interpreter in float64; it defines the semantics of the
``trapezoid`` quadrature mode.  Only programs whose comparisons are strict are used:
the numpy backend honours ``<=`` while the C backend (and TEGP) treat it as ``<``.
Output: tests/golden/vectors_numpy/<scene>_N<samples>.npz with inputs [n_args, n] and outputs [k, n].
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('TEG_REFERENCE_PATH', '/root/reference')
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
sys.setrecursionlimit(100000)

from teg.eval import evaluate  # noqa: E402
from teg.lang.base import Var  # noqa: E402

from oracle.scenes import SCENES  # noqa: E402
from teg_b200.tegp import Program, from_teg, label_of  # noqa: E402
from teg_b200.workloads import bind, scene_values  # noqa: E402

CASES = [('readme_primal', 50, 40), ('readme_primal', 7, 40), ('tri_primal', 8, 49), ('shaded_primal', 6, 36),
         ('nested2d_value', 9, 30), ('nested3d_value', 4, 12), ('tri_grad', 8, 49), ('nested2d_grad', 9, 30),
         ('bilinear_lerp', 10, 36), ('circle', 8, 25)]


def main():
    out_dir = os.path.join(HERE, 'vectors_numpy')
    os.makedirs(out_dir, exist_ok=True)
    for scene, ns, n in CASES:
        Var.global_uid = 1000
        program, arglist = SCENES[scene]()
        prog = from_teg(program, arglist, name=scene)
        committed = Program.load(os.path.join(HERE, 'programs', f'{scene}.tegp'))
        assert prog.dumps() == committed.dumps(), f'{scene}: frontend output differs from the committed fixture'
        side = int(round(np.sqrt(n)))
        vals = bind(prog.args, scene_values(scene, res=(side, side), n=n))
        cnt = max(np.size(v) for v in vals)
        soa = np.empty((len(vals), cnt))
        for i, v in enumerate(vals):
            soa[i] = v
        live = {lab: prog.live[lab] for lab in prog.args}
        outs = []
        for j in range(cnt):
            b = {live[lab]: float(soa[i, j]) for i, lab in enumerate(prog.args)}
            r = evaluate(program, b, backend='numpy', num_samples=ns)
            outs.append(np.atleast_1d(np.asarray(r, dtype=np.float64)))
        out = np.stack(outs, axis=1)
        path = os.path.join(out_dir, f'{scene}_N{ns}.npz')
        np.savez_compressed(path, inputs=soa, outputs=out, num_samples=ns)
        print(path, soa.shape, out.shape, 'nnz', np.count_nonzero(out))


if __name__ == '__main__':
    main()
