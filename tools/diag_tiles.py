import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch
from conftest import load_program
from teg_b200 import CUDA_EvalMode
from teg_b200.workloads import TRIANGLE
for scene in ('tri_primal', 'tri_grad'):
    prog = load_program(scene)
    mode = CUDA_EvalMode(prog, num_samples=16)
    params = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in prog.args[4:]}
    red = scene == 'tri_grad'
    mode.render(prog.args[:4], (4096, 4096), ((0, 1), (0, 1)), params, reduce=red)
    torch.cuda.synchronize()
    ks, p = mode.compiled([0, 1, 2, 3], [0, 1, 2, 3], reduce=red)
    n_tiles = (4096 // ks.tile_cols) * (4096 // ks.tile_rows)
    t = p.tile_table(n_tiles * ks.tile_conds).reshape(ks.tile_conds, -1)
    for k in range(ks.tile_conds):
        print(scene, k, {v: float(np.mean(t[k] == v)) for v in (0, 1, 2)})
    print(scene, 'any unresolved', float(np.mean((t == 2).any(axis=0))))
