"""Experiment: primal kernel time vs fraction of samples using the min3 form (pipe balancing)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from teg_b200 import CUDA_EvalMode
from teg_b200.tegp import Program
from teg_b200.workloads import TRIANGLE
prog = Program.load(os.path.join(ROOT, 'tests/golden/programs/tri_primal.tegp'))
res = 4096
params = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in prog.args[4:]}
out = torch.empty((1, res, res), device='cuda')
ref = None
for ratio in (0.0, 0.125, 0.25, 0.3125, 0.375, 0.5, 0.625, 1.0):
    for bs in (128, 256):
        m = CUDA_EvalMode(prog, num_samples=16, pred_mix=ratio, block_size=bs, use_cache=False)
        for _ in range(3):
            m.render(prog.args[:4], (res, res), ((0, 1), (0, 1)), params, out=out)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); m.render(prog.args[:4], (res, res), ((0, 1), (0, 1)), params, out=out); b.record()
            torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
        if ref is None:
            ref = out.clone()
        print(f'pred_mix={ratio:<6} block={bs} {best:.4f} ms  identical={torch.equal(out, ref)}', flush=True)
