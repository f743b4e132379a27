"""Experiment: primal kernel time vs register cap / outer unroll."""
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
for opts, kw in (('', {}), ('--maxrregcount=64', {}), ('--maxrregcount=56', {}), ('--maxrregcount=48', {}), ('--maxrregcount=80', {}),
                 ('', {'unroll_outer': 2}), ('', {'unroll_outer': 4}), ('--maxrregcount=64', {'unroll_outer': 2}),
                 ('', {'block_size': 64}), ('', {'block_size': 96}), ('--maxrregcount=64', {'block_size': 64})):
    m = CUDA_EvalMode(prog, num_samples=16, use_cache=False, nvrtc_opts=opts, **kw)
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
    print(f'opts={opts!r:22} {kw!s:22} {best:.4f} ms  identical={torch.equal(out, ref)}', flush=True)
