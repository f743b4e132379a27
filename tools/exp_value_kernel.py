#!/usr/bin/env python
"""Experiment: what bounds the tiled value kernel?  Times tiles + main for triangles with no / all / some unresolved tiles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from teg_b200 import CUDA_EvalMode
from teg_b200.tegp import Program

prog = Program.load(os.path.join(ROOT, 'tests', 'golden', 'programs', 'tri_primal.tegp'))
opts = {}
for a in sys.argv[1:]:                       # e.g. --min-blocks=6 --tile-rows=8
    k, _, v = a.lstrip('-').partition('=')
    opts[k.replace('-', '_')] = int(v)
print('workers/SM', os.environ.get('TEGB200_QUEUE_WORKERS', 'default'), opts)
mode = CUDA_EvalMode(prog, num_samples=16, **opts)
res = 4096
box = prog.args[:4]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
scenes = {'default': [0.15, 0.2, 0.85, 0.3, 0.4, 0.9], 'outside': [2.15, 2.2, 2.85, 2.3, 2.4, 2.9],
          'covering': [-5.0, -5.0, 9.0, -5.0, -5.0, 9.0], 'small': [0.50, 0.50, 0.53, 0.5, 0.51, 0.53],
          'sliver': [0.0, 0.5, 1.0, 0.5004, 0.0, 0.5008], 'axis_rect_like': [0.1, 0.1, 0.9, 0.1, 0.1, 0.9]}
stream = torch.cuda.current_stream()
for name, v in scenes.items():
    pr = mode.prepare_render(box, (res, res), ((0, 1), (0, 1)), dict(zip(prog.args[4:], v)))
    pr.prog.enable_timing(True)
    for _ in range(3):
        pr.run(stream.cuda_stream)
    torch.cuda.synchronize()
    ts, ks = [], []
    for k in range(20):
        flush.fill_(k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pr.run(stream.cuda_stream); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
        ks.append(pr.prog.main_kernel_ms())
    n_t = (res // mode.codegen_opts.get('tile_rows', 16)) * (res // 32)
    t = pr.prog.tile_table(n_t)
    print(f'{name:16s} tiles+main {np.mean(ts)*1e3:7.2f} us (min {np.min(ts)*1e3:6.2f})  main {np.mean(ks)*1e3:6.2f} us  unresolved tiles {int((t == 2).sum()):6d}  inside {int((t == 1).sum()):6d}  area {float(pr.out.double().sum()):.6f}')
out = torch.empty((res, res), device='cuda')
ts = []
for k in range(20):
    flush.fill_(k)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out.fill_(0.0); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f'torch fill_ of the image: {np.mean(ts)*1e3:.2f} us')
