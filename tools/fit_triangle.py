#!/usr/bin/env python
"""Optimisation loop on the device: fit triangle vertices to a target image.

    loss(v) = 1/2 * sum_p (I_v(p) - target(p))^2,   I_v(p) = per-pixel coverage integral of the triangle

Every iteration is two launches of reduced Teg programs through the CUDA backend: the value image
(``tri_primal``) and the reverse derivative with the upstream per-pixel gradient dL = I - target
(``tri_color_grad``-style programs take dL as an argument; here the colour is fixed to 1), reduced to the
[6] parameter gradient inside the kernel.  Under torchrun the rows are sharded and the gradient is
all-reduced (one collective per iteration).  Adam runs on the 6 numbers with torch.

    python tools/fit_triangle.py --res 512 --iters 200
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def fit(res=256, iters=150, lr=4e-3, device=0, start=None, target_verts=None, verbose=False):
    import torch
    import torch.distributed as dist
    from teg_b200 import CUDA_EvalMode
    from teg_b200.parallel import allreduce_gradients, shard_rows
    from teg_b200.tegp import Program
    from teg_b200.workloads import base_name

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    pdir = os.path.join(ROOT, 'tests', 'golden', 'programs')
    primal = Program.load(os.path.join(pdir, 'tri_color_primal.tegp'))
    grad = Program.load(os.path.join(pdir, 'tri_color_grad.tegp'))
    m_val = CUDA_EvalMode(primal, num_samples=16, device=device)
    m_grd = CUDA_EvalMode(grad, num_samples=16, device=device)
    dev = torch.device('cuda', device)
    rows = shard_rows(res, world, rank)
    names = ['px0', 'py0', 'px1', 'py1', 'px2', 'py2']
    target_verts = np.array([0.15, 0.2, 0.85, 0.3, 0.4, 0.9]) if target_verts is None else np.asarray(target_verts)
    start = np.array([0.25, 0.25, 0.75, 0.35, 0.45, 0.8]) if start is None else np.asarray(start)
    bounds = ((0.0, 1.0), (0.0, 1.0))

    def labels(prog):
        return {base_name(lab): lab for lab in prog.args}

    lv, lg = labels(primal), labels(grad)
    box_v = [lv[b] for b in ('x_lb', 'x_ub', 'y_lb', 'y_ub')]
    box_g = [lg[b] for b in ('x_lb', 'x_ub', 'y_lb', 'y_ub')]

    def image(v):
        b = {lv[n]: float(x) for n, x in zip(names, v)}
        b[lv['col']] = 1.0
        return m_val.render(box_v, (res, res), bounds, b, rows=rows)[0]

    target = image(target_verts).clone()
    # the gradient program reads dL per pixel: grouped launch with a single group = the whole band
    windows = torch.tensor([[rows[0], rows[1], 0, res]], dtype=torch.int32, device=dev)
    v = torch.tensor(start, dtype=torch.float64)
    m = torch.zeros(6, dtype=torch.float64)
    s = torch.zeros(6, dtype=torch.float64)
    losses = []
    for it in range(iters):
        img = image(v.numpy())
        dL = (img - target).contiguous()
        gb = {lg[n]: torch.full((1,), float(x), dtype=torch.float32, device=dev) for n, x in zip(names, v.numpy())}
        gb[lg['col']] = torch.ones(1, dtype=torch.float32, device=dev)
        g = m_grd.render_groups(box_g, (res, res), bounds, gb, windows, rows[1] - rows[0], res,
                                pixel_bindings={lg['dL']: dL}, rows=rows, reduce=True)[0, :6].double()
        loss = 0.5 * (dL.double() ** 2).sum()
        if world > 1:
            allreduce_gradients(g)
            dist.all_reduce(loss)
        # pixel values are integrals over a pixel of area 1/res^2: rescale so that lr is resolution free
        g = g.cpu() * res * res
        losses.append(float(loss.item()) * res * res)
        m = 0.9 * m + 0.1 * g
        s = 0.999 * s + 0.001 * g * g
        mh, sh = m / (1 - 0.9 ** (it + 1)), s / (1 - 0.999 ** (it + 1))
        v = v - lr * mh / (sh.sqrt() + 1e-12)
        if verbose and rank == 0 and it % 10 == 0:
            print(f'iter {it:4d} loss {losses[-1]:.6f} |v - v*| {float((v - torch.tensor(target_verts)).abs().max()):.4f}')
    return losses, v.numpy(), target_verts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--res', type=int, default=512)
    ap.add_argument('--iters', type=int, default=200)
    ap.add_argument('--lr', type=float, default=4e-3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    losses, v, vt = fit(args.res, args.iters, args.lr, device=local, verbose=True)
    if int(os.environ.get('RANK', '0')) == 0:
        print(f'loss {losses[0]:.6f} -> {losses[-1]:.6f}; max vertex error {np.abs(v - vt).max():.4f}')
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
