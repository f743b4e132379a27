#!/usr/bin/env python
"""Optimisation loop on the device: fit triangle vertices to a target image.

    python tools/fit_triangle.py --res 1024 --iters 200
    python -m torch.distributed.run --nproc-per-node N tools/fit_triangle.py        # rows sharded over N GPUs

Thin command-line front of ``teg_b200.optim.fit_triangle``: loss, upstream gradient, Adam state and the parameters stay
on the device, every iteration is one replay of a captured CUDA graph, nothing is read back inside the loop.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def fit(res=256, iters=150, lr=4e-3, device=0, start=None, target_verts=None, verbose=False, graph=True, peers=None):
    """(losses, fitted vertices, target vertices) -- the signature the earlier CPU-side loop had"""
    from teg_b200.optim import fit_triangle
    losses, v, vt, _ = fit_triangle(res, iters, lr, device, start, target_verts, graph=graph, peers=peers)
    return list(losses), v, vt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--res', type=int, default=1024)
    ap.add_argument('--iters', type=int, default=200)
    ap.add_argument('--lr', type=float, default=4e-3)
    ap.add_argument('--graph', type=int, default=1)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    peers = None
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        from teg_b200.parallel import PeerGroup
        peers = PeerGroup(local, max_values=16)
    t0 = time.perf_counter()
    losses, v, vt = fit(args.res, args.iters, args.lr, device=local, graph=bool(args.graph), peers=peers)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if int(os.environ.get('RANK', '0')) == 0:
        print(f'loss {losses[0]:.6f} -> {losses[-1]:.6f}; max vertex error {np.abs(v - vt).max():.4f}; '
              f'{args.iters} iterations at {args.res}^2 in {dt:.2f} s (set-up and JIT included)')
    if peers is not None:
        peers.close()
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
