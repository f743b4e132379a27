"""world_size = 2 on CPU (gloo): instance sharding + the single gradient all-reduce of the N > 1 path.

Each rank evaluates ITS rows of the pixel grid (with the oracle -- no GPU here), reduces its gradient
locally and all-reduces the [6] vector; the result must equal the single-process gradient over the whole
grid, and the value image bands must tile the full image exactly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, program_text

from teg_b200.parallel import allreduce_gradients, shard_rows


def test_shard_rows_partitions_exactly():
    for n in (0, 1, 7, 4096, 4097):
        for w in (1, 2, 3, 8):
            parts = [shard_rows(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, res, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import Oracle
    from teg_b200.tegp import Program
    from teg_b200.workloads import bind, scene_values
    rows = shard_rows(res, world, rank)
    outs = {}
    for scene in ('tri_primal', 'tri_grad'):
        prog = Program.loads(program_text(scene))
        inputs = bind(prog.args, scene_values(scene, res=(res, res), rows=rows))
        outs[scene] = Oracle(program_text(scene)).eval(inputs, 16, np.float64)
    g = torch.from_numpy(outs['tri_grad'].sum(axis=1))
    allreduce_gradients(g)
    np.save(os.path.join(out_dir, f'img_{rank}.npy'), outs['tri_primal'])
    np.save(os.path.join(out_dir, f'grad_{rank}.npy'), g.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_process(tmp_path):
    from oracle import Oracle
    from teg_b200.tegp import Program
    from teg_b200.workloads import bind, scene_values
    res, world = 48, 2
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, res, str(tmp_path)), nprocs=world, join=True)
    full = {}
    for scene in ('tri_primal', 'tri_grad'):
        prog = Program.loads(program_text(scene))
        inputs = bind(prog.args, scene_values(scene, res=(res, res)))
        full[scene] = Oracle(program_text(scene)).eval(inputs, 16, np.float64)
    img = np.concatenate([np.load(tmp_path / f'img_{r}.npy') for r in range(world)], axis=1)
    assert np.array_equal(img, full['tri_primal'])
    g0, g1 = np.load(tmp_path / 'grad_0.npy'), np.load(tmp_path / 'grad_1.npy')
    assert np.array_equal(g0, g1)
    assert np.allclose(g0, full['tri_grad'].sum(axis=1), rtol=1e-12, atol=1e-15)
    # analytic gradient of the triangle area wrt its vertices
    x0, y0, x1, y1, x2, y2 = 0.15, 0.2, 0.85, 0.3, 0.4, 0.9
    want = np.array([(y1 - y2) / 2, (x2 - x1) / 2, (y2 - y0) / 2, (x0 - x2) / 2, (y0 - y1) / 2, (x1 - x0) / 2])
    assert np.allclose(g0, want, atol=1e-2)      # quadrature error at 48x48, N=16
