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


def _interleave_worker(rank, world, port, res, tile_rows, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import Oracle
    from teg_b200.eval import interleaved_rows
    from teg_b200.tegp import Program
    from teg_b200.workloads import TRIANGLE, bind, pixel_boxes
    rows = interleaved_rows(res, tile_rows, world, rank)            # what render(interleave=(world, rank)) renders
    outs = {}
    for scene in ('tri_primal', 'tri_grad'):
        prog = Program.loads(program_text(scene))
        parts = [pixel_boxes((res, res), rows=(r, r + 1)) for r in rows]
        box = [np.concatenate([p[k] for p in parts]) for k in range(4)]
        vals = dict(zip(('x_lb', 'x_ub', 'y_lb', 'y_ub'), box), **TRIANGLE)
        outs[scene] = Oracle(program_text(scene)).eval(bind(prog.args, vals), 16, np.float64)
    g = torch.from_numpy(outs['tri_grad'].sum(axis=1))
    allreduce_gradients(g)
    np.save(os.path.join(out_dir, f'rows_{rank}.npy'), np.asarray(rows))
    np.save(os.path.join(out_dir, f'img_{rank}.npy'), outs['tri_primal'][0].reshape(len(rows), res))
    np.save(os.path.join(out_dir, f'grad_{rank}.npy'), g.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_interleaved_sharding_tiles_the_image_and_sums_the_gradient(tmp_path):
    """bench.py's default N > 1 layout: rank r owns the tile rows r, r + world, ...  The shards of all ranks are
    disjoint, cover every row, reassemble to the single-process image, and their all-reduced gradient is the full one."""
    from oracle import Oracle
    from teg_b200.eval import interleaved_rows
    from teg_b200.tegp import Program
    from teg_b200.workloads import bind, scene_values
    res, world, tile_rows = 40, 2, 16                  # 3 tile rows (the last one short): rank 0 gets two, rank 1 one
    assert sorted(interleaved_rows(res, tile_rows, world, 0) + interleaved_rows(res, tile_rows, world, 1)) == list(range(res))
    assert interleaved_rows(res, tile_rows, world, 1) == list(range(16, 32))
    assert interleaved_rows(res, tile_rows, 8, 5) == []                        # more ranks than tile rows
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_interleave_worker, args=(world, port, res, tile_rows, str(tmp_path)), nprocs=world, join=True)
    full = {}
    for scene in ('tri_primal', 'tri_grad'):
        prog = Program.loads(program_text(scene))
        full[scene] = Oracle(program_text(scene)).eval(bind(prog.args, scene_values(scene, res=(res, res))), 16, np.float64)
    img = np.zeros((res, res))
    for r in range(world):
        img[np.load(tmp_path / f'rows_{r}.npy')] = np.load(tmp_path / f'img_{r}.npy')
    assert np.array_equal(img, full['tri_primal'][0].reshape(res, res))
    g0, g1 = np.load(tmp_path / 'grad_0.npy'), np.load(tmp_path / 'grad_1.npy')
    assert np.array_equal(g0, g1) and np.allclose(g0, full['tri_grad'].sum(axis=1), rtol=1e-12, atol=1e-15)


def _peer_fail_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from teg_b200.parallel import PeerGroup
    outcome = 'created'
    try:
        PeerGroup(0, max_values=8)
    except RuntimeError as e:
        outcome = f'raised: {e}'
    with open(os.path.join(out_dir, f'peer_{rank}.txt'), 'w') as f:
        f.write(outcome)
    dist.barrier()
    dist.destroy_process_group()


def test_peer_group_fails_on_every_rank_together_without_a_device(tmp_path):
    """No CUDA device here: the mailbox allocation fails locally, the ranks agree on it and ALL raise (nobody is
    left waiting inside a collective); bench.py then falls back to the NCCL all-reduce."""
    from teg_b200 import runtime
    if runtime.device_count() > 0:
        import pytest
        pytest.skip('a CUDA device is present')
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_peer_fail_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        text = (tmp_path / f'peer_{r}.txt').read_text()
        assert text.startswith('raised: peer mailboxes unavailable'), text
