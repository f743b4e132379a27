"""Gradient all-reduce over NVLink peer memory, fused into the column-sum kernel (teg_b200/parallel.PeerGroup,
csrc/tegb200.cu colsum_stage2_peer).  World size 1 runs on any GPU box; the 2-rank test needs two GPUs and is
skipped otherwise (the CPU suite covers the N > 1 host logic with gloo, tests/test_distributed_gloo.py)."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT, load_program

from teg_b200.workloads import TRIANGLE

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl' if world > 1 else 'gloo', rank=rank, world_size=world,
                            **({'device_id': torch.device('cuda', rank)} if world > 1 else {}))
    from teg_b200 import CUDA_EvalMode
    from teg_b200.parallel import PeerGroup, shard_rows
    prog = load_program('tri_grad')
    res = 512
    rows = shard_rows(res, world, rank)
    params = {lab: TRIANGLE[lab.rsplit('_', 1)[0]] for lab in prog.args[4:]}
    mode = CUDA_EvalMode(prog, num_samples=16, device=rank)
    local = mode.render(prog.args[:4], (res, res), ((0, 1), (0, 1)), params, rows=rows, reduce=True).clone()
    ref = local.clone()
    if world > 1:
        dist.all_reduce(ref)                                     # the collective-library answer
    peers = PeerGroup(rank, max_values=16)
    pr = mode.prepare_render(prog.args[:4], (res, res), ((0, 1), (0, 1)), params, rows=rows, reduce=True)
    peers.attach(pr.prog)
    stream = torch.cuda.current_stream().cuda_stream
    outs = []
    for _ in range(50):                                          # many calls back to back: the two banks alternate
        outs.append(pr.run(stream).clone())
    torch.cuda.synchronize()
    peers.check()
    assert all(torch.equal(o, outs[0]) for o in outs)
    # stand-alone all-reduce of a vector, fp32 and fp64, ranks deliberately out of step
    v32 = torch.arange(1, 9, dtype=torch.float32, device=f'cuda:{rank}') * (rank + 1)
    v64 = torch.arange(1, 9, dtype=torch.float64, device=f'cuda:{rank}') / (rank + 3)
    w32, w64 = v32.clone(), v64.clone()
    if rank == 0:
        torch.cuda._sleep(50_000_000)
    peers.allreduce_(v32, stream)
    peers.allreduce_(v64, stream)
    if world > 1:
        dist.all_reduce(w32)
        dist.all_reduce(w64)
    torch.cuda.synchronize()
    peers.check()
    # an exchange inside a captured CUDA graph: every replay is a NEW exchange (the sequence number lives on the device),
    # so replays with changing inputs return the current sums, not a stale mailbox
    gbuf = torch.zeros(8, dtype=torch.float32, device=f'cuda:{rank}')
    gsrc = torch.zeros(8, dtype=torch.float32, device=f'cuda:{rank}')
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        gbuf.copy_(gsrc)
        peers.allreduce_(gbuf, side.cuda_stream)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        gbuf.copy_(gsrc)
        peers.allreduce_(gbuf, torch.cuda.current_stream().cuda_stream)
    graph_ok = True
    for it in range(6):
        gsrc.copy_(torch.arange(8, dtype=torch.float32, device=f'cuda:{rank}') * (it + 1) + rank)
        want_g = gsrc.clone()
        if world > 1:
            dist.all_reduce(want_g)
        if rank == it % world:
            torch.cuda._sleep(20_000_000)                        # ranks out of step
        graph.replay()
        torch.cuda.synchronize()
        graph_ok = graph_ok and bool(torch.equal(gbuf, want_g))
    peers.check()
    peers.detach(pr.prog)
    again = pr.run(stream).clone()
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f'peer_{rank}.npy'), outs[0].cpu().numpy())
    np.save(os.path.join(out_dir, f'ref_{rank}.npy'), ref.cpu().numpy())
    ok = (graph_ok and torch.equal(again, local) and torch.equal(v32, w32) and
          torch.allclose(v64, w64, rtol=1e-15, atol=0))
    np.save(os.path.join(out_dir, f'ok_{rank}.npy'), np.array([int(ok)]))
    peers.close()
    dist.destroy_process_group()


def _run(world, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    peer = [np.load(tmp_path / f'peer_{r}.npy') for r in range(world)]
    ref = [np.load(tmp_path / f'ref_{r}.npy') for r in range(world)]
    assert all(int(np.load(tmp_path / f'ok_{r}.npy')[0]) == 1 for r in range(world))
    for r in range(world):
        assert np.array_equal(peer[r], peer[0])                  # the same bits on every rank
        # rank-order sum vs the library's: identical for two ranks, rounding-level otherwise
        assert np.allclose(peer[r], ref[r], rtol=1e-6, atol=1e-9)
    return peer[0]


def test_peer_group_world_1(cuda_device, tmp_path):
    g = _run(1, tmp_path)
    assert abs(g[0] + 0.3) < 1e-2


def test_peer_allreduce_two_ranks(cuda_device, tmp_path):
    from teg_b200 import runtime
    if runtime.device_count() < 2:
        pytest.skip('needs two GPUs (run with gpurun --gpus 2)')
    g = _run(2, tmp_path)
    x0, y0, x1, y1, x2, y2 = (TRIANGLE[k] for k in ('px0', 'py0', 'px1', 'py1', 'px2', 'py2'))
    want = np.array([(y1 - y2) / 2, (x2 - x1) / 2, (y2 - y0) / 2, (x0 - x2) / 2, (y0 - y1) / 2, (x1 - x0) / 2])
    assert np.allclose(g, want, atol=6e-3)
