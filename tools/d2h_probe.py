#!/usr/bin/env python
"""What bounds the end-to-end figure at N > 1: N ranks copy one 4096^2 fp32 image each from their GPU
to pinned host memory at the same time -- plain cudaMemcpyAsync, one copy per call.  Variants: the default pinned
allocation; the rank pinned to its own slice of the host cores BEFORE it allocates and touches the buffer; several
buffers in flight.  Prints one JSON line (rank 0).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/d2h_probe.py
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    nbytes = 4096 * 4096 * 4
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)

    def agg(x):
        t = torch.tensor([x, -x], dtype=torch.float64, device=dev)
        if world > 1:
            s = t.clone()
            dist.all_reduce(s)
            m = t.clone()
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
            return float(s[0]), -float(m[1])
        return x, x

    def run(dsts, reps=8):
        for d in dsts:
            d.copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for k in range(reps):
            dsts[k % len(dsts)].copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        return agg(nbytes * reps / dt / 1e9)

    out = {'world': world, 'bytes_per_copy': nbytes, 'host_cpus': os.cpu_count(),
           'affinity_before': sorted(os.sched_getaffinity(0))[:4] + ['...'] if len(os.sched_getaffinity(0)) > 4 else sorted(os.sched_getaffinity(0))}
    a = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    out['default'] = dict(zip(('GBps_aggregate', 'GBps_per_rank_min'), run([a])))
    cpus = sorted(os.sched_getaffinity(0))
    per = max(1, len(cpus) // world)
    mine = cpus[rank * per:(rank + 1) * per] or cpus
    os.sched_setaffinity(0, mine)
    b = torch.empty(nbytes, dtype=torch.uint8)
    b.fill_(1)                                   # first touch on this rank's cores
    b = b.pin_memory()
    out['rank_pinned_to_its_cores_then_allocated'] = dict(zip(('GBps_aggregate', 'GBps_per_rank_min'), run([b])))
    c = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(3)]
    out['three_buffers_in_flight'] = dict(zip(('GBps_aggregate', 'GBps_per_rank_min'), run(c, reps=9)))
    # the opposite direction for reference
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(8):
        src.copy_(h, non_blocking=True)
    torch.cuda.synchronize(dev)
    out['host_to_device'] = dict(zip(('GBps_aggregate', 'GBps_per_rank_min'), agg(nbytes * 8 / (time.perf_counter() - t0) / 1e9)))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
