#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_multi_triangle.py tests/test_gpu_fit.py -m gpu -x -q 2>&1 | tail -4
for g in 1 0; do
python bench.py --steps 40 --warmup 5 --no-cpu --no-extras --graph $g > $O/r02t_bench_g$g.json 2>$O/r02t_bench_g$g.err; echo "rc=$?"; tail -3 $O/r02t_bench_g$g.err
python -c "
import json; d=json.load(open('$O/r02t_bench_g$g.json')); print('graph $g', d['value'], d['ms_per_step'], d['phase_ms'], d['roofline']['kernel_ms'], d['check']['ok'], d['gpu_launches'], 'e2e', d['e2e']['ms_per_step'], d['e2e_grad_only']['ms_per_step'])"
done
