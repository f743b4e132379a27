#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_triangle.py tests/test_gpu_fit.py tests/test_gpu_peer.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "libm" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --workload multi_tri_4096 --no-extras > $O/r02x_mt_n1.json 2>$O/r02x_mt_n1.err; echo "mt rc=$?"; tail -3 $O/r02x_mt_n1.err
python -c "
import json; d=json.load(open('$O/r02x_mt_n1.json')); print('mt', d['value'], d['ms_per_step'], d['phase_ms'], d['check'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'])"
