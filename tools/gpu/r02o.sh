#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_cull.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_threads.py tests/test_gpu_peer.py -m gpu -x -q > $O/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02o_pytest.log
tail -6 $O/r02o_pytest.log
for w in 4 2 3 6; do
echo "== workers per SM $w"
TEGB200_QUEUE_WORKERS=$w python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extras > $O/r02o_bench.json 2>$O/r02o_bench.err
python -c "
import json; d=json.load(open('$O/r02o_bench.json')); print('tri', d['value'], d['ms_per_step'], d['phase_ms'], d['roofline']['kernel_ms'], d['check']['ok'])"
done
