#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_cull.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_threads.py tests/test_gpu_peer.py tests/test_ref_suite_replay.py -m gpu -x -q > $O/r02v_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02v_pytest.log
tail -6 $O/r02v_pytest.log
for w in 2 1 3; do
echo "== workers per SM $w (x2 for the gradient)"
TEGB200_QUEUE_WORKERS=$w python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e --no-extras > $O/r02v_bench.json 2>$O/r02v_bench.err
python -c "
import json; d=json.load(open('$O/r02v_bench.json')); print('tri', d['value'], d['ms_per_step'], d['phase_ms'], d['roofline']['kernel_ms'], d['check']['ok'])"
done
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extras --graph 0"
$B > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $O/r02v_launches.csv $B > $O/r02v_ncu_l.log 2>&1
