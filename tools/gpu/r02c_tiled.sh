#!/bin/bash
# round 2, call c: new plain tiled main kernel (packed order, ordered states, 16-byte fill, rows in flight)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_cull.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_threads.py -m gpu -x -q > $O/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02c_pytest.log
tail -5 $O/r02c_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e"
$B > $O/r02c_bench.json 2> $O/r02c_bench.err && \
ncu --set full --clock-control none --import-source on -k regex:'tri_(primal|grad)_main' -s 6 -c 2 -o $O/r02c_prof_tri $B > $O/r02c_ncu_tri.log 2>&1
python -c "
import json; d=json.load(open('$O/r02c_bench.json')); print(d['value'], d['ms_per_step'], d['phase_ms'], d['roofline']['kernel_ms'])"
