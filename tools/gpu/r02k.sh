#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_cull.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_threads.py tests/test_gpu_peer.py -m gpu -x -q > $O/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02k_pytest.log
tail -6 $O/r02k_pytest.log
python tools/exp_value_kernel.py
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extras > $O/r02k_bench.json 2>$O/r02k_bench.err
python -c "
import json; d=json.load(open('$O/r02k_bench.json')); print('tri', d['value'], d['ms_per_step'], d['phase_ms'], d['roofline']['kernel_ms'], d['check']['ok'])"
