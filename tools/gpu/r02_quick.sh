#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > $O/r02_quick_bench.json 2>$O/r02_quick_bench.err; python -c "
import json; d=json.load(open('$O/r02_quick_bench.json')); print(d['value'], d['ms_per_step'], d['check']['ok'], d['phase_ms'], d['gpu_launches'], d['roofline']['kernel_ms'])"
timeout 900 python -m pytest tests/test_gpu_cull.py tests/test_gpu_threads.py tests/test_gpu_edge_cases.py tests/test_gpu_fit.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > $O/r02_quick_bench2.json 2>$O/r02_quick_bench.err; python -c "
import json; d=json.load(open('$O/r02_quick_bench2.json')); print(d['value'], d['ms_per_step'], d['check']['ok'])"
