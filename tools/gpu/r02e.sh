#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02e_pytest.log
tail -4 $O/r02e_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > $O/r02e_bench.json 2>$O/r02e_bench.err
python -c "
import json; d=json.load(open('$O/r02e_bench.json')); print('tri', d['value'], d['ms_per_step'], d['phase_ms'], d['roofline']['kernel_ms'])"
python tools/bench_multi_tri.py --steps 10 --no-cpu --graph 1 --min-blocks 4,4 > $O/r02e_mt.json 2>$O/r02e_mt.err
python -c "
import json; d=json.load(open('$O/r02e_mt.json')); print('mt', d['value'], d['ms_per_step'], d['phase_ms'], d['check'])"
