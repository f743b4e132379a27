#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r02p_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02p_pytest.log
tail -5 $O/r02p_pytest.log
python bench.py --steps 20 --warmup 3 > $O/r02p_bench_n1.json 2>$O/r02p_bench_n1.err; echo "bench rc=$?"; tail -3 $O/r02p_bench_n1.err
python -c "
import json; d=json.load(open('$O/r02p_bench_n1.json')); print('tri', d['value'], d['ms_per_step'], d['phase_ms'], {k:d['roofline'][k] for k in ('kernel_ms','frac','call_ms_classification_plus_main','frac_with_classification')}, d['check']['ok']); print(d['workloads']['multi_tri_4096']['value'], d['workloads']['multi_tri_4096']['ms_per_step']); print(d['e2e']['value'], d['e2e_grad_only']['value'])"
python -c "import __graft_entry__ as g; g.smoke()"
