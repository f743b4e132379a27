#!/bin/bash
# forward pass of config 5 as one gather launch: parity tests, then with / without
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_triangle.py tests/test_gpu_fit.py -m gpu -x -q > $O/r02y_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02y_pytest.log
: > $O/r02y_sweep.txt
for v in "--gather 1" "--gather 0" "--gather 1 --tile-rows 32,32"; do
  timeout 600 python tools/bench_multi_tri.py --steps 10 --no-cpu --graph 1 $v > $O/tmp.json 2>$O/tmp.err
  python - "$v" <<'PY' >> $O/r02y_sweep.txt
import json,sys
try:
    d=json.load(open('gpurun_out/tmp.json')); print('multi_tri', sys.argv[1], '| step', round(d['ms_per_step'],4), 'phase', {k:round(v,4) for k,v in d['phase_ms'].items()}, 'G pairs/s', round(d['value']/1e9,2), d['check'])
except Exception as e:
    print('multi_tri', sys.argv[1], 'FAILED', e, open('gpurun_out/tmp.err').read()[-600:])
PY
done
cat $O/r02y_sweep.txt; tail -25 $O/r02y_pytest.log
