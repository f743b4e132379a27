#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_triangle.py tests/test_gpu_fit.py -m gpu -x -q 2>&1 | tail -4
for v in "--tile-rows 16,32" "--tile-rows 16,16" "--tile-rows 32,32"; do
python tools/bench_multi_tri.py --steps 10 --no-cpu --graph 1 --min-blocks 4,4 $v > $O/tmp.json 2>$O/tmp.err
python - "$v" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/tmp.json')); print('multi_tri', sys.argv[1], '| step', round(d['ms_per_step'],4), 'phase', {k:round(v,4) for k,v in d['phase_ms'].items()}, 'G pairs/s', round(d['value']/1e9,2), d['check'])
except Exception as e:
    print('multi_tri', sys.argv[1], 'FAILED', e, open('gpurun_out/tmp.err').read()[-900:])
PY
done
