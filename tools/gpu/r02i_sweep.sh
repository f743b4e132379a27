#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
: > $O/r02i_sweep.txt
for v in "--tile-rows 16 --coop-groups 4" "--tile-rows 16 --coop-groups 1" "--tile-rows 8 --coop-groups 4" "--tile-rows 32 --coop-groups 4" "--tile-rows 16 --coop-groups 2"; do
  python tools/bench_multi_tri.py --steps 10 --no-cpu --graph 1 --min-blocks 4,4 $v > $O/tmp.json 2>$O/tmp.err
  python - "$v" <<'PY' >> $O/r02i_sweep.txt
import json,sys
try:
    d=json.load(open('gpurun_out/tmp.json')); print('multi_tri', sys.argv[1], '| step', round(d['ms_per_step'],4), 'phase', {k:round(v,4) for k,v in d['phase_ms'].items()}, 'G pairs/s', round(d['value']/1e9,2), d['check'])
except Exception as e:
    print('multi_tri', sys.argv[1], 'FAILED', e, open('gpurun_out/tmp.err').read()[-600:])
PY
done
for v in "--coop-groups 4" "--coop-groups 1" "--coop-groups 2"; do
  python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extras $v > $O/tmp.json 2>$O/tmp.err
  python - "$v" <<'PY' >> $O/r02i_sweep.txt
import json,sys
try:
    d=json.load(open('gpurun_out/tmp.json')); print('tri', sys.argv[1], '| step', round(d['ms_per_step'],4), 'phase', {k:round(v,4) for k,v in d['phase_ms'].items()}, 'value-kernel', round(d['roofline']['kernel_ms'],4))
except Exception as e:
    print('tri', sys.argv[1], 'FAILED', e, open('gpurun_out/tmp.err').read()[-400:])
PY
done
timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_multi_triangle.py -m gpu -x -q > $O/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02i_pytest.log
cat $O/r02i_sweep.txt; tail -15 $O/r02i_pytest.log
