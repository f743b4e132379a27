#!/bin/bash
# round 2, call d: geometry sweep of the plain tiled kernels + launch-bound sweep of the grouped kernels
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
: > $O/r02d_sweep.txt
for v in "--tile-rows 16" "--tile-rows 32" "--tile-rows 64" "--tile-rows 32 --block 256" "--tile-rows 16 --block 256" "--tile-rows 16 --min-blocks 6" "--tile-rows 32 --min-blocks 4" "--tile-rows 8"; do
  python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e $v > $O/tmp.json 2>$O/tmp.err
  python - "$v" <<'PY' >> $O/r02d_sweep.txt
import json,sys
try:
    d=json.load(open('gpurun_out/tmp.json')); print(sys.argv[1], '| step', round(d['ms_per_step'],4), 'phase', {k:round(v,4) for k,v in d['phase_ms'].items()}, 'value-kernel', round(d['roofline']['kernel_ms'],4))
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('gpurun_out/tmp.err').read()[-400:])
PY
done
for mb in "0,0" "4,3" "4,4" "6,4" "8,4" "8,6" "6,3"; do
  python tools/bench_multi_tri.py --steps 10 --no-cpu --graph 1 --min-blocks $mb > $O/tmp.json 2>$O/tmp.err
  python - "$mb" <<'PY' >> $O/r02d_sweep.txt
import json,sys
try:
    d=json.load(open('gpurun_out/tmp.json')); print('multi_tri min_blocks', sys.argv[1], '| step', round(d['ms_per_step'],4), 'phase', {k:round(v,4) for k,v in d['phase_ms'].items()}, 'G pairs/s', round(d['value']/1e9,2), d['check'])
except Exception as e:
    print('multi_tri', sys.argv[1], 'FAILED', e, open('gpurun_out/tmp.err').read()[-600:])
PY
done
timeout 600 python -m pytest tests/test_gpu_multi_triangle.py tests/test_gpu_fit.py -m gpu -x -q > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02d_pytest.log
cat $O/r02d_sweep.txt; tail -5 $O/r02d_pytest.log
