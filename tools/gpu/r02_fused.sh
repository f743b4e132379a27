#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for f in 1 0; do
python bench.py --steps 40 --warmup 5 --no-cpu --no-extras --no-e2e --fused $f > $O/tmp.json 2>$O/tmp.err; echo "rc=$?"; tail -3 $O/tmp.err
python -c "
import json; d=json.load(open('$O/tmp.json')); print('fused $f', d['value'], d['ms_per_step'], d['phase_ms'], d['roofline']['kernel_ms'], d['check']['ok'], d['gpu_launches'])"
done
