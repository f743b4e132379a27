#!/bin/bash
# measurement only: the fp64 line and the config-5 headline line for profiles/
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 200 python bench.py --steps 20 --warmup 5 --dtype f64 --no-cpu --no-extras > $O/r02_f64.json 2>$O/r02_f64.err; echo "f64 rc=$?"
timeout 200 python bench.py --steps 20 --warmup 5 --workload multi_tri_4096 --no-cpu --no-extras > $O/r02_mt_headline.json 2>$O/r02_mt_headline.err; echo "mt rc=$?"
python -c "
import json
for f in ('r02_f64','r02_mt_headline'):
    d=json.load(open('$O/'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['check'].get('ok'), d.get('e2e',{}).get('value'))"
