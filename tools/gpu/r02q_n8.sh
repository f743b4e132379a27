#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > $O/r02q_bench_n$N.json 2>$O/r02q_bench_n$N.err; echo "n$N rc=$?"; tail -4 $O/r02q_bench_n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.load(open(f'gpurun_out/r02q_bench_n{N}.json'))
print('headline', d['value'], d['ms_per_step'], d['check']['ok'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'grad-only', d['e2e_grad_only']['value'])
for k,v in d['workloads'].items():
    if isinstance(v, dict): print(k, v.get('ms_per_step'), v.get('value'), v.get('phase_ms'), (v.get('check') or {}).get('ok'), v.get('error'), v.get('GBps_aggregate'), v.get('GBps_per_rank_min'))
PY
