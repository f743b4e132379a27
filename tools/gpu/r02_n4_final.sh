#!/bin/bash
# 4 GPUs: the driver's line (weak scaling, all workloads) -- one run, kept short
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu > $O/r02_n4_bench.json 2>$O/r02_n4_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$O/r02_n4_bench.json')); print(d['value'], d['ms_per_step'], d['check']['ok'], d['e2e']['value']); w=d.get('workloads',{}); print({k:(v.get('ms_per_step'), v.get('value'), v.get('check',{}).get('ok') if isinstance(v.get('check'),dict) else None) for k,v in w.items() if isinstance(v,dict)})"
