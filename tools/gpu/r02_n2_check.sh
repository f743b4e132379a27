#!/bin/bash
# 2 GPUs: the peer-exchange tests and both scaling modes of the bench (checks inside the line)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > $O/r02_n2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_n2_pytest.log; tail -5 $O/r02_n2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > $O/r02_n2_bench.json 2>$O/r02_n2_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$O/r02_n2_bench.json')); print(d['value'], d['ms_per_step'], d['check']); w=d.get('workloads',{}); print({k:(v.get('ms_per_step'), v.get('check',{}).get('ok') if isinstance(v.get('check'),dict) else None) for k,v in w.items() if isinstance(v,dict)})"
