#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_triangle.py tests/test_gpu_peer.py tests/test_gpu_fit.py -m gpu -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > $O/r02g_bench_n2.json 2>$O/r02g_bench_n2.err; echo "n2 rc=$?"; tail -4 $O/r02g_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench_n2.json'))
print(d['value'], d['ms_per_step'], d['check']['ok'])
for k,v in d['workloads'].items():
    if isinstance(v, dict): print(k, v.get('ms_per_step'), v.get('value'), v.get('phase_ms'), v.get('check',{}).get('ok'), v.get('error'))
PY
