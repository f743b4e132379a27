#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
N=8
P=29530
for v in "" "--graph 0" "--allreduce nccl" "--shard band" "--scaling strong"; do
P=$((P+1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --no-e2e --no-extras $v > $O/tmp_n8.json 2>$O/tmp_n8.err
python - "$v" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/tmp_n8.json')); print('N=8', sys.argv[1], '| step', round(d['ms_per_step'],4), 'phase', {k:round(v,4) for k,v in d['phase_ms'].items()}, 'value-kernel', round(d['roofline']['kernel_ms'],4), d['check']['ok'])
except Exception as e:
    print('N=8', sys.argv[1], 'FAILED', e, open('gpurun_out/tmp_n8.err').read()[-600:])
PY
done
nproc; cat /proc/cpuinfo | grep "model name" | head -1; cat /sys/fs/cgroup/cpu.max 2>/dev/null
