#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  python bench.py --steps 20 --warmup 3 > $O/r02f_bench_n1.json 2>$O/r02f_bench_n1.err; echo "n1 rc=$?"; tail -5 $O/r02f_bench_n1.err
  python bench.py --impl reference --steps 3 --warmup 1 > $O/r02f_ref_n1.json 2>$O/r02f_ref_n1.err; echo "ref rc=$?"
  python bench.py --workload multi_tri_4096 --steps 10 --warmup 3 --no-extras > $O/r02f_mt_n1.json 2>$O/r02f_mt_n1.err; echo "mt rc=$?"; tail -5 $O/r02f_mt_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $O/r02f_bench_n$N.json 2>$O/r02f_bench_n$N.err; echo "n$N rc=$?"; tail -8 $O/r02f_bench_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-extras > $O/r02f_strong_n$N.json 2>$O/r02f_strong_n$N.err; echo "strong n$N rc=$?"; tail -5 $O/r02f_strong_n$N.err
  timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q 2>&1 | tail -3
fi
