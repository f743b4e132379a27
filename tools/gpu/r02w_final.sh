#!/bin/bash
# final measurements of round 2 for profiles/: bench.py exactly as the driver runs it
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $O/r02w_ref_n1.json 2>$O/r02w_ref_n1.err; echo "ref rc=$?"
  python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02w_bench_n1.json 2>$O/r02w_bench_n1.err; echo "n1 rc=$?"; tail -3 $O/r02w_bench_n1.err
  python bench.py --gpus 1 --steps 20 --warmup 5 --workload multi_tri_4096 --no-extras > $O/r02w_mt_n1.json 2>$O/r02w_mt_n1.err; echo "mt rc=$?"
  python bench.py --gpus 1 --steps 20 --warmup 5 --dtype f64 --no-extras > $O/r02w_f64_n1.json 2>$O/r02w_f64_n1.err; echo "f64 rc=$?"; tail -3 $O/r02w_f64_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02w_bench_n$N.json 2>$O/r02w_bench_n$N.err; echo "n$N rc=$?"; tail -3 $O/r02w_bench_n$N.err
fi
