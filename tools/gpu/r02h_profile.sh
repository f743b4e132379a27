#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi_triangle.py -m gpu -x -q 2>&1 | tail -3
M="python tools/bench_multi_tri.py --steps 3 --no-cpu --graph 0 --min-blocks 4,4"
$M > $O/r02h_mt_plain.json 2> $O/r02h_mt_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 40 --csv --log-file $O/r02h_mt_launches.csv $M > $O/r02h_ncu_l.log 2>&1
$M > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'tri_color_(primal|grad)_main' -s 30 -c 10 -o $O/r02h_prof_mt $M > $O/r02h_ncu_mt.log 2>&1
tail -3 $O/r02h_ncu_mt.log
