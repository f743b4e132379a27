#!/bin/bash
# round 2, call a: profile the kernels exactly as round 1 shipped them (64-register launch bounds)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
$B > $O/r02a_bench_plain.json 2> $O/r02a_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02a_launches.csv $B > $O/r02a_ncu_launches.log 2>&1
$B > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'tri_(primal|grad)_main' -s 6 -c 4 -o $O/r02a_prof_tri $B > $O/r02a_ncu_tri.log 2>&1
M="python tools/bench_multi_tri.py --steps 3 --no-cpu --graph 0"
$M > $O/r02a_mt_plain.json 2> $O/r02a_mt_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:'tri_color_(primal|grad)_main' -s 30 -c 11 -o $O/r02a_prof_mt $M > $O/r02a_ncu_mt.log 2>&1
python bench.py --steps 20 --warmup 3 > $O/r02a_bench_full.json 2> $O/r02a_bench_full.err
ls -la $O | tail -20
