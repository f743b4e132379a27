#!/bin/bash
# round 2: evidence for profiles/ from the code as committed (run AFTER the same commands exited 0 without ncu)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extras --graph 0"
$B > $O/r02u_plain.json 2> $O/r02u_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file $O/r02u_launches_tri.csv $B > $O/r02u_ncu_l.log 2>&1
$B > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'tri_(primal|grad)_(main|tiles)' -s 12 -c 4 -o $O/r02u_prof_tri $B > $O/r02u_ncu_tri.log 2>&1
M="python tools/bench_multi_tri.py --steps 3 --no-cpu --graph 0"
$M > $O/r02u_mt_plain.json 2> $O/r02u_mt_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file $O/r02u_launches_mt.csv $M > $O/r02u_ncu_lm.log 2>&1
$M > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'tri_color_(primal|grad)_(main|tiles|gather)' -s 8 -c 8 -o $O/r02u_prof_mt $M > $O/r02u_ncu_mt.log 2>&1
ls -la $O | grep r02u
