#!/bin/bash
# config 5 with the gather forward: launch list + one full capture of the gather kernel (after a plain run exited 0)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
M="python tools/bench_multi_tri.py --steps 3 --no-cpu --graph 0"
$M > $O/r02z_mt_plain.json 2> $O/r02z_mt_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 30 --csv --log-file $O/r02z_launches_mt.csv $M > $O/r02z_ncu_lm.log 2>&1
$M > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'tri_color_primal_(gather|tiles)' -s 4 -c 2 -o $O/r02z_prof_gather $M > $O/r02z_ncu_g.log 2>&1
ls -la $O | grep r02z
