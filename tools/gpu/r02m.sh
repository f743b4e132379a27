#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extras"
$B > $O/r02m_plain.json 2> $O/r02m_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/r02m_launches.csv $B > $O/r02m_ncu_l.log 2>&1
$B > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'tri_(primal|grad)_main' -s 6 -c 2 -o $O/r02m_prof_tri $B > $O/r02m_ncu_tri.log 2>&1
tail -2 $O/r02m_ncu_tri.log
