#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras --graph 0"
$B > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'tri_grad_main' -s 3 -c 1 -o $O/r02_prof_grad_sparse $B > $O/r02_ncu_grad.log 2>&1
ls -la $O | grep prof_grad
