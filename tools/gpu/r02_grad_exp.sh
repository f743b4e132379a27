#!/bin/bash
# gradient kernel of the headline workload (all outputs summed): scenes x worker blocks per SM x tile height x launch bounds
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
: > $O/r02_grad_exp.txt
for w in 2 3 4; do
  for o in "" "--tile-rows=8" "--min-blocks=4" "--tile-rows=8 --min-blocks=4" "--tile-rows=8 --min-blocks=6"; do
    TEGB200_QUEUE_WORKERS=$w timeout 300 python tools/exp_grad_kernel.py $o 2>&1 | grep -v "^small\|^sliver" >> $O/r02_grad_exp.txt
  done
done
cat $O/r02_grad_exp.txt
