#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
: > $O/r02_value_exp.txt
for o in "" "--min-blocks=6" "--min-blocks=4" "--min-blocks=5"; do
  timeout 300 python tools/exp_value_kernel.py $o 2>&1 | grep -v "^small\|^sliver\|^axis" >> $O/r02_value_exp.txt
done
TEGB200_QUEUE_WORKERS=3 timeout 300 python tools/exp_value_kernel.py --min-blocks=6 2>&1 | grep -v "^small\|^sliver\|^axis" >> $O/r02_value_exp.txt
cat $O/r02_value_exp.txt
