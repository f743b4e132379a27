#!/bin/bash
# round 2, call b: the whole -m gpu suite + a default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02b_pytest.log
tail -15 $O/r02b_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 > $O/r02b_bench.json 2> $O/r02b_bench.err; echo "bench rc=$?"
tail -3 $O/r02b_bench.err
