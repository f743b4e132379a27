#!/bin/bash
# what the driver runs at round end: the -m gpu suite, smoke(), both arms of bench.py
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > $O/r02_validate_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_validate_pytest.log; tail -4 $O/r02_validate_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02_validate_ref.json 2>$O/r02_validate_ref.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_validate_bench.json 2>$O/r02_validate_bench.err; echo "bench rc=$?"; tail -2 $O/r02_validate_bench.err
python -c "
import json; d=json.load(open('$O/r02_validate_bench.json')); r=json.load(open('$O/r02_validate_ref.json')); print(d['value'], d['ms_per_step'], d['check']['ok'], d['e2e']['value'], 'ref', r['value'], r['ms_per_step'], 'same config', d['config']==r['config'])"
