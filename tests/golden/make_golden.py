"""Generate tests/golden/vectors/*.npz: outputs of the REFERENCE C backend itself.

Run where /root/reference exists, after `python oracle/build_ref.py`:

    python tests/golden/make_golden.py

For every (scene, precision, num_samples) variant built into oracle/_ref/ (the reference's own
``to_c`` output compiled with its own flags), evaluate a small seeded batch and store inputs + outputs.
The fixtures travel with the repository; tests/test_oracle_pinning.py holds the CPU oracle to them bit
for bit, and tests/test_gpu_parity.py holds the CUDA path to them on the GPU box (where neither
/root/reference nor a rebuilt _ref is needed).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import RefLib  # noqa: E402
from oracle.build_ref import HIGH_N_INSTANCES, VARIANTS, is_high_n  # noqa: E402
from teg_b200.tegp import Program  # noqa: E402
from teg_b200.workloads import bind, scene_values  # noqa: E402

OUT = os.path.join(HERE, 'vectors')
OUT_HIGH = os.path.join(HERE, 'vectors_high_n')       # config 4 at the top of its sample-count sweep: 64 instances


def main():
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(OUT_HIGH, exist_ok=True)
    only_high = '--high-n' in sys.argv
    for scene, variants in VARIANTS.items():
        prog = Program.load(os.path.join(HERE, 'programs', f'{scene}.tegp'))
        for prec, ns in variants:
            high = is_high_n(scene, ns)
            if only_high and not high:
                continue
            vals = scene_values(scene, res=(24, 24), n=HIGH_N_INSTANCES if high else 576)
            inputs = bind(prog.args, vals)
            n = max(np.size(v) for v in inputs)
            dt = np.float32 if prec == 'f32' else np.float64
            ref = RefLib(scene, dt, ns)
            out = ref.eval(inputs, threads=os.cpu_count() or 1)
            soa = np.empty((len(inputs), n), dtype=np.float64)
            for i, v in enumerate(inputs):
                soa[i] = v
            path = os.path.join(OUT_HIGH if high else OUT, f'{scene}_{prec}_N{ns}.npz')
            np.savez_compressed(path, inputs=soa, outputs=out, num_samples=ns, args=np.array(prog.args))
            print(f'{path}: inputs {soa.shape} outputs {out.shape} nnz={np.count_nonzero(out)}')


if __name__ == '__main__':
    main()
