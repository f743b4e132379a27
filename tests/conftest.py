import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PROGRAM_DIR = os.path.join(ROOT, 'tests', 'golden', 'programs')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_program(scene: str):
    from teg_b200.tegp import Program
    return Program.load(os.path.join(PROGRAM_DIR, f'{scene}.tegp'))


def program_text(scene: str) -> str:
    with open(os.path.join(PROGRAM_DIR, f'{scene}.tegp')) as f:
        return f.read()


@pytest.fixture(scope='session')
def cuda_device():
    from teg_b200 import runtime
    if runtime.device_count() < 1:
        pytest.fail('gpu test selected but no CUDA device is visible (the CUDA path must not be skipped)')
    return 0
