import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'needs_reference: imports the upstream checkout (build container only)')


def pytest_collection_modifyitems(config, items):
    from oracle import reference_harness
    have_ref = reference_harness.available()
    have_gpu = torch.cuda.is_available()
    for item in items:
        if 'needs_reference' in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason='upstream checkout not present'))
        if 'gpu' in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason='no CUDA device'))


def load_golden(name):
    with np.load(GOLDEN / f'{name}.npz') as f:
        return {k: torch.from_numpy(f[k]) for k in f.files}


def load_configs(name):
    d = json.loads((GOLDEN / f'{name}_configs.json').read_text())
    return d['configs'], d['model_configs']


@pytest.fixture(scope='session')
def golden():
    return load_golden


@pytest.fixture(scope='session')
def golden_configs():
    return load_configs
