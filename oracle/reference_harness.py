"""Import the UNMODIFIED upstream Simple-RF modules for pinning the oracle (build container only).

Test infrastructure.  `/root/reference` does not exist on the GPU box, so nothing in the
`-m gpu` tests, `smoke()` or `bench.py` may call into this module; it is used by
`oracle/generate_golden.py` and by the CPU tests marked `needs_reference` (auto-skipped when
the checkout is absent).

Recipe: the model modules need numpy, torch and utils.CommonUtils04, whose only
obstacle is an unused `from matplotlib import pyplot` (src/utils/CommonUtils04.py:9) -> stub.
"""
import copy
import json
import os
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get('SIMPLE_RF_REFERENCE', '/root/reference'))


def available():
    return (REFERENCE_ROOT / 'src' / 'models' / 'ModelFactory02.py').exists()


def _install_stubs():
    for name in ('matplotlib', 'matplotlib.pyplot'):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules['matplotlib'], 'pyplot'):
        sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    # the data preprocessors import skimage for image I/O only (src/data_preprocessors/DataPreprocessor10.py:10-11)
    for name in ('skimage', 'skimage.io', 'skimage.transform'):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    for sub in ('io', 'transform'):
        if not hasattr(sys.modules['skimage'], sub):
            setattr(sys.modules['skimage'], sub, sys.modules[f'skimage.{sub}'])


def import_reference():
    """Put <reference>/src on sys.path and return (get_model, CommonUtils04)."""
    if not available():
        raise RuntimeError(f'reference checkout not found at {REFERENCE_ROOT}')
    _install_stubs()
    src = str(REFERENCE_ROOT / 'src')
    if src not in sys.path:
        sys.path.insert(0, src)
    from models.ModelFactory02 import get_model
    from utils import CommonUtils04
    return get_model, CommonUtils04


def load_configs(train_num, scene):
    """Shipped fixtures: runs/training/train{NNNN}/Configs.json + <scene>/ModelConfigs.json."""
    run = REFERENCE_ROOT / 'runs' / 'training' / f'train{train_num:04d}'
    configs = json.loads((run / 'Configs.json').read_text())
    model_configs = json.loads((run / scene / 'ModelConfigs.json').read_text())
    # shipped TensoRF configs name loss modules that do not exist on disk
    for loss in configs.get('losses', []):
        loss['name'] = {'TotalVariationLoss05': 'TotalVariationLoss04',
                        'MassConcentrationLoss07': 'MassConcentrationLoss06'}.get(loss['name'], loss['name'])
    return configs, model_configs


def shrink(model_configs, factor):
    """Scale resolution + intrinsics down so CPU fixtures stay small."""
    mc = copy.deepcopy(model_configs)
    h, w = mc['resolution']
    mc['resolution'] = [h // factor, w // factor]
    for k in mc['intrinsics']:
        for r in range(2):
            for c in range(3):
                k[r][c] = k[r][c] / factor
    return mc


def build_model(configs, model_configs):
    get_model, _ = import_reference()
    return get_model(configs, model_configs=model_configs)
