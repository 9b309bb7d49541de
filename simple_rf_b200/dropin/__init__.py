"""Make the drop-in classes discoverable by the reference's UNMODIFIED factory.

src/models/ModelFactory02.py:10-22 imports `models.<name>` and instantiates the class called `<name>` minus
its two-digit suffix.  `install()` appends a directory of one-line shim modules (`SimpleNeRF91.py`,
`SimpleTensoRF91.py`) to the `__path__` of the reference's already-importable `models` package, so a config
with `"model": {"name": "SimpleNeRF91", ...}` resolves to simple_rf_b200's class.  The alternative with no
Python call at all is to copy / symlink the two shim files into `<reference>/src/models/`.
"""
from pathlib import Path

SHIM_DIR = Path(__file__).resolve().parent / 'models_shims'
LOSS_SHIM_DIR = Path(__file__).resolve().parent / 'loss_shims'
PREPROC_SHIM_DIR = Path(__file__).resolve().parent / 'preproc_shims'


def install():
    """Models (`SimpleNeRF91`, `SimpleTensoRF91`) and the fused patch-reprojection losses (`AugmentationsDepthLoss91`,
    `CoarseFineConsistencyLoss91`, resolved by src/loss_functions/LossComputer03.py:21-32 the same way)."""
    import models  # the reference's package: <reference>/src must already be on sys.path
    if str(SHIM_DIR) not in list(models.__path__):
        models.__path__.append(str(SHIM_DIR))
    try:
        import loss_functions
        if str(LOSS_SHIM_DIR) not in list(loss_functions.__path__):
            loss_functions.__path__.append(str(LOSS_SHIM_DIR))
    except ImportError:
        pass
    try:                                   # `DataPreprocessor91` (fused batch assembly), src/data_preprocessors/DataPreprocessorFactory01.py:15-22
        import data_preprocessors
        if str(PREPROC_SHIM_DIR) not in list(data_preprocessors.__path__):
            data_preprocessors.__path__.append(str(PREPROC_SHIM_DIR))
    except ImportError:
        pass
    return SHIM_DIR
