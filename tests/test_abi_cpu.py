"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports every symbol the header declares.
No compute call is made here (there is no GPU), only the loud-failure contract is exercised."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope='module')
def lib_path():
    from simple_rf_b200 import build
    return build.build()


def header_symbols():
    text = (ROOT / 'include' / 'simple_rf_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(srf_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_entry_points():
    syms = header_symbols()
    assert 'srf_composite_fwd' in syms and 'srf_sample_pdf_merge' in syms and len(syms) >= 7


def test_integration_guide_maps_every_entry_point():
    """INTEGRATION.md's table names the reference lines each exported symbol replaces (or says that it has no counterpart)."""
    guide = (ROOT / 'INTEGRATION.md').read_text()
    assert [s for s in header_symbols() if f'`{s}`' not in guide] == []


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(str(lib_path))
    for s in header_symbols():
        assert hasattr(lib, s), f'{s} declared in include/simple_rf_b200.h but not exported'


def test_python_binding_covers_header(lib_path):
    from simple_rf_b200 import _lib
    assert sorted(_lib.declared_symbols()) == header_symbols()
    lib = _lib.load()
    assert lib.srf_abi_version() >= 1


def test_no_cpu_fallback(lib_path):
    """CPU tensors are refused; nothing silently routes through PyTorch or the oracle."""
    from simple_rf_b200 import _lib, ops
    z = torch.rand(4, 8)
    with pytest.raises(_lib.SimpleRFNativeError):
        ops.composite(z, None, z, z[:, :3], z[:, :3], z[:, :3], ndc=True)
    with pytest.raises(_lib.SimpleRFNativeError):
        ops.sample_pdf_merge(z, z, 4)


def test_product_does_not_import_oracle():
    pkg = ROOT / 'simple_rf_b200'
    for f in pkg.rglob('*.py'):
        src = f.read_text()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
        assert 'reference_harness' not in src, f
