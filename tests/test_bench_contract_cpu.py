"""CPU: the reference arm of bench.py (`--impl reference`: the UNMODIFIED reference from baseline/_ref on the host cores) prints ONE JSON
line with the contract's keys; the product arm must refuse to run without a GPU instead of falling back to anything."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, str(ROOT / 'bench.py'), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    if not (ROOT / 'baseline' / '_ref' / 'src' / 'Tester07.py').exists():
        pytest.skip('baseline/_ref not installed (tools/install_reference.sh)')
    r = _run('--impl', 'reference', '--steps', '1', '--warmup', '1')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and 'unavailable' not in d
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
                'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['metric'] == 'rendered_rays_per_sec' and d['unit'] == 'rays/s' and d['higher_is_better'] is True
    assert d['steps'] == 1 and d['warmup'] == 1 and d['n_gpus'] == 1 and d['vs_baseline'] is None
    assert d['config']['workload'] == 'simple_nerf_frame_render' and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'reference' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 100 < d['value'] < 1e6                                  # a CPU rate, not a GPU one


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    r = _run('--main-only', '--steps', '1', '--warmup', '1', timeout=300)
    assert r.returncode != 0
    assert not any(ln.startswith('{') and '"value"' in ln for ln in r.stdout.splitlines())
