"""Drive the reference's UNMODIFIED callers — `Trainer.train_one_iter` (src/Trainer10.py:65-115) and
`NerfTester.predict_frame` (src/Tester07.py:153-173) — on synthetic scenes, with either the reference's own model classes
(`SimpleNeRF17`, `SimpleTensoRF09`) or the drop-in classes of this package (`SimpleNeRF91`, `SimpleTensoRF91`,
`DataPreprocessor91`, `*Loss91`), selected purely by the names in the config dict (the reference's factories do the rest).

Integration tooling, not product path: the kernels never need it.  It locates an upstream tree ($SIMPLE_RF_REFERENCE,
`baseline/_ref` as installed by tools/install_reference.sh — git-ignored, travels with gpurun — or /root/reference), puts
`<tree>/src` on sys.path, stubs the four optional imports this image lacks (matplotlib, skimage, simplejson, deepdiff: image
I/O and config dumps only, src/Trainer10.py:15-20) and builds the `raw_data_dict` the disk loaders would have produced
(keys consumed at src/data_preprocessors/DataPreprocessor10.py:116-120,179,192-200; recipe: SURVEY.md §8c).

The synthetic scene is analytic — a slanted textured background plane plus a textured sphere in front of it, rendered by
exact ray intersection for every training camera — so images, sparse depths and camera poses are mutually consistent and a
model trained on it converges to a real surface (sharp weights, large sigma), which is what the bf16 tolerance
measurements on a *trained* field need.
"""
import copy
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy

REPO_ROOT = Path(__file__).resolve().parents[2]


# ------------------------------------------------------------------------------------------------ locating / importing
def reference_root():
    """First existing of $SIMPLE_RF_REFERENCE, <repo>/baseline/_ref, /root/reference (None if there is none)."""
    cands = [os.environ.get('SIMPLE_RF_REFERENCE'), REPO_ROOT / 'baseline' / '_ref', '/root/reference']
    for c in cands:
        if c and (Path(c) / 'src' / 'models' / 'ModelFactory02.py').exists():
            return Path(c)
    return None


def available():
    return reference_root() is not None


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return __import__(name, fromlist=['_'])
    except Exception:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m


def install_import_stubs():
    """Optional third-party imports of the reference that this image lacks; none of them is on the compute path."""
    mpl = _stub('matplotlib')
    plt = _stub('matplotlib.pyplot')
    if not hasattr(mpl, 'pyplot'):
        mpl.pyplot = plt
    sk = _stub('skimage')
    for sub in ('io', 'transform'):
        m = _stub(f'skimage.{sub}')
        if not hasattr(sk, sub):
            setattr(sk, sub, m)
    _stub('simplejson', dump=json.dump, dumps=json.dumps, load=json.load, loads=json.loads)
    _stub('deepdiff', DeepDiff=lambda a, b, **k: {} if a == b else {'values_changed': True})


class force_cpu:
    """Context manager: the reference picks its device with `torch.cuda.is_available()` (src/utils/CommonUtils04.py:25-26,
    torch.nn.DataParallel likewise), so hiding CUDA behind that one call runs its unmodified code on the host CPU inside a
    process that also uses the GPU (bench.py's cpu_baseline)."""

    def __enter__(self):
        import torch
        self._saved = torch.cuda.is_available
        torch.cuda.is_available = lambda: False
        return self

    def __exit__(self, *exc):
        import torch
        torch.cuda.is_available = self._saved


def prepare(install_dropin=True):
    """sys.path + stubs (+ shim directories of the drop-in classes).  Returns the reference root."""
    root = reference_root()
    if root is None:
        raise RuntimeError('no upstream Simple-RF tree found: run tools/install_reference.sh (needs /root/reference) or set '
                           'SIMPLE_RF_REFERENCE')
    install_import_stubs()
    src = str(root / 'src')
    if src not in sys.path:
        sys.path.insert(0, src)
    if install_dropin:
        from . import install
        install()
    return root


def load_shipped_configs(train_num, scene=None):
    """runs/training/train{NNNN}/Configs.json (+ <scene>/ModelConfigs.json when `scene` is given) of the upstream tree, with
    the two loss names that do not exist on disk patched."""
    root = reference_root()
    run = root / 'runs' / 'training' / f'train{train_num:04d}'
    configs = json.loads((run / 'Configs.json').read_text())
    for loss in configs.get('losses', []):
        loss['name'] = {'TotalVariationLoss05': 'TotalVariationLoss04',
                        'MassConcentrationLoss07': 'MassConcentrationLoss06'}.get(loss['name'], loss['name'])
    if scene is None:
        return configs
    return configs, json.loads((run / str(scene) / 'ModelConfigs.json').read_text())


DROPIN_NAMES = {'SimpleNeRF17': 'SimpleNeRF91', 'SimpleTensoRF09': 'SimpleTensoRF91', 'DataPreprocessor10': 'DataPreprocessor91',
                'AugmentationsDepthLoss11': 'AugmentationsDepthLoss91', 'CoarseFineConsistencyLoss34': 'CoarseFineConsistencyLoss91',
                'TotalVariationLoss04': 'TotalVariationLoss91'}


def use_dropin(configs, preprocessor=True, losses=True):
    """The config edit of INTEGRATION.md: same dict, names of the fused implementations."""
    cfg = copy.deepcopy(configs)
    cfg['model']['name'] = DROPIN_NAMES.get(cfg['model']['name'], cfg['model']['name'])
    if preprocessor:
        n = cfg['data_loader'].get('data_preprocessor_name', 'DataPreprocessor10')
        cfg['data_loader']['data_preprocessor_name'] = DROPIN_NAMES.get(n, n)
    if losses:
        for loss in cfg.get('losses', []):
            loss['name'] = DROPIN_NAMES.get(loss['name'], loss['name'])
    return cfg


# ------------------------------------------------------------------------------------------------ synthetic scene
def _colmap_w2c(tx, ty, tz, yaw):
    """world->camera, COLMAP/RE10K camera frame (x right, y down, z forward), as the disk loaders deliver extrinsics."""
    c, s = numpy.cos(yaw), numpy.sin(yaw)
    c2w = numpy.eye(4)
    c2w[:3, :3] = numpy.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    c2w[:3, 3] = [tx, ty, tz]
    return numpy.linalg.inv(c2w)


def _texture(u, v, phase):
    """Smooth + mid-frequency procedural rgb in [0,1]."""
    r = 0.5 + 0.35 * numpy.sin(2.1 * u + phase) * numpy.cos(1.3 * v) + 0.15 * numpy.sin(9.0 * u + 5.0 * v)
    g = 0.5 + 0.35 * numpy.cos(1.7 * u - 0.6 * v + phase) + 0.15 * numpy.sin(7.0 * v - 3.0 * u)
    b = 0.5 + 0.30 * numpy.sin(1.1 * (u + v) + 2.0 * phase) + 0.20 * numpy.cos(6.0 * u) * numpy.sin(6.0 * v)
    return numpy.clip(numpy.stack([r, g, b], -1), 0.0, 1.0)


def render_scene_view(w2c, k, h, w):
    """Exact image + z-depth of the analytic scene for one camera (all in the raw, un-normalised world)."""
    c2w = numpy.linalg.inv(w2c)
    ys, xs = numpy.meshgrid(numpy.arange(h, dtype=numpy.float64), numpy.arange(w, dtype=numpy.float64), indexing='ij')
    dirs_cam = numpy.stack([(xs - k[0, 2]) / k[0, 0], (ys - k[1, 2]) / k[1, 1], numpy.ones_like(xs)], -1)   # z = 1 => t is z-depth
    d = dirs_cam @ c2w[:3, :3].T
    o = c2w[:3, 3]
    # background plane  n . p = c  (slanted, 6-8 units away)
    n, c = numpy.array([0.18, -0.06, 1.0]), 7.0
    t_plane = (c - n @ o) / (d @ n)
    hit = o + t_plane[..., None] * d
    rgb = _texture(hit[..., 0], hit[..., 1], 0.3)
    depth = t_plane
    # sphere in front
    centre, rad = numpy.array([0.15, 0.05, 3.6]), 0.85
    oc = o - centre
    a = (d * d).sum(-1)
    b = 2.0 * (d @ oc)
    cc = oc @ oc - rad * rad
    disc = b * b - 4 * a * cc
    t_s = numpy.where(disc > 0, (-b - numpy.sqrt(numpy.maximum(disc, 0))) / (2 * a), numpy.inf)
    on = (t_s > 0) & (t_s < t_plane)
    ph = o + numpy.where(on, t_s, 0.0)[..., None] * d - centre
    rgb_s = _texture(3.0 * numpy.arctan2(ph[..., 0], -ph[..., 2]), 3.0 * ph[..., 1], 1.7)
    rgb = numpy.where(on[..., None], rgb_s, rgb)
    depth = numpy.where(on, t_s, depth)
    return (rgb * 255.0 + 0.5).astype(numpy.uint8), depth


def synthetic_raw_data(kind='llff', num_views=3, resolution=None, sparse_points=2000, seed=0, tensorf=False):
    """The dict `get_data_preprocessor(cfg, mode='train', raw_data_dict=...)` consumes.  kind: 'llff' (756x1008, f 815.13) or
    're10k' (576x1024, f 493.91); `resolution` (h, w) shrinks the frame and scales the intrinsics (seconds-scale tests)."""
    import pandas
    if kind == 'llff':
        h0, w0, f0 = 756, 1008, 815.1316
    else:
        h0, w0, f0 = 576, 1024, 493.9102
    h, w = (h0, w0) if resolution is None else (int(resolution[0]), int(resolution[1]))
    f = f0 * h / h0
    k = numpy.array([[f, 0.0, w / 2.0], [0.0, f, h / 2.0], [0.0, 0.0, 1.0]])
    frame_nums = numpy.arange(num_views) * 5 + 3
    rng = numpy.random.RandomState(seed)
    images, w2cs, sparse = [], [], {}
    lo, hi = numpy.inf, 0.0
    for i, fn in enumerate(frame_nums):
        s = i - (num_views - 1) / 2.0
        w2c = _colmap_w2c(0.45 * s, 0.05 * ((i % 2) * 2 - 1), 0.02 * s, 0.035 * s)
        img, depth = render_scene_view(w2c, k, h, w)
        images.append(img)
        w2cs.append(w2c)
        n = min(sparse_points, h * w)
        flat = rng.choice(h * w, size=n, replace=False)
        ys, xs = flat // w, flat % w
        sparse[int(fn)] = pandas.DataFrame({'x': xs.astype(numpy.float64), 'y': ys.astype(numpy.float64), 'depth': depth[ys, xs],
                                            'reprojection_error': rng.uniform(0.2, 1.5, size=n)})
        lo, hi = min(lo, float(depth.min())), max(hi, float(depth.max()))
    raw = {'frame_nums': frame_nums,
           'nerf_data': {'images': numpy.stack(images), 'extrinsics': numpy.stack(w2cs), 'intrinsics': numpy.stack([k] * num_views),
                         'bounds': numpy.array([0.9 * lo, 1.1 * hi]), 'resolution': (h, w)},
           'sparse_depth_data': sparse}
    if tensorf:
        raw['tensorf_data'] = {'bounding_box': [[-1.5, -1.67, -1.0], [1.5, 1.67, 1.0]]}
    return raw


def test_pose(raw_data_dict, t=0.37):
    """A w2c test pose (raw world) on a small arc between the training cameras, for NerfTester.predict_frame."""
    a = 2 * numpy.pi * t
    return _colmap_w2c(0.3 * numpy.cos(a), 0.04 * numpy.sin(a), 0.03 * numpy.sin(2 * a), 0.03 * numpy.cos(a))


# ------------------------------------------------------------------------------------------------ the callers
def complete_configs(configs, device_ids, seed=0):
    cfg = copy.deepcopy(configs)
    cfg['device'] = list(device_ids)
    cfg['seed'] = seed
    cfg.setdefault('train_num', 0)
    cfg['data_loader'].setdefault('scene_id', 'synthetic')
    return cfg


def make_trainer(configs, raw_data_dict, out_dir=None, seed=0):
    """-> (trainer, model, model_configs): the objects src/Trainer10.py:536-571 (start_training) builds, with the disk loader
    replaced by `raw_data_dict`.  `configs['device']` decides CPU ([] / no CUDA) or GPU ([0])."""
    prepare()
    import torch
    import Trainer10
    from data_preprocessors.DataPreprocessorFactory01 import get_data_preprocessor
    from loss_functions.LossComputer03 import LossComputer
    from models.ModelFactory02 import get_model
    Trainer10.init_seeds(seed)
    pre = get_data_preprocessor(configs, mode='train', raw_data_dict=copy.deepcopy(raw_data_dict))
    model_configs = pre.get_model_configs()
    device = Trainer10.CommonUtils.get_device(configs['device'])
    model = get_model(configs, model_configs=model_configs).to(device)
    model = torch.nn.DataParallel(model, device_ids=configs['device'] if device.type == 'cuda' else None)
    loss_computer = LossComputer(configs)
    optimizers, lr_decayers = Trainer10.get_optimizers(configs, model)
    out_dir = Path(out_dir or tempfile.mkdtemp(prefix='srf_trainer_'))
    trainer = Trainer10.Trainer(configs, model_configs, pre, None, model, loss_computer, optimizers, lr_decayers, out_dir,
                                configs['device'], verbose_log=False)
    return trainer, model, model_configs


def step_learning_rates(trainer, iter_num):
    """The learning-rate update Trainer.train() applies after every train_one_iter (src/Trainer10.py:303-308)."""
    for key in trainer.optimizers.keys():
        scale = trainer.lr_decayers[key].get_learning_rate_scale(iter_num)
        for group in trainer.optimizers[key].param_groups:
            group['lr'] = group['lr'] * scale


def make_tester(train_configs, model_configs, device_ids):
    """-> NerfTester (src/Tester07.py:30-48) with a fresh model of the class `train_configs` names."""
    root = prepare()
    import Tester07
    test_configs = {'device': list(device_ids), 'database_dirpath': train_configs.get('database_dirpath', 'synthetic')}
    return Tester07.NerfTester(copy.deepcopy(train_configs), copy.deepcopy(model_configs), test_configs, root)
