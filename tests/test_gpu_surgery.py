"""GPU parity of the grid-surgery kernels (csrc/tensorf_surgery.cu) and of the drop-in tensor's in-forward model modifications:
bit-exact occupancy volumes / crop windows / boxes against the committed outputs of the unmodified reference and the CPU oracle,
resampled planes within 1e-6 (relative to 1 + |value|), and the alpha-mask rebuild at the benchmarked 331x368x220 grid through size-independent stages."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import fixtures as FX
from oracle import surgery as SG
from oracle import tensorf as TF

_bits, run_oracle_schedule = SG.pack_volume, SG.replay_golden_schedule

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _tensor(configs, mc, t, index=None):
    """One drop-in VM tensor on the GPU holding the fixture's parameters."""
    from simple_rf_b200.models.SimpleTensoRF91 import SimpleTensoRF
    model = SimpleTensoRF(configs, mc)
    module = model.coarse_model if index is None else model.augmented_models[index]['coarse_model']
    sd = dict(module.named_parameters())
    assert set(sd) == set(t['params'])
    for k, v in t['params'].items():
        sd[k].data.copy_(v)
    return model.to(DEV), module


def test_dropin_tensor_follows_the_reference_schedule(golden, golden_configs):
    """run_model_modifications at iterations 2 (rebuild + crop), 4 (resample + optimiser re-grouping) and 6 (rebuild against
    the previous mask on the old grid / old box) — every volume, window, box bit-exact; resampled planes <= 1e-6."""
    g = golden('tensorf_surgery')
    configs, mc = golden_configs('tensorf_surgery')
    o = run_oracle_schedule(configs)
    model, t = _tensor(configs, mc, FX.surgery_sets(configs, seed=41)['coarse_model'])
    opt_cfg = next(c for c in configs['optimizers'] if c['name'] == 'optimizer_main')
    opt = torch.optim.Adam(model.get_trainable_parameters(opt_cfg), betas=(opt_cfg['beta1'], opt_cfg['beta2']))
    model.optimizers = {'optimizer_nerf': opt}
    model.train()

    model.eval(); t.run_model_modifications(2); assert t.alpha_mask is None       # nothing happens outside training (:822)
    model.train()
    t.run_model_modifications(1); assert t.alpha_mask is None
    t.run_model_modifications(2)
    vol = t.alpha_mask.alpha_volume
    assert vol.dtype == torch.bool and list(vol.shape) == [1, 1, *g['volume1_shape'].tolist()]
    assert torch.equal(_bits(vol.cpu()), g['volume1_bits'])
    assert t.alpha_mask.resolution.tolist() == o['geo0']['resolution'].tolist()
    assert torch.equal(t.alpha_mask.bounding_box.cpu(), o['geo0']['bbox'])
    assert torch.equal(t.resolution.cpu(), g['shrink_resolution']) and torch.equal(t.bounding_box.cpu(), g['shrink_bbox'])
    assert int(t.num_samples) == int(g['shrink_num_samples'])
    for k, v in t.named_parameters():
        if k.startswith(('matrices', 'vectors')):
            assert v.is_contiguous() and torch.equal(v.detach().cpu(), o['params1'][k]), k

    held = {id(p) for grp in opt.param_groups for p in grp['params']}
    t.run_model_modifications(4)
    assert torch.equal(t.resolution.cpu(), g['upsample_resolution']) and int(t.num_samples) == int(g['upsample_num_samples'])
    worst = 0.0
    for k, v in t.named_parameters():
        if k.startswith(('matrices', 'vectors')):
            assert v.shape == o['params2'][k].shape, k
            want = o['params2'][k]                       # carved density planes hold values of 50: the bound is relative
            worst = max(worst, float(((v.detach().cpu() - want).abs() / (1 + want.abs())).max()))
    assert worst <= 1e-6, worst
    want = g['upsampled_matrices_density_0']
    assert float(((t.matrices_density[0].detach().cpu() - want).abs() / (1 + want.abs())).max()) <= 1e-6
    now = {id(p) for grp in opt.param_groups for p in grp['params']}
    assert {id(p) for p in t.parameters()} <= now and now != held                 # the optimiser holds the new planes
    assert [grp['name'] for grp in opt.param_groups][-2:] == ['coarse_model_tensor_params', 'coarse_model_network_params']

    # iteration 6: identical planes on both sides (the oracle's bit-exact upsampled planes), previous mask at the old resolution
    for k, v in t.named_parameters():
        if k.startswith(('matrices', 'vectors')):
            v.data.copy_(o['params2'][k])
    t.run_model_modifications(6)
    assert torch.equal(_bits(t.alpha_mask.alpha_volume.cpu()), g['volume2_bits'])
    assert torch.equal(t.resolution.cpu(), g['upsample_resolution'])             # no crop on the second rebuild (:824)
    assert torch.equal(t.bounding_box.cpu(), g['shrink_bbox'])


def test_occupied_box_is_the_reference_box(golden, golden_configs):
    from simple_rf_b200 import grid_surgery as GS
    g = golden('tensorf_surgery')
    configs, mc = golden_configs('tensorf_surgery')
    model, t = _tensor(configs, mc, FX.surgery_sets(configs, seed=41)['coarse_model'])
    model.train()
    box = t.rebuild_alpha_mask()
    assert torch.equal(box.cpu(), g['box1'])
    lo, hi, new_box = GS.crop_window(t.bounding_box, t.voxel_length, t.resolution, box, t.alpha_mask.resolution)
    assert torch.equal(lo, g['window_lo']) and torch.equal(hi, g['window_hi'])


def test_rebuild_raises_when_nothing_is_occupied(golden_configs):
    configs, mc = golden_configs('tensorf_surgery')
    sets = FX.surgery_sets(configs, seed=41)
    for i in range(3):
        sets['coarse_model']['params'][f'matrices_density.{i}'].zero_()
    model, t = _tensor(configs, mc, sets['coarse_model'])
    model.train()
    with pytest.raises(RuntimeError):
        t.rebuild_alpha_mask()


@pytest.mark.parametrize('shape,window,out', [((5, 37, 53), None, (74, 101)), ((16, 49, 59), (3, 7, 40, 45), (40, 45)),
                                              ((4, 64, 1), None, (150, 1)), ((12, 30, 30), (0, 0, 30, 30), (30, 30)),
                                              ((3, 9, 11), (2, 3, 5, 6), (17, 1)), ((48, 368, 331), None, (400, 360))])
def test_resample_plane(shape, window, out):
    """Window copy is exact; bilinear resampling equals ATen's own CUDA kernel bit for bit and the CPU oracle (F.interpolate on
    the host) to 1e-6."""
    from simple_rf_b200 import grid_surgery as GS
    g = torch.Generator().manual_seed(sum(shape))
    src = torch.randn(1, *shape, generator=g)
    got = GS.resample(src.to(DEV), out, window)
    y0, x0, h, w = window if window is not None else (0, 0, shape[1], shape[2])
    ref_src = src[..., y0:y0 + h, x0:x0 + w]
    want = F.interpolate(ref_src, size=out, mode='bilinear', align_corners=True)
    assert got.shape == want.shape and got.is_contiguous()
    if (h, w) == tuple(out):
        assert torch.equal(got.cpu(), ref_src)
    else:
        assert float((got.cpu() - want).abs().max()) <= 1e-6
        on_device = F.interpolate(ref_src.to(DEV).contiguous(), size=out, mode='bilinear', align_corners=True)
        assert torch.equal(got, on_device)


def test_pack_bits_u8_equals_float_pack():
    from simple_rf_b200 import tensorf_ops as T
    g = torch.Generator().manual_seed(0)
    for n in (1, 31, 32, 33, 190 ** 3 // 7 + 5):
        vol = (torch.rand(n, generator=g) < 0.3)
        a = T.pack_alpha_bits(vol.to(DEV))
        b = T.pack_alpha_bits(vol.float().to(DEV))
        assert torch.equal(a, b)


def test_full_size_rebuild_331x368x220():
    """BASELINE config 3 size.  The rebuild is (pointwise density -> bit) o (3^3 dilation) o (axis projections): the pointwise
    stage is compared with the oracle on 400 000 random voxels plus every voxel of 3 full grid rows-planes, the dilation with
    F.max_pool3d of the kernel's own raw bits on the host (all 26.8 M voxels), the box with the oracle's amin / amax."""
    from simple_rf_b200 import _lib as L
    from simple_rf_b200 import grid_surgery as GS
    from simple_rf_b200 import tensorf_ops as T
    from simple_rf_b200.synthetic import blocky_alpha_volume
    g = torch.Generator().manual_seed(77)
    bbox = torch.tensor([[-1.5, -1.67, -1.0], [1.5, 1.67, 1.0]])
    res = TF.vm_resolution(300 ** 3, bbox)
    assert res.tolist() == [331, 368, 220]
    t = {'params': TF.init_vm_params(res, [16, 4, 4], [4, 4, 4], generator=g), 'resolution': res, 'bbox': bbox}
    for i in range(3):
        t['params'][f'matrices_density.{i}'] *= 6.0
    FX.carve_empty_border(FX.sparsify_density(t))
    geo = SG.tensor_geometry(res, bbox)
    prev = blocky_alpha_volume(190, 10, 0.5, 0.01, g).view(1, 1, 190, 190, 190)
    prev_box = bbox * 1.02
    planes = [t['params'][f'matrices_density.{i}'].to(DEV) for i in range(3)]
    lines = [t['params'][f'vectors_density.{i}'].to(DEV) for i in range(3)]
    previous = {'bits': T.pack_alpha_bits(prev.bool().to(DEV)), 'res': [190, 190, 190], 'box_min': prev_box[0].tolist(),
                'box_size': (prev_box[1] - prev_box[0]).tolist()}
    geometry = {'box': bbox.to(DEV), 'box_min': bbox[0].tolist(), 'box_size': geo['size'].tolist(), 'res': res.tolist()}
    kw = dict(step_size=float(geo['step_size']), threshold=1e-4, softplus=False, density_offset=-10.0, previous=previous)
    volume, box = GS.rebuild_occupancy(planes, lines, geometry, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    volume, box = GS.rebuild_occupancy(planes, lines, geometry, **kw)
    e1.record()
    torch.cuda.synchronize()
    print(f'alpha-mask rebuild at 331x368x220 with a 190^3 previous mask: {e0.elapsed_time(e1):.2f} ms, occupied {volume.float().mean().item():.4f}')
    X, Y, Z = res.tolist()
    assert volume.shape == (Z, Y, X) and volume.dtype == torch.bool

    # stage 1, pointwise: re-run the occupancy kernel alone to get the raw (pre-dilation) bits
    cl_planes, cl_lines = T.to_channels_last(planes, lines)
    coords = GS.axis_coordinates(bbox.to(DEV), res.tolist())
    c_res = T._i3(res.tolist())
    words = L.load().srf_alpha_grid_words(c_res)
    raw = torch.empty((words,), dtype=torch.int32, device=DEV)
    L.call('srf_alpha_grid_occupancy', T._ptrs(cl_planes), T._ptrs(cl_lines), (ctypes.c_int * 3)(16, 4, 4), c_res, T._f3(bbox[0]),
           T._f3(geo['size']), L.ptr(coords[0]), L.ptr(coords[1]), L.ptr(coords[2]), L.ptr(previous['bits']), T._i3([190] * 3),
           T._f3(previous['box_min']), T._f3(previous['box_size']), 0, -10.0, float(geo['step_size']), 1e-4, L.ptr(raw), L.stream_handle())
    pitch = (X + 31) // 32
    raw_np = raw.cpu().numpy().view(np.uint32).reshape(Z, Y, pitch)
    raw_bits = ((raw_np[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(Z, Y, pitch * 32)[..., :X].astype(bool)
    ix = torch.cat([torch.randint(0, X, (400000,), generator=g), torch.arange(X).repeat_interleave(Y), torch.randint(0, X, (Y * Z,), generator=g)])
    iy = torch.cat([torch.randint(0, Y, (400000,), generator=g), torch.arange(Y).repeat(X), torch.arange(Y).repeat(Z)])
    iz = torch.cat([torch.randint(0, Z, (400000,), generator=g), torch.full((X * Y,), Z // 3), torch.arange(Z).repeat_interleave(Y)])
    cx, cy, cz = [c.cpu() for c in coords]
    ref_axes = [bbox[0, a] * (1 - torch.linspace(0, 1, n)) + bbox[1, a] * torch.linspace(0, 1, n) for a, n in enumerate((X, Y, Z))]
    assert all(torch.equal(a, b) for a, b in zip((cx, cy, cz), ref_axes))
    xyz = torch.stack([cx[ix], cy[iy], cz[iz]], -1)
    alpha = SG.compute_alpha(t['params'], bbox, xyz, geo['step_size'], prev, prev_box).clamp(0, 1)
    want = (alpha >= 1e-4).numpy()
    got = raw_bits[iz.numpy(), iy.numpy(), ix.numpy()]
    mismatch = int((want != got).sum())
    print(f'pointwise stage: {mismatch} of {want.size} voxels differ from the oracle ({want.mean():.4f} occupied)')
    assert 0.002 < want.mean() < 0.6
    assert mismatch == 0

    # stage 2, dilation + projection on all voxels
    pooled = F.max_pool3d(torch.from_numpy(raw_bits).float()[None, None], kernel_size=3, padding=1, stride=1)[0, 0] > 0.5
    assert torch.equal(volume.cpu(), pooled)
    occ = pooled.nonzero()
    want_box = torch.stack([torch.stack([cx[occ[:, 2]].amin(), cy[occ[:, 1]].amin(), cz[occ[:, 0]].amin()]),
                            torch.stack([cx[occ[:, 2]].amax(), cy[occ[:, 1]].amax(), cz[occ[:, 0]].amax()])])
    assert torch.equal(box.cpu(), want_box)
