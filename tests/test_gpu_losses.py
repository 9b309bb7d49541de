"""GPU parity for the "next" row f1: fused patch-reprojection masks + the drop-in loss classes (through the C ABI) against
the reference's golden outputs and the oracle."""
import pytest
import torch

from oracle import generate_golden as GG
from oracle import losses as OL

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _dev(g, *keys):
    return [g[k].to(DEV) for k in keys]


@pytest.mark.parametrize('tag,rule', [('aug', True), ('cf', False)])
def test_masks_and_rmse_vs_reference_golden(golden, tag, rule):
    """Masks identical to the reference's (stated tolerance: at most 1e-3 of the rays may differ - a reprojection within one
    ulp of a half-integer or two RMSEs within rounding of each other; measured 0), RMSEs within 1e-5."""
    from simple_rf_b200.loss_functions import patch_reprojection as PR
    g = golden('patch_loss')
    m = g['mask_nerf']
    sub = {k: g[k][m] for k in ('rays_o', 'rays_d', 'depth1', 'depth2', 'pixel_id')}
    m1, m2, r1, r2 = PR.patch_reprojection_masks(*_dev(sub, 'rays_o', 'rays_d', 'depth1', 'depth2', 'pixel_id'), *_dev(g, 'poses', 'k', 'images'),
                                                 (5, 5), 0.1, rule, return_rmse=True)
    n = int(m.sum())
    bad = int((m1.cpu() != g[f'{tag}_mask1']).sum()) + int((m2.cpu() != g[f'{tag}_mask2']).sum())
    print('mask mismatches', bad, 'of', 2 * n)
    assert bad <= 1e-3 * 2 * n
    fin = torch.isfinite(g[f'{tag}_rmse1']) & torch.isfinite(g[f'{tag}_rmse2'])
    assert (r1.cpu() - g[f'{tag}_rmse1'])[fin].abs().max().item() <= 1e-5
    assert (r2.cpu() - g[f'{tag}_rmse2'])[fin].abs().max().item() <= 1e-5


def _dicts(g, depth1, depth2):
    n = depth1.shape[0]
    input_dict = {'target_rgb': torch.zeros(n, 3, device=DEV), 'common_data': {'images': g['images'].to(DEV), 'resolution': (int(g['h']), int(g['w']))},
                  'indices_mask_nerf': g['mask_nerf'].to(DEV), 'indices_mask_sparse_depth': (~g['mask_nerf']).to(DEV),
                  'pixel_id': g['pixel_id'].to(DEV), 'iter_num': 0}
    output_dict = {'rays_o': g['rays_o'].to(DEV), 'rays_d': g['rays_d'].to(DEV), 'intrinsics': g['k'].to(DEV)[None].expand(n, 3, 3),
                   'extrinsics_all': g['poses'].to(DEV), 'depth_coarse': depth1, 'depth_fine': depth2,
                   'points_augmentation_depth_coarse': depth2}
    return input_dict, output_dict


def test_dropin_loss_classes_vs_reference_golden(golden):
    """Loss values and depth gradients of the drop-in classes against the reference's (fp32 reductions in a different order:
    1e-5 relative; gradient of the second depth is exactly zero, as in the reference)."""
    from simple_rf_b200.loss_functions.AugmentationsDepthLoss91 import AugmentationsDepthLoss
    from simple_rf_b200.loss_functions.CoarseFineConsistencyLoss91 import CoarseFineConsistencyLoss
    g = golden('patch_loss')
    lcfg = {'patch_size': [5, 5], 'rmse_threshold': 0.1}
    configs = {'model': {'coarse_model': {}, 'fine_model': {}, 'augmentations': [{'name': 'points_augmentation', 'coarse_model': {}}]},
               'data_loader': {}}
    for tag, obj in (('aug', AugmentationsDepthLoss(configs, lcfg)), ('cf', CoarseFineConsistencyLoss(configs, lcfg))):
        d1 = g['depth1'].to(DEV).requires_grad_()
        d2 = g['depth2'].to(DEV).requires_grad_()
        input_dict, output_dict = _dicts(g, d1, d2)
        out = obj.compute_loss(input_dict, output_dict, None, return_loss_maps=True)
        ref = g[f'{tag}_loss'].item()
        assert abs(out['loss_value'].item() - ref) <= 1e-5 * abs(ref), (tag, out['loss_value'].item(), ref)
        g1, g2 = torch.autograd.grad(out['loss_value'], [d1, d2], allow_unused=True)
        assert (g1.cpu() - g[f'{tag}_g1']).abs().max().item() <= 1e-6 * g[f'{tag}_g1'].abs().max().item()
        assert g2 is None or float(g2.abs().max()) == 0.0
        assert len(out['loss_maps']) == 2


def test_masks_vs_oracle_at_training_batch_size():
    """A 4096-ray batch on a 378 x 504 frame (the reference's down-scaled LLFF frames), including depths that send the
    reprojection behind the camera or far outside the frame."""
    from simple_rf_b200.loss_functions import patch_reprojection as PR
    a = GG.patch_loss_inputs(num_rays=4096, h=378, w=504, seed=3)
    m1r, m2r, r1r, r2r = OL.patch_reprojection_masks(a['rays_o'], a['rays_d'], a['depth1'], a['depth2'], a['pixel_id'], a['poses'], a['k'],
                                                     a['images'], (5, 5), 0.1, True)
    d = lambda k: a[k].to(DEV)
    m1, m2 = PR.patch_reprojection_masks(d('rays_o'), d('rays_d'), d('depth1'), d('depth2'), d('pixel_id'), d('poses'), d('k'), d('images'),
                                         (5, 5), 0.1, True)
    bad = int((m1.cpu() != m1r).sum()) + int((m2.cpu() != m2r).sum())
    assert bad <= 8, bad
    # empty batch
    e = lambda k: a[k][:0].to(DEV)
    m1, m2 = PR.patch_reprojection_masks(e('rays_o'), e('rays_d'), e('depth1'), e('depth2'), e('pixel_id'), d('poses'), d('k'), d('images'),
                                         (5, 5), 0.1, True)
    assert m1.shape == (0,)


def test_masks_with_non_contiguous_and_non_fp32_inputs():
    """Strided views (RGBA images sliced `[..., :3]` as DataPreprocessor10.py:812 produces, column slices of wider ray tensors)
    and fp64 depths: every argument becomes a converted TEMPORARY; the temporaries must all stay alive until the launch
    (a freed one would be handed to the next same-sized temporary and two kernel arguments would alias)."""
    from simple_rf_b200.loss_functions import patch_reprojection as PR
    a = GG.patch_loss_inputs(num_rays=2048, h=120, w=160, seed=5)
    d = lambda k: a[k].to(DEV)
    want = PR.patch_reprojection_masks(d('rays_o'), d('rays_d'), d('depth1'), d('depth2'), d('pixel_id'), d('poses'), d('k'), d('images'),
                                       (5, 5), 0.1, True, return_rmse=True)
    rgba = torch.cat([d('images'), torch.ones_like(d('images')[..., :1])], -1)
    wide = torch.cat([d('rays_o'), d('rays_d')], 1)                               # [R,6]: both ray tensors are column slices
    got = PR.patch_reprojection_masks(wide[:, :3], wide[:, 3:], d('depth1').double(), d('depth2').double(), d('pixel_id'),
                                      d('poses').double(), d('k'), rgba[..., :3], (5, 5), 0.1, True, return_rmse=True)
    for w_, g_ in zip(want, got):
        assert torch.equal(w_, g_) if w_.dtype != torch.float32 else torch.equal(torch.nan_to_num(w_), torch.nan_to_num(g_))


def test_fused_tv_loss_vs_oracle_and_autograd():
    """srf_tv_loss (loss + gradient in one launch) against the restated TotalVariationLoss04.compute_tv_loss differentiated by
    autograd: the shipped plane shapes of the augmentation tensor, a 1-row / 1-column plane (numel clamp :105-106) and a
    non-contiguous plane."""
    from simple_rf_b200.loss_functions.TotalVariationLoss91 import tv_loss
    g = torch.Generator().manual_seed(3)
    shapes = [(1, 4, 37, 41), (1, 4, 29, 41), (1, 48, 29, 37), (1, 12, 1, 9), (1, 12, 7, 1), (1, 16, 180, 163)]
    planes = [torch.randn(s, generator=g).to(DEV).requires_grad_() for s in shapes]
    planes.append(torch.randn(1, 8, 20, 22, generator=g).to(DEV).transpose(2, 3).requires_grad_())       # strided view
    ref_planes = [p.detach().cpu().double().requires_grad_() for p in planes]
    w = 0.37
    mine = tv_loss(planes, w)
    ref = OL.tv_loss(ref_planes, w)
    assert abs(mine.item() - ref.item()) <= 1e-5 * abs(ref.item())
    (mine * 1.7).backward()
    (ref * 1.7).backward()
    for p, r in zip(planes, ref_planes):
        assert p.grad.shape == r.grad.shape
        err = (p.grad.cpu().double() - r.grad).abs().max().item()
        assert err <= 1e-5 * r.grad.abs().max().item(), err
