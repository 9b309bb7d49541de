"""CPU suite for the "next" row f1 (patch-reprojection depth losses): oracle vs the reference's golden outputs, and the
drop-in loss classes resolve through the reference's unmodified LossComputer."""
import pytest
import torch

from oracle import losses as OL


@pytest.mark.parametrize('tag,rule', [('aug', True), ('cf', False)])
def test_oracle_matches_reference_golden(golden, tag, rule):
    g = golden('patch_loss')
    m = g['mask_nerf']
    d1 = g['depth1'].clone().requires_grad_()
    d2 = g['depth2'].clone().requires_grad_()
    m1, m2, r1, r2 = OL.patch_reprojection_masks(g['rays_o'][m], g['rays_d'][m], d1[m], d2[m], g['pixel_id'][m], g['poses'], g['k'],
                                                 g['images'], (5, 5), 0.1, rule)
    assert torch.equal(m1, g[f'{tag}_mask1']) and torch.equal(m2, g[f'{tag}_mask2'])
    loss, map1, map2 = OL.masked_depth_loss(d1[m], d2[m], m1, m2)
    assert torch.equal(loss.detach(), g[f'{tag}_loss']) and torch.equal(map1.detach(), g[f'{tag}_map1'])
    g1, g2 = torch.autograd.grad(loss, [d1, d2], allow_unused=True)
    assert torch.equal(g1, g[f'{tag}_g1'])
    assert float(g[f'{tag}_g2'].abs().max()) == 0.0 and (g2 is None or float(g2.abs().max()) == 0.0)   # the reference's aliasing quirk
    assert not bool((m1 & m2).any())
    assert 0.2 < float(m1.float().mean()) < 0.5 and 0.2 < float(m2.float().mean()) < 0.5              # the fixture exercises both masks


@pytest.mark.needs_reference
def test_dropin_losses_resolve_through_reference_loss_computer():
    from oracle import reference_harness as H
    from simple_rf_b200 import dropin
    H.import_reference()
    dropin.install()
    from loss_functions.LossComputer03 import LossComputer
    configs = {'model': {'coarse_model': {}, 'fine_model': {}, 'augmentations': [{'name': 'points_augmentation', 'coarse_model': {}}]},
               'data_loader': {'sparse_depth': {}},
               'losses': [{'name': 'AugmentationsDepthLoss91', 'weight': 0.1, 'patch_size': [5, 5], 'rmse_threshold': 0.1},
                          {'name': 'CoarseFineConsistencyLoss91', 'weight': 0.1, 'patch_size': [5, 5], 'rmse_threshold': 0.1}]}
    lc = LossComputer(configs)
    assert type(lc.losses['AugmentationsDepthLoss91']).__module__.startswith('simple_rf_b200.loss_functions')
    assert type(lc.losses['CoarseFineConsistencyLoss91']).__name__ == 'CoarseFineConsistencyLoss'
    assert lc.get_loss_weight(lc.losses['AugmentationsDepthLoss91'], 0) == 0.1


def test_dropin_losses_refuse_cpu(golden):
    """No CPU fallback: the fused masks raise on CPU tensors."""
    from simple_rf_b200.loss_functions import patch_reprojection as PR
    g = golden('patch_loss')
    with pytest.raises(RuntimeError):
        PR.patch_reprojection_masks(g['rays_o'], g['rays_d'], g['depth1'], g['depth2'], g['pixel_id'], g['poses'], g['k'], g['images'],
                                    (5, 5), 0.1, True)


@pytest.mark.needs_reference
def test_oracle_tv_loss_equals_reference_class():
    """oracle.losses.tv_loss against TotalVariationLoss04.compute_tv_loss (static, src/loss_functions/TotalVariationLoss04.py:97)."""
    from oracle import losses as OL
    from oracle import reference_harness as H
    H.import_reference()
    from loss_functions.TotalVariationLoss04 import TotalVariationLoss
    g = torch.Generator().manual_seed(0)
    planes = [torch.randn(1, 4, 13, 17, generator=g), torch.randn(1, 12, 1, 6, generator=g), torch.randn(1, 6, 9, 1, generator=g)]
    ref = TotalVariationLoss.compute_tv_loss(planes, 0.3, False)['loss_value']
    assert torch.equal(torch.as_tensor(ref), torch.as_tensor(OL.tv_loss(planes, 0.3)))
