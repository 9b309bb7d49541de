"""GPU parity: training-side pieces of the fused NeRF MLP (saved activation tiles, ReLU masks, dgrad, wgrad)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import nerf_mlp as M

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _setup(golden_configs, variant, R, S, seed=0):
    from simple_rf_b200 import nerf_program
    configs, mc = golden_configs('nerf')
    m = configs['model']
    cfg = {'main': m['coarse_model'], 'points_augmentation': m['augmentations'][0]['coarse_model'],
           'views_augmentation': m['augmentations'][1]['coarse_model']}[variant]
    g = torch.Generator().manual_seed(seed)
    params = M.init_mlp_params(cfg, g)
    params['pts_output_linear.bias'][0] += 1.0
    o = torch.rand(R, 3, generator=g) - .5
    d = torch.rand(R, 3, generator=g) - .5
    vd = F.normalize(torch.randn(R, 3, generator=g), dim=-1)
    z = torch.rand(R, S, generator=g)
    packed = nerf_program.PackedMLP(cfg).refresh({k: v.to(DEV) for k, v in params.items()})
    return cfg, params, packed, o, d, vd, z


def _hidden(params, cfg, pts, vdirs):
    """fp32 activations of every layer (oracle arithmetic)."""
    pts_in, _ = M.variant_dims(cfg)
    enc = M.positional_encoding(pts, cfg['points_positional_encoding_degree'])
    x_in = enc[:, :pts_in]
    h, hs = x_in, []
    for i in range(cfg['points_net_depth']):
        h = F.relu(F.linear(h, params[f'pts_linears.{i}.weight'], params[f'pts_linears.{i}.bias']))
        hs.append(h)
        if i == 4:
            h = torch.cat([x_in, h], -1)
    out = {'enc': enc, 'h': hs}
    if cfg['view_dependent_rgb']:
        feat = F.linear(hs[-1], params['feature_linear.weight'], params['feature_linear.bias'])
        ev = M.positional_encoding(vdirs, cfg['views_positional_encoding_degree'])
        hv = F.relu(F.linear(torch.cat([feat, enc[:, pts_in:], ev], -1), params['views_linears.0.weight'], params['views_linears.0.bias']))
        out.update(feat=feat, ev=ev, hv=hv)
    return out


@pytest.mark.parametrize('variant', ['main', 'views_augmentation'])
def test_forward_saves_activation_tiles_and_masks(golden_configs, variant):
    from simple_rf_b200 import tile_images as TI
    R, S = 50, 64                                     # 3200 rows = 25 tiles
    cfg, params, packed, o, d, vd, z = _setup(golden_configs, variant, R, S)
    sigma, rgb, acts, masks = packed.forward(o.to(DEV), d.to(DEV), z.to(DEV), vd.to(DEV), save=True)
    s2, c2 = packed.forward(o.to(DEV), d.to(DEV), z.to(DEV), vd.to(DEV))
    assert torch.equal(sigma, s2) and torch.equal(rgb, c2)
    pts = (o[:, None] + d[:, None] * z[..., None]).reshape(-1, 3)
    vflat = vd[:, None].expand(R, S, 3).reshape(-1, 3)
    ref = _hidden(params, cfg, pts, vflat)
    prog = packed.program
    n = R * S
    enc = TI.decode(acts, 0, 1)[:n].cpu()
    assert (enc[:, :63] - ref['enc']).abs().max().item() <= 1e-2          # bf16 rounding of values up to ~1
    assert (enc[:, 63] == 0).all()
    for l in (0, 4, 7):
        h = TI.decode(acts, prog.layers[l].save_slot, 4)[:n].cpu()
        tol = 2e-2 * max(1.0, ref['h'][l].abs().max().item())
        assert (h - ref['h'][l]).abs().max().item() <= tol, l
        bits = masks[:, l].reshape(-1, 8)[:n].cpu()
        got = torch.stack([(bits[:, w] >> b) & 1 for w in range(8) for b in range(32)], 1).bool()
        assert torch.equal(got, TI.decode(acts, prog.layers[l].save_slot, 4)[:n].cpu() > 0), l
    if cfg['view_dependent_rgb']:
        feat = TI.decode(acts, prog.layers[8].save_slot, 4)[:n].cpu()
        assert (feat - ref['feat']).abs().max().item() <= 2e-2 * max(1.0, ref['feat'].abs().max().item())
        hv = TI.decode(acts, prog.layers[9].save_slot, 2)[:n].cpu()
        assert (hv - ref['hv']).abs().max().item() <= 2e-2 * max(1.0, ref['hv'].abs().max().item())
        ev = TI.decode(acts, prog.v_slot, 1)[:n].cpu()
        assert (ev[:, :27] - ref['ev']).abs().max().item() <= 1e-2


def test_tile_image_codec_roundtrip():
    from simple_rf_b200 import tile_images as TI
    x = torch.randn(3 * 128, 256, device=DEV).to(torch.bfloat16).float()
    assert torch.equal(TI.decode(TI.encode(x, 3), 0, 4), x)
