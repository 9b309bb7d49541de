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
def test_forward_saves_activation_tiles(golden_configs, variant):
    from simple_rf_b200 import tile_images as TI
    R, S = 50, 64                                     # 3200 rows = 25 tiles
    cfg, params, packed, o, d, vd, z = _setup(golden_configs, variant, R, S)
    sigma, rgb, acts = packed.forward(o.to(DEV), d.to(DEV), z.to(DEV), vd.to(DEV), save=True)
    s2, c2 = packed.forward(o.to(DEV), d.to(DEV), z.to(DEV), vd.to(DEV))
    assert torch.equal(sigma, s2) and torch.equal(rgb, c2)
    pts = (o[:, None] + d[:, None] * z[..., None]).reshape(-1, 3)
    vflat = vd[:, None].expand(R, S, 3).reshape(-1, 3)
    ref = _hidden(params, cfg, pts, vflat)
    prog = packed.program
    n = R * S
    # the non-zero mask words written next to the tiles (what the dgrad chain reads as ReLU masks): bit-exact against the host
    # restatement, for every image an epilogue saved (the encoding images E / V carry no mask)
    from simple_rf_b200.nerf_program import act_tile_images
    assert acts.shape[1] == act_tile_images(prog.act_slots)
    host = TI.add_masks(acts[:, :prog.act_slots].contiguous())
    words, hwords = (x[:, prog.act_slots:].reshape(acts.shape[0], -1).view(torch.int32) for x in (acts, host))
    for l in range(prog.num_layers):
        s0 = prog.layers[l].save_slot if prog.layers[l].relu else -1      # the words are ReLU masks: x > 0
        for s_ in range(s0, s0 + prog.layers[l].n // 64 if s0 >= 0 else s0):
            base = (s_ // 16) * 4096 + (s_ % 16) * 256
            assert torch.equal(words[:, base:base + 256], hwords[:, base:base + 256]), (l, s_)
    enc = TI.decode(acts, 0, 1)[:n].cpu()
    assert (enc[:, :63] - ref['enc']).abs().max().item() <= 1e-2          # bf16 rounding of values up to ~1
    assert (enc[:, 63] == 0).all()
    for l in (0, 4, 7):
        h = TI.decode(acts, prog.layers[l].save_slot, 4)[:n].cpu()
        tol = 2e-2 * max(1.0, ref['h'][l].abs().max().item())
        assert (h - ref['h'][l]).abs().max().item() <= tol, l
        assert torch.equal(h > 0, ref['h'][l] > 0) or ((h > 0) != (ref['h'][l] > 0)).float().mean() < 1e-3, l   # ReLU pattern
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


def test_wgrad_kernel_vs_matmul():
    """dW = dZ^T X and db = colsum(dZ) from tile images, against torch on the same bf16-rounded operands."""
    import ctypes
    from simple_rf_b200 import _lib, nerf_program as NP, tile_images as TI
    assert _lib.load().srf_wgrad_item_bytes() == ctypes.sizeof(NP.WgradItem)
    g = torch.Generator(device=DEV).manual_seed(0)
    tiles = 37
    M_ = tiles * 128
    X = torch.randn(M_, 6 * 64, device=DEV, generator=g).to(torch.bfloat16).float()          # slots 0..5
    dZ = (torch.randn(M_, 6 * 64, device=DEV, generator=g) * 0.1).to(torch.bfloat16).float()  # slots 0..5
    dZ[-40:] = 0                                                                              # padded rows carry no gradient
    acts, dz = TI.add_masks(TI.encode(X, tiles)), TI.encode(dZ, tiles)
    # item 0: a 256x256 hidden layer (dz 0..3, x 1..4) with bias; item 1: its 63-wide encoding block (x slot 0, columns 0..62
    # -> weight columns 0..62 of a [256, 319] matrix); item 2: a 3-row head over 128 inputs (dz 4..5, x 4..5)
    grads = torch.zeros(256 * 256 + 256 + 256 * 319 + 3 * 128 + 3, device=DEV)
    o_w1, o_b1, o_w2, o_w3, o_b3 = 0, 256 * 256, 256 * 256 + 256, 256 * 256 + 256 + 256 * 319, 256 * 256 + 256 + 256 * 319 + 3 * 128
    items = [NP.WgradItem(0, 4, 1, 4, 256, 0, 256, 0, 256, 1, o_w1, o_b1),
             NP.WgradItem(0, 4, 0, 1, 256, 0, 63, 0, 319, 0, o_w2, 0),
             NP.WgradItem(4, 2, 4, 2, 3, 0, 128, 0, 128, 1, o_w3, o_b3)]
    NP.run_wgrad(items, acts, dz, grads)
    torch.cuda.synchronize()
    ref1 = dZ[:, :256].T @ X[:, 64:320]
    ref2 = dZ[:, :256].T @ X[:, :63]
    ref3 = dZ[:, 256:259].T @ X[:, 256:384]
    tol = lambda r: 2e-3 * r.abs().max().item()
    assert (grads[o_w1:o_b1].view(256, 256) - ref1).abs().max().item() <= tol(ref1)
    assert (grads[o_b1:o_w2] - dZ[:, :256].sum(0)).abs().max().item() <= 2e-3 * dZ[:, :256].sum(0).abs().max().item()
    got2 = grads[o_w2:o_w3].view(256, 319)
    assert (got2[:, :63] - ref2).abs().max().item() <= tol(ref2) and (got2[:, 63:] == 0).all()
    assert (grads[o_w3:o_b3].view(3, 128) - ref3).abs().max().item() <= tol(ref3)
    assert (grads[o_b3:] - dZ[:, 256:259].sum(0)).abs().max().item() <= 2e-3 * dZ[:, 256:259].sum(0).abs().max().item()


class _RoundBF16(torch.autograd.Function):
    """bf16 rounding of a value in the forward and of its gradient in the backward (the kernels' arithmetic model)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


def _bf16_model_forward(p, cfg, pts, vdirs):
    """MLP forward with the tensor-core kernels' rounding points: bf16 weights and A operands, fp32 accumulation,
    fp32 bias/activation, heads on the un-rounded fp32 activations; gradients rounded where the kernels round dZ."""
    rnd = _RoundBF16.apply
    rw = lambda w: w + (w.to(torch.bfloat16).float() - w).detach()        # bf16 value, identity gradient to the fp32 weight
    pts_in, _ = M.variant_dims(cfg)
    enc = M.positional_encoding(pts, cfg['points_positional_encoding_degree']).to(torch.bfloat16).float()
    x_in = enc[:, :pts_in]
    h = x_in
    h32 = None
    for i in range(cfg['points_net_depth']):
        h32 = F.relu(F.linear(h, rw(p[f'pts_linears.{i}.weight']), p[f'pts_linears.{i}.bias']))
        h = rnd(h32)
        if i == 4:
            h = torch.cat([x_in, h], -1)
    head = F.linear(h32, p['pts_output_linear.weight'], p['pts_output_linear.bias'])
    sigma = F.relu(head[:, 0:1])
    if not cfg['view_dependent_rgb']:
        return sigma, torch.sigmoid(head[:, 1:4])
    feat = rnd(F.linear(h, rw(p['feature_linear.weight']), p['feature_linear.bias']))
    ev = M.positional_encoding(vdirs, cfg['views_positional_encoding_degree']).to(torch.bfloat16).float()
    hv32 = F.relu(F.linear(torch.cat([feat, enc[:, pts_in:], ev], -1), rw(p['views_linears.0.weight']), p['views_linears.0.bias']))
    rgb = torch.sigmoid(F.linear(hv32, p['views_output_linear.weight'], p['views_output_linear.bias']))
    return sigma, rgb


@pytest.mark.parametrize('variant', ['main', 'points_augmentation', 'views_augmentation'])
@pytest.mark.parametrize('R,S', [(40, 64), (37, 192)])
def test_mlp_backward_vs_autograd(golden_configs, variant, R, S):
    """Parameter gradients from the hand-written dgrad + wgrad kernels.
    (a) against autograd through a torch model with the kernels' bf16 rounding points: <= 5e-2 of max |g| per tensor (measured 0.3-3 %);
    (b) against fp32 autograd through the oracle MLP: the stated looser bound of the all-bf16-operand path (activations,
        weights and layer gradients are bf16 tensor-core operands): relative L2 error <= 0.15 per tensor (measured 0.3 % on
        the heads, 2-5 % on the top layers, up to 11 % on the lowest layers after 9 bf16 dgrad GEMMs)."""
    import ctypes
    from simple_rf_b200 import _lib, nerf_program as NP
    assert _lib.load().srf_dgrad_program_bytes() == ctypes.sizeof(NP.DgradProgram)
    cfg, params, packed, o, d, vd, z = _setup(golden_configs, variant, R, S, seed=3)
    g = torch.Generator().manual_seed(9)
    g_sigma = torch.randn(R, S, 1, generator=g) * 0.1
    g_rgb = torch.randn(R, S, 3, generator=g)
    pts = (o[:, None] + d[:, None] * z[..., None]).reshape(-1, 3)
    vflat = vd[:, None].expand(R, S, 3).reshape(-1, 3)
    leaves = {k: v.clone().requires_grad_() for k, v in params.items()}
    ref = M.mlp_forward(leaves, cfg, pts, vflat if cfg['use_view_dirs'] else None, None)
    ((ref['sigma'] * g_sigma.reshape(-1, 1)).sum() + (ref['rgb'] * g_rgb.reshape(-1, 3)).sum()).backward()
    leaves_b = {k: v.clone().requires_grad_() for k, v in params.items()}
    sb, cb = _bf16_model_forward(leaves_b, cfg, pts, vflat)
    ((sb * g_sigma.reshape(-1, 1)).sum() + (cb * g_rgb.reshape(-1, 3)).sum()).backward()

    sigma, rgb, acts = packed.forward(o.to(DEV), d.to(DEV), z.to(DEV), vd.to(DEV), save=True)
    flat_grad, dz = NP.mlp_backward(packed, packed.flat, acts, sigma, rgb, g_sigma.to(DEV), g_rgb.to(DEV))
    torch.cuda.synchronize()
    err_model, err_fp32 = {}, {}
    off = 0
    for name in packed.param_names:
        n = leaves[name].numel()
        got = flat_grad[off:off + n].view(leaves[name].shape).cpu()
        off += n
        gb, gf = leaves_b[name].grad, leaves[name].grad
        err_model[name] = (got - gb).abs().max().item() / max(gb.abs().max().item(), 1e-12)
        err_fp32[name] = ((got - gf).norm() / gf.norm().clamp_min(1e-12)).item()
    print(variant, R, S, 'vs bf16-model (max-abs rel):', {k: round(v, 4) for k, v in err_model.items()})
    print(variant, R, S, 'vs fp32 (rel L2):', {k: round(v, 4) for k, v in err_fp32.items()})
    for name in packed.param_names:
        assert err_model[name] <= 5e-2, ('bf16 model', name, err_model[name])
        assert err_fp32[name] <= 0.15, ('fp32', name, err_fp32[name])


def test_rows_mlp_backward_with_device_count_matches_exact_size():
    """The TensoRF colour branch sizes every buffer for the worst case and passes the number of valid rows on the device
    (srf_mlp_rows_fwd / srf_nerf_mlp_dgrad / srf_nerf_mlp_wgrad `count`): gradients must equal those of an exactly sized call,
    also for an empty row set and for counts that end inside a tile."""
    from simple_rf_b200.nerf_program import PackedRowsMLP
    g = torch.Generator().manual_seed(3)
    m = PackedRowsMLP(72, 27, 3, prefix='mlp')
    lin = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5).to(DEV)
    params = {'mlp.0.weight': lin(128, 30), 'mlp.0.bias': lin(128, 1)[:, 0], 'mlp.2.weight': lin(128, 128), 'mlp.2.bias': lin(128, 1)[:, 0],
              'mlp.4.weight': lin(3, 128), 'mlp.4.bias': torch.zeros(3, device=DEV)}
    m.refresh(params, lin(27, 72))
    capacity = 128 * 37 + 5
    rows_all = (torch.randn(capacity, 80, generator=g) * 0.3).to(torch.bfloat16).to(DEV)
    g_all = torch.randn(capacity, 3, generator=g).to(DEV)
    for n in (0, 1, 128, 1000, 128 * 20 + 77, capacity):
        count = torch.tensor([n], dtype=torch.int32, device=DEV)
        rgb_c, acts_c = m.forward(rows_all, count, capacity, save=True)
        grads_c, g_rows_c = m.backward(acts_c, rgb_c, g_all, capacity, count=count)
        if n == 0:
            assert float(grads_c.abs().max()) == 0.0
            continue
        rgb_e, acts_e = m.forward(rows_all[:n].contiguous(), None, n, save=True)
        grads_e, g_rows_e = m.backward(acts_e, rgb_e, g_all[:n].contiguous(), n)
        assert torch.equal(rgb_c[:n], rgb_e)
        assert torch.equal(g_rows_c[:n], g_rows_e[:n])
        # the weight gradients are sums over tiles in a different CTA split: fp32 reassociation only
        assert (grads_c - grads_e).abs().max().item() <= 1e-5 * max(1.0, grads_e.abs().max().item()), n
