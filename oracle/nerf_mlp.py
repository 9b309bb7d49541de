"""Oracle: positional encoding and the NeRF MLP variants (test infrastructure).

Follows src/models/SimpleNeRF17.py:581-613 (PositionalEncoder), :616-667 (MLP.__init__ layer
shapes), :696-724 (MLP.forward), :726-755 (view-independent trunk, sigma head, noise) and
:757-785 (view-dependent head).  Parameters are passed as a flat dict with the reference's
state-dict names (`pts_linears.0.weight`, `feature_linear.bias`, ...).
"""
import torch
import torch.nn.functional as F


def positional_encoding(x, degree):
    """SimpleNeRF17.py:589-613 with include_input, log_sampling, [sin, cos]:
    [x | sin(2^0 x) | cos(2^0 x) | ... | sin(2^(L-1) x) | cos(2^(L-1) x)]  -> (.., 3*(2L+1))."""
    parts = [x]
    if degree > 0:
        for f in 2. ** torch.linspace(0., degree - 1, steps=degree):
            parts.append(torch.sin(x * f))
            parts.append(torch.cos(x * f))
    return torch.cat(parts, -1)


def variant_dims(cfg):
    """Layer input widths implied by an mlp config dict (SimpleNeRF17.py:630-636)."""
    full = (2 * cfg['points_positional_encoding_degree'] + 1) * 3
    views = (2 * cfg['views_positional_encoding_degree'] + 1) * 3 if cfg['use_view_dirs'] else 0
    pts_in = full
    if 'points_sigma_positional_encoding_degree' in cfg:
        pts_in = (2 * cfg['points_sigma_positional_encoding_degree'] + 1) * 3
        views += full - pts_in
    return pts_in, views


def mlp_forward(params, cfg, pts, view_dirs=None, noise=None, skips=(4,)):
    """pts [M,3], view_dirs [M,3] or None, noise [M,1] or None (the reference's
    torch.randn(sigma.shape) * raw_noise_std, SimpleNeRF17.py:739-741, already scaled).
    Returns dict(sigma [M,1], rgb [M,3], + rgb_view_(in)dependent)."""
    depth = cfg['points_net_depth']
    pts_in, _ = variant_dims(cfg)
    enc = positional_encoding(pts, cfg['points_positional_encoding_degree'])
    x_in = enc[:, :pts_in]
    h = x_in
    for i in range(depth):
        h = F.relu(F.linear(h, params[f'pts_linears.{i}.weight'], params[f'pts_linears.{i}.bias']))
        if i in skips:
            h = torch.cat([x_in, h], -1)
    head = F.linear(h, params['pts_output_linear.weight'], params['pts_output_linear.bias'])
    raw_sigma = head[..., 0:1]
    if noise is not None:
        raw_sigma = raw_sigma + noise
    out = {'sigma': F.relu(raw_sigma)}
    view_dep = cfg['view_dependent_rgb']
    if not view_dep:
        out['rgb_view_independent'] = torch.sigmoid(head[..., 1:4])
        out['rgb'] = out['rgb_view_independent']
        return out
    feat = F.linear(h, params['feature_linear.weight'], params['feature_linear.bias'])
    feat = torch.cat([feat, enc[:, pts_in:]], dim=1)
    enc_v = positional_encoding(view_dirs, cfg['views_positional_encoding_degree'])
    hv = torch.cat([feat, enc_v], -1)
    n_views = cfg['views_net_depth']
    for i in range(n_views):
        hv = F.relu(F.linear(hv, params[f'views_linears.{i}.weight'], params[f'views_linears.{i}.bias']))
    rgb = torch.sigmoid(F.linear(hv, params['views_output_linear.weight'], params['views_output_linear.bias'])[..., 0:3])
    out['rgb_view_dependent'] = rgb
    out['rgb'] = rgb
    return out


def init_mlp_params(cfg, generator=None, width=None):
    """Default nn.Linear initialisation (kaiming-uniform a=sqrt(5) == U(-1/sqrt(in), 1/sqrt(in)))
    for the layer shapes of SimpleNeRF17.py:644-666; used for synthetic, weight-independent benches."""
    wp = cfg['points_net_width'] if width is None else width
    wv = cfg['views_net_width']
    pts_in, views_in = variant_dims(cfg)
    shapes = {}
    for i in range(cfg['points_net_depth']):
        fan_in = pts_in if i == 0 else (wp + pts_in if (i - 1) in (4,) else wp)
        shapes[f'pts_linears.{i}'] = (wp, fan_in)
    view_dep = cfg['view_dependent_rgb']
    shapes['pts_output_linear'] = (1 if view_dep else 4, wp)
    if view_dep:
        shapes['feature_linear'] = (wp, wp)
        for i in range(cfg['views_net_depth']):
            shapes[f'views_linears.{i}'] = (wv, views_in + wp if i == 0 else wv)
        shapes['views_output_linear'] = (3, wv)
    params = {}
    for name, (o, i) in shapes.items():
        bound = 1.0 / (i ** 0.5)
        params[f'{name}.weight'] = (torch.rand(o, i, generator=generator) * 2 - 1) * bound
        params[f'{name}.bias'] = (torch.rand(o, generator=generator) * 2 - 1) * bound
    return params
