"""Oracle: Simple-TensoRF vector-matrix (VM) and CANDECOMP/PARAFAC (CP) tensor evaluation (test infrastructure).

Follows src/models/SimpleTensoRF09.py:701-761 (LowRankTensor.forward), :763-765
(normalize_points), :1214-1239 (VmDecomposedTensor.get_volume_density), :1241-1272
(get_color), :1342-1349 (AlphaGridMask.sample_alpha / normalize_points) and :1370-1421
(MlpFeaturesColorPredictor, PE degree 0 == identity); for the CP tensor :1043-1062 (CpDecomposedTensor.get_volume_density)
and :1064-1089 (get_color).  A parameter dict without `matrices_*` entries is a CP tensor.  Parameters are passed as a dict with
the reference's state-dict names (`matrices_density.0`, `vectors_color.2`,
`basis_matrix_color.weight`, `color_predictor.mlp.0.weight`, ...).
"""
import torch
import torch.nn.functional as F

from .composite import composite

MATRIX_AXES = [[0, 1], [0, 2], [1, 2]]     # SimpleTensoRF09.py:1131
VECTOR_AXES = [2, 1, 0]                    # SimpleTensoRF09.py:1132


def bbox_mask(pts, bbox):
    """SimpleTensoRF09.py:705."""
    return ((bbox[0] <= pts) & (pts <= bbox[1])).all(dim=-1)


def normalize(pts, bbox):
    """SimpleTensoRF09.py:763-765 / :1347-1349."""
    return ((pts - bbox[0]) / (bbox[1] - bbox[0])) * 2 - 1


def sample_alpha(alpha_volume, alpha_bbox, pts):
    """SimpleTensoRF09.py:1342-1345: trilinear grid_sample(align_corners=True) of the
    [1,1,Z,Y,X] float {0,1} occupancy volume at pts [N,3]."""
    pn = normalize(pts, alpha_bbox)
    return F.grid_sample(alpha_volume, pn.view(1, -1, 1, 1, 3), align_corners=True).view(-1)


def validity_mask(pts, bbox, alpha_volume=None, alpha_bbox=None):
    """SimpleTensoRF09.py:705-710: in-box test, then AND with occupancy > 0 on the in-box points."""
    mask = bbox_mask(pts, bbox)
    if alpha_volume is not None:
        occ = sample_alpha(alpha_volume, alpha_bbox, pts[mask]) > 0
        mask = mask.clone()
        mask[mask.clone()] &= occ
    return mask


def _plane_line_coords(p):
    """Grid coordinates are DETACHED (SimpleTensoRF09.py:1224-1228, :1251-1253; CP: :1054, :1075): with learnable cameras the sample points
    carry no gradient into the tensors' interpolation."""
    plane = torch.stack([p[..., MATRIX_AXES[0]], p[..., MATRIX_AXES[1]], p[..., MATRIX_AXES[2]]]).detach().view(3, -1, 1, 2)
    line = torch.stack([p[..., VECTOR_AXES[0]], p[..., VECTOR_AXES[1]], p[..., VECTOR_AXES[2]]])
    line = torch.stack([torch.zeros_like(line), line], dim=-1).detach().view(3, -1, 1, 2)
    return plane, line


def vm_density(params, pts_norm, mask, density_predictor='ReLU', density_offset=-10.0):
    """SimpleTensoRF09.py:1214-1239.  pts_norm [R,S,3] in [-1,1]; returns sigma [R,S,1]."""
    sigma = torch.zeros([*pts_norm.shape[:-1], 1], dtype=pts_norm.dtype)
    p = pts_norm[mask]
    if p.any():
        cp, cl = _plane_line_coords(p)
        feat = torch.zeros((p.shape[0],), dtype=p.dtype)
        for i in range(3):
            pc = F.grid_sample(params[f'matrices_density.{i}'], cp[[i]], align_corners=True).view(-1, p.shape[0])
            lc = F.grid_sample(params[f'vectors_density.{i}'], cl[[i]], align_corners=True).view(-1, p.shape[0])
            feat = feat + torch.sum(pc * lc, dim=0)
        if density_predictor == 'ReLU':
            val = F.relu(feat)
        else:
            val = F.softplus(feat + density_offset)
        sigma[mask] = val[..., None]
    return sigma


def is_cp(params):
    return 'matrices_density.0' not in params


def cp_products(params, kind, p):
    """SimpleTensoRF09.py:1053-1058 / :1074-1079: product of the three 1-D grid_samples, [C, N]."""
    _, cl = _plane_line_coords(p)
    out = F.grid_sample(params[f'vectors_{kind}.0'], cl[[0]], align_corners=True).view(-1, p.shape[0])
    out = out * F.grid_sample(params[f'vectors_{kind}.1'], cl[[1]], align_corners=True).view(-1, p.shape[0])
    out = out * F.grid_sample(params[f'vectors_{kind}.2'], cl[[2]], align_corners=True).view(-1, p.shape[0])
    return out


def cp_density(params, pts_norm, mask, density_predictor='ReLU', density_offset=-10.0):
    """SimpleTensoRF09.py:1043-1062."""
    sigma = torch.zeros([*pts_norm.shape[:-1], 1], dtype=pts_norm.dtype)
    p = pts_norm[mask]
    if p.any():
        feat = torch.sum(cp_products(params, 'density', p), dim=0)
        val = F.relu(feat) if density_predictor == 'ReLU' else F.softplus(feat + density_offset)
        sigma[mask] = val[..., None]
    return sigma


def density(params, pts_norm, mask, density_predictor='ReLU', density_offset=-10.0):
    fn = cp_density if is_cp(params) else vm_density
    return fn(params, pts_norm, mask, density_predictor, density_offset)


def color_products(params, p):
    """The input of basis_matrix_color, [N, in_features]."""
    return cp_products(params, 'color', p).T if is_cp(params) else vm_color_products(params, p)


def vm_color_products(params, p):
    """SimpleTensoRF09.py:1252-1262: (plane*line)^T [N, sum(C)], the input of basis_matrix_color."""
    cp, cl = _plane_line_coords(p)
    pcs, lcs = [], []
    for i in range(3):
        pcs.append(F.grid_sample(params[f'matrices_color.{i}'], cp[[i]], align_corners=True).view(-1, p.shape[0]))
        lcs.append(F.grid_sample(params[f'vectors_color.{i}'], cl[[i]], align_corners=True).view(-1, p.shape[0]))
    return (torch.cat(pcs) * torch.cat(lcs)).T


def vm_color_features(params, p):
    """SimpleTensoRF09.py:1252-1263: products -> basis matrix -> [N, 27]."""
    return F.linear(color_products(params, p), params['basis_matrix_color.weight'])


def color_mlp(params, features, view_dirs):
    """SimpleTensoRF09.py:1389-1393, :1411-1421 with both PE degrees 0 (identity encodings); view_dirs None: `use_view_dirs` /
    `view_dependent_color` false (:1384, :1414)."""
    x = features if view_dirs is None else torch.cat([features, view_dirs], dim=-1)
    x = F.relu(F.linear(x, params['color_predictor.mlp.0.weight'], params['color_predictor.mlp.0.bias']))
    x = F.relu(F.linear(x, params['color_predictor.mlp.2.weight'], params['color_predictor.mlp.2.bias']))
    return torch.sigmoid(F.linear(x, params['color_predictor.mlp.4.weight'], params['color_predictor.mlp.4.bias']))


def vm_color(params, pts_norm, mask, view_dirs):
    """SimpleTensoRF09.py:1241-1272.  view_dirs [R,3] (expanded per sample, :733-736) or None."""
    rgb = torch.zeros([*pts_norm.shape[:-1], 3], dtype=pts_norm.dtype)
    if mask.any():
        p = pts_norm[mask]
        vd = None if view_dirs is None else view_dirs[:, None].expand(pts_norm.shape)[mask]
        rgb[mask] = color_mlp(params, vm_color_features(params, p), vd)
    return rgb


def tensor_forward(params, bbox, pts, z, rays_o, rays_d, rays_d_ndc, view_dirs, *, ndc=True,
                   alpha_volume=None, alpha_bbox=None, distance_scale=25.0, weight_threshold=1e-4,
                   white_bkgd=False, density_predictor='ReLU', density_offset=-10.0):
    """SimpleTensoRF09.py:701-761 for one VM or CP tensor: mask -> density -> weights -> surface mask
    -> colour -> composite.  `white_bkgd` folds in the training-time coin of :746."""
    mask = validity_mask(pts, bbox, alpha_volume, alpha_bbox)
    pn = normalize(pts, bbox)
    sigma = density(params, pn, mask, density_predictor, density_offset)
    vr = composite(sigma[..., 0], None, z, rays_o, rays_d, rays_d_ndc, ndc=ndc, distance_scale=distance_scale)
    surface = vr['weights'] > weight_threshold
    rgb = vm_color(params, pn, surface, view_dirs)
    rgb_map = torch.sum(vr['weights'][..., None] * rgb, dim=-2)
    if white_bkgd:
        rgb_map = rgb_map + (1 - vr['acc'][..., None])
    out = dict(vr)
    out['rgb'] = rgb_map
    out['raw_sigma'] = sigma
    out['raw_rgb'] = rgb
    out['validity_mask'] = mask
    out['surface_mask'] = surface
    return out


def vm_resolution(num_voxels, bbox):
    """SimpleTensoRF09.py:626-638 (cubical voxels)."""
    size = bbox[1] - bbox[0]
    voxel = (size.prod() / num_voxels).pow(1 / 3)
    return (size / voxel).long()


def vm_num_samples(resolution, voxels_per_sample=0.5, num_samples_max=1e6):
    """SimpleTensoRF09.py:640-650."""
    n = (torch.linalg.norm(resolution.float()) / voxels_per_sample).round().long()
    return int(min(num_samples_max, n))


def init_vm_params(resolution, comps_density, comps_color, feat_dim=27, units=128, generator=None, scale=0.1, use_views=True):
    """Shapes of SimpleTensoRF09.py:1154-1165 + :1151 + :1389-1393 (0.1*randn planes/lines)."""
    res = [int(r) for r in resolution]
    p = {}
    for kind, comps in (('density', comps_density), ('color', comps_color)):
        for i in range(3):
            a0, a1 = MATRIX_AXES[i]
            p[f'matrices_{kind}.{i}'] = scale * torch.randn(1, comps[i], res[a1], res[a0], generator=generator)
            p[f'vectors_{kind}.{i}'] = scale * torch.randn(1, comps[i], res[VECTOR_AXES[i]], 1, generator=generator)

    _init_color_network(p, sum(comps_color), feat_dim, units, generator, use_views)
    return p


def _init_color_network(p, in_features, feat_dim, units, generator, use_views=True):
    """basis_matrix_color (:1151 / :986) + MlpFeaturesColorPredictor (:1389-1393), torch.nn.Linear-style uniform init."""
    def lin(o, i, bias=True):
        b = 1.0 / i ** 0.5
        w = (torch.rand(o, i, generator=generator) * 2 - 1) * b
        return w, ((torch.rand(o, generator=generator) * 2 - 1) * b if bias else None)
    p['basis_matrix_color.weight'], _ = lin(feat_dim, in_features, bias=False)
    p['color_predictor.mlp.0.weight'], p['color_predictor.mlp.0.bias'] = lin(units, feat_dim + (3 if use_views else 0))
    p['color_predictor.mlp.2.weight'], p['color_predictor.mlp.2.bias'] = lin(units, units)
    p['color_predictor.mlp.4.weight'], _ = lin(3, units)
    p['color_predictor.mlp.4.bias'] = torch.zeros(3)


def init_cp_params(resolution, comps_density, comps_color, feat_dim=27, units=128, generator=None, scale=0.1, use_views=True):
    """Shapes of SimpleTensoRF09.py:979-996 (every line holds num_components[0] components) + the shared colour predictor."""
    res = [int(r) for r in resolution]
    p = {}
    for kind, comps in (('density', comps_density), ('color', comps_color)):
        for i in range(3):
            p[f'vectors_{kind}.{i}'] = scale * torch.randn(1, comps[0], res[VECTOR_AXES[i]], 1, generator=generator)
    _init_color_network(p, sum(comps_color), feat_dim, units, generator, use_views)
    return p
