"""Oracle: the whole per-ray render of one chunk, Simple-NeRF and Simple-TensoRF (test infrastructure).

Stitches the stage oracles together in the reference's order, INCLUDING the order in which the
CPU default generator is consumed, so that with the same seed the
outputs equal the reference's `render_rays` (src/models/SimpleNeRF17.py:159-328,
src/models/SimpleTensoRF09.py:194-351) bit for bit on the same host.
"""
import torch

from . import composite as C
from . import nerf_mlp as M
from . import rays as RY
from . import sampling as SP
from . import tensorf as TF


def _run_mlp_chunked(params, cfg, pts, view_dirs, netchunk, noise_std, training):
    """SimpleNeRF17.py:419-484: flatten, expand view dirs, loop over `netchunk` points; the sigma
    noise is drawn per chunk as torch.randn([m,1]) * raw_noise_std (:739-741)."""
    R, S, _ = pts.shape
    flat = pts.reshape(-1, 3)
    vflat = None
    if cfg['use_view_dirs']:
        vflat = view_dirs[:, None].expand(pts.shape).reshape(-1, 3)
    outs = {}
    for i in range(0, flat.shape[0], netchunk):
        p = flat[i:i + netchunk]
        v = None if vflat is None else vflat[i:i + netchunk]
        noise = None
        if training and noise_std > 0.:
            noise = torch.randn([p.shape[0], 1]) * noise_std
        o = M.mlp_forward(params, cfg, p, v, noise)
        for k, t in o.items():
            outs.setdefault(k, []).append(t)
    return {k: torch.cat(v, 0).reshape(R, S, -1) for k, v in outs.items()}


def nerf_render_chunk(models, configs, model_configs, pixel_id, *, training, retraw=True, extrinsics=None):
    """models: {'coarse_model': params, 'fine_model': params,
                'augmentations': [(name, cfg, params), ...]}  (aug models are coarse-only, as shipped).
    Mirrors SimpleNeRF.render_rays for ndc=True/False, no visibility prediction.  `extrinsics` [V,4,4]: the (differentiable) view
    matrices of a learnable pose correction (RY.pose_correction) instead of the fixed ones of model_configs."""
    mc = configs['model']
    ndc = configs['data_loader']['ndc']
    K = torch.tensor(model_configs['intrinsics'], dtype=torch.float32)
    E = torch.tensor(model_configs['extrinsics'], dtype=torch.float32) if extrinsics is None else extrinsics
    h, w = model_configs['resolution']
    R = pixel_id.shape[0]
    out = {}
    rays_o, rays_d = RY.camera_rays(pixel_id, K, E, half_pixel=False, flip_x=False)
    out['rays_o'], out['rays_d'] = rays_o, rays_d
    img = pixel_id[:, 0].long()
    if ndc:
        o_ndc, d_ndc = RY.ndc_rays(rays_o, rays_d, h, w, K[img, 0, 0], K[img, 1, 1], model_configs['near'])
        out['rays_o_ndc'], out['rays_d_ndc'] = o_ndc, d_ndc
        near, far = model_configs['near_ndc'], model_configs['far_ndc']
        so, sd = o_ndc, d_ndc
    else:
        d_ndc = None
        near, far = model_configs['near'], model_configs['far']
        so, sd = rays_o, rays_d
    vd = RY.view_dirs(rays_d)
    out['view_dirs'] = vd
    perturb = training and mc['perturb']
    noise_std = mc['raw_noise_std']
    netchunk = mc['netchunk']

    def render(params, cfg, z, tag, prefix=''):
        pts = so[..., None, :] + sd[..., None, :] * z[..., :, None]
        raw = _run_mlp_chunked(params, cfg, pts, vd, netchunk, noise_std, training)
        vr = C.composite(raw['sigma'][..., 0], raw['rgb'], z, rays_o, rays_d, d_ndc, ndc=ndc,
                         white_bkgd=mc['white_bkgd'])
        for k, t in vr.items():
            out[f'{prefix}{k}_{tag}'] = t
        if retraw:
            for k, t in raw.items():
                out[f'{prefix}raw_{k}_{tag}'] = t
        return vr

    S_c = mc['coarse_model']['num_samples']
    ladder = SP.coarse_depths(S_c, near, far, mc['lindisp'])
    jitter = torch.rand([R, S_c]) if perturb else None
    z_c = SP.stratified_depths(ladder, R, jitter)
    out['z_vals_coarse'] = z_c
    vr_c = render(models['coarse_model'], mc['coarse_model'], z_c, 'coarse')
    if training:
        for name, cfg, params in models.get('augmentations', []):
            render(params, cfg, z_c, 'coarse', prefix=f'{name}_')
    if 'fine_model' in models:
        N_f = mc['fine_model']['num_samples']
        u = torch.rand([R, N_f]) if perturb else SP.det_u(R, N_f)
        z_f, samples, below, above = SP.fine_depths(z_c, vr_c['weights'], u)
        out['z_vals_fine'] = z_f
        out['_fine_u'], out['_fine_below'], out['_fine_above'], out['_fine_samples'] = u, below, above, samples
        render(models['fine_model'], mc['fine_model'], z_f, 'fine')
    return out


def tensorf_render_chunk(tensors, configs, model_configs, pixel_id, *, training, retraw=True, extrinsics=None):
    """tensors: {'coarse_model': dict(params=, bbox=, num_samples=, alpha_volume=None, alpha_bbox=None),
                 'augmentations': [(name, cfg, dict(...)), ...]}.  NDC path of SimpleTensoRF.render_rays
    (SimpleTensoRF09.py:194-296): half-pixel rays, x-flip, view dirs from the NDC direction, samples
    from the MAIN tensor's num_samples for every tensor, background coin per tensor in training."""
    mc = configs['model']
    K = torch.tensor(model_configs['intrinsics'], dtype=torch.float32)
    E = torch.tensor(model_configs['extrinsics'], dtype=torch.float32) if extrinsics is None else extrinsics     # as in nerf_render_chunk
    h, w = model_configs['resolution']
    R = pixel_id.shape[0]
    out = {}
    ndc = bool(configs['data_loader']['ndc'])
    rays_o, rays_d = RY.camera_rays(pixel_id, K, E, half_pixel=True, flip_x=True)
    main = tensors['coarse_model']
    S = main['num_samples']
    perturb = training and mc['perturb']
    if ndc:
        img = pixel_id[:, 0].long()
        o_ndc, d_ndc = RY.ndc_rays(rays_o, rays_d, h, w, K[img, 0, 0], K[img, 1, 1], model_configs['near'])
        vd = RY.view_dirs(d_ndc)
        out.update(rays_o=rays_o, rays_d=rays_d, rays_o_ndc=o_ndc, rays_d_ndc=d_ndc, view_dirs=vd)
        ladder = SP.coarse_depths(S, model_configs['near_ndc'], model_configs['far_ndc'], mc['lindisp'])
        jitter = torch.rand([R, S]) if perturb else None
        z = SP.stratified_depths(ladder, R, jitter)
        pts = o_ndc[..., None, :] + d_ndc[..., None, :] * z[..., :, None]
    else:
        # world-space box marching (SimpleTensoRF09.py:388-400): one jitter draw per ray, steps of the MAIN tensor's step_size
        d_ndc = None
        vd = RY.view_dirs(rays_d)
        out.update(rays_o=rays_o, rays_d=rays_d, view_dirs=vd)
        res = main['resolution'].long()
        step_size = torch.mean((main['bbox'][1] - main['bbox'][0]).float() / (res - 1)) * mc['coarse_model']['num_voxels_per_sample']
        jitter = torch.rand([R, 1]) if perturb else None
        z = SP.box_march_depths(rays_o, rays_d, main['bbox'], model_configs['near'], model_configs['far'], step_size, S, jitter)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    out['z_vals_coarse'] = z

    def run(t, cfg, prefix):
        white = mc['white_bkgd'] or bool(training and (torch.rand((1,)) < 0.5))
        views = vd if (cfg['use_view_dirs'] and cfg['view_dependent_color']) else None       # SimpleTensoRF09.py:732, :1267, :1414
        res = TF.tensor_forward(t['params'], t['bbox'], pts, z, rays_o, rays_d, d_ndc, views, ndc=ndc,
                                alpha_volume=t.get('alpha_volume'), alpha_bbox=t.get('alpha_bbox'),
                                distance_scale=cfg['distance_scale'],
                                weight_threshold=cfg['ray_marching_weight_threshold'], white_bkgd=white,
                                density_predictor=cfg['density_predictor'], density_offset=cfg['density_offset'])
        for k, v in res.items():
            if k.startswith('raw_') and not retraw:
                continue
            out[f'{prefix}{k}_coarse'] = v

    run(main, mc['coarse_model'], '')
    if training:
        for name, cfg, t in tensors.get('augmentations', []):
            run(t, cfg, f'{name}_')
    return out
