"""Drop-in Simple-NeRF model class backed by the sm_100a kernels.

Resolvable by the reference's unmodified factory (src/models/ModelFactory02.py:10-22): module
`SimpleNeRF91` -> class `SimpleNeRF` (module name minus its two-digit suffix).  Mirrors the public
surface of src/models/SimpleNeRF17.py: constructor, `forward(input_batch, *, retraw, sec_views_vis, mode)`
dict contract (:104-127, :159-328), `get_trainable_parameters` (:77-90),
`rebuild_camera_params_learners` (:92-102), `augmented_models`, and state-dict compatible parameter
names (`coarse_model.pts_linears.0.weight`, `augmented_models_nn.0.coarse_model...`), so reference
checkpoints interchange.  All per-ray arithmetic runs in hand-written CUDA kernels through the C ABI
(simple_rf_b200/ops.py, nerf_program.py); there is no eager fallback.

Random numbers: in `rng_mode='reference'` (default) the stratified jitter, the sigma noise and the
inverse-CDF draws are taken from the CPU default generator in the reference's order
and uploaded, so seeded runs reproduce the reference's samples; `rng_mode='device'` draws them on the
GPU (torch's CUDA generator: no host->device copies, capturable in a CUDA graph, see train_graph.py) and is not
seed-compatible.
"""
import numpy
import torch
from torch.nn import ModuleDict, ModuleList

from .. import _lib as L
from .. import ops
from ..nerf_program import PackedMLP


class SimpleNeRF(torch.nn.Module):
    def __init__(self, configs: dict, model_configs: dict):
        super().__init__()
        self.configs = configs
        self.model_configs = model_configs
        self.ndc = self.configs['data_loader']['ndc']
        self.coarse_model_needed = 'coarse_model' in self.configs['model']
        self.fine_model_needed = 'fine_model' in self.configs['model']
        self.predict_visibility = (self.coarse_model_needed and self.configs['model']['coarse_model']['predict_visibility']) or \
                                  (self.fine_model_needed and self.configs['model']['fine_model']['predict_visibility'])
        if self.predict_visibility:
            raise NotImplementedError('predict_visibility is dead in every shipped config (SURVEY.md App. C11)')
        self.rng_mode = self.configs['model'].get('rng_mode', 'reference')
        self.eval_chunk = int(self.configs['model'].get('eval_chunk', 1 << 16))

        self.coarse_model = None
        self.fine_model = None
        self.augmentations_needed = 'augmentations' in configs['model']
        if self.augmentations_needed:
            self.augmented_models = []
            self.augmented_models_nn = []
        self._camera_tables = None
        self.build_nerf()
        self.optimizers = None          # Trainer10.py:61-62 hands the optimiser dict over through this attribute

    def __setattr__(self, name, value):
        super().__setattr__(name, value)
        if name == 'optimizers' and value is not None:
            # optimiser tail: Adam optimisers step through one fused kernel over flat buffers, with ONE
            # all-reduce of the flat gradient bucket on several ranks; other optimisers get the generic all-reduce hook
            from .. import optim, parallel
            super().__setattr__('_fused_adam', optim.attach(value) if optim.enabled() else [])
            super().__setattr__('_grad_allreduce', parallel.attach_gradient_allreduce(value))

    # ------------------------------------------------------------------ construction (SimpleNeRF17.py:36-75)
    def build_nerf(self):
        mc = self.configs['model']
        if self.coarse_model_needed:
            self.coarse_model = MLP('coarse_model', self.configs, mc['coarse_model'], self.model_configs)
        if self.fine_model_needed:
            self.fine_model = MLP('fine_model', self.configs, mc['fine_model'], self.model_configs)
        self.intrinsics_learner = IntrinsicsLearner(numpy.array(self.model_configs['intrinsics']),
                                                    learn_focal=mc['learn_camera_focal_length'])
        self.extrinsics_learner = ExtrinsicsLearner(numpy.array(self.model_configs['extrinsics']),
                                                    learn_rotation=mc['learn_camera_rotation'],
                                                    learn_translation=mc['learn_camera_translation'])
        if self.augmentations_needed:
            for aug in mc['augmentations']:
                nn_dict = ModuleDict()
                entry = {'name': aug['name'], 'coarse_model': None, 'fine_model': None}
                for tag in ('coarse_model', 'fine_model'):
                    if tag in aug:
                        entry[tag] = MLP(f"{aug['name']}_{tag}", self.configs, aug[tag], self.model_configs)
                        nn_dict[tag] = entry[tag]
                self.augmented_models.append(entry)
                self.augmented_models_nn.append(nn_dict)
            self.augmented_models_nn = ModuleList(self.augmented_models_nn)

    def get_trainable_parameters(self, optimizer_configs):
        groups = []
        for m in (self.coarse_model, self.fine_model):
            if m is not None:
                groups.extend(m.get_trainable_parameters(optimizer_configs))
        if self.augmentations_needed:
            for aug in self.augmented_models:
                for tag in ('coarse_model', 'fine_model'):
                    if aug[tag] is not None:
                        groups.extend(aug[tag].get_trainable_parameters(optimizer_configs))
        return groups

    def rebuild_camera_params_learners(self, *, intrinsics: numpy.ndarray = None, extrinsics=None, device):
        mc = self.configs['model']
        if intrinsics is not None:
            self.intrinsics_learner = IntrinsicsLearner(intrinsics, learn_focal=mc['learn_camera_focal_length']).to(device)
        if extrinsics is not None:
            self.extrinsics_learner = ExtrinsicsLearner(extrinsics, learn_rotation=mc['learn_camera_rotation'],
                                                        learn_translation=mc['learn_camera_translation']).to(device)
        self._camera_tables = None

    # ------------------------------------------------------------------ forward (SimpleNeRF17.py:104-127)
    def forward(self, input_batch: dict, *, retraw: bool = False, sec_views_vis: bool = False, mode: str = None):
        pixel_id = input_batch['pixel_id']
        L.require_cuda(pixel_id)
        image_id = pixel_id[:, 0].long()
        intrinsics = self.intrinsics_learner(image_id)
        extrinsics = self.extrinsics_learner(image_id)
        all_extrinsics = self.extrinsics_learner(torch.arange(input_batch['num_frames'], device=pixel_id.device))
        if mode == 'camera_params_only':
            out = {}
        else:
            out = self.render(input_batch, retraw=retraw or self.training, mode=mode)
        out['intrinsics'] = intrinsics
        out['extrinsics'] = extrinsics
        out['extrinsics_all'] = all_extrinsics
        return out

    def _tables(self, device, intrinsics=None, extrinsics=None):
        """Per-view K^-1 / c2w / focal tables for the raygen kernel (cameras are frozen in every shipped config; with learnable
        cameras and gradients enabled `_rays` attaches the rays to the learner's autograd graph instead, camera_grad.py)."""
        if intrinsics is not None:
            return ops.camera_tables(intrinsics, extrinsics, device)
        # keyed on the camera tensors' storage + version: load_state_dict() copies new cameras in place
        cams = (self.intrinsics_learner.initial_intrinsics, self.extrinsics_learner.initial_extrinsics,
                self.extrinsics_learner.r, self.extrinsics_learner.t)
        key = (str(device),) + tuple((c.data_ptr(), c._version) for c in cams)
        if self._camera_tables is None or self._camera_tables[0] != key:
            self._camera_tables = (key, ops.camera_tables(self.intrinsics_learner.initial_intrinsics,
                                                          self.extrinsics_learner.view_matrices(), device))
        return self._camera_tables[1]

    def _rays(self, pixel_id, h, w):
        """get_rays_tr + get_ndc_rays_tr + get_view_dirs_tr (SimpleNeRF17.py:170-190) in one launch.  Learnable cameras under autograd:
        same kernel, same values, attached to the graph of ExtrinsicsLearner.forward (:817-842) so that r / t receive gradients."""
        flags = dict(half_pixel=False, flip_x=False, ndc=self.ndc, viewdirs_from_ndc=False)
        learner = self.extrinsics_learner
        if torch.is_grad_enabled() and (learner.r.requires_grad or learner.t.requires_grad):
            from .. import camera_grad
            dev = pixel_id.device
            k = self.intrinsics_learner.initial_intrinsics
            key = (str(dev), k.data_ptr(), k._version)
            if getattr(self, '_intrinsic_tables', None) is None or self._intrinsic_tables[0] != key:
                k_inv, _, focal = ops.camera_tables(k, torch.eye(4)[None].expand(k.shape[0], 4, 4), dev)
                self._intrinsic_tables = (key, k_inv, focal)
            views = learner(torch.arange(learner.num_frames, device=dev))          # [V, 4, 4], differentiable w.r.t. r and t
            return camera_grad.rays_with_camera_gradient(views, pixel_id, k, self._intrinsic_tables[1], self._intrinsic_tables[2], h, w,
                                                         self.model_configs['near'], **flags)
        return ops.raygen(pixel_id, self._tables(pixel_id.device), h, w, self.model_configs['near'], **flags)

    def render(self, input_dict: dict, *, retraw: bool, mode: str):
        """batchify_rays (SimpleNeRF17.py:133-157): chunks of `chunk` rays in training (keeps the
        reference's random-number order), of `eval_chunk` rays otherwise (results do not depend on it)."""
        from .. import parallel
        pixel_id = input_dict['pixel_id']
        num_rays = pixel_id.shape[0]
        chunk = self.configs['model']['chunk'] if self.training else max(self.eval_chunk, self.configs['model']['chunk'])
        # test time on several ranks (one process per GPU): every rank receives the full frame from create_test_data
        # (DataPreprocessor10.py:736-743), renders its row band and all-gathers the per-ray maps, so Tester07 needs no edit
        band = None if self.training else parallel.eval_band(num_rays, self.configs['model'])
        if band is not None:
            pixel_id = pixel_id[band[0]:band[1]]
        parts = []
        for i in range(0, pixel_id.shape[0], chunk):
            parts.append(self.render_rays(pixel_id[i:i + chunk], input_dict, retraw=retraw, mode=mode))
        out = parts[0] if len(parts) == 1 else {k: torch.cat([p[k] for p in parts], dim=0) for k in parts[0]}
        if band is not None:
            out = parallel.gather_ray_outputs(out, num_rays)
        return out

    def render_rays(self, pixel_id, input_dict, *, retraw, mode):
        mc = self.configs['model']
        dev = pixel_id.device
        R = pixel_id.shape[0]
        h, w = self.model_configs['resolution']
        out = {}
        rays_o, rays_d, o_ndc, d_ndc, view_dirs = self._rays(pixel_id, h, w)
        out['rays_o'], out['rays_d'] = rays_o, rays_d
        if self.ndc:
            out['rays_o_ndc'], out['rays_d_ndc'] = o_ndc, d_ndc
            near, far = self.model_configs['near_ndc'], self.model_configs['far_ndc']
            so, sd = o_ndc, d_ndc
        else:
            near, far = self.model_configs['near'], self.model_configs['far']
            so, sd = rays_o, rays_d
        uses_views = any(m is not None and m.mlp_configs['use_view_dirs'] for m in (self.coarse_model, self.fine_model))
        if uses_views:
            if mode == 'static_camera':
                cd = input_dict['common_data']
                k_view = cd['view_intrinsic'] if cd.get('view_intrinsic') is not None else None
                pose = cd['processed_view_pose']
                if pose.dim() == 3:
                    pose = pose[0]
                if k_view is not None and k_view.dim() == 3:
                    k_view = k_view[0]
                nviews = self.intrinsics_learner.initial_intrinsics.shape[0]
                ks = self.intrinsics_learner.initial_intrinsics.detach() if k_view is None else k_view[None].expand(nviews, 3, 3)
                tabs = self._tables(dev, ks, pose[None].expand(nviews, 4, 4))
                view_dirs = ops.raygen(pixel_id, tabs, h, w, self.model_configs['near'], half_pixel=False, flip_x=False,
                                       ndc=False, viewdirs_from_ndc=False)[4]
            out['view_dirs'] = view_dirs
        perturb = self.training and mc['perturb']
        from .. import parallel
        shard = input_dict.get('srf_shard') if self.training else None

        def cpu_draw(draw):                      # the reference's CPU draw for R rays (this rank's rows of it on several ranks)
            return parallel.rows_of_global_draw(draw, R, shard, mc['chunk']).to(dev)

        if self.rng_mode != 'reference':
            parallel.decorrelate_device_rng()
        aug_active = self.augmentations_needed and self.training and (mode != 'test_camera_params_optimization')

        def run(model, z, tag, prefix=''):
            noise = model.draw_noise(R, z.shape[1], dev, self.training, self.rng_mode, cpu_draw)
            sigma, rgb = model.evaluate(so, sd, z, view_dirs, noise)
            vr = ops.composite(sigma[..., 0], rgb, z, rays_o, rays_d, d_ndc, ndc=self.ndc,
                               white_bkgd=mc['white_bkgd'], per_sample=retraw)
            for k, v in vr.items():
                out[f'{prefix}{k}_{tag}'] = v
            if retraw:
                out[f'{prefix}raw_sigma_{tag}'] = sigma
                out[f'{prefix}raw_rgb_{tag}'] = rgb
                if model.view_dep_rgb:
                    out[f'{prefix}raw_rgb_view_dependent_{tag}'] = rgb
                else:
                    out[f'{prefix}raw_rgb_view_independent_{tag}'] = rgb
            return vr

        z_coarse = weights_coarse = None
        if self.coarse_model_needed:
            S = mc['coarse_model']['num_samples']
            ladder = coarse_ladder_on(dev, S, near, far, mc['lindisp'])
            if perturb and self.rng_mode == 'reference':
                z_coarse = ops.stratified_z(ladder, R, jitter=cpu_draw(lambda n: torch.rand([n, S])))       # SimpleNeRF17.py:355
            elif perturb:
                z_coarse = ops.stratified_z(ladder, R, jitter=torch.rand([R, S], device=dev))     # torch's CUDA generator: CUDA-graph safe
            else:
                z_coarse = ops.stratified_z(ladder, R)
            out['z_vals_coarse'] = z_coarse
            weights_coarse = run(self.coarse_model, z_coarse, 'coarse')['weights']
            if aug_active:
                for aug in self.augmented_models:
                    if aug['coarse_model'] is not None:
                        run(aug['coarse_model'], z_coarse, 'coarse', prefix=f"{aug['name']}_")
        if self.fine_model_needed:
            N = mc['fine_model']['num_samples']
            w_det = weights_coarse.detach()
            if perturb and self.rng_mode == 'reference':
                z_fine = ops.sample_pdf_merge(z_coarse, w_det, N, u=cpu_draw(lambda n: torch.rand([n, N])))   # SimpleNeRF17.py:397
            elif perturb:
                z_fine = ops.sample_pdf_merge(z_coarse, w_det, N, u=torch.rand([R, N], device=dev))
            else:
                z_fine = ops.sample_pdf_merge(z_coarse, w_det, N, u=_on_device('linspace', int(N), dev, lambda: torch.linspace(0., 1., steps=N)))
            out['z_vals_fine'] = z_fine
            run(self.fine_model, z_fine, 'fine')
            if aug_active:
                for aug in self.augmented_models:
                    if aug['fine_model'] is not None:
                        run(aug['fine_model'], z_fine, 'fine', prefix=f"{aug['name']}_")
        if not retraw:                                                                  # SimpleNeRF17.py:315-326
            for k in [k for k in out if k.startswith('z_vals_') or '_alpha_' in f'_{k}' or '_visibility_' in f'_{k}'
                      or '_weights_' in f'_{k}']:
                del out[k]
        return out


def _coarse_ladder(num_samples, near, far, lindisp):
    """SimpleNeRF17.py:341-345, evaluated with the same CPU torch ops so the ladder is bit-identical."""
    t = torch.linspace(0., 1., steps=num_samples)
    if not lindisp:
        return near * (1. - t) + far * t
    return 1. / (1. / near * (1. - t) + 1. / far * t)


_DEVICE_CONSTANTS = {}


def _on_device(kind, key, device, make):
    """Small per-call constants (depth ladders, the deterministic u of sample_pdf) are built once on the host with the
    reference's CPU ops and cached per device: a pageable host->device copy in every forward() synchronises the stream and
    drains the launch queue."""
    k = (kind, key, str(device))
    t = _DEVICE_CONSTANTS.get(k)
    if t is None:
        t = _DEVICE_CONSTANTS[k] = make().to(device)
    return t


def coarse_ladder_on(device, num_samples, near, far, lindisp):
    return _on_device('ladder', (int(num_samples), float(near), float(far), bool(lindisp)), device,
                      lambda: _coarse_ladder(num_samples, near, far, lindisp))


class MLP(torch.nn.Module):
    """Parameter container with the reference's layer names and shapes (SimpleNeRF17.py:616-667); evaluation
    goes through the fused tcgen05 kernel on a packed-weight cache that is rebuilt when parameters change."""

    def __init__(self, name, configs, mlp_configs, model_configs):
        super().__init__()
        self.name = name
        self.configs = configs
        self.mlp_configs = mlp_configs
        self.model_configs = model_configs
        self.Dp, self.Dv = mlp_configs['points_net_depth'], mlp_configs['views_net_depth']
        self.Wp, self.Wv = mlp_configs['points_net_width'], mlp_configs['views_net_width']
        full = (2 * mlp_configs['points_positional_encoding_degree'] + 1) * 3
        self.pts_input_dim = full
        self.views_input_dim = 0
        if mlp_configs['use_view_dirs']:
            self.views_input_dim = (2 * mlp_configs['views_positional_encoding_degree'] + 1) * 3
        if 'points_sigma_positional_encoding_degree' in mlp_configs:
            self.pts_input_dim = (2 * mlp_configs['points_sigma_positional_encoding_degree'] + 1) * 3
            self.views_input_dim += full - self.pts_input_dim
        self.skips = [4]
        self.view_dep_rgb = mlp_configs['view_dependent_rgb']
        self.predict_visibility = mlp_configs['predict_visibility']
        self.view_dep_outputs = self.view_dep_rgb or self.predict_visibility
        self.raw_noise_std = configs['model']['raw_noise_std']
        Lin = torch.nn.Linear
        self.pts_linears = ModuleList(
            [Lin(self.pts_input_dim, self.Wp)] +
            [Lin(self.Wp, self.Wp) if i not in self.skips else Lin(self.Wp + self.pts_input_dim, self.Wp)
             for i in range(self.Dp - 1)])
        if self.view_dep_outputs:
            self.views_linears = ModuleList([Lin(self.views_input_dim + self.Wp, self.Wv)] +
                                            [Lin(self.Wv, self.Wv) for _ in range(self.Dv - 1)])
        self.pts_output_linear = Lin(self.Wp, 1 if self.view_dep_rgb else 4)
        if self.view_dep_outputs:
            self.feature_linear = Lin(self.Wp, self.Wp)
            self.views_output_linear = Lin(self.Wv, 3)
        self._packed = PackedMLP(mlp_configs)
        self._packed_version = None

    def get_trainable_parameters(self, optimizer_configs):
        params = torch.nn.ParameterList()
        params.extend(self.parameters())
        return [{'name': f'{self.name}_network_params', 'params': params, 'lr': optimizer_configs['lr_initial']}]

    def named_param_dict(self):
        return dict(self.named_parameters())

    def packed(self, params=None):
        params = self.named_param_dict() if params is None else params
        version = tuple((p.data_ptr(), p._version) for p in params.values())
        if version != self._packed_version:
            self._packed.refresh(params)
            self._packed_version = version
        return self._packed

    def draw_noise(self, num_rays, num_samples, device, training, rng_mode, cpu_draw=None):
        """sigma pre-activation noise (SimpleNeRF17.py:739-741).  Reference mode draws torch.randn on the CPU
        generator per `netchunk` points, exactly as the reference's chunk loop does (:460, :740); `cpu_draw` (several
        ranks) hands this rank its rows of the global draw."""
        if not (training and self.raw_noise_std > 0.):
            return None
        if rng_mode != 'reference':
            return torch.randn(num_rays * num_samples, device=device) * self.raw_noise_std

        def draw(n):
            count = n * num_samples
            netchunk = self.configs['model']['netchunk'] or count
            parts = [torch.randn([min(netchunk, count - i), 1]) * self.raw_noise_std for i in range(0, count, netchunk)]
            return torch.cat(parts, 0).reshape(n, num_samples)
        noise = cpu_draw(draw) if cpu_draw is not None else draw(num_rays).to(device)
        return noise.reshape(-1)

    def evaluate(self, rays_o, rays_d, z, view_dirs, noise):
        """-> sigma [R,S,1], rgb [R,S,3]; differentiable w.r.t. the parameters when grad is enabled.
        `configs['model']['mlp_precision']`: 'bf16' (default: bf16 operands, fp32 accumulation; stated looser tolerance) or
        'bf16x3' (split-bf16 operands: fp32-contract accuracy at three MMAs per K block; gradient-free evaluation only —
        a differentiable call keeps the bf16 program, whose saved tiles the backward kernels read)."""
        params = self.named_param_dict()        # ONE walk of the module tree per call (it costs ~0.1 ms of host time)
        packed = self.packed(params)
        rays_grad = any(t is not None and t.requires_grad for t in (rays_o, rays_d, view_dirs))        # learnable cameras
        if torch.is_grad_enabled() and (rays_grad or any(p.requires_grad for p in params.values())):
            return _FusedMLP.apply(packed, rays_o, rays_d, z, view_dirs, noise, *[params[n] for n in packed.param_names])
        split = self.configs['model'].get('mlp_precision', 'bf16') == 'bf16x3'
        return packed.forward(rays_o, rays_d, z, view_dirs if packed.use_views else None, noise, split=split)


class _FusedMLP(torch.autograd.Function):
    """Forward: the fused tcgen05 kernel, also saving the activation tiles.  Backward: the hand-written
    dgrad chain (srf_nerf_mlp_dgrad) and weight-gradient GEMMs (srf_nerf_mlp_wgrad) on the tensor cores."""

    @staticmethod
    def forward(ctx, packed, rays_o, rays_d, z, view_dirs, noise, *params):
        sigma, rgb, acts = packed.forward(rays_o, rays_d, z, view_dirs if packed.use_views else None, noise, save=True)
        ctx.packed = packed
        ctx.flat = packed.flat
        ctx.shapes = [p.shape for p in params]
        ctx.rays_grad = any(t is not None and t.requires_grad for t in (rays_o, rays_d, view_dirs))
        ctx.save_for_backward(acts, sigma, rgb, *((rays_o, rays_d, z, view_dirs) if ctx.rays_grad else ()))
        return sigma, rgb

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        from ..nerf_program import mlp_backward
        acts, sigma, rgb, *rays = ctx.saved_tensors
        flat_grad, dz = mlp_backward(ctx.packed, ctx.flat, acts, sigma, rgb, g_sigma, g_rgb)
        g_o = g_d = g_v = None
        if ctx.rays_grad:                   # learnable cameras: the gradient of the sample points / view directions (SimpleNeRF17.py:210-214)
            from ..nerf_program import mlp_input_backward
            g_o, g_d, g_v = mlp_input_backward(ctx.packed, ctx.flat, dz, *rays)
        grads, o = [], 0
        for shp in ctx.shapes:
            n = 1
            for d in shp:
                n *= d
            grads.append(flat_grad[o:o + n].view(shp))
            o += n
        return (None, g_o, g_d, None, g_v, None, *grads)


class IntrinsicsLearner(torch.nn.Module):
    """SimpleNeRF17.py:788-814 (focal learning is NotImplemented upstream too)."""

    def __init__(self, initial_intrinsics, learn_focal):
        super().__init__()
        self.name = self.__class__.__name__
        k = torch.as_tensor(numpy.asarray(initial_intrinsics).astype(numpy.float32))
        self.initial_intrinsics = torch.nn.Parameter(k, requires_grad=False)
        self.learn_focal = learn_focal
        if self.learn_focal:
            raise NotImplementedError

    def forward(self, cam_id):
        return self.initial_intrinsics[cam_id]

    def get_trainable_parameters(self, optimizer_configs):
        return [{'name': f'{self.name}_params', 'params': self.parameters(),
                 'lr': optimizer_configs['lr_initial'] if optimizer_configs is not None else None}]


class ExtrinsicsLearner(torch.nn.Module):
    """SimpleNeRF17.py:817-912.  With r = t = 0 and nothing learnable (every shipped config) the pose
    correction inv([Exp(r)|t]) is the identity, so view matrices equal `initial_extrinsics`."""

    def __init__(self, initial_extrinsics, learn_rotation, learn_translation):
        super().__init__()
        self.name = self.__class__.__name__
        e = torch.as_tensor(numpy.asarray(initial_extrinsics).astype(numpy.float32))
        self.num_frames = e.shape[0]
        self.initial_extrinsics = torch.nn.Parameter(e, requires_grad=False)
        self.learn_rotation = learn_rotation
        self.learn_translation = learn_translation
        self.r = torch.nn.Parameter(torch.zeros((self.num_frames, 3), dtype=torch.float32), requires_grad=learn_rotation)
        self.t = torch.nn.Parameter(torch.zeros((self.num_frames, 3), dtype=torch.float32), requires_grad=learn_translation)

    def view_matrices(self):
        return self.forward(torch.arange(self.num_frames, device=self.initial_extrinsics.device)).detach()

    def _is_identity_correction(self):
        if self.learn_rotation or self.learn_translation:
            return False
        key = (self.r._version, self.t._version, self.r.data_ptr())
        if getattr(self, '_ident_key', None) != key:
            self._ident_key = key
            self._ident = not bool(self.r.any() or self.t.any())
        return self._ident

    def forward(self, cam_id):
        if self._is_identity_correction():
            # r = t = 0: inv([Exp(0)|0]) is exactly I, and M @ I == M, so skip the per-ray 4x4 inverse (:831-842)
            return self.initial_extrinsics[cam_id]
        r, t = self.r[cam_id], self.t[cam_id]
        return self.initial_extrinsics[cam_id] @ self.make_extrinsics(r, t)

    def get_trainable_parameters(self, optimizer_configs):
        return [{'name': f'{self.name}_params', 'params': self.parameters(),
                 'lr': optimizer_configs['lr_initial'] if optimizer_configs is not None else None}]

    def make_extrinsics(self, r, t):
        rot = self.Exp(r)
        c2w = torch.cat([rot, t.unsqueeze(2)], dim=2)
        bottom = torch.zeros_like(c2w[:, 0:1])
        bottom[:, 0, 3] = 1.0
        return torch.linalg.inv(torch.cat([c2w, bottom], dim=1))

    @staticmethod
    def vec2skew(v):
        zero = torch.zeros((v.shape[0], 1), dtype=torch.float32, device=v.device)
        c0 = torch.cat([zero, -v[:, 2:3], v[:, 1:2]], dim=1)
        c1 = torch.cat([v[:, 2:3], zero, -v[:, 0:1]], dim=1)
        c2 = torch.cat([-v[:, 1:2], v[:, 0:1], zero], dim=1)
        return torch.stack([c0, c1, c2], dim=2)

    @classmethod
    def Exp(cls, r):
        skew = cls.vec2skew(r)
        n = r.norm(dim=1) + 1e-15
        eye = torch.eye(3, dtype=torch.float32, device=r.device)
        return eye[None] + (torch.sin(n) / n)[:, None, None] * skew + ((1 - torch.cos(n)) / n ** 2)[:, None, None] * (skew @ skew)
