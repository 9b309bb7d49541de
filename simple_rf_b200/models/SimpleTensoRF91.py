"""Drop-in Simple-TensoRF model class backed by the sm_100a kernels.

Resolvable by the reference's unmodified factory (src/models/ModelFactory02.py:10-22): module
`SimpleTensoRF91` -> class `SimpleTensoRF`.  Mirrors the public surface of src/models/SimpleTensoRF09.py:
`forward` dict contract (:133-162, :194-351), `get_trainable_parameters` with the `*_tensor_params` /
`*_network_params` groups (:1167-1194), `optimizers` hand-off (:98-113), in-forward model surgery
(alphaMask rebuild, shrink, upsampling, optimiser reconfiguration, :821-944), `augmented_models`, and
state-dict compatible names (`coarse_model.matrices_density.0`, `coarse_model.alpha_mask.alpha_volume`, ...).

Per-ray arithmetic (ray generation, stratified depths, occupancy test, compaction, VM density / appearance
gathers, colour MLP, compositing) runs in hand-written CUDA kernels; planes stay `nn.Parameter`s of logical
shape [1,C,H,W] (TotalVariationLoss04.py:85-95 differentiates them directly), channels-last and bit-packed
copies are derived per call.  Grid surgery (5 iterations out of 25 000) is a precomputed plan executed on the kernels of
csrc/tensorf_surgery.cu (grid_surgery.py).
"""
import numpy
import torch
from torch.nn import ModuleDict, ModuleList

from .. import _lib as L
from .. import grid_surgery as GS
from .. import ops
from .. import tensorf_ops as T
from ..nerf_program import PackedRowsMLP
from .SimpleNeRF91 import ExtrinsicsLearner, IntrinsicsLearner, coarse_ladder_on


class SimpleTensoRF(torch.nn.Module):
    def __init__(self, configs: dict, model_configs: dict):
        super().__init__()
        self.configs = configs
        self.model_configs = model_configs
        self.ndc = self.configs['data_loader']['ndc']          # False: world-space box marching (SimpleTensoRF09.py:388-400)
        mc = self.configs['model']
        self.coarse_model_needed = 'coarse_model' in mc
        self.fine_model_needed = 'fine_model' in mc
        if self.fine_model_needed:
            raise NotImplementedError('shipped Simple-TensoRF configs have no fine model')
        if self.coarse_model_needed and mc['coarse_model']['predict_visibility']:
            raise NotImplementedError('predict_visibility raises upstream as well (SimpleTensoRF09.py:1274-1275)')
        self.rng_mode = mc.get('rng_mode', 'reference')
        # rays per launch group at test time: 65 536 rays x 1083 samples keep ~14 GB of per-sample intermediates live (of 180 GB)
        # and render a frame 4 % faster than 16 384-ray groups; results do not depend on it
        self.eval_chunk = int(mc.get('eval_chunk', 1 << 16))
        self.coarse_model = None
        self.fine_model = None
        self.augmentations_needed = 'augmentations' in mc
        if self.augmentations_needed:
            self.augmented_models = []
            self.augmented_models_nn = []
        self._camera_tables = None
        self.build_nerf()
        self.optimizers = None
        self.train_data_preprocessor = None

    def build_nerf(self):
        mc = self.configs['model']
        if self.coarse_model_needed:
            self.coarse_model = get_tensor_model('coarse_model', self.configs, mc['coarse_model'], self.model_configs)
        self.intrinsics_learner = IntrinsicsLearner(numpy.array(self.model_configs['intrinsics']),
                                                    learn_focal=mc['learn_camera_focal_length'])
        self.extrinsics_learner = ExtrinsicsLearner(numpy.array(self.model_configs['extrinsics']),
                                                    learn_rotation=mc['learn_camera_rotation'],
                                                    learn_translation=mc['learn_camera_translation'])
        if self.augmentations_needed:
            for aug in mc['augmentations']:
                nn_dict = ModuleDict()
                entry = {'name': aug['name'], 'coarse_model': None, 'fine_model': None}
                if 'coarse_model' in aug:
                    entry['coarse_model'] = get_tensor_model(f"{aug['name']}_coarse_model", self.configs, aug['coarse_model'],
                                                             self.model_configs)
                    nn_dict['coarse_model'] = entry['coarse_model']
                if 'fine_model' in aug:
                    raise NotImplementedError
                self.augmented_models.append(entry)
                self.augmented_models_nn.append(nn_dict)
            self.augmented_models_nn = ModuleList(self.augmented_models_nn)

    def _tensors(self):
        out = [self.coarse_model] if self.coarse_model is not None else []
        if self.augmentations_needed:
            out += [a['coarse_model'] for a in self.augmented_models if a['coarse_model'] is not None]
        return out

    def get_trainable_parameters(self, optimizer_configs):
        groups = []
        for t in self._tensors():
            groups.extend(t.get_trainable_parameters(optimizer_configs))
        return groups

    def __setattr__(self, name, value):
        super().__setattr__(name, value)
        if name == 'optimizers' and value is not None:           # SimpleTensoRF09.py:98-113
            for t in self._tensors():
                t.optimizers = value
            # optimiser tail: Adam optimisers step through one fused kernel over flat buffers, with ONE
            # all-reduce of the flat gradient bucket on several ranks; other optimisers get the generic all-reduce hook
            from .. import optim, parallel
            super().__setattr__('_fused_adam', optim.attach(value) if optim.enabled() else [])
            super().__setattr__('_grad_allreduce', parallel.attach_gradient_allreduce(value))

    def rebuild_camera_params_learners(self, *, intrinsics: numpy.ndarray = None, extrinsics=None, device):
        mc = self.configs['model']
        if intrinsics is not None:
            self.intrinsics_learner = IntrinsicsLearner(intrinsics, learn_focal=mc['learn_camera_focal_length']).to(device)
        if extrinsics is not None:
            self.extrinsics_learner = ExtrinsicsLearner(extrinsics, learn_rotation=mc['learn_camera_rotation'],
                                                        learn_translation=mc['learn_camera_translation']).to(device)
        self._camera_tables = None

    def forward(self, input_batch: dict, *, retraw: bool = False, sec_views_vis: bool = False, mode: str = None):
        pixel_id = input_batch['pixel_id']
        L.require_cuda(pixel_id)
        if self.training and mode != 'test_camera_params_optimization' and input_batch['sub_batch_index'] == 0:
            self.run_model_modifications(input_batch['iter_num'])          # SimpleTensoRF09.py:141-145
        image_id = pixel_id[:, 0].long()
        intrinsics = self.intrinsics_learner(image_id)
        # the reference flips x of the per-ray extrinsics in place through a view (SimpleTensoRF09.py:206, App. C5)
        extrinsics = self.extrinsics_learner(image_id).clone()
        extrinsics[:, 0, 3] *= -1
        all_extrinsics = self.extrinsics_learner(torch.arange(input_batch['num_frames'], device=pixel_id.device))
        if mode == 'camera_params_only':
            out = {}
        else:
            out = self.render(input_batch, retraw=retraw or self.training, mode=mode)
        out['intrinsics'] = intrinsics
        out['extrinsics'] = extrinsics
        out['extrinsics_all'] = all_extrinsics
        return out

    def run_model_modifications(self, iter_num):
        for t in self._tensors():
            t.run_model_modifications(iter_num)

    def _rays(self, pixel_id, h, w):
        """get_rays_tr (+0.5 px, x flip) + get_ndc_rays_tr + get_view_dirs_tr (SimpleTensoRF09.py:205-239) in one launch.  Learnable cameras under
        autograd: same kernel, same values, attached to the graph of ExtrinsicsLearner.forward (camera_grad.py); the gradient reaches the pose
        through the view directions of the colour MLP, |d| under delta and the NDC -> world depths — in world space: the entry depth of the
        box march, `_render_rays_world` — (the grid coordinates are detached upstream, :1054, :1075)."""
        ndc = self.ndc
        flags = dict(half_pixel=True, flip_x=True, ndc=ndc, viewdirs_from_ndc=ndc)
        learner = self.extrinsics_learner
        if torch.is_grad_enabled() and (learner.r.requires_grad or learner.t.requires_grad):
            from .. import camera_grad
            dev = pixel_id.device
            k = self.intrinsics_learner.initial_intrinsics
            key = (str(dev), k.data_ptr(), k._version)
            if getattr(self, '_intrinsic_tables', None) is None or self._intrinsic_tables[0] != key:
                k_inv, _, focal = ops.camera_tables(k, torch.eye(4)[None].expand(k.shape[0], 4, 4), dev)
                self._intrinsic_tables = (key, k_inv, focal)
            views = learner(torch.arange(learner.num_frames, device=dev))
            return camera_grad.rays_with_camera_gradient(views, pixel_id, k, self._intrinsic_tables[1], self._intrinsic_tables[2], h, w,
                                                         self.model_configs['near'], **flags)
        return ops.raygen(pixel_id, self._tables(pixel_id.device), h, w, self.model_configs['near'], **flags)

    def _tables(self, device, intrinsics=None, extrinsics=None):
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

    def render(self, input_dict: dict, *, retraw: bool, mode: str):
        from .. import parallel
        pixel_id = input_dict['pixel_id']
        num_rays = pixel_id.shape[0]
        chunk = self.configs['model']['chunk'] if self.training else max(self.eval_chunk, self.configs['model']['chunk'])
        # test time on several ranks: this rank's row band of the frame, then one all-gather of the per-ray maps (SURVEY.md §8e iii)
        band = None if self.training else parallel.eval_band(num_rays, self.configs['model'])
        if band is not None:
            pixel_id = pixel_id[band[0]:band[1]]
        parts = [self.render_rays(pixel_id[i:i + chunk], input_dict, retraw=retraw, mode=mode)
                 for i in range(0, pixel_id.shape[0], chunk)]
        out = parts[0] if len(parts) == 1 else {k: torch.cat([p[k] for p in parts], dim=0) for k in parts[0]}
        if band is not None:
            out = parallel.gather_ray_outputs(out, num_rays)
        return out

    def render_rays(self, pixel_id, input_dict, *, retraw, mode):
        mc = self.configs['model']
        dev = pixel_id.device
        R = pixel_id.shape[0]
        h, w = self.model_configs['resolution']
        ndc = self.ndc
        rays_o, rays_d, o_ndc, d_ndc, view_dirs = self._rays(pixel_id, h, w)
        if mode == 'static_camera':                                          # SimpleTensoRF09.py:222-233
            cd = input_dict['common_data']
            pose = cd['processed_view_pose']
            pose = pose[0] if pose.dim() == 3 else pose
            k_view = cd.get('view_intrinsic')
            if k_view is not None and k_view.dim() == 3:
                k_view = k_view[0]
            nviews = self.intrinsics_learner.initial_intrinsics.shape[0]
            ks = self.intrinsics_learner.initial_intrinsics.detach() if k_view is None else k_view[None].expand(nviews, 3, 3)
            tabs = self._tables(dev, ks, pose[None].expand(nviews, 4, 4))
            view_dirs = ops.raygen(pixel_id, tabs, h, w, self.model_configs['near'], half_pixel=False, flip_x=False, ndc=ndc,
                                   viewdirs_from_ndc=ndc)[4]
        out = {'rays_o': rays_o, 'rays_d': rays_d, 'view_dirs': view_dirs}
        if ndc:
            out['rays_o_ndc'], out['rays_d_ndc'] = o_ndc, d_ndc
        main = self.coarse_model
        S = main.host_geometry()['num_samples']
        perturb = self.training and mc['perturb']
        if not ndc:
            return self._render_rays_world(out, pixel_id, input_dict, S, perturb, retraw=retraw, mode=mode)
        ladder = coarse_ladder_on(dev, S, self.model_configs['near_ndc'], self.model_configs['far_ndc'], mc['lindisp'])
        # test time without per-sample outputs: the depths are ONE ladder for all rays -> fused march, z[R,S] never materialised
        if not self.training and not retraw and not torch.is_grad_enabled() and mc.get('fused_eval', True) and main.has_fused_march:
            rays = dict(rays_o=rays_o, rays_d=rays_d, rays_o_ndc=o_ndc, rays_d_ndc=d_ndc, view_dirs=view_dirs, z=None, ladder=ladder)
            for k, v in main(rays, False, white_bkgd=mc['white_bkgd']).items():
                out[f'{k}_coarse'] = v
            return out
        from .. import parallel
        shard = input_dict.get('srf_shard') if self.training else None
        if perturb and self.rng_mode == 'reference':
            jitter = parallel.rows_of_global_draw(lambda n: torch.rand([n, S]), R, shard, mc['chunk'])
            z = ops.stratified_z(ladder, R, jitter=jitter.to(dev))                                  # SimpleTensoRF09.py:379
        elif perturb:
            parallel.decorrelate_device_rng()
            z = ops.stratified_z(ladder, R, jitter=torch.rand([R, S], device=dev))     # torch's CUDA generator: CUDA-graph safe
        else:
            z = ops.stratified_z(ladder, R)
        out['z_vals_coarse'] = z
        rays = dict(rays_o=rays_o, rays_d=rays_d, rays_o_ndc=o_ndc, rays_d_ndc=d_ndc, view_dirs=view_dirs, z=z)
        for k, v in main(rays, retraw, white_bkgd=mc['white_bkgd']).items():
            out[f'{k}_coarse'] = v
        if self.augmentations_needed and self.training and mode != 'test_camera_params_optimization':
            for aug in self.augmented_models:
                if aug['coarse_model'] is not None:                        # same z_vals as the main tensor (:283-296)
                    for k, v in aug['coarse_model'](rays, retraw, white_bkgd=mc['white_bkgd']).items():
                        out[f"{aug['name']}_{k}_coarse"] = v
        if not retraw:
            for k in [k for k in out if k.startswith('z_vals_') or '_alpha_' in f'_{k}' or '_visibility_' in f'_{k}'
                      or '_weights_' in f'_{k}']:
                del out[k]
        return out

    def _render_rays_world(self, out, pixel_id, input_dict, S, perturb, *, retraw, mode):
        """`data_loader.ndc = False` (SimpleTensoRF09.py:388-400, :263): every ray marches from its entry into the main tensor's
        box in steps of the tensor's `step_size`; sample points, view directions and compositing (last interval to 1e10) live in
        world space.  Depths differ per ray, so test time takes the same per-sample path as training."""
        from .. import parallel
        mc = self.configs['model']
        main = self.coarse_model
        rays_o, rays_d = out['rays_o'], out['rays_d']
        R = pixel_id.shape[0]
        jitter = None
        if perturb and self.rng_mode == 'reference':          # one draw per ray on the CPU generator (:397-398)
            shard = input_dict.get('srf_shard') if self.training else None
            jitter = parallel.rows_of_global_draw(lambda n: torch.rand([n, 1]), R, shard, mc['chunk']).to(rays_o.device)
        elif perturb:
            parallel.decorrelate_device_rng()
            jitter = torch.rand([R, 1], device=rays_o.device)
        z = ops.box_march_z(rays_o, rays_d, S, main.host_geometry()['box'], self.model_configs['near'], self.model_configs['far'],
                            float(main.step_size), jitter)
        if rays_o.requires_grad or rays_d.requires_grad:       # learnable cameras: the entry depth moves with the pose (:390-394)
            from .. import camera_grad
            t = camera_grad.box_entry_depth(rays_o, rays_d, main.host_geometry()['box'], self.model_configs['near'], self.model_configs['far'])
            z = z + (t - t.detach())[:, None]
        out['z_vals_coarse'] = z
        rays = dict(rays_o=rays_o, rays_d=rays_d, rays_o_ndc=None, rays_d_ndc=None, view_dirs=out['view_dirs'], z=z)
        for k, v in main(rays, retraw, white_bkgd=mc['white_bkgd']).items():
            out[f'{k}_coarse'] = v
        if self.augmentations_needed and self.training and mode != 'test_camera_params_optimization':
            for aug in self.augmented_models:
                if aug['coarse_model'] is not None:
                    for k, v in aug['coarse_model'](rays, retraw, white_bkgd=mc['white_bkgd']).items():
                        out[f"{aug['name']}_{k}_coarse"] = v
        if not retraw:
            for k in [k for k in out if k.startswith('z_vals_') or '_alpha_' in f'_{k}' or '_visibility_' in f'_{k}'
                      or '_weights_' in f'_{k}']:
                del out[k]
        return out


class AlphaGridMask(torch.nn.Module):
    """Binary occupancy volume [1,1,Z,Y,X] + the box it was built on (SimpleTensoRF09.py:1323-1367).  Buffer names, shapes and
    the bool checkpoint form are the reference's; in memory the volume stays bool (the reference keeps fp32 only because
    `grid_sample` needs it) and the kernels read a 1-bit-per-voxel cache derived from it."""

    def __init__(self, alpha_volume, bounding_box):
        super().__init__()
        volume = alpha_volume if alpha_volume.dtype == torch.bool else alpha_volume > 0
        z, y, x = volume.shape[-3:]
        self.register_buffer('alpha_volume', volume.reshape(1, 1, z, y, x))
        self.register_buffer('bounding_box', bounding_box)
        self.register_buffer('bounding_box_size', bounding_box[1] - bounding_box[0])
        self.register_buffer('resolution', torch.tensor([x, y, z], dtype=torch.long, device=volume.device))
        self._bits = None

    def packed(self):
        """1 bit / voxel derived cache + host-side box description for the occupancy tests."""
        key = (self.alpha_volume.data_ptr(), self.alpha_volume._version)
        if self._bits is None or self._bits[0] != key:
            self._bits = (key, {'bits': T.pack_alpha_bits(self.alpha_volume), 'res': self.resolution.tolist(),
                                'box_min': self.bounding_box[0].tolist(), 'box_size': self.bounding_box_size.tolist()})
        return self._bits[1]


class MlpFeaturesColorPredictor(torch.nn.Module):
    """SimpleTensoRF09.py:1370-1421 with both encoding degrees 0 (the shipped setting): parameter container with
    the reference's `mlp.{0,2,4}` names; evaluated by the tensor-core rows MLP."""

    def __init__(self, tensor_configs, features_dim, num_units):
        super().__init__()
        if tensor_configs['features_positional_encoding_degree'] != 0 or tensor_configs['views_positional_encoding_degree'] != 0:
            raise NotImplementedError('colour predictor encodings other than degree 0 are not shipped')
        self.input_dim = features_dim + (3 if tensor_configs['use_view_dirs'] else 0)
        Lin = torch.nn.Linear
        self.mlp = torch.nn.Sequential(Lin(self.input_dim, num_units), torch.nn.ReLU(inplace=True), Lin(num_units, num_units),
                                       torch.nn.ReLU(inplace=True), Lin(num_units, 3), torch.nn.Sigmoid())
        torch.nn.init.constant_(self.mlp[-2].bias, 0)
        self._packed = PackedRowsMLP(sum(tensor_configs['num_components_color']), features_dim,
                                     3 if tensor_configs['use_view_dirs'] else 0, prefix='mlp', units=num_units)
        self._version = None

    def packed(self, basis):
        """basis: basis_matrix_color.weight [F, sum(C)] (:1151), folded into the first tensor-core layer."""
        params = dict(self.named_parameters())
        version = tuple((p.data_ptr(), p._version) for p in [*params.values(), basis])
        if version != self._version:
            self._packed.refresh(params, basis)
            self._version = version
        return self._packed


class _VmColor(torch.autograd.Function):
    """The appearance branch on the compacted surface samples, forward and backward hand-written:
    gather (srf_vm_color_features_fwd, or srf_cp_color_features_fwd when n_planes == 0: a CP tensor's three lines) -> basis o
    colour MLP on the tensor cores (srf_mlp_rows_fwd, activations saved as tile images) | data-gradient chain + weight gradients
    on the tensor cores (srf_nerf_mlp_dgrad / _wgrad) -> scatter of d loss / d products into the planes / lines.  What autograd
    derives for SimpleTensoRF09.py:1241-1272 (VM) / :1064-1089 (CP) + :1411-1421."""

    @staticmethod
    def forward(ctx, predictor, geom, comp, view_dirs, n_planes, basis, *params):
        # the surface count stays on the device (no read-back: a synchronisation here drains the launch queue twice per
        # iteration); every buffer is sized for the worst case comp.total and every kernel stops at the device-side count
        if n_planes == 0:
            mlp = params[3:]
            rows, tables = T.cp_color_rows(geom, comp, view_dirs, list(params[:3]))
        else:
            planes, lines, mlp = params[:n_planes], params[n_planes:2 * n_planes], params[2 * n_planes:]
            rows, tables = T.vm_color_rows(geom, comp, view_dirs, list(planes), list(lines))
        packed = predictor.packed(basis)
        rgb, acts = packed.forward(rows, comp.count, rows.shape[0], save=True)
        ctx.packed, ctx.flat, ctx.geom, ctx.comp, ctx.tables, ctx.n_planes = packed, packed.flat, geom, comp, tables, n_planes
        ctx.view_grad = (view_dirs.shape[0], geom.z.shape[1]) if (view_dirs is not None and view_dirs.requires_grad and packed.num_view == 3) else None
        ctx.save_for_backward(rgb, acts, basis, mlp[0])
        return rgb

    @staticmethod
    def backward(ctx, g_rgb):
        rgb, acts, basis, w0 = ctx.saved_tensors
        packed = ctx.packed
        g_flat, g_rows, dz = packed.backward(acts, rgb, g_rgb, rgb.shape[0], flat=ctx.flat, count=ctx.comp.count, return_dz=True)
        g_view = None
        if ctx.view_grad is not None:     # learnable cameras: every surface sample hands its view-direction gradient to its ray
            num_rays, num_samples = ctx.view_grad
            per_row = packed.view_dirs_backward(dz, rgb.shape[0], flat=ctx.flat, count=ctx.comp.count)       # zero beyond the count
            ray = torch.div(ctx.comp.idx[:per_row.shape[0]].long(), num_samples, rounding_mode='floor').clamp_(0, num_rays - 1)
            g_view = torch.zeros((num_rays, 3), dtype=torch.float32, device=per_row.device).index_add_(0, ray, per_row)
        if ctx.n_planes == 0:
            g_tensor = T.cp_color_rows_backward(ctx.geom, ctx.comp, ctx.tables, g_rows)
        else:
            gp, gl = T.vm_color_rows_backward(ctx.geom, ctx.comp, ctx.tables, g_rows)
            g_tensor = [*gp, *gl]
        g_w0, g_basis = packed.split_first_layer_grad(g_flat, w0.detach().float(), basis.detach().float())
        g_mlp = [g_w0] + [packed.grad_of(g_flat, nm) for nm in packed.names[1:]]
        return (None, None, None, g_view, None, g_basis, *g_tensor, *g_mlp)


def get_tensor_model(name, configs, tensor_configs, model_configs):
    """SimpleTensoRF09.py:535-544."""
    kind = tensor_configs['decomposition_type']
    if kind == 'VectorMatrix':
        return VmDecomposedTensor(name, configs, tensor_configs, model_configs)
    if kind == 'CandecompParafac':
        return CpDecomposedTensor(name, configs, tensor_configs, model_configs)
    raise RuntimeError(f'Unknown tensor decomposition: {kind}')


class LowRankTensor(torch.nn.Module):
    """SimpleTensoRF09.py:582-961: what the two decompositions share (geometry bookkeeping, the forward orchestration, the grid
    surgery schedule, the colour predictor).  Subclasses provide the factor parameters and their gathers."""
    has_fused_march = False

    def __init__(self, name, configs, tensor_configs, model_configs):
        super().__init__()
        self.name = name
        self.configs = configs
        self.tensor_configs = tensor_configs
        self.model_configs = model_configs
        self.ndc = configs['data_loader']['ndc']
        self.predict_visibility = tensor_configs['predict_visibility']
        for buf, val in (('bounding_box', torch.zeros(2, 3)), ('resolution', torch.zeros(3)), ('num_samples', torch.tensor(0)),
                         ('bounding_box_size', torch.zeros(3)), ('voxel_length', torch.zeros(3)), ('step_size', torch.tensor(0))):
            self.register_buffer(buf, val)
        bbox = torch.tensor(tensor_configs.get('bounding_box', model_configs['bounding_box']))
        self.matrix_axes = T.MATRIX_AXES
        self.vector_axes = T.VECTOR_AXES
        self.alpha_mask = None
        self.optimizers = None
        self.update_tensor_params(self.compute_resolution_in_voxels(tensor_configs['num_voxels_initial'], bbox), bbox)
        tc = tensor_configs
        self.build_tensor()
        self.basis_matrix_color = torch.nn.Linear(sum(tc['num_components_color']), tc['features_dimension_color'], bias=False)
        if tc['density_predictor'] not in ('ReLU', 'SoftPlus'):
            raise NotImplementedError
        self.density_predictor = tc['density_predictor']
        if tc['color_predictor'] != 'MLP_Features':
            raise NotImplementedError
        self.color_predictor = MlpFeaturesColorPredictor(tc, tc['features_dimension_color'], tc['num_units_color_predictor'])
        self._register_load_state_dict_pre_hook(self._load_hook)

    def get_trainable_parameters(self, optimizer_configs):
        tensor_params, network_params = torch.nn.ParameterList(), torch.nn.ParameterList()
        for group in self.tensor_parameter_lists():
            tensor_params.extend(group)
        network_params.extend(self.basis_matrix_color.parameters())
        network_params.extend(self.color_predictor.parameters())
        return [{'name': f'{self.name}_tensor_params', 'params': tensor_params, 'lr': optimizer_configs['lr_initial_tensor']},
                {'name': f'{self.name}_network_params', 'params': network_params, 'lr': optimizer_configs['lr_initial_network']}]

    # ---------------------------------------------------------------- geometry bookkeeping (:625-665)
    @staticmethod
    def compute_resolution_in_voxels(num_voxels, bounding_box):
        lo, hi = bounding_box
        voxel = ((hi - lo).prod() / num_voxels).pow(1 / 3)
        return ((hi - lo) / voxel).long()

    def compute_num_samples(self, resolution, voxels_per_sample):
        n = (torch.linalg.norm(resolution.float()) / voxels_per_sample).round().long()
        return min(self.tensor_configs['num_samples_max'], n)

    def update_tensor_params(self, new_resolution, new_bounding_box):
        self.bounding_box = new_bounding_box.float()
        self.bounding_box_size = self.bounding_box[1] - self.bounding_box[0]
        self.resolution = new_resolution.long()
        self.voxel_length = self.bounding_box_size / (self.resolution - 1)
        self.step_size = torch.mean(self.voxel_length) * self.tensor_configs['num_voxels_per_sample']
        self.num_samples = self.compute_num_samples(self.resolution, self.tensor_configs['num_voxels_per_sample'])
        self._host_geom = None          # freed buffers may be re-allocated at the same address: do not trust the pointer key here

    def _load_hook(self, state, prefix, *args, **kwargs):
        """:946-961 + :1196-1212: rebuild the alpha mask and resize planes/lines before loading."""
        if f'{prefix}alpha_mask.alpha_volume' in state:
            self.alpha_mask = AlphaGridMask(state[f'{prefix}alpha_mask.alpha_volume'], state[f'{prefix}alpha_mask.bounding_box'])
        self._resize_parameters(state[f'{prefix}resolution'])

    # ---------------------------------------------------------------- forward (:701-761)
    def host_geometry(self):
        """Host copies of the box / resolution / sample-count buffers (read once per change, not once per chunk: a
        `.tolist()` on a device buffer is a stream synchronisation)."""
        bufs = (self.bounding_box, self.bounding_box_size, self.resolution, self.num_samples)
        key = tuple((b.data_ptr(), b._version) for b in bufs)
        cached = getattr(self, '_host_geom', None)
        if cached is None or cached[0] != key:
            box = self.bounding_box.tolist()
            cached = (key, {'box': box, 'box_min': box[0], 'box_size': self.bounding_box_size.tolist(),
                            'res': [int(v) for v in self.resolution.tolist()], 'num_samples': int(self.num_samples)})
            self._host_geom = cached
        return cached[1]

    def _geometry(self, rays_o_s, rays_d_s, z):
        hg = self.host_geometry()
        return T.VmGeometry(rays_o_s, rays_d_s, z, hg['box_min'], hg['box_size'], hg['res'])

    def forward(self, rays: dict, retraw: bool, white_bkgd=False):
        tc = self.tensor_configs
        if rays.get('z') is None:
            return self.forward_fused_eval(rays, white_bkgd=white_bkgd)
        z = rays['z']
        R, S = z.shape
        ndc = self.ndc
        so, sd = (rays['rays_o_ndc'], rays['rays_d_ndc']) if ndc else (rays['rays_o'], rays['rays_d'])     # what the samples ride on (:263)
        alpha = self.alpha_mask.packed() if self.alpha_mask is not None else None
        valid = T.validity_compact(so, sd, z, self.host_geometry()['box'], alpha)
        geom = self._geometry(so, sd, z)
        sigma = self.density(geom, valid)
        with torch.no_grad():                                               # weights only decide where colour is read (:726)
            w0 = ops.composite(sigma.detach()[..., 0], None, z, rays['rays_o'], rays['rays_d'], sd if ndc else None, ndc=ndc,
                               distance_scale=tc['distance_scale'], per_sample=False)['weights']
        surface = T.threshold_compact(w0, tc['ray_marching_weight_threshold'])
        cp = self.color_predictor
        basis = self.basis_matrix_color.weight
        factors = self.color_factors()                                      # planes then lines (VM) / the three lines (CP)
        color_params = [*factors, *[cp.mlp[i].weight if j == 0 else cp.mlp[i].bias for i in (0, 2, 4) for j in (0, 1)]]
        if torch.is_grad_enabled() and any(p is not None and p.requires_grad for p in [basis, rays['view_dirs']] + color_params):
            rgb_rows = _VmColor.apply(cp, geom, surface, rays['view_dirs'], self.num_color_planes, basis, *color_params)
        else:
            rows, _ = self.color_rows(geom, surface, rays['view_dirs'])
            rgb_rows = cp.packed(basis).forward(rows, surface.count, rows.shape[0])
        rgb = T._ScatterRows.apply(surface, rgb_rows, R * S).view(R, S, 3)
        device_coin = self.training and not white_bkgd and self.configs['model'].get('rng_mode', 'reference') != 'reference'
        white = white_bkgd or bool(self.training and not device_coin and (torch.rand((1,)) < 0.5))      # :746
        vr = ops.composite(sigma[..., 0], rgb, z, rays['rays_o'], rays['rays_d'], sd if ndc else None, ndc=ndc, white_bkgd=white,
                           distance_scale=tc['distance_scale'], per_sample=retraw)      # alpha / visibility are dropped unless retraw
        if device_coin:          # the background coin (:746) drawn on the device: no host decision inside a captured iteration
            coin = (torch.rand((), device=z.device) < 0.5).to(vr['rgb'].dtype)
            vr['rgb'] = vr['rgb'] + coin * (1. - vr['acc'])[:, None]
        out = {k: vr[k] for k in ('acc', 'alpha', 'visibility', 'weights', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc', 'rgb') if k in vr}
        if retraw:
            out['raw_sigma'] = sigma
            out['raw_rgb'] = rgb
        return out

    # ---------------------------------------------------------------- grid surgery (grid_surgery.py; reference :821-944, :1277-1320)
    def run_model_modifications(self, iter_num):
        """Execute this iteration's steps of the precomputed plan (training only, :822,:827)."""
        if not self.training:
            return
        if getattr(self, '_surgery_plan', None) is None:
            self._surgery_plan = GS.build_plan(self.tensor_configs)
        for step, arg in self._surgery_plan.get(iter_num, ()):
            if step == 'occupancy':
                occupied_box = self.rebuild_alpha_mask()
                if arg:
                    self.crop_to(occupied_box)
            else:
                self.resample_to(self.compute_resolution_in_voxels(arg, self.bounding_box))
                self.regroup_optimizer()

    def rebuild_alpha_mask(self):
        """New occupancy mask on the current grid (:849-876) -> bounding box of the occupied voxels."""
        hg = self.host_geometry()
        planes, lines = self.density_factors()
        volume, occupied_box = GS.rebuild_occupancy(
            planes, lines,
            {'box': self.bounding_box, 'box_min': hg['box_min'], 'box_size': hg['box_size'], 'res': hg['res']},
            step_size=float(self.step_size), threshold=self.tensor_configs['alpha_mask_threshold'],
            softplus=self.density_predictor == 'SoftPlus', density_offset=self.tensor_configs['density_offset'],
            previous=self.alpha_mask.packed() if self.alpha_mask is not None else None)
        self.alpha_mask = AlphaGridMask(volume, self.bounding_box)
        return occupied_box

    def crop_to(self, occupied_box):
        """Cut the tensor down to the voxel window around `occupied_box` (:899-914, :1299-1320 / :1113-1124)."""
        lo, hi, box = GS.crop_window(self.bounding_box, self.voxel_length, self.resolution, occupied_box, self.alpha_mask.resolution)
        dev = self.bounding_box.device
        self._crop_parameters(lo, hi)
        self.update_tensor_params((hi - lo).to(dev), box.to(dev))

    def resample_to(self, resolution):
        """Bilinear resampling of every plane / line to `resolution` voxels (:832-835, :1277-1297 / :1094-1111)."""
        self.update_tensor_params(resolution, self.bounding_box)
        self._resize_parameters(self.resolution)

    def regroup_optimizer(self):
        """The trainer's optimiser must now hold the new Parameter objects (:916-944)."""
        opt_cfg = next(c for c in self.configs['optimizers'] if c['name'] == 'optimizer_main')
        GS.regroup_optimizer(self.optimizers['optimizer_nerf'], self.get_trainable_parameters(opt_cfg))


class VmDecomposedTensor(LowRankTensor):
    """SimpleTensoRF09.py:1126-1320: three (plane, line) pairs per quantity."""
    has_fused_march = True
    num_color_planes = 3

    def build_tensor(self):
        tc = self.tensor_configs
        self.matrices_density, self.vectors_density = self.create_decomposed_tensor(tc['num_components_density'], self.resolution, 0.1)
        self.matrices_color, self.vectors_color = self.create_decomposed_tensor(tc['num_components_color'], self.resolution, 0.1)

    def create_decomposed_tensor(self, comps, resolution, scale):
        mats, vecs = [], []
        for i in range(3):
            a0, a1 = self.matrix_axes[i]
            mats.append(torch.nn.Parameter(scale * torch.randn((1, comps[i], resolution[a1], resolution[a0]))))
            vecs.append(torch.nn.Parameter(scale * torch.randn((1, comps[i], resolution[self.vector_axes[i]], 1))))
        return torch.nn.ParameterList(mats), torch.nn.ParameterList(vecs)

    def tensor_parameter_lists(self):                       # group order of :1170-1173
        return (self.vectors_density, self.matrices_density, self.vectors_color, self.matrices_color)

    def density(self, geom, comp):
        return T.vm_density(geom, comp, list(self.matrices_density), list(self.vectors_density),
                            softplus=self.density_predictor == 'SoftPlus', offset=self.tensor_configs['density_offset'])

    def density_factors(self):
        return self.matrices_density, self.vectors_density

    def color_factors(self):
        return [*self.matrices_color, *self.vectors_color]

    def color_rows(self, geom, comp, view_dirs):
        return T.vm_color_rows(geom, comp, view_dirs, list(self.matrices_color), list(self.vectors_color))

    def forward_fused_eval(self, rays: dict, white_bkgd=False):
        """LowRankTensor.forward (:701-761) at test time through the fused march (csrc/tensorf_march.cu): per-ray maps and the
        surface list in one pass over the shared depth ladder, colour on the surface samples, per-ray accumulation."""
        tc = self.tensor_configs
        hg = self.host_geometry()
        alpha = self.alpha_mask.packed() if self.alpha_mask is not None else None
        m = T.march(rays['rays_o_ndc'], rays['rays_d_ndc'], rays['rays_o'], rays['rays_d'], rays['ladder'], hg['box'], hg['box_size'], alpha,
                    list(self.matrices_density), list(self.vectors_density), hg['res'], softplus=self.density_predictor == 'SoftPlus',
                    offset=tc['density_offset'], distance_scale=tc['distance_scale'], threshold=tc['ray_marching_weight_threshold'])
        geom = T.VmGeometry(rays['rays_o_ndc'], rays['rays_d_ndc'], rays['ladder'], hg['box_min'], hg['box_size'], hg['res'])
        rows, _ = T.vm_color_rows(geom, m.surface, rays['view_dirs'], list(self.matrices_color), list(self.vectors_color))
        rgb_rows = self.color_predictor.packed(self.basis_matrix_color.weight).forward(rows, m.surface.count, rows.shape[0])
        out = dict(m.maps)
        out['rgb'] = T.ray_accumulate(rgb_rows, m, white_bkgd)
        return out

    def _crop_parameters(self, lo, hi):
        self.matrices_density, self.vectors_density = GS.crop_vm(self.matrices_density, self.vectors_density, lo, hi)
        self.matrices_color, self.vectors_color = GS.crop_vm(self.matrices_color, self.vectors_color, lo, hi)

    def _resize_parameters(self, resolution):
        self.matrices_density, self.vectors_density = GS.resample_vm(self.matrices_density, self.vectors_density, resolution)
        self.matrices_color, self.vectors_color = GS.resample_vm(self.matrices_color, self.vectors_color, resolution)


class CpDecomposedTensor(LowRankTensor):
    """SimpleTensoRF09.py:964-1124 (`decomposition_type = "CandecompParafac"`; selected by no shipped configuration): three line
    factors per component, `num_components[0]` components in every line (:992), on the kernels of csrc/tensorf_cp.cu.  Test time
    takes the per-sample path (there is no fused march for this decomposition)."""
    num_color_planes = 0

    def build_tensor(self):
        tc = self.tensor_configs
        if tc['num_components_density'][0] % 4 or tc['num_components_color'][0] % 4:
            raise NotImplementedError('CP component counts must be multiples of 4 (16-byte texel vectors)')
        self.vectors_density = self.create_decomposed_tensor(tc['num_components_density'], self.resolution, 0.1)
        self.vectors_color = self.create_decomposed_tensor(tc['num_components_color'], self.resolution, 0.1)

    def create_decomposed_tensor(self, comps, resolution, scale):
        return torch.nn.ParameterList([torch.nn.Parameter(scale * torch.randn((1, comps[0], resolution[self.vector_axes[i]], 1)))
                                       for i in range(3)])

    def tensor_parameter_lists(self):                       # :1002-1003
        return (self.vectors_density, self.vectors_color)

    def density(self, geom, comp):
        return T.cp_density(geom, comp, list(self.vectors_density), softplus=self.density_predictor == 'SoftPlus',
                            offset=self.tensor_configs['density_offset'])

    def density_factors(self):
        return None, self.vectors_density

    def color_factors(self):
        return list(self.vectors_color)

    def color_rows(self, geom, comp, view_dirs):
        return T.cp_color_rows(geom, comp, view_dirs, list(self.vectors_color))

    def _crop_parameters(self, lo, hi):
        self.vectors_density = GS.crop_lines(self.vectors_density, lo, hi)
        self.vectors_color = GS.crop_lines(self.vectors_color, lo, hi)

    def _resize_parameters(self, resolution):
        self.vectors_density = GS.resample_lines(self.vectors_density, resolution)
        self.vectors_color = GS.resample_lines(self.vectors_color, resolution)
