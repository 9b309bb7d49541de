"""Host side of the fused NeRF MLP kernel: layer program + packed-weight cache for one MLP variant.

The reference MLP (models/SimpleNeRF17.py:616-785) is described to the kernel as a list of layers whose
A operand is a concatenation of 64-column shared-memory K-blocks:
    region 0 = E  point encoding (63 columns + zero pad)      regions 1..4 = H0..H3 (256 hidden units)
    region 5 = V  view-direction encoding (27 columns + zero pad)
`nn.Linear` weights stay the fp32 `nn.Parameter`s the optimiser and checkpoints see; the bf16,
128-byte-swizzled K-block images the tensor cores read are a derived cache, rebuilt by one gather
whenever the parameters change (PackedMLP.refresh).
"""
import ctypes

import numpy as np
import torch

from . import _lib as L

MAX_LAYERS = 12
MAX_KBLOCKS = 6


class MlpLayer(ctypes.Structure):
    _fields_ = [('num_kblocks', ctypes.c_int32), ('kblock_region', ctypes.c_int32 * MAX_KBLOCKS),
                ('kblock_ksteps', ctypes.c_int32 * MAX_KBLOCKS), ('n', ctypes.c_int32), ('relu', ctypes.c_int32),
                ('write_h', ctypes.c_int32), ('head', ctypes.c_int32), ('bias_offset', ctypes.c_int32),
                ('head_offset', ctypes.c_int32), ('save_slot', ctypes.c_int32), ('weight_offset', ctypes.c_int64)]


class MlpProgram(ctypes.Structure):
    _fields_ = [('num_layers', ctypes.c_int32), ('points_degree', ctypes.c_int32), ('views_degree', ctypes.c_int32),
                ('side_count', ctypes.c_int32), ('layers', MlpLayer * MAX_LAYERS), ('lo_offset', ctypes.c_int64)]


def _enc_width(degree):
    return 3 * (2 * degree + 1)


class PackedMLP:
    """Program + gather indices for one reference MLP (any of the three shipped variants).

    `param_names` is the flat parameter order the gather indices refer to; `refresh(params)` takes the
    live fp32 tensors (dict name -> tensor on the device) and rebuilds the packed blob + side table."""

    def __init__(self, mlp_configs, width=256, views_width=None, depth=None, skips=(4,)):
        cfg = mlp_configs
        if views_width is None:                       # the view layer may be 128 (shipped) or 256 wide: both are layer widths of the kernel
            views_width = int(cfg.get('views_net_width', 128)) if cfg.get('use_view_dirs') else 128
        assert views_width in (128, 256), 'the view layer must be 128 or 256 wide'

        assert width == 256 and cfg['points_net_width'] == 256, 'kernel is specialised for 256-wide trunks'
        # any trunk depth the layer program holds (the shipped models use 8; the skip connection after layer 4 exists from depth 6 on,
        # SimpleNeRF17.py:638-647); depth + feature layer + view layer <= MAX_LAYERS
        depth = int(cfg['points_net_depth']) if depth is None else depth
        assert cfg['points_net_depth'] == depth and 2 <= depth <= MAX_LAYERS - 2, 'trunk depth must be 2..10'
        assert depth != 5, 'a 5-layer trunk concatenates the encoding behind its LAST layer and fails upstream too (SimpleNeRF17.py:732-733, :652)'
        self.cfg = cfg
        self.pdeg = cfg['points_positional_encoding_degree']
        full = _enc_width(self.pdeg)
        assert full <= 63
        self.use_views = bool(cfg['use_view_dirs'])
        self.view_dep = bool(cfg['view_dependent_rgb'])
        assert self.use_views == self.view_dep, 'view dirs are only consumed by the view-dependent head'
        self.vdeg = cfg['views_positional_encoding_degree'] if self.use_views else -1
        if self.use_views:
            assert _enc_width(self.vdeg) <= 32 and cfg['views_net_depth'] == 1 and cfg['views_net_width'] == views_width
        pts_in = full
        if 'points_sigma_positional_encoding_degree' in cfg:
            pts_in = _enc_width(cfg['points_sigma_positional_encoding_degree'])
        extra = full - pts_in                      # encoding columns routed to the view branch instead

        # ---- flat parameter order
        names = []
        for i in range(depth):
            names += [f'pts_linears.{i}.weight', f'pts_linears.{i}.bias']
        names += ['pts_output_linear.weight', 'pts_output_linear.bias']
        if self.view_dep:
            names += ['feature_linear.weight', 'feature_linear.bias', 'views_linears.0.weight', 'views_linears.0.bias',
                      'views_output_linear.weight', 'views_output_linear.bias']
        self.param_names = names
        shapes = {}
        for i in range(depth):
            k = pts_in if i == 0 else (width + pts_in if (i - 1) in skips else width)
            shapes[f'pts_linears.{i}.weight'] = (width, k)
            shapes[f'pts_linears.{i}.bias'] = (width,)
        shapes['pts_output_linear.weight'] = (1 if self.view_dep else 4, width)
        shapes['pts_output_linear.bias'] = (1 if self.view_dep else 4,)
        if self.view_dep:
            shapes['feature_linear.weight'] = (width, width)
            shapes['feature_linear.bias'] = (width,)
            vin = width + extra + _enc_width(self.vdeg)
            shapes['views_linears.0.weight'] = (views_width, vin)
            shapes['views_linears.0.bias'] = (views_width,)
            shapes['views_output_linear.weight'] = (3, views_width)
            shapes['views_output_linear.bias'] = (3,)
        self.shapes = shapes
        offs, o = {}, 0
        for n in names:
            offs[n] = o
            o += int(np.prod(shapes[n]))
        self.flat_size = o
        zero = o                                   # index of the appended zero element

        def widx(name, row, col):
            return offs[name] + row * shapes[name][1] + col

        # column maps: for a layer, a list of K-blocks; each K-block = (region, 64 source columns or -1)
        def enc_block(ncols, src0=0, first=0):
            """E block exposing encoding columns [first, first+ncols) as weight columns src0.."""
            cols = [-1] * 64
            for j in range(ncols):
                cols[first + j] = src0 + j
            return (0, cols)

        def h_blocks(src0):
            return [(1 + b, [src0 + b * 64 + j for j in range(64)]) for b in range(4)]

        layers = []                                # (weight name, n, kblocks, relu, write_h, head)
        for i in range(depth):
            name = f'pts_linears.{i}.weight'
            if i == 0:
                kbs = [enc_block(pts_in)]
            elif (i - 1) in skips:
                kbs = [enc_block(pts_in)] + h_blocks(pts_in)       # cat([input_pts, h]) (:732-733)
            else:
                kbs = h_blocks(0)
            last = i == depth - 1
            head = 0
            if last:
                head = 1 if self.view_dep else 2
            layers.append((name, width, kbs, 1, 0 if (last and not self.view_dep) else 1, head))
        if self.view_dep:
            layers.append(('feature_linear.weight', width, h_blocks(0), 0, 1, 0))
            kbs = h_blocks(0)
            if extra > 0:                                           # [feature | enc[pts_in:] | enc_view] (:703, :765)
                kbs.append(enc_block(extra, src0=width, first=pts_in))
            vcols = [-1] * 64
            for j in range(_enc_width(self.vdeg)):
                vcols[j] = width + extra + j
            kbs.append((5, vcols))
            layers.append(('views_linears.0.weight', views_width, kbs, 1, 0, 3))
        assert len(layers) <= MAX_LAYERS

        self.program, self.blob_elems, self._gather_np, self._side_np = build_program(
            layers, shapes, offs, zero, self.pdeg, self.vdeg,
            head_names={1: 'pts_output_linear', 2: 'pts_output_linear', 3: 'views_output_linear'})
        self._dev = None
        self.blob = None
        self.side = None
        self.offs = offs
        # split-bf16 ("bf16x3") inference program for the 1e-3 fp32 contract: same layers, lo weight images behind the hi images
        self.program_split = MlpProgram.from_buffer_copy(self.program)
        self.program_split.lo_offset = self.blob_elems * 2
        self.program_split.act_slots, self.program_split.v_slot = self.program.act_slots, self.program.v_slot
        self.blob_split = None
        self.backward_plan = build_backward_plan(layers, shapes, offs, zero, self.program)
        macs = 0
        for name, n, kbs, *_ in layers:
            macs += shapes[name][0] * shapes[name][1]
        macs += shapes['pts_output_linear.weight'][0] * width + (3 * views_width if self.view_dep else 0)
        self.macs_per_sample = macs          # unpadded, as SURVEY.md §8d counts them

    def refresh(self, params):
        """params: dict name -> fp32 CUDA tensor (the live nn.Parameters).  Rebuilds blob + side table."""
        dev = params[self.param_names[0]].device
        L.require_cuda(params[self.param_names[0]])
        if self._dev != dev:
            self._gather = torch.from_numpy(self._gather_np).to(dev)
            self._side_idx = torch.from_numpy(self._side_np).to(dev)
            self._dev = dev
        flat = torch.cat([params[n].detach().reshape(-1).float() for n in self.param_names] +
                         [torch.zeros(1, dtype=torch.float32, device=dev)])
        assert flat.numel() == self.flat_size + 1, 'parameter shapes do not match the MLP config'
        self.flat = flat
        self.blob = flat[self._gather].to(torch.bfloat16).contiguous()
        self.side = flat[self._side_idx].contiguous()
        self.blob_split = None
        return self

    def split_blob(self):
        """[hi images | lo images]: lo = bf16(w - float(bf16(w))), built on first use after a refresh."""
        if self.blob_split is None:
            w = self.flat[self._gather]
            self.blob_split = torch.cat([self.blob, (w - self.blob.float()).to(torch.bfloat16)]).contiguous()
        return self.blob_split

    def forward(self, rays_o, rays_d, z, view_dirs=None, noise=None, save=False, split=False):
        """rays_o/rays_d [R,3] (the origin/direction the sample points are built from), z [R,S].
        Returns sigma [R,S,1], rgb [R,S,3] (post-activation, as MLP.forward returns them); with save=True also
        acts uint8 [tiles, act_slots, 16384] (the saved activation tile images) for the backward kernels.
        split=True: split-bf16 operands (three MMAs per K block, fp32-contract accuracy; inference only)."""
        assert self.blob is not None, 'call refresh(params) first'
        assert not (split and save), 'the split-bf16 program is inference-only'
        L.require_cuda(rays_o, rays_d, z, view_dirs, noise)
        rays_o, rays_d, z = L.f32c(rays_o), L.f32c(rays_d), L.f32c(z)
        view_dirs, noise = L.f32c(view_dirs), L.f32c(noise)
        R, S = z.shape
        sigma = torch.empty((R, S, 1), dtype=torch.float32, device=z.device)
        rgb = torch.empty((R, S, 3), dtype=torch.float32, device=z.device)
        if self.use_views and view_dirs is None:
            raise L.SimpleRFNativeError('this MLP variant needs view_dirs')
        acts = None
        prog = self.program_split if split else self.program
        blob = self.split_blob() if split else self.blob
        if save:
            tiles = (R * S + 127) // 128
            acts = torch.empty((tiles, act_tile_images(prog.act_slots), 16384), dtype=torch.uint8, device=z.device)
        L.call('srf_nerf_mlp_fwd', ctypes.addressof(prog), L.ptr(blob), L.ptr(self.side), L.ptr(rays_o),
               L.ptr(rays_d), L.ptr(z), L.ptr(view_dirs if self.use_views else None), L.ptr(noise), R, S,
               L.ptr(sigma), L.ptr(rgb), L.ptr(acts), prog.act_slots if save else 0, 0, max(prog.v_slot, 0),
               L.stream_handle(), work=2.0 * self.macs_per_sample * R * S)    # algorithmic FLOPs (the split program issues 3x the MMAs)
        if save:
            return sigma, rgb, acts
        return sigma, rgb


def build_program(layers, shapes, offs, zero, points_degree, views_degree, head_names):
    """layers: [(weight name, n, [(region, 64 source columns or -1)], relu, write_h, head)].
    Returns (MlpProgram, number of bf16 blob elements, blob gather indices, side-table gather indices);
    indices address the flat parameter vector described by `offs` / `shapes`, `zero` = index of a 0 element."""
    assert len(layers) <= MAX_LAYERS
    prog = MlpProgram()
    prog.num_layers = len(layers)
    prog.points_degree = points_degree
    prog.views_degree = views_degree
    gather, side, woff = [], [], 0
    slot = 1                                                   # saved-tile image slots: 0 = E, layer outputs, then V
    e = np.arange(128 * 64)
    r = e // 64
    unit = (e % 64) // 8
    c = ((unit ^ (r & 7)) * 8) + e % 8
    for li, (name, n, kbs, relu, write_h, head) in enumerate(layers):
        Lr = prog.layers[li]
        Lr.num_kblocks = len(kbs)
        Lr.n, Lr.relu, Lr.write_h, Lr.head = n, relu, write_h, head
        Lr.weight_offset = woff * 2
        Lr.save_slot = slot
        slot += n // 64
        for bi, (region, cols) in enumerate(kbs):
            Lr.kblock_region[bi] = region
            used = max(j for j, cc in enumerate(cols) if cc >= 0) + 1
            Lr.kblock_ksteps[bi] = (used + 15) // 16
        # weight images in streaming order: [128-row half of n][K block] -> 128 x 64 bf16, 128B swizzle
        for nh in range(n // 128):
            for bi, (region, cols) in enumerate(kbs):
                src = np.asarray(cols)[c]
                gather.append(np.where(src >= 0, offs[name] + (nh * 128 + r) * shapes[name][1] + np.maximum(src, 0), zero))
                woff += 128 * 64
        bname = name.replace('.weight', '.bias')
        side += [zero] * (-len(side) % 4)                      # float4 loads in the epilogue
        Lr.bias_offset = len(side)
        side += [offs[bname] + j for j in range(n)]
        if head:
            hname = head_names[head]
            rows = shapes[f'{hname}.weight'][0]
            side += [zero] * (-len(side) % 4)
            Lr.head_offset = len(side)
            side += [offs[f'{hname}.weight'] + r_ * shapes[f'{hname}.weight'][1] + c_ for r_ in range(rows) for c_ in range(n)]
            side += [offs[f'{hname}.bias'] + r_ for r_ in range(rows)]
    prog.side_count = len(side)
    prog.v_slot = slot if views_degree >= 0 else -1             # python-side attributes (not part of the C struct)
    prog.act_slots = slot + (1 if views_degree >= 0 else 0)
    return prog, woff, np.concatenate(gather).astype(np.int64), np.asarray(side, dtype=np.int64)


class PackedRowsMLP:
    """The TensoRF colour predictor (reference models/SimpleTensoRF09.py:1389-1393: Linear(F [+3], 128) ReLU
    Linear(128,128) ReLU Linear(128,3) Sigmoid) composed with basis_matrix_color (:1151, :1263, Linear(sum(C), F) without
    bias): the first tensor-core layer uses W0' = [W0[:, :F] @ B | W0[:, F:]] over the bf16 rows
    [products (sum(C)) | view_dirs (3) | 0] (8..128 wide) that srf_vm_color_features_fwd writes."""

    def __init__(self, num_products, features_dim, num_view=3, prefix='color_predictor.mlp', units=128):
        assert num_products + num_view <= 128 and units == 128
        self.num_products, self.features_dim, self.num_view = num_products, features_dim, num_view
        self.in_cols = num_products + num_view
        self.names = [f'{prefix}.0.weight', f'{prefix}.0.bias', f'{prefix}.2.weight', f'{prefix}.2.bias',
                      f'{prefix}.4.weight', f'{prefix}.4.bias']
        shapes = {self.names[0]: (units, self.in_cols), self.names[1]: (units,), self.names[2]: (units, units),
                  self.names[3]: (units,), self.names[4]: (3, units), self.names[5]: (3,)}
        offs, o = {}, 0
        for n in self.names:
            offs[n] = o
            o += int(np.prod(shapes[n]))
        self.flat_size = o
        two = self.in_cols > 64
        kbs0 = [(0, [j if j < self.in_cols else -1 for j in range(64)])]
        if two:
            kbs0.append((5, [j if j < self.in_cols else -1 for j in range(64, 128)]))
        layers = [(self.names[0], units, kbs0, 1, 1, 0),
                  (self.names[2], units, [(1, list(range(64))), (2, list(range(64, 128)))], 1, 0, 3)]
        self.program, self.blob_elems, self._gather_np, self._side_np = build_program(
            layers, shapes, offs, o, 0, -2 if two else -1, head_names={3: f'{prefix}.4'})
        self._dev = None
        self.blob = self.side = None
        self.macs_per_row = units * self.in_cols + units * units + 3 * units
        self.units = units
        self.offs, self.shapes = offs, shapes
        self.e_slot, self.v_slot, self.act_slots = 0, (5 if two else -1), (6 if two else 5)     # E | L0 out (1,2) | L1 out (3,4) | V
        assert self.program.layers[0].save_slot == 1 and self.program.layers[1].save_slot == 3
        self.backward_plan = self._build_backward_plan(o, two)

    def _build_backward_plan(self, zero, two):
        """dgrad chain dZ1 -> dZ0 -> d(rows) and the weight-gradient work items, in the formats of nerf_mlp_dgrad.cu /
        nerf_mlp_wgrad.cu (transposed-weight images [K block of outputs] of 128 inputs x 64 outputs)."""
        n, offs, names = self.units, self.offs, self.names
        plan = BackwardPlan()
        prog = DgradProgram()
        prog.num_fwd_layers, prog.num_layers = 2, 2
        prog.head_slot, prog.top_slot, prog.top_width = 0, 2, n
        prog.top_mask_slot = 3
        prog.head_kind = 1
        prog.head_w_offset = 0
        side = [offs[names[4]] + r * n + c for r in range(3) for c in range(n)]
        e = np.arange(128 * 64)
        r = e // 64
        unit = (e % 64) // 8
        c = ((unit ^ (r & 7)) * 8) + e % 8
        gather, woff = [], 0
        # backward layer 0: dH0 = dZ1 W1, masked by the L0 activations -> dZ0 (dz slots 4, 5)
        L0 = prog.layers[0]
        L0.num_kblocks, L0.mask_slot, L0.rank1_offset, L0.dz_slot, L0.n_out, L0.rows_cols = n // 64, 1, -1, 4, n, 0
        L0.weight_offset = 0
        for kb in range(n // 64):
            gather.append(offs[names[2]] + (kb * 64 + c) * n + r)
            woff += 128 * 64
        # backward layer 1: d(rows) = dZ0 W0' (no mask; only the product columns are written, as fp32 rows)
        L1 = prog.layers[1]
        L1.num_kblocks, L1.mask_slot, L1.rank1_offset, L1.dz_slot, L1.n_out, L1.rows_cols = n // 64, -1, -1, -1, 128, self.num_products
        L1.weight_offset = woff * 2
        for kb in range(n // 64):
            gather.append(np.where(r < self.in_cols, offs[names[0]] + (kb * 64 + c) * self.in_cols + np.minimum(r, self.in_cols - 1), zero))
            woff += 128 * 64
        prog.side_count = len(side)
        plan.program, plan.dz_slots = prog, 6
        plan.gather_t = np.concatenate(gather).astype(np.int64)
        plan.side_idx = np.asarray(side, dtype=np.int64)
        e_cols = min(64, self.in_cols)
        items = [WgradItem(2, 2, 1, 2, n, 0, n, 0, n, 1, offs[names[2]], offs[names[3]]),
                 WgradItem(4, 2, 0, 1, n, 0, e_cols, 0, self.in_cols, 1, offs[names[0]], offs[names[1]])]
        if two:
            items.append(WgradItem(4, 2, 5, 1, n, 0, self.in_cols - 64, 64, self.in_cols, 0, offs[names[0]], offs[names[1]]))
        items.append(WgradItem(0, 2, 3, 2, 3, 0, n, 0, n, 1, offs[names[4]], offs[names[5]]))
        plan.items = items
        plan._dev = None
        return plan

    def composed_first_layer(self, w0, basis):
        """W0' [units, sum(C) + num_view] (differentiable in torch: the interim backward and the tests use it too)."""
        f = self.features_dim
        return torch.cat([w0[:, :f] @ basis, w0[:, f:f + self.num_view]], dim=1)

    def refresh(self, params, basis):
        dev = params[self.names[0]].device
        if self._dev != dev:
            self._gather = torch.from_numpy(self._gather_np).to(dev)
            self._side_idx = torch.from_numpy(self._side_np).to(dev)
            self._dev = dev
        w0c = self.composed_first_layer(params[self.names[0]].detach().float(), basis.detach().float())
        flat = torch.cat([w0c.reshape(-1)] + [params[n].detach().reshape(-1).float() for n in self.names[1:]] +
                         [torch.zeros(1, dtype=torch.float32, device=dev)])
        assert flat.numel() == self.flat_size + 1
        self.flat = flat
        self.blob = flat[self._gather].to(torch.bfloat16).contiguous()
        self.side = flat[self._side_idx].contiguous()
        return self

    def forward(self, rows, count, max_rows, save=False):
        """rows [max_rows, pitch] bf16, count int32[1] on the device (or None: all rows) -> rgb [max_rows, 3] (rows >=
        count undefined).  save=True also returns the saved activation tile images for `backward`."""
        assert rows.dtype == torch.bfloat16 and rows.shape[1] >= self.in_cols and rows.is_contiguous()
        rgb = torch.empty((max_rows, 3), dtype=torch.float32, device=rows.device)
        acts = torch.empty(((max_rows + 127) // 128, act_tile_images(self.act_slots), 16384), dtype=torch.uint8,
                           device=rows.device) if save else None
        L.call('srf_mlp_rows_fwd', ctypes.addressof(self.program), L.ptr(self.blob), L.ptr(self.side), L.ptr(rows),
               rows.shape[1], L.ptr(count), max_rows, L.ptr(rgb), L.ptr(acts), self.act_slots, self.e_slot, self.v_slot,
               L.stream_handle(), work=(count, 2.0 * self.macs_per_row) if count is not None else 2.0 * self.macs_per_row * max_rows)
        return (rgb, acts) if save else rgb

    def backward(self, acts, rgb, g_rgb, num_rows, flat=None, count=None, return_dz=False):
        """Hand-written backward on the tensor cores (nerf_mlp_dgrad.cu / nerf_mlp_wgrad.cu) of a forward(save=True) over
        num_rows rows (count: device int32[1] with the number of valid rows when num_rows is a worst-case capacity): returns (flat gradient in the layout [W0' | b0 | W1 | b1 | W2 | b2], g_rows fp32
        [num_rows, pitch4] whose first num_products columns are d loss / d product rows)."""
        plan = self.backward_plan
        dev = acts.device
        if plan._dev != dev:
            plan._gather = torch.from_numpy(plan.gather_t).to(dev)
            plan._side = torch.from_numpy(plan.side_idx).to(dev)
            plan._dev = dev
        flat = self.flat if flat is None else flat
        wt = flat[plan._gather].to(torch.bfloat16).contiguous()
        side = flat[plan._side].contiguous()
        tiles = acts.shape[0]
        dz = torch.empty((tiles, plan.dz_slots, 16384), dtype=torch.uint8, device=dev)
        pitch = -(-self.num_products // 4) * 4
        g_rows = torch.empty((max(num_rows, 1), pitch), dtype=torch.float32, device=dev)
        L.call('srf_nerf_mlp_dgrad', ctypes.addressof(plan.program), L.ptr(wt), L.ptr(side), L.ptr(acts), act_data_slots(acts.shape[1]), None, L.ptr(rgb),
               None, L.ptr(L.f32c(g_rgb)), num_rows, L.ptr(count), L.ptr(dz), plan.dz_slots, L.ptr(g_rows), pitch, L.stream_handle(),
               work=2.0 * (2 * self.units * self.units) * num_rows)
        grads = torch.zeros(self.flat_size, dtype=torch.float32, device=dev)
        run_wgrad(plan.items, acts, dz, grads, count=count)
        return (grads, g_rows, dz) if return_dz else (grads, g_rows)

    def view_dirs_backward(self, dz, num_rows, flat=None, count=None):
        """d loss / d view_dirs per ROW [num_rows, 3] (zero beyond `count`) from the dZ0 images `backward(return_dz=True)` left in HBM:
        dZ0 W0'[:, sum(C):sum(C)+3] (learnable cameras only: the view directions are what ties the colour branch to the pose,
        SimpleTensoRF09.py:236-239, :1417-1418)."""
        assert self.num_view == 3
        flat = self.flat if flat is None else flat
        src = InputGradSource(4, self.units // 64, self.in_cols, 0, self.offs[self.names[0]])
        for j in range(64):
            src.cols[j] = self.num_products + j if j < 3 else -1
        arr = (InputGradSource * 1)(src)
        g = torch.empty((max(num_rows, 1), 3), dtype=torch.float32, device=dz.device)
        L.call('srf_nerf_mlp_input_grad', ctypes.addressof(arr), 1, L.ptr(flat), L.ptr(dz), dz.shape[1], None, None, None, None,
               num_rows, L.ptr(count), 0, 0, -1, L.ptr(g), None, L.stream_handle())
        return g

    def split_first_layer_grad(self, g_flat, w0, basis):
        """(dW0, dB) from dW0' = d loss / d [W0[:, :F] B | W0[:, F:]] (chain rule through composed_first_layer)."""
        f, ct = self.features_dim, self.num_products
        gw = g_flat[self.offs[self.names[0]]:self.offs[self.names[0]] + self.units * self.in_cols].view(self.units, self.in_cols)
        g_w0 = torch.cat([gw[:, :ct] @ basis.t(), gw[:, ct:ct + self.num_view]], dim=1)
        g_basis = w0[:, :f].t() @ gw[:, :ct]
        return g_w0, g_basis

    def grad_of(self, g_flat, name):
        o = self.offs[name]
        shape = self.shapes[name]
        return g_flat[o:o + int(np.prod(shape))].view(shape)


def act_tile_images(act_slots):
    """Images per tile of a saved-activation buffer: the data images plus one mask image per 16 of them (the forward writes,
    for every saved image, one 32-bit non-zero mask per (row, 32-column group): csrc/common.cuh act_mask_offset)."""
    return act_slots + (act_slots + 15) // 16


def act_data_slots(tile_images):
    """Inverse of act_tile_images."""
    d = tile_images
    while act_tile_images(d) > tile_images:
        d -= 1
    assert act_tile_images(d) == tile_images, 'not a saved-activation buffer'
    return d


class WgradItem(ctypes.Structure):
    _fields_ = [('dz_slot', ctypes.c_int32), ('dz_images', ctypes.c_int32), ('x_slot', ctypes.c_int32), ('x_images', ctypes.c_int32),
                ('out_rows', ctypes.c_int32), ('in_col0', ctypes.c_int32), ('in_cols', ctypes.c_int32), ('w_col0', ctypes.c_int32),
                ('w_stride', ctypes.c_int32), ('bias', ctypes.c_int32), ('dw_offset', ctypes.c_int64), ('db_offset', ctypes.c_int64)]


def run_wgrad(items, acts, dz, grads, count=None):
    """items: list of WgradItem; acts / dz uint8 [tiles, slots, 16384]; grads flat fp32 (accumulated into)."""
    arr = (WgradItem * len(items))(*items)
    tiles = acts.shape[0]
    flops = sum(2.0 * 64 * it.dz_images * 64 * it.x_images * 128 * tiles for it in items)
    L.call('srf_nerf_mlp_wgrad', ctypes.addressof(arr), len(items), L.ptr(acts), act_data_slots(acts.shape[1]), L.ptr(dz), dz.shape[1], tiles,
           L.ptr(count), L.ptr(grads), L.stream_handle(), work=flops)


class DgradLayer(ctypes.Structure):
    _fields_ = [('num_kblocks', ctypes.c_int32), ('mask_slot', ctypes.c_int32), ('rank1_offset', ctypes.c_int32),
                ('dz_slot', ctypes.c_int32), ('n_out', ctypes.c_int32), ('rows_cols', ctypes.c_int32), ('weight_offset', ctypes.c_int64)]


class DgradProgram(ctypes.Structure):
    _fields_ = [('num_layers', ctypes.c_int32), ('num_fwd_layers', ctypes.c_int32), ('top_width', ctypes.c_int32),
                ('top_mask_slot', ctypes.c_int32), ('top_slot', ctypes.c_int32), ('head_slot', ctypes.c_int32),
                ('head_kind', ctypes.c_int32), ('head_w_offset', ctypes.c_int32), ('side_count', ctypes.c_int32),
                ('pad_', ctypes.c_int32), ('layers', DgradLayer * MAX_LAYERS)]


class InputGradSource(ctypes.Structure):
    """One layer that consumes an encoding image (csrc/nerf_mlp_input_grad.cu): its dZ images, its fp32 weight matrix in the flat
    parameter vector, and which weight column every column of the encoding image multiplies."""
    _fields_ = [('dz_slot', ctypes.c_int32), ('dz_images', ctypes.c_int32), ('in_total', ctypes.c_int32), ('target', ctypes.c_int32),
                ('w_offset', ctypes.c_int64), ('cols', ctypes.c_int32 * 64)]


class BackwardPlan:
    """Everything the dgrad / wgrad kernels need for one MLP variant (built once from the forward layer list)."""
    __slots__ = ('program', 'dz_slots', 'items', 'gather_t', 'side_idx', '_dev', '_gather', '_side', 'input_sources')


def build_backward_plan(layers, shapes, offs, zero, fwd_prog):
    nl = len(layers)
    top = nl - 1
    name_t, n_t, kbs_t, relu_t, _, head_t = layers[top]
    plan = BackwardPlan()
    prog = DgradProgram()
    prog.num_fwd_layers = nl
    prog.head_slot = 0
    prog.top_slot = 2
    prog.top_width = n_t
    prog.top_mask_slot = fwd_prog.layers[top].save_slot
    assert relu_t and head_t in (2, 3)
    prog.head_kind = 1 if head_t == 3 else 2
    hname = 'views_output_linear' if head_t == 3 else 'pts_output_linear'
    rows = shapes[f'{hname}.weight'][0]
    side = [offs[f'{hname}.weight'] + r * n_t + c for r in range(rows) for c in range(n_t)]
    prog.head_w_offset = 0
    # dz slots: 0,1 head images; top; then the output of every backward layer (dZ of forward layer f-1)
    dz_of = {top: 2}
    slot = 2 + n_t // 64
    e = np.arange(128 * 64)
    r = e // 64
    unit = (e % 64) // 8
    c = ((unit ^ (r & 7)) * 8) + e % 8
    gather, woff, bl = [], 0, 0
    sigma_layer = next((i for i, L_ in enumerate(layers) if L_[5] == 1), None)      # forward layer carrying the sigma head
    for f in range(top, 0, -1):
        name, n, kbs, relu, write_h, head = layers[f]
        hb = [(reg, cols) for reg, cols in kbs if 1 <= reg <= 4]
        assert len(hb) == 4 and [reg for reg, _ in hb] == [1, 2, 3, 4], 'hidden input of every layer above 0 is 256 wide'
        src0 = hb[0][1][0]
        in_total = shapes[name][1]
        Lr = prog.layers[bl]
        Lr.num_kblocks = n // 64
        below = layers[f - 1]
        Lr.mask_slot = fwd_prog.layers[f - 1].save_slot if below[3] else -1
        Lr.rank1_offset = -1
        if sigma_layer is not None and sigma_layer == f - 1:
            side += [zero] * (-len(side) % 4)
            Lr.rank1_offset = len(side)
            side += [offs['pts_output_linear.weight'] + j for j in range(256)]
        dz_of[f - 1] = slot
        Lr.dz_slot = slot
        Lr.n_out = 256
        Lr.rows_cols = 0
        slot += 4
        Lr.weight_offset = woff * 2
        for nh in range(2):
            for kb in range(n // 64):
                gather.append(offs[name] + (kb * 64 + c) * in_total + src0 + nh * 128 + r)
                woff += 128 * 64
        bl += 1
    prog.num_layers = bl
    prog.side_count = len(side)
    plan.program = prog
    plan.dz_slots = slot
    plan.gather_t = np.concatenate(gather).astype(np.int64)
    plan.side_idx = np.asarray(side, dtype=np.int64)
    # ---- weight-gradient work items
    items = []
    for f in range(nl):
        name, n, kbs, relu, write_h, head = layers[f]
        in_total = shapes[name][1]
        bname = name.replace('.weight', '.bias')
        first = True
        groups = []                                   # (x_slot, x_images, in_col0, in_cols, w_col0)
        hb = [(reg, cols) for reg, cols in kbs if 1 <= reg <= 4]
        if hb:
            prev_slot = fwd_prog.layers[f - 1].save_slot
            groups.append((prev_slot, len(hb), 0, 64 * len(hb), hb[0][1][0]))
        for reg, cols in kbs:
            if reg in (0, 5):
                valid = [i for i, cc in enumerate(cols) if cc >= 0]
                assert valid == list(range(valid[0], valid[-1] + 1)) and [cols[i] for i in valid] == list(range(cols[valid[0]], cols[valid[0]] + len(valid)))
                groups.append((0 if reg == 0 else fwd_prog.v_slot, 1, valid[0], len(valid), cols[valid[0]]))
        for x_slot, x_images, in_col0, in_cols, w_col0 in groups:
            items.append(WgradItem(dz_of[f], n // 64, x_slot, x_images, n, in_col0, in_cols, w_col0, in_total, 1 if first else 0,
                                   offs[name], offs[bname]))
            first = False
    if prog.head_kind == 1:
        items.append(WgradItem(1, 2, fwd_prog.layers[sigma_layer].save_slot, 4, 1, 0, 256, 0, 256, 1,
                               offs['pts_output_linear.weight'], offs['pts_output_linear.bias']))
        items.append(WgradItem(0, 2, fwd_prog.layers[top].save_slot, n_t // 64, 3, 0, n_t, 0, n_t, 1,
                               offs['views_output_linear.weight'], offs['views_output_linear.bias']))
    else:
        items.append(WgradItem(0, 2, fwd_prog.layers[top].save_slot, 4, 4, 0, 256, 0, 256, 1,
                               offs['pts_output_linear.weight'], offs['pts_output_linear.bias']))
    plan.items = items
    # ---- consumers of the encoding images (input gradients: only learnable cameras ask for them)
    sources = []
    for f in range(nl):
        name, n, kbs, relu, write_h, head = layers[f]
        for reg, cols in kbs:
            if reg in (0, 5):
                src = InputGradSource(dz_of[f], n // 64, shapes[name][1], 0 if reg == 0 else 1, offs[name])
                for j in range(64):
                    src.cols[j] = cols[j]
                sources.append(src)
    plan.input_sources = sources
    plan._dev = None
    return plan


def mlp_backward(packed, params_flat, acts, sigma, rgb, g_sigma, g_rgb):
    """Hand-written backward of one fused-MLP evaluation: dgrad chain then weight gradients.
    params_flat: the flat fp32 parameter vector (+ trailing zero) used for the forward's refresh.
    Returns the flat fp32 gradient (same layout as the parameters)."""
    plan = packed.backward_plan
    dev = acts.device
    if plan._dev != dev:
        plan._gather = torch.from_numpy(plan.gather_t).to(dev)
        plan._side = torch.from_numpy(plan.side_idx).to(dev)
        plan._dev = dev
    wt = params_flat[plan._gather].to(torch.bfloat16).contiguous()
    side = params_flat[plan._side].contiguous()
    tiles = acts.shape[0]
    rows = sigma.numel()
    dz = torch.empty((tiles, plan.dz_slots, 16384), dtype=torch.uint8, device=dev)
    gs = None if g_sigma is None else L.f32c(g_sigma).reshape(-1)
    gc = None if g_rgb is None else L.f32c(g_rgb).reshape(-1, 3)
    L.call('srf_nerf_mlp_dgrad', ctypes.addressof(plan.program), L.ptr(wt), L.ptr(side), L.ptr(acts), act_data_slots(acts.shape[1]), L.ptr(sigma), L.ptr(rgb),
           L.ptr(gs), L.ptr(gc), rows, None, L.ptr(dz), plan.dz_slots, None, 0, L.stream_handle(), work=2.0 * packed.macs_per_sample * rows)
    grads = torch.zeros(packed.flat_size, dtype=torch.float32, device=dev)
    run_wgrad(plan.items, acts, dz, grads)
    return grads, dz


def mlp_input_backward(packed, params_flat, dz, rays_o, rays_d, z, view_dirs):
    """Gradient of one fused-MLP evaluation w.r.t. what its sample points and view directions were built from (learnable cameras,
    SimpleNeRF17.py:817-842): `dz` are the dZ images `mlp_backward` returned.  pts = o + z d, so
    g_o = sum_s g_pts, g_d = sum_s z g_pts, g_view_dirs = sum_s g_views.  Returns (g_rays_o, g_rays_d, g_view_dirs or None), each [R, 3]."""
    plan = packed.backward_plan
    R, S = z.shape
    rows = R * S
    rays_o, rays_d, z = L.f32c(rays_o.detach()), L.f32c(rays_d.detach()), L.f32c(z.detach())
    use_views = packed.use_views and view_dirs is not None
    view_dirs = L.f32c(view_dirs.detach()) if use_views else None
    g_pts = torch.empty((R, S, 3), dtype=torch.float32, device=z.device)
    g_views = torch.empty((R, S, 3), dtype=torch.float32, device=z.device) if use_views else None
    sources = plan.input_sources if use_views else [s for s in plan.input_sources if s.target == 0]
    arr = (InputGradSource * len(sources))(*sources)
    L.call('srf_nerf_mlp_input_grad', ctypes.addressof(arr), len(sources), L.ptr(params_flat), L.ptr(dz), dz.shape[1],
           L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.ptr(view_dirs), rows, None, S, int(packed.pdeg), int(packed.vdeg),
           L.ptr(g_pts), L.ptr(g_views), L.stream_handle())
    g_o = g_pts.sum(1)
    g_d = (g_pts * z[..., None]).sum(1)
    return g_o, g_d, (g_views.sum(1) if use_views else None)
