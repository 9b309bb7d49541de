"""Drop-in for src/loss_functions/TotalVariationLoss04.py: same constructor, `compute_loss` contract, iteration weight
(:118-120) and — quirk preserved — the regulariser only covers the AUGMENTED tensors (:44-78: everything sits inside the
`if self.augmentations_needed` loop); forward and gradient of all planes of a tensor come from ONE launch (csrc/losses.cu,
srf_tv_loss) instead of ~25 eager launches per plane, which made the Simple-TensoRF training iteration host-bound."""
import ctypes
from pathlib import Path

import torch

from .. import _lib as L

this_filename = Path(__file__).stem


class _TvLoss(torch.autograd.Function):
    """sum over planes of 2 (mean dh^2 + mean dw^2) * iter_weight; planes [1,C,H,W] fp32 CUDA parameters."""

    @staticmethod
    def forward(ctx, iter_weight, *planes):
        L.require_cuda(*planes)
        xs = [L.f32c(p.detach()) for p in planes]
        grads = [torch.empty_like(x) for x in xs]
        dims = (ctypes.c_int * (3 * len(xs)))(*[d for x in xs for d in (x.shape[0] * x.shape[1], x.shape[2], x.shape[3])])
        loss = torch.zeros((1,), dtype=torch.float64, device=xs[0].device)
        pp = (ctypes.c_void_p * len(xs))(*[L.ptr(x) for x in xs])
        gg = (ctypes.c_void_p * len(xs))(*[L.ptr(g) for g in grads])
        L.call('srf_tv_loss', pp, gg, dims, len(xs), float(iter_weight), L.ptr(loss), L.stream_handle())
        ctx.grads = grads
        return loss[0].float()

    @staticmethod
    def backward(ctx, g_out):
        grads = ctx.grads
        torch._foreach_mul_(grads, g_out)
        return (None, *grads)


def tv_loss(planes, iter_weight):
    planes = list(planes)
    out = None
    for i in range(0, len(planes), 12):
        part = _TvLoss.apply(iter_weight, *planes[i:i + 12])
        out = part if out is None else out + part
    return out


class TotalVariationLoss:
    def __init__(self, configs: dict, loss_configs: dict):
        self.configs = configs
        self.loss_configs = loss_configs
        self.augmentations_needed = 'augmentations' in self.configs['model']
        self.lr_decay_ratio, self.lr_decay_iters = None, None
        for optimizer_configs in configs['optimizers']:                     # :28-37
            if optimizer_configs['name'] == 'optimizer_main':
                self.lr_decay_ratio = optimizer_configs['lr_decay_ratio']
                self.lr_decay_iters = optimizer_configs['lr_decay_iters'] if optimizer_configs['lr_decay_iters'] is not None \
                    else self.configs['num_iterations']

    def get_iter_weight(self, iter_num):
        return self.lr_decay_ratio ** ((iter_num + 1) / self.lr_decay_iters)

    @staticmethod
    def get_components(tensor):
        """:85-95: the regularised factors are the lines of a CP tensor ([1,C,L,1]: only the difference along L exists, the empty
        one contributes 0 over max(numel, 1)) and the planes of a VM tensor."""
        if tensor.__class__.__name__ == 'CpDecomposedTensor':
            return tensor.vectors_density, tensor.vectors_color
        if tensor.__class__.__name__ == 'VmDecomposedTensor':
            return tensor.matrices_density, tensor.matrices_color
        raise RuntimeError

    def compute_loss(self, input_dict: dict, output_dict: dict, model, return_loss_maps: bool = False):
        total_loss = torch.zeros((), dtype=torch.float32, device=input_dict['target_rgb'].device)
        iter_weight = self.get_iter_weight(input_dict['iter_num'])
        module = model.module if hasattr(model, 'module') else model
        if self.augmentations_needed:
            for aug_cfg in self.configs['model']['augmentations']:
                matches = [a for a in module.augmented_models if a['name'] == aug_cfg['name']]
                if len(matches) != 1:
                    raise RuntimeError
                for tag in ('coarse_model', 'fine_model'):
                    if tag not in aug_cfg:
                        continue
                    tensor = matches[0][tag]
                    density, color = self.get_components(tensor)
                    wd, wc = self.loss_configs['weight_density'], self.loss_configs['weight_color']
                    if wd == wc:
                        total_loss = total_loss + wd * tv_loss([*density, *color], iter_weight)
                    else:
                        total_loss = total_loss + wd * tv_loss(density, iter_weight) + wc * tv_loss(color, iter_weight)
        loss_dict = {'loss_value': total_loss}
        if return_loss_maps:
            loss_dict['loss_maps'] = {}
        return loss_dict
