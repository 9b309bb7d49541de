"""Drop-in for src/loss_functions/AugmentationsDepthLoss11.py (depth MSE between the main model and an augmented model,
the more accurate depth chosen by patch reprojection error): same constructor, `compute_loss` contract and loss-map
naming, masks computed by one fused kernel (csrc/losses.cu) instead of 75 fancy-index gathers per patch set."""
from pathlib import Path

import torch

from .patch_reprojection import batch_rows as _rows, consistency_loss_nerf

this_filename = Path(__file__).stem


class AugmentationsDepthLoss:
    def __init__(self, configs: dict, loss_configs: dict):
        self.configs = configs
        self.loss_configs = loss_configs
        self.coarse_model_needed = 'coarse_model' in self.configs['model']
        self.fine_model_needed = 'fine_model' in self.configs['model']
        self.augmentations_needed = 'augmentations' in self.configs['model']
        self.patch_size = tuple(self.loss_configs['patch_size'])
        self.rmse_threshold = self.loss_configs['rmse_threshold']

    def compute_loss(self, input_dict: dict, output_dict: dict, model, return_loss_maps: bool = False):
        total_loss = torch.zeros((), dtype=input_dict['target_rgb'].dtype, device=input_dict['target_rgb'].device)   # no pageable host->device copy
        loss_maps = {}
        common = (input_dict['indices_mask_nerf'], output_dict['rays_o'], output_dict['rays_d'], output_dict['extrinsics_all'].detach(),
                  input_dict['common_data']['images'], input_dict['pixel_id'], output_dict['intrinsics'].detach())
        if self.augmentations_needed:
            for stage, needed in (('coarse', self.coarse_model_needed), ('fine', self.fine_model_needed)):
                if not needed:
                    continue
                depth_main = output_dict[f'depth_{stage}']
                for aug in self.configs['model']['augmentations']:
                    if f'{stage}_model' not in aug:
                        continue
                    name = aug['name']
                    loss, map1, map2 = consistency_loss_nerf(depth_main, output_dict[f'{name}_depth_{stage}'], *common, self.patch_size,
                                                             self.rmse_threshold, both_invalid_rule=True, rows_nerf=_rows(input_dict, 'nerf'))
                    total_loss = total_loss + loss
                    if return_loss_maps:
                        loss_maps[f'{this_filename}_{name}_{stage}_main'] = map1
                        loss_maps[f'{this_filename}_{name}_{stage}_augmented'] = map2
        loss_dict = {'loss_value': total_loss}
        if return_loss_maps:
            loss_dict['loss_maps'] = loss_maps
        return loss_dict
