"""Drop-in for src/loss_functions/CoarseFineConsistencyLoss34.py (depth MSE between the main coarse and fine models, the
more accurate depth chosen by patch reprojection error, plus fine -> coarse supervision on the sparse-depth rays): same
constructor, `compute_loss` contract and loss-map naming, masks from the fused kernel (csrc/losses.cu)."""
from pathlib import Path

import torch

from .patch_reprojection import batch_rows as _rows, consistency_loss_nerf

this_filename = Path(__file__).stem


class CoarseFineConsistencyLoss:
    def __init__(self, configs: dict, loss_configs: dict) -> None:
        self.configs = configs
        self.loss_configs = loss_configs
        self.coarse_model_needed = 'coarse_model' in self.configs['model']
        self.fine_model_needed = 'fine_model' in self.configs['model']
        self.sparse_depth_needed = 'sparse_depth' in self.configs['data_loader']
        self.patch_size = tuple(self.loss_configs['patch_size'])
        self.rmse_threshold = self.loss_configs['rmse_threshold']

    def compute_loss(self, input_dict: dict, output_dict: dict, model, return_loss_maps: bool = False) -> dict:
        total_loss = torch.zeros((), dtype=input_dict['target_rgb'].dtype, device=input_dict['target_rgb'].device)   # no pageable host->device copy
        if not self.coarse_model_needed or not self.fine_model_needed:
            return {'loss_value': total_loss}
        depth_coarse, depth_fine = output_dict['depth_coarse'], output_dict['depth_fine']
        loss, map_coarse, map_fine = consistency_loss_nerf(
            depth_coarse, depth_fine, input_dict['indices_mask_nerf'], output_dict['rays_o'], output_dict['rays_d'],
            output_dict['extrinsics_all'].detach(), input_dict['common_data']['images'], input_dict['pixel_id'],
            output_dict['intrinsics'].detach(), self.patch_size, self.rmse_threshold, both_invalid_rule=False,
            rows_nerf=_rows(input_dict, 'nerf'))
        total_loss = total_loss + loss
        map_sd = None
        if self.sparse_depth_needed:                     # CoarseFineConsistencyLoss34.py:170-187: the fine depth supervises the coarse one
            mask_sd = input_dict.get('indices_mask_sparse_depth', None)
            rows_sd = _rows(input_dict, 'sparse_depth')
            if rows_sd is not None:
                map_sd = torch.square(depth_coarse.index_select(0, rows_sd) - depth_fine.index_select(0, rows_sd).detach())
                if map_sd.numel() > 0:
                    total_loss = total_loss + map_sd.mean()
            elif mask_sd is not None:
                map_sd = torch.square(depth_coarse[mask_sd] - depth_fine[mask_sd].detach())
                if map_sd.numel() > 0:
                    total_loss = total_loss + map_sd.mean()
        loss_dict = {'loss_value': total_loss}
        if return_loss_maps:
            loss_dict['loss_maps'] = {f'{this_filename}_coarse': map_coarse, f'{this_filename}_fine': map_fine}
            if map_sd is not None:
                loss_dict['loss_maps'][f'{this_filename}_coarse_sparse_depth'] = map_sd
        return loss_dict
