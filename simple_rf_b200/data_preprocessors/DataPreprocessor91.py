"""Drop-in for src/data_preprocessors/DataPreprocessor10.py ("next" row f2): everything is inherited from the
reference's class — loading, caching, index shuffling with the reference's numpy RNG order — except the per-iteration batch
assembly (:530-549, :568-595), which becomes one fused gather (csrc/batch.cu) instead of ~30 small launches with two
mask-size synchronisations.  Select it with `configs['data_loader']['data_preprocessor_name'] = 'DataPreprocessor91'`
(resolved by the unmodified src/data_preprocessors/DataPreprocessorFactory01.py:15-22 after `simple_rf_b200.dropin.install()`,
or by copying the shim file into <reference>/src/data_preprocessors/).  Needs the reference on sys.path (it subclasses it)."""
import torch

from data_preprocessors.DataPreprocessor10 import DataPreprocessor as _ReferenceDataPreprocessor

from ..batch import assemble_batch


class DataPreprocessor(_ReferenceDataPreprocessor):
    def _tables(self):
        t = getattr(self, '_srf_tables', None)
        nerf = self.preprocessed_data_dict['nerf_data']
        key = (nerf['pixel_id'].data_ptr(), nerf['target_rgb'].data_ptr())
        if t is None or t['key'] != key:
            dev = nerf['pixel_id'].device
            t = {'key': key, 'pixel': nerf['pixel_id'].to(torch.int32).contiguous(), 'rgb': nerf['target_rgb'].float().contiguous()}
            sd = self.preprocessed_data_dict.get('sparse_depth_data') if self.sparse_depth_needed else None
            if sd is not None:
                t.update(depth=sd['depths'].float().contiguous().to(dev), error=sd['reprojection_errors'].float().contiguous().to(dev),
                         points=sd['points_3d'].float().contiguous().to(dev))
            self._srf_tables = t
        return t

    def load_nerf_cached_batch(self, iter_num, indices_dict):
        t = self._tables()
        mask_sd = indices_dict.get('indices_mask_sparse_depth')
        fields = assemble_batch(indices_dict['indices'], mask_sd, t['pixel'], t['rgb'],
                                *( (t['depth'], t['error'], t['points']) if (mask_sd is not None and 'depth' in t) else () ))
        self._srf_batch_fields = fields
        return {'iter_num': iter_num, 'num_frames': self.preprocessed_data_dict['frame_nums'].size,
                'pixel_id': fields['pixel_id'], 'target_rgb': fields['target_rgb']}

    def load_sparse_depth_cached_batch(self, indices_dict, batch_dict):
        if 'indices_mask_sparse_depth' not in indices_dict:
            return {}
        f = self._srf_batch_fields
        return {'pixel_id': batch_dict['pixel_id'], 'sparse_depth_values': f['sparse_depth_values'],
                'sparse_depth_errors': f['sparse_depth_errors'], 'sparse_depth_points3d': f['sparse_depth_points3d']}
