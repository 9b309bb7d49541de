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
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in (nerf['pixel_id'], nerf['target_rgb']))
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

    # ------------------------------------------------------------------ multi-GPU: this rank's share of the seeded batch
    def select_batch_indices(self, iter_num, image_num):
        """src/data_preprocessors/DataPreprocessor10.py:496-528, then (several ranks, one process per GPU) this rank's slice.

        Every rank runs the reference's index bookkeeping unchanged — all ranks share `init_seeds(seed)` (src/Trainer10.py:547),
        hence the same numpy permutations and the same wrap-around reshuffles — and keeps rows `shard_bounds(block, rank, world)` of
        the image-ray block and of the sparse-depth block separately, so every rank's loss means weigh the two ray kinds like the
        global batch does.  `configs['data_loader']['rank_sharding']`:
          'strong' (default): the global batch stays `num_rays` (+ sparse-depth `num_rays`); each rank gets 1/world of it;
          'weak': every rank gets a full-size batch — the reference's selection runs `world` times per iteration and rank r keeps
                  the r-th draw — so the global batch is `world` times larger.
        The rank's rows of the global batch travel in the batch dict as `srf_shard` = (rank, world, global_rows, row_indices):
        the drop-in models use it to consume THEIR rows of the single global CPU random stream (App. B), which makes a
        multi-rank run reproduce the single-GPU batch statistics exactly."""
        from .. import parallel
        import numpy
        if image_num is not None:
            return super().select_batch_indices(iter_num, image_num)
        if not parallel.is_distributed():
            n_img = len(self.preprocessed_data_dict['indices'][self.i_batch: self.i_batch + self.num_rays])
            d = super().select_batch_indices(iter_num, image_num)
            return self._with_rows(d, n_img, int(d['indices'].shape[0]))
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        mode = self.configs['data_loader'].get('rank_sharding', 'strong')
        if mode == 'weak':
            picked = None
            for r in range(world):
                n_img = len(self.preprocessed_data_dict['indices'][self.i_batch: self.i_batch + self.num_rays])
                d = super().select_batch_indices(iter_num, image_num)
                if r == rank:
                    picked = self._with_rows(d, n_img, int(d['indices'].shape[0]))
            return picked
        # block sizes from the host-side index arrays (the masks live on the device: reading them would synchronise)
        n_img = len(self.preprocessed_data_dict['indices'][self.i_batch: self.i_batch + self.num_rays])
        d = super().select_batch_indices(iter_num, image_num)
        n_all = int(d['indices'].shape[0])
        parts = []
        for start, size in ((0, n_img), (n_img, n_all - n_img)):
            s_, e_ = parallel.shard_bounds(size, rank, world)
            parts.append(numpy.arange(start + s_, start + e_, dtype=numpy.int64))
        rows = numpy.concatenate(parts)
        rows_dev = torch.from_numpy(rows).to(d['indices'].device)
        out = {k: (v[rows_dev] if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
        out['srf_shard'] = (rank, world, n_all, rows)
        return self._with_rows(out, len(parts[0]), len(rows))

    def _with_rows(self, d, n_img, n_all):
        """`srf_rows`: the row indices of the image rays / sparse-depth rays of the batch (they are [0, n_img) and [n_img, n_all)
        by construction, :520-521) as device index tensors, so the fused losses can index_select instead of boolean-mask
        indexing, which synchronises the stream once per indexed tensor.  Only when the batch is one sub-batch (the trainer
        slices tensors per sub-batch, src/Trainer10.py:88-96)."""
        if n_all <= self.configs.get('sub_batch_size', n_all):
            dev = d['indices'].device
            cache = getattr(self, '_srf_rows_cache', None)
            if cache is None or cache[0] != (n_img, n_all, dev):
                cache = ((n_img, n_all, dev), {'nerf': torch.arange(0, n_img, device=dev), 'sparse_depth': torch.arange(n_img, n_all, device=dev)})
                self._srf_rows_cache = cache
            d['srf_rows'] = dict(cache[1])
        return d

    # ------------------------------------------------------------------ output tail
    def retrieve_inference_outputs(self, network_outputs: dict):
        """src/data_preprocessors/DataPreprocessor10.py:775-803 with the conversion on the device: one kernel (csrc/output.cu)
        writes the uint8 image and the clipped depth maps into one record, one copy brings it to pinned host memory.  The
        reference moves every tensor of the output dict (~190 B/ray) through pageable memory and converts on the CPU."""
        import numpy
        from .. import _lib as L
        h, w = self.model_configs['resolution']
        if 'fine_model' in self.configs['model']:
            suffix = '_fine'
        elif 'coarse_model' in self.configs['model']:
            suffix = '_coarse'
        else:
            raise RuntimeError
        rgb = network_outputs[f'rgb{suffix}']
        if f'visibility2{suffix}' in network_outputs or not rgb.is_cuda:
            return super().retrieve_inference_outputs(network_outputs)
        n = h * w
        maps = [network_outputs[f'depth{suffix}'], network_outputs[f'depth_var{suffix}']]
        if self.ndc:
            maps += [network_outputs[f'depth_ndc{suffix}'], network_outputs[f'depth_var_ndc{suffix}']]
        tensors = [L.f32c(rgb.detach()).reshape(n, 3)] + [L.f32c(m.detach()).reshape(n) for m in maps] + [None] * (4 - len(maps))
        lib = L.load()
        nbytes = int(lib.srf_frame_record_bytes(n, 4))
        record = torch.empty((nbytes,), dtype=torch.uint8, device=rgb.device)
        L.call('srf_frame_outputs', *[L.ptr(t) for t in tensors], n, L.ptr(record), L.stream_handle())
        host = torch.empty((nbytes,), dtype=torch.uint8, pin_memory=True)
        host.copy_(record, non_blocking=True)
        torch.cuda.current_stream(rgb.device).synchronize()
        buf = host.numpy()                                    # the arrays below are views that keep the pinned block alive
        img_bytes, map_bytes = (n * 3 + 15) // 16 * 16, (n * 4 + 15) // 16 * 16
        out = {'image': buf[:n * 3].reshape(h, w, 3)}
        names = ['depth', 'depth_var', 'depth_ndc', 'depth_var_ndc'][:len(maps)]
        for i, name in enumerate(names):
            o = img_bytes + i * map_bytes
            out[name] = buf[o:o + n * 4].view(numpy.float32).reshape(h, w)
        return out
