#!/usr/bin/env python
"""Benchmark of the Simple-RF per-ray rendering hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--main-only]

Main line — workload `simple_nerf_frame_render` (BASELINE.json configs[0]/[1] shape): Simple-NeRF, synthetic LLFF-shaped scene
(2 input views, 756x1008, focal 815.13, NDC), 64 coarse + 128 fine samples, default-initialised weights; one STEP renders one
full 762 048-ray frame from a new pose.  Multi-GPU: one process per GPU (torchrun), every rank renders its own frame per step —
independent units, no data-path collective — so scaling is weak and `value` is all ranks' rays over the max-over-ranks time.

`value`   : device-resident pixel ids, the drop-in model's forward(), CUDA-event timed.
`e2e`     : the reference's UNMODIFIED caller `NerfTester.predict_frame` (src/Tester07.py:153-173) driving the drop-in classes:
            host pose in, `create_test_data` (host pixel ids -> device), camera rebuild, forward, output record -> pinned host
            numpy image + depth maps.  Falls back to forward() with pinned host buffers when no upstream tree is installed.
`roofline`: fused tcgen05 MLP kernel — algorithmic FLOPs (unpadded MACs x 2) per launch over the CUDA-event
            duration of each launch, against the measured bf16 peak in MEASURED_PEAKS.json.
`workloads`: the rest of BASELINE.json's `metric`, measured in the same run at every N, each with its own roofline:
            Simple-NeRF training it/s (weak: 4096 rays per rank; strong: 4096 global), Simple-TensoRF training it/s,
            Simple-TensoRF trajectory frames/s, the main frame with `mlp_precision='bf16x3'` (the fp32-contract program), and a single
            frame sharded over the ranks — training steps are ONE CUDA graph each and go through the fused flat Adam with ONE NCCL
            all-reduce of the gradient bucket (its bytes and time are reported).
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference classes from baseline/_ref (tools/install_reference.sh) on the host
            CPU (kind "reference"); the CPU oracle port (kind "port") only if no upstream tree is installed.
`gpu_reference_bar`: the reference's own eager PyTorch path on the same B200 (N = 1 only) — the bar the kernels must beat.
"""
import argparse
import copy
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FRAME_H, FRAME_W = 756, 1008
S_COARSE, N_FINE = 64, 128
NERF_FLOP_PER_RAY_RENDER = (S_COARSE + S_COARSE + N_FINE) * 1186816.0


def measured_peaks():
    f = ROOT / 'MEASURED_PEAKS.json'
    if f.exists():
        d = json.loads(f.read_text())
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'], 'bf16_tflops_sustained': d['bf16_tflops_sustained'],
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


L2_GBS = 12000.0       # measured L2-resident read bandwidth of this part (tools/l2_probe.py, profiles/r01_hbm_microbench.md)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}


def main_config(world):
    """The `config` of the main line — identical in both arms (`--impl ours` and `--impl reference`)."""
    return {'workload': 'simple_nerf_frame_render', 'frame': [FRAME_H, FRAME_W], 'rays_per_step_per_gpu': FRAME_H * FRAME_W,
            'samples': f'{S_COARSE} coarse + {N_FINE} fine (fine pass evaluates {S_COARSE + N_FINE})', 'views': 2, 'ndc': True,
            'weights': 'random-init', 'parallelism': f'ray-sharded x{world} (one frame per rank)',
            'l2': 'per-step working set (~3 GB of per-sample intermediates) exceeds the 126 MB L2; no flush needed'}


# ------------------------------------------------------------------------------------------------ plumbing
class Dist:
    def __init__(self, rank, world, device):
        self.rank, self.world, self.device = rank, world, device

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_ms(self, ms):
        if self.world > 1:
            t = torch.tensor([ms], device=self.device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = t.item()
        return ms


def timed_steps(dist, fn, steps, warmup, collect=False):
    """W untimed + K timed calls of fn(step) between barrier + synchronize; returns (max-over-ranks ms for the K steps,
    per-launch timing records of this rank, launches of this rank's kernels inside the timed region)."""
    from simple_rf_b200 import _lib
    for s in range(warmup):
        fn(s)
    dist.barrier()
    _lib.LAUNCHES.clear()
    _lib.TIMING = [] if collect else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        fn(warmup + s)
    e1.record()
    dist.barrier()
    timing, _lib.TIMING = _lib.TIMING, None
    return dist.max_ms(e0.elapsed_time(e1)), timing, sum(_lib.LAUNCHES.values())


def kernel_table(timing):
    """name -> (total ms, total algorithmic work, launches) from the per-launch CUDA events."""
    from simple_rf_b200 import _lib
    table = {}
    for name, e0, e1, work in timing or []:
        t, w, n = table.get(name, (0.0, 0.0, 0))
        table[name] = (t + e0.elapsed_time(e1), w + _lib.resolve_work(work), n + 1)
    return table


def roofline_of(table, names, bound, peak, unit, step_ms_total, peak_source, kernel_label, traffic=None):
    ms = sum(table[n][0] for n in names if n in table)
    work = sum(table[n][1] for n in names if n in table)
    launches = sum(table[n][2] for n in names if n in table)
    scale = 1e12 if unit == 'TFLOP/s' else 1e9
    achieved = work / (ms * 1e-3) / scale if ms > 0 else 0.0
    return {'kernel': kernel_label, 'bound': bound, 'achieved': achieved, 'peak': peak, 'unit': unit, 'frac': achieved / peak,
            'traffic': traffic, 'peak_source': peak_source, 'kernel_share_of_step': ms / step_ms_total if step_ms_total else None,
            'launches': launches}


def traffic_of(kernel):
    tf = ROOT / 'profiles' / 'roofline_traffic.json'
    if tf.exists():
        return json.loads(tf.read_text()).get(kernel, {}).get('dram_bytes_per_launch')
    return None


def reference_available():
    from simple_rf_b200.dropin import callers as C
    return C.available()


# ------------------------------------------------------------------------------------------------ main workload
def build_render_setup(device, local):
    """-> dict(model, frame(step) -> pose install, tester or None, raw, model_configs)."""
    from simple_rf_b200 import synthetic
    if reference_available():
        from simple_rf_b200.dropin import callers as C
        C.prepare()
        cfg = C.use_dropin(C.complete_configs(C.load_shipped_configs(1061), [local], seed=0))
        raw = C.synthetic_raw_data('llff', 2, sparse_points=2000, seed=0)
        from data_preprocessors.DataPreprocessorFactory01 import get_data_preprocessor
        import Trainer10
        Trainer10.init_seeds(0)
        pre = get_data_preprocessor(cfg, mode='train', raw_data_dict=copy.deepcopy(raw))
        mc = pre.get_model_configs()
        del pre
        tester = C.make_tester(cfg, mc, [local])
        tester.model.eval()
        # main line = one INDEPENDENT frame per rank (weak scaling): the drop-in's test-time row-band sharding (on by default under
        # torch.distributed, for an unedited Tester07 on several ranks) is measured separately by sharded_frame()
        tester.model.module.configs['model']['shard_eval_rays'] = False
        return {'model': tester.model.module, 'tester': tester, 'raw': raw, 'model_configs': mc, 'pose': lambda t: C.test_pose(raw, t)}
    from simple_rf_b200.models.SimpleNeRF91 import SimpleNeRF
    configs = synthetic.nerf_configs()
    mc = synthetic.scene_model_configs('llff', num_views=2)
    torch.manual_seed(0)
    model = SimpleNeRF(configs, mc).to(device).eval()
    model.configs['model']['shard_eval_rays'] = False
    return {'model': model, 'tester': None, 'raw': None, 'model_configs': mc,
            'pose': lambda t: np.asarray(synthetic.trajectory_pose(mc, t), dtype=np.float32)}


def run_main(args, dist):
    from simple_rf_b200 import synthetic
    device = dist.device
    setup = build_render_setup(device, device.index)
    model, tester, mc = setup['model'], setup['tester'], setup['model_configs']
    h, w = mc['resolution']
    assert (h, w) == (FRAME_H, FRAME_W)
    pid_host = torch.from_numpy(synthetic.frame_pixel_ids(h, w, view=0)).pin_memory()
    pid_dev = pid_host.to(device)
    num_rays = pid_host.shape[0]
    k = np.asarray(mc['intrinsics'][:1], dtype=np.float32)

    def pose_of(step):
        return setup['pose'](((step * 8 + dist.rank) % 120) / 120.0)

    def processed_pose(step):
        if tester is None:
            return np.asarray(pose_of(step), dtype=np.float32)
        return tester.data_preprocessor.preprocess_poses(
            {'poses': pose_of(step)[None].copy(), 'translation_scale': mc['translation_scale'], 'average_pose': np.array(mc['average_pose'])},
            train_mode=False)['poses'][0]

    poses = {}

    def render_device(step):
        p = poses.get(step)
        if p is None:
            p = poses[step] = processed_pose(step)
        model.rebuild_camera_params_learners(intrinsics=k, extrinsics=p[None], device=device)      # Tester07.py:164
        with torch.no_grad():
            return model({'pixel_id': pid_dev, 'num_frames': 1})

    out_rgb = torch.empty((num_rays, 3), dtype=torch.float32).pin_memory()
    out_depth = torch.empty((num_rays,), dtype=torch.float32).pin_memory()
    e2e_bytes = {}

    def render_e2e(step):
        if tester is not None:
            frame = tester.predict_frame(pose_of(step))                                             # the unmodified caller
            e2e_bytes['d2h'] = int(sum(v.nbytes for v in frame.values()))
            return frame
        model.rebuild_camera_params_learners(intrinsics=k, extrinsics=processed_pose(step)[None], device=device)
        with torch.no_grad():
            out = model({'pixel_id': pid_host.to(device, non_blocking=True), 'num_frames': 1})
        out_rgb.copy_(out['rgb_fine'], non_blocking=True)
        out_depth.copy_(out['depth_fine'], non_blocking=True)
        e2e_bytes['d2h'] = int(out_rgb.numel() * 4 + out_depth.numel() * 4)
        return out

    for s in range(args.warmup + args.steps):
        poses[s] = processed_pose(s)
    with ClockSampler(torch.cuda.current_device()) as clocks:
        ms_dev, timing, launches = timed_steps(dist, render_device, args.steps, args.warmup, collect=True)
    ms_e2e, _, _ = timed_steps(dist, render_e2e, args.steps, args.warmup)
    total_rays = num_rays * args.steps * dist.world
    peaks = measured_peaks()
    table = kernel_table(timing)
    line = {
        'metric': 'rendered_rays_per_sec', 'value': total_rays / (ms_dev * 1e-3), 'unit': 'rays/s', 'n_gpus': dist.world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16 (MLP operands; fp32 accumulate and everything else fp32)',
        'data': 'synthetic', 'config': main_config(dist.world),
        'e2e': {'value': total_rays / (ms_e2e * 1e-3), 'unit': 'rays/s', 'h2d_bytes_per_step': int(pid_host.numel() * 4),
                'd2h_bytes_per_step': e2e_bytes.get('d2h'),
                'through': 'Tester07.NerfTester.predict_frame (unmodified) -> SimpleNeRF91 + DataPreprocessor91' if tester is not None
                else 'SimpleNeRF91.forward with pinned host pixel ids / maps (no upstream tree installed)'},
        'gpu_launches': launches,
        'roofline': roofline_of(table, ['srf_nerf_mlp_fwd'], 'tensor', peaks['bf16_tflops_sustained'], 'TFLOP/s', ms_dev,
                                peaks['source'] + ', sustained bf16', 'nerf_mlp_fwd_kernel', traffic_of('nerf_mlp_fwd_kernel')),
        'clocks': clocks.summary(),
    }
    return line, setup


# ------------------------------------------------------------------------------------------------ training workloads
def reference_shaped_losses(kind, configs, images, resolution):
    """The shipped loss sets as one function of (batch, model output): MSE14 (:30-90) on the image rays of every model,
    SparseDepthMSE14 (:30-95) x 0.1 on the sparse-depth rays, and the two fused patch-reprojection consistency losses
    (AugmentationsDepthLoss91 / CoarseFineConsistencyLoss91) x 0.1 — no per-loss `.item()`.  TensoRF: + TV (TotalVariationLoss04)."""
    from simple_rf_b200.loss_functions.AugmentationsDepthLoss91 import AugmentationsDepthLoss
    from simple_rf_b200.loss_functions.CoarseFineConsistencyLoss91 import CoarseFineConsistencyLoss
    lcfg = {'patch_size': [5, 5], 'rmse_threshold': 0.1}
    aug_loss = AugmentationsDepthLoss(configs, lcfg)
    cf_loss = CoarseFineConsistencyLoss(configs, lcfg) if 'fine_model' in configs['model'] else None
    aug_names = [a['name'] for a in configs['model'].get('augmentations', [])]
    rgb_keys = ['rgb_coarse'] + (['rgb_fine'] if 'fine_model' in configs['model'] else []) + [f'{n}_rgb_coarse' for n in aug_names]
    depth_keys = ['depth_coarse'] + [f'{n}_depth_coarse' for n in aug_names]

    def losses(batch, out, model):
        r_img, r_sd = batch['srf_rows']['nerf'], batch['srf_rows']['sparse_depth']       # index_select: no mask-size read-back
        target = batch['target_rgb'].index_select(0, r_img)
        total = sum(torch.square(out[k].index_select(0, r_img) - target).mean() for k in rgb_keys)
        gt = batch['sparse_depth_values'][:, 0].index_select(0, r_sd)
        total = total + 0.1 * sum(torch.square(out[k].index_select(0, r_sd) - gt).mean() for k in depth_keys)
        inp = dict(batch, common_data={'images': images, 'resolution': resolution})
        total = total + 0.1 * aug_loss.compute_loss(inp, out, model)['loss_value']
        if cf_loss is not None:
            total = total + 0.1 * cf_loss.compute_loss(inp, out, model)['loss_value']
        if kind == 'tensorf':                                   # TotalVariationLoss04.py:44-78: the augmented tensors only
            from simple_rf_b200.loss_functions.TotalVariationLoss91 import tv_loss
            for a in model.augmented_models:
                t = a['coarse_model']
                total = total + 0.01 * tv_loss([*t.matrices_density, *t.matrices_color], 1.0)
        return total
    return losses


def synthetic_batch(rays_img, rays_sd, num_views, h, w, device, seed):
    g = torch.Generator().manual_seed(seed)
    n = rays_img + rays_sd
    pid = torch.stack([torch.randint(0, num_views, (n,), generator=g), torch.randint(2, w - 2, (n,), generator=g),
                       torch.randint(2, h - 2, (n,), generator=g)], 1).int()
    m_img = torch.zeros(n, dtype=torch.bool)
    m_img[:rays_img] = True
    return {'pixel_id': pid.to(device), 'target_rgb': torch.rand(n, 3, generator=g).to(device),
            'sparse_depth_values': (1.0 + 3.0 * torch.rand(n, 1, generator=g)).to(device),
            'indices_mask_nerf': m_img.to(device), 'indices_mask_sparse_depth': (~m_img).to(device), 'num_frames': num_views,
            'srf_rows': {'nerf': torch.arange(0, rays_img, device=device), 'sparse_depth': torch.arange(rays_img, n, device=device)},
            'iter_num': 0, 'sub_batch_index': 0}


def train_workload(kind, dist, scaling, steps=20, warmup=4, use_graph=True):
    """One optimiser step per STEP: forward of every model of the shipped training config + the four (five) reference-shaped
    loss terms + hand-written backward + ONE all-reduce of the flat gradient bucket + fused Adam.  weak: 2048 + 2048 rays per
    rank; strong: 2048 + 2048 rays over all ranks."""
    from simple_rf_b200 import _lib, synthetic
    device = dist.device
    if kind == 'nerf':
        from simple_rf_b200.models.SimpleNeRF91 import SimpleNeRF as Model
        cfg = synthetic.nerf_configs(rng_mode='device')
        mc = synthetic.scene_model_configs('llff', num_views=3)
    else:
        from simple_rf_b200.models.SimpleTensoRF91 import AlphaGridMask, SimpleTensoRF as Model
        cfg = synthetic.tensorf_configs(num_voxels=300 ** 3, augmentations=True, rng_mode='device')
        cfg['model']['augmentations'][0]['coarse_model']['num_voxels_initial'] = 160 ** 3
        mc = synthetic.scene_model_configs('re10k', num_views=3)
    cfg['data_loader']['sparse_depth'] = {'num_rays': 2048}
    torch.manual_seed(0)                                               # replicated parameters on every rank
    model = Model(cfg, mc).to(device).train()
    if kind == 'tensorf':
        t = model.coarse_model
        with torch.no_grad():
            for p_ in t.matrices_density:
                p_.mul_(6.0)
        vol = synthetic.blocky_alpha_volume(190, 10, 0.10, 0.004, torch.Generator().manual_seed(1))
        t.alpha_mask = AlphaGridMask(vol, t.bounding_box.cpu()).to(device)
    ocfg = cfg['optimizers'][0]
    opt = torch.optim.Adam(model.get_trainable_parameters(ocfg), betas=(ocfg['beta1'], ocfg['beta2']))
    model.optimizers = {'optimizer_nerf': opt}                         # Trainer10.py:59-62: attaches the fused Adam + all-reduce
    h, w = mc['resolution']
    per_rank = 2048 if scaling == 'weak' else 2048 // dist.world
    batch = synthetic_batch(per_rank, per_rank, 3, h, w, device, seed=100 + dist.rank)
    images = torch.rand(3, h, w, 3, generator=torch.Generator().manual_seed(5)).to(device)
    losses = reference_shaped_losses(kind, cfg, images, (h, w))

    def step(i=None):
        opt.zero_grad(set_to_none=True)
        out = model(batch)
        loss = losses(batch, out, model)
        loss.backward()
        opt.step()
        return loss

    ms_eager, _, launches = timed_steps(dist, step, steps, warmup)                # one launch per kernel: bound by the host's issue rate
    ms_prof, timing, _ = timed_steps(dist, step, steps, 1, collect=True)          # second pass: per-launch CUDA events for the rooflines
    ms, graphed, graph_error = ms_eager, False, None
    if use_graph:
        try:                                                                      # the whole iteration as ONE CUDA graph (train_graph.py)
            from simple_rf_b200.train_graph import GraphedStep
            g = GraphedStep(step, {'optimizer_nerf': opt}, warmup=2)
            ms, _, _ = timed_steps(dist, lambda i: g.replay(), steps, warmup)
            graphed = True
        except Exception as e:
            graph_error = repr(e)[:300]
    table = kernel_table(timing)
    ms_it = ms / steps
    peaks = measured_peaks()
    res = {'metric': 'train_iters_per_sec', 'value': 1e3 / ms_it, 'unit': 'it/s', 'ms_per_step': ms_it, 'steps': steps, 'warmup': warmup,
           'cuda_graph': graphed, 'ms_per_step_eager': ms_eager / steps, 'cuda_graph_error': graph_error,
           'scaling': scaling, 'n_gpus': dist.world, 'rays_per_iter_per_gpu': 2 * per_rank, 'global_rays_per_iter': 2 * per_rank * dist.world,
           'rays_per_sec': 2 * per_rank * dist.world * 1e3 / ms_it, 'gpu_launches_per_step': launches / steps,
           'losses': 'MSE + sparse-depth MSE + AugmentationsDepthLoss91 + ' + ('CoarseFineConsistencyLoss91' if kind == 'nerf' else 'TV'),
           'kernels_ms_per_step': {n: round(v[0] / steps, 4) for n, v in sorted(table.items(), key=lambda kv: -kv[1][0])}}
    fused = getattr(model, '_fused_adam', [])
    if 'nccl_all_reduce' in table:
        t_ar, b_ar, n_ar = table['nccl_all_reduce']
        res['all_reduce'] = {'bytes_per_step': b_ar / steps, 'ms_per_step': t_ar / steps, 'calls_per_step': n_ar / steps,
                             'algbw_gbs': b_ar / (t_ar * 1e-3) / 1e9 if t_ar > 0 else None,
                             'busbw_gbs': b_ar * 2 * (dist.world - 1) / dist.world / (t_ar * 1e-3) / 1e9 if t_ar > 0 else None}
    elif fused:
        res['all_reduce'] = {'bytes_per_step': fused[0].bytes_reduced_last}
    if kind == 'nerf':
        names = ['srf_nerf_mlp_fwd', 'srf_nerf_mlp_dgrad', 'srf_nerf_mlp_wgrad']
        res['config'] = {'workload': 'simple_nerf_training', 'views': 3, 'frame': [h, w], 'models': 'main coarse + fine, points + views augmentation',
                         'samples': '64 coarse + 128 fine', 'rng': 'device (Philox)'}
        res['roofline'] = roofline_of(table, names, 'tensor', peaks['bf16_tflops_sustained'], 'TFLOP/s', ms_prof, peaks['source'] + ', sustained bf16',
                                      'nerf_mlp_fwd + dgrad + wgrad kernels (1.322 GFLOP/ray algorithmic)')
        res['rooflines'] = {n: roofline_of(table, [n], 'tensor', peaks['bf16_tflops_sustained'], 'TFLOP/s', ms_prof, peaks['source'], n) for n in names}
        comp = ['srf_composite_fwd', 'srf_composite_bwd']
        res['rooflines']['composite_fwd_bwd'] = roofline_of(table, comp, 'hbm', peaks['hbm_gbs'], 'GB/s', ms_prof, peaks['source'], 'composite_fwd/bwd kernels')
    else:
        t = model.coarse_model
        res['config'] = {'workload': 'simple_tensorf_training', 'views': 3, 'frame': [h, w], 'grid': [int(v) for v in t.resolution.tolist()],
                         'samples_per_ray': int(t.num_samples), 'alpha_mask': '190^3, ~14 % occupied', 'augmentation_grid': '160^3 voxels',
                         'rng': 'device (Philox)'}
        gathers = ['srf_vm_density_fwd', 'srf_vm_density_bwd', 'srf_vm_color_features_fwd', 'srf_vm_color_features_bwd']
        res['roofline'] = roofline_of(table, gathers, 'hbm', L2_GBS, 'GB/s', ms_prof, 'requested texel bytes against the measured L2-resident read '
                                      'bandwidth (tools/l2_probe.py): the planes (<= 41 MB) stay in the 126 MB L2 (SURVEY.md §8d)',
                                      'vm_density / vm_color_features gather + scatter kernels')
        res['rooflines'] = {n: roofline_of(table, [n], 'hbm', L2_GBS, 'GB/s', ms_prof, 'L2-resident (requested texel bytes)', n) for n in gathers}
        comp = ['srf_composite_fwd', 'srf_composite_bwd']
        res['rooflines']['composite_fwd_bwd'] = roofline_of(table, comp, 'hbm', peaks['hbm_gbs'], 'GB/s', ms_prof, peaks['source'], 'composite_fwd/bwd kernels')
        res['rooflines']['mask_compaction'] = roofline_of(table, ['srf_tensorf_mask'], 'hbm', peaks['hbm_gbs'], 'GB/s', ms_prof, peaks['source'], 'tensorf_mask_kernel')
    del model, opt
    torch.cuda.empty_cache()
    return res


def tensorf_trajectory(dist, frames=6, warmup=2):
    """BASELINE.json configs[3]: Simple-TensoRF test-trajectory rendering (rgb + depth), one 576x1024 frame per rank per step."""
    from simple_rf_b200 import synthetic
    from simple_rf_b200.models.SimpleTensoRF91 import AlphaGridMask, SimpleTensoRF
    device = dist.device
    cfg = synthetic.tensorf_configs(num_voxels=300 ** 3, augmentations=False)
    mc = synthetic.scene_model_configs('re10k', num_views=3)
    torch.manual_seed(0)
    model = SimpleTensoRF(cfg, mc).to(device).eval()
    model.configs['model']['shard_eval_rays'] = False        # the ranks render different frames of the trajectory (weak scaling)
    t = model.coarse_model
    with torch.no_grad():
        for p_ in t.matrices_density:
            p_.mul_(6.0)
    t.alpha_mask = AlphaGridMask(synthetic.blocky_alpha_volume(190, 10, 0.10, 0.004, torch.Generator().manual_seed(1)), t.bounding_box.cpu()).to(device)
    h, w = mc['resolution']
    pid = torch.from_numpy(synthetic.frame_pixel_ids(h, w, view=0)).to(device)
    k = np.asarray(mc['intrinsics'][:1], dtype=np.float32)

    def render(step):
        pose = np.asarray(synthetic.trajectory_pose(mc, ((step * 8 + dist.rank) % 120) / 120.0), dtype=np.float32)
        model.rebuild_camera_params_learners(intrinsics=k, extrinsics=pose[None], device=device)
        with torch.no_grad():
            model({'pixel_id': pid, 'num_frames': 1})

    ms, _, launches = timed_steps(dist, render, frames, warmup)
    ms_prof, timing, _ = timed_steps(dist, render, frames, 1, collect=True)
    table = kernel_table(timing)
    peaks = measured_peaks()
    ms_f = ms / frames
    rays = pid.shape[0]
    S = int(t.num_samples)
    gathers = ['srf_vm_density_fwd', 'srf_vm_color_features_fwd']
    res = {'metric': 'rendered_rays_per_sec', 'value': rays * dist.world * 1e3 / ms_f, 'unit': 'rays/s', 'frames_per_sec': dist.world * 1e3 / ms_f,
           'ms_per_step': ms_f, 'steps': frames, 'warmup': warmup, 'scaling': 'weak', 'n_gpus': dist.world, 'gpu_launches_per_step': launches / frames,
           'config': {'workload': 'simple_tensorf_trajectory_render', 'frame': [h, w], 'grid': [int(v) for v in t.resolution.tolist()],
                      'samples_per_ray': S, 'alpha_mask': '190^3, ~14 % occupied', 'outputs': 'rgb + depth'},
           'kernels_ms_per_step': {n: round(v[0] / frames, 4) for n, v in sorted(table.items(), key=lambda kv: -kv[1][0])},
           # the kernel of the frame that has a hardware peak to be held against: the colour MLP on the tensor cores (the march is
           # issue-bound, and the colour gather re-uses texels in registers along runs of consecutive samples, so its requested texel
           # bytes per second are no longer bounded by the L2 figure - see `rooflines`)
           'roofline': roofline_of(table, ['srf_mlp_rows_fwd'], 'tensor', peaks['bf16_tflops_sustained'], 'TFLOP/s', ms_prof,
                                   peaks['source'] + ', sustained bf16', 'nerf_mlp_fwd_kernel (rows mode)'),
           # compulsory HBM bytes of the whole frame (SURVEY.md §8d i): per sample 4 B depth in; per ray 72 B rays + 28 B maps out —
           # the fused march never materialises the [R,S] depths, so this is a bound nobody is near: reported for the record
           'compulsory_hbm': {'bytes_per_frame': rays * S * 4.0 + rays * 100.0, 'GB/s': (rays * S * 4.0 + rays * 100.0) / (ms_f * 1e-3) / 1e9,
                              'frac_of_hbm_peak': (rays * S * 4.0 + rays * 100.0) / (ms_f * 1e-3) / 1e9 / peaks['hbm_gbs']},
           'rooflines': {n: roofline_of(table, [n], 'hbm', L2_GBS, 'GB/s', ms_prof, 'L2-resident (requested texel bytes)', n) for n in gathers}}
    res['rooflines']['mlp_rows_fwd'] = roofline_of(table, ['srf_mlp_rows_fwd'], 'tensor', peaks['bf16_tflops_sustained'], 'TFLOP/s', ms_prof,
                                                   peaks['source'] + ', sustained bf16', 'nerf_mlp_fwd_kernel (rows mode)')
    res['rooflines']['srf_vm_color_features_fwd']['note'] = ('requested bytes (1728 B per surface sample); the run-merged gather re-loads a texel only when '
                                                             'the footprint of consecutive samples changes, so this rate may exceed the L2 figure')
    del model
    torch.cuda.empty_cache()
    return res


def sharded_frame(dist, setup, steps=4, warmup=2):
    """Strong scaling of ONE Simple-NeRF frame: the drop-in's render() takes this rank's row band and all-gathers the per-ray
    maps (the only collective of the test path; SURVEY.md §8e iii)."""
    from simple_rf_b200 import synthetic
    model, mc = setup['model'], setup['model_configs']
    device = dist.device
    pid = torch.from_numpy(synthetic.frame_pixel_ids(*mc['resolution'], view=0)).to(device)
    model.configs['model']['shard_eval_rays'] = True

    def render(step):
        with torch.no_grad():
            model({'pixel_id': pid, 'num_frames': 1})
    try:
        ms, _, _ = timed_steps(dist, render, steps, warmup)
    finally:
        model.configs['model']['shard_eval_rays'] = False
    return {'metric': 'rendered_rays_per_sec', 'value': pid.shape[0] * 1e3 / (ms / steps), 'unit': 'rays/s', 'ms_per_step': ms / steps,
            'scaling': 'strong', 'n_gpus': dist.world, 'steps': steps, 'warmup': warmup,
            'config': {'workload': 'simple_nerf_frame_render_sharded', 'frame': mc['resolution'], 'collective': 'one all_gather_into_tensor of the per-ray record per frame'}}


def split_precision_frame(dist, setup, steps=2, warmup=1):
    """The main workload with `mlp_precision = 'bf16x3'` (split-bf16 operands, three MMAs per K block): the program that meets
    the 1e-3 fp32 contract on trained fields.  One independent frame per rank, as the main line."""
    from simple_rf_b200 import synthetic
    model, mc = setup['model'], setup['model_configs']
    pid = torch.from_numpy(synthetic.frame_pixel_ids(*mc['resolution'], view=0)).to(dist.device)
    model.configs['model']['mlp_precision'] = 'bf16x3'

    def render(step):
        with torch.no_grad():
            model({'pixel_id': pid, 'num_frames': 1})
    try:
        ms, timing, launches = timed_steps(dist, render, steps, warmup, collect=True)
    finally:
        model.configs['model']['mlp_precision'] = 'bf16'
    peaks = measured_peaks()
    res = {'metric': 'rendered_rays_per_sec', 'value': pid.shape[0] * dist.world * 1e3 / (ms / steps), 'unit': 'rays/s', 'ms_per_step': ms / steps,
           'scaling': 'weak', 'n_gpus': dist.world, 'steps': steps, 'warmup': warmup, 'dtype': 'split bf16 (hi + lo operand pairs, fp32 accumulate)',
           'config': {'workload': 'simple_nerf_frame_render', 'mlp_precision': 'bf16x3', 'frame': mc['resolution']},
           'roofline': roofline_of(kernel_table(timing), ['srf_nerf_mlp_fwd'], 'tensor', peaks['bf16_tflops_sustained'], 'TFLOP/s', ms,
                                   peaks['source'] + ', sustained bf16; algorithmic FLOPs (the program issues 3x the MMAs)', 'nerf_mlp_fwd_kernel<split>', None)}
    return res


# ------------------------------------------------------------------------------------------------ reference (CPU / eager CUDA)
def reference_render_setup(device_ids, chunk):
    """The UNMODIFIED reference classes for the main workload: NerfTester (src/Tester07.py:30-48) with SimpleNeRF17.  Call under
    `callers.force_cpu()` for the host-CPU arm (the reference picks its device with torch.cuda.is_available())."""
    from simple_rf_b200.dropin import callers as C
    C.prepare()
    cfg = C.complete_configs(C.load_shipped_configs(1061), device_ids, seed=0)
    cfg['model']['chunk'] = chunk
    raw = C.synthetic_raw_data('llff', 2, sparse_points=2000, seed=0)
    from data_preprocessors.DataPreprocessorFactory01 import get_data_preprocessor
    import Trainer10
    Trainer10.init_seeds(0)
    with C.force_cpu():                                       # the train-mode caches are only needed for the model configs: keep them on the host
        pre = get_data_preprocessor(cfg, mode='train', raw_data_dict=copy.deepcopy(raw))
    mc = pre.get_model_configs()
    del pre
    tester = C.make_tester(cfg, mc, device_ids)
    tester.model.eval()
    return tester, raw, mc, C


def reference_render_rays(tester, raw, C, num_rays, step):
    """`num_rays` rays of the frame at pose `step` through the reference's own path: create_test_data -> rebuild cameras ->
    model forward on the row-major prefix of the frame (what predict_frame does, src/Tester07.py:156-171, on a bounded sample)."""
    inp = tester.data_preprocessor.create_test_data(pose=C.test_pose(raw, (step % 120) / 120.0))
    stride = max(1, inp['pixel_id'].shape[0] // num_rays)
    inp['pixel_id'] = inp['pixel_id'][::stride][:num_rays].contiguous()
    pose = inp['common_data']['processed_pose'][0].cpu().numpy()
    k = inp['common_data']['intrinsic'][0].cpu().numpy()
    tester.model.module.rebuild_camera_params_learners(intrinsics=k[None], extrinsics=pose[None], device=tester.device)
    with torch.no_grad():
        out = tester.model(inp)
    return int(inp['pixel_id'].shape[0]), out


def cpu_baseline_main(budget_s):
    """Reference CPU arm of the main workload on a bounded sample (1024-ray batches: BASELINE.json configs[0])."""
    torch.set_num_threads(os.cpu_count() or 1)
    if reference_available():
        from simple_rf_b200.dropin import callers
        with callers.force_cpu():
            tester, raw, mc, C = reference_render_setup([0], chunk=1024)
            reference_render_rays(tester, raw, C, 1024, 0)
            t0 = time.perf_counter()
            n, _ = reference_render_rays(tester, raw, C, 1024, 1)
            dt = time.perf_counter() - t0
            rays = int(min(65536, max(1024, (budget_s / max(dt, 1e-3)) * 1024 // 1024 * 1024)))
            t0 = time.perf_counter()
            n, _ = reference_render_rays(tester, raw, C, rays, 2)
            dt = time.perf_counter() - t0
        return {'value': n / dt, 'unit': 'rays/s', 'cores': torch.get_num_threads(), 'kind': 'reference',
                'sample': f'{n} rays of the same frame in 1024-ray batches through the unmodified SimpleNeRF17 (baseline/_ref), {dt:.1f} s'}
    return cpu_baseline_port(budget_s)


def cpu_baseline_port(budget_s):
    from oracle import nerf_mlp as M
    from oracle import pipeline as P
    from simple_rf_b200 import synthetic
    configs = synthetic.nerf_configs(augmentations=False)
    model_configs = synthetic.scene_model_configs('llff', num_views=2)
    g = torch.Generator().manual_seed(0)
    sets = {'coarse_model': M.init_mlp_params(configs['model']['coarse_model'], g), 'fine_model': M.init_mlp_params(configs['model']['fine_model'], g)}

    def sample(num_rays):
        pid = torch.from_numpy(synthetic.frame_pixel_ids(FRAME_H, FRAME_W, view=0)[::max(1, FRAME_H * FRAME_W // num_rays)][:num_rays])
        t0 = time.perf_counter()
        with torch.no_grad():
            for i in range(0, pid.shape[0], 1024):
                P.nerf_render_chunk(sets, configs, model_configs, pid[i:i + 1024], training=False, retraw=False)
        return pid.shape[0], time.perf_counter() - t0
    sample(1024)
    n, dt = sample(1024)
    n, dt = sample(int(min(65536, max(1024, (budget_s / max(dt, 1e-3)) * 1024 // 1024 * 1024))))
    return {'value': n / dt, 'unit': 'rays/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{n} rays of the same frame in 1024-ray batches, oracle/pipeline.py (no upstream tree installed), {dt:.1f} s'}


def reference_train_iteration(kind, on_cpu, rays, iters):
    """Seconds per UNMODIFIED `Trainer.train_one_iter` (src/Trainer10.py:65) with the reference's own classes on the host CPU or
    on the current GPU (eager PyTorch), shipped config, `rays` image rays + `rays` sparse-depth rays per iteration."""
    from simple_rf_b200.dropin import callers as C
    if on_cpu:
        with C.force_cpu():
            return _reference_train_iteration(kind, [0], rays, iters, False)
    return _reference_train_iteration(kind, [torch.cuda.current_device()], rays, iters, True)


def _reference_train_iteration(kind, device_ids, rays, iters, sync):
    from simple_rf_b200.dropin import callers as C
    C.prepare()
    if kind == 'nerf':
        cfg = C.complete_configs(C.load_shipped_configs(1142), device_ids, seed=0)
        raw = C.synthetic_raw_data('llff', 3, resolution=(378, 504), sparse_points=2000, seed=0)
    else:
        cfg = C.complete_configs(C.load_shipped_configs(212), device_ids, seed=0)
        cfg['model']['coarse_model']['num_voxels_initial'] = 300 ** 3
        cfg['model']['coarse_model']['num_voxels_final'] = 300 ** 3
        cfg['model']['augmentations'][0]['coarse_model']['num_voxels_initial'] = 160 ** 3
        cfg['model']['augmentations'][0]['coarse_model']['num_voxels_final'] = 160 ** 3
        raw = C.synthetic_raw_data('re10k', 3, resolution=(288, 512), sparse_points=2000, seed=0, tensorf=True)
    cfg['data_loader']['num_rays'] = rays
    cfg['data_loader']['sparse_depth']['num_rays'] = rays
    for loss in cfg['losses']:
        if 'iter_weights' in loss:
            loss['iter_weights'] = {'0': 0.1}
    trainer, model, mc = C.make_trainer(cfg, raw, seed=0)
    trainer.train_one_iter(0)
    if sync:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(iters):
        trainer.train_one_iter(1 + it)
    if sync:
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters


def training_baselines(kind, on_gpu):
    """cpu_baseline (and, on one GPU, the eager-CUDA reference bar) of a training workload, scaled to 4096-ray iterations."""
    out = {}
    if not reference_available():
        return out
    try:
        rays = 512
        s = reference_train_iteration(kind, True, rays, 1)
        out['cpu_baseline'] = {'value': (2 * rays / 4096.0) / s, 'unit': 'it/s', 'cores': torch.get_num_threads(), 'kind': 'reference',
                               'sample': f'one unmodified Trainer.train_one_iter (src/Trainer10.py:65) with {rays} + {rays} rays on the host CPU '
                                         f'({s:.1f} s), scaled to the 2048 + 2048-ray iteration'}
        if on_gpu:
            s = reference_train_iteration(kind, False, 2048, 3)
            out['gpu_reference_bar'] = {'value': 1.0 / s, 'unit': 'it/s', 'kind': 'reference eager PyTorch on the same B200',
                                        'sample': 'unmodified Trainer.train_one_iter, 2048 + 2048 rays, mean of 3 iterations'}
    except Exception as e:                                     # informational: never take the contract line down
        out['baseline_error'] = repr(e)[:300]
    return out


def dropin_trainer_e2e(kind, device, iters=10):
    """it/s of the UNMODIFIED Trainer.train_one_iter driving the drop-in classes (what a user of the reference gets): host index
    selection, fused batch assembly, forward, the shipped loss classes with their per-loss `.item()`, backward, fused Adam."""
    from simple_rf_b200.dropin import callers as C
    C.prepare()
    if kind == 'nerf':
        cfg = C.use_dropin(C.complete_configs(C.load_shipped_configs(1142), [device.index], seed=0))
        raw = C.synthetic_raw_data('llff', 3, resolution=(378, 504), sparse_points=2000, seed=0)
    else:
        cfg = C.use_dropin(C.complete_configs(C.load_shipped_configs(212), [device.index], seed=0))
        cfg['model']['coarse_model']['num_voxels_initial'] = 300 ** 3
        cfg['model']['coarse_model']['num_voxels_final'] = 300 ** 3
        cfg['model']['augmentations'][0]['coarse_model']['num_voxels_initial'] = 160 ** 3
        cfg['model']['augmentations'][0]['coarse_model']['num_voxels_final'] = 160 ** 3
        raw = C.synthetic_raw_data('re10k', 3, resolution=(288, 512), sparse_points=2000, seed=0, tensorf=True)
    cfg['model']['rng_mode'] = 'device'
    for loss in cfg['losses']:
        if 'iter_weights' in loss:
            loss['iter_weights'] = {'0': 0.1}
    trainer, model, mc = C.make_trainer(cfg, raw, seed=0)
    for it in range(3):
        trainer.train_one_iter(it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(iters):
        trainer.train_one_iter(3 + it)
    torch.cuda.synchronize()
    s = (time.perf_counter() - t0) / iters
    return {'value': 1.0 / s, 'unit': 'it/s', 'ms_per_step': s * 1e3,
            'through': 'Trainer10.Trainer.train_one_iter (unmodified) -> drop-in model + DataPreprocessor91 + *Loss91, 2048 + 2048 rays'}


# ------------------------------------------------------------------------------------------------ arms
def run_ours(args, rank, world, device):
    dist = Dist(rank, world, device)
    line, setup = run_main(args, dist)
    if not args.main_only:
        wl = {}

        def guarded(name, fn):
            try:
                wl[name] = fn()
            except Exception as e:                            # a secondary measurement must never take the contract line down
                import traceback
                wl[name] = {'error': repr(e)[:300], 'trace': traceback.format_exc()[-600:]}
            dist.barrier()
        guarded('simple_nerf_train_weak', lambda: train_workload('nerf', dist, 'weak'))
        if world > 1:
            guarded('simple_nerf_train_strong', lambda: train_workload('nerf', dist, 'strong'))
        guarded('simple_tensorf_train_weak', lambda: train_workload('tensorf', dist, 'weak'))
        guarded('simple_tensorf_trajectory', lambda: tensorf_trajectory(dist))
        guarded('simple_nerf_frame_bf16x3', lambda: split_precision_frame(dist, setup))
        if world > 1:
            guarded('simple_nerf_frame_sharded', lambda: sharded_frame(dist, setup))
        if world == 1 and rank == 0 and reference_available():
            for kind, key in (('nerf', 'simple_nerf_train_weak'), ('tensorf', 'simple_tensorf_train_weak')):
                if 'error' not in wl[key]:
                    try:
                        wl[key]['e2e'] = dropin_trainer_e2e(kind, device)
                    except Exception as e:
                        wl[key]['e2e'] = {'error': repr(e)[:300]}
                    wl[key].update(training_baselines(kind, on_gpu=True))
        line['workloads'] = wl
    if rank == 0:
        if world == 1:
            if reference_available():
                try:
                    tester, raw, mc, C = reference_render_setup([device.index], chunk=4096)
                    reference_render_rays(tester, raw, C, 16384, 0)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    n, _ = reference_render_rays(tester, raw, C, 65536, 1)
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    line['gpu_reference_bar'] = {'value': n / dt, 'unit': 'rays/s', 'kind': 'reference eager PyTorch on the same B200',
                                                 'sample': f'{n} rays of the same frame through the unmodified SimpleNeRF17 on cuda (chunk 4096), {dt:.2f} s'}
                    del tester
                except Exception as e:
                    line['gpu_reference_bar'] = {'error': repr(e)[:300]}
            line['cpu_baseline'] = cpu_baseline_main(budget_s=12.0)
        print(json.dumps(line), flush=True)


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the main workload on the host cores (rank 0 only)."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    rays_per_step = 16384        # ~3 s of host work per step: the per-step frame set-up (create_test_data of 762 048 pixel ids) stays < 3 %
    if reference_available():
        from simple_rf_b200.dropin import callers
        callers.force_cpu().__enter__()                       # for the rest of this process: the reference arm is the host-CPU arm
        tester, raw, mc, C = reference_render_setup([0], chunk=1024)
        kind, what = 'reference', 'the unmodified SimpleNeRF17 through NerfTester (baseline/_ref)'

        def step(i):
            return reference_render_rays(tester, raw, C, rays_per_step, i)[0]
    else:
        from oracle import nerf_mlp as M
        from oracle import pipeline as P
        from simple_rf_b200 import synthetic
        configs = synthetic.nerf_configs(augmentations=False)
        model_configs = synthetic.scene_model_configs('llff', num_views=2)
        g = torch.Generator().manual_seed(0)
        sets = {'coarse_model': M.init_mlp_params(configs['model']['coarse_model'], g), 'fine_model': M.init_mlp_params(configs['model']['fine_model'], g)}
        pid = torch.from_numpy(synthetic.frame_pixel_ids(FRAME_H, FRAME_W, view=0)[::FRAME_H * FRAME_W // rays_per_step][:rays_per_step])
        kind, what = 'port', 'oracle/pipeline.py (no upstream tree installed)'

        def step(i):
            with torch.no_grad():
                for j in range(0, pid.shape[0], 1024):
                    P.nerf_render_chunk(sets, configs, model_configs, pid[j:j + 1024], training=False, retraw=False)
            return pid.shape[0]
    for i in range(max(1, min(args.warmup, 2))):
        step(i)
    t0 = time.perf_counter()
    n = 0
    for i in range(args.steps):
        n += step(args.warmup + i)
    dt = time.perf_counter() - t0
    v = n / dt
    cb = {'value': v, 'unit': 'rays/s', 'cores': torch.get_num_threads(), 'kind': kind,
          'sample': f'{rays_per_step} rays of the 762 048-ray frame per step (a rate: rays/s does not depend on the sample size) in 1024-ray '
                    f'batches, {what}'}
    print(json.dumps({
        'impl': 'reference', 'metric': 'rendered_rays_per_sec', 'value': v, 'unit': 'rays/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic', 'config': main_config(world),
        'cpu_baseline': cb, 'e2e': {'value': v, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--main-only', action='store_true', help='skip the secondary workloads (training, TensoRF, sharded frame)')
    ap.add_argument('--no-extras', dest='main_only', action='store_true', help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        sys.exit('bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)')
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.distributed.init_process_group('nccl', device_id=device)
    from simple_rf_b200 import build
    if not build.LIB.exists():
        build.build()
    run_ours(args, rank, world, device)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
