#!/usr/bin/env python
"""Benchmark of the Simple-RF per-ray rendering hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload nerf_render]

Workload `nerf_render` (BASELINE.json configs[0]/[1] shape): Simple-NeRF, synthetic LLFF-shaped scene
(3 views, 1008x756, focal 815.13, NDC), 64 coarse + 128 fine samples, default-initialised weights; one
STEP renders one full 762 048-ray frame from a new pose through the drop-in model's public forward().
Multi-GPU: one process per GPU (torchrun), every rank renders its own frame per step — independent units,
no data-path collective — so scaling is weak and `value` is all ranks' rays over the max-over-ranks time.

`value`  : device-resident inputs (pixel ids already in HBM), CUDA-event timed.
`e2e`    : same call with pixel ids in pinned HOST memory (H2D inside the timed region) and the rendered
           rgb + depth maps copied back to the host (D2H) every step.
`roofline`: fused tcgen05 MLP kernel — algorithmic FLOPs (unpadded MACs x 2) per launch over
           the CUDA-event duration of each launch, against the measured bf16 peak in MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the CPU oracle port of the reference algorithm (oracle/pipeline.py, same
           ATen CPU kernels the reference runs) on a bounded ray sample with all host threads.
"""
import argparse
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


def measured_peaks():
    f = ROOT / 'MEASURED_PEAKS.json'
    if f.exists():
        d = json.loads(f.read_text())
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'], 'bf16_tflops_sustained': d['bf16_tflops_sustained'],
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


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


def build_model(device):
    from simple_rf_b200 import synthetic
    from simple_rf_b200.models.SimpleNeRF91 import SimpleNeRF
    configs = synthetic.nerf_configs()
    model_configs = synthetic.scene_model_configs('llff', num_views=3)
    torch.manual_seed(0)
    model = SimpleNeRF(configs, model_configs).to(device).eval()
    return model, configs, model_configs


def frame_pose(model, model_configs, step, rank, device):
    """New test pose per (step, rank): installed the way Tester07.predict_frame does (:164)."""
    from simple_rf_b200 import synthetic
    pose = synthetic.trajectory_pose(model_configs, ((step * 8 + rank) % 120) / 120.0)
    k = np.asarray(model_configs['intrinsics'][:1], dtype=np.float32)
    model.rebuild_camera_params_learners(intrinsics=k, extrinsics=np.asarray(pose, dtype=np.float32)[None], device=device)


def run_ours(args, rank, world, device):
    from simple_rf_b200 import _lib, synthetic
    model, configs, model_configs = build_model(device)
    pid_host = torch.from_numpy(synthetic.frame_pixel_ids(FRAME_H, FRAME_W, view=0)).pin_memory()
    pid_dev = pid_host.to(device)
    num_rays = pid_host.shape[0]
    out_rgb = torch.empty((num_rays, 3), dtype=torch.float32).pin_memory()
    out_depth = torch.empty((num_rays,), dtype=torch.float32).pin_memory()

    def render(step, host_io):
        frame_pose(model, model_configs, step, rank, device)
        pid = pid_host.to(device, non_blocking=True) if host_io else pid_dev
        with torch.no_grad():
            out = model({'pixel_id': pid, 'num_frames': 1})
        if host_io:
            out_rgb.copy_(out['rgb_fine'], non_blocking=True)
            out_depth.copy_(out['depth_fine'], non_blocking=True)
        return out

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(host_io, collect):
        for s in range(args.warmup):
            render(s, host_io)
        barrier()
        _lib.LAUNCHES.clear()
        _lib.TIMING = [] if collect else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(args.steps):
            render(args.warmup + s, host_io)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        timing, _lib.TIMING = _lib.TIMING, None
        if world > 1:
            t = torch.tensor([ms], device=device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = t.item()
        return ms, timing, sum(_lib.LAUNCHES.values())

    with ClockSampler(torch.cuda.current_device()) as clocks:
        ms_dev, timing, launches = timed(host_io=False, collect=True)
    ms_e2e, _, _ = timed(host_io=True, collect=False)
    total_rays = num_rays * args.steps * world
    peaks = measured_peaks()
    mlp = [(e0.elapsed_time(e1), w) for name, e0, e1, w in timing if name == 'srf_nerf_mlp_fwd']
    mlp_ms = sum(t for t, _ in mlp)
    mlp_flops = sum(w for _, w in mlp)
    achieved = mlp_flops / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
    peak = peaks['bf16_tflops_sustained']
    traffic = None
    tf = ROOT / 'profiles' / 'roofline_traffic.json'
    if tf.exists():
        traffic = json.loads(tf.read_text()).get('nerf_mlp_fwd_kernel', {}).get('dram_bytes_per_launch')
    line = {
        'metric': 'rendered_rays_per_sec', 'value': total_rays / (ms_dev * 1e-3), 'unit': 'rays/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16 (MLP operands; fp32 accumulate and everything else fp32)',
        'data': 'synthetic',
        'config': {'workload': 'simple_nerf_frame_render', 'frame': [FRAME_H, FRAME_W], 'rays_per_step_per_gpu': num_rays,
                   'samples': f'{S_COARSE} coarse + {N_FINE} fine (fine pass evaluates {S_COARSE + N_FINE})',
                   'views': 3, 'ndc': True, 'weights': 'random-init', 'parallelism': f'ray-sharded x{world} (one frame per rank)',
                   'l2': 'per-step working set (~3 GB of per-sample intermediates) exceeds the 126 MB L2; no flush needed'},
        'e2e': {'value': total_rays / (ms_e2e * 1e-3), 'unit': 'rays/s', 'h2d_bytes_per_step': int(pid_host.numel() * 4),
                'd2h_bytes_per_step': int(out_rgb.numel() * 4 + out_depth.numel() * 4)},
        'gpu_launches': launches,
        'roofline': {'kernel': 'nerf_mlp_fwd_kernel', 'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                     'frac': achieved / peak, 'traffic': traffic, 'peak_source': peaks['source'] + ', sustained bf16',
                     'kernel_share_of_step': mlp_ms / ms_dev, 'launches': len(mlp)},
        'clocks': clocks.summary(),
    }
    if world == 1 and not args.no_extras:
        line['extras'] = extras(device)
    if rank == 0:
        if world == 1:
            line['cpu_baseline'] = cpu_baseline(budget_s=12.0)
        print(json.dumps(line), flush=True)


def _time_steps(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def extras(device):
    """Secondary, informational measurements of the other BASELINE.json configs (single GPU, device-timed):
    Simple-TensoRF frame render (configs[3] shape), one Simple-NeRF training iteration (configs[1] shape: 4096 rays,
    main coarse + fine + both augmented MLPs, forward + hand-written tcgen05 backward + Adam step) and one Simple-TensoRF
    training iteration (configs[2] shape)."""
    from simple_rf_b200 import synthetic
    from simple_rf_b200.models.SimpleNeRF91 import SimpleNeRF
    from simple_rf_b200.models.SimpleTensoRF91 import AlphaGridMask, SimpleTensoRF
    out = {}
    # ---- Simple-TensoRF: 576x1024 frame, 300^3-class grid (331x368x220 -> 1083 samples/ray), 5 % occupancy mask
    try:
        cfg = synthetic.tensorf_configs(num_voxels=300 ** 3, augmentations=False)
        mc = synthetic.scene_model_configs('re10k', num_views=3)
        torch.manual_seed(0)
        model = SimpleTensoRF(cfg, mc).to(device).eval()
        t = model.coarse_model
        for p_ in t.matrices_density:
            p_.data.mul_(6.0)
        g = torch.Generator().manual_seed(1)
        vol = (torch.rand(190, 190, 190, generator=g) < 0.05).float()
        t.alpha_mask = AlphaGridMask(vol, t.bounding_box.cpu()).to(device)
        h, w = mc['resolution']
        pid = torch.from_numpy(synthetic.frame_pixel_ids(h, w, view=0)).to(device)

        def render():
            with torch.no_grad():
                model({'pixel_id': pid, 'num_frames': 3})
        ms = _time_steps(render, steps=2, warmup=1)
        out['simple_tensorf_frame_render'] = {'rays_per_sec': pid.shape[0] / (ms * 1e-3), 'ms_per_frame': ms, 'frame': [h, w],
                                              'grid': [int(v) for v in t.resolution.tolist()], 'samples_per_ray': int(t.num_samples),
                                              'alpha_mask_occupancy': 0.05}
        del model, pid
    except Exception as e:                                    # informational: never take the contract line down
        out['simple_tensorf_frame_render'] = {'error': repr(e)[:200]}
    # ---- Simple-NeRF training iteration: 4096 rays, main coarse+fine + both augmentations, Adam step
    try:
        cfg = synthetic.nerf_configs(rng_mode='device')
        mc = synthetic.scene_model_configs('llff', num_views=3)
        torch.manual_seed(0)
        model = SimpleNeRF(cfg, mc).to(device).train()
        opt = torch.optim.Adam(model.get_trainable_parameters(cfg['optimizers'][0]), betas=(0.9, 0.999))
        model.optimizers = {'optimizer_nerf': opt}        # what Trainer10.py:59-62 does: the drop-in attaches its fused Adam step
        g = torch.Generator().manual_seed(2)
        pid = torch.stack([torch.randint(0, 3, (4096,), generator=g), torch.randint(0, FRAME_W, (4096,), generator=g),
                           torch.randint(0, FRAME_H, (4096,), generator=g)], 1).int().to(device)
        target = torch.rand(4096, 3, device=device)

        def step():
            opt.zero_grad(set_to_none=True)
            o = model({'pixel_id': pid, 'num_frames': 3, 'iter_num': 0, 'sub_batch_index': 0})
            loss = sum(((o[k] - target) ** 2).mean() for k in ('rgb_coarse', 'rgb_fine', 'points_augmentation_rgb_coarse',
                                                                 'views_augmentation_rgb_coarse'))
            loss = loss + 0.1 * (o['depth_coarse'] - o['points_augmentation_depth_coarse'].detach()).square().mean()
            loss.backward()
            opt.step()
        ms = _time_steps(step, steps=3, warmup=2)
        out['simple_nerf_train_iteration'] = {'iters_per_sec': 1e3 / ms, 'ms_per_iter': ms, 'rays_per_iter': 4096,
                                              'note': 'forward + dgrad + wgrad on tcgen05 kernels, compositing backward hand-written; synthetic MSE + depth-consistency loss'}
    except Exception as e:
        out['simple_nerf_train_iteration'] = {'error': repr(e)[:200]}
    # ---- Simple-TensoRF training iteration (configs[2] shape): 4096 rays, 300^3-class main tensor + points-augmentation
    # tensor, alpha mask set, Adam step; every kernel of forward and backward is hand-written
    try:
        cfg = synthetic.tensorf_configs(num_voxels=300 ** 3, augmentations=True, rng_mode='device')
        mc = synthetic.scene_model_configs('re10k', num_views=3)
        torch.manual_seed(0)
        model = SimpleTensoRF(cfg, mc).to(device).train()
        t = model.coarse_model
        with torch.no_grad():
            for p_ in t.matrices_density:
                p_.mul_(6.0)
        vol = (torch.rand(190, 190, 190, generator=torch.Generator().manual_seed(1)) < 0.05).float()
        t.alpha_mask = AlphaGridMask(vol, t.bounding_box.cpu()).to(device)
        opt = torch.optim.Adam(model.get_trainable_parameters(cfg['optimizers'][0]), betas=(0.9, 0.99))
        model.optimizers = {'optimizer_nerf': opt}        # what Trainer10.py:59-62 does: the drop-in attaches its fused Adam step
        h, w = mc['resolution']
        g = torch.Generator().manual_seed(2)
        pid = torch.stack([torch.randint(0, 3, (4096,), generator=g), torch.randint(0, w, (4096,), generator=g),
                           torch.randint(0, h, (4096,), generator=g)], 1).int().to(device)
        target = torch.rand(4096, 3, device=device)

        def tstep():
            opt.zero_grad(set_to_none=True)
            o = model({'pixel_id': pid, 'num_frames': 3, 'iter_num': 0, 'sub_batch_index': 0})
            loss = sum(((o[k] - target) ** 2).mean() for k in ('rgb_coarse', 'points_augmentation_rgb_coarse'))
            loss = loss + 0.1 * (o['depth_coarse'] - o['points_augmentation_depth_coarse'].detach()).square().mean()
            loss.backward()
            opt.step()
        ms = _time_steps(tstep, steps=3, warmup=2)
        out['simple_tensorf_train_iteration'] = {'iters_per_sec': 1e3 / ms, 'ms_per_iter': ms, 'rays_per_iter': 4096,
                                                 'samples_per_ray': int(t.num_samples),
                                                 'note': 'VM gathers / scatters, colour MLP forward + dgrad + wgrad on tcgen05, compositing fwd/bwd; synthetic MSE + depth-consistency loss'}
    except Exception as e:
        out['simple_tensorf_train_iteration'] = {'error': repr(e)[:200]}
    return out


def cpu_render_sample(num_rays, configs, model_configs, sets):
    from oracle import pipeline as P
    from simple_rf_b200 import synthetic
    pid = torch.from_numpy(synthetic.frame_pixel_ids(FRAME_H, FRAME_W, view=0)[::max(1, FRAME_H * FRAME_W // num_rays)][:num_rays])
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in range(0, pid.shape[0], 1024):              # the reference's 1024-ray batches (BASELINE configs[0])
            P.nerf_render_chunk(sets, configs, model_configs, pid[i:i + 1024], training=False, retraw=False)
    return pid.shape[0], time.perf_counter() - t0


def cpu_setup():
    from oracle import nerf_mlp as M
    from simple_rf_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    configs = synthetic.nerf_configs(augmentations=False)
    model_configs = synthetic.scene_model_configs('llff', num_views=3)
    g = torch.Generator().manual_seed(0)
    sets = {'coarse_model': M.init_mlp_params(configs['model']['coarse_model'], g),
            'fine_model': M.init_mlp_params(configs['model']['fine_model'], g)}
    return configs, model_configs, sets


def cpu_baseline(budget_s):
    configs, model_configs, sets = cpu_setup()
    cpu_render_sample(1024, configs, model_configs, sets)                       # warm-up
    n, dt = cpu_render_sample(1024, configs, model_configs, sets)
    rays = int(min(65536, max(1024, (budget_s / max(dt, 1e-3)) * 1024 // 1024 * 1024)))
    n, dt = cpu_render_sample(rays, configs, model_configs, sets)
    return {'value': n / dt, 'unit': 'rays/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{n} rays of the same frame in 1024-ray batches, oracle/pipeline.py (torch CPU ops), {dt:.1f} s'}


def run_reference(args, rank, world):
    if rank != 0:
        return
    configs, model_configs, sets = cpu_setup()
    rays_per_step = 4096
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_render_sample(1024, configs, model_configs, sets)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        m, _ = cpu_render_sample(rays_per_step, configs, model_configs, sets)
        n += m
    dt = time.perf_counter() - t0
    v = n / dt
    cb = {'value': v, 'unit': 'rays/s', 'cores': torch.get_num_threads(), 'kind': 'port',
          'sample': f'{rays_per_step} rays per step in 1024-ray batches, oracle/pipeline.py (CPU restatement of the reference; '
                    'the reference itself is Python and cannot travel to the GPU box)'}
    print(json.dumps({
        'impl': 'reference', 'metric': 'rendered_rays_per_sec', 'value': v, 'unit': 'rays/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
        'config': {'workload': 'simple_nerf_frame_render', 'frame': [FRAME_H, FRAME_W], 'rays_per_step': rays_per_step,
                   'samples': f'{S_COARSE} coarse + {N_FINE} fine', 'note': 'bounded sample of the frame; host CPU only'},
        'cpu_baseline': cb, 'e2e': {'value': v, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-extras', action='store_true', help='skip the informational secondary measurements')
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
