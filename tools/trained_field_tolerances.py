#!/usr/bin/env python
"""bf16 tensor-core MLP error on a TRAINED field (VERDICT r1, weak #2 / next #1b).

Trains Simple-NeRF with THIS repo (drop-in model + DataPreprocessor91 + fused losses, unmodified Trainer.train_one_iter) on
the analytic synthetic LLFF scene for `--iters` iterations, then compares, on that checkpoint,
  * a test-time render (sigma, rgb, weights, depth, depth_ndc, acc) and
  * the parameter gradients of one training iteration (same batch, same CPU random stream)
of the drop-in (bf16 operands, fp32 accumulate) against the reference's own eager fp32 model (SimpleNeRF17 from
baseline/_ref) holding the same state dict, on the same GPU.  Recipe is seed-deterministic: scene seed 3, training seed 11.

    python tools/trained_field_tolerances.py --iters 3000 --out gpurun_out/trained_field.json
"""
import argparse
import copy
import json
import sys
import time
from pathlib import Path

import numpy
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from simple_rf_b200.dropin import callers as C  # noqa: E402


def configs(name_dropin, rng_mode):
    cfg = C.complete_configs(C.load_shipped_configs(1142), [0], seed=11)
    for loss in cfg['losses']:
        if 'iter_weights' in loss:
            loss['iter_weights'] = {'0': 0.0, '1000': 0.1}
    if name_dropin:
        cfg = C.use_dropin(cfg)
        cfg['model']['rng_mode'] = rng_mode
    return cfg


def stats(got, want, name):
    got, want = got.float(), want.float()
    diff = (got - want).abs()
    return {f'{name}_max_abs': float(diff.max()), f'{name}_rel_l2': float(diff.norm() / want.norm().clamp_min(1e-12)),
            f'{name}_ref_max': float(want.abs().max())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=3000)
    ap.add_argument('--res', type=int, nargs=2, default=[378, 504])
    ap.add_argument('--out', default=str(ROOT / 'gpurun_out' / 'trained_field.json'))
    args = ap.parse_args()
    raw = C.synthetic_raw_data('llff', 3, resolution=tuple(args.res), sparse_points=2000, seed=3)
    trainer, model, mc = C.make_trainer(configs(True, 'device'), raw, seed=11)
    t0 = time.time()
    curve = []
    for it in range(args.iters):
        losses = trainer.train_one_iter(it)
        C.step_learning_rates(trainer, it)
        if it % 250 == 0 or it == args.iters - 1:
            curve.append((it, {k: float(v) for k, v in losses.items()}))
            print(it, curve[-1][1], flush=True)
    torch.cuda.synchronize()
    train_s = time.time() - t0
    state = copy.deepcopy(model.state_dict())
    report = {'iters': args.iters, 'resolution': list(args.res), 'train_seconds': train_s, 'it_per_s': args.iters / train_s,
              'loss_curve': curve}

    # ---- test-time render of a held-out pose: drop-in vs eager fp32 reference, same weights
    h, w = mc['resolution']
    ys, xs = numpy.meshgrid(numpy.arange(0, h, 2, dtype=numpy.int32), numpy.arange(0, w, 2, dtype=numpy.int32), indexing='ij')
    pid = torch.from_numpy(numpy.stack([numpy.ones(xs.size, dtype=numpy.int32), xs.reshape(-1), ys.reshape(-1)], 1)).cuda()
    outs = {}
    for tag, cfg in (('dropin', configs(True, 'reference')), ('dropin_bf16x3', configs(True, 'reference')), ('reference', configs(False, None))):
        if tag == 'dropin_bf16x3':
            cfg['model']['mlp_precision'] = 'bf16x3'             # split-bf16 operands: the fp32-contract program
        tester = C.make_tester(cfg, mc, [0])
        tester.model.load_state_dict(state)
        tester.model.eval()
        with torch.no_grad():
            outs[tag] = tester.model.module({'pixel_id': pid, 'num_frames': 3}, retraw=True)
        del tester
    b = outs['reference']
    sig = b['raw_sigma_fine'].float()
    for tag, key in (('dropin', 'eval'), ('dropin_bf16x3', 'eval_bf16x3')):
        a = outs[tag]
        ev = {}
        for k in ('rgb_coarse', 'rgb_fine', 'acc_fine', 'depth_fine', 'depth_ndc_fine', 'depth_ndc_coarse', 'weights_fine', 'weights_coarse',
                  'raw_sigma_fine', 'raw_sigma_coarse', 'raw_rgb_fine', 'raw_rgb_coarse', 'z_vals_fine'):
            ev.update(stats(a[k], b[k], k))
        ev['sigma_fine_p50'], ev['sigma_fine_p99'], ev['sigma_fine_max'] = (float(torch.quantile(sig.flatten()[:4_000_000], q)) for q in (0.5, 0.99, 1.0))
        ev['weights_fine_peak_median'] = float(b['weights_fine'].max(dim=1)[0].median())
        # rays whose fine depths coincide (the inverse-CDF resampling is discontinuous: a coarse weight moving by 1e-6 can move
        # a fine sample to the neighbouring bin) carry the clean per-sample comparison of the fine stage
        same = (a['z_vals_fine'] - b['z_vals_fine']).abs().amax(dim=1) <= 1e-6
        ev['rays_with_identical_fine_depths'] = float(same.float().mean())
        if same.any():
            for k in ('rgb_fine', 'depth_fine', 'weights_fine', 'raw_sigma_fine', 'raw_rgb_fine'):
                ev.update(stats(a[k][same], b[k][same], f'same_z_{k}'))
        # sigma error where it matters: relative to max(sigma, 1) on samples carrying weight
        wmask = b['weights_coarse'] > 1e-3
        sc = b['raw_sigma_coarse'].float()[..., 0]
        rel = ((a['raw_sigma_coarse'][..., 0] - sc).abs() / sc.clamp_min(1.0))[wmask]
        ev['sigma_coarse_rel_err_on_weighted_samples_max'] = float(rel.max()) if rel.numel() else 0.0
        ev['sigma_coarse_rel_err_on_weighted_samples_p99'] = float(torch.quantile(rel[:4_000_000], 0.99)) if rel.numel() else 0.0
        report[key] = ev
        print(tag, json.dumps(ev, indent=1), flush=True)
    del outs, a, b

    # ---- gradients of one training iteration on the trained weights: same batch, same CPU random stream
    grads, losses = {}, {}
    for tag, cfg in (('dropin', configs(True, 'reference')), ('reference', configs(False, None))):
        cfg['losses'] = [dict(l, iter_weights={'0': 0.1}) if 'iter_weights' in l else l for l in cfg['losses']]
        tr, mdl, _ = C.make_trainer(cfg, raw, seed=77)
        mdl.load_state_dict(state)
        losses[tag] = {k: float(v) for k, v in tr.train_one_iter(0).items()}
        grads[tag] = {n: p.grad.detach().float().clone() for n, p in mdl.module.named_parameters() if p.grad is not None}
        del tr, mdl
    gr = {}
    for n, g in grads['reference'].items():
        d = grads['dropin'][n]
        gr[n] = float((d - g).norm() / g.norm().clamp_min(1e-20))
    report['train'] = {'losses_dropin': losses['dropin'], 'losses_reference': losses['reference'],
                       'gradient_relative_l2': gr, 'gradient_relative_l2_worst': max(gr.values()),
                       'gradient_relative_l2_median': float(numpy.median(list(gr.values())))}
    print('losses', losses, '\nworst gradient rel-L2', sorted(gr.items(), key=lambda kv: -kv[1])[:6], flush=True)
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(report, indent=1))


if __name__ == '__main__':
    main()
