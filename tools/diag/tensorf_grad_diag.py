"""Diagnostic: per-tensor gradient agreement (relative L2, cosine, norm ratio) of the drop-in Simple-TensoRF against the
reference's eager fp32 model through the unmodified Trainer.train_one_iter, at iterations 0 and 1, for two batch sizes."""
import copy, json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
from simple_rf_b200.dropin import callers as C
import test_gpu_reference_callers as T

out = {}
for rays in (2048, 128):
    raw = C.synthetic_raw_data('re10k', 3, resolution=(144, 256), sparse_points=600, seed=7, tensorf=True)
    cfg_ref = T._tensorf_configs()
    cfg_ref['data_loader']['num_rays'] = rays
    cfg_ref['data_loader']['sparse_depth']['num_rays'] = rays
    grads = {}
    for tag, cfg in (('ref', cfg_ref), ('mine', C.use_dropin(cfg_ref))):
        trainer, model, mc = C.make_trainer(cfg, raw, seed=cfg['seed'])
        g = []
        for it in range(2):
            trainer.train_one_iter(it)
            g.append({n: p.grad.detach().float().cpu().clone() for n, p in model.module.named_parameters() if p.grad is not None})
            C.step_learning_rates(trainer, it)
        grads[tag] = g
    for it in range(2):
        rows = {}
        for n, gr in grads['ref'][it].items():
            d = grads['mine'][it][n]
            rows[n] = (float((d - gr).norm() / gr.norm().clamp_min(1e-20)), float(torch.nn.functional.cosine_similarity(d.flatten(), gr.flatten(), dim=0)),
                       float(d.norm() / gr.norm().clamp_min(1e-20)), float(gr.norm()))
        out[f'rays{rays}_it{it}'] = rows
        print(f'--- rays {rays} iteration {it}')
        for n, v in sorted(rows.items(), key=lambda kv: -kv[1][0])[:14]:
            print(f'{n:60s} rel {v[0]:.4f} cos {v[1]:.5f} norm ratio {v[2]:.4f} |g| {v[3]:.3e}')
Path(ROOT / 'gpurun_out' / 'tensorf_grad_diag.json').write_text(json.dumps(out, indent=1))
