"""Where the host time of the UNMODIFIED Trainer.train_one_iter goes when it drives the drop-in (cProfile, cumulative)."""
import cProfile
import pstats
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from simple_rf_b200.dropin import callers as C

kind = sys.argv[1] if len(sys.argv) > 1 else 'nerf'
C.prepare()
if kind == 'nerf':
    cfg = C.use_dropin(C.complete_configs(C.load_shipped_configs(1142), [0], seed=0))
    raw = C.synthetic_raw_data('llff', 3, resolution=(378, 504), sparse_points=2000, seed=0)
else:
    cfg = C.use_dropin(C.complete_configs(C.load_shipped_configs(212), [0], seed=0))
    raw = C.synthetic_raw_data('re10k', 3, resolution=(288, 512), sparse_points=2000, seed=0, tensorf=True)
cfg['model']['rng_mode'] = 'device'
for loss in cfg['losses']:
    if 'iter_weights' in loss:
        loss['iter_weights'] = {'0': 0.1}
trainer, model, mc = C.make_trainer(cfg, raw, seed=0)
for it in range(5):
    trainer.train_one_iter(it)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for it in range(20):
    trainer.train_one_iter(5 + it)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(45)
