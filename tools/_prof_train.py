import sys, torch
sys.path.insert(0, '.')
sys.argv = ["tensorf_train_step.py"]
import runpy
from torch.profiler import profile, ProfilerActivity
ns = runpy.run_path('tools/tensorf_train_step.py')
step = ns['step']
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=25, max_name_column_width=60))
