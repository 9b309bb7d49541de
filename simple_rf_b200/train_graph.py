"""One training iteration as ONE CUDA graph.

A Simple-NeRF iteration is ~8 ms of GPU work issued as several hundred launches (4 MLP evaluations with hand-written forward /
dgrad / wgrad kernels, compositing, sampling, the loss terms, the all-reduce and the optimiser step); on one GPU the host
needs longer to issue them than the GPU needs to run them, and with the batch sharded over 8 ranks the GPU work per rank
shrinks to ~1 ms while the issue time stays.  `GraphedStep` captures `step_fn` — forward, losses, backward, gradient
all-reduce, fused Adam — once and replays it with one launch per iteration.

Requirements (checked where possible):
  * `rng_mode='device'`: random numbers come from torch's CUDA generator, whose Philox offset torch advances per replay
    (the reference-order CPU random stream cannot be replayed from a graph);
  * the optimisers go through `optim.FusedFlatAdam.make_capturable()` (step count and learning rate in device memory);
  * inputs are STATIC tensors: write the next batch into them (`copy_`) before `replay()`;
  * no model surgery inside the step (TensoRF: re-capture after `run_model_modifications` changed the grids).
"""
import torch


class GraphedStep:
    def __init__(self, step_fn, optimizers=(), warmup=3):
        """step_fn(): one full iteration on static inputs, returns a tensor (e.g. the loss) or None.  optimizers: the
        torch optimisers stepped inside (their `_srf_fused` wrappers are switched to the capturable form)."""
        self.step_fn = step_fn
        self.fused = []
        for opt in (optimizers.values() if isinstance(optimizers, dict) else optimizers):
            f = getattr(opt, '_srf_fused', None)
            if f is None:
                raise RuntimeError('GraphedStep needs optimisers wrapped by simple_rf_b200.optim.FusedFlatAdam (assign them to model.optimizers)')
            self.fused.append(f.make_capturable())
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):                       # builds the flat buffers, the device-side step / lr, every lazy cache
                step_fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: NCCL's watchdog thread may query events while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode='thread_local'):
            self.output = step_fn()
        self.replays = 0

    def replay(self):
        for f in self.fused:
            f.sync_hyperparameters()
        self.graph.replay()
        for f in self.fused:
            f.after_replay()
        self.replays += 1
        return self.output
