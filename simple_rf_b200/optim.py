"""Optimiser tail of a training iteration ("next" row f4): the trainer's `torch.optim.Adam` instances
(src/optimizers/OptimizerFactory02.py:9-22; handed to the model through `model.optimizers`, src/Trainer10.py:59-62, and
stepped at src/Trainer10.py:109-110) keep their identity, parameter groups, `state_dict()` format and learning-rate
handling (src/Trainer10.py:303-308 rescales `param_group['lr']` every iteration), but `step()` becomes: one gather of the
gradients into a flat fp32 bucket, ONE all-reduce of that bucket when running on several ranks, and one fused Adam kernel
launch per parameter group (csrc/optim.cu) over flat parameter / moment buffers the parameters and optimiser states are
views of.  torch's default path runs ~10 multi-tensor passes over ~100 small tensors (56 launches per step)."""
import ctypes

import torch

from . import _lib as L


class _FlatGroup:
    def __init__(self, group, optimizer):
        params = [p for p in group['params'] if p.requires_grad]
        self.params = params
        self.sizes = [p.numel() for p in params]
        self.offsets, o = [], 0
        for n in self.sizes:
            self.offsets.append(o)
            o += -(-n // 4) * 4                       # every parameter starts 16-byte aligned
        self.total = o
        dev = params[0].device
        self.flat_p = torch.zeros(o, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(o, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(o, dtype=torch.float32, device=dev)
        # gradient bucket + one 'this rank holds a gradient' flag per parameter behind it: both travel in the SAME all-reduce
        self.flat_g = torch.zeros(o + -(-len(params) // 4) * 4, dtype=torch.float32, device=dev)
        self.flags = self.flat_g[o:o + len(params)]
        self.steps = []                                # per parameter, as torch keeps them (they differ after TensoRF's
        for p, off, n in zip(params, self.offsets, self.sizes):       # reconfigure_optimizer deletes states by index)
            st = optimizer.state.get(p, {})
            self.flat_p[off:off + n].copy_(p.data.reshape(-1))
            step = 0
            if 'exp_avg' in st:                        # resumed / already stepped: adopt the moments
                self.flat_m[off:off + n].copy_(st['exp_avg'].reshape(-1))
                self.flat_v[off:off + n].copy_(st['exp_avg_sq'].reshape(-1))
                step = int(float(st['step']))
            self.steps.append(step)
            p.data = self.flat_p[off:off + n].view(p.shape)
            optimizer.state[p] = {'step': torch.tensor(float(step)), 'exp_avg': self.flat_m[off:off + n].view(p.shape),
                                  'exp_avg_sq': self.flat_v[off:off + n].view(p.shape)}
        self.key = self.make_key(group, optimizer)

    @staticmethod
    def make_key(group, optimizer):
        key = []
        for p in group['params']:
            if p.requires_grad:
                st = optimizer.state.get(p)
                key.append((id(p), p.data_ptr(), st['exp_avg'].data_ptr() if st and 'exp_avg' in st else 0))
        return tuple(key)


class FusedFlatAdam:
    """Wraps ONE torch.optim.Adam instance.  `optimizer.step()` keeps working for every caller; `optimizer.state_dict()` /
    `load_state_dict()` keep the torch format (the state tensors are views of the flat buffers; after a `load_state_dict`
    or a change of the parameter set — TensoRF swaps its planes when it upsamples, src/models/SimpleTensoRF09.py:916-944 —
    the flat buffers are rebuilt at the next step)."""

    def __init__(self, optimizer, process_group=None):
        assert supports(optimizer), 'FusedFlatAdam needs a plain torch.optim.Adam over CUDA fp32 parameters'
        self.optimizer = optimizer
        self.process_group = process_group
        self.capturable = False            # make_capturable(): step count and learning rate move to device memory (CUDA graphs)
        self.groups = [None] * len(optimizer.param_groups)
        self.bytes_reduced_last = 0
        self._torch_step = optimizer.step
        optimizer.step = self.step
        optimizer._srf_fused = self

    def detach(self):
        self.optimizer.step = self._torch_step
        self.optimizer._srf_fused = None

    # ---------------------------------------------------------------- CUDA-graph support (train_graph.GraphedStep)
    def make_capturable(self):
        """From now on `step()` reads the step count and the learning rate of every group from device memory
        (srf_adam_advance + srf_adam_step_capturable), so it can sit inside a captured CUDA graph.  Needs one (already
        performed or pending) flat-buffer build per group, equal step counts inside a group and a gradient on every parameter —
        the steady state of training.  `sync_hyperparameters()` pushes `param_group['lr']` (the trainer rescales it every
        iteration, src/Trainer10.py:303-308) to the device; `after_replay()` keeps the host mirrors current."""
        self.capturable = True
        return self

    def _device_state(self, fg, group):
        if getattr(fg, 'step_dev', None) is None:
            dev = fg.flat_p.device
            assert len(set(fg.steps)) <= 1, 'capturable mode needs equal step counts inside a parameter group'
            fg.step_dev = torch.full((1,), fg.steps[0] if fg.steps else 0, dtype=torch.int64, device=dev)
            fg.lr_dev = torch.full((1,), float(group['lr']), dtype=torch.float32, device=dev)
        return fg.step_dev, fg.lr_dev

    def sync_hyperparameters(self):
        """Outside a capture: copy every group's current `lr` into its device scalar (a fill kernel, no host->device copy)."""
        for group, fg in zip(self.optimizer.param_groups, self.groups):
            if fg is not None and getattr(fg, 'lr_dev', None) is not None:
                fg.lr_dev.fill_(float(group['lr']))

    def after_replay(self):
        """A captured step ran without this Python code: advance the host-side step mirrors and bump the parameters' version
        counters (the packed-weight / channels-last caches of the models are keyed on them)."""
        params = []
        for fg in self.groups:
            if fg is None or getattr(fg, 'step_dev', None) is None:
                continue
            fg.steps = [s + 1 for s in fg.steps]
            params.extend(fg.params)
        if params:
            torch.autograd.graph.increment_version(params)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        opt = self.optimizer
        if len(self.groups) != len(opt.param_groups):
            self.groups = [None] * len(opt.param_groups)
        work = []
        for gi, group in enumerate(opt.param_groups):
            if not any(p.requires_grad for p in group['params']):
                continue
            fg = self.groups[gi]
            if fg is None or fg.key != _FlatGroup.make_key(group, opt):
                fg = self.groups[gi] = _FlatGroup(group, opt)
            work.append((group, fg))
        from . import parallel
        world = torch.distributed.get_world_size(self.process_group) if parallel.is_distributed() else 1
        self.bytes_reduced_last = 0
        for group, fg in work:
            # gather the gradients (one batched copy).  A parameter is stepped iff SOME rank holds a gradient for it — what
            # torch.optim.Adam does on one GPU (parameters whose .grad is None are skipped: the stale planes TensoRF leaves in
            # the optimiser between shrink_tensor and the next reconfigure_optimizer, SimpleTensoRF09.py:821-830, must not
            # move, and their step counters must not advance).  The per-parameter 'has gradient' flags ride behind the
            # bucket in the same all-reduce; they are read back (one synchronisation) only by a rank that itself misses a
            # gradient — a rank holding all of them already knows the answer.
            have = [p.grad is not None for p in fg.params]
            if not any(have) and world == 1:
                continue
            if world > 1 or not all(have):
                fg.flat_g.zero_()
            dst = [fg.flat_g[o:o + n].view(p.shape) for p, o, n, h in zip(fg.params, fg.offsets, fg.sizes, have) if h]
            src = [p.grad for p, h in zip(fg.params, have) if h]
            if dst:
                torch._foreach_copy_(dst, src)
            have_step = have
            if world > 1:
                if all(have):
                    fg.flags.fill_(1.0)                  # a kernel, not a pageable host->device copy (which would drain the launch queue)
                else:
                    fg.flags.copy_(torch.tensor([1.0 if h else 0.0 for h in have]))
                timed = L.TIMING is not None
                if timed:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                torch.distributed.all_reduce(fg.flat_g, op=torch.distributed.ReduceOp.SUM, group=self.process_group)
                if timed:                              # NCCL runs on its own stream: the events bracket the wait of the compute stream
                    e1.record()
                    L.TIMING.append(('nccl_all_reduce', e0, e1, float(fg.flat_g.numel() * 4)))
                self.bytes_reduced_last += fg.flat_g.numel() * 4
                if not all(have):
                    have_step = [f > 0 for f in fg.flags.tolist()]
                fg.flat_g[:fg.total].mul_(1.0 / world)
                for p, o, n in zip(fg.params, fg.offsets, fg.sizes):          # the averaged gradient stays visible in .grad
                    if p.grad is not None:
                        p.grad.copy_(fg.flat_g[o:o + n].view(p.shape))
            if not any(have_step):
                continue
            beta1, beta2 = group['betas']
            if self.capturable:
                assert all(have_step), 'capturable mode: every parameter of the group needs a gradient'
                step_dev, lr_dev = self._device_state(fg, group)
                if not torch.cuda.is_current_stream_capturing():
                    lr_dev.fill_(float(group['lr']))
                L.call('srf_adam_advance', L.ptr(step_dev), L.stream_handle())
                L.call('srf_adam_step_capturable', L.ptr(fg.flat_p), L.ptr(fg.flat_g), L.ptr(fg.flat_m), L.ptr(fg.flat_v), fg.total,
                       L.ptr(lr_dev), float(beta1), float(beta2), float(group['eps']), float(group['weight_decay']), L.ptr(step_dev),
                       L.stream_handle())
                if not torch.cuda.is_current_stream_capturing():      # a capture only records: GraphedStep.after_replay keeps the mirrors
                    fg.steps = [st + 1 for st in fg.steps]
                    torch.autograd.graph.increment_version(fg.params)
                continue
            for a, b in _runs(have_step, fg.steps):     # one launch per run of stepped parameters with equal step counts
                step = fg.steps[a] + 1
                lo = fg.offsets[a]
                hi = fg.offsets[b - 1] + -(-fg.sizes[b - 1] // 4) * 4
                L.call('srf_adam_step', L.ptr(fg.flat_p[lo:]), L.ptr(fg.flat_g[lo:]), L.ptr(fg.flat_m[lo:]), L.ptr(fg.flat_v[lo:]),
                       hi - lo, float(group['lr']), float(beta1), float(beta2), float(group['eps']), float(group['weight_decay']),
                       step, L.stream_handle())
                stepf = torch.tensor(float(step))
                for i in range(a, b):
                    fg.steps[i] = step
                    opt.state[fg.params[i]]['step'] = stepf
                # the kernel wrote through raw pointers: bump the version counters the way torch's in-place ops do (the
                # packed-weight / channels-last caches of the models are keyed on them)
                torch.autograd.graph.increment_version(fg.params[a:b])
        return loss


def _runs(have, steps):
    """Maximal runs [a, b) of consecutive entries with have[i] and equal steps[i]."""
    runs, a = [], None
    for i in range(len(have) + 1):
        ok = i < len(have) and have[i]
        if a is not None and (not ok or steps[i] != steps[a]):
            runs.append((a, i))
            a = None
        if ok and a is None:
            a = i
    return runs


def enabled():
    """SIMPLE_RF_B200_FUSED_ADAM=0 keeps torch's own Adam step (debugging aid)."""
    import os
    return os.environ.get('SIMPLE_RF_B200_FUSED_ADAM', '1') != '0'


def supports(optimizer):
    if type(optimizer) is not torch.optim.Adam:
        return False
    for g in optimizer.param_groups:
        if g.get('amsgrad') or g.get('maximize') or g.get('capturable') or g.get('differentiable'):
            return False
        if isinstance(g['lr'], torch.Tensor):
            return False
        for p in g['params']:
            if p.requires_grad and (not p.is_cuda or p.dtype != torch.float32 or p.is_sparse):
                return False
    return any(p.requires_grad for g in optimizer.param_groups for p in g['params'])


def attach(optimizers, process_group=None):
    """optimizers: the trainer's dict name -> torch optimiser (src/Trainer10.py:507-537).  Every plain Adam over CUDA fp32
    parameters gets the fused step (idempotent); anything else is left alone (and, on several ranks, gets the generic
    flat-bucket all-reduce hook of parallel.py).  Returns the wrappers."""
    out = []
    for opt in optimizers.values():
        if opt is None:
            continue
        if getattr(opt, '_srf_fused', None) is not None:
            out.append(opt._srf_fused)
        elif supports(opt):
            out.append(FusedFlatAdam(opt, process_group))
    return out
