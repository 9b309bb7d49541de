"""Ray-sharded multi-GPU execution: one process per GPU (torchrun), NCCL over NVLink for the single collective.

The reference only has `nn.DataParallel` on one device (src/Trainer10.py:566).  Here:
  * inference: every rank renders a contiguous band of the frame's rays (no collective on the data path; an
    optional gather of the per-ray maps to rank 0 for the caller that wants the whole frame);
  * training: every rank takes its slice of the globally seeded batch — the image-ray block and the
    sparse-depth-ray block are split separately so per-rank loss means average to the global mean
     — and ONE all-reduce of a flat fp32 gradient bucket runs before `optimizer.step()`.
The gradient hook is attached when the trainer hands the optimisers to the model (`model.optimizers = ...`,
src/Trainer10.py:61-62), so Trainer10 itself stays unmodified.
"""
import os

import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (no-op for a single process)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1 or dist.is_initialized():
        return int(os.environ.get('RANK', '0')), world
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(local)
        dist.init_process_group(backend, device_id=torch.device('cuda', local))
    else:
        dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size()


def shard_bounds(n, rank, world):
    """Contiguous, balanced [start, end) of `n` items for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(indices_mask_nerf, rank, world):
    """Row indices of this rank's share of a training batch: the image-ray rows and the sparse-depth rows
    (indices_mask_nerf false) are each split evenly, keeping every rank's proportions equal to the global ones."""
    mask = torch.as_tensor(indices_mask_nerf).bool()
    rows = torch.arange(mask.numel(), device=mask.device)
    parts = []
    for block in (rows[mask], rows[~mask]):
        s, e = shard_bounds(block.numel(), rank, world)
        parts.append(block[s:e])
    return torch.sort(torch.cat(parts))[0]


class FlatGradAllReduce:
    """Sums the gradients of every parameter of the optimiser's groups across ranks in ONE collective on a
    flat fp32 bucket, then scales by 1/world, right before `optimizer.step()` (step pre-hook).  The bucket is
    rebuilt lazily whenever the parameter set changes (TensoRF swaps its plane parameters when it upsamples,
    src/models/SimpleTensoRF09.py:916-944)."""

    def __init__(self, optimizer, group=None):
        self.optimizer = optimizer
        self.group = group
        self._key = None
        self._bucket = None
        self.handle = optimizer.register_step_pre_hook(self._hook)
        self.bytes_last = 0

    def _params(self):
        return [p for g in self.optimizer.param_groups for p in g['params'] if p.requires_grad]

    def _hook(self, optimizer, args, kwargs):
        self.reduce()

    def reduce(self):
        if not is_distributed():
            return
        world = dist.get_world_size(self.group)
        params = self._params()
        key = tuple((p.data_ptr(), p.numel()) for p in params)
        if key != self._key:
            total = sum(p.numel() for p in params)
            self._bucket = torch.zeros(total, dtype=torch.float32, device=params[0].device)
            self._key = key
        bucket = self._bucket
        views, o = [], 0
        for p in params:
            v = bucket[o:o + p.numel()].view_as(p)
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
            views.append(v)
            o += p.numel()
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group)
        bucket.mul_(1.0 / world)
        for p, v in zip(params, views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        self.bytes_last = bucket.numel() * 4


def attach_gradient_allreduce(optimizers):
    """optimizers: the trainer's dict name -> torch optimiser.  Returns the hooks (kept alive by the caller)."""
    if not is_distributed():
        return []
    # optimisers stepping through optim.FusedFlatAdam reduce their own flat gradient bucket
    return [FlatGradAllReduce(opt) for opt in optimizers.values() if opt is not None and getattr(opt, '_srf_fused', None) is None]


def gather_ray_outputs(out, num_rays):
    """All-gather of a dict of per-ray tensors ([rows_of_this_rank, ...], rows = shard_bounds(num_rays, rank, world)) into
    full-frame tensors on every rank: ONE collective over a [longest_band, total_width] fp32 record (bands differ by at most
    one row; the pad row is dropped).  This is the only collective of a sharded render and it moves 16-28 B/ray/key."""
    rank, world = dist.get_rank(), dist.get_world_size()
    spans = [shard_bounds(num_rays, r, world) for r in range(world)]
    longest = max(e - s for s, e in spans)
    keys = sorted(out.keys())
    flat = [out[k].reshape(out[k].shape[0], -1).float() for k in keys]
    widths = [f.shape[1] for f in flat]
    rec = torch.zeros((longest, sum(widths)), dtype=torch.float32, device=flat[0].device)
    rec[:flat[0].shape[0]] = torch.cat(flat, 1)
    full = torch.empty((world, longest, sum(widths)), dtype=torch.float32, device=rec.device)
    dist.all_gather_into_tensor(full.view(world * longest, -1), rec)
    rows = torch.cat([full[r, :e - s] for r, (s, e) in enumerate(spans)], 0)
    res, o = {}, 0
    for k, wd in zip(keys, widths):
        res[k] = rows[:, o:o + wd].reshape((num_rays,) + tuple(out[k].shape[1:])).to(out[k].dtype)
        o += wd
    return res


def eval_band(num_rays, configs_model):
    """(start, end) of this rank's row band of a test-time frame, or None when the render is not sharded (single process,
    `configs['model']['shard_eval_rays'] = False`, or fewer rays than ranks)."""
    if not is_distributed() or not configs_model.get('shard_eval_rays', True) or num_rays < dist.get_world_size():
        return None
    return shard_bounds(num_rays, dist.get_rank(), dist.get_world_size())


def rows_of_global_draw(draw, rows_local, shard, chunk):
    """`draw(n)` makes the reference's CPU random tensor for n rays (first dim n).  On one rank: draw(rows_local).  On several
    ranks with a `srf_shard` record from DataPreprocessor91 (rank, world, global_rows, row_indices): this rank's rows of the
    single global draw, i.e. exactly the numbers the same rays get in a single-GPU run — as long as the global
    batch is one `chunk` (the shipped configuration: 4096 of 4096)."""
    if shard is None or shard[2] > chunk or len(shard[3]) != rows_local:
        return draw(rows_local)
    return draw(shard[2])[torch.from_numpy(shard[3])]


_DEVICE_RNG_DECORRELATED = False


def decorrelate_device_rng():
    """`rng_mode='device'` draws jitter / noise from torch's CUDA generator; every rank seeds it identically
    (init_seeds, src/Trainer10.py:443-450), which would give different rays the same random numbers on every rank.
    Offsets the generator of this process by its rank, once."""
    global _DEVICE_RNG_DECORRELATED
    if _DEVICE_RNG_DECORRELATED or not is_distributed() or not torch.cuda.is_available():
        return
    torch.cuda.manual_seed(torch.cuda.initial_seed() + 7919 * dist.get_rank())
    _DEVICE_RNG_DECORRELATED = True


def rank_seed(seed):
    """Decorrelates the in-kernel Philox streams of the ranks (`rng_mode='device'`): every rank draws the same CPU seed."""
    if not is_distributed():
        return seed
    return (seed ^ (dist.get_rank() * 0x5bd1e995)) & 0x7fffffff


def render_sharded(model, input_batch, gather_keys=None, **forward_kwargs):
    """Render this rank's band of `input_batch['pixel_id']`; when `gather_keys` is given, all-gather those
    per-ray outputs so every rank (rank 0 included) holds the full-frame tensors."""
    rank = dist.get_rank() if is_distributed() else 0
    world = dist.get_world_size() if is_distributed() else 1
    pid = input_batch['pixel_id']
    s, e = shard_bounds(pid.shape[0], rank, world)
    local = dict(input_batch)
    local['pixel_id'] = pid[s:e]
    out = model(local, **forward_kwargs)
    if world == 1 or not gather_keys:
        return out
    n = pid.shape[0]
    longest = max(shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world))
    full = dict(out)
    for k in gather_keys:
        t = out[k]
        pad = torch.zeros((longest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        gathered = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(gathered, pad)
        pieces = []
        for r in range(world):
            rs, re = shard_bounds(n, r, world)
            pieces.append(gathered[r][:re - rs])
        full[k] = torch.cat(pieces, 0)
    return full
