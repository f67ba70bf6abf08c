"""Data parallelism for the SAE step: one process per GPU, images sharded across ranks, parameters and
Adam state replicated.  The only exchange is an all-reduce of the flat buffer svb_*_step_grads fills — gradients (SUM),
loss partial sums and per-feature activity counts (SUM), per-channel max / -min (MAX) — between the two halves of
the step; projection + Adam + renormalisation then run identically on every rank (svb_*_step_apply).

The reference has no distributed code at all; this module adds it around the drop-in boundary.
`all_reduce_flat` is written against torch.distributed only, so it runs on NCCL (GPU) and on gloo (CPU tests).
"""
import torch
import torch.distributed as dist


def shard_images(n_images, rank, world):
    """Contiguous image shard [lo, hi) of `rank`; remainders go to the first ranks.  Tokens of one image never
    straddle ranks, so per-image sparsity / variance-explained stay local."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_flat(flat, n_sum, n_max, group=None):
    """In-place reduction of the flat step buffer: first n_sum elements SUM, next n_max elements MAX."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return flat
    dist.all_reduce(flat[:n_sum], op=dist.ReduceOp.SUM, group=group)
    if n_max:
        dist.all_reduce(flat[n_sum:n_sum + n_max], op=dist.ReduceOp.MAX, group=group)
    return flat


def global_counts(n_images_local, hw, group=None, device=None):
    """(global images, global tokens) over the data-parallel group."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return n_images_local, n_images_local * hw
    t = torch.tensor([n_images_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    n = int(t.item())
    return n, n * hw


class DataParallelStep:
    """grads() on the local shard -> all-reduce -> apply().  `kind` is 'sae_mlp' or 'gated_sae'."""

    def __init__(self, kind, group=None):
        self.kind = kind
        self.group = group

    def step(self, x_local, params, adam_m, adam_v, step, lr, lam, expansion_factor, optimizer, betas,
             global_images, global_tokens, want_dec=True, eps=1e-8):
        from . import ops
        ss = ops.SplitStep(self.kind, x_local, params, lam, want_dec=want_dec)
        addr, n_sum, n_max = ss.grads(global_tokens=global_tokens)
        flat = ops.wrap_device_buffer(addr, n_sum + n_max, x_local.device)
        all_reduce_flat(flat, n_sum, n_max, self.group)
        return ss.apply(adam_m, adam_v, step, lr, expansion_factor, optimizer, betas, eps=eps,
                        global_tokens=global_tokens, global_images=global_images)
