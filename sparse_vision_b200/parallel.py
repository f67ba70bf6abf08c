"""Data parallelism for the SAE step: one process per GPU, images sharded across ranks, parameters and
Adam state replicated.  The only exchange is an all-reduce of the flat buffer svb_*_step_grads fills — gradients (SUM),
loss partial sums and per-feature activity counts (SUM), per-channel max / -min (MAX) — between the two halves of
the step; projection + Adam + renormalisation then run identically on every rank (svb_*_step_apply).
The encoder-side gradients form the leading part of the buffer and are final before the decoder weight-gradient GEMM
starts: with `overlap=True` their all-reduce is issued on a communication stream the library releases at that point
(svb_set_comm_stream), so it runs over NVLink while that GEMM is still computing.

The reference has no distributed code at all; this module adds it around the drop-in boundary.
`all_reduce_flat` is written against torch.distributed only, so it runs on NCCL (GPU) and on gloo (CPU tests).
"""
import ctypes

import torch
import torch.distributed as dist


def shard_images(n_images, rank, world):
    """Contiguous image shard [lo, hi) of `rank`; remainders go to the first ranks.  Tokens of one image never
    straddle ranks, so per-image sparsity / variance-explained stay local."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_flat(flat, n_sum, n_max, group=None, early=0, comm_stream=None):
    """In-place reduction of the flat step buffer: first n_sum elements SUM, next n_max elements MAX.
    early > 0 with a comm_stream: the leading `early` elements are reduced on that stream (which the library has made
    wait for exactly those elements), overlapping the work still queued on the current stream."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return flat
    work = None
    if early > 0 and comm_stream is not None:
        with torch.cuda.stream(comm_stream):
            work = dist.all_reduce(flat[:early], op=dist.ReduceOp.SUM, group=group, async_op=True)
        dist.all_reduce(flat[early:n_sum], op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(flat[:n_sum], op=dist.ReduceOp.SUM, group=group)
    if n_max:
        dist.all_reduce(flat[n_sum:n_sum + n_max], op=dist.ReduceOp.MAX, group=group)
    if work is not None:
        work.wait()   # the current stream waits for the early bucket
    return flat


def global_counts(n_images_local, hw, group=None, device=None):
    """(global images, global tokens) over the data-parallel group."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return n_images_local, n_images_local * hw
    t = torch.tensor([n_images_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    n = int(t.item())
    return n, n * hw


def connect_peer_memory(device, n_floats, group=None):
    """Sets up the library's peer-memory exchange buffer on `device` for all ranks of `group` (one node, <= 8 GPUs):
    svb_comm_alloc -> all-gather of the 64-byte CUDA IPC handles -> svb_comm_connect.  Returns True when EVERY rank
    succeeded (otherwise all ranks drop the buffer again and the caller falls back to torch.distributed)."""
    from . import _lib as L
    lib, h = L.load(), L.handle(device)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    cap = ctypes.c_int64()
    lib.svb_comm_capacity(h, ctypes.byref(cap))
    if cap.value >= n_floats:      # an earlier step on this device already connected a large enough buffer
        return True
    if cap.value > 0:
        lib.svb_comm_destroy(h)    # too small: every rank sees the same sizes, so every rank rebuilds it
    handle = ctypes.create_string_buffer(64)
    ok = world <= 8 and lib.svb_comm_alloc(h, int(n_floats), handle) == 0
    gathered = [None] * world
    dist.all_gather_object(gathered, (bool(ok), handle.raw), group=group)
    ok = all(g[0] for g in gathered)
    if ok:
        blob = ctypes.create_string_buffer(b"".join(g[1] for g in gathered), 64 * world)
        ok = lib.svb_comm_connect(h, rank, world, blob) == 0
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    ok = bool(flag.item())
    if not ok:
        lib.svb_comm_destroy(h)
    return ok


class DataParallelStep:
    """grads() on the local shard -> all-reduce -> apply().  `kind` is 'sae_mlp' or 'gated_sae'.

    exchange = 'peer' (default on CUDA): the flat buffer lives in CUDA-IPC peer memory and ONE kernel of the library
    reduces it over NVLink (svb_comm_allreduce); falls back to 'nccl' when peer memory cannot be set up.
    exchange = 'nccl': torch.distributed all-reduce (also the CPU / gloo path); overlap=True additionally sends the
    encoder-side gradients on a communication stream while the decoder weight-gradient GEMM runs."""

    def __init__(self, kind, group=None, exchange="peer", overlap=False):
        self.kind = kind
        self.group = group
        self.exchange = exchange
        self.overlap = overlap
        self.comm_stream = None
        self.peer = None   # None: not tried yet, True / False afterwards

    def step(self, x_local, params, adam_m, adam_v, step, lr, lam, expansion_factor, optimizer, betas,
             global_images, global_tokens, want_dec=True, eps=1e-8):
        from . import ops
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        if self.overlap and self.exchange != "peer" and multi and x_local.is_cuda and self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=x_local.device)
            ops.set_comm_stream(x_local.device, self.comm_stream)
        ss = ops.SplitStep(self.kind, x_local, params, lam, want_dec=want_dec)
        addr, n_sum, n_max = ss.grads(global_tokens=global_tokens)
        if multi and x_local.is_cuda and self.exchange == "peer" and self.peer is None:
            self.peer = connect_peer_memory(x_local.device, n_sum + n_max, self.group)
            if self.peer:   # once: rebuild this step's buffer inside the exchange region
                addr, n_sum, n_max = ss.grads(global_tokens=global_tokens)
        if self.peer:
            ss.peer_allreduce()
        else:
            flat = ops.wrap_device_buffer(addr, n_sum + n_max, x_local.device)
            early = ss.early_elems() if self.comm_stream is not None else 0
            all_reduce_flat(flat, n_sum, n_max, self.group, early=early, comm_stream=self.comm_stream)
        self.device = x_local.device
        return ss.apply(adam_m, adam_v, step, lr, expansion_factor, optimizer, betas, eps=eps,
                        global_tokens=global_tokens, global_images=global_images)

    def check(self):
        """Raises if a peer never arrived at one of the exchanges so far (the parameters are undefined from that step
        on).  Synchronises with the device: call it per logging interval / before a checkpoint, not per step."""
        if self.peer and getattr(self, "device", None) is not None:
            from . import ops
            st = ops.comm_status(self.device)
            if st:
                raise RuntimeError(f"data-parallel exchange: rank {st - 1} never arrived at an all-reduce "
                                   f"(timeout; SVB_COMM_TIMEOUT_S)")
