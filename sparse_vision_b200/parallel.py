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


_SYMM_KEEPALIVE = {}   # device index -> (symmetric tensor, rendezvous handle): the region must outlive the library's use of it


def connect_symmetric_memory(device, n_floats, group=None):
    """Sets up the exchange buffer in torch SYMMETRIC memory (torch.distributed._symmetric_memory does the VMM allocation,
    the file-descriptor exchange between the ranks and the multicast binding) and hands every rank's mapping plus the
    multicast pointer to the library (svb_comm_attach): with a multicast pointer the SUM section of the all-reduce goes
    through the NVSwitch (multimem.ld_reduce / multimem.st, "NVLS").  Returns 'nvls', 'symm' (symmetric memory without
    multicast: plain peer loads / stores) or None when it is unavailable on ANY rank -- the caller then falls back to CUDA
    IPC (connect_peer_memory).  SVB_DP_NVLS=0 skips it."""
    import os
    from . import _lib as L
    if os.environ.get("SVB_DP_NVLS", "1") == "0":
        return None
    lib, h = L.load(), L.handle(device)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    key = device.index if device.index is not None else torch.cuda.current_device()
    cap = ctypes.c_int64()
    lib.svb_comm_capacity(h, ctypes.byref(cap))
    if cap.value >= n_floats:      # an earlier step on this device already connected a large enough buffer
        return _SYMM_KEEPALIVE[key][2] if key in _SYMM_KEEPALIVE else None    # None: it is an IPC buffer, keep it
    if cap.value > 0:
        lib.svb_comm_destroy(h)
        _SYMM_KEEPALIVE.pop(key, None)
    nbytes, capf = ctypes.c_int64(), ctypes.c_int64()
    L.check(lib.svb_comm_region_bytes(int(n_floats), ctypes.byref(nbytes), ctypes.byref(capf)), "svb_comm_region_bytes")
    ok, t, hd, ptrs, mc = world <= 8, None, None, None, 0
    if ok:
        try:
            import torch.distributed._symmetric_memory as symm
            with torch.cuda.device(device):
                t = symm.empty(nbytes.value // 4, dtype=torch.float32, device=device)
                t.zero_()
                torch.cuda.synchronize(device)
                hd = symm.rendezvous(t, (group if group is not None else dist.group.WORLD).group_name)
            ptrs = [int(p) for p in hd.buffer_ptrs]
            mc = int(hd.multicast_ptr or 0)
            off = t.data_ptr() - ptrs[rank]          # the tensor may sit at an offset inside its symmetric block
            ptrs = [p + off for p in ptrs]
            mc = mc + off if mc else 0
            ok = len(ptrs) == world and off >= 0
        except Exception:
            ok = False
    flags = torch.tensor([1 if ok else 0, 1 if mc else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN, group=group)      # also: every rank has zeroed its region by now
    ok, use_mc = bool(flags[0].item()), bool(flags[1].item())
    if ok:
        arr = (ctypes.c_void_p * world)(*ptrs)
        ok = lib.svb_comm_attach(h, rank, world, int(n_floats), arr, ctypes.c_void_p(mc if use_mc else 0)) == 0
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if not bool(flag.item()):
        lib.svb_comm_destroy(h)
        return None
    mode = "nvls" if use_mc else "symm"
    _SYMM_KEEPALIVE[key] = (t, hd, mode)
    return mode


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

    exchange = 'peer' (default on CUDA): the flat buffer lives in memory every rank maps and ONE kernel of the library
    reduces it over NVLink (svb_comm_allreduce) -- in torch symmetric memory with the SUM section reduced by the NVSwitch
    (`mode` 'nvls') when that is available, else in CUDA-IPC peer memory (`mode` 'ipc'); falls back to 'nccl' when
    neither can be set up.
    exchange = 'nccl': torch.distributed all-reduce (also the CPU / gloo path); overlap=True additionally sends the
    encoder-side gradients on a communication stream while the decoder weight-gradient GEMM runs."""

    def __init__(self, kind, group=None, exchange="peer", overlap=False):
        self.kind = kind
        self.group = group
        self.exchange = exchange
        self.overlap = overlap
        self.comm_stream = None
        self.peer = None   # None: not tried yet, True / False afterwards
        self.mode = None   # 'nvls' | 'symm' | 'ipc' once the peer exchange is connected

    def step(self, x_local, params, adam_m, adam_v, step, lr, lam, expansion_factor, optimizer, betas,
             global_images, global_tokens, want_dec=True, eps=1e-8, step_dev=None, dec_out=None):
        from . import ops
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        if self.overlap and self.exchange != "peer" and multi and x_local.is_cuda and self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=x_local.device)
            ops.set_comm_stream(x_local.device, self.comm_stream)
        ss = ops.SplitStep(self.kind, x_local, params, lam, want_dec=want_dec, dec_out=dec_out)
        addr, n_sum, n_max = ss.grads(global_tokens=global_tokens)
        if multi and x_local.is_cuda and self.exchange == "peer" and self.peer is None:
            self.mode = connect_symmetric_memory(x_local.device, n_sum + n_max, self.group)
            if self.mode is None and connect_peer_memory(x_local.device, n_sum + n_max, self.group):
                self.mode = "ipc"
            self.peer = self.mode is not None
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
                        global_tokens=global_tokens, global_images=global_images, step_dev=step_dev)

    def check(self):
        """Raises if a peer never arrived at one of the exchanges so far (the parameters are undefined from that step
        on).  Synchronises with the device: call it per logging interval / before a checkpoint, not per step."""
        if self.peer and getattr(self, "device", None) is not None:
            from . import ops
            st = ops.comm_status(self.device)
            if st:
                raise RuntimeError(f"data-parallel exchange: rank {st - 1} never arrived at an all-reduce "
                                   f"(timeout; SVB_COMM_TIMEOUT_S)")
