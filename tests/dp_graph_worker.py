"""torchrun worker (2 GPUs): ModelPipeline(data_parallel=True) with the training batch captured in a CUDA graph against
the same pipeline launched eagerly -- every rank on its image shard, same global batch.  The SAE parameters, the dead
masks and the per-batch scalars must be bit-identical between the two runs and across the ranks."""
import collections
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    import sparse_vision_b200.models.sae_mlp as M
    from sparse_vision_b200.model_pipeline import ModelPipeline
    from sparse_vision_b200.parallel import shard_images
    C, k, B = 64, 4, 6 * world

    def make(graph):
        torch.manual_seed(0)
        base = nn.Sequential(collections.OrderedDict(
            conv=nn.Conv2d(3, C, 3, padding=1), act=nn.ReLU(), c2=nn.Conv2d(C, 32, 3, padding=1), r2=nn.ReLU(),
            gap=nn.AdaptiveAvgPool2d(1), fl=nn.Flatten(), fc=nn.Linear(32, 10))).eval().to(dev)
        sae = M.SaeMLP(C, k)
        with torch.no_grad():
            sae.encoder.bias[:9] = -50.0
        sae = sae.to(dev)
        pipe = ModelPipeline(base, sae, "sae_mlp", "act", "constrained_adam", 1e-3, 5.0, k, data_parallel=True,
                             global_batch_images=B, compare_in_one_pass=True, cuda_graph=graph)
        pipe.register_hooks(train_sae=True)
        return pipe, sae

    lo, hi = shard_images(B, rank, world)
    xs = [torch.randn(B, 3, 16, 16, generator=torch.Generator().manual_seed(7 + i))[lo:hi].to(dev) for i in range(9)]
    ys = [torch.randint(0, 10, (B,), generator=torch.Generator().manual_seed(70 + i))[lo:hi].to(dev) for i in range(9)]
    runs = []
    for graph in (False, True):
        pipe, sae = make(graph)
        log = []
        for x, y in zip(xs, ys):
            out, _ = pipe.train_batch(x, targets=y)
            log.append((pipe.batch_scalars(), out.clone(), pipe._last.dead.clone()))
        pipe.dp.check()
        runs.append((pipe, sae, log))
    (pe, se, le), (pg, sg, lg) = runs
    assert pe._graph is None and pg._graph is not None, "the data-parallel batch was not captured"
    for i, ((s0, o0, d0), (s1, o1, d1)) in enumerate(zip(le, lg)):
        assert s0 == s1, (i, s0, s1)
        assert torch.equal(o0, o1) and torch.equal(d0, d1), i
    for a, b in zip(se.param_list(), sg.param_list()):
        assert torch.equal(a, b)
        other = b.detach().clone()
        dist.broadcast(other, src=0)
        assert torch.equal(other, b.detach()), "replicas differ across ranks"
    dist.barrier()
    if rank == 0:
        print("dp graph parity ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
