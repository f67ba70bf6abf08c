"""torchrun worker for the 2-GPU data-parallel parity test: two ranks, each with half of the images, must produce
the same parameters / stats as one rank stepping on the whole batch (up to fp32 summation order)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sae_oracle as O  # noqa: E402
from sparse_vision_b200 import ops  # noqa: E402
from sparse_vision_b200.parallel import DataParallelStep, shard_images  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cases = [(kind, keys, init, exch) for exch in ("peer", "nccl")
             for kind, keys, init in (("sae_mlp", O.SAE_MLP_KEYS, O.init_sae_mlp),
                                      ("gated_sae", O.GATED_KEYS, O.init_gated_sae))]
    for kind, keys, init, exch in cases:
        C, k, B, H, W = 64, 4, 6, 7, 7
        torch.manual_seed(0)
        p = init(C, k)
        p["decoder.bias"].normal_(0, 0.05, generator=torch.Generator().manual_seed(2))
        x = torch.relu(torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(11))).bfloat16()
        lo, hi = shard_images(B, rank, world)
        params = [p[kk].clone().to(dev) for kk in keys]
        ms = [torch.zeros_like(q) for q in params]
        vs = [torch.zeros_like(q) for q in params]
        # 'peer': the library's own all-reduce kernel over CUDA-IPC peer memory; 'nccl': torch.distributed, with the
        # encoder-side gradients sent early on a communication stream
        dp = DataParallelStep(kind, exchange=exch, overlap=(exch == "nccl"))
        for step in (1, 2, 3):     # several steps: the flag epochs of the peer exchange must keep working
            res = dp.step(x[lo:hi].to(dev), params, ms, vs, step, 1e-3, 0.5, k, "constrained_adam", (0.9, 0.999), B,
                          B * H * W)
        if exch == "peer":
            assert dp.peer, "peer-memory exchange could not be set up on this box"
            if rank == 0:
                print(f"{kind}: peer exchange mode = {dp.mode}", flush=True)
        got = res.scalars()
        if rank == 0:
            ref_params = [p[kk].clone().to(dev) for kk in keys]
            rm = [torch.zeros_like(q) for q in ref_params]
            rv = [torch.zeros_like(q) for q in ref_params]
            fn = ops.sae_train_step if kind == "sae_mlp" else ops.gated_train_step
            ops.set_comm_stream(dev, None)
            for step in (1, 2, 3):
                ref = fn(x.to(dev), ref_params, rm, rv, step, 1e-3, 0.5, k, optimizer="constrained_adam")
            want = ref.scalars()
            for key in ("loss", "rec", "l1", "nrmse", "rmse", "aux", "var_expl", "sparsity", "n_dead"):
                assert abs(got[key] - want[key]) <= 1e-4 * max(abs(want[key]), 1e-3), (kind, key, got[key], want[key])
            assert torch.equal(res.dead, ref.dead), kind
            for a, b, kk in zip(params, ref_params, keys):
                d = (a - b).abs().max().item()
                assert d <= 6.3e-3, (kind, kk, d)            # a sign flip of a ~0 gradient moves a weight by 2*lr per step
                assert (a - b).abs().mean().item() <= 2e-5, (kind, kk)
        # replicas must stay bit-identical across ranks
        for q in params:
            other = q.clone()
            dist.broadcast(other, src=0)
            assert torch.equal(other, q), (kind, exch, "replicas diverged")
        ops.set_comm_stream(dev, None)
    dist.barrier()
    if rank == 0:
        print("dp parity ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
