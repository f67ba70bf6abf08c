"""torchrun worker for the 2-GPU attribution test: two ranks, each with half of the images of every batch, must give
the same averages and node-IE values as one process on the whole batches (compute_ie.py:95-226, :365-472; the
reference's sample-weighted running mean equals the global mean)."""
import collections
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparse_vision_b200.compute_ie import IE  # noqa: E402
from sparse_vision_b200.models.sae_mlp import SaeMLP  # noqa: E402
from sparse_vision_b200.parallel import shard_images  # noqa: E402


def build(dev):
    torch.manual_seed(3)
    net = nn.Sequential(collections.OrderedDict(
        c1=nn.Conv2d(3, 64, 3, padding=1), r1=nn.ReLU(), p1=nn.MaxPool2d(2),
        c2=nn.Conv2d(64, 128, 3, padding=1), r2=nn.ReLU(), p2=nn.MaxPool2d(2),
        gap=nn.AdaptiveAvgPool2d(1), fl=nn.Flatten(), fc=nn.Linear(128, 10))).eval().to(dev)
    names = {"r1": (64, 4), "r2": (128, 4)}
    torch.manual_seed(5)
    saes = {n: SaeMLP(c, k).to(dev) for n, (c, k) in names.items()}
    mods = dict(net.named_modules())
    return IE(net, {n: mods[n] for n in names}, saes, {n: k for n, (_, k) in names.items()}), names


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B = 6     # batches of 6 and 5 images: the second one shards unevenly (3 + 2)
    batches = [(torch.randn(B - i, 3, 16, 16, generator=torch.Generator().manual_seed(40 + i)),
                torch.randint(0, 10, (B - i,), generator=torch.Generator().manual_seed(50 + i))) for i in range(2)]
    # single-process result on the whole batches, BEFORE the process group exists (no collectives)
    ie, names = build(dev)
    avg1 = ie.compute_average([x for x, _ in batches])
    f1, e1, n1 = ie.compute_node_ie(batches, avg1)

    dist.init_process_group("nccl", device_id=dev)
    ie, names = build(dev)
    local_batches = []
    for x, y in batches:
        lo, hi = shard_images(x.shape[0], rank, world)
        local_batches.append((x[lo:hi], y[lo:hi]))
    avg2 = ie.compute_average([x for x, _ in local_batches])
    f2, e2, n2 = ie.compute_node_ie(local_batches, avg2)
    for n in names:
        for key in ("encoder_output_average", "sae_error_average", "original_layer_output_average"):
            a, b = avg2[key][n].float(), avg1[key][n].float()
            assert (a - b).norm() <= 1e-5 * b.norm() + 1e-7, (n, key, float((a - b).norm()), float(b.norm()))
        assert torch.equal(avg2["dead_units"][n], avg1["dead_units"][n]), n
        assert abs(avg2["sparsity"][n] - avg1["sparsity"][n]) <= 1e-6 * max(abs(avg1["sparsity"][n]), 1e-6), n
        # the averages differ in the last bits (summation order), so the IE values agree to ~1e-3, not bit for bit
        assert (f2[n] - f1[n]).norm() <= 5e-3 * f1[n].norm(), (n, "features")
        assert abs(float(e2[n]) - float(e1[n])) <= 5e-3 * abs(float(e1[n])), (n, "error")
        assert (n2[n] - n1[n]).norm() <= 5e-3 * n1[n].norm(), (n, "neurons")
        # every rank ends with the same global result
        other = f2[n].clone()
        dist.broadcast(other, src=0)
        assert torch.equal(other, f2[n]), (n, "ranks disagree")
    dist.barrier()
    if rank == 0:
        print("dp ie parity ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
