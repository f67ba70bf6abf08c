"""Kernel-time table (torch.profiler) of ONE ModelPipeline.train_batch as the e2e leg of bench.py runs it (frozen
GoogLeNet bf16 / channels_last / folded / fused forward, SaeMLP on inception3a, same-pass comparison with the
original model), launched eagerly so that every kernel is attributed.  Diagnostic only.

    python tools/prof_e2e.py [--batch 256] [--fuse 1]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_vision_b200.model_pipeline import ModelPipeline  # noqa: E402
from sparse_vision_b200.models.sae_mlp import SaeMLP  # noqa: E402
from sparse_vision_b200.producer import synthetic_googlenet, to_producer_format  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--fuse", type=int, default=1)
    ap.add_argument("--rows", type=int, default=40)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    base = to_producer_format(synthetic_googlenet(seed=0), dev, torch.bfloat16, channels_last=True, fold_bn=True,
                              fuse=bool(args.fuse))
    torch.manual_seed(0)
    sae = SaeMLP(256, 8).to(dev)
    pipe = ModelPipeline(base, sae, "sae_mlp", "inception3a", "constrained_adam", 1e-3, 5.0, 8, compare_in_one_pass=True,
                         cuda_graph=False)
    pipe.register_hooks(train_sae=True)
    x = torch.randn(args.batch, 3, 224, 224, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    tgt = torch.randint(0, 1000, (args.batch,), device=dev)
    for _ in range(3):
        pipe.train_batch(x, targets=tgt)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        pipe.train_batch(x, targets=tgt)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=args.rows, max_name_column_width=100))


if __name__ == "__main__":
    main()
