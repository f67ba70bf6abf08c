"""Kernel-time table (torch.profiler) of ONE IE.compute_node_ie batch as bench.py's ie_pipeline section runs it
(three GoogLeNet layers, 64 images, base model from producer.to_attribution_format).  Diagnostic only.

    python tools/prof_ie.py [--batch 64] [--rows 40]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_vision_b200.compute_ie import IE  # noqa: E402
from sparse_vision_b200.models.sae_mlp import SaeMLP  # noqa: E402
from sparse_vision_b200.producer import (GOOGLENET_LAYERS, hooked_layers, synthetic_googlenet,  # noqa: E402
                                         to_attribution_format)

IE_LAYERS = {"mixed3a": 8, "mixed4c": 4, "mixed5b": 4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--rows", type=int, default=40)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    base = to_attribution_format(synthetic_googlenet(seed=0), dev, "mixed3a", torch.bfloat16)
    saes = {}
    for j, (n, k) in enumerate(IE_LAYERS.items()):
        torch.manual_seed(5 + j)
        saes[n] = SaeMLP(GOOGLENET_LAYERS[n][1], k).to(dev)
    x = torch.randn(args.batch, 3, 224, 224, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (args.batch,), device=dev)
    ie = IE(base, hooked_layers(base, list(IE_LAYERS)), saes, dict(IE_LAYERS), device=dev)
    avg = ie.compute_average([x])
    for _ in range(3):
        ie.compute_node_ie([(x, y)], avg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ie.compute_node_ie([(x, y)], avg)
    torch.cuda.synchronize()
    print(f"wall per batch: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ie.compute_node_ie([(x, y)], avg)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=args.rows, max_name_column_width=100))


if __name__ == "__main__":
    main()
