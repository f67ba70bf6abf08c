"""Where does the producer's time go?  Kernel-time table (torch.profiler) of the frozen GoogLeNet forward in the
format the e2e leg of bench.py runs it in (bf16, channels_last, BatchNorm folded, 256 images), plus event-timed
variants: cudnn.benchmark on / off, 3-channel vs padded stem.  Diagnostic only (not on the product path).

    python tools/prof_producer.py [--batch 256] [--benchmark 0|1] [--table 1]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_vision_b200.producer import synthetic_googlenet, to_producer_format  # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--benchmark", type=int, default=0)
    ap.add_argument("--table", type=int, default=1)
    ap.add_argument("--fuse", type=int, default=0, help="1: producer.fuse_forward (libsvb pool / bias+relu+concat kernels)")
    args = ap.parse_args()
    torch.backends.cudnn.benchmark = bool(args.benchmark)
    dev = torch.device("cuda:0")
    model = to_producer_format(synthetic_googlenet(seed=0), dev, torch.bfloat16, channels_last=True, fold_bn=True,
                               fuse=bool(args.fuse))
    x = torch.randn(args.batch, 3, 224, 224, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)

    with torch.no_grad():
        ms = timed(lambda: model(x))
        print(f"full forward B={args.batch} benchmark={args.benchmark} fuse={args.fuse}: {ms:.3f} ms")
        # per-stage times: run the children one after the other
        feats = x
        mods = [(n, m) for n, m in model.named_children() if not n.startswith("aux") and n not in ("dropout", "fc")]
        for name, m in mods:
            if name == "avgpool":
                break
            inp = feats
            t = timed(lambda: m(inp), n=10, warm=2)
            feats = m(inp)
            gb = (inp.numel() + feats.numel()) * 2 / 1e9
            print(f"  {name:12s} {t:7.3f} ms  in {tuple(inp.shape)} out {tuple(feats.shape)}  {gb:.3f} GB in+out "
                  f"-> {gb / (t * 1e-3):7.0f} GB/s")
        if args.fuse:
            from sparse_vision_b200 import ops
            c1 = model.conv1
            packed = ops.conv1_pack_weights(c1.conv.weight)
            t = timed(lambda: ops.conv1_stem(x, packed, c1.conv.bias), n=20)
            flop = 2.0 * args.batch * 112 * 112 * 64 * 147
            print(f"  conv1_stem kernel alone: {t:.3f} ms  ({flop / t / 1e9:.0f} useful TFLOP/s, "
                  f"{(x.numel() + args.batch * 64 * 112 * 112) * 2 / t / 1e6:.0f} GB/s in+out)")
        if args.table:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                model(x)
                torch.cuda.synchronize()
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=90))


if __name__ == "__main__":
    main()
