"""Times svb_node_ie_layer on the cfg5 / mixed3a shape (64 images, C=256, 28x28, F=2048) for NCHW and channels_last bf16
inputs, fused and un-fused.  Run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel breakdown."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_vision_b200 import _lib as L, ops  # noqa: E402

B, C, H, W, k = 64, 256, 28, 28, 8
F = C * k
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.relu(torch.randn(B, C, H, W, generator=g)).to(torch.bfloat16).to(dev)
gr = (torch.randn(B, C, H, W, generator=g) * 0.1).to(torch.bfloat16).to(dev)
params = [(torch.randn(F, C, generator=g) / 16).to(dev), torch.zeros(F, device=dev), (torch.randn(C, F, generator=g) / 45).to(dev),
          torch.zeros(C, device=dev)]
avg, err_avg, x_avg = torch.rand(F, H, W, device=dev), torch.zeros(C, H, W, device=dev), x.float().mean(0)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lib = L.load()
for fmt in ("nchw", "channels_last"):
    xs = x if fmt == "nchw" else x.contiguous(memory_format=torch.channels_last)
    gs = gr if fmt == "nchw" else gr.contiguous(memory_format=torch.channels_last)
    for fused in (1, 0):
        lib.svb_set_tuning(5, fused)
        for _ in range(3):
            ops.node_ie_layer(xs, gs, params, avg, err_avg, x_avg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.node_ie_layer(xs, gs, params, avg, err_avg, x_avg)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"node_ie_layer {fmt:13s} fused={fused}: {ms:.4f} ms per 64-image layer = {B / ms * 1e3:.0f} images/s")
lib.svb_set_tuning(5, 1)

