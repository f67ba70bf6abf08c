"""Launches each libsvb producer kernel a few times at GoogLeNet's shapes (256 images) so that ncu can capture them:
    ncu --set full --clock-control none --import-source on -k regex:'maxpool|bias_relu|relu_grad|conv1_7x7' -c 30 \
        -o gpurun_out/producer_full python tools/producer_kernels_run.py
Also prints event-timed bandwidths (outside ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_vision_b200 import ops  # noqa: E402


def nhwc(*shape):
    return torch.randn(*shape, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)


def timed(fn, n=20, warm=3):
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
    B = int(os.environ.get("B", 256))
    quick = bool(int(os.environ.get("QUICK", "0")))       # under ncu: one launch each
    n = 1 if quick else 20
    cases = [("maxpool1 3/2 C=64 112x112", (B, 64, 112, 112), 3, 2, 0), ("maxpool2 3/2 C=192 56x56", (B, 192, 56, 56), 3, 2, 0),
             ("branch pool 3/1 C=256 28x28", (B, 256, 28, 28), 3, 1, 1), ("branch pool 3/1 C=512 14x14", (B, 512, 14, 14), 3, 1, 1),
             ("branch pool 3/1 C=832 7x7", (B, 832, 7, 7), 3, 1, 1), ("maxpool4 2/2 C=832 14x14", (B, 832, 14, 14), 2, 2, 0)]
    for name, shape, k, s, p in cases:
        x = nhwc(*shape)
        y = ops.maxpool_nhwc(x, k, s, p, True)
        t = timed(lambda: ops.maxpool_nhwc(x, k, s, p, True), n=n, warm=0 if quick else 3)
        gb = (x.numel() + y.numel()) * 2 / 1e9
        print(f"{name:32s} {t * 1e3:8.1f} us  {gb / t * 1e3:7.0f} GB/s (in + out)")
        del x, y
    for name, shape, chans in [("bias_relu conv3 C=192 56x56 in place", (B, 192, 56, 56), None),
                               ("bias_relu 3a merged 176 -> 64|96|16", (B, 176, 28, 28), (64, 96, 16)),
                               ("bias_relu 4c 3x3 256 -> concat 512", (B, 256, 14, 14), "concat")]:
        src = nhwc(*shape)
        bias = torch.randn(shape[1], device="cuda").bfloat16()
        if chans is None:
            dests = [(src, 0, shape[1])]
        elif chans == "concat":
            dests = [(nhwc(B, 512, 14, 14), 128, 256)]
        else:
            dests = [(nhwc(B, c, *shape[2:]), 0, c) for c in chans]
        t = timed(lambda: ops.bias_relu_scatter(src, bias, dests), n=n, warm=0 if quick else 3)
        gb = src.numel() * 2 * 2 / 1e9
        print(f"{name:40s} {t * 1e3:8.1f} us  {gb / t * 1e3:7.0f} GB/s (read + write)")
        del src, dests
    # the differentiable pair and the ReLU-mask gather of the IE passes (64 images per batch there)
    Bi = max(B // 4, 1)
    for name, shape, k, s, p in [("pool fwd+argmax 3/1 C=480 14x14", (Bi, 480, 14, 14), 3, 1, 1),
                                 ("pool fwd+argmax 3/2 C=480 28x28", (Bi, 480, 28, 28), 3, 2, 0)]:
        x = nhwc(*shape)
        y, arg = ops.maxpool_nhwc_with_argmax(x, k, s, p, True)
        go = torch.randn_like(y)
        t = timed(lambda: ops.maxpool_nhwc_with_argmax(x, k, s, p, True), n=n, warm=0 if quick else 3)
        tb = timed(lambda: ops.maxpool_nhwc_backward(go, arg, x.shape, k, s, p), n=n, warm=0 if quick else 3)
        gb = (x.numel() + y.numel()) * 2 / 1e9
        print(f"{name:32s} {t * 1e3:8.1f} us fwd {tb * 1e3:8.1f} us bwd  {gb / t * 1e3:7.0f} / {gb / tb * 1e3:7.0f} GB/s (in + out)")
        del x, y, arg, go
    out = nhwc(Bi, 512, 14, 14).relu_()
    go = nhwc(Bi, 512, 14, 14)
    t = timed(lambda: ops.relu_grad_gather([(go, 128, out, 128, 256)], out), n=n, warm=0 if quick else 3)
    print(f"relu_grad_gather 256 of 512 ch 14x14     {t * 1e3:8.1f} us")
    del out, go
    x = nhwc(B, 3, 224, 224)
    w = (torch.randn(64, 3, 7, 7, device="cuda") * 0.1).bfloat16()
    bias = torch.randn(64, device="cuda").bfloat16()
    packed = ops.conv1_pack_weights(w)
    t = timed(lambda: ops.conv1_stem(x, packed, bias), n=n, warm=0 if quick else 3)
    print(f"conv1 stem {t * 1e3:8.1f} us  {2.0 * B * 112 * 112 * 64 * 147 / t / 1e9:6.0f} useful TFLOP/s  "
          f"{(x.numel() + B * 64 * 112 * 112) * 2 / t / 1e6:6.0f} GB/s (in + out)")


if __name__ == "__main__":
    main()
