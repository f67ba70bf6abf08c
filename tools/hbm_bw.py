"""Development probe: write-only / copy / read-only HBM bandwidth of plain torch kernels (1 GiB, fp32 elements)."""
import torch
dev = torch.device("cuda", 0)
n = 1 << 28          # fp32 elements = 1 GiB
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)


def t(f, it=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


ms = t(lambda: a.zero_()); print("fill 1 GiB (fp32): %.3f ms  %.0f GB/s write" % (ms, 4 * n / ms * 1e-6))
ms = t(lambda: torch.cuda.memset if False else a.fill_(1.5)); print("fill_(1.5) 1 GiB: %.3f ms  %.0f GB/s write" % (ms, 4 * n / ms * 1e-6))
ms = t(lambda: b.copy_(a)); print("copy 1 GiB: %.3f ms  %.0f GB/s read+write" % (ms, 8 * n / ms * 1e-6))
ms = t(lambda: a.sum()); print("sum 1 GiB: %.3f ms  %.0f GB/s read" % (ms, 4 * n / ms * 1e-6))
