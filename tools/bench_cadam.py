"""Development timing of the ConstrainedAdam decoder kernel (utils.py:65-81) on the decoder shapes of the reference."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_vision_b200 import ops

def main():
    dev = torch.device("cuda", 0)
    flush = torch.empty(2048 << 20, dtype=torch.uint8, device=dev)
    for C, F in [(256, 2048), (480, 1920), (512, 2048), (832, 3328), (1024, 4096), (512, 8192)]:
        w = torch.randn(C, F, device=dev); w /= w.norm(dim=0, keepdim=True)
        g = torch.randn(C, F, device=dev) * 1e-3
        m = torch.zeros_like(w); v = torch.zeros_like(w)
        for i in range(3):
            ops.adam_step([w], [g], [m], [v], i + 1, 1e-3, (0.9, 0.999), optimizer="constrained_adam", decoder_index=0)
        ts = []
        for i in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.adam_step([w], [g], [m], [v], i + 4, 1e-3, (0.9, 0.999), optimizer="constrained_adam", decoder_index=0)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print(f"C={C} F={F}: {t * 1e3:.1f} us  ({C * F * 28 / t * 1e-6:.0f} GB/s algorithmic, cold L2)", flush=True)

if __name__ == "__main__":
    main()
