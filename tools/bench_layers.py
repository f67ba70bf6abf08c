"""Development timing of the SaeMLP training step on the InceptionV1 layer shapes the reference trains SAEs on
(utils.py:2662-2741: mixed3a k=8, the others k=4; torchvision GoogLeNet as utils.py:280 loads it -- mixed4a/4b/4c all
have 512 channels), B=256 images per GPU, bf16 NCHW input."""
import sys, os
import ctypes as C_
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sae_oracle as O
from sparse_vision_b200 import ops, _lib as L

LAYERS = [("mixed3a", 256, 28, 8), ("mixed3b", 480, 28, 4), ("mixed4a", 512, 14, 4),
          ("mixed4d", 528, 14, 4), ("mixed4e", 832, 14, 4), ("mixed5a", 832, 7, 4), ("mixed5b", 1024, 7, 4)]

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dt = torch.float32 if len(sys.argv) > 2 and sys.argv[2] == "fp32" else torch.bfloat16   # activation dtype (in and out)
    only = sys.argv[3] if len(sys.argv) > 3 else None
    dev = torch.device("cuda", 0)
    for name, C, S, k in LAYERS:
        if only and name != only:
            continue
        torch.manual_seed(0)
        p = O.init_sae_mlp(C, k)
        params = [p[kk].clone().to(dev) for kk in O.SAE_MLP_KEYS]
        ms = [torch.zeros_like(q) for q in params]
        vs = [torch.zeros_like(q) for q in params]
        xs = [torch.relu(torch.randn(B, C, S, S, device=dev)).to(dt) for _ in range(2)]
        try:
            for i in range(3):
                ops.sae_train_step(xs[i % 2], params, ms, vs, i + 1, 1e-3, 0.1, k, optimizer="constrained_adam")
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record()
            for i in range(n):
                ops.sae_train_step(xs[i % 2], params, ms, vs, i + 4, 1e-3, 0.1, k, optimizer="constrained_adam")
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / n
            T, F = B * S * S, C * k
            print(f"{name}: C={C} {S}x{S} F={F} T={T}: {t:.3f} ms/step  {T / t * 1e3 / 1e6:.1f} M act-vec/s  "
                  f"{10 * C * F * T / t * 1e-9:.0f} TFLOP/s", flush=True)
            # per-phase times from the library's CUDA-event profiler (adds a few us per phase)
            lib, h = L.load(), L.handle(dev)
            L.check(lib.svb_profile_enable(h, 1), "svb_profile_enable")
            for i in range(5):
                ops.sae_train_step(xs[i % 2], params, ms, vs, i + 20, 1e-3, 0.1, k, optimizer="constrained_adam")
            torch.cuda.synchronize()
            ph = (C_.c_float * 16)()
            n_ph, n_st = C_.c_int32(0), C_.c_int32(0)
            L.check(lib.svb_profile_read(h, 16, ph, C_.byref(n_ph), C_.byref(n_st)), "svb_profile_read")
            L.check(lib.svb_profile_enable(h, 0), "svb_profile_enable")
            print("    " + "  ".join(f"{lib.svb_profile_phase_name(i).decode()}={ph[i]:.3f}" for i in range(n_ph.value)), flush=True)
        except Exception as exc:  # report and go on to the next layer
            print(f"{name}: C={C} {S}x{S}: FAILED {type(exc).__name__}: {exc}", flush=True)

if __name__ == "__main__":
    main()
