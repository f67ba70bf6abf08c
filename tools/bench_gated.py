"""Development timing of the GatedSae training step at cfg3 (C=512, 14x14, k=16, B=256 images per GPU)."""
import sys, os, time
import ctypes as C_
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sae_oracle as O
from sparse_vision_b200 import ops, _lib as L

def main():
    B, C, H, W, k = 256, 512, 14, 14, 16
    if len(sys.argv) > 1:
        B, C, H, W, k = [int(v) for v in sys.argv[1:6]]
    kind = sys.argv[6] if len(sys.argv) > 6 else "gated_sae"
    torch.manual_seed(0)
    p = O.init_gated_sae(C, k) if kind == "gated_sae" else O.init_sae_mlp(C, k)
    keys = O.GATED_KEYS if kind == "gated_sae" else O.SAE_MLP_KEYS
    dev = torch.device("cuda", 0)
    params = [p[kk].clone().to(dev) for kk in keys]
    ms = [torch.zeros_like(q) for q in params]
    vs = [torch.zeros_like(q) for q in params]
    xs = [torch.relu(torch.randn(B, C, H, W, device=dev)).bfloat16() for _ in range(2)]
    fn = ops.gated_train_step if kind == "gated_sae" else ops.sae_train_step
    lam = 0.1 if kind == "gated_sae" else 5.0
    for i in range(3):
        res = fn(xs[i % 2], params, ms, vs, i + 1, 1e-3, lam, k, optimizer="constrained_adam")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for i in range(n):
        res = fn(xs[i % 2], params, ms, vs, i + 4, 1e-3, lam, k, optimizer="constrained_adam")
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / n
    T, F = B * H * W, C * k
    flops = (12 if kind == "gated_sae" else 10) * C * F * T
    print(f"{kind} B={B} C={C} HW={H}x{W} k={k}: {ms_step:.3f} ms/step  {T / ms_step * 1e3 / 1e6:.1f} M act-vec/s  "
          f"{flops / ms_step * 1e-9:.0f} TFLOP/s algorithmic  stats={ {kk: round(v, 4) for kk, v in res.scalars().items() if kk in ('loss', 'rec', 'l1', 'aux')} }")

    lib, h = L.load(), L.handle(dev)
    L.check(lib.svb_profile_enable(h, 1), "svb_profile_enable")
    for i in range(5):
        fn(xs[i % 2], params, ms, vs, i + 20, 1e-3, lam, k, optimizer="constrained_adam")
    torch.cuda.synchronize()
    ph = (C_.c_float * 16)()
    n_ph, n_st = C_.c_int32(0), C_.c_int32(0)
    L.check(lib.svb_profile_read(h, 16, ph, C_.byref(n_ph), C_.byref(n_st)), "svb_profile_read")
    L.check(lib.svb_profile_enable(h, 0), "svb_profile_enable")
    print("    " + "  ".join(f"{lib.svb_profile_phase_name(i).decode()}={ph[i]:.3f}" for i in range(n_ph.value)), flush=True)

if __name__ == "__main__":
    main()
