import os, sys, torch
sys.path.insert(0, "/root/repo")
from oracle import sae_oracle as O
from sparse_vision_b200 import ops, _lib as L
B, C, S, k = 256, 256, 28, 8
dev = torch.device("cuda", 0)
torch.manual_seed(0)
p = O.init_sae_mlp(C, k)
params = [p[kk].clone().to(dev) for kk in O.SAE_MLP_KEYS]
ms = [torch.zeros_like(q) for q in params]; vs = [torch.zeros_like(q) for q in params]
xs = [torch.relu(torch.randn(B, C, S, S, device=dev)).bfloat16().contiguous(memory_format=torch.channels_last) for _ in range(2)]
lib, h = L.load(), L.handle(dev)
def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(n):
        ops.sae_train_step(xs[i % 2], params, ms, vs, i + 1, 1e-3, 5.0, k, optimizer="constrained_adam")
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
run(5)
for rep in range(3):
    lib.svb_profile_enable(h, 0); off = run(50)
    lib.svb_profile_enable(h, 1); on = run(50)
    print(f"profiler off {off:.4f} ms/step, on {on:.4f} ms/step")
