import torch, sys, os
sys.path.insert(0, '.')
from oracle import sae_oracle as O
from sparse_vision_b200 import ops
B,C,H,W,k = 3,64,12,12,4
torch.manual_seed(0)
p = O.init_sae_mlp(C, k)
x = torch.relu(torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(1234))).bfloat16().float()
params = [p[key].clone().cuda() for key in O.SAE_MLP_KEYS]
ms = [torch.zeros_like(q) for q in params]; vs = [torch.zeros_like(q) for q in params]
try:
    res = ops.sae_train_step(x.cuda().bfloat16(), params, ms, vs, 1, 1e-3, 5.0, k, optimizer="constrained_adam")
    torch.cuda.synchronize()
    st = O.new_adam_state(p, O.SAE_MLP_KEYS)
    ref = O.train_step("sae_mlp", p, st, x, 5.0, "constrained_adam", 1e-3, k)
    sc = res.scalars()
    print({k_: (round(sc[k_],5), round(float(ref[k_]),5)) for k_ in ("loss","rec","l1","var_expl","rmse","nrmse")})
    d = res.dec.float().cpu(); r = ref["dec"]
    print("dec relerr", float((d-r).norm()/r.norm()), "max abs", float((d-r).abs().max()))
except Exception as e:
    print("FAILED:", str(e)[:200])
