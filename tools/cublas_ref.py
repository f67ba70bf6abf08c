# Reference point for the K=256 output-bound GEMM: what does cuBLAS (torch.matmul, bf16) reach on the same shapes?
import torch, time
torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = True
dev = "cuda"
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
T, C, F = 200704, 256, 2048
X = torch.randn(T, C, device=dev, dtype=torch.bfloat16)
We = torch.randn(F, C, device=dev, dtype=torch.bfloat16)
Wd = torch.randn(C, F, device=dev, dtype=torch.bfloat16)
b = torch.randn(F, device=dev, dtype=torch.bfloat16)
E = torch.empty(T, F, device=dev, dtype=torch.bfloat16)
D = torch.empty(T, C, device=dev, dtype=torch.bfloat16)
ms = bench(lambda: torch.matmul(X, We.t(), out=E)); print(f"enc  X@We^T        {ms:.3f} ms {2*T*C*F/ms*1e-9:.0f} TFLOP/s")
ms = bench(lambda: torch.addmm(b, X, We.t(), out=E)); print(f"enc  addmm(bias)   {ms:.3f} ms {2*T*C*F/ms*1e-9:.0f} TFLOP/s")
ms = bench(lambda: torch.matmul(E, Wd.t(), out=D)); print(f"dec  E@Wd^T        {ms:.3f} ms {2*T*C*F/ms*1e-9:.0f} TFLOP/s")
ms = bench(lambda: torch.matmul(D, Wd, out=E)); print(f"dE   D@Wd          {ms:.3f} ms {2*T*C*F/ms*1e-9:.0f} TFLOP/s")
G = torch.empty(C, F, device=dev, dtype=torch.bfloat16)
ms = bench(lambda: torch.matmul(D.t(), E, out=G)); print(f"dWd  D^T@E         {ms:.3f} ms {2*T*C*F/ms*1e-9:.0f} TFLOP/s")
G2 = torch.empty(F, C, device=dev, dtype=torch.bfloat16)
ms = bench(lambda: torch.matmul(E.t(), X, out=G2)); print(f"dWe  E^T@X         {ms:.3f} ms {2*T*C*F/ms*1e-9:.0f} TFLOP/s")
ms = bench(lambda: E.copy_(E)); print(f"copy 822MB r+w     {ms:.3f} ms {2*E.numel()*2/ms*1e-6:.0f} GB/s")
ms = bench(lambda: E.fill_(1.0)); print(f"fill 822MB         {ms:.3f} ms {E.numel()*2/ms*1e-6:.0f} GB/s")
