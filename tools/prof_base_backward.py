"""Forward + backward of the frozen GoogLeNet behind the first hooked layer (leaf at inception3a), channels_last against
NCHW, 64 images: event-timed and as a torch.profiler kernel table.  This is the measurement behind "with the stem out of
the backward, channels_last is the faster format" (IE passes).  Diagnostic only.

    python tools/prof_base_backward.py
"""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from sparse_vision_b200.producer import synthetic_googlenet, to_producer_format
dev = torch.device('cuda:0')
for cl in (True, False):
    m = to_producer_format(synthetic_googlenet(seed=0), dev, torch.bfloat16, channels_last=cl, fold_bn=True)
    x = torch.randn(64, 3, 224, 224, device=dev).bfloat16()
    if cl: x = x.contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (64,), device=dev)
    leaf = {}
    def hook(_m, _i, out):
        leaf['x'] = out.detach().requires_grad_(True); return leaf['x']
    h = m.inception3a.register_forward_hook(hook)
    def step():
        out = m(x); torch.nn.functional.cross_entropy(out.float(), y).backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    print('channels_last' if cl else 'nchw', e0.elapsed_time(e1) / 5, 'ms per fwd+bwd')
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=90))
    h.remove()
