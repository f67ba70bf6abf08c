"""Kernel-time table (torch.profiler) of ONE IE.compute_average batch (three GoogLeNet layers, 64 images, eager launches):
where the averages pass spends its time (this is what pointed at svb_measure_inactive's many-row path).  Diagnostic only.

    python tools/prof_ie_average.py
"""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from sparse_vision_b200.compute_ie import IE
from sparse_vision_b200.models.sae_mlp import SaeMLP
from sparse_vision_b200.producer import GOOGLENET_LAYERS, hooked_layers, synthetic_googlenet, to_attribution_format
IE_LAYERS = {"mixed3a": 8, "mixed4c": 4, "mixed5b": 4}
dev = torch.device("cuda:0")
base = to_attribution_format(synthetic_googlenet(seed=0), dev, "mixed3a", torch.bfloat16)
saes = {}
for j, (n, k) in enumerate(IE_LAYERS.items()):
    torch.manual_seed(5 + j); saes[n] = SaeMLP(GOOGLENET_LAYERS[n][1], k).to(dev)
x = torch.randn(64, 3, 224, 224, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
ie = IE(base, hooked_layers(base, list(IE_LAYERS)), saes, dict(IE_LAYERS), device=dev)
for _ in range(3): ie.compute_average([x])
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ie.compute_average([x]); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=90))
