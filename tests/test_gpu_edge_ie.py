"""Edge indirect effects and circuit faithfulness (compute_ie.py:476-711, :715-944) on a small
frozen classifier with three hooked layers: the GPU path (one forward cut into segments, closed-form cotangents, one
vector-Jacobian product per downstream node, svb_node_ie_layer upstream) against a LITERAL autograd restatement of the
reference's procedure on the CPU (oracle.edge_ie_pass / faithfulness_pass)."""
import collections
import os
import sys

import numpy as np
import pytest
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sae_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu

NAMES = {"r1": (64, 4), "r2": (128, 4), "r3": (256, 2)}      # layer -> (channels, expansion factor)
FEATS = {"r1": [3, 70, 200, 11], "r2": [5, 100, 300], "r3": [7, 8, 400, 20, 33]}


def _net():
    torch.manual_seed(3)
    return nn.Sequential(collections.OrderedDict(
        c1=nn.Conv2d(3, 64, 3, padding=1), r1=nn.ReLU(), p1=nn.MaxPool2d(2),
        c2=nn.Conv2d(64, 128, 3, padding=1), r2=nn.ReLU(), p2=nn.MaxPool2d(2),
        c3=nn.Conv2d(128, 256, 3, padding=1), r3=nn.ReLU(),
        gap=nn.AdaptiveAvgPool2d(1), fl=nn.Flatten(), fc=nn.Linear(256, 10))).eval()


def _forward_from(net):
    mods = list(net.named_children())

    def run(name, x):
        start = 0 if name is None else [n for n, _ in mods].index(name) + 1
        outs = {}
        for n, m in mods[start:]:
            x = m(x)
            if n in NAMES:
                outs[n] = x
        outs["out"] = x
        return outs
    return run


def _run_with(net):
    def run(inputs, fn):
        hs = [m.register_forward_hook(lambda _m, _i, o, n=n: fn(n, o)) for n, m in net.named_children() if n in NAMES]
        try:
            return net(inputs)
        finally:
            for h in hs:
                h.remove()
    return run


def _setup():
    from sparse_vision_b200.models.sae_mlp import SaeMLP
    net = _net()
    for q in net.parameters():
        q.requires_grad = False
    torch.manual_seed(5)
    saes = {n: SaeMLP(c, k) for n, (c, k) in NAMES.items()}
    cpu_p = {n: {key: v.detach().clone() for key, v in s.state_dict().items()} for n, s in saes.items()}
    B = 6
    batches = [(torch.randn(B, 3, 16, 16, generator=torch.Generator().manual_seed(40 + i)),
                torch.randint(0, 10, (B,), generator=torch.Generator().manual_seed(50 + i))) for i in range(2)]
    # averages from the oracle (both sides use the same ones: this test is about the edge / faithfulness passes)
    run = _forward_from(net)
    avg, seen = {}, 0
    with torch.no_grad():
        for x, _ in batches:
            seen += B
            acts = run(None, x)
            for n, (_, k) in NAMES.items():
                la = O.layer_averages(cpu_p[n], acts[n], k)
                avg[n] = {q: la[q] for q in ("enc_avg", "err_avg", "x_avg")} if n not in avg else \
                    {q: O.running_mean_update(avg[n][q], la[q], seen, B) for q in ("enc_avg", "err_avg", "x_avg")}
    return net, saes, cpu_p, batches, avg


def _gpu_ie(net, saes, avg):
    from sparse_vision_b200.compute_ie import IE
    net_g = net.cuda()
    ie = IE(net_g, {n: dict(net_g.named_modules())[n] for n in NAMES}, {n: s.cuda() for n, s in saes.items()},
            {n: k for n, (_, k) in NAMES.items()})
    averages = {"encoder_output_average": {n: avg[n]["enc_avg"].cuda() for n in NAMES},
                "sae_error_average": {n: avg[n]["err_avg"].cuda() for n in NAMES},
                "original_layer_output_average": {n: avg[n]["x_avg"].cuda() for n in NAMES}}
    return ie, averages


def test_edge_ie_matches_literal_autograd_restatement():
    net, saes, cpu_p, batches, avg = _setup()
    layers = list(NAMES)
    ref = O.edge_ie_pass(layers, cpu_p, FEATS, {n: avg[n]["enc_avg"] for n in NAMES}, {n: avg[n]["err_avg"] for n in NAMES},
                         _forward_from(net), nn.CrossEntropyLoss(), batches)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ie, averages = _gpu_ie(net, saes, avg)
        got = ie.compute_edge_ie(batches, averages, layers, FEATS)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        net.cpu()
    for i, nu in enumerate(layers):
        g, r = got[nu].cpu(), ref[nu]
        n_d = len(FEATS[layers[i + 1]]) if i + 1 < len(layers) else 0
        assert g.shape == r.shape == (len(FEATS[nu]) + 1, n_d + 1)
        assert float(r.abs().max()) > 0
        for col in range(r.shape[1]):            # every downstream node separately: their scales differ by orders of magnitude
            assert (g[:, col] - r[:, col]).norm() <= 3e-2 * r[:, col].norm() + 1e-12, (nu, col, g[:, col], r[:, col])
        # the strongest edge into every downstream node is the same one
        assert torch.equal(g[:-1].argmax(dim=0), r[:-1].argmax(dim=0)) or \
            all(abs(r[g[:-1, c].argmax(), c] - r[:-1, c].max()) <= 3e-2 * r[:-1, c].max() for c in range(r.shape[1]))


@pytest.mark.parametrize("mode", ["sae", "model"])
def test_faithfulness_matches_oracle(mode):
    net, saes, cpu_p, batches, avg = _setup()
    layers = list(NAMES)
    # node IEs from the oracle (any fixed values would do; these are the real ones)
    run = _forward_from(net)
    ie_ref = {}
    for x, y in batches[:1]:
        leaves = {}
        xx, prev = x, None
        for n in layers:
            leaves[n] = run(prev, xx)[n].detach().requires_grad_(True)
            xx, prev = leaves[n], n
        nn.CrossEntropyLoss()(run(prev, xx)["out"], y).backward()
        # gradients w.r.t. the last layer only reach it; chain the rest
        grads = {layers[-1]: leaves[layers[-1]].grad}
        for i in range(len(layers) - 2, -1, -1):
            od = run(layers[i], leaves[layers[i]])[layers[i + 1]]
            grads[layers[i]] = torch.autograd.grad(od, leaves[layers[i]], grad_outputs=grads[layers[i + 1]])[0]
        for n in layers:
            ie_ref[n] = O.node_ie_layer(cpu_p[n], leaves[n].detach(), grads[n], avg[n]["enc_avg"], avg[n]["err_avg"],
                                        avg[n]["x_avg"])
    ie_feat = {n: v[0] for n, v in ie_ref.items()}
    ie_err = {n: v[1] for n, v in ie_ref.items()}
    ie_neur = {n: v[2] for n, v in ie_ref.items()}
    thr = float(torch.cat([v.flatten() for v in ie_feat.values()]).quantile(0.7))      # ~30 % of the features stay
    want = O.faithfulness_pass(layers, cpu_p, {n: avg[n]["enc_avg"] for n in NAMES}, {n: avg[n]["err_avg"] for n in NAMES},
                               {n: avg[n]["x_avg"] for n in NAMES}, ie_feat, ie_err, ie_neur, _run_with(net),
                               nn.CrossEntropyLoss(), batches, thr, model_or_sae=mode)
    ie, averages = _gpu_ie(net, saes, avg)
    try:
        got = ie.compute_faithfulness(batches, averages, ({n: v.cuda() for n, v in ie_feat.items()},
                                                          {n: v.cuda() for n, v in ie_err.items()},
                                                          {n: v.cuda() for n, v in ie_neur.items()}), thr, model_or_sae=mode)
    finally:
        net.cpu()
    keys = ["m_C", "m_empty", "m_M"] + (["m_C_zero", "m_C_mean"] if mode == "sae" else [])
    for k in keys:
        assert abs(got[k] - want[k]) <= 1e-2 * abs(want[k]), (k, got[k], want[k])
    fk = ["faithfulness"] + (["faithfulness_sae_errors_zero_ablated", "faithfulness_sae_errors_mean_ablated"] if mode == "sae" else [])
    den = abs(want["m_M"] - want["m_empty"])
    for k in fk:      # a ratio of loss differences: the 1e-2 loss tolerance is amplified by |m| / |m(M) - m(empty)|
        assert abs(got[k] - want[k]) <= 2e-2 * max(abs(want[k]), 1.0) * max(1.0, abs(want["m_M"]) / den * 0.1), (k, got[k], want[k])
    if mode == "sae":
        assert got["nodes_in_circuit"] == {n: int((ie_feat[n].abs() > thr).sum()) for n in layers}
