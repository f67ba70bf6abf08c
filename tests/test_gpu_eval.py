"""Eval-epoch metrics on the device (model_pipeline.py:335-360, utils.py:1445-1481, :1934-1963,
:1996-2010): spatial means, per-batch top-k / small-k with bit-exact indices, the running merge against the fixtures
the real reference produced, histograms against torch.histc, and the whole eval batch against the CPU oracle."""
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


@pytest.mark.parametrize("B,F,H,W", [(6, 72, 5, 7), (3, 2048, 28, 28), (9, 40, 1, 1), (2, 256, 14, 14)])
def test_spatial_mean_matches_torch(B, F, H, W):
    from sparse_vision_b200 import ops, utils as U
    t = torch.randn(B, F, H, W, generator=torch.Generator().manual_seed(1)).cuda()
    want = t.mean(dim=(2, 3))
    np.testing.assert_allclose(ops.spatial_mean(t).cpu().numpy(), want.cpu().numpy(), rtol=1e-5, atol=1e-6)
    tok = t.permute(0, 2, 3, 1).reshape(-1, F).contiguous()
    if F % 8 == 0:
        got = ops.spatial_mean(tok, n_images=B)
        np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-5, atol=1e-6)
        got16 = ops.spatial_mean(tok.bfloat16(), n_images=B)
        want16 = tok.bfloat16().float().reshape(B, H * W, F).mean(1)
        np.testing.assert_allclose(got16.cpu().numpy(), want16.cpu().numpy(), rtol=1e-5, atol=1e-6)
        # average_over_W_H on the [B,F,H,W] VIEW of a token-major tensor takes the zero-copy route
        view = tok.reshape(B, H, W, F).permute(0, 3, 1, 2)
        a, b = U.average_over_W_H(view, None)
        assert b is None
        np.testing.assert_allclose(a.cpu().numpy(), want.cpu().numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("B,F,k", [(256, 2048, 25), (64, 72, 25), (7, 16, 7), (300, 40, 200), (1000, 8, 3), (5, 6, 4)])
@pytest.mark.parametrize("largest", [True, False])
def test_topk_columns_bit_exact_vs_torch(B, F, k, largest):
    from sparse_vision_b200 import ops
    vals = torch.randn(B, F, generator=torch.Generator().manual_seed(B + F + k)).cuda()
    want_v, want_i = torch.topk(vals, k=k, dim=0, largest=largest)
    got_v, got_i, got_f = ops.topk_columns(vals, k, largest=largest)
    assert got_f is None
    assert torch.equal(got_v, want_v)
    assert torch.equal(got_i, want_i)          # no ties in continuous random data: indices are bit-exact


def test_topk_columns_ties_and_special_values():
    from sparse_vision_b200 import ops
    vals = torch.tensor([[0.0, 1.0, -0.0, 2.0], [0.0, 1.0, 0.0, float("inf")], [1.0, 1.0, -1.0, float("nan")],
                         [0.0, -3.0, 0.0, -float("inf")], [1.0, 1.0, 5.0, 2.0]]).cuda()
    for largest in (True, False):
        got_v, got_i, _ = ops.topk_columns(vals, 3, largest=largest)
        want_v, _ = torch.topk(vals, 3, dim=0, largest=largest)
        assert torch.equal(got_v.nan_to_num(nan=9e9), want_v.nan_to_num(nan=9e9))       # the VALUES are torch's
        assert torch.equal(torch.gather(vals, 0, got_i).nan_to_num(nan=9e9), got_v.nan_to_num(nan=9e9))
    got_v, got_i, _ = ops.topk_columns(vals, 3, largest=True)
    assert got_i[:, 0].tolist() == [2, 4, 0] and got_i[:, 1].tolist() == [0, 1, 2]      # ties: lower row first
    assert got_i[0, 3].item() == 2                                                     # NaN is the largest, like torch


def test_running_topk_merge_vs_reference_golden(golden_dir):
    """utils.get_top_k_samples on CUDA tensors (one merge kernel) against the REAL reference's outputs."""
    from sparse_vision_b200 import utils as U
    g = dict(np.load(os.path.join(golden_dir, "utils_small.npz")))
    T = lambda name: torch.from_numpy(g[name]).cuda()
    k, bs, F = 4, 5, 6
    for largest in (True, False):
        state = (torch.empty(0, F).cuda(), torch.empty(0, F, dtype=torch.long).cuda(), bs,
                 torch.empty(0, F, dtype=torch.long).cuda())
        for batch in (1, 2, 3):
            tag = f"topk_{int(largest)}_{batch}"
            state = U.get_top_k_samples(state, T(tag + "_v").clone(), T(tag + "_i").clone(), T(tag + "_f").clone(),
                                        batch, largest, k)
            assert torch.equal(state[0], T(tag + "_out_v")), tag
            assert torch.equal(state[1], T(tag + "_out_i")), tag        # top-k indices: bit-exact
            assert torch.equal(state[3], T(tag + "_out_f")), tag
            assert state[2] == bs


def test_histogram_matches_torch_histc():
    from sparse_vision_b200 import utils as U
    B, F, bins = 300, 50, 100
    vals = torch.randn(B, F, generator=torch.Generator().manual_seed(3)).cuda()
    units = [3, 17, 4, 49, 0]
    tops = torch.tensor([vals[:, u].max().item() for u in units])
    smalls = torch.tensor([vals[:, u].min().item() for u in units])
    tops[1], smalls[1] = 0.5, -0.5                 # values outside the range are ignored
    tops[2] = smalls[2] = 0.25                      # min == max: torch uses the data's own range
    info = {("l", "orig"): (torch.zeros(bins, len(units)), tops, smalls, units)}
    want = torch.zeros(bins, len(units))
    for rep in range(2):                           # accumulates over batches
        info = U.update_histogram(info, "l", "orig", vals, "cuda")
        for j, u in enumerate(units):
            want[:, j] += torch.histc(vals[:, u], bins=bins, min=smalls[j].item(), max=tops[j].item()).cpu()
    assert torch.equal(info[("l", "orig")][0].cpu(), want)
    info[("l", "sae")] = info[("l", "orig")]
    with pytest.raises(ValueError):                # the SAE's histogram is over the pre-ReLU output (output_2)
        U.update_histogram(info, "l", "sae", vals, "cuda")


def test_eval_batches_top_samples_vs_oracle():
    """Three eval batches through ModelPipeline.hook (SAE in inference mode): losses, dead-unit AND and the running
    top / small k samples of every unit against the CPU oracle.  Index parity: exact for every unit whose oracle values
    are separated by more than the bf16 error of the spatial means; near ties may swap, never anything else."""
    import sparse_vision_b200.models.sae_mlp as M
    from sparse_vision_b200.model_pipeline import ModelPipeline

    C, k_exp, B, hw, K = 64, 4, 12, 8, 5
    torch.manual_seed(0)
    base = nn.Sequential(collections.OrderedDict(conv=nn.Conv2d(3, C, 3, padding=1), act=nn.ReLU())).eval()
    sae = M.SaeMLP(C, k_exp)
    p = {key: v.detach().clone() for key, v in sae.state_dict().items()}
    xs = [torch.randn(B, 3, hw, hw, generator=torch.Generator().manual_seed(200 + i)) for i in range(3)]
    base, sae = base.cuda(), sae.cuda()
    with torch.no_grad():
        acts = [base(x.cuda()).cpu() for x in xs]
    pipe = ModelPipeline(base, sae, "sae_mlp", "act", "constrained_adam", 1e-3, 5.0, k_exp)
    pipe.record_top_samples, pipe.k = True, K
    pipe.register_hooks(train_sae=False)
    for x in xs:
        out = pipe.eval_batch(x.cuda())
        assert out.shape == (B, C, hw, hw)
    F = C * k_exp
    ref_means, ref_dead = [], None
    for a in acts:
        loss, rec, l1, nrmse, rmse, aux, enc, pre, dec = O.sae_inference_and_loss("sae_mlp", p, a, 5.0)
        ref_means.append(pre.mean(dim=(2, 3)))
        dead, _, _ = O.measure_inactive_units(enc, k_exp)
        ref_dead = dead if ref_dead is None else ref_dead & dead
    sc = pipe.batch_scalars()
    assert abs(sc["loss"] - float(loss)) <= 1e-2 * float(loss) and abs(sc["rec"] - float(rec)) <= 1e-2 * float(rec)
    assert torch.equal(pipe.eval_dead_neurons[("act", "sae")].cpu().bool(), ref_dead)
    allm = torch.cat(ref_means, 0)                                   # [3B, F]; dataset index = row (batch_size = B)
    err = 0.0
    for largest, state in ((True, pipe.top_k_samples[("act", "sae")]), (False, pipe.small_k_samples[("act", "sae")])):
        vals, idx, bs, files = state
        assert vals.shape == (K, F) and bs == B
        assert torch.equal(idx, files)                               # default filename indices = dataset positions
        want_v, want_i = torch.topk(allm, K + 1, dim=0, largest=largest)
        e = (vals.cpu() - want_v[:K]).abs().max().item()             # bf16 GEMM operands vs the fp32 oracle
        err = max(err, e)
        # the sample recorded at rank j must be the oracle's rank-j sample, or one whose oracle value is within the bf16
        # error of it (a near tie the two precisions may order differently)
        picked = torch.gather(allm, 0, idx.cpu())
        assert (picked - want_v[:K]).abs().max().item() <= 2 * e + 1e-6
        exact = (idx.cpu() == want_i[:K]).float().mean().item()
        assert exact >= 0.9, exact
    assert err <= 1e-2 * allm.abs().max().item()
