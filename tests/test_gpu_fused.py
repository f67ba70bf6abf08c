"""The fused encoder backward (dE GEMM -> ReLU mask -> dW_enc GEMM in one kernel, single-CTA and SM-pair variants), the
SM-pair streaming GEMMs and the fused node-IE layer against the plain kernels and against the oracle.

svb_set_tuning (include/svb.h) selects the kernels at run time, so one process can run the same seeded step through every
combination.  All variants round dPre' to bf16 in the same place, so their gradients differ only by fp32 summation order
(and by the column sums, which the fused kernel takes from the un-rounded values): they are held to 2e-3 of each other in
Frobenius norm, far inside the 2e-2 the bf16 path is allowed against the fp32 reference (model_pipeline.py:385)."""
import contextlib

import numpy as np
import pytest
import torch

from oracle import sae_oracle as O

pytestmark = pytest.mark.gpu

FUSED_BWD, FBW_2CTA, GEMM_PAIRS, ENC_2CTA, FBW_PREFETCH, FUSED_IE = range(6)


@contextlib.contextmanager
def tuning(**kw):
    from sparse_vision_b200 import _lib as L
    lib = L.load()
    keys = {"fused_bwd": FUSED_BWD, "fbw_2cta": FBW_2CTA, "pairs": GEMM_PAIRS, "enc_2cta": ENC_2CTA, "prefetch": FBW_PREFETCH,
            "fused_ie": FUSED_IE}
    old = {k: lib.svb_get_tuning(keys[k]) for k in kw}
    try:
        for k, v in kw.items():
            L.check(lib.svb_set_tuning(keys[k], int(v)), "svb_set_tuning")
        yield
    finally:
        for k, v in old.items():
            lib.svb_set_tuning(keys[k], v)


def _setup(B, C, H, W, k, channels_last):
    torch.manual_seed(0)
    p = O.init_sae_mlp(C, k)
    F = C * k
    dead_idx = torch.randperm(F, generator=torch.Generator().manual_seed(1))[:max(F // 20, 1)]
    p["encoder.bias"][dead_idx] = -50.0
    p["decoder.bias"].normal_(0, 0.05, generator=torch.Generator().manual_seed(2))
    x = torch.relu(torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(1234))).bfloat16().float()
    xg = x.cuda().bfloat16()
    if channels_last:
        xg = xg.contiguous(memory_format=torch.channels_last)
    return p, x, xg


def _flat_grads(xg, p, lam):
    from sparse_vision_b200 import ops
    params = [p[key].clone().cuda() for key in O.SAE_MLP_KEYS]
    ss = ops.SplitStep("sae_mlp", xg, params, lam)
    addr, n_sum, n_max = ss.grads()
    flat = ops.wrap_device_buffer(addr, n_sum + n_max, xg.device).clone().cpu().numpy()
    flags = ss.lib.svb_last_step_flags(ss.h)
    return flat, flags


def _sections(flat, C, F):
    FC = F * C
    return {"encoder.weight": flat[:FC], "encoder.bias": flat[FC:FC + F], "decoder.weight": flat[FC + F:2 * FC + F],
            "decoder.bias": flat[2 * FC + F:2 * FC + F + C], "rest": flat[2 * FC + F + C:]}


def _fro(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("B,C,H,W,k,channels_last", [
    (6, 256, 14, 14, 4, True),    # SM-pair kernel, several token blocks per CTA (F = 1024), zero-copy token input
    (6, 256, 14, 14, 4, False),   # ... NCHW input: slab-major X / DIFF workspaces
    (5, 128, 9, 9, 8, False),     # SM pairs with a token tail (405 tokens)
    (4, 64, 12, 12, 4, True),     # C = 64: single-CTA kernel (C % 128 != 0)
    (3, 192, 10, 10, 4, False),   # C = 192: single-CTA kernel, three k-blocks
    (2, 256, 28, 28, 8, True),    # the cfg2 layer shape at a small batch
    (1, 128, 5, 5, 3, False),     # 25 tokens (less than one block, one slot), F = 384: the second CTA of the last pair is half empty
    (1, 256, 6, 6, 1, True),      # F = 256 = C (expansion 1): one pair tile, 36 tokens
    (7, 64, 3, 3, 12, False),     # 63 tokens, single-CTA kernel with six feature tiles
])
def test_fused_backward_and_pair_gemms_match_the_plain_kernels(B, C, H, W, k, channels_last):
    p, x, xg = _setup(B, C, H, W, k, channels_last)
    F, lam = C * k, 5.0
    with tuning(fused_bwd=0, pairs=0):
        base, fl = _flat_grads(xg, p, lam)
        assert fl & 1 == 0
    variants = {
        "fused (default)": dict(fused_bwd=1, fbw_2cta=1, pairs=1),
        "fused single-CTA": dict(fused_bwd=1, fbw_2cta=0, pairs=0),
        "fused with L2 prefetch": dict(fused_bwd=1, fbw_2cta=1, pairs=1, prefetch=2),
        "un-fused on SM pairs": dict(fused_bwd=0, pairs=1),
    }
    sb = _sections(base, C, F)
    for name, kw in variants.items():
        with tuning(**kw):
            got, fl = _flat_grads(xg, p, lam)
        assert bool(fl & 1) == bool(kw.get("fused_bwd")), name
        sg = _sections(got, C, F)
        for key in ("encoder.weight", "encoder.bias", "decoder.weight", "decoder.bias"):
            assert _fro(sg[key], sb[key]) <= 2e-3, f"{name}: {key} differs by {_fro(sg[key], sb[key])}"
        # loss sums, per-channel statistics and the activity counts do not depend on the backward kernels at all
        assert np.allclose(sg["rest"], sb["rest"], rtol=1e-5, atol=1e-6), name


@pytest.mark.parametrize("B,C,H,W,k", [(6, 256, 14, 14, 4), (4, 64, 12, 12, 4)])
def test_fused_backward_gradients_vs_oracle_autograd(B, C, H, W, k):
    """The flat gradient buffer of the fused path against torch autograd through the oracle's forward + loss
    (sae_mlp.py:49-52, sparse_loss.py:35,41): 2e-2 in Frobenius norm, the tolerance of the golden-gradient test."""
    p, x, xg = _setup(B, C, H, W, k, True)
    F, lam = C * k, 5.0
    flat, fl = _flat_grads(xg, p, lam)
    assert fl & 1
    leaves = {key: p[key].clone().requires_grad_(True) for key in O.SAE_MLP_KEYS}
    O.sae_inference_and_loss("sae_mlp", leaves, x, lam)[0].backward()
    sg = _sections(flat, C, F)
    for key in O.SAE_MLP_KEYS:
        want = leaves[key].grad.numpy().reshape(-1)
        assert _fro(sg[key], want) <= 2e-2, f"{key}: {_fro(sg[key], want)}"


@pytest.mark.parametrize("B,C,H,W,k", [
    (5, 256, 28, 28, 8),     # the cfg5 / mixed3a layer shape: F = 2048, 16-byte loads of the running average
    (9, 128, 7, 7, 4),       # 7x7 maps: HW % 4 != 0 (scalar loads of the average), 441 tokens (token tail)
    (3, 256, 14, 14, 3),     # F = 768: three pair tiles, several slots
])
def test_fused_node_ie_layer_vs_oracle_and_unfused(B, C, H, W, k):
    """svb_node_ie_layer with a and G = g W_dec kept in TMEM (fused_ie_sm100.cuh) against the oracle
    (compute_ie.py:420-453, utils.py:2574-2637) and against the GEMM + reduction path; identical top-k feature sets."""
    from sparse_vision_b200 import ops
    torch.manual_seed(0)
    p = O.init_sae_mlp(C, k)
    F = C * k
    p["decoder.bias"].normal_(0, 0.05, generator=torch.Generator().manual_seed(2))
    gen = torch.Generator().manual_seed(77)
    x = torch.relu(torch.randn(B, C, H, W, generator=gen)).bfloat16().float()
    g = (torch.randn(B, C, H, W, generator=gen) * 0.1).bfloat16().float()
    # plant a margin between the leading features so that the top-k sets do not depend on bf16 rounding
    lead = torch.randperm(F, generator=gen)[:5]
    for r, f in enumerate(lead):
        p["encoder.weight"][f] *= 3.0 + 0.6 * r
    enc_avg = torch.rand(F, H, W, generator=gen)
    err_avg = torch.randn(C, H, W, generator=gen) * 0.05
    x_avg = x.mean(0)
    params = [p[key].clone().cuda() for key in O.SAE_MLP_KEYS]
    args = (x.cuda().bfloat16().contiguous(memory_format=torch.channels_last),
            g.cuda().bfloat16().contiguous(memory_format=torch.channels_last), params, enc_avg.cuda(), err_avg.cuda(), x_avg.cuda())
    with tuning(fused_ie=1):
        feat, err, neur = [t.float().cpu() for t in ops.node_ie_layer(*args)]
    with tuning(fused_ie=0):
        feat0, err0, neur0 = [t.float().cpu() for t in ops.node_ie_layer(*args)]
    rf, re, rn = O.node_ie_layer(p, x, g, enc_avg, err_avg, x_avg)
    assert (feat - rf).abs().max() <= 2e-2 * rf.abs().max() and (feat0 - rf).abs().max() <= 2e-2 * rf.abs().max()
    assert (feat - feat0).abs().max() <= 2e-2 * rf.abs().max()
    assert abs(float(err) - float(re)) <= 1e-2 * abs(float(re)) and abs(float(err0) - float(re)) <= 1e-2 * abs(float(re))
    assert (neur - rn).abs().max() <= 1e-2 * rn.abs().max() and torch.equal(neur, neur0)
    top = set(torch.topk(rf, 5).indices.tolist())
    assert set(torch.topk(feat, 5).indices.tolist()) == top == set(torch.topk(feat0, 5).indices.tolist())


def test_fused_backward_on_plain_token_matrices():
    """cfg1-style 2-D activations [T, C] (no image structure, hw = 1): the fused backward reads them in place."""
    from sparse_vision_b200 import ops
    T, C, k, lam = 300, 256, 4, 5.0
    torch.manual_seed(0)
    p = O.init_sae_mlp(C, k)
    x = torch.relu(torch.randn(T, C, generator=torch.Generator().manual_seed(5))).bfloat16().float()
    xg = x.cuda().bfloat16()
    with tuning(fused_bwd=0, pairs=0):
        base, fl0 = _flat_grads(xg, p, lam)
    got, fl1 = _flat_grads(xg, p, lam)
    assert fl0 & 1 == 0 and fl1 & 1 == 1
    sb, sg = _sections(base, C, C * k), _sections(got, C, C * k)
    for key in ("encoder.weight", "encoder.bias", "decoder.weight", "decoder.bias"):
        assert _fro(sg[key], sb[key]) <= 2e-3, key
    leaves = {key: p[key].clone().requires_grad_(True) for key in O.SAE_MLP_KEYS}
    O.sae_inference_and_loss("sae_mlp", leaves, x, lam)[0].backward()
    for key in O.SAE_MLP_KEYS:
        assert _fro(sg[key], leaves[key].grad.numpy().reshape(-1)) <= 2e-2, key
