"""The memory-bound layers of the frozen activation producer (utils.py:277-281,
model_pipeline.py:445-475): svb_maxpool_nhwc against torch.nn.functional.max_pool2d (bit-exact, every window GoogLeNet
uses plus ragged sizes), svb_bias_relu_scatter against add_ + relu_ + torch.cat (bit-exact), and the fused forward of
the whole network against torchvision's eager forward of the same frozen model."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _nhwc(t):
    return t.cuda().bfloat16().contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("B,C,H,W,k,s,p,ceil", [
    (3, 64, 112, 112, 3, 2, 0, True),     # maxpool1
    (2, 192, 56, 56, 3, 2, 0, True),      # maxpool2
    (2, 480, 28, 28, 3, 2, 0, True),      # maxpool3
    (2, 832, 14, 14, 2, 2, 0, True),      # maxpool4
    (2, 192, 28, 28, 3, 1, 1, True),      # inception branch4 pools
    (2, 528, 14, 14, 3, 1, 1, True),
    (3, 832, 7, 7, 3, 1, 1, True),
    (2, 8, 5, 9, 3, 2, 0, False),         # ragged: floor mode, odd sizes, one channel vector
    (1, 24, 6, 11, 3, 2, 1, True),
    (2, 16, 3, 3, 2, 2, 0, True),
    (1, 40, 1, 1, 3, 1, 1, False),
    (2, 72, 13, 2, 3, 1, 1, False),
])
def test_maxpool_nhwc_is_exact(B, C, H, W, k, s, p, ceil):
    from sparse_vision_b200 import ops
    x = _nhwc(torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(C + H)))
    want = F.max_pool2d(x, k, s, p, ceil_mode=ceil)
    got = ops.maxpool_nhwc(x, k, s, p, ceil)
    assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, want)
    # and against the CPU implementation in fp32 (max is exact in any precision)
    assert torch.equal(got.float().cpu(), F.max_pool2d(x.float().cpu(), k, s, p, ceil_mode=ceil))


def test_maxpool_nhwc_nan_and_inf():
    from sparse_vision_b200 import ops
    x = torch.randn(1, 8, 6, 6, generator=torch.Generator().manual_seed(3))
    x[0, 0, 2, 2] = float("nan")
    x[0, 1, 0, 0] = float("-inf")
    x[0, 2, 5, 5] = float("inf")
    x = _nhwc(x)
    want, got = F.max_pool2d(x, 3, 1, 1, ceil_mode=True), ops.maxpool_nhwc(x, 3, 1, 1, True)
    assert torch.equal(torch.isnan(got), torch.isnan(want)) and torch.isnan(got).sum() == 9
    assert torch.equal(torch.nan_to_num(got, nan=0.0), torch.nan_to_num(want, nan=0.0))


def test_maxpool_nhwc_rejects_what_it_cannot_do():
    from sparse_vision_b200 import ops
    from sparse_vision_b200._lib import SvbError
    with pytest.raises(ValueError):
        ops.maxpool_nhwc(torch.randn(1, 8, 4, 4).cuda(), 3, 2)                      # fp32
    with pytest.raises(ValueError):
        ops.maxpool_nhwc(torch.randn(1, 8, 4, 4).cuda().bfloat16(), 3, 2)           # NCHW-contiguous
    with pytest.raises(SvbError):
        ops.maxpool_nhwc(_nhwc(torch.randn(1, 12, 4, 4)), 3, 2)                     # C % 8
    with pytest.raises(SvbError):
        ops.maxpool_nhwc(_nhwc(torch.randn(1, 8, 9, 9)), 5, 1, 2)                   # window GoogLeNet does not have


@pytest.mark.parametrize("B,H,W,chans", [(4, 28, 28, (64, 96, 16)), (3, 14, 14, (112, 144, 32)), (2, 7, 7, (384,)),
                                         (5, 3, 5, (8, 8, 8, 8)), (1, 1, 1, (16, 24))])
def test_bias_relu_scatter_matches_add_relu_cat(B, H, W, chans):
    """One source convolution output split over its destinations (the merged 1x1 convolution of an inception block),
    each landing at a channel offset inside a wider tensor whose other channels must stay untouched."""
    from sparse_vision_b200 import ops
    g = torch.Generator().manual_seed(sum(chans))
    C = sum(chans)
    src = _nhwc(torch.randn(B, C, H, W, generator=g) * 3)
    bias = (torch.randn(C, generator=g)).cuda().bfloat16()
    want = src.clone().add_(bias.view(1, C, 1, 1)).relu_()
    dests, begin = [], 0
    for i, c in enumerate(chans):
        wide = _nhwc(torch.full((B, c + 16 * (i + 1), H, W), -7.0))
        dests.append((wide, 8 * (i + 1), c))
    ops.bias_relu_scatter(src, bias, dests)
    for (wide, off, c) in dests:
        assert torch.equal(wide[:, off:off + c], want[:, begin:begin + c])
        rest = torch.cat([wide[:, :off], wide[:, off + c:]], 1)
        assert bool((rest == -7.0).all())
        begin += c
    # in place, without relu
    y = src.clone(memory_format=torch.channels_last)
    ops.bias_relu_scatter(y, bias, [(y, 0, C)], relu=False)
    assert torch.equal(y, src.clone().add_(bias.view(1, C, 1, 1)))


def test_bias_relu_scatter_checks_its_arguments():
    from sparse_vision_b200 import ops
    from sparse_vision_b200._lib import SvbError
    src = _nhwc(torch.randn(2, 16, 4, 4))
    bias = torch.zeros(16).cuda().bfloat16()
    with pytest.raises(SvbError):
        ops.bias_relu_scatter(src, bias, [(src, 0, 8)])                     # does not cover C
    with pytest.raises(SvbError):
        ops.bias_relu_scatter(src, bias, [(src, 4, 16)])                    # offset not a multiple of 8 / outside the row
    with pytest.raises(ValueError):
        ops.bias_relu_scatter(src, bias.float(), [(src, 0, 16)])
    with pytest.raises(ValueError):
        ops.bias_relu_scatter(src, bias, [(_nhwc(torch.randn(2, 16, 5, 4)), 0, 16)])


def test_fused_forward_matches_the_eager_model():
    """fuse_forward changes which kernels run, not the function.  The yardstick is the SAME frozen model in fp32 (bf16
    weights widened, TF32 off): at every hooked layer and at the logits the fused bf16 forward must be as close to it
    as torchvision's eager bf16 forward is (both round every activation to bf16; the fused stem rounds once where
    conv + add_ round twice, the merged 1x1 convolutions may sum in another order).  The state_dict and module names
    are unchanged and the hooks fire on the same modules."""
    import copy
    from sparse_vision_b200 import _lib as L
    from sparse_vision_b200.producer import GOOGLENET_LAYERS, fuse_forward, synthetic_googlenet, to_producer_format
    dev = torch.device("cuda:0")
    eager = to_producer_format(synthetic_googlenet(seed=0), dev, torch.bfloat16, channels_last=True, fold_bn=True)
    fused = fuse_forward(copy.deepcopy(eager))
    exact = copy.deepcopy(eager).float()
    assert list(fused.state_dict().keys()) == list(eager.state_dict().keys())
    assert [n for n, _ in fused.named_modules()] == [n for n, _ in eager.named_modules()]
    x = _nhwc(torch.randn(6, 3, 224, 224, generator=torch.Generator().manual_seed(5)))
    seen = {}

    def grab(tag):
        def hook(mod, inp, out):
            seen.setdefault(tag, {})[mod._svb_name] = out
        return hook
    for tag, model in (("eager", eager), ("fused", fused), ("exact", exact)):
        for name, (mod_name, _, _) in GOOGLENET_LAYERS.items():
            m = dict(model.named_modules())[mod_name]
            m._svb_name = name
            m.register_forward_hook(grab(tag))
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            ref = exact(x.float())
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    launches0 = L.load().svb_launch_count()
    with torch.no_grad():
        want, got = eager(x), fused(x)
    assert L.load().svb_launch_count() - launches0 >= 3 + 4 + 9 * 5
    report = []
    for name, (_, C, hw) in GOOGLENET_LAYERS.items():
        r, a, b = seen["exact"][name], seen["eager"][name].float(), seen["fused"][name].float()
        assert tuple(b.shape[1:2]) == (C,) and b.shape[2] * b.shape[3] == hw
        assert seen["fused"][name].is_contiguous(memory_format=torch.channels_last)
        e_eager, e_fused = (a - r).abs().mean().item(), (b - r).abs().mean().item()
        report.append((name, e_eager / r.abs().mean().item(), e_fused / r.abs().mean().item()))
        assert e_fused <= 1.25 * e_eager + 1e-6, report
        assert (b - r).abs().max().item() <= 1.5 * (a - r).abs().max().item() + 1e-6, report
    e_eager, e_fused = (want.float() - ref).abs().mean().item(), (got.float() - ref).abs().mean().item()
    assert e_fused <= 1.25 * e_eager + 1e-6, (report, e_eager, e_fused)
    print("relative mean error vs fp32 (layer, eager bf16, fused bf16):", report)
    # conv3's bias + relu pass runs BEHIND maxpool2 on the fused model (relu(max(y) + b) == max(relu(y + b))): the pooled
    # tensor is bit-identical to conv -> add_ -> relu_ -> max_pool2d
    h = _nhwc(torch.randn(3, 64, 56, 56, generator=torch.Generator().manual_seed(8)))
    with torch.no_grad():
        raw = fused.conv3(h)
        assert getattr(raw, "_svb_pending_bias", None) is not None
        assert torch.equal(fused.maxpool2(raw), eager.maxpool2(eager.conv3(h)))
    # a tensor that asks for gradients (what the IE passes send through the layers behind a hooked one) takes
    # torchvision's own forward
    xg = x[:2].clone().requires_grad_(True)
    fused(xg).float().sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad.float()).all()


@pytest.mark.parametrize("B", [1, 5])
def test_conv1_stem_matches_conv2d(B):
    """GoogLeNet's 7x7 / stride-2 stem: against F.conv2d in fp32 on the same bf16-rounded operands (the kernel
    accumulates in fp32 and rounds once), every border row / column included; bit-exact on integer-valued data."""
    from sparse_vision_b200 import ops
    g = torch.Generator().manual_seed(B)
    w = (torch.randn(64, 3, 7, 7, generator=g) * 0.1).cuda().bfloat16()
    bias = torch.randn(64, generator=g).cuda().bfloat16()
    x = _nhwc(torch.randn(B, 3, 224, 224, generator=g))
    for wt in (w, w.contiguous(memory_format=torch.channels_last)):
        got = ops.conv1_stem(x, ops.conv1_pack_weights(wt), bias)
        want = F.relu(F.conv2d(x.float(), w.float(), bias.float(), stride=2, padding=3))
        assert got.shape == (B, 64, 112, 112) and got.is_contiguous(memory_format=torch.channels_last)
        err = (got.float() - want).abs()
        assert err.max().item() <= 2e-2 * want.abs().max().item(), err.max().item()       # bf16 output rounding + TF32-free fp32 reference
        assert err.mean().item() <= 2e-3 * want.abs().mean().item()
    # small integers: every product and partial sum is exact in fp32 and the result is exact in bf16 where |y| <= 256
    xi = _nhwc(torch.randint(-2, 3, (B, 3, 224, 224), generator=g).float())
    wi = torch.randint(-1, 2, (64, 3, 7, 7), generator=g).float().cuda().bfloat16()
    bi = torch.randint(-3, 4, (64,), generator=g).float().cuda().bfloat16()
    got = ops.conv1_stem(xi, ops.conv1_pack_weights(wi), bi, relu=False).float()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        want = F.conv2d(xi.float(), wi.float(), bi.float(), stride=2, padding=3)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert want.abs().max().item() <= 256
    assert torch.equal(got, want)


def test_conv1_stem_rejects_other_shapes():
    from sparse_vision_b200 import ops
    w = torch.zeros(64, 3, 7, 7).cuda().bfloat16()
    with pytest.raises(ValueError):
        ops.conv1_stem(_nhwc(torch.zeros(1, 3, 96, 96)), ops.conv1_pack_weights(w), torch.zeros(64).cuda().bfloat16())
    with pytest.raises(ValueError):
        ops.conv1_pack_weights(torch.zeros(64, 3, 3, 3).cuda().bfloat16())


def test_pipeline_on_the_fused_producer_trains_the_same_sae():
    """ModelPipeline.train_batch (hook on inception3a, same-pass comparison with the original model) on the fused
    producer: eager launches and the CUDA-graphed batch are bit-identical to each other, and the SAE statistics follow
    the run on torchvision's eager forward (the activations at mixed3a differ by bf16 rounding only)."""
    import copy
    from sparse_vision_b200.model_pipeline import ModelPipeline
    from sparse_vision_b200.models.sae_mlp import SaeMLP
    from sparse_vision_b200.producer import fuse_forward, synthetic_googlenet, to_producer_format
    dev = torch.device("cuda:0")
    eager = to_producer_format(synthetic_googlenet(seed=0), dev, torch.bfloat16, channels_last=True, fold_bn=True)
    B = 8
    xs = [_nhwc(torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(50 + i))) for i in range(6)]
    ys = [torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(90 + i)).cuda() for i in range(6)]

    def run(base, graph):
        torch.manual_seed(0)
        sae = SaeMLP(256, 8).to(dev)
        pipe = ModelPipeline(base, sae, "sae_mlp", "inception3a", "constrained_adam", 1e-3, 5.0, 8,
                             compare_in_one_pass=True, cuda_graph=graph)
        pipe.register_hooks(train_sae=True)
        log = []
        for x, y in zip(xs, ys):
            out, _ = pipe.train_batch(x, targets=y)
            log.append((pipe.batch_scalars(), pipe.batch_model_stats.clone(), out.clone()))
        pipe.remove_hooks()
        return pipe, sae, log

    _, s_ref, l_ref = run(copy.deepcopy(eager), False)
    _, s_fe, l_fe = run(fuse_forward(copy.deepcopy(eager)), False)
    p_fg, s_fg, l_fg = run(fuse_forward(copy.deepcopy(eager)), True)
    assert p_fg._graph is not None and p_fg.graph_svb_launches >= 60   # 16 of the SAE step + the producer's own kernels
    for (sc_e, ms_e, out_e), (sc_g, ms_g, out_g) in zip(l_fe, l_fg):
        assert sc_e == sc_g and torch.equal(ms_e, ms_g) and torch.equal(out_e, out_g)
    for a, b in zip(s_fe.param_list(), s_fg.param_list()):
        assert torch.equal(a, b)
    for step, ((sc_r, ms_r, _), (sc_f, ms_f, _)) in enumerate(zip(l_ref, l_fe)):
        for key in ("loss", "rec", "l1", "var_expl"):
            assert abs(sc_f[key] - sc_r[key]) <= 3e-2 * max(abs(sc_r[key]), 1e-3), (step, key, sc_f[key], sc_r[key])


@pytest.mark.parametrize("B,C,H,W,k,s,p", [(2, 480, 28, 28, 3, 2, 0), (2, 256, 28, 28, 3, 1, 1), (2, 832, 14, 14, 2, 2, 0),
                                           (3, 832, 7, 7, 3, 1, 1), (1, 16, 5, 6, 3, 2, 0), (2, 8, 9, 4, 3, 1, 1)])
def test_maxpool_nhwc_autograd_matches_torch(B, C, H, W, k, s, p):
    """Forward with indices + gather backward against torch: the forward is exact; the gradient lands on the same
    positions as max_pool2d's (ties go to the first maximum in window order -- post-ReLU data with many exact zeros
    and a coarse grid of positive values makes ties the common case here) and equals the fp32 backward on the same
    values up to the one bf16 rounding of the sum."""
    from sparse_vision_b200 import ops
    g = torch.Generator().manual_seed(B * C + H)
    x = _nhwc(torch.relu(torch.randn(B, C, H, W, generator=g)).mul(4).round().div(4))      # ties: zeros and a 0.25 grid
    xa = x.clone().requires_grad_(True)
    ya = ops.maxpool_nhwc_autograd(xa, k, s, p, True)
    xt = x.float().requires_grad_(True)
    yt = F.max_pool2d(xt, k, s, p, ceil_mode=True)
    assert torch.equal(ya.float(), yt) and ya.is_contiguous(memory_format=torch.channels_last)
    go = _nhwc(torch.randn(yt.shape, generator=g))
    ya.backward(go)
    yt.backward(go.float())
    want = xt.grad
    assert xa.grad.dtype == torch.bfloat16 and xa.grad.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(xa.grad.float() != 0, want.bfloat16().float() != 0)
    assert torch.equal(xa.grad, want.bfloat16())
    # and against ATen's bf16 backward of the same pool (it accumulates overlapping windows in bf16: a few ulps apart)
    xb = x.clone().requires_grad_(True)
    F.max_pool2d(xb, k, s, p, ceil_mode=True).backward(go)
    assert (xa.grad.float() - xb.grad.float()).abs().max().item() <= 0.05 * go.float().abs().max().item() + 1e-6


def test_node_ie_on_the_attribution_format_model_eager_graph_and_nchw():
    """IE.compute_node_ie on producer.to_attribution_format (bf16 channels_last, fused forward-only head, libsvb's
    differentiable max-pool behind the leaf): the CUDA-graphed batches reproduce the eager ones bit for bit (also the
    second sweep, which only replays), and both follow the same frozen model run in NCHW on torchvision's forward and
    ATen's pools (same function, other kernels: bf16 rounding apart)."""
    import copy
    from sparse_vision_b200.compute_ie import IE
    from sparse_vision_b200.models.sae_mlp import SaeMLP
    from sparse_vision_b200.producer import (GOOGLENET_LAYERS, hooked_layers, synthetic_googlenet, to_attribution_format,
                                             to_producer_format)
    dev = torch.device("cuda:0")
    layers = {"mixed3a": 8, "mixed4c": 4, "mixed5b": 4}
    raw = synthetic_googlenet(seed=0)
    fast = to_attribution_format(copy.deepcopy(raw), dev)
    plain = to_producer_format(copy.deepcopy(raw), dev, torch.bfloat16, channels_last=False, fold_bn=True)
    saes = {}
    for j, (n, k) in enumerate(layers.items()):
        torch.manual_seed(5 + j)
        saes[n] = SaeMLP(GOOGLENET_LAYERS[n][1], k).to(dev)
    g = torch.Generator().manual_seed(2)
    batches = [(torch.randn(6, 3, 224, 224, generator=g).cuda().bfloat16(), torch.randint(0, 1000, (6,), generator=g).cuda())
               for _ in range(3)]
    batches_cl = [(x.contiguous(memory_format=torch.channels_last), y) for x, y in batches]

    def run(model, data, graph):
        ie = IE(model, hooked_layers(model, list(layers)), saes, dict(layers), device=dev, cuda_graph=graph)
        avg = ie.compute_average([b[0] for b in data])
        first = ie.compute_node_ie(data, avg)
        ie.last_avg = avg
        return first, ie.compute_node_ie(data, avg), ie

    (f_e, e_e, n_e), _, ie_e = run(fast, batches_cl, False)
    (f_g, e_g, n_g), (f_g2, e_g2, n_g2), ie_g = run(fast, batches_cl, True)
    assert len(ie_g._graphs) == 2                      # one capture for compute_average's batches, one for compute_node_ie's
    for key in ("encoder_output_average", "sae_error_average", "original_layer_output_average", "dead_units"):
        for name in layers:
            assert torch.equal(ie_e.last_avg[key][name], ie_g.last_avg[key][name]), (key, name)
    for name in layers:
        assert abs(ie_e.last_avg["sparsity"][name] - ie_g.last_avg["sparsity"][name]) <= 1e-9
    for name in layers:
        for a, b, c in ((f_e, f_g, f_g2), (n_e, n_g, n_g2)):
            assert torch.equal(a[name], b[name]) and torch.equal(a[name], c[name]), name
        assert torch.equal(e_e[name], e_g[name]) and torch.equal(e_e[name], e_g2[name])
    (f_p, e_p, n_p), _, ie_p = run(plain, batches, False)
    # edge IE and faithfulness cut the network into segments at the hooked layers (leaves): the fused blocks' hand-written
    # backward is then called once per downstream node with retain_graph -- same numbers as on the plain model up to the
    # bf16 rounding both carry
    feats = {"mixed4c": [3, 40, 77], "mixed5b": [5, 9]}
    edge_f = ie_e.compute_edge_ie(batches_cl[:1], ie_e.last_avg, ["mixed4c", "mixed5b"], feats)
    edge_p = ie_p.compute_edge_ie(batches[:1], ie_p.last_avg, ["mixed4c", "mixed5b"], feats)
    for name in edge_p:
        assert edge_f[name].shape == edge_p[name].shape and torch.isfinite(edge_f[name]).all()
        assert (edge_f[name] - edge_p[name]).abs().max().item() <= 0.25 * edge_p[name].abs().max().item() + 1e-9, name
    for name in layers:
        scale = f_p[name].abs().max().item()
        assert (f_e[name] - f_p[name]).abs().max().item() <= 0.15 * scale, name    # a random-weight net amplifies bf16 rounding
        assert (f_e[name] - f_p[name]).abs().mean().item() <= 0.05 * f_p[name].abs().mean().item(), name


def test_relu_grad_gather_matches_threshold_backward():
    from sparse_vision_b200 import ops
    g = torch.Generator().manual_seed(4)
    B, H, W = 3, 14, 14
    out = _nhwc(torch.relu(torch.randn(B, 512, H, W, generator=g)))          # a block output (post-ReLU: exact zeros)
    t3 = _nhwc(torch.relu(torch.randn(B, 128, H, W, generator=g)))
    go = _nhwc(torch.randn(B, 512, H, W, generator=g))
    g3 = _nhwc(torch.randn(B, 128, H, W, generator=g))
    got = ops.relu_grad_gather([(go, 0, out, 0, 128), (g3, 0, t3, 0, 128), (go, 448, out, 448, 64)], out)
    want = torch.cat([go[:, :128] * (out[:, :128] > 0), g3 * (t3 > 0), go[:, 448:] * (out[:, 448:] > 0)], 1)
    assert got.shape == (B, 320, H, W) and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, want)


@pytest.mark.parametrize("name,shape", [("inception3b", (4, 256, 28, 28)), ("inception4d", (3, 512, 14, 14)),
                                        ("inception5b", (5, 832, 7, 7))])
def test_fused_inception_backward_matches_autograd(name, shape):
    """The hand-written backward of the fused inception block (IE passes) against torch autograd through torchvision's
    forward of the same frozen block: output exact to bf16 rounding, input gradient as close to the fp32 gradient as
    torch's bf16 autograd is."""
    import copy
    from sparse_vision_b200.producer import fuse_forward, synthetic_googlenet, to_producer_format
    dev = torch.device("cuda:0")
    eager = to_producer_format(synthetic_googlenet(seed=0), dev, torch.bfloat16, channels_last=True, fold_bn=True)
    blk_e = getattr(eager, name)
    blk_f = getattr(fuse_forward(copy.deepcopy(eager)), name)
    blk_x = copy.deepcopy(blk_e).float()
    g = torch.Generator().manual_seed(shape[1])
    x = _nhwc(torch.relu(torch.randn(*shape, generator=g)))
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        res = {}
        for tag, blk, xin in (("eager", blk_e, x.clone()), ("fused", blk_f, x.clone()), ("exact", blk_x, x.float())):
            xin.requires_grad_(True)
            y = blk(xin)
            go = torch.randn(y.shape, generator=torch.Generator().manual_seed(1)).cuda().to(y.dtype)
            go = go.contiguous(memory_format=torch.channels_last)
            y.backward(go)
            res[tag] = (y.detach().float(), xin.grad.float())
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert type(blk_f).__name__ == "FusedInception"
    (y_e, g_e), (y_f, g_f), (y_x, g_x) = res["eager"], res["fused"], res["exact"]
    assert (y_f - y_x).abs().mean().item() <= 1.25 * (y_e - y_x).abs().mean().item() + 1e-6
    e_eager, e_fused = (g_e - g_x).abs().mean().item(), (g_f - g_x).abs().mean().item()
    assert e_fused <= 1.25 * e_eager + 1e-6, (e_eager, e_fused, g_x.abs().mean().item())
    assert (g_f - g_x).abs().max().item() <= 1.5 * (g_e - g_x).abs().max().item() + 1e-6
