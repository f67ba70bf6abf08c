"""cfg4-style check of the hook-driven training loop (model_pipeline.py:363-432, :744-793): a frozen base model, the
SAE trained through the forward hook with the fused step, dead-unit masks AND-ed over steps and dead units
re-initialised on the reference's schedule -- against the CPU oracle run on the same activations and RNG stream."""
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


@pytest.mark.parametrize("hw,steps", [(16, 11), (7, 8)])
def test_hook_training_with_reinit_matches_oracle(tmp_path, monkeypatch, hw, steps):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import sparse_vision_b200.models.sae_mlp as M
    from sparse_vision_b200.model_pipeline import ModelPipeline

    C, k, B, n_dead, every = 64, 4, 6, 13, 3
    torch.manual_seed(0)
    base = nn.Sequential(collections.OrderedDict(conv=nn.Conv2d(3, C, 3, padding=1), act=nn.ReLU())).eval()
    sae = M.SaeMLP(C, k)
    with torch.no_grad():
        sae.encoder.bias[:n_dead] = -50.0          # planted dead units
    p = {key: v.detach().clone() for key, v in sae.state_dict().items()}
    assert list(p) == list(O.SAE_MLP_KEYS)
    base, sae = base.cuda(), sae.cuda()
    # the fixtures of the oracle come from the CPU generator: draw the re-initialisation matrices there too
    orig = M.draw_reinit
    monkeypatch.setattr(M, "draw_reinit", lambda w, b, d, dead, draw_device=None: orig(w, b, d, dead, "cpu"))
    pipe = ModelPipeline(base, sae, "sae_mlp", "act", "constrained_adam", 1e-3, 5.0, k, dead_neurons_steps=every,
                         reinit_index_dir=str(tmp_path))
    pipe.register_hooks(train_sae=True)
    xs = [torch.randn(B, 3, hw, hw, generator=torch.Generator().manual_seed(100 + i)) for i in range(steps)]
    pipe.remove_hooks()
    with torch.no_grad():                              # raw layer outputs for the oracle (no hook installed)
        acts = [base(x.cuda()).cpu().bfloat16().float() for x in xs]   # the fused step consumes bf16-rounded activations
    pipe.register_hooks(train_sae=True)

    torch.manual_seed(777)
    got = []
    for x in xs:
        out, action = pipe.train_batch(x.cuda())
        assert out.shape == (B, C, hw, hw) and out.dtype == torch.float32      # the hook returns dec in the layer's format
        got.append((pipe.batch_scalars(), pipe._last.dead.cpu().bool().clone(), action))
    assert any(a == "reinit" for _, _, a in got), "the schedule never fired in this test"
    assert len(os.listdir(tmp_path)) == sum(a == "reinit" for _, _, a in got)   # one index file per re-initialisation

    torch.manual_seed(777)
    st = O.new_adam_state(p, O.SAE_MLP_KEYS)
    acc = None
    for i, x in enumerate(xs):
        ref = O.train_step("sae_mlp", p, st, acts[i], 5.0, "constrained_adam", 1e-3, k)
        sc, dead, action = got[i]
        for key in ("loss", "rec", "l1", "var_expl"):
            assert abs(sc[key] - float(ref[key])) <= 2e-2 * max(abs(float(ref[key])), 1e-3), (i, key, sc[key], float(ref[key]))
        assert torch.equal(dead, ref["dead"]), f"dead-unit mask differs at step {i}"
        acc = ref["dead"].clone() if acc is None else acc & ref["dead"]
        want = O.dead_neuron_action(i + 1, every)
        assert action == want, (i, action, want)
        if want == "reinit":
            n = O.reset_encoder_weights(p, st, acc)
            assert n >= n_dead
            acc = None
        elif want == "clear":
            acc = None
    for key, q in zip(O.SAE_MLP_KEYS, sae.param_list()):
        d = (q.detach().cpu() - p[key]).abs()
        assert d.max().item() <= 2e-2 and d.mean().item() <= 5e-4, (key, d.max().item(), d.mean().item())
    # Adam moments of re-initialised units were reset on both sides
    m_gpu = pipe.sae_optimizer.state[sae.encoder.weight]["exp_avg"].cpu()
    assert np.allclose(m_gpu.numpy(), st["m"]["encoder.weight"].numpy(), atol=5e-3)


def test_checkpoint_round_trip_resumes_bit_exact(tmp_path):
    """model_pipeline.py:233-263,1266-1280: a checkpoint with the reference's keys, written mid-training and loaded into
    a fresh pipeline, continues exactly like the uninterrupted run (the fused step is deterministic)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import sparse_vision_b200.models.sae_mlp as M
    from sparse_vision_b200.model_pipeline import ModelPipeline

    C, k, B = 64, 4, 4

    def make():
        torch.manual_seed(0)
        base = nn.Sequential(collections.OrderedDict(conv=nn.Conv2d(3, C, 3, padding=1), act=nn.ReLU())).eval().cuda()
        sae = M.SaeMLP(C, k).cuda()
        pipe = ModelPipeline(base, sae, "sae_mlp", "act", "constrained_adam", 1e-3, 5.0, k)
        pipe.register_hooks(train_sae=True)
        return pipe, sae

    xs = [torch.randn(B, 3, 16, 16, generator=torch.Generator().manual_seed(7 + i)).cuda() for i in range(6)]
    pipe_a, sae_a = make()
    for x in xs:
        pipe_a.train_batch(x)
    pipe_b, sae_b = make()
    for x in xs[:3]:
        pipe_b.train_batch(x)
    path = os.path.join(str(tmp_path), "ckpt.pth")
    pipe_b.save_checkpoint(path, epoch=2)
    ck = torch.load(path, map_location="cpu")
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "training_step"}      # reference keys
    assert list(ck["model_state_dict"]) == ["encoder.weight", "encoder.bias", "decoder.weight", "decoder.bias"]
    pipe_c, sae_c = make()
    assert pipe_c.load_checkpoint(path) == 2 and pipe_c.train_batch_idx == 3
    for x in xs[3:]:
        pipe_c.train_batch(x)
    for a, c in zip(sae_a.param_list(), sae_c.param_list()):
        assert torch.equal(a, c), "resumed training diverged from the uninterrupted run"


def test_original_model_comparison_one_pass_equals_unhooked_copy():
    """model_pipeline.py:694-708 (KL divergence / same classification / loss difference against the unhooked original):
    carrying the original activation through the SAME forward pass (the hook hands [reconstruction; original] on, batch
    2B) must give what a second forward of an unhooked copy gives, and must not change the SAE training."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import copy
    import sparse_vision_b200.models.sae_mlp as M
    from sparse_vision_b200.model_pipeline import ModelPipeline

    C, k, B = 64, 4, 8

    def make(one_pass):
        torch.manual_seed(0)
        base = nn.Sequential(collections.OrderedDict(
            conv=nn.Conv2d(3, C, 3, padding=1), act=nn.ReLU(), c2=nn.Conv2d(C, 32, 3, padding=1), r2=nn.ReLU(),
            gap=nn.AdaptiveAvgPool2d(1), fl=nn.Flatten(), fc=nn.Linear(32, 10))).eval().cuda()
        sae = M.SaeMLP(C, k).cuda()
        pipe = ModelPipeline(base, sae, "sae_mlp", "act", "constrained_adam", 1e-3, 5.0, k,
                             model_copy=None if one_pass else copy.deepcopy(base), compare_in_one_pass=one_pass)
        pipe.register_hooks(train_sae=True)
        return pipe, sae

    xs = [torch.randn(B, 3, 16, 16, generator=torch.Generator().manual_seed(7 + i)).cuda() for i in range(3)]
    ys = [torch.randint(0, 10, (B,), generator=torch.Generator().manual_seed(70 + i)).cuda() for i in range(3)]
    (pa, sa), (pb, sb) = make(False), make(True)
    for x, y in zip(xs, ys):
        oa, _ = pa.train_batch(x, targets=y)
        ob, _ = pb.train_batch(x, targets=y)
        assert oa.shape == ob.shape == (B, 10)
        assert torch.allclose(oa, ob, rtol=1e-5, atol=1e-6)
        assert pa.batch_model_stats.shape == (3,)
        assert torch.allclose(pa.batch_model_stats, pb.batch_model_stats, rtol=1e-4, atol=1e-6), \
            (pa.batch_model_stats, pb.batch_model_stats)
        assert float(pa.batch_model_stats[0]) > 0          # the SAE does change the model's distribution
    for a, b in zip(sa.param_list(), sb.param_list()):
        assert torch.equal(a, b)


def test_cuda_graphed_training_batches_equal_eager(tmp_path, monkeypatch):
    """the whole training batch (frozen forward, fused SAE step in the hook, comparison with the
    original model) captured in a CUDA graph and replayed must reproduce the eager run bit for bit -- including the Adam
    bias corrections (step count on the device) and a dead-unit re-initialisation between replays."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import sparse_vision_b200.models.sae_mlp as M
    from sparse_vision_b200.model_pipeline import ModelPipeline

    C, k, B = 64, 4, 8

    def make(graph):
        torch.manual_seed(0)
        base = nn.Sequential(collections.OrderedDict(
            conv=nn.Conv2d(3, C, 3, padding=1), act=nn.ReLU(), c2=nn.Conv2d(C, 32, 3, padding=1), r2=nn.ReLU(),
            gap=nn.AdaptiveAvgPool2d(1), fl=nn.Flatten(), fc=nn.Linear(32, 10))).eval().cuda()
        sae = M.SaeMLP(C, k)
        with torch.no_grad():
            sae.encoder.bias[:9] = -50.0
        sae = sae.cuda()
        pipe = ModelPipeline(base, sae, "sae_mlp", "act", "constrained_adam", 1e-3, 5.0, k, dead_neurons_steps=3,
                             compare_in_one_pass=True, cuda_graph=graph)
        pipe.register_hooks(train_sae=True)
        return pipe, sae

    xs = [torch.randn(B, 3, 16, 16, generator=torch.Generator().manual_seed(7 + i)).cuda() for i in range(10)]
    ys = [torch.randint(0, 10, (B,), generator=torch.Generator().manual_seed(70 + i)).cuda() for i in range(10)]
    runs = []
    for graph in (False, True):
        pipe, sae = make(graph)
        torch.manual_seed(123)                       # the re-initialisation draws
        log = []
        for x, y in zip(xs, ys):
            out, action = pipe.train_batch(x, targets=y)
            log.append((pipe.batch_scalars(), pipe.batch_model_stats.clone(), out.clone(), action,
                        pipe._last.dead.clone()))
        runs.append((pipe, sae, log))
    (pe, se, le), (pg, sg, lg) = runs
    assert pg._graph is not None and pe._graph is None
    assert any(a == "reinit" for _, _, _, a, _ in lg)
    for i, ((s0, m0, o0, a0, d0), (s1, m1, o1, a1, d1)) in enumerate(zip(le, lg)):
        assert a0 == a1 and s0 == s1, (i, s0, s1)
        assert torch.equal(m0, m1) and torch.equal(o0, o1) and torch.equal(d0, d1), i
    for a, b in zip(se.param_list(), sg.param_list()):
        assert torch.equal(a, b)
    for p0, p1 in zip(se.param_list(), sg.param_list()):
        s0, s1 = pe.sae_optimizer.state[p0], pg.sae_optimizer.state[p1]
        assert int(s0["step"]) == int(s1["step"]) == 10 and torch.equal(s0["exp_avg"], s1["exp_avg"])
    assert int(pg._step_dev.item()) == 10
