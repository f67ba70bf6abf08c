"""Pins oracle/sae_oracle.py against fixtures produced by the REAL reference (oracle/gen_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import sae_oracle as O


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _params(g, prefix, keys):
    return {k: torch.from_numpy(g[prefix + k]).clone() for k in keys}


SCALARS = ["loss", "rec", "l1", "nrmse", "rmse", "aux", "sparsity", "var_expl"]


def _run_train(g, kind, opt_name, keys):
    act, k, lam, lr, seed, plant = g["meta"]
    k = int(k)
    p = _params(g, "init.", keys)
    st = O.new_adam_state(p, keys)
    outs = []
    for i in range(g["x"].shape[0]):
        x = torch.from_numpy(g["x"][i])
        outs.append(O.train_step(kind, p, st, x, float(lam), opt_name, float(lr), k))
    return p, st, outs


@pytest.mark.parametrize("name,kind,opt,keys", [
    ("cfg1_mlp_adam.npz", "sae_mlp", "adam", O.SAE_MLP_KEYS),
    ("conv_mlp_cadam.npz", "sae_mlp", "constrained_adam", O.SAE_MLP_KEYS),
    ("conv_gated_cadam.npz", "gated_sae", "constrained_adam", O.GATED_KEYS),
])
def test_train_steps_match_reference(golden_dir, name, kind, opt, keys):
    g = _load(golden_dir, name)
    p, st, outs = _run_train(g, kind, opt, keys)
    for i, o in enumerate(outs):
        ref = g[f"step{i}.scalars"]
        got = np.array([o[s] for s in SCALARS])
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-7, err_msg=f"step {i} scalars {SCALARS}")
        assert np.array_equal(o["dead"].numpy(), g[f"step{i}.dead"]), f"dead mask step {i}"
        np.testing.assert_allclose(o["freq"].numpy(), g[f"step{i}.freq"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(outs[0]["enc"].numpy(), g["step0.enc"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(outs[0]["dec"].numpy(), g["step0.dec"], rtol=1e-5, atol=1e-6)
    if "step0.pre" in g:
        np.testing.assert_allclose(outs[0]["pre"].numpy(), g["step0.pre"], rtol=1e-5, atol=1e-6)
    for k in keys:
        np.testing.assert_allclose(outs[0]["grads"][k].numpy(), g["step0.grad." + k], rtol=1e-4, atol=1e-8,
                                   err_msg=f"grad {k}")
        np.testing.assert_allclose(p[k].numpy(), g["final." + k], rtol=1e-4, atol=2e-6, err_msg=f"final {k}")


def test_init_matches_reference_rng_stream(golden_dir):
    g = _load(golden_dir, "cfg1_mlp_adam.npz")
    torch.manual_seed(0)
    p = O.init_sae_mlp(16, 4)
    for k in O.SAE_MLP_KEYS:
        assert np.array_equal(p[k].numpy(), g["init." + k]), k
    g = _load(golden_dir, "conv_gated_cadam.npz")
    torch.manual_seed(0)
    p = O.init_gated_sae(32, 4)
    for k in ("W_gate", "decoder.weight"):
        assert np.array_equal(p[k].numpy(), g["init." + k]), k


def test_reset_encoder_weights_matches_reference(golden_dir):
    g = _load(golden_dir, "conv_mlp_cadam.npz")
    keys = O.SAE_MLP_KEYS
    p = _params(g, "final.", keys)
    st = {"step": 4, "m": {k: torch.from_numpy(g["pre_reset.m." + k]).clone() for k in keys},
          "v": {k: torch.from_numpy(g["pre_reset.v." + k]).clone() for k in keys}}
    dead = torch.from_numpy(g["step3.dead"])
    assert dead.sum() >= 6
    torch.manual_seed(77)
    n = O.reset_encoder_weights(p, st, dead)
    assert n == int(dead.sum())
    for k in keys:
        np.testing.assert_allclose(p[k].numpy(), g["reset." + k], rtol=1e-6, atol=1e-7, err_msg=k)
        np.testing.assert_allclose(st["m"][k].numpy(), g["reset.m." + k], rtol=0, atol=0, err_msg="m " + k)
        np.testing.assert_allclose(st["v"][k].numpy(), g["reset.v." + k], rtol=0, atol=0, err_msg="v " + k)


def test_ie_reductions_and_apply_sae(golden_dir):
    g = _load(golden_dir, "ie_small.npz")
    p = _params(g, "init.", O.SAE_MLP_KEYS)
    x, grad = torch.from_numpy(g["x"]), torch.from_numpy(g["g"])
    enc, dec, new_dec = O.apply_sae(p, x)
    np.testing.assert_allclose(enc.numpy(), g["enc"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dec.numpy(), g["dec"], rtol=1e-5, atol=1e-6)
    enc_avg, err_avg, x_avg = (torch.from_numpy(g[k]) for k in ("enc_avg", "err_avg", "x_avg"))
    ie_feat, ie_err, ie_neur = O.node_ie_layer(p, x, grad, enc_avg, err_avg, x_avg)
    np.testing.assert_allclose(ie_feat.numpy(), g["ie_feat"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ie_err.numpy(), g["ie_err"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ie_neur.numpy(), g["ie_neur"], rtol=1e-5, atol=1e-7)
    nodes = torch.from_numpy(g["nodes"])
    _, _, abl = O.apply_sae(p, x, nodes=nodes, ablation=enc_avg)
    np.testing.assert_allclose(abl.numpy(), g["ablated_dec"], rtol=1e-5, atol=1e-6)


def test_intervention_gradient_identity(golden_dir):
    """nnsight_intervention_check.py:194-195,212-213: enc.grad == layer_grad @ W_dec under stop-grad+pass-through."""
    g = _load(golden_dir, "ie_small.npz")
    p = {k: v.requires_grad_(True) for k, v in _params(g, "init.", O.SAE_MLP_KEYS).items()}
    x = torch.from_numpy(g["x"])
    head = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(24 * 4 * 6, 5))
    tgt = torch.tensor([1, 0, 3])
    enc, enc_grad, grad_orig = O.node_ie_layer_via_autograd(
        p, x, lambda t: torch.nn.functional.cross_entropy(head(torch.tanh(t)), tgt))
    want = O.to_tokens(grad_orig)[0] @ p["decoder.weight"].detach()
    np.testing.assert_allclose(enc_grad.numpy(), want.numpy(), rtol=1e-5, atol=1e-8)


def test_dead_neuron_schedule(golden_dir):
    g = _load(golden_dir, "schedule.npz")
    for n, upto in ((9912, 100000), (8, 70)):
        reinit = [i for i in range(1, upto + 1) if O.dead_neuron_action(i, n) == "reinit"]
        wait = [i for i in range(1, upto + 1) if O.dead_neuron_action(i, n) == "clear"]
        assert reinit == list(g[f"reinit_{n}"]) and wait == list(g[f"wait_{n}"])
    # the values the survey quotes from the reference script's output
    assert list(g["reinit_9912"])[:2] == [19825, 39649] and list(g["wait_9912"])[:2] == [9912, 29736]
