"""CPU oracle: a plain-PyTorch (fp32, CPU) restatement of the reference's SAE training-and-attribution path.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product package `sparse_vision_b200`; the only
callers are tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

Parity pinning: the reference (jasper3100/sparse-vision) ships NO golden vectors or tests for this path
(SURVEY.md §4, §8c).  This restatement is therefore pinned against outputs of the reference ITSELF, produced in the
build container by oracle/gen_golden.py (which imports the real reference modules from /root/reference) and frozen
under tests/golden/*.npz; tests/test_oracle_golden.py checks every function here against those fixtures.  The
nnsight-driven IE driver (compute_ie.py) cannot run anywhere without nnsight + pretrained GoogLeNet; `node_ie_pass`
below restates it with plain torch hooks and is "parity unpinned" beyond the gradient identities the reference's own
scripts assert (supplementary_files_2/nnsight_intervention_check.py:194-213).

All arithmetic lives in PyTorch (unpinned version in the reference; the container's torch 2.11 is the oracle
runtime).  Parameters travel as dicts keyed like the reference's state_dict.
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------- initialisation
def _kaiming_uniform(rows, cols):
    # torch.nn.init.kaiming_uniform_(empty(rows, cols)) with a=0: bound = sqrt(6 / fan_in)   (sae_mlp.py:29,35)
    return torch.nn.init.kaiming_uniform_(torch.empty(rows, cols))


def init_sae_mlp(act_size, expansion_factor):
    """models/sae_mlp.py:20-40.  Consumes the global RNG in the reference's order (the nn.Linear constructors at
    :26 and :32 draw too) so that the same torch.manual_seed gives the same parameters."""
    hidden = int(act_size * expansion_factor)
    torch.nn.Linear(act_size, hidden)          # :26 (draws weight + bias, then discarded by :29-30)
    w_enc = _kaiming_uniform(hidden, act_size)  # :29
    torch.nn.Linear(hidden, act_size)          # :32
    w_dec = _kaiming_uniform(act_size, hidden)  # :35
    w_dec = w_dec / w_dec.norm(dim=0, keepdim=True)  # :39 unit-norm columns
    return {
        "encoder.weight": w_enc, "encoder.bias": torch.zeros(hidden),
        "decoder.weight": w_dec, "decoder.bias": torch.zeros(act_size),
    }


def init_gated_sae(act_size, expansion_factor):
    """models/gated_sae.py:4-26."""
    hidden = int(act_size * expansion_factor)
    w_gate = _kaiming_uniform(hidden, act_size)  # :11
    torch.nn.Linear(hidden, act_size)            # :18
    w_dec = _kaiming_uniform(act_size, hidden)   # :21
    w_dec = w_dec / w_dec.norm(dim=0, keepdim=True)  # :25
    return {
        "W_gate": w_gate, "b_gate": torch.zeros(hidden), "b_mag": torch.zeros(hidden),
        "r_mag": torch.zeros(hidden), "decoder.weight": w_dec, "decoder.bias": torch.zeros(act_size),
    }


SAE_MLP_KEYS = ("encoder.weight", "encoder.bias", "decoder.weight", "decoder.bias")
GATED_KEYS = ("W_gate", "b_gate", "b_mag", "r_mag", "decoder.weight", "decoder.bias")


# ----------------------------------------------------------------------------------------------- layout helpers
def to_tokens(x):
    """utils.py:2770-2774 reshape_tensor: [B,C,H,W] -> [(b h w), C]; 2-D passes through."""
    if x.dim() == 4:
        b, c, h, w = x.shape
        return x.permute(0, 2, 3, 1).reshape(b * h * w, c), True
    return x, False


def from_tokens(t, b, h, w):
    """'(b h w) c -> b c h w'  (utils.py:2462-2480)."""
    return t.reshape(b, h, w, t.shape[1]).permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------------------------- forward passes
def sae_mlp_forward(p, x):
    """models/sae_mlp.py:42-53 -> (encoder_output, decoder_output, encoder_output_prerelu), all token-major."""
    x_tok, _ = to_tokens(x)
    x_cent = x_tok - p["decoder.bias"]
    pre = F.linear(x_cent, p["encoder.weight"], p["encoder.bias"])
    enc = torch.relu(pre)
    dec = F.linear(enc, p["decoder.weight"], p["decoder.bias"])
    return enc, dec, pre


def gated_forward(p, x):
    """models/gated_sae.py:28-56 -> (encoder_output, decoder_output, relu_pi_gate, via_gate)."""
    x_tok, _ = to_tokens(x)
    x_cent = x_tok - p["decoder.bias"]
    pi_gate = F.linear(x_cent, p["W_gate"], p["b_gate"])
    # heaviside with value 0.5 at exactly 0, detached (:39)
    f_gate = torch.heaviside(pi_gate, torch.tensor([0.5])).detach()
    w_mag = torch.exp(p["r_mag"][:, None]) * p["W_gate"]          # :43 weight sharing
    f_mag = torch.relu(F.linear(x_cent, w_mag, p["b_mag"]))        # :44
    enc = f_gate * f_mag
    dec = F.linear(enc, p["decoder.weight"], p["decoder.bias"])
    relu_pi = torch.relu(pi_gate)
    with torch.no_grad():                                          # :53-54 frozen decoder, no gradient at all
        via = F.linear(relu_pi, p["decoder.weight"].detach().clone(), p["decoder.bias"].detach().clone())
    return enc, dec, relu_pi, via


# ----------------------------------------------------------------------------------------------- losses
def rmse_nrmse(decoded, targets):
    """losses/sparse_loss.py:4-21: per-channel RMSE over tokens, NRMSE = RMSE / (max - min of the target)."""
    mse_c = ((decoded - targets) ** 2).mean(dim=0)
    rng_c = targets.max(dim=0)[0] - targets.min(dim=0)[0]
    rmse_c = mse_c.sqrt()
    return rmse_c.mean(), (rmse_c / rng_c).mean()


def sparse_loss(encoded, decoded, targets):
    """losses/sparse_loss.py:30-61 -> (mse, mean|enc|, nrmse, rmse)."""
    assert decoded.shape == targets.shape and decoded.dim() == 2
    rec = ((decoded - targets) ** 2).mean()
    l1 = encoded.abs().mean()
    rmse, nrmse = rmse_nrmse(decoded, targets)
    return rec, l1, nrmse, rmse


def gated_loss(relu_pi, via, decoded, targets):
    """losses/sparse_loss.py:68-76 -> (mse, mean|relu_pi|, nrmse, rmse, aux mse)."""
    rec = ((decoded - targets) ** 2).mean()
    l1 = relu_pi.abs().mean()
    aux = ((via - targets) ** 2).mean()
    rmse, nrmse = rmse_nrmse(decoded, targets)
    return rec, l1, nrmse, rmse, aux


def sae_inference_and_loss(kind, p, output, lam):
    """utils.py:2448-2482.  Returns the reference's 9-tuple:
    (loss, rec, l1, nrmse, rmse, aux, enc [B,F,H,W], prerelu [B,F,H,W] | None, dec [B,C,H,W])."""
    sae_in, transformed = to_tokens(output)
    if kind == "sae_mlp":
        enc, dec, pre = sae_mlp_forward(p, sae_in)
    elif kind == "gated_sae":
        enc, dec, relu_pi, via = gated_forward(p, sae_in)
        pre = None
    else:
        raise ValueError(f"Unknown SAE model name {kind}.")
    if transformed:
        b, _, h, w = output.shape
        enc = from_tokens(enc, b, h, w)
        if pre is not None:
            pre = from_tokens(pre, b, h, w)
    if kind == "sae_mlp":
        rec, l1, nrmse, rmse = sparse_loss(enc, dec, sae_in)
        aux = torch.tensor(0)
        loss = rec + lam * l1
    else:
        rec, l1, nrmse, rmse, aux = gated_loss(relu_pi, via, dec, sae_in)
        loss = rec + lam * l1 + aux
    if transformed:
        dec = from_tokens(dec, b, h, w)
        assert dec.shape == output.shape
    return loss, rec, l1, nrmse, rmse, aux, enc, pre, dec


# ----------------------------------------------------------------------------------------------- activity metrics
def measure_inactive_units(output, expansion_factor):
    """utils.py:2032-2069 -> (dead mask [units] bool, sparsity float, activity frequency [units])."""
    zero = output == 0
    if output.dim() == 4:
        inactive = zero.all(dim=3).all(dim=2)      # a channel is inactive for an image iff every pixel is 0
    elif output.dim() == 2:
        inactive = zero
    else:
        raise ValueError(f"Output has unexpected shape {output.dim()}.")
    n_units = inactive.shape[1]
    dead = inactive.all(dim=0)
    freq = 1 - inactive.float().mean(dim=0)
    n_active = n_units - inactive.sum(dim=1)
    sparsity = (n_active / (n_units / expansion_factor)).mean().item()
    return dead, sparsity, freq


def variance_explained(output, decoder_output):
    """utils.py:2012-2030: 1 - mean Var_hw(dec) / mean Var_hw(x)   (unbiased variance)."""
    if output.dim() == 4:
        if decoder_output.dim() != 4:
            raise ValueError("Decoder output has unexpected shape.")
        var = output.var(dim=(2, 3)).mean()
        mod = decoder_output.var(dim=(2, 3)).mean()
    elif output.dim() == 2:
        if decoder_output.dim() != 2:
            raise ValueError("Decoder output has unexpected shape.")
        var = output.var(dim=1).mean()
        mod = decoder_output.var(dim=1).mean()
    else:
        raise ValueError("Output has unexpected shape.")
    return 1 - mod / var


def average_over_w_h(output, output_2=None):
    """utils.py:1996-2010."""
    if output.dim() == 4:
        output = output.mean(dim=(2, 3))
    if output_2 is not None and output_2.dim() == 4:
        output_2 = output_2.mean(dim=(2, 3))
    return output, output_2


# ----------------------------------------------------------------------------------------------- optimiser
def new_adam_state(p, keys):
    return {"step": 0, "m": {k: torch.zeros_like(p[k]) for k in keys},
            "v": {k: torch.zeros_like(p[k]) for k in keys}}


def adam_update(p, grads, state, keys, lr, betas, eps=1e-8):
    """torch.optim.Adam single-tensor math (no weight decay, no amsgrad), in place on p / state."""
    b1, b2 = betas
    state["step"] += 1
    t = state["step"]
    bc1 = 1 - b1 ** t
    bc2 = 1 - b2 ** t
    for k in keys:
        g = grads.get(k)
        if g is None:
            continue
        m, v = state["m"][k], state["v"][k]
        m.lerp_(g, 1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p[k].addcdiv_(m, denom, value=-(lr / bc1))


def optimizer_step(name, p, grads, state, keys, lr):
    """utils.py:50-97.  'adam' = Adam(betas=(0.9, 0.9999)); 'constrained_adam' = project the decoder-weight
    gradient orthogonal to each (unit) column, Adam(betas=(0.9, 0.999)), then renormalise the columns."""
    if name == "adam":
        adam_update(p, grads, state, keys, lr, (0.9, 0.9999))
    elif name == "constrained_adam":
        w = p["decoder.weight"]
        g = grads.get("decoder.weight")
        if g is not None:
            normed = w / w.norm(dim=0, keepdim=True)
            g = g - (g * normed).sum(dim=0, keepdim=True) * normed     # :72-74
            grads = dict(grads)
            grads["decoder.weight"] = g
        adam_update(p, grads, state, keys, lr, (0.9, 0.999))           # :75
        w /= w.norm(dim=0, keepdim=True)                                # :78-79
    else:
        raise ValueError(f"Unsupported optimizer: {name}")


# ----------------------------------------------------------------------------------------------- train step
def train_step(kind, p, state, output, lam, optimizer_name, lr, expansion_factor):
    """One pass of ModelPipeline.hook's train branch (model_pipeline.py:363-432) on one batch `output`
    ([B,C,H,W] or [N,C]): forward, loss, backward, optimiser step, then the per-batch metrics the hook stores.
    Mutates p / state in place.  Returns a dict of python floats / tensors."""
    keys = SAE_MLP_KEYS if kind == "sae_mlp" else GATED_KEYS
    output = output.detach()
    leaf = {k: p[k].detach().requires_grad_(True) for k in keys}
    loss, rec, l1, nrmse, rmse, aux, enc, pre, dec = sae_inference_and_loss(kind, leaf, output, lam)
    loss.backward()                                                          # :385
    grads = {k: leaf[k].grad for k in keys}
    optimizer_step(optimizer_name, p, grads, state, keys, lr)                 # :386
    enc, dec = enc.detach(), dec.detach()
    dead, sparsity, freq = measure_inactive_units(enc, expansion_factor)      # :418 -> :312
    var_expl = variance_explained(output, dec).item()                        # :420
    return {
        "loss": loss.item(), "rec": rec.item(), "l1": l1.item(), "nrmse": nrmse.item(), "rmse": rmse.item(),
        "aux": float(aux.item()), "dead": dead, "sparsity": sparsity, "freq": freq, "var_expl": var_expl,
        "enc": enc, "pre": None if pre is None else pre.detach(), "dec": dec, "grads": grads,
    }


# ----------------------------------------------------------------------------------------------- dead-neuron logic
def dead_neuron_action(train_batch_idx, dead_neurons_steps):
    """model_pipeline.py:771,792 evaluated AFTER train_batch_idx was incremented for the batch.
    'reinit' -> re-initialise the units dead over the last n steps, then clear the mask;
    'clear'  -> only clear the mask (the wait phase ended);  None -> keep AND-accumulating."""
    n, i = dead_neurons_steps, train_batch_idx
    if (i - 1) % n == 0 and ((i - 1) // n) % 2 == 0 and (i - 1) != 0:
        return "reinit"
    if i == n or (i > n and i % n == 0 and (i // n) % 2 == 1):
        return "clear"
    return None


def reset_encoder_weights(p, state, dead):
    """models/sae_mlp.py:79-184 without the index-file write.  `dead` is a bool [F] mask.  Draws from the global
    RNG exactly like the reference (full W_enc-shaped draw, then full W_dec-shaped draw).  Mutates p/state."""
    idx = torch.nonzero(dead).squeeze(-1)
    if idx.numel() == 0:
        return 0
    w_enc, b_enc, w_dec = p["encoder.weight"], p["encoder.bias"], p["decoder.weight"]
    new_w_enc = torch.nn.init.kaiming_uniform_(torch.zeros_like(w_enc))      # :106
    new_w_dec = torch.nn.init.kaiming_uniform_(torch.zeros_like(w_dec))      # :107
    live = torch.nonzero(~dead).squeeze(-1)
    avg_l2_enc = torch.norm(w_enc[live, :], p=2, dim=1).mean().item()        # :113-115
    avg_l2_dec = torch.norm(w_dec[:, live], p=2, dim=1).mean().item()        # :118-119 (dim=1 as in the reference)
    l2_b = b_enc[live].abs().mean().item()                                   # :121
    new_w_enc = new_w_enc / torch.norm(new_w_enc, p=2, dim=1, keepdim=True) * avg_l2_enc   # :125-126
    new_w_dec = new_w_dec / torch.norm(new_w_dec, p=2, dim=1, keepdim=True) * avg_l2_dec   # :127-128
    w_enc[idx, :] = new_w_enc[idx, :]                                        # :133
    w_dec[:, idx] = new_w_dec[:, idx]                                        # :134
    b_enc[idx] = l2_b                                                        # :130,135
    w_dec[:] = w_dec / w_dec.norm(dim=0, keepdim=True)                       # :138 renormalise ALL columns
    for key, rows in (("encoder.weight", True), ("encoder.bias", None), ("decoder.weight", False)):
        for mom in ("m", "v"):                                               # :148-176; Adam `step` untouched (:146)
            t = state[mom][key]
            if rows is None:
                t[idx] = 0
            elif rows:
                t[idx, :] = 0
            else:
                t[:, idx] = 0
    return int(idx.numel())


# ----------------------------------------------------------------------------------------------- indirect effects
def compute_ie_channel_wise(encoder_outputs, encoder_output_average, encoder_gradients, batch_size):
    """utils.py:2606-2637: ie[f] = mean_t | g[t,f] * (avg[f,h(t),w(t)] - a[t,f]) |, token order (b h w).
    encoder_outputs / encoder_gradients: [B*H*W, F]; encoder_output_average: [F,H,W]."""
    f, h, w = encoder_output_average.shape
    avg_tok = encoder_output_average.permute(1, 2, 0).reshape(1, h * w, f).expand(batch_size, h * w, f)
    avg_tok = avg_tok.reshape(batch_size * h * w, f)                # utils.py:2776-2782 without the copy
    return (encoder_gradients * (avg_tok - encoder_outputs)).abs().mean(dim=0)


def compute_ie_all_channels(sae_errors, sae_error_average, model_gradients, batch_size):
    """utils.py:2574-2602: mean_t | sum_c g[b,c,h,w] * (avg[c,h,w] - err[b,c,h,w]) |  -> scalar."""
    diff = sae_error_average.unsqueeze(0) - sae_errors
    return (model_gradients * diff).sum(dim=1).abs().mean()


def apply_sae(p, model_output, nodes=None, ablation=None):
    """utils.py:2786-2820 -> (encoder_output [T,F], decoder_output [B,C,H,W], new_decoder_output [B,C,H,W])."""
    b, _, h, w = model_output.shape
    x_tok, _ = to_tokens(model_output)
    enc, dec, _ = sae_mlp_forward(p, x_tok)
    if nodes is not None:
        new_enc = from_tokens(enc.clone(), b, h, w).clone()
        new_enc[..., ~nodes, :, :] = ablation[~nodes, :, :]
        new_tok, _ = to_tokens(new_enc)
        new_dec = F.linear(new_tok, p["decoder.weight"], p["decoder.bias"])
    else:
        new_dec = dec
    return enc, from_tokens(dec, b, h, w), from_tokens(new_dec, b, h, w)


def running_mean_update(old, new, num_samples, batch_size):
    """compute_ie.py:198-202,460-462 sample-weighted running average."""
    return (old * (num_samples - batch_size) + new * batch_size) / num_samples


def layer_averages(p, layer_output, expansion_factor):
    """compute_ie.py:146-162 for one layer and one batch -> batch means + dead mask + sparsity."""
    b, _, h, w = layer_output.shape
    x_tok, _ = to_tokens(layer_output)
    enc, dec, _ = sae_mlp_forward(p, x_tok)
    err = x_tok - dec
    dead, sparsity, _ = measure_inactive_units(enc, expansion_factor)    # NOTE: 2-D call, as in the reference (:155)
    return {
        "enc_avg": from_tokens(enc, b, h, w).mean(dim=0), "err_avg": from_tokens(err, b, h, w).mean(dim=0),
        "x_avg": layer_output.mean(dim=0), "dead": dead, "sparsity": sparsity,
    }


def node_ie_layer(p, x, grad_original, enc_avg, err_avg, x_avg):
    """compute_ie.py:242-267 + :442-453 for one layer of one batch, using the identity the reference's own check
    script asserts (nnsight_intervention_check.py:194-195,212-213): with stop-gradient on the SAE error and the
    pass-through gradient, d loss / d enc == rearrange(grad_original) @ W_dec exactly.
    x, grad_original: [B,C,H,W].  Returns (ie_sae_features [F], ie_sae_error scalar, ie_model_neurons [C])."""
    b = x.shape[0]
    x_tok, _ = to_tokens(x)
    g_tok, _ = to_tokens(grad_original)
    enc, dec, _ = sae_mlp_forward(p, x_tok)
    err = x - from_tokens(dec, *([b] + list(x.shape[2:])))
    enc_grad = g_tok @ p["decoder.weight"]
    ie_feat = compute_ie_channel_wise(enc, enc_avg, enc_grad, b)
    ie_err = compute_ie_all_channels(err, err_avg, grad_original, b)
    ie_neur = compute_ie_channel_wise(x_tok, x_avg, g_tok, b)
    return ie_feat, ie_err, ie_neur


def node_ie_layer_via_autograd(p, x, downstream):
    """The intervention itself (compute_ie.py:242-267) written with plain autograd instead of nnsight, for checking
    `node_ie_layer`'s identity: x_d = dec + (x - dec).detach(); loss = downstream(x_d); with the pass-through the
    gradient arriving at x_d equals the original model's d loss / d x.  Returns (enc, enc.grad, grad_original)."""
    x0 = x.detach().clone().requires_grad_(True)
    downstream(x0).backward()
    grad_original = x0.grad.detach()
    b, _, h, w = x.shape
    x_tok, _ = to_tokens(x.detach())
    x_cent = x_tok - p["decoder.bias"]
    pre = F.linear(x_cent, p["encoder.weight"], p["encoder.bias"])
    enc = torch.relu(pre)
    enc.retain_grad()
    enc_leaf = enc
    dec = from_tokens(F.linear(enc_leaf, p["decoder.weight"], p["decoder.bias"]), b, h, w)
    x_d = dec + (x.detach() - dec).detach()
    # pass-through gradient: overwrite whatever flows into x_d with grad_original (compute_ie.py:265)
    x_d.backward(grad_original)
    return enc.detach(), enc_leaf.grad.detach(), grad_original


# ----------------------------------------------------------------------------------------------- edge IE / faithfulness
def edge_ie_pass(layers, saes, feature_indices, enc_avg, err_avg, forward_from, loss_fn, inputs_list):
    """compute_ie.py:476-711 (compute_edge_ie) with plain autograd instead of nnsight -- a LITERAL restatement: for
    every pair of consecutive layers (u, d) the upstream layer is intervened on with the stop-gradient
    x_u~ = dec_u + (x_u - dec_u).detach() (:242-267, no pass-through), the downstream SAE is applied WITHOUT
    stop-gradient, and for every selected downstream feature j the scalar mean_t(grad_m_d[t, j] * enc_d[t, j]) is
    back-propagated to enc_u and dec_u (:589-611); the SAE error of d likewise (:633-650); the last layer's downstream
    node is the model loss (:672-700).  Parity unpinned (the reference needs nnsight + pretrained GoogLeNet).

    layers: ordered names; saes: {name: param dict}; forward_from(name_or_None, tensor) -> {name: raw layer output for
    every LATER layer, "out": logits} running the base model from the output of layer `name` (None: from the inputs);
    loss_fn(logits, targets); inputs_list: [(inputs, targets)].  Returns {name_u: [n_u + 1, n_d + 1]}."""
    vals = {}
    for i, nu in enumerate(layers):
        n_d = len(feature_indices[layers[i + 1]]) if i + 1 < len(layers) else 0
        vals[nu] = torch.zeros(len(feature_indices[nu]) + 1, n_d + 1)
    for batch_idx, (inputs, targets) in enumerate(inputs_list, start=1):
        b = inputs.shape[0]
        # gradients of the model loss w.r.t. every layer output of the un-intervened model (get_grad_original :270-311)
        acts = forward_from(None, inputs.detach())
        leaves = {}
        x = inputs.detach()
        outs = {}
        prev = None
        for name in layers:                                   # re-run layer by layer so that every output is a leaf
            o = forward_from(prev, x)[name]
            leaves[name] = o.detach().requires_grad_(True)
            outs[name] = leaves[name]
            x, prev = leaves[name], name
        logits = forward_from(prev, x)["out"]
        grad_orig = {}
        g = torch.autograd.grad(loss_fn(logits, targets), leaves[layers[-1]])[0]
        grad_orig[layers[-1]] = g
        for i in range(len(layers) - 2, -1, -1):              # chain the segments backwards
            nu, nd = layers[i], layers[i + 1]
            od = forward_from(nu, leaves[nu])[nd]
            grad_orig[nu] = torch.autograd.grad(od, leaves[nu], grad_outputs=grad_orig[nd])[0]
        del acts

        def upstream(nu, x_u):
            p = saes[nu]
            h, w = x_u.shape[2:]
            x_u = x_u.detach().requires_grad_(True)            # the reference's inputs require grad (:404)
            x_tok, _ = to_tokens(x_u)
            enc = torch.relu(F.linear(x_tok - p["decoder.bias"], p["encoder.weight"], p["encoder.bias"]))
            enc.retain_grad()
            dec = from_tokens(F.linear(enc, p["decoder.weight"], p["decoder.bias"]), b, h, w)
            dec.retain_grad()
            err = (x_u - dec).detach()                         # stop-gradient (:256-259)
            return enc, dec, err, dec + err

        def update(nu, enc_u, err_u, g_feat, g_err, col):
            sel = feature_indices[nu]
            ie_f = compute_ie_channel_wise(enc_u[:, sel], enc_avg[nu][sel], g_feat, b)
            ie_e = compute_ie_all_channels(err_u, err_avg[nu], g_err, b)
            batch_ie = torch.cat((ie_f, ie_e.reshape(1)))
            vals[nu][:, col] = batch_ie if batch_idx == 1 else (vals[nu][:, col] * (batch_idx - 1) + batch_ie) / batch_idx

        for i in range(len(layers) - 1):
            nu, nd = layers[i], layers[i + 1]
            pd = {k: v.detach() for k, v in saes[nd].items()}
            x_u = leaves[nu].detach()
            # grad of the loss w.r.t. the downstream encoder output under the node-IE intervention (:561-566)
            g_d = grad_orig[nd].detach()
            grad_m_d = to_tokens(g_d)[0] @ pd["decoder.weight"]
            enc_u, dec_u, err_u, x_t = upstream(nu, x_u)
            x_d = forward_from(nu, x_t)[nd]
            xd_tok, _ = to_tokens(x_d)
            enc_d = torch.relu(F.linear(xd_tok - pd["decoder.bias"], pd["encoder.weight"], pd["encoder.bias"]))
            dec_d = F.linear(enc_d, pd["decoder.weight"], pd["decoder.bias"])
            err_d = xd_tok - dec_d                              # no stop-gradient downstream (:583-586)
            for col, j in enumerate(feature_indices[nd]):
                prod = (grad_m_d[:, j] * enc_d[:, j]).mean()
                enc_u.grad = None
                dec_u.grad = None
                prod.backward(retain_graph=True)
                update(nu, enc_u.detach(), err_u, enc_u.grad[:, feature_indices[nu]].clone(), dec_u.grad.clone(), col)
            prod = (to_tokens(g_d)[0] * err_d).sum(dim=1).mean()
            enc_u.grad = None
            dec_u.grad = None
            prod.backward()
            update(nu, enc_u.detach(), err_u, enc_u.grad[:, feature_indices[nu]].clone(), dec_u.grad.clone(), -1)
        # last layer: the downstream node is the model loss (:672-700)
        nu = layers[-1]
        enc_u, dec_u, err_u, x_t = upstream(nu, leaves[nu].detach())
        loss_fn(forward_from(nu, x_t)["out"], targets).backward()
        update(nu, enc_u.detach(), err_u, enc_u.grad[:, feature_indices[nu]].clone(), dec_u.grad.clone(), 0)
    return vals


def faithfulness_pass(layers, saes, enc_avg, err_avg, x_avg, ie_feat, ie_err, ie_neur, run_with, loss_fn, inputs_list,
                      threshold, model_or_sae="sae"):
    """compute_ie.py:715-944 (compute_faithfulness): losses of the circuit (features / errors whose |IE| exceeds the
    threshold kept, the rest mean-ablated), of the circuit with all SAE errors zero- or mean-ablated, of the empty
    circuit and of the full model, averaged over batches; faithfulness = (m(C) - m(empty)) / (m(M) - m(empty)).
    run_with(inputs, fn) runs the base model replacing every layer's output by fn(name, output)."""
    nodes = {n: ie_feat[n].abs() > threshold for n in layers}
    err_nodes = {n: bool(abs(float(ie_err[n])) > threshold) for n in layers}
    neur_nodes = {n: ie_neur[n].abs() > threshold for n in layers}
    sums = {"zero": 0.0, "mean": 0.0, "C": 0.0, "empty": 0.0, "M": 0.0}
    with torch.no_grad():
        for inputs, targets in inputs_list:
            m = lambda fn: float(loss_fn(run_with(inputs, fn), targets))
            if model_or_sae == "sae":
                def circuit(kind):
                    def fn(name, x):
                        keep = nodes[name] if kind != "empty" else torch.zeros_like(nodes[name])
                        _, dec, new_dec = apply_sae(saes[name], x, nodes=keep, ablation=enc_avg[name])
                        if kind == "zero":
                            return new_dec
                        if kind in ("mean", "empty"):
                            return new_dec + err_avg[name]
                        err = x - dec
                        if not err_nodes[name]:
                            err = err_avg[name]
                        return new_dec + err
                    return fn
                for kind in ("zero", "mean", "C", "empty"):
                    sums[kind] += m(circuit(kind))
            else:
                def c_model(name, x):
                    x = x.clone()
                    x[:, ~neur_nodes[name]] = x_avg[name][~neur_nodes[name]]
                    return x
                sums["C"] += m(c_model)
                sums["empty"] += m(lambda name, x: x_avg[name].unsqueeze(0).expand_as(x).clone())
            sums["M"] += m(lambda name, x: x)
    n = len(inputs_list)
    avg = {k: v / n for k, v in sums.items()}
    out = {"m_C": avg["C"], "m_empty": avg["empty"], "m_M": avg["M"],
           "faithfulness": (avg["C"] - avg["empty"]) / (avg["M"] - avg["empty"])}
    if model_or_sae == "sae":
        out["faithfulness_sae_errors_zero_ablated"] = (avg["zero"] - avg["empty"]) / (avg["M"] - avg["empty"])
        out["faithfulness_sae_errors_mean_ablated"] = (avg["mean"] - avg["empty"]) / (avg["M"] - avg["empty"])
        out["m_C_zero"], out["m_C_mean"] = avg["zero"], avg["mean"]
    return out
