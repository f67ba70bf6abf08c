"""Generates tests/golden/*.npz by running the REAL reference modules (imported from /root/reference through
oracle/ref_import.py) on small seeded inputs.  Run in the build container only:

    python -m oracle.gen_golden

TEST INFRASTRUCTURE ONLY.  The fixtures are what pins oracle/sae_oracle.py (and, through it, the CUDA path) to the
reference's behaviour; the reference has no golden vectors of its own.
Every fixture stores the inputs, the reference's initial parameters and the reference's outputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _np(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy().copy()   # copy: parameters are updated in place later
    return np.asarray(t)


def _sd(model, prefix):
    return {f"{prefix}{k}": _np(v) for k, v in model.state_dict().items()}


def _planted_input(b, c, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn(b, c, h, w, generator=g))


def _train_scenario(ref, kind, x_batches, act_size, k, lam, opt_name, lr, seed, plant_dead=0):
    """Runs the reference hook's train branch (model_pipeline.py:380-420) step by step."""
    torch.manual_seed(seed)
    model = ref.SaeMLP(act_size, k) if kind == "sae_mlp" else ref.GatedSae(act_size, k)
    if plant_dead:
        with torch.no_grad():
            if kind == "sae_mlp":
                model.encoder.bias[:plant_dead] = -50.0
            else:
                model.b_gate[:plant_dead] = -50.0
                model.b_mag[:plant_dead] = -50.0
    out = _sd(model, "init.")
    crit_name = "sae_loss" if kind == "sae_mlp" else "gated_sae_loss"
    crit = ref.get_criterion(crit_name)
    opt, _ = ref.get_optimizer(opt_name, model, lr)
    for i, x in enumerate(x_batches):
        res = ref.sae_inference_and_loss(kind, model, crit_name, x, crit, lam)
        loss, rec, l1, nrmse, rmse, aux, enc, pre, dec = res
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if i == 0:
            for n, p in model.named_parameters():
                out[f"step0.grad.{n}"] = _np(p.grad)
        opt.step()
        opt.zero_grad(set_to_none=True)
        dead, sparsity, freq = ref.measure_inactive_units(enc.detach(), k)
        var_expl = ref.variance_explained(x, dec.detach()).item()
        out[f"step{i}.scalars"] = np.array(
            [loss.item(), rec.item(), l1.item(), nrmse.item(), rmse.item(), float(aux.item()), sparsity, var_expl],
            dtype=np.float64)
        out[f"step{i}.dead"] = _np(dead)
        out[f"step{i}.freq"] = _np(freq)
        if i == 0:
            out["step0.enc"] = _np(enc)
            out["step0.dec"] = _np(dec)
            if pre is not None:
                out["step0.pre"] = _np(pre)
    out.update(_sd(model, "final."))
    out["x"] = np.stack([_np(x) for x in x_batches])
    asz = int(np.prod(act_size)) if isinstance(act_size, (tuple, list)) else int(act_size)
    out["meta"] = np.array([asz, k, lam, lr, seed, plant_dead], dtype=np.float64)
    return out, model, opt


def main():
    ref = ref_import.load()
    os.makedirs(OUT, exist_ok=True)

    # ---- cfg1: SaeMLP((16,),4) on custom_mlp_9.fc1 activations of MNIST-shaped noise, plain 'adam', lambda 0.1
    torch.manual_seed(1234)
    base = ref.CustomMLP9((1, 28, 28))
    acts = []
    for _ in range(5):
        with torch.no_grad():
            inp = torch.randn(64, 1, 28, 28)
            acts.append(base.fc1(inp.view(-1, 784)))      # forward-hook output of fc1: [64,16]
    out, _, _ = _train_scenario(ref, "sae_mlp", acts, (16,), 4, 0.1, "adam", 1e-3, seed=0)
    np.savez_compressed(os.path.join(OUT, "cfg1_mlp_adam.npz"), **out)

    # ---- conv-shaped SaeMLP, constrained_adam, lambda 5 (cfg2 in miniature), with planted dead units
    xs = [_planted_input(4, 32, 5, 5, 100 + i) for i in range(4)]
    out, model, opt = _train_scenario(ref, "sae_mlp", xs, 32, 4, 5.0, "constrained_adam", 1e-3, seed=0, plant_dead=6)
    # dead-neuron re-initialisation right after these steps (sae_mlp.py:79-184)
    dead = torch.from_numpy(out["step3.dead"])
    for n, p in model.named_parameters():
        out[f"pre_reset.m.{n}"] = _np(opt.state[p]["exp_avg"])
        out[f"pre_reset.v.{n}"] = _np(opt.state[p]["exp_avg_sq"])
    torch.manual_seed(77)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            model.reset_encoder_weights(dead, "cpu", opt, 0, 4, 4, os.path.join(td, "idx.txt"))
    out.update(_sd(model, "reset."))
    for n, p in model.named_parameters():
        out[f"reset.m.{n}"] = _np(opt.state[p]["exp_avg"])
        out[f"reset.v.{n}"] = _np(opt.state[p]["exp_avg_sq"])
    np.savez_compressed(os.path.join(OUT, "conv_mlp_cadam.npz"), **out)

    # ---- GatedSae, constrained_adam, lambda 0.1 (cfg3 in miniature)
    xs = [_planted_input(4, 32, 5, 5, 200 + i) for i in range(3)]
    out, _, _ = _train_scenario(ref, "gated_sae", xs, 32, 4, 0.1, "constrained_adam", 1e-3, seed=0, plant_dead=5)
    np.savez_compressed(os.path.join(OUT, "conv_gated_cadam.npz"), **out)

    # ---- indirect-effect reductions + apply_sae with node ablation (utils.py:2574-2660, 2786-2820)
    torch.manual_seed(5)
    sae = ref.SaeMLP(24, 4)
    with torch.no_grad():
        sae.decoder.bias.normal_(0, 0.1)
        sae.encoder.bias.normal_(0, 0.1)
    x = _planted_input(3, 24, 4, 6, 300)
    g = torch.randn(3, 24, 4, 6, generator=torch.Generator().manual_seed(301))
    with torch.no_grad():
        enc, dec, new_dec = ref.apply_sae(sae, x)
        enc_avg = torch.randn(96, 4, 6, generator=torch.Generator().manual_seed(302)).abs()
        err_avg = torch.randn(24, 4, 6, generator=torch.Generator().manual_seed(303)) * 0.1
        x_avg = x.mean(dim=0)
        enc_grad = ref.reshape_tensor(g)[0] @ sae.decoder.weight
        ie_feat = ref.compute_ie_channel_wise(enc, enc_avg, enc_grad, 3)
        ie_err = ref.compute_ie_all_channels(x - dec, err_avg, g, 3)
        ie_neur = ref.compute_ie_channel_wise(ref.reshape_tensor(x)[0], x_avg, ref.reshape_tensor(g)[0], 3)
        nodes = torch.rand(96, generator=torch.Generator().manual_seed(304)) > 0.3
        enc2, dec2, new_dec2 = ref.apply_sae(sae, x, nodes=nodes, ablation=enc_avg)
    out = _sd(sae, "init.")
    out.update(x=_np(x), g=_np(g), enc=_np(enc), dec=_np(dec), enc_avg=_np(enc_avg), err_avg=_np(err_avg),
               x_avg=_np(x_avg), ie_feat=_np(ie_feat), ie_err=_np(ie_err), ie_neur=_np(ie_neur), nodes=_np(nodes),
               ablated_dec=_np(new_dec2))
    np.savez_compressed(os.path.join(OUT, "ie_small.npz"), **out)

    # ---- SaeConv API shell (models/sae_conv.py): forward of the 3x3 conv pair
    torch.manual_seed(9)
    conv = ref.SaeConv((8, 6, 6), 2)
    xc = torch.randn(2, 8, 6, 6, generator=torch.Generator().manual_seed(400))
    with torch.no_grad():
        ce, cd = conv(xc)
    out = _sd(conv, "init.")
    out.update(x=_np(xc), enc=_np(ce), dec=_np(cd))
    np.savez_compressed(os.path.join(OUT, "sae_conv.npz"), **out)

    # ---- re-init / wait schedule truth table (supplementary_files_1/reinitalize_dead_neurons_times.py)
    def ref_schedule(n, upto):
        # executes the reference's own truth-table script with its two constants substituted
        import contextlib, io
        path = os.path.join(ref_import.REFERENCE_ROOT, "supplementary_files_1", "reinitalize_dead_neurons_times.py")
        src = open(path).read().replace("dead_neurons_steps = 9912", f"dead_neurons_steps = {n}")
        src = src.replace("100000", str(upto))
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            exec(compile(src, path, "exec"), {})
        reinit = [int(l.split()[1]) for l in buf.getvalue().splitlines() if l.startswith("Re-initialize")]
        wait = [int(l.split()[1]) for l in buf.getvalue().splitlines() if l.startswith("Wait")]
        return np.array(reinit), np.array(wait)
    r1, w1 = ref_schedule(9912, 100000)
    r2, w2 = ref_schedule(8, 70)
    np.savez_compressed(os.path.join(OUT, "schedule.npz"), reinit_9912=r1, wait_9912=w1, reinit_8=r2, wait_8=w2)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
