"""Node indirect-effect (IE) attribution — the hot slice of the reference's compute_ie.py (IE.compute_average :95-226,
IE.intervention :242-267, IE.get_grad_original :270-311, IE.compute_node_ie :365-472) restated without nnsight.

The reference runs, per batch, one traced forward+backward of the base model to get d loss / d layer_output for every
layer, and then ANOTHER full forward+backward per SAE layer with the intervention  x_d = dec + (x - dec).detach()  and
the pass-through gradient.  Under exactly that intervention  d loss / d enc == rearrange(grad_original) @ W_dec  (the
reference's own check: supplementary_files_2/nnsight_intervention_check.py:194-195,212-213), so ONE forward+backward
with plain torch hooks yields x_l and g_l for every layer and the per-layer work becomes three GEMMs and three
reductions on the GPU (svb_node_ie_layer).  Edge IE / faithfulness (:476-944) are out of scope.

Data parallel: shard the images of each batch across ranks; every rank accumulates un-normalised sums (scale = 1)
and token counts; one all-reduce(SUM) per layer at the end gives the same global means as the reference's
sample-weighted running average (:455-462).  The batch-mean criterion is compensated per batch (see compute_node_ie);
tests/dp_ie_worker.py checks two ranks against one.
"""
import torch
import torch.distributed as dist

from . import ops
from .utils import measure_inactive_units


class IE:
    def __init__(self, model, layers, saes, exp_fac, model_criterion=None, device=None):
        """model: frozen base classifier (eval mode); layers: ordered {name: module} of the hooked layers
        (compute_ie.py:52); saes: {name: SaeMLP} (:63-72); exp_fac: {name: expansion factor}."""
        self.model = model
        self.layers = dict(layers)
        self.saes = saes
        self.exp_fac = exp_fac
        self.model_criterion = model_criterion or torch.nn.CrossEntropyLoss()
        self.device = device or next(model.parameters()).device
        for p in self.model.parameters():
            p.requires_grad = False

    # ------------------------------------------------------------------ activations (+ gradients) of every layer
    def _forward_collect(self, inputs, targets=None):
        acts, grads, handles = {}, {}, []

        def make_hook(name):
            def hook(_m, _i, out):
                acts[name] = out
                if targets is not None and out.requires_grad:
                    out.register_hook(lambda g, n=name: grads.__setitem__(n, g))
            return hook

        for name, module in self.layers.items():
            handles.append(module.register_forward_hook(make_hook(name)))
        try:
            if targets is None:
                with torch.no_grad():
                    self.model(inputs)
            else:
                inputs = inputs.detach().requires_grad_(True)        # compute_ie.py:404
                with torch.enable_grad():
                    out = self.model(inputs)
                    self.model_criterion(out, targets).backward()    # get_grad_original :299-311
        finally:
            for h in handles:
                h.remove()
        return {k: v.detach() for k, v in acts.items()}, grads

    # ------------------------------------------------------------------ compute_average (:95-226)
    def compute_average(self, batches):
        """batches: iterable of input tensors (or (inputs, ...) tuples).  Returns dicts keyed by layer:
        encoder_output_average [F,H,W], sae_error_average [C,H,W], original_layer_output_average [C,H,W],
        dead_units bool [F], sparsity float."""
        sums, n_samples = {}, 0
        for batch in batches:
            inputs = (batch[0] if isinstance(batch, (tuple, list)) else batch).to(self.device)
            if inputs.shape[0] == 0:
                continue
            bs = inputs.shape[0]
            n_samples += bs
            acts, _ = self._forward_collect(inputs)
            for name, x in acts.items():
                sae = self.saes[name]
                b, c, h, w = x.shape
                enc, dec, _ = ops.sae_forward(x, *[p.detach() for p in sae.param_list()], want_pre=False)
                dead, sparsity, _ = measure_inactive_units(enc, self.exp_fac[name])        # 2-D call, as at :155
                err_tok = x.permute(0, 2, 3, 1).reshape(-1, c).float() - dec
                enc_sum = enc.reshape(b, h * w, -1).sum(0).t().reshape(-1, h, w).contiguous()   # all-reduced later
                err_sum = err_tok.reshape(b, h * w, c).sum(0).t().reshape(c, h, w).contiguous()
                x_sum = x.float().sum(0)
                if name not in sums:
                    sums[name] = {"enc": enc_sum, "err": err_sum, "x": x_sum, "dead": dead, "sp": sparsity * bs}
                else:
                    s = sums[name]
                    s["enc"] += enc_sum
                    s["err"] += err_sum
                    s["x"] += x_sum
                    s["dead"] = s["dead"] & dead                                          # :201
                    s["sp"] += sparsity * bs
        n_global = self._all_reduce_counts(n_samples)
        out = {"encoder_output_average": {}, "sae_error_average": {}, "original_layer_output_average": {},
               "dead_units": {}, "sparsity": {}}
        if self._dp():
            # every rank must issue the same collectives, also one whose shards were all empty: such a rank learns the
            # layer shapes from its peers and contributes zeros (dead = all-True is the neutral element of the AND)
            shapes = {n: (s["enc"].shape[0],) + tuple(s["x"].shape) for n, s in sums.items()}
            gathered = [None] * dist.get_world_size()
            dist.all_gather_object(gathered, shapes)
            for g in gathered:
                for n, (f, c, h, w) in g.items():
                    if n not in sums:
                        z = lambda *s: torch.zeros(*s, device=self.device)
                        sums[n] = {"enc": z(f, h, w), "err": z(c, h, w), "x": z(c, h, w),
                                   "dead": torch.ones(f, dtype=torch.bool, device=self.device), "sp": 0.0}
        for name in self.layers:
            if name not in sums:
                continue
            s = sums[name]
            sp = torch.tensor([s["sp"]], device=self.device, dtype=torch.float64)
            dead_i = s["dead"].to(torch.int32)
            if self._dp():
                for t in (s["enc"], s["err"], s["x"], sp):
                    dist.all_reduce(t, op=dist.ReduceOp.SUM)
                dist.all_reduce(dead_i, op=dist.ReduceOp.MIN)
            out["encoder_output_average"][name] = s["enc"] / n_global
            out["sae_error_average"][name] = s["err"] / n_global
            out["original_layer_output_average"][name] = s["x"] / n_global
            out["dead_units"][name] = dead_i.bool()
            out["sparsity"][name] = float(sp.item()) / n_global
        return out

    # ------------------------------------------------------------------ compute_node_ie (:365-472)
    def compute_node_ie(self, batches, averages):
        """batches: iterable of (inputs, targets).  averages: the dict compute_average returned (what :372 loads).
        Returns (ie_sae_features {layer: [F]}, ie_sae_error {layer: scalar}, ie_model_neurons {layer: [C]})."""
        feat, err, neur, tokens = {}, {}, {}, {}
        for inputs, targets in batches:
            # The reference's criterion is a MEAN over the batch (utils.py:128-129), so d loss / d x carries 1 / B.  A
            # rank that holds bs of the batch's bs_global images gets 1 / bs from its local mean; IE is linear in the
            # gradient, so the rank's sums are rescaled by bs / bs_global.  (Collective: every rank calls it once per
            # batch, also with an empty shard.)
            bs = inputs.shape[0]
            bs_global = self._all_reduce_counts(bs)
            if bs == 0:
                continue
            ratio = bs / bs_global
            inputs, targets = inputs.to(self.device), targets.to(self.device)
            acts, grads = self._forward_collect(inputs, targets)
            for name, x in acts.items():
                sae = self.saes[name]
                f, e, n = ops.node_ie_layer(
                    x.float() if x.dtype not in (torch.float32, torch.bfloat16) else x, grads[name].to(x.dtype),
                    [p.detach() for p in sae.param_list()], averages["encoder_output_average"][name],
                    averages["sae_error_average"][name], averages["original_layer_output_average"][name], scale=1.0)
                if ratio != 1.0:
                    f, e, n = f * ratio, e * ratio, n * ratio
                t = x.shape[0] * x.shape[2] * x.shape[3]
                if name not in feat:
                    feat[name], err[name], neur[name], tokens[name] = f, e.reshape(1).clone(), n, t
                else:
                    feat[name] += f
                    err[name] += e
                    neur[name] += n
                    tokens[name] += t
        for name in self.layers:          # the same collectives on every rank (see compute_average)
            if name not in feat:
                if not self._dp():
                    continue
                n_f = self.saes[name].hidden_size
                n_c = self.saes[name].act_size
                feat[name] = torch.zeros(n_f, device=self.device)
                err[name] = torch.zeros(1, device=self.device)
                neur[name] = torch.zeros(n_c, device=self.device)
                tokens[name] = 0
            tg = self._all_reduce_counts(tokens[name])
            if tg == 0:                   # no rank saw this layer
                del feat[name], err[name], neur[name]
                continue
            if self._dp():
                for t in (feat[name], err[name], neur[name]):
                    dist.all_reduce(t, op=dist.ReduceOp.SUM)
            feat[name] = feat[name] / tg
            err[name] = (err[name] / tg)[0]
            neur[name] = neur[name] / tg
        return feat, err, neur

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _dp():
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _all_reduce_counts(self, n):
        if not self._dp():
            return n
        t = torch.tensor([n], device=self.device, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())
