"""Node indirect-effect (IE) attribution — the hot slice of the reference's compute_ie.py (IE.compute_average :95-226,
IE.intervention :242-267, IE.get_grad_original :270-311, IE.compute_node_ie :365-472) restated without nnsight.

The reference runs, per batch, one traced forward+backward of the base model to get d loss / d layer_output for every
layer, and then ANOTHER full forward+backward per SAE layer with the intervention  x_d = dec + (x - dec).detach()  and
the pass-through gradient.  Under exactly that intervention  d loss / d enc == rearrange(grad_original) @ W_dec  (the
reference's own check: supplementary_files_2/nnsight_intervention_check.py:194-195,212-213), so ONE forward+backward
with plain torch hooks yields x_l and g_l for every layer and the per-layer work becomes three GEMMs and three
reductions on the GPU (svb_node_ie_layer).  Edge IE (:476-711) and faithfulness (:715-944) are restated on the same
building blocks (compute_edge_ie, compute_faithfulness below).

Data parallel: shard the images of each batch across ranks; every rank accumulates un-normalised sums (scale = 1)
and token counts; one all-reduce(SUM) per layer at the end gives the same global means as the reference's
sample-weighted running average (:455-462).  The batch-mean criterion is compensated per batch (see compute_node_ie);
tests/dp_ie_worker.py checks two ranks against one.
"""
import torch
import torch.distributed as dist

from . import ops
from .utils import measure_inactive_units_device


class IE:
    def __init__(self, model, layers, saes, exp_fac, model_criterion=None, device=None, cuda_graph=False):
        """model: frozen base classifier (eval mode); layers: ordered {name: module} of the hooked layers
        (compute_ie.py:52); saes: {name: SaeMLP} (:63-72); exp_fac: {name: expansion factor}.
        cuda_graph: compute_node_ie captures the work of one batch -- frozen forward, backward down to the first hooked
        layer, svb_node_ie_layer per layer: ~500 launches whose host side (Python, autograd, ctypes) takes longer than the
        3 ms they run -- ONCE per batch shape in a CUDA graph and replays it for every later batch of that shape."""
        self.cuda_graph = bool(cuda_graph)
        self._graphs = {}
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
        """One forward of the frozen model collecting the hooked layers' outputs and, with `targets`, the gradient of the
        loss at each of them (get_grad_original, compute_ie.py:270-311).  The reference makes the INPUT IMAGES require
        grad (:404) and back-propagates to them; only the gradients at the hooked layers are used, so here the first
        hooked layer the forward reaches hands a detached leaf on: the layers in front of it run without an autograd
        graph (and on the fused producer kernels when the model has them) and the backward stops there instead of going
        through the 112x112 / 56x56 stem.  The gradients at the hooked layers are the same numbers."""
        acts, grads, handles = {}, {}, []

        def make_hook(name):
            def hook(_m, _i, out):
                ret = None
                if targets is not None and not out.requires_grad:
                    out = ret = out.detach().requires_grad_(True)
                acts[name] = out
                if targets is not None:
                    out.register_hook(lambda g, n=name: grads.__setitem__(n, g))
                return ret
            return hook

        for name, module in self.layers.items():
            handles.append(module.register_forward_hook(make_hook(name)))
        try:
            if targets is None:
                with torch.no_grad():
                    self.model(inputs)
            else:
                with torch.enable_grad():
                    out = self.model(inputs.detach())
                    self.model_criterion(out, targets).backward()    # get_grad_original :299-311
        finally:
            for h in handles:
                h.remove()
        return {k: v.detach() for k, v in acts.items()}, grads

    # ------------------------------------------------------------------ compute_average (:95-226)
    def _batch_average(self, inputs):
        """One batch: {layer: (sum of encoder outputs [F,H,W], of SAE errors [C,H,W], of layer outputs [C,H,W] over the
        images, dead units bool [F], sparsity as a device scalar)} -- no host synchronisation."""
        acts, _ = self._forward_collect(inputs)
        out = {}
        for name, x in acts.items():
            sae = self.saes[name]
            b, c, h, w = x.shape
            enc, dec, _ = ops.sae_forward(x, *[p.detach() for p in sae.param_list()], want_pre=False,
                                          out_dtype=torch.bfloat16)
            dead, sparsity, _ = measure_inactive_units_device(enc, self.exp_fac[name])     # 2-D call, as at :155
            # per-position sums over the images on libsvb (one read of each tensor); the SAE error x - dec is
            # never materialised: its sum is the difference of the two sums
            enc_sum = ops.image_sum(enc, b).t().reshape(-1, h, w).contiguous()             # [F,H,W]; all-reduced later
            x_sum = ops.image_sum(x if x.dtype in (torch.float32, torch.bfloat16) else x.float(), b)   # [C,H,W]
            err_sum = x_sum - ops.image_sum(dec, b).t().reshape(c, h, w)
            out[name] = (enc_sum, err_sum, x_sum, dead, sparsity)
        return out

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
            per_layer = (self._graphed(self._batch_average, inputs) if self.cuda_graph and inputs.is_cuda
                         else self._batch_average(inputs))
            for name, (enc_sum, err_sum, x_sum, dead, sp) in per_layer.items():
                if name not in sums:
                    sums[name] = {"enc": enc_sum.clone(), "err": err_sum.clone(), "x": x_sum.clone(), "dead": dead.clone(),
                                  "sp": sp.double() * bs}
                else:
                    s = sums[name]
                    s["enc"] += enc_sum
                    s["err"] += err_sum
                    s["x"] += x_sum
                    s["dead"] = s["dead"] & dead                                          # :201
                    s["sp"] = s["sp"] + sp.double() * bs
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
                                   "dead": torch.ones(f, dtype=torch.bool, device=self.device),
                                   "sp": torch.zeros((), dtype=torch.float64, device=self.device)}
        for name in self.layers:
            if name not in sums:
                continue
            s = sums[name]
            sp = s["sp"].reshape(1).to(device=self.device, dtype=torch.float64)
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
    def _batch_node_ie(self, inputs, targets, averages):
        """One batch: {layer: (ie_features [F], ie_error [1], ie_neurons [C], tokens)} as un-normalised sums."""
        acts, grads = self._forward_collect(inputs, targets)
        out = {}
        for name, x in acts.items():
            sae = self.saes[name]
            f, e, n = ops.node_ie_layer(
                x.float() if x.dtype not in (torch.float32, torch.bfloat16) else x, grads[name].to(x.dtype),
                [p.detach() for p in sae.param_list()], averages["encoder_output_average"][name],
                averages["sae_error_average"][name], averages["original_layer_output_average"][name], scale=1.0)
            out[name] = (f, e.reshape(1), n, x.shape[0] * x.shape[2] * x.shape[3])
        return out

    def _graphed(self, fn, *tensors, extra=(), key=()):
        """fn(*tensors, *extra) captured once per (function, tensor shapes / strides, key) in a CUDA graph and replayed:
        the tensors are copied into the capture's static inputs, the result is the capture's static output (read it
        before the next replay).  Two eager runs on a side stream come first (cuDNN plans, autograd buffers, growth of
        the library's workspace); a re-allocated workspace invalidates the capture."""
        from . import _lib as L
        dev = tensors[0].device
        k = (fn.__name__, tuple((tuple(t.shape), t.dtype, t.stride()) for t in tensors), tuple(key))
        ws = L.load().svb_workspace_bytes(L.handle(dev))
        entry = self._graphs.get(k)
        if entry is not None and entry[3] != ws:
            entry = None
        if entry is None:
            static = [t.clone() for t in tensors]
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn(*static, *extra)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = fn(*static, *extra)
            entry = (graph, static, static_out, L.load().svb_workspace_bytes(L.handle(dev)))
            self._graphs[k] = entry
        graph, static, static_out, _ = entry
        for dst, src in zip(static, tensors):
            dst.copy_(src)
        graph.replay()
        return static_out

    def _graphed_batch_node_ie(self, inputs, targets, averages):
        key = tuple(averages["encoder_output_average"][n].data_ptr() for n in self.layers
                    if n in averages["encoder_output_average"])      # the averages are baked into the capture by address
        return self._graphed(self._batch_node_ie, inputs, targets, extra=(averages,), key=key)

    def compute_node_ie(self, batches, averages):
        """batches: iterable of (inputs, targets).  averages: the dict compute_average returned (what :372 loads).
        Returns (ie_sae_features {layer: [F]}, ie_sae_error {layer: scalar}, ie_model_neurons {layer: [C]})."""
        feat, err, neur, tokens = {}, {}, {}, {}
        # Data parallel with the batches at hand as a list: ONE all-reduce of all local batch sizes up front instead of
        # one (host-synchronising) exchange per batch -- at 2.3 ms of device work per batch that exchange was a fifth of
        # the 8-GPU pass.  Every rank passes the same number of batches either way (see below).
        global_sizes = None
        if self._dp() and isinstance(batches, (list, tuple)) and len(batches) > 0:
            t = torch.tensor([int(b[0].shape[0]) for b in batches], device=self.device, dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            global_sizes = t.tolist()
        for i_batch, (inputs, targets) in enumerate(batches):
            # The reference's criterion is a MEAN over the batch (utils.py:128-129), so d loss / d x carries 1 / B.  A
            # rank that holds bs of the batch's bs_global images gets 1 / bs from its local mean; IE is linear in the
            # gradient, so the rank's sums are rescaled by bs / bs_global.  (Collective: every rank calls it once per
            # batch, also with an empty shard.)
            bs = inputs.shape[0]
            bs_global = global_sizes[i_batch] if global_sizes is not None else self._all_reduce_counts(bs)
            if bs == 0:
                continue
            ratio = bs / bs_global
            inputs, targets = inputs.to(self.device), targets.to(self.device)
            per_layer = (self._graphed_batch_node_ie if self.cuda_graph and inputs.is_cuda
                         else self._batch_node_ie)(inputs, targets, averages)
            for name, (f, e, n, t) in per_layer.items():
                if ratio != 1.0:
                    f, e, n = f * ratio, e * ratio, n * ratio
                if name not in feat:
                    feat[name], err[name], neur[name], tokens[name] = f.clone(), e.reshape(1).clone(), n.clone(), t
                else:
                    feat[name] += f
                    err[name] += e.reshape(1)
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

    # ------------------------------------------------------------------ compute_edge_ie (:476-711)
    def _forward_segments(self, inputs, names):
        """One forward of the frozen model in which every layer of `names` hands a detached leaf on: outs[n] is the
        layer's raw output (its graph reaches back to the previous leaf only), leaves[n] the leaf the rest of the
        network consumed.  The network is thereby cut into differentiable segments leaf_i -> out_{i+1}."""
        outs, leaves, handles = {}, {}, []

        def make_hook(name):
            def hook(_m, _i, out):
                outs[name] = out
                leaves[name] = out.detach().requires_grad_(True)
                return leaves[name]
            return hook

        for name in names:
            handles.append(self.layers[name].register_forward_hook(make_hook(name)))
        try:
            with torch.enable_grad():
                logits = self.model(inputs)
        finally:
            for h in handles:
                h.remove()
        return outs, leaves, logits

    def compute_edge_ie(self, batches, averages, custom_layers, feature_indices):
        """compute_ie.py:476-711.  custom_layers: ordered subset of the hooked layers; feature_indices: {layer: [SAE
        feature indices]} (:81-88).  For every pair of consecutive layers (u, d), every selected downstream feature j
        and the downstream SAE error, the reference back-propagates  mean_t(dloss/d node_d * node_d)  from d to the
        upstream encoder output / SAE error (one traced forward + backward per node) and reduces the result with
        compute_ie_channel_wise / compute_ie_all_channels; the last layer's downstream node is the model loss.

        Here ONE forward per batch cuts the network into segments; the cotangent of every downstream node at the
        downstream layer output is written down in closed form
            feature j : 1/T * (g_d W_dec)[t, j] * 1[enc_d[t, j] > 0] * W_enc[j, :]
            SAE error : 1/T * (g_d - ((g_d W_dec) * 1[enc_d > 0]) W_enc)          (no stop-gradient downstream, :583-586)
        pulled back to the upstream layer by one vector-Jacobian product of the segment (cuDNN, the base model is a
        library), and the upstream part -- encoder, g W_dec, the three reductions -- is svb_node_ie_layer.
        Returns {name_u: float32 [len(features_u) + 1, len(features_d) + 1]} (rows: features then SAE error; columns:
        downstream features then downstream SAE error; one column for the model loss at the last layer), averaged over
        the batches like :357-360 (every batch counts once)."""
        from . import ops as O
        names = list(custom_layers)
        vals = {}
        for i, nu in enumerate(names):
            n_d = len(feature_indices[names[i + 1]]) if i + 1 < len(names) else 0
            vals[nu] = torch.zeros(len(feature_indices[nu]) + 1, n_d + 1, device=self.device)
        sel = {n: torch.as_tensor(list(feature_indices[n]), dtype=torch.long, device=self.device) for n in names}
        bf = torch.bfloat16
        batch_idx = 0
        for inputs, targets in batches:
            if inputs.shape[0] == 0:
                continue
            batch_idx += 1
            inputs, targets = inputs.to(self.device), targets.to(self.device)
            outs, leaves, logits = self._forward_segments(inputs, names)
            with torch.enable_grad():
                loss = self.model_criterion(logits, targets)
            g = {names[-1]: torch.autograd.grad(loss, leaves[names[-1]], retain_graph=True)[0]}
            for i in range(len(names) - 2, -1, -1):                      # get_grad_original (:270-311), segment by segment
                g[names[i]] = torch.autograd.grad(outs[names[i + 1]], leaves[names[i]], grad_outputs=g[names[i + 1]],
                                                  retain_graph=True)[0]

            def upstream_ie(nu, grad_u):
                x_u = leaves[nu].detach()
                x_u = x_u.float() if x_u.dtype not in (torch.float32, bf) else x_u
                f, e, _ = O.node_ie_layer(x_u, grad_u.to(x_u.dtype), [p.detach() for p in self.saes[nu].param_list()],
                                          averages["encoder_output_average"][nu], averages["sae_error_average"][nu],
                                          averages["original_layer_output_average"][nu])
                return torch.cat((f[sel[nu]], e.reshape(1)))

            def update(nu, col, batch_ie):
                vals[nu][:, col] = batch_ie if batch_idx == 1 else (vals[nu][:, col] * (batch_idx - 1) + batch_ie) / batch_idx

            for i in range(len(names) - 1):
                nu, nd = names[i], names[i + 1]
                sae_d = self.saes[nd]
                w_enc, b_enc, w_dec, b_dec = [p.detach() for p in sae_d.param_list()]
                x_d = outs[nd]
                b, c, h, w = x_d.shape
                t_d = b * h * w
                xd = x_d.detach()
                enc_d, _, _ = O.sae_forward(xd.float() if xd.dtype not in (torch.float32, bf) else xd, w_enc, b_enc, w_dec,
                                            b_dec, want_pre=False, want_dec=False)
                active = enc_d > 0                                              # relu'(pre) = 1[pre > 0] = 1[enc > 0]
                g_tok = g[nd].detach().permute(0, 2, 3, 1).reshape(t_d, c).float()
                g_enc = O.gemm_bf16(g_tok.to(bf), w_dec.to(bf), b_mn=True)      # [T, F] = g W_dec (d loss / d enc_d, :561-566)

                def pull_back(v_tok):
                    v = v_tok.reshape(b, h, w, c).permute(0, 3, 1, 2).to(x_d.dtype)
                    return torch.autograd.grad(x_d, leaves[nu], grad_outputs=v, retain_graph=True)[0]

                for col, j in enumerate(feature_indices[nd]):
                    coef = g_enc[:, j] * active[:, j] / t_d                     # [T]
                    update(nu, col, upstream_ie(nu, pull_back(coef[:, None] * w_enc[j][None, :])))
                back = O.gemm_bf16((g_enc * active).to(bf), w_enc.to(bf), b_mn=True)   # [T, C] = ((g W_dec) * mask) W_enc
                update(nu, -1, upstream_ie(nu, pull_back((g_tok - back) / t_d)))
                del enc_d, active, g_enc, back
            update(names[-1], 0, upstream_ie(names[-1], g[names[-1]]))           # downstream node = model loss (:672-700)
        return vals

    # ------------------------------------------------------------------ compute_faithfulness (:715-944)
    def _loss_with(self, inputs, targets, replace):
        """Model loss with every hooked layer's output replaced by replace(name, output) (layers in network order, so
        later layers see the modified activations, as in the reference's sequential nnsight interventions)."""
        handles = [m.register_forward_hook(lambda _m, _i, out, n=n: replace(n, out)) for n, m in self.layers.items()]
        try:
            with torch.no_grad():
                return self.model_criterion(self.model(inputs), targets)
        finally:
            for h in handles:
                h.remove()

    def compute_faithfulness(self, batches, averages, node_ie, feature_node_threshold=0.0, model_or_sae="sae"):
        """compute_ie.py:715-944.  node_ie = (ie_sae_features, ie_sae_error, ie_model_neurons) as compute_node_ie returns
        them.  Nodes whose |IE| exceeds the threshold form the circuit C; everything else is mean-ablated
        (apply_sae(nodes=, ablation=), utils.py:2795-2807).  Returns the batch-averaged losses m(C), m(C) with all SAE
        errors zero- / mean-ablated, m(empty), m(M) and the faithfulness values (m(C) - m(empty)) / (m(M) - m(empty))."""
        from .utils import apply_sae
        ie_feat, ie_err, ie_neur = node_ie
        thr = feature_node_threshold                                           # error threshold = feature threshold (:722)
        nodes = {n: ie_feat[n].abs() > thr for n in self.layers}
        err_nodes = {n: bool(abs(float(ie_err[n])) > thr) for n in self.layers}
        neur_nodes = {n: ie_neur[n].abs() > thr for n in self.layers}
        enc_avg, err_avg = averages["encoder_output_average"], averages["sae_error_average"]
        x_avg = averages["original_layer_output_average"]
        sums = torch.zeros(5, device=self.device, dtype=torch.float64)          # zero, mean, C, empty, M
        n_batches = 0

        def circuit(kind):
            def fn(name, x):
                keep = nodes[name] if kind != "empty" else torch.zeros_like(nodes[name])
                _, dec, new_dec = apply_sae(self.saes[name], x, nodes=keep, ablation=enc_avg[name])
                if kind == "zero":
                    out = new_dec
                elif kind in ("mean", "empty"):
                    out = new_dec + err_avg[name]
                else:
                    out = new_dec + ((x - dec) if err_nodes[name] else err_avg[name])
                return out.to(x.dtype)
            return fn

        def model_circuit(name, x):
            x = x.clone()
            x[:, ~neur_nodes[name]] = x_avg[name][~neur_nodes[name]].to(x.dtype)
            return x

        for inputs, targets in batches:
            if inputs.shape[0] == 0:
                continue
            n_batches += 1
            inputs, targets = inputs.to(self.device), targets.to(self.device)
            if model_or_sae == "sae":
                for k, kind in enumerate(("zero", "mean", "C", "empty")):
                    sums[k] += self._loss_with(inputs, targets, circuit(kind)).double()
            else:
                sums[2] += self._loss_with(inputs, targets, model_circuit).double()
                sums[3] += self._loss_with(inputs, targets,
                                           lambda n, x: x_avg[n].to(x.dtype).unsqueeze(0).expand_as(x).clone()).double()
            with torch.no_grad():
                sums[4] += self.model_criterion(self.model(inputs), targets).double()
        zero, mean, m_c, empty, m_m = (sums / max(n_batches, 1)).tolist()          # ONE device->host copy
        out = {"m_C": m_c, "m_empty": empty, "m_M": m_m, "faithfulness": (m_c - empty) / (m_m - empty),
               "feature_node_threshold": thr, "error_node_threshold": thr}
        if model_or_sae == "sae":
            out["m_C_zero"], out["m_C_mean"] = zero, mean
            out["faithfulness_sae_errors_zero_ablated"] = (zero - empty) / (m_m - empty)
            out["faithfulness_sae_errors_mean_ablated"] = (mean - empty) / (m_m - empty)
            out["nodes_in_circuit"] = {n: int(v.sum()) for n, v in nodes.items()}
        return out

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
