"""The SAE-training slice of the reference's model_pipeline.py: the forward-hook train step (model_pipeline.py:363-432),
the per-batch metric capture (:278-360), the dead-neuron AND-accumulation and re-initialisation schedule (:744-793)
and the per-epoch checkpoint (:233-263, :1266-1280).  Base-model training, evaluation aggregation, plotting and MIS
are out of scope.

What changes relative to the reference: the train branch of the hook is ONE C-ABI call (svb_*_train_step) instead of
forward + criterion + autograd backward + optimizer.step() + three metric passes; its scalars stay on the device in a
stats block (one D2H copy when `batch_*` values are read) instead of six .item() syncs per step; the dead-unit mask is
AND-accumulated on the device.  The optimizer object is still a torch.optim.Adam-shaped ConstrainedAdam / Adam whose
state tensors (exp_avg / exp_avg_sq / step) the fused step updates in place, so reset_encoder_weights, state_dict()
and checkpoints behave as in the reference.
"""
import os

import torch

from . import ops
from ._lib import STAT, is_channels_last_tokens
from .parallel import DataParallelStep, global_counts
from .utils import (average_over_W_H, batch_top_k, get_criterion, get_optimizer, get_top_k_samples,
                    measure_inactive_units, sae_inference_and_loss, update_histogram, variance_explained)


def dead_neuron_action(train_batch_idx, dead_neurons_steps):
    """The two conditions of model_pipeline.py:771 and :792, evaluated after train_batch_idx was incremented.
    'reinit': re-initialise units dead over the last n steps, then clear; 'clear': clear only; None: keep measuring."""
    n, i = dead_neurons_steps, train_batch_idx
    if (i - 1) % n == 0 and ((i - 1) // n) % 2 == 0 and (i - 1) != 0:
        return "reinit"
    if i == n or (i > n and i % n == 0 and (i // n) % 2 == 1):
        return "clear"
    return None


class ModelPipeline:
    """Trains an SAE on the output of one layer of a frozen base model through a forward hook."""

    def __init__(self, model, sae_model, sae_model_name, sae_layer, sae_optimizer_name="constrained_adam",
                 sae_learning_rate=1e-3, sae_lambda_sparse=5.0, sae_expansion_factor=8, dead_neurons_steps=None,
                 device=None, reinit_index_dir=None, data_parallel=False, global_batch_images=None, model_copy=None,
                 model_criterion=None, compare_in_one_pass=False, cuda_graph=False):
        self.model = model
        # the unhooked original the modified model is compared with per batch (model_pipeline.py:694-708)
        self.model_copy = model_copy
        self.model_criterion = model_criterion or torch.nn.CrossEntropyLoss()
        self.batch_model_stats = None
        # compare_in_one_pass: instead of a second forward of an unhooked copy, the hook hands BOTH the reconstruction
        # and the original activation on (batch 2B) and the rest of the frozen network processes them together: the
        # layers in front of the SAE layer run once instead of twice and the logits of the original model fall out of
        # the same pass (a frozen eval-mode network treats samples independently, so the results are the same)
        self.compare_in_one_pass = compare_in_one_pass
        # cuda_graph: after three eager batches the whole training batch -- frozen base-model forward, the fused SAE
        # step inside the hook, the comparison with the original model -- is captured ONCE in a CUDA graph and replayed
        # for every later batch of the same shape (one launch instead of several hundred).  The Adam step count then
        # lives on the device (svb_opt_config::step_dev); the per-batch results are static tensors that the next
        # replay overwrites.  Data parallel: only with a fixed global batch (no per-step host exchange of the counts)
        # and the peer-memory exchange, whose exchange counter lives on the device too; every rank replays its graph
        # once per batch, in step.
        self.cuda_graph = bool(cuda_graph) and (not data_parallel or global_batch_images is not None)
        self._graph = None
        self._graph_eager_left = 3
        self._graph_static = None
        self._step_dev = None
        self._capturing = False
        self._graph_ws = 0
        self.graph_svb_launches = 0
        self.sae_model = sae_model
        self.sae_model_name = sae_model_name
        if sae_model_name not in ("sae_mlp", "gated_sae"):
            raise ValueError(f"Unknown SAE model name {sae_model_name}.")
        self.sae_criterion_name = "sae_loss" if sae_model_name == "sae_mlp" else "gated_sae_loss"
        self.sae_criterion = get_criterion(self.sae_criterion_name)
        self.sae_layer = sae_layer
        self.sae_optimizer_name = sae_optimizer_name
        self.sae_lambda_sparse = sae_lambda_sparse
        self.sae_expansion_factor = sae_expansion_factor
        self.dead_neurons_steps = dead_neurons_steps
        self.device = device or next(sae_model.parameters()).device
        self.sae_optimizer, _ = get_optimizer(sae_optimizer_name, sae_model, sae_learning_rate)
        if self.sae_optimizer.__class__.__name__ not in ("Adam", "ConstrainedAdam"):
            raise ValueError("the fused SAE step supports 'adam' and 'constrained_adam'")
        self.reinit_index_dir = reinit_index_dir
        self.dp = DataParallelStep(sae_model_name) if data_parallel else None
        # data parallel with a FIXED global batch (equal shards, drop_last): the caller states the global image count
        # and the hook skips the per-step count exchange (one tiny all-reduce + host sync, parallel.global_counts)
        self.global_batch_images = global_batch_images
        self.train_batch_idx = 0
        self.epoch_batch_idx = 0
        self.train_dead_neurons = {}
        self.hooks = []
        self.train_sae = True
        self._last = None
        for p in self.model.parameters():
            p.requires_grad = False
        if self.model_copy is not None:
            for p in self.model_copy.parameters():
                p.requires_grad = False
        # per-batch quantities, keyed like the reference (model_pipeline.py:394-420)
        self.batch_dead_units, self.batch_sparsity, self.batch_neuron_frequency = {}, {}, {}
        # eval-epoch recording (model_pipeline.py:80-104): top / small k samples per unit and activation histograms
        self.record_top_samples = False
        self.k = 25
        self.get_histogram = False
        self.histogram_info = {}
        self.batch_top_k_values, self.batch_top_k_indices = {}, {}
        self.batch_small_k_values, self.batch_small_k_indices = {}, {}
        self.top_k_samples, self.small_k_samples = {}, {}
        self.eval_dead_neurons = {}
        self.used_batch_size = None

    # ------------------------------------------------------------------ optimizer state shared with the fused step
    def _adam_tensors(self):
        params = self.sae_model.param_list()
        ms, vs = [], []
        for p in params:
            st = self.sae_optimizer.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            ms.append(st["exp_avg"])
            vs.append(st["exp_avg_sq"])
        steps = set()
        for p in params:
            st = self.sae_optimizer.state[p]
            if st["step"].is_cuda:         # load_state_dict(map_location=cuda) leaves the counters on the GPU: a .item()
                st["step"] = st["step"].cpu()   # there would synchronise the host with the device on every step
            if not self._capturing:        # a capture records the step, it does not take one
                st["step"] += 1
            steps.add(int(st["step"].item()))
        if len(steps) != 1:
            raise ValueError(f"the fused SAE step needs one Adam step count for all parameters, found {sorted(steps)}")
        return [p.data for p in params], ms, vs, steps.pop()

    # ------------------------------------------------------------------ the hook (model_pipeline.py:363-432)
    def hook(self, module, input, output, name, use_sae=True, train_sae=True):
        output = output.detach()
        if not (use_sae and name == self.sae_layer):
            return output
        group = self.sae_optimizer.param_groups[0]
        if train_sae:
            params, ms, vs, step = self._adam_tensors()
            kw = dict(optimizer=self.sae_optimizer_name, betas=group["betas"], eps=group["eps"])
            if self._capturing:
                kw["step_dev"] = self._step_dev
            both = None
            if self.compare_in_one_pass and output.dim() == 4:
                # [reconstruction; original] for the rest of the network: the step writes its reconstruction straight
                # into the first half of the 2B batch, the original is copied behind it -- no torch.cat pass
                fmt = torch.channels_last if is_channels_last_tokens(output) else torch.contiguous_format
                x_in = output if fmt == torch.channels_last else output.contiguous()
                both = torch.empty((2 * output.shape[0],) + tuple(output.shape[1:]), device=output.device,
                                   dtype=output.dtype, memory_format=fmt)
                kw["dec_out"] = both[:output.shape[0]]
                output = x_in
            if self.dp is not None:
                n_img = output.shape[0]
                hw = output.shape[2] * output.shape[3] if output.dim() == 4 else 1
                if self.global_batch_images:
                    g_img, g_tok = int(self.global_batch_images), int(self.global_batch_images) * hw
                else:
                    g_img, g_tok = global_counts(n_img, hw, device=output.device)
                if self._capturing and not self.dp.peer:
                    raise RuntimeError("cuda_graph=True with data_parallel=True needs the peer-memory exchange")
                res = self.dp.step(output, params, ms, vs, step, group["lr"], self.sae_lambda_sparse,
                                   self.sae_expansion_factor, self.sae_optimizer_name, group["betas"], g_img, g_tok,
                                   eps=group["eps"], step_dev=kw.get("step_dev"), dec_out=kw.get("dec_out"))
            elif self.sae_model_name == "sae_mlp":
                res = ops.sae_train_step(output, params, ms, vs, step, group["lr"], self.sae_lambda_sparse,
                                         self.sae_expansion_factor, **kw)
            else:
                res = ops.gated_train_step(output, params, ms, vs, step, group["lr"], self.sae_lambda_sparse,
                                           self.sae_expansion_factor, **kw)
            self._last = res
            self.batch_dead_units[(name, "sae")] = res.dead          # uint8 on the device; AND-ed in train_batch()
            self.batch_neuron_frequency[(name, "sae")] = res.freq
            if both is not None:
                both[output.shape[0]:].copy_(output)
                return both
            if self.compare_in_one_pass:
                return torch.cat((res.dec, output.to(res.dec.dtype)), dim=0)
            return res.dec                                           # replaces the layer output (:425,432)
        with torch.no_grad():
            r = sae_inference_and_loss(self.sae_model_name, self.sae_model, self.sae_criterion_name, output,
                                       self.sae_criterion, self.sae_lambda_sparse)
        loss, rec, l1, nrmse, rmse, aux, enc, pre, dec = r
        self._last = {"loss": loss, "rec": rec, "l1": l1, "nrmse": nrmse, "rmse": rmse, "aux": aux,
                      "var_expl": variance_explained(output, dec)}
        self.compute_and_store_batch_wise_metrics("sae", enc, name, self.sae_expansion_factor, output_2=pre)
        if (name, "sae") in self.batch_sparsity:
            self._last["sparsity"] = self.batch_sparsity[(name, "sae")]
        return dec.to(output.dtype)

    # ------------------------------------------------------------------ per-batch metrics (model_pipeline.py:278-360)
    def compute_and_store_batch_wise_metrics(self, model_key, output, name, expansion_factor=1, output_2=None):
        """Spatial means of the (pre-ReLU) activations -> histograms or activity metrics, and the k largest / smallest
        samples of every unit in this batch, all on the device."""
        output_avg, output_2_avg = average_over_W_H(output, output_2)
        if self.get_histogram:
            self.histogram_info = update_histogram(self.histogram_info, name, model_key, output_avg, self.device,
                                                   output_2=output_2_avg)
        else:
            dead, sparsity, freq = measure_inactive_units(output, expansion_factor)
            self.batch_sparsity[(name, model_key)] = sparsity
            self.batch_dead_units[(name, model_key)] = dead.to(torch.uint8)
            self.batch_neuron_frequency[(name, model_key)] = freq
        if self.record_top_samples:
            use_output = output_2_avg if output_2 is not None else output_avg          # :349-354
            k = min(self.k, use_output.shape[0])
            tv, ti, sv, si = batch_top_k(use_output, k)
            key = (name, model_key)
            self.batch_top_k_values[key], self.batch_top_k_indices[key] = tv, ti
            self.batch_small_k_values[key], self.batch_small_k_indices[key] = sv, si

    def eval_batch(self, inputs, filename_indices=None):
        """model_pipeline.py:603-664 + :862-909 for one evaluation batch: frozen forward with the SAE in inference
        mode, dead-unit AND over the epoch, and the merge of this batch's top / small k samples into the running ones.
        filename_indices: int64 [batch] dataset positions of the batch's samples (default: running sample numbers)."""
        self.epoch_batch_idx += 1
        if self.used_batch_size is None:
            self.used_batch_size = inputs.shape[0]
        with torch.no_grad():
            outputs = self.model(inputs)
        for key, dead in self.batch_dead_units.items():                                  # :870-878
            self.eval_dead_neurons[key] = dead if key not in self.eval_dead_neurons else self.eval_dead_neurons[key] & dead
        if self.record_top_samples:
            if filename_indices is None:
                filename_indices = torch.arange(inputs.shape[0], device=inputs.device) + \
                    (self.epoch_batch_idx - 1) * self.used_batch_size
            filename_indices = filename_indices.to(device=inputs.device, dtype=torch.int64)
            for key in self.batch_top_k_values:
                ti, si = self.batch_top_k_indices[key], self.batch_small_k_indices[key]
                if key not in self.top_k_samples:                                        # :881-893
                    self.top_k_samples[key] = (self.batch_top_k_values[key], ti, self.used_batch_size, filename_indices[ti])
                    self.small_k_samples[key] = (self.batch_small_k_values[key], si, self.used_batch_size,
                                                 filename_indices[si])
                else:                                                                    # :895-909
                    self.top_k_samples[key] = get_top_k_samples(
                        self.top_k_samples[key], self.batch_top_k_values[key], ti, filename_indices[ti],
                        self.epoch_batch_idx, largest=True, k=self.k)
                    self.small_k_samples[key] = get_top_k_samples(
                        self.small_k_samples[key], self.batch_small_k_values[key], si, filename_indices[si],
                        self.epoch_batch_idx, largest=False, k=self.k)
        return outputs

    def register_hooks(self, train_sae=True):
        """model_pipeline.py:445-475: a forward hook on the SAE layer."""
        self.remove_hooks()
        self.train_sae = train_sae
        module = dict(self.model.named_modules())[self.sae_layer]
        name = self.sae_layer
        self.hooks.append(module.register_forward_hook(
            lambda m, i, o: self.hook(m, i, o, name, use_sae=True, train_sae=self.train_sae)))

    def remove_hooks(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []

    def batch_scalars(self):
        """loss / rec / l1 / nrmse / rmse / aux / var_expl / sparsity of the last batch as python floats
        (the values the reference stores in batch_sae_* at model_pipeline.py:394-399,420) — one D2H copy."""
        if self._last is None:
            return {}
        if isinstance(self._last, dict):
            return {k: float(v) for k, v in self._last.items()}
        return self._last.scalars()

    # ------------------------------------------------------------------ one training batch + dead-neuron schedule
    def compare_with_original(self, inputs, outputs, targets=None, out_orig=None):
        """model_pipeline.py:694-708: the unhooked copy of the base model on the same inputs -> KL divergence between the
        two class distributions (sum over classes and images / batch size), the share of images both classify alike and
        (with targets) the loss difference.  Stays on the device: batch_model_stats = float32[3] (kld, same, loss_diff)."""
        with torch.no_grad():
            if out_orig is None:
                out_orig = self.model_copy(inputs)
            lp_orig = torch.nn.functional.log_softmax(out_orig.float(), dim=1)
            lp_mod = torch.nn.functional.log_softmax(outputs.float(), dim=1)
            kld = torch.nn.functional.kl_div(lp_orig, lp_mod, reduction="sum", log_target=True) / inputs.size(0)
            same = (out_orig.argmax(dim=1) == outputs.argmax(dim=1)).float().mean()
            if targets is not None:
                diff = self.model_criterion(outputs.float(), targets) - self.model_criterion(out_orig.float(), targets)
            else:
                diff = torch.zeros((), device=kld.device)
            self.batch_model_stats = torch.stack([kld, same, diff])
        return self.batch_model_stats

    def _forward_and_compare(self, inputs, targets):
        with torch.no_grad():
            outputs = self.model(inputs)
        if self.compare_in_one_pass and self.hooks and self.train_sae:
            outputs, out_orig = outputs[:inputs.shape[0]], outputs[inputs.shape[0]:]
            self.compare_with_original(inputs, outputs, targets, out_orig=out_orig)
        elif self.model_copy is not None:
            self.compare_with_original(inputs, outputs, targets)
        return outputs

    def _graphed_forward(self, inputs, targets):
        """Capture on the fourth batch, replay afterwards (see `cuda_graph` in __init__)."""
        from . import _lib as L
        ws = L.load().svb_workspace_bytes(L.handle(inputs.device))
        if self._graph is not None and ws != self._graph_ws:
            # another call on this device grew (= re-allocated) the library's workspace: the captured launches point into
            # the old one.  Drop the graph and capture again.
            self._graph, self._graph_eager_left = None, 1
        if self._graph is None:
            if self._graph_eager_left > 0:                 # cuDNN heuristics, workspace growth, lazy optimizer state
                self._graph_eager_left -= 1
                return self._forward_and_compare(inputs, targets)
            if self.dp is not None and not self.dp.peer:
                # the exchange fell back to torch.distributed (no peer memory on this machine: the same on every rank):
                # the data-parallel batch stays eager
                self.cuda_graph = False
                return self._forward_and_compare(inputs, targets)
            params = self.sae_model.param_list()
            host_step = int(self.sae_optimizer.state[params[0]]["step"].item())
            self._step_dev = torch.tensor([host_step], dtype=torch.int32, device=inputs.device)
            static_in = inputs.clone()
            static_tgt = None if targets is None else targets.clone()
            torch.cuda.synchronize(inputs.device)
            self._graph = torch.cuda.CUDAGraph()
            self._capturing = True
            launches0 = L.load().svb_launch_count()
            try:
                with torch.cuda.graph(self._graph):
                    static_out = self._forward_and_compare(static_in, static_tgt)
            finally:
                self._capturing = False
            # kernels of libsvb inside one replay (svb_launch_count only sees launches issued through the C ABI)
            self.graph_svb_launches = L.load().svb_launch_count() - launches0
            self._graph_static = (static_in, static_tgt, static_out, self._last, self.batch_model_stats,
                                  dict(self.batch_dead_units), dict(self.batch_neuron_frequency))
            self._graph_ws = L.load().svb_workspace_bytes(L.handle(inputs.device))
        static_in, static_tgt, static_out, last, model_stats, dead, freq = self._graph_static
        if tuple(inputs.shape) != tuple(static_in.shape):
            raise ValueError(f"cuda_graph=True replays batches of shape {tuple(static_in.shape)}; got {tuple(inputs.shape)}")
        static_in.copy_(inputs)
        if static_tgt is not None:
            static_tgt.copy_(targets)
        self._graph.replay()
        for p in self.sae_model.param_list():              # the host-side counters follow (checkpoints, re-initialisation)
            self.sae_optimizer.state[p]["step"] += 1
        self._last, self.batch_model_stats = last, model_stats
        self.batch_dead_units, self.batch_neuron_frequency = dict(dead), dict(freq)
        return static_out

    def train_batch(self, inputs, epoch=0, targets=None):
        """model_pipeline.py:603-793 for one batch: frozen base-model forward (the hook trains the SAE), the comparison
        with the unhooked copy when one was given, then train_batch_idx bookkeeping, dead-mask accumulation and the
        re-initialisation schedule."""
        self.epoch_batch_idx += 1
        graphed = self.cuda_graph and self.hooks and self.train_sae
        outputs = self._graphed_forward(inputs, targets) if graphed else self._forward_and_compare(inputs, targets)
        self.train_batch_idx += 1
        for key, dead in self.batch_dead_units.items():                      # :744-748 (AND == product of bools)
            if key not in self.train_dead_neurons:                          # (a replayed graph overwrites `dead`: copy)
                self.train_dead_neurons[key] = dead.clone() if graphed else dead
            else:
                self.train_dead_neurons[key] = self.train_dead_neurons[key] & dead
        action = None
        if self.dead_neurons_steps:
            action = dead_neuron_action(self.train_batch_idx, self.dead_neurons_steps)
            if action == "reinit":
                dead = self.train_dead_neurons[(self.sae_layer, "sae")].bool()
                file_path = None
                if self.reinit_index_dir:
                    os.makedirs(self.reinit_index_dir, exist_ok=True)
                    file_path = os.path.join(
                        self.reinit_index_dir,
                        f"epoch_{epoch}_train_batch_idx_{self.train_batch_idx}_epoch_batch_idx_{self.epoch_batch_idx}.txt")
                self.sae_model.reset_encoder_weights(dead, self.device, self.sae_optimizer, epoch,
                                                     self.train_batch_idx, self.epoch_batch_idx, file_path)
                self.train_dead_neurons = {}
            elif action == "clear":
                self.train_dead_neurons = {}
        return outputs, action

    # ------------------------------------------------------------------ checkpoints (reference dict keys)
    def save_checkpoint(self, path, epoch):
        torch.save({"epoch": epoch, "model_state_dict": self.sae_model.state_dict(),
                    "optimizer_state_dict": self.sae_optimizer.state_dict(),
                    "training_step": self.train_batch_idx}, path)                 # model_pipeline.py:1268-1273

    def load_checkpoint(self, path):
        ckpt = torch.load(path, map_location=self.device)
        self.sae_model.load_state_dict(ckpt["model_state_dict"])
        self.sae_optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        for st in self.sae_optimizer.state.values():      # Adam's step counters live on the host (no per-step sync)
            if torch.is_tensor(st.get("step")) and st["step"].is_cuda:
                st["step"] = st["step"].cpu()
        self.train_batch_idx = ckpt["training_step"]                              # model_pipeline.py:255-262
        return ckpt["epoch"]


STAT_NAMES = tuple(STAT)
