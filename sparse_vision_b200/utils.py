"""Hot-path slice of the reference's utils.py with the same names, argument meaning and error behaviour, running on
libsvb: optimisers (utils.py:50-97), criteria (:127-137), SAE step glue (:2448-2482), activity metrics (:1996-2069),
indirect-effect reductions (:2574-2660) and apply_sae (:2786-2820).  Everything else in the reference's utils.py
(data loading, logging, plotting, MIS) is out of scope.
"""
import logging

import os

import torch
import torch.nn as nn

from . import ops
from .losses.sparse_loss import GatedSAELoss, SparseLoss
from .models.gated_sae import GatedSae  # noqa: F401
from .models.sae_conv import SaeConv  # noqa: F401
from .models.sae_mlp import SaeMLP  # noqa: F401


# --------------------------------------------------------------------------------------------------- optimisers
class _SvbAdamMixin:
    """torch.optim.Adam front (param_groups, state dict layout, load/save) whose step() runs on libsvb."""

    _svb_optimizer = "adam"
    _svb_constrained = None

    def _svb_step(self):
        for group in self.param_groups:
            b1, b2 = group["betas"]
            # torch.optim.Adam keeps one step count PER PARAMETER (a parameter that had no gradient in some step, or
            # state loaded from a checkpoint, may be at a different count): one libsvb call per distinct count
            by_step = {}
            for p in group["params"]:
                if p.grad is None and p is not self._svb_constrained:
                    continue
                if not p.is_cuda:
                    raise ValueError("sparse_vision_b200 optimizers update CUDA parameters only (no CPU fallback)")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if st["step"].is_cuda:                        # e.g. after load_state_dict(map_location=cuda)
                    st["step"] = st["step"].cpu()
                if p.grad is not None:
                    st["step"] += 1
                # the constrained tensor without a gradient is only renormalised (utils.py:76-79): any step count does
                by_step.setdefault(max(int(st["step"].item()), 1), []).append(p)
            for step_no, plist in by_step.items():
                grads = [p.grad for p in plist]
                dec_index = next((i for i, p in enumerate(plist) if p is self._svb_constrained), -1)
                new_grads = ops.adam_step([p.data for p in plist], grads, [self.state[p]["exp_avg"] for p in plist],
                                          [self.state[p]["exp_avg_sq"] for p in plist], step_no, group["lr"], (b1, b2),
                                          group["eps"], optimizer=self._svb_optimizer, decoder_index=dec_index)
                # the projected decoder gradient is visible on p.grad afterwards, as in the reference (utils.py:74)
                if dec_index >= 0 and grads[dec_index] is not None and new_grads[dec_index] is not grads[dec_index]:
                    grads[dec_index].copy_(new_grads[dec_index])


class Adam(_SvbAdamMixin, torch.optim.Adam):
    """torch.optim.Adam drop-in (class name 'Adam' is what sae_mlp.py:143 checks)."""

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._svb_step()
        return loss


class ConstrainedAdam(_SvbAdamMixin, torch.optim.Adam):
    """Adam whose `constrained_params` tensor (the decoder weight) keeps unit-norm columns: the gradient is projected
    orthogonal to each column before the step and the columns are renormalised after it (utils.py:50-81)."""

    _svb_optimizer = "constrained_adam"

    def __init__(self, params, constrained_params, lr):
        super().__init__(params, lr=lr, betas=(0.9, 0.999))
        self.p = constrained_params
        self._svb_constrained = constrained_params

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if self.p.grad is not None and self.p.norm(dim=0).min() < 1e-6:
            logging.warning(f"Constrained parameter {self.p} has a norm smaller than 1e-6")
        self._svb_step()
        return loss


def get_optimizer(optimizer_name, model, learning_rate):
    """utils.py:84-97 -> (optimizer, scheduler | None)."""
    if optimizer_name == "adam":
        return Adam(model.parameters(), lr=learning_rate, betas=(0.9, 0.9999)), None
    if optimizer_name == "sgd":
        return torch.optim.SGD(model.parameters(), lr=learning_rate), None
    if optimizer_name == "sgd_w_scheduler":
        opt = torch.optim.SGD(model.parameters(), lr=learning_rate, momentum=0.9)
        return opt, torch.optim.lr_scheduler.StepLR(opt, step_size=7, gamma=0.1)
    if optimizer_name == "constrained_adam":
        return ConstrainedAdam(model.parameters(), model.decoder.weight, lr=learning_rate), None
    raise ValueError(f"Unsupported optimizer: {optimizer_name}")


class CustomCrossEntropyLoss(nn.Module):
    """utils.py:99-125: negative log-likelihood of already-normalised class probabilities."""

    def forward(self, logits, targets):
        picked = torch.gather(logits, 1, targets.unsqueeze(1)).squeeze(1)
        return torch.mean(-torch.log(picked + 1e-40))


def get_criterion(criterion_name):
    """utils.py:127-137."""
    if criterion_name == "cross_entropy":
        return nn.CrossEntropyLoss()
    if criterion_name == "sae_loss":
        return SparseLoss()
    if criterion_name == "gated_sae_loss":
        return GatedSAELoss()
    if criterion_name == "negative_log_likelihood":
        return CustomCrossEntropyLoss()
    raise ValueError(f"Unsupported criterion: {criterion_name}")


def load_model(model_name, img_size=None, num_classes=None, expansion_factor=None, execution_location=None):
    """The SAE branches of utils.py:227-316 (base models are out of scope)."""
    if model_name == "sae_conv":
        return SaeConv(img_size, expansion_factor)
    if model_name == "sae_mlp":
        return SaeMLP(img_size, expansion_factor)
    if model_name == "gated_sae":
        return GatedSae(img_size, expansion_factor)
    raise ValueError(f"Unexpected model name: {model_name} (only the SAE models are part of sparse_vision_b200)")


# --------------------------------------------------------------------------------------------------- layout glue
def _to_bchw(t, b, h, w):
    return t.reshape(b, h, w, t.shape[1]).permute(0, 3, 1, 2)


def reshape_tensor(tensor):
    """utils.py:2770-2774: 'b c h w -> (b h w) c' (a view; the CUDA path never materialises it)."""
    if tensor.dim() == 4:
        return tensor.permute(0, 2, 3, 1).reshape(-1, tensor.shape[1]), True
    return tensor, False


def reshape_encoder_output_average(tensor, batch_size):
    """utils.py:2776-2782 -> [1, N*H*W, C*K] (expanded view instead of the reference's repeat copy)."""
    f, h, w = tensor.shape
    t = tensor.permute(1, 2, 0).reshape(1, h * w, f).expand(batch_size, h * w, f)
    return t.reshape(1, batch_size * h * w, f)


def sae_inference_and_loss(sae_model_name, sae_model, sae_criterion_name, output, sae_criterion, sae_lambda_sparse):
    """utils.py:2448-2482 -> (loss, rec, l1, nrmse, rmse, aux, enc [B,F,H,W], prerelu [B,F,H,W] | None, dec [B,C,H,W]).
    Module-granularity path (differentiable through the autograd Functions of the models); the fused training step
    in model_pipeline.py does not go through here."""
    sae_input, transformed = reshape_tensor(output)
    if sae_model_name == "sae_mlp":
        encoder_output, decoder_output, encoder_output_prerelu = sae_model(output)
    elif sae_model_name == "gated_sae":
        encoder_output, decoder_output, relu_pi_gate, via_gate = sae_model(output)
        encoder_output_prerelu = None
    else:
        raise ValueError(f"Unknown SAE model name {sae_model_name}.")
    if transformed:
        b, _, h, w = output.shape
        encoder_output = _to_bchw(encoder_output, b, h, w)
        if encoder_output_prerelu is not None:
            encoder_output_prerelu = _to_bchw(encoder_output_prerelu, b, h, w)
    if sae_model_name == "sae_mlp" and sae_criterion_name == "sae_loss":
        rec_loss, l1_loss, nrmse_loss, rmse_loss = sae_criterion(encoder_output, decoder_output, sae_input)
        aux_loss = torch.tensor(0)
        loss = rec_loss + sae_lambda_sparse * l1_loss
    elif sae_model_name == "gated_sae" and sae_criterion_name == "gated_sae_loss":
        rec_loss, l1_loss, nrmse_loss, rmse_loss, aux_loss = sae_criterion(relu_pi_gate, via_gate, decoder_output,
                                                                           sae_input)
        loss = rec_loss + sae_lambda_sparse * l1_loss + aux_loss
    else:
        raise ValueError(f"Unknown combination of SAE criterion name {sae_criterion_name} and SAE model name "
                         f"{sae_model_name}.")
    if transformed:
        decoder_output = _to_bchw(decoder_output, b, h, w)
        assert decoder_output.shape == output.shape
    return (loss, rec_loss, l1_loss, nrmse_loss, rmse_loss, aux_loss, encoder_output, encoder_output_prerelu,
            decoder_output)


# --------------------------------------------------------------------------------------------------- activity
def _spatial_mean(t):
    if t.dim() != 4:
        return t
    if not t.is_cuda:
        return torch.mean(t, dim=(2, 3))
    tok = t.permute(0, 2, 3, 1)
    if tok.is_contiguous():            # a [B,F,H,W] VIEW of the token-major tensor the SAE kernels emit: no copy
        return ops.spatial_mean(tok.reshape(-1, t.shape[1]), n_images=t.shape[0])
    return ops.spatial_mean(t)


def average_over_W_H(output, output_2):
    """utils.py:1996-2010: spatial means [B, #units] of conv-shaped activations (one HBM pass on libsvb)."""
    return _spatial_mean(output), (None if output_2 is None else _spatial_mean(output_2))


def variance_explained(output, decoder_output):
    """utils.py:2012-2030 (module-granularity form; the fused step returns it in its stats block)."""
    if output.dim() == 4:
        if decoder_output.dim() != 4:
            raise ValueError(f"Decoder output has unexpected shape {decoder_output.dim()}.")
        var, mod = torch.var(output, dim=(2, 3)).mean(), torch.var(decoder_output, dim=(2, 3)).mean()
    elif output.dim() == 2:
        if decoder_output.dim() != 2:
            raise ValueError(f"Decoder output has unexpected shape {decoder_output.dim()}.")
        var, mod = torch.var(output, dim=1).mean(), torch.var(decoder_output, dim=1).mean()
    else:
        raise ValueError(f"Output has unexpected shape {output.dim()}.")
    return 1 - mod / var


def measure_inactive_units_device(output, expansion_factor):
    """measure_inactive_units with the sparsity left on the device as a 0-dim tensor: no host synchronisation (callers
    that accumulate over batches, CUDA-graph capture)."""
    if output.dim() not in (2, 4):
        raise ValueError(f"Output has unexpected shape {output.dim()}.")
    dead, freq, n_active = ops.measure_inactive(output)
    n_units = output.shape[1]
    return dead.bool(), torch.mean(n_active / (n_units / expansion_factor)), freq


def measure_inactive_units(output, expansion_factor):
    """utils.py:2032-2069 -> (dead units bool [#units], sparsity float, activity frequency [#units])."""
    dead, sparsity, freq = measure_inactive_units_device(output, expansion_factor)
    return dead, sparsity.item(), freq


def get_top_k_samples(top_k_samples, batch_top_k_values, batch_top_k_indices, batch_filename_indices, eval_batch_idx,
                      largest, k):
    """utils.py:1445-1481: merge the running top-k (values, dataset indices, batch_size, filename indices) with one
    batch's top-k."""
    prev_values, prev_indices, batch_size, prev_files = top_k_samples
    batch_top_k_indices += (eval_batch_idx - 1) * batch_size
    if prev_values.is_cuda and prev_values.shape[0] + batch_top_k_values.shape[0] >= k:
        # one kernel: candidates = previous rows followed by the batch's, sorted per unit, payloads gathered along
        if prev_values.shape[0] == 0:
            new_values, new_indices, new_files = ops.topk_columns(
                batch_top_k_values, k, largest, indices=batch_top_k_indices, files=batch_filename_indices)
        else:
            new_values, new_indices, new_files = ops.topk_columns(
                prev_values, k, largest, indices=prev_indices, files=prev_files, values2=batch_top_k_values,
                indices2=batch_top_k_indices, files2=batch_filename_indices)
        return new_values, new_indices, batch_size, new_files
    values = torch.cat((prev_values, batch_top_k_values), dim=0)
    indices = torch.cat((prev_indices, batch_top_k_indices), dim=0)
    files = torch.cat((prev_files, batch_filename_indices), dim=0)
    if values.shape[0] < k:
        return values, indices, batch_size, files
    new_values, pos = torch.topk(values, k=k, dim=0, largest=largest)
    return new_values, torch.gather(indices, 0, pos), batch_size, torch.gather(files, 0, pos)


def batch_top_k(use_output, k):
    """model_pipeline.py:357-360: the k largest and k smallest values of every unit over the batch and the in-batch
    indices of the samples they come from -> (top values, top indices, small values, small indices), each [k, #units].
    (The reference calls torch.topk four times; here two launches of the device top-k.)"""
    top_v, top_i, _ = ops.topk_columns(use_output, k, largest=True)
    small_v, small_i, _ = ops.topk_columns(use_output, k, largest=False)
    return top_v, top_i, small_v, small_i


def update_histogram(histogram_info, name, model_key, output, device, output_2=None):
    """utils.py:1934-1963: add this batch's activations (spatial means [B, #units]; the pre-ReLU encoder output for the
    SAE) of the selected units to their histograms.  histogram_info[(name, model_key)] =
    (histogram_matrix [bins, U], top_values [U], small_values [U], neuron_indices [U])."""
    histogram_matrix, top_values, small_values, neuron_indices = histogram_info[(name, model_key)]
    if output_2 is not None and model_key == "sae":
        activations = output_2
    elif output_2 is None and model_key != "sae":
        activations = output
    else:
        raise ValueError(f"model_key is {model_key} but output_2 is {output_2}")
    idx = torch.as_tensor(neuron_indices, dtype=torch.int64, device=activations.device)
    histogram_matrix = histogram_matrix.to(activations.device)
    ops.histogram_update(histogram_matrix, activations, idx, small_values.to(activations.device),
                         top_values.to(activations.device))
    histogram_info[(name, model_key)] = (histogram_matrix, top_values, small_values, neuron_indices)
    return histogram_info


# --------------------------------------------------------------------------------------------------- indirect effects
def compute_ie_channel_wise(encoder_outputs, encoder_output_average, encoder_gradients, batch_sizes):
    """utils.py:2606-2637: ie[f] = mean over tokens of |grad * (average - activation)|; one HBM pass on libsvb."""
    return ops.ie_channelwise(encoder_outputs, encoder_output_average, encoder_gradients, batch_sizes)


def compute_ie_all_channels(sae_errors, sae_error_average, model_gradients, batch_size):
    """utils.py:2574-2602: mean over tokens of |sum_c grad * (average - error)| -> scalar tensor."""
    return ops.ie_allchannels(sae_errors, sae_error_average, model_gradients, batch_size)


def apply_sae(sae, model_output, nodes=None, ablation=None):
    """utils.py:2786-2820 -> (encoder_output [T,F], decoder_output [B,C,H,W], new_decoder_output [B,C,H,W]); with
    `nodes` (bool [F], True = keep) the other features are replaced by `ablation` [F,H,W] before decoding."""
    b, _, h, w = model_output.shape
    encoder_output, decoder_output, _ = sae(model_output)
    if nodes is not None:
        new_enc = _to_bchw(encoder_output.detach().clone(), b, h, w).clone()
        new_enc[..., ~nodes, :, :] = ablation[~nodes, :, :].to(new_enc.dtype)
        new_tok = new_enc.permute(0, 2, 3, 1).reshape(-1, new_enc.shape[1]).contiguous()
        new_decoder_output = ops.gemm_bf16(new_tok, sae.decoder.weight.detach(), bias=sae.decoder.bias.detach())
    else:
        new_decoder_output = decoder_output
    return encoder_output, _to_bchw(decoder_output, b, h, w), _to_bchw(new_decoder_output, b, h, w)

# --------------------------------------------------------------------------------------------------- a17: trained SAEs
def get_file_path(folder_path=None, sae_layer=None, params=None, file_name=None, params2=None):
    """utils.py:151-185: `<folder>/<layer>[_<params>[_<params2>]]<ending>` where a file name that starts with '.' is an
    extension and anything else is appended with '_' (None included, as in the reference).  Dict parameters are
    joined by '_' in insertion order with None spelled "None"; the folder is created."""
    ending = file_name if (file_name is not None and file_name.startswith(".")) else f"_{file_name}"
    if folder_path is not None:
        os.makedirs(folder_path, exist_ok=True)

    def joined(prm):
        return "_".join("None" if v is None else str(v) for v in prm.values()) if isinstance(prm, dict) else prm

    stem = sae_layer if params is None else (
        f"{sae_layer}_{joined(params)}" if params2 is None else f"{sae_layer}_{joined(params)}_{joined(params2)}")
    name = f"{stem}{ending}"
    return name if folder_path is None else os.path.join(folder_path, name)


# layer suffix -> (dead_neurons_steps, checkpoint epoch, lambda_sparse, expansion factor); utils.py:2671-2724.  The
# learning rate (0.001) and the SAE batch size ('256') are the same for every layer.
_SAE_LAYER_TABLE = {
    "3a": (626, 7, 5.0, 8), "3b": (625, 6, 0.1, 4), "4a": (625, 6, 0.1, 4), "4b": (625, 6, 0.1, 4),
    "4c": (625, 5, 0.1, 4), "4d": (625, 7, 0.1, 4), "4e": (625, 9, 0.1, 4), "5a": (625, 5, 0.1, 4),
    "5b": (625, 12, 0.1, 4),
}


def get_specific_sae_params(layer_name, sae_model_name, model_params_temp, sae_optimizer_name):
    """utils.py:2662-2741: the hyper-parameters a layer's SAE was trained with and the two parameter strings that name
    its checkpoint / MIS files.  Returns (params_string_ie, checkpoint_epoch, expansion_factor, params_string_mis,
    dead_neurons_steps).  Layer names are `mixedXY` or `inceptionXY`."""
    for prefix in ("mixed", "inception"):
        if layer_name.startswith(prefix) and layer_name[len(prefix):] in _SAE_LAYER_TABLE:
            steps, ckpt_epoch, lam, k = _SAE_LAYER_TABLE[layer_name[len(prefix):]]
            break
    else:
        raise UnboundLocalError(f"no SAE parameters are recorded for layer {layer_name}")   # the reference fails the same way
    base = "_".join(model_params_temp.values())
    tail = [str(v) for v in (0.001, "256", sae_optimizer_name, k, lam, steps)]
    ie = f"{base}_{sae_model_name}_" + "_".join(tail) + f"_sae_checkpoint_epoch_{ckpt_epoch}"
    sae_epoch = "11" if layer_name[-2:] == "3a" else "13"       # part of the MIS file name only (:2733-2736)
    mis = f"{base}_{sae_model_name}_{sae_epoch}_" + "_".join(tail) + f"_mis_epoch_{ckpt_epoch}"
    return ie, ckpt_epoch, k, mis, steps


def get_specific_sae_model(layer_name, layer_size, sae_model_name, sae_weights_folder_path, model_params_temp, device,
                           sae_optimizer_name):
    """utils.py:2745-2767: build the layer's SAE, load `<folder>/<layer>_<params>.pth` ('model_state_dict', the
    checkpoint format of model_pipeline.py:1268-1273) and freeze it in eval mode.
    Returns (sae_model, params_string, expansion_factor)."""
    params_string, ckpt_epoch, k, _, _ = get_specific_sae_params(layer_name, sae_model_name, model_params_temp,
                                                                 sae_optimizer_name)
    sae_model = load_model(sae_model_name, img_size=layer_size, expansion_factor=k)
    if ckpt_epoch > 0:
        path = get_file_path(sae_weights_folder_path, layer_name, params=params_string, file_name=".pth")
        sae_model.load_state_dict(torch.load(path, map_location=device)["model_state_dict"])
        print(f"Use SAE on layer {layer_name} from epoch {ckpt_epoch}")
    sae_model = sae_model.to(device).eval()
    for param in sae_model.parameters():
        param.requires_grad = False
    return sae_model, params_string, int(k)
