"""SaeMLP — drop-in for the reference's models/sae_mlp.py, running on libsvb (tcgen05 GEMMs, fused epilogues).

Same constructor, parameter names / order (`encoder.weight [F,C]`, `encoder.bias`, `decoder.weight [C,F]`,
`decoder.bias`), state_dict keys, return tuples and error behaviour as the reference class (sae_mlp.py:4-199), so
reference checkpoints load unchanged.  The forward is one C-ABI call (svb_sae_forward); its autograd backward uses
the same tcgen05 GEMM kernel (svb_gemm_bf16).  The training hot path does NOT go through autograd: the trainer in
model_pipeline.py calls the fused step (svb_sae_train_step).  CPU tensors are rejected — there is no fallback.
"""
import math

import torch
import torch.nn as nn

from .. import ops


def _kaiming_uniform(rows, cols, device=None):
    # same draw as torch.nn.init.kaiming_uniform_(torch.empty(rows, cols)): U(-b, b), b = sqrt(6 / fan_in)
    return nn.init.kaiming_uniform_(torch.empty(rows, cols, device=device))


class _SaeMlpFunction(torch.autograd.Function):
    """(x, W_enc, b_enc, W_dec, b_dec) -> (enc [T,F], dec [T,C], pre [T,F]); sae_mlp.py:49-52."""

    @staticmethod
    def forward(ctx, x, w_enc, b_enc, w_dec, b_dec):
        enc, dec, pre = ops.sae_forward(x, w_enc, b_enc, w_dec, b_dec)
        ctx.save_for_backward(x, w_enc, w_dec, b_dec, enc)
        return enc, dec, pre

    @staticmethod
    def backward(ctx, g_enc, g_dec, g_pre):
        x, w_enc, w_dec, b_dec, enc = ctx.saved_tensors
        bf = torch.bfloat16
        T, F = enc.shape
        C = w_enc.shape[1]
        x_cent = (x - b_dec).to(bf)
        enc_b = enc.to(bf)
        d_enc = g_enc if g_enc is not None else None
        gw_dec = gb_dec = None
        if g_dec is not None:
            gd = g_dec.to(bf).contiguous()
            gw_dec = ops.gemm_bf16(gd, enc_b, a_mn=True, b_mn=True)          # [C,F] = g_dec^T enc
            gb_dec = g_dec.sum(0)
            de = ops.gemm_bf16(gd, w_dec.to(bf), b_mn=True)                  # [T,F] = g_dec W_dec
            d_enc = de if d_enc is None else d_enc + de
        d_pre = torch.zeros_like(enc) if d_enc is None else d_enc * (enc > 0)
        if g_pre is not None:
            d_pre = d_pre + g_pre
        dp = d_pre.to(bf).contiguous()
        gw_enc = ops.gemm_bf16(dp, x_cent, a_mn=True, b_mn=True)             # [F,C] = d_pre^T x_cent
        gb_enc = d_pre.sum(0)
        d_xc = ops.gemm_bf16(dp, w_enc.to(bf), b_mn=True)                    # [T,C] = d_pre W_enc
        gb_dec = (gb_dec if gb_dec is not None else 0) - d_xc.sum(0)         # b_dec also feeds x - b_dec
        gx = d_xc if ctx.needs_input_grad[0] else None
        return gx, gw_enc, gb_enc, gw_dec, gb_dec


def draw_reinit(w_enc, b_enc, w_dec, dead, draw_device=None):
    """The random part of reset_encoder_weights (sae_mlp.py:106-130): full Kaiming draws (W_enc-shaped first, then
    W_dec-shaped, like the reference), rows rescaled to the mean row norm of the live units, and the replacement
    encoder bias.  The draws use the generator of `draw_device` (default: the parameters' device, which is what the
    reference's torch.nn.init call does; tests pass "cpu" to reproduce fixtures made on the CPU generator).
    Returns (new_W_enc [F,C], new_W_dec [C,F], new_b_enc float)."""
    dev = w_enc.device
    dd = dev if draw_device is None else torch.device(draw_device)
    new_w_enc = nn.init.kaiming_uniform_(torch.zeros(w_enc.shape, device=dd)).to(dev)
    new_w_dec = nn.init.kaiming_uniform_(torch.zeros(w_dec.shape, device=dd)).to(dev)
    live = torch.nonzero(~dead.bool()).squeeze(-1)
    avg_enc = torch.norm(w_enc[live, :], p=2, dim=1).mean().item()            # :113-115
    avg_dec = torch.norm(w_dec[:, live], p=2, dim=1).mean().item()            # :118-119 (dim=1, as in the reference)
    new_b = b_enc[live].abs().mean().item()                                   # :121
    new_w_enc = (new_w_enc / torch.norm(new_w_enc, p=2, dim=1, keepdim=True) * avg_enc).contiguous()
    new_w_dec = (new_w_dec / torch.norm(new_w_dec, p=2, dim=1, keepdim=True) * avg_dec).contiguous()
    return new_w_enc, new_w_dec, new_b


class SaeMLP(nn.Module):
    def __init__(self, img_size, expansion_factor):
        """img_size: int or tuple (act_size = prod); expansion_factor: hidden_size = act_size * k (sae_mlp.py:4-40)."""
        super().__init__()
        self.img_size = img_size
        self.act_size = int(torch.prod(torch.tensor(self.img_size)).item())
        self.hidden_size = int(self.act_size * expansion_factor)
        # nn.Linear first (it draws from the RNG, as in the reference), then the explicit Kaiming re-draws
        self.encoder = nn.Linear(self.act_size, self.hidden_size)
        self.encoder.weight = nn.Parameter(_kaiming_uniform(self.hidden_size, self.act_size))
        self.encoder.bias = nn.Parameter(torch.zeros(self.hidden_size))
        self.sae_act = nn.ReLU()
        self.decoder = nn.Linear(self.hidden_size, self.act_size)
        self.decoder.bias = nn.Parameter(torch.zeros(self.act_size))
        w_dec = _kaiming_uniform(self.act_size, self.hidden_size)
        self.decoder.weight = nn.Parameter(w_dec / w_dec.norm(dim=0, keepdim=True))   # unit-norm columns

    def forward(self, x):
        """[B,C,H,W] (pixels become tokens, order (b h w)) or [N,C] -> (encoder_output, decoder_output,
        encoder_output_prerelu), all 2-D token-major like the reference (sae_mlp.py:42-53)."""
        if not x.is_cuda:
            raise ValueError("sparse_vision_b200.SaeMLP runs on CUDA (B200) tensors only; there is no CPU fallback")
        if x.dim() not in (2, 4):
            raise ValueError(f"Output has unexpected shape {x.dim()}.")
        x_tok = x.permute(0, 2, 3, 1).reshape(-1, x.shape[1]) if x.dim() == 4 else x
        needs_grad = torch.is_grad_enabled() and (
            x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad:
            return _SaeMlpFunction.apply(x_tok.contiguous().float(), self.encoder.weight, self.encoder.bias,
                                         self.decoder.weight, self.decoder.bias)
        return ops.sae_forward(x, self.encoder.weight.detach(), self.encoder.bias.detach(),
                               self.decoder.weight.detach(), self.decoder.bias.detach())

    # parameters in the order the fused step and the optimizer expect (== nn.Module.parameters() order)
    def param_list(self):
        return [self.encoder.weight, self.encoder.bias, self.decoder.weight, self.decoder.bias]

    def reset_encoder_weights(self, dead_neurons_sae, device, optimizer, epoch, train_batch_idx, epoch_batch_idx,
                              file_path):
        """Re-initialise the units flagged in the bool mask `dead_neurons_sae` and reset their Adam moments
        (sae_mlp.py:79-184).  Same messages and ValueErrors as the reference."""
        idx = torch.nonzero(dead_neurons_sae)
        where = f"Epoch {epoch}, train batch index {train_batch_idx}, epoch batch index {epoch_batch_idx}"
        if idx.dim() != 2:
            raise ValueError(f"{where}: The indices_of_dead_neurons tensor has unexpected shape.")
        if idx.shape[0] == 0 or idx.shape[1] == 0:
            print(f"{where}: No dead neurons in the SAE --> no re-initialization necessary")
            return
        if idx.shape[1] != 1:
            raise ValueError(f"{where}: The indices_of_dead_neurons tensor has unexpected value in second dimension.")
        idx = idx.squeeze(-1)
        if file_path is not None:
            with open(file_path, "w") as fh:
                fh.write("".join(f"{int(i)}\n" for i in idx.tolist()))
        name = optimizer.__class__.__name__
        if name not in ("Adam", "ConstrainedAdam"):
            raise ValueError(f"The optimizer {name} is not supported for re-initializing dead neurons.")
        w_enc, b_enc, w_dec = self.encoder.weight.data, self.encoder.bias.data, self.decoder.weight.data
        dead = dead_neurons_sae.to(w_enc.device)
        new_w_enc, new_w_dec, new_b = draw_reinit(w_enc, b_enc, w_dec, dead)
        ms, vs = [], []
        for p in self.param_list():
            st = optimizer.state.get(p, {})
            if "exp_avg" not in st:   # optimizer has not stepped yet: nothing to reset
                ms = vs = None
                break
            ms.append(st["exp_avg"])
            vs.append(st["exp_avg_sq"])
        ops.reinit_dead([p.data for p in self.param_list()], ms, vs, dead.to(torch.uint8), new_w_enc, new_w_dec, new_b)
        print(f"{where}: Re-initialized {len(idx)} dead neurons in the SAE and reset optimizer parameters.")

    def intervene_on_decoder_weights(self, unit_index, value):
        """sae_mlp.py:187-199: overwrite one decoder column (one SAE feature direction)."""
        self.decoder.weight.data[:, unit_index] = value
