"""GatedSae — drop-in for the reference's models/gated_sae.py (Gated SAE, Rajamanoharan et al.) on libsvb.

Same constructor, parameter names / order (`W_gate`, `b_gate`, `b_mag`, `r_mag`, `decoder.weight`, `decoder.bias`)
and 4-tuple return as the reference (gated_sae.py:4-56).  The forward is one C-ABI call: a single tcgen05 GEMM
x W_gate^T feeds both the gate and the weight-shared magnitude path (W_mag = exp(r_mag) * W_gate is never
materialised), followed by the decoder GEMM on the gated output and on relu(pi) (the frozen-decoder `via_gate`).
Like the reference has none, there is no reset_encoder_weights here.
"""
import torch
import torch.nn as nn

from .. import ops


class _GatedFunction(torch.autograd.Function):
    """(x, W_gate, b_gate, b_mag, r_mag, W_dec, b_dec) -> (enc, dec, relu_pi, via); via carries no gradient
    (it is computed under no_grad in the reference, gated_sae.py:53-54) and f_gate is detached (:39)."""

    @staticmethod
    def forward(ctx, x, w_gate, b_gate, b_mag, r_mag, w_dec, b_dec):
        enc, dec, rp, via = ops.gated_forward(x, w_gate, b_gate, b_mag, r_mag, w_dec, b_dec)
        ctx.save_for_backward(x, w_gate, b_mag, r_mag, w_dec, b_dec, enc, rp)
        ctx.mark_non_differentiable(via)
        return enc, dec, rp, via

    @staticmethod
    def backward(ctx, g_enc, g_dec, g_rp, _g_via):
        x, w_gate, b_mag, r_mag, w_dec, b_dec, enc, rp = ctx.saved_tensors
        bf = torch.bfloat16
        x_cent = (x - b_dec).to(bf)
        d_enc = g_enc
        gw_dec = gb_dec = None
        if g_dec is not None:
            gd = g_dec.to(bf).contiguous()
            gw_dec = ops.gemm_bf16(gd, enc.to(bf), a_mn=True, b_mn=True)     # [C,F]
            gb_dec = g_dec.sum(0)
            de = ops.gemm_bf16(gd, w_dec.to(bf), b_mn=True)                  # [T,F]
            d_enc = de if d_enc is None else d_enc + de
        d_mag = torch.zeros_like(enc) if d_enc is None else d_enc * (enc > 0)   # gate in {0,1} where enc > 0
        d_pi = torch.zeros_like(enc) if g_rp is None else g_rp * (rp > 0)
        er = torch.exp(r_mag)
        a = (d_pi + er * d_mag).to(bf).contiguous()
        gw_gate = ops.gemm_bf16(a, x_cent, a_mn=True, b_mn=True)             # [F,C]
        gb_gate = d_pi.sum(0)
        gb_mag = d_mag.sum(0)
        gr = (d_mag * (enc - b_mag)).sum(0)                                  # d mag_pre / d r = mag_pre - b_mag
        d_xc = ops.gemm_bf16(a, w_gate.to(bf), b_mn=True)                    # [T,C]
        gb_dec = (gb_dec if gb_dec is not None else 0) - d_xc.sum(0)
        gx = d_xc if ctx.needs_input_grad[0] else None
        return gx, gw_gate, gb_gate, gb_mag, gr, gw_dec, gb_dec


class GatedSae(nn.Module):
    def __init__(self, img_size, expansion_factor):
        super().__init__()
        self.img_size = img_size
        self.act_size = int(torch.prod(torch.tensor(self.img_size)).item())
        self.hidden_size = int(self.act_size * expansion_factor)
        self.W_gate = nn.Parameter(nn.init.kaiming_uniform_(torch.empty(self.hidden_size, self.act_size)))
        self.b_gate = nn.Parameter(torch.zeros(self.hidden_size))
        self.b_mag = nn.Parameter(torch.zeros(self.hidden_size))
        self.r_mag = nn.Parameter(torch.zeros(self.hidden_size))
        self.decoder = nn.Linear(self.hidden_size, self.act_size)
        self.decoder.bias = nn.Parameter(torch.zeros(self.act_size))
        w_dec = nn.init.kaiming_uniform_(torch.empty(self.act_size, self.hidden_size))
        self.decoder.weight = nn.Parameter(w_dec / w_dec.norm(dim=0, keepdim=True))

    def param_list(self):
        return [self.W_gate, self.b_gate, self.b_mag, self.r_mag, self.decoder.weight, self.decoder.bias]

    def forward(self, x):
        """-> (encoder_output, decoder_output, relu_pi_gate, via_gate), 2-D token-major (gated_sae.py:28-56)."""
        if not x.is_cuda:
            raise ValueError("sparse_vision_b200.GatedSae runs on CUDA (B200) tensors only; there is no CPU fallback")
        if x.dim() not in (2, 4):
            raise ValueError(f"Output has unexpected shape {x.dim()}.")
        needs_grad = torch.is_grad_enabled() and (
            x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad:
            x_tok = x.permute(0, 2, 3, 1).reshape(-1, x.shape[1]) if x.dim() == 4 else x
            return _GatedFunction.apply(x_tok.contiguous().float(), *self.param_list())
        return ops.gated_forward(x, *[p.detach() for p in self.param_list()])
