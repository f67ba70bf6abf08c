"""SparseLoss / GatedSAELoss — module-level drop-ins for the reference's losses/sparse_loss.py.

These are the differentiable, module-granularity forms (tiny reductions on tensors the caller already holds); the
training hot path never calls them — the fused step computes the same quantities inside the GEMM epilogues
(csrc/epilogues.cuh: sum (d-x)^2, sum |enc|, per-channel stats) and returns them in the stats block.
"""
import torch
import torch.nn as nn


def compute_rmse_nrmse(decoded, targets):
    """sparse_loss.py:4-21: per-channel RMSE over the batch axis; NRMSE divides by the channel's target range."""
    mse_c = torch.mean(torch.square(decoded - targets), dim=0)
    rng_c = torch.max(targets, dim=0)[0] - torch.min(targets, dim=0)[0]
    rmse_c = torch.sqrt(mse_c)
    return torch.mean(rmse_c), torch.mean(rmse_c / rng_c)


class SparseLoss(nn.Module):
    def forward(self, encoded, decoded, targets):
        """-> (reconstruction mse, mean |encoded|, nrmse, rmse)   (sparse_loss.py:30-61)."""
        assert decoded.shape == targets.shape
        assert len(decoded.shape) == 2
        rec = nn.functional.mse_loss(decoded, targets)
        l1 = torch.mean(torch.abs(encoded))
        rmse, nrmse = compute_rmse_nrmse(decoded, targets)
        return rec, l1, nrmse, rmse


class GatedSAELoss(nn.Module):
    def forward(self, relu_pi_gate, via_gate, decoded, targets):
        """-> (reconstruction mse, mean |relu_pi_gate|, nrmse, rmse, aux mse)   (sparse_loss.py:68-76)."""
        rec = nn.functional.mse_loss(decoded, targets)
        l1 = torch.mean(torch.abs(relu_pi_gate))
        aux = nn.functional.mse_loss(via_gate, targets)
        rmse, nrmse = compute_rmse_nrmse(decoded, targets)
        return rec, l1, nrmse, rmse, aux
