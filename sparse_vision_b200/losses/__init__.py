from .sparse_loss import SparseLoss, GatedSAELoss, compute_rmse_nrmse  # noqa: F401
