"""Loader for the REAL reference (jasper3100/sparse-vision) — usable only where /root/reference exists.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py (and by tests that pin the oracle when the reference is
present) to run the reference's own code; nothing here is imported by the product package, by `-m gpu` tests,
by smoke() or by bench.py (the reference does not exist on the GPU box).

Recipe:
  * models/sae_conv.py, losses/sparse_loss.py, models/custom_mlp.py import directly (torch only);
  * models/sae_mlp.py and models/gated_sae.py do `from utils import *` and need only torch/nn/F/rearrange, so a
    four-name shim module is registered as `utils` while they are imported;
  * utils.py itself cannot be imported (h5py, lucent, webdataset, nnsight, ... are absent); the functions on the
    hot path are extracted from its AST and executed in a namespace seeded with what they use.
No reference source is copied into this repository: the code is read from /root/reference at run time.
"""
import ast
import importlib
import logging
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F
from einops import rearrange

REFERENCE_ROOT = os.environ.get("SVB_REFERENCE_ROOT", "/root/reference")

_UTILS_NAMES = (
    "ConstrainedAdam", "get_optimizer", "get_criterion", "CustomCrossEntropyLoss", "average_over_W_H",
    "variance_explained", "measure_inactive_units", "sae_inference_and_loss", "compute_ie_all_channels",
    "compute_ie_channel_wise", "reshape_tensor", "reshape_encoder_output_average", "apply_sae",
    "get_top_k_samples",
)


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "sae_mlp.py"))


_cache = None


def load():
    """Returns a namespace with the reference's SaeMLP, GatedSae, SaeConv, SparseLoss, GatedSAELoss,
    compute_rmse_nrmse, CustomMLP9 and the hot-path functions of utils.py."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    saved_path = list(sys.path)
    saved_utils = sys.modules.get("utils")
    saved_pkgs = {k: sys.modules.get(k) for k in ("models", "losses")}
    try:
        sys.path.insert(0, REFERENCE_ROOT)
        for k in list(sys.modules):
            if k == "models" or k.startswith("models.") or k == "losses" or k.startswith("losses."):
                del sys.modules[k]
        shim = types.ModuleType("utils")
        shim.torch, shim.nn, shim.F, shim.rearrange = torch, nn, F, rearrange
        shim.__all__ = ["torch", "nn", "F", "rearrange"]
        sys.modules["utils"] = shim
        sae_mlp = importlib.import_module("models.sae_mlp")
        gated = importlib.import_module("models.gated_sae")
        sae_conv = importlib.import_module("models.sae_conv")
        custom_mlp = importlib.import_module("models.custom_mlp")
        sparse_loss = importlib.import_module("losses.sparse_loss")

        ns = {
            "torch": torch, "nn": nn, "F": F, "rearrange": rearrange, "logging": logging,
            "SparseLoss": sparse_loss.SparseLoss, "GatedSAELoss": sparse_loss.GatedSAELoss,
            "SaeMLP": sae_mlp.SaeMLP, "GatedSae": gated.GatedSae, "SaeConv": sae_conv.SaeConv,
        }
        with open(os.path.join(REFERENCE_ROOT, "utils.py")) as fh:
            tree = ast.parse(fh.read())
        wanted = [n for n in tree.body
                  if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in _UTILS_NAMES]
        mod = ast.Module(body=wanted, type_ignores=[])
        exec(compile(mod, os.path.join(REFERENCE_ROOT, "utils.py"), "exec"), ns)

        out = types.SimpleNamespace(
            SaeMLP=sae_mlp.SaeMLP, GatedSae=gated.GatedSae, SaeConv=sae_conv.SaeConv,
            SparseLoss=sparse_loss.SparseLoss, GatedSAELoss=sparse_loss.GatedSAELoss,
            compute_rmse_nrmse=sparse_loss.compute_rmse_nrmse, CustomMLP9=custom_mlp.CustomMLP9,
            **{k: ns[k] for k in _UTILS_NAMES},
        )
        _cache = out
        return out
    finally:
        sys.path[:] = saved_path
        if saved_utils is not None:
            sys.modules["utils"] = saved_utils
        else:
            sys.modules.pop("utils", None)
        for k in list(sys.modules):
            if k == "models" or k.startswith("models.") or k == "losses" or k.startswith("losses."):
                del sys.modules[k]
        for k, v in saved_pkgs.items():
            if v is not None:
                sys.modules[k] = v
