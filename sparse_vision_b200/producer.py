"""The frozen base model that PRODUCES the activations the SAE path consumes.

The reference loads torchvision's GoogLeNet with ImageNet weights (`utils.py:277-281`:
`torchvision.models.googlenet(pretrained=True, aux_logits=True)`) and hooks its inception blocks, which it calls
`mixed3a` ... `mixed5b` (`compute_ie.py:52`, `utils.py:2671-2724`).  There is no network here, so the weights are a
seeded random initialisation; to keep every hooked layer at the O(1), post-ReLU, roughly half-sparse statistics of
a trained network (instead of the 1e-4 ... 1e-10 magnitudes an uncalibrated random GoogLeNet has ten layers deep)
the BatchNorm running statistics are calibrated once on a seeded batch.

With `channels_last=True` (and bf16) cuDNN emits NHWC activations, i.e. the hooked `[B,C,H,W]` tensor already IS the
`[(b h w), C]` token matrix the SAE kernels read: no layout copy anywhere between the convolution and the encoder GEMM
(what `models/sae_mlp.py:44` / `utils.py:2770-2774` pay a permute copy for).
"""
import torch

# reference layer name -> torchvision module name, and the (C, H*W) each one emits for 224x224 / 229x229 inputs
GOOGLENET_LAYERS = {
    "mixed3a": ("inception3a", 256, 784), "mixed3b": ("inception3b", 480, 784),
    "mixed4a": ("inception4a", 512, 196), "mixed4b": ("inception4b", 512, 196), "mixed4c": ("inception4c", 512, 196),
    "mixed4d": ("inception4d", 528, 196), "mixed4e": ("inception4e", 832, 196),
    "mixed5a": ("inception5a", 832, 49), "mixed5b": ("inception5b", 1024, 49),
}


def module_name(layer_name):
    """`mixedXY` (the reference's name) or `inceptionXY` (torchvision's) -> torchvision module name."""
    if layer_name in GOOGLENET_LAYERS:
        return GOOGLENET_LAYERS[layer_name][0]
    return layer_name


def synthetic_googlenet(seed=0, calibration_images=8, image_size=224):
    """GoogLeNet as the reference builds it (`aux_logits=True`), seeded random weights, calibrated BatchNorm statistics,
    eval mode, frozen.  Deterministic on the CPU generator: the same seed gives bit-identical weights everywhere."""
    import torchvision
    gen_state = torch.get_rng_state()
    try:
        torch.manual_seed(seed)
        model = torchvision.models.googlenet(weights=None, aux_logits=True, init_weights=True)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.momentum = None            # cumulative average: one pass sets the running statistics to the batch's
                m.reset_running_stats()
        model.train()
        with torch.no_grad():
            model(torch.randn(calibration_images, 3, image_size, image_size))
        model.eval()
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.momentum = 0.1
    finally:
        torch.set_rng_state(gen_state)
    for p in model.parameters():
        p.requires_grad = False
    return model


def fold_batchnorm(model):
    """Folds every eval-mode BatchNorm into the convolution in front of it (exact algebra for a frozen model): one cuDNN
    kernel and one pass over the activations per layer instead of three.  In place; returns the model."""
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    for m in model.modules():
        conv, bn = getattr(m, "conv", None), getattr(m, "bn", None)
        if isinstance(conv, torch.nn.Conv2d) and isinstance(bn, torch.nn.BatchNorm2d):
            m.conv = fuse_conv_bn_eval(conv.eval(), bn.eval())
            m.bn = torch.nn.Identity()
    for p in model.parameters():
        p.requires_grad = False
    return model


def to_producer_format(model, device, dtype=torch.bfloat16, channels_last=True, fold_bn=False):
    """Moves the frozen base model to `device` in the format the SAE kernels read without a copy."""
    if fold_bn:
        model = fold_batchnorm(model.float())
    model = model.to(device=device, dtype=dtype)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    return model


def hooked_layers(model, names):
    """{reference layer name: module} for IE(...) / ModelPipeline(...)."""
    mods = dict(model.named_modules())
    return {n: mods[module_name(n)] for n in names}
