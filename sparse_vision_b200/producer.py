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


def to_producer_format(model, device, dtype=torch.bfloat16, channels_last=True, fold_bn=False, fuse=False):
    """Moves the frozen base model to `device` in the format the SAE kernels read without a copy.  fuse=True (with
    fold_bn, bf16 and channels_last) also swaps in the fused forward below."""
    if fold_bn:
        model = fold_batchnorm(model.float())
    model = model.to(device=device, dtype=dtype)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    if fuse:
        model = fuse_forward(model)
    return model


def hooked_layers(model, names):
    """{reference layer name: module} for IE(...) / ModelPipeline(...)."""
    mods = dict(model.named_modules())
    return {n: mods[module_name(n)] for n in names}


# ------------------------------------------------------------------------------------------------ fused forward
# torchvision's eager forward of the frozen network, bf16 / channels_last / BatchNorm folded, 256 images, spends 4.5 of
# its 10.6 ms in ATen's NHWC max-pool, 2.4 ms in the bias add_ and relu_ behind every cuDNN convolution and 0.65 ms in
# the torch.cat of the nine inception blocks -- kernels that only move bytes (tools/prof_producer.py).  fuse_forward()
# swaps the classes of those modules for the ones below: the module tree, the names the hooks are registered under and
# the state_dict stay what they were; the convolutions still go to cuDNN; everything between them is two kernels of
# libsvb (svb_maxpool_nhwc, svb_bias_relu_scatter).  Inputs that are not bf16 channels_last CUDA tensors take
# torchvision's own forward.
def _fast_input(x):
    return (x.is_cuda and x.dim() == 4 and x.dtype == torch.bfloat16 and not x.requires_grad
            and x.is_contiguous(memory_format=torch.channels_last))


def _conv_nobias(x, conv):
    return torch.nn.functional.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)


def _make_fused_classes():
    from torchvision.models.googlenet import BasicConv2d, Inception
    from . import ops

    class FusedBasicConv2d(BasicConv2d):
        """conv (cuDNN, no bias) -> relu(. + folded BatchNorm bias) in place, one pass instead of two."""

        def _is_stem(self, x):
            c = self.conv
            return (tuple(x.shape[1:]) == (3, 224, 224) and tuple(c.weight.shape) == (64, 3, 7, 7) and c.stride == (2, 2)
                    and c.padding == (3, 3) and c.dilation == (1, 1) and c.groups == 1)

        def forward(self, x):
            if not (_fast_input(x) and isinstance(self.bn, torch.nn.Identity) and self.conv.bias is not None):
                return super().forward(x)
            if self._is_stem(x):          # conv1: libsvb's own implicit-GEMM kernel, bias + relu in its epilogue
                w = self.conv.weight
                key = (w.data_ptr(), w._version, w.device)
                if getattr(self, "_svb_key", None) != key:
                    self._svb_packed, self._svb_key = ops.conv1_pack_weights(w), key
                return ops.conv1_stem(x, self._svb_packed, self.conv.bias)
            y = _conv_nobias(x, self.conv)
            if getattr(self, "_svb_defer_bias", False) and not self._forward_hooks:     # a hook must see the finished output
                # the next module is a max-pool: relu(max(y) + b) == max(relu(y + b)) exactly (a per-channel constant,
                # a monotone rounding and a monotone relu commute with max), so the bias pass runs on the POOLED tensor
                # (a quarter of the bytes).  The raw convolution output is tagged; FusedMaxPool2d finishes it.
                y._svb_pending_bias = self.conv.bias
                return y
            ops.bias_relu_scatter(y, self.conv.bias, [(y, 0, y.shape[1])])
            return y

    class FusedMaxPool2d(torch.nn.MaxPool2d):
        def forward(self, x):
            k, s, p = self.kernel_size, self.stride, self.padding
            geom_ok = (isinstance(k, int) and isinstance(s, int) and isinstance(p, int) and self.dilation == 1
                       and not self.return_indices and (k, s) in ((3, 1), (3, 2), (2, 2)))
            pending = getattr(x, "_svb_pending_bias", None)
            if geom_ok and pending is None and x.requires_grad and torch.is_grad_enabled() and _fast_input(x.detach()):
                # the IE passes differentiate through the pools behind the first hooked layer: libsvb's forward with
                # indices + gather backward instead of ATen's (2.5 of 5.5 ms of a 64-image forward + backward)
                return ops.maxpool_nhwc_autograd(x, k, s, p, self.ceil_mode)
            if not (_fast_input(x) and geom_ok):
                if pending is not None:          # cannot pool this one here: finish the deferred bias + relu first
                    x = x.add(pending.view(1, -1, 1, 1)).relu_()
                return super().forward(x)
            y = ops.maxpool_nhwc(x, k, s, p, self.ceil_mode)
            if pending is not None:
                ops.bias_relu_scatter(y, pending, [(y, 0, y.shape[1])])
            return y

    def _conv_grad_input(grad, like_input, conv_weight, padding):
        return torch.ops.aten.convolution_backward(grad, like_input, conv_weight, None, [1, 1], list(padding), [1, 1], False,
                                                   [0, 0], 1, [True, False, False])[0]

    class _InceptionFn(torch.autograd.Function):
        """The fused inception forward for a tensor that requires grad (the IE passes behind the first hooked layer),
        with its backward written out: per branch the ReLU-mask gather of the block gradient's channel range
        (svb_relu_grad_gather), cuDNN's data gradient of the convolution, libsvb's max-pool backward for branch 4; the
        three 1x1 convolutions on the block input are ONE data-gradient GEMM over the merged channels."""

        @staticmethod
        def forward(ctx, x, block):
            out, t3, t5, arg = block._fused(x, with_argmax=True)
            ctx.block = block
            ctx.save_for_backward(x, out, t3, t5, arg)
            return out

        @staticmethod
        def backward(ctx, grad_out):
            x, out, t3, t5, arg = ctx.saved_tensors
            blk = ctx.block
            go = grad_out.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            c1, c3r, c3, c5r, c5, cp = blk._widths()
            w1, _ = blk._merged()
            g = ops.relu_grad_gather([(go, c1 + c3 + c5, out, c1 + c3 + c5, cp)], x)              # pool projection
            g = _conv_grad_input(g, x, blk.branch4[1].conv.weight, (0, 0))
            pool = blk.branch4[0]
            gx = ops.maxpool_nhwc_backward(g.contiguous(memory_format=torch.channels_last), arg, x.shape, pool.kernel_size,
                                           pool.stride, pool.padding)
            g3 = ops.relu_grad_gather([(go, c1, out, c1, c3)], x)
            g3 = _conv_grad_input(g3, t3, blk.branch2[1].conv.weight, blk.branch2[1].conv.padding)
            g5 = ops.relu_grad_gather([(go, c1 + c3, out, c1 + c3, c5)], x)
            g5 = _conv_grad_input(g5, t5, blk.branch3[1].conv.weight, blk.branch3[1].conv.padding)
            gm = ops.relu_grad_gather([(go, 0, out, 0, c1),
                                       (g3.contiguous(memory_format=torch.channels_last), 0, t3, 0, c3r),
                                       (g5.contiguous(memory_format=torch.channels_last), 0, t5, 0, c5r)], x)
            gx += _conv_grad_input(gm, x, w1, (0, 0))
            return gx, None

    class FusedInception(Inception):
        """The three 1x1 convolutions that read the block's input run as ONE convolution (weights concatenated once);
        every branch's relu(conv + bias) is written straight into its channel range of the block output."""

        def _merged(self):
            convs = [self.branch1.conv, self.branch2[0].conv, self.branch3[0].conv]
            # rebuilt when a weight was re-assigned, moved or written in place (load_state_dict after fuse_forward)
            key = tuple((c.weight.data_ptr(), c.weight._version, c.bias.data_ptr(), c.bias._version) for c in convs)
            if getattr(self, "_svb_key", None) != key:
                self._svb_w1 = torch.cat([c.weight for c in convs]).contiguous(memory_format=torch.channels_last)
                self._svb_b1 = torch.cat([c.bias for c in convs]).contiguous()
                self._svb_key = key
            return self._svb_w1, self._svb_b1

        def _convs(self):
            return [self.branch1, self.branch2[0], self.branch2[1], self.branch3[0], self.branch3[1], self.branch4[1]]

        def _widths(self):
            return tuple(m.conv.out_channels for m in self._convs())

        def _fused(self, x, with_argmax=False):
            b, _, h, w = x.shape
            c1, c3r, c3, c5r, c5, cp = self._widths()
            new = lambda c: torch.empty((b, c, h, w), device=x.device, dtype=x.dtype,   # noqa: E731
                                        memory_format=torch.channels_last)
            out, t3, t5 = new(c1 + c3 + c5 + cp), new(c3r), new(c5r)
            w1, b1 = self._merged()
            y = torch.nn.functional.conv2d(x, w1, None)
            ops.bias_relu_scatter(y, b1, [(out, 0, c1), (t3, 0, c3r), (t5, 0, c5r)])
            y = _conv_nobias(t3, self.branch2[1].conv)
            ops.bias_relu_scatter(y, self.branch2[1].conv.bias, [(out, c1, c3)])
            y = _conv_nobias(t5, self.branch3[1].conv)
            ops.bias_relu_scatter(y, self.branch3[1].conv.bias, [(out, c1 + c3, c5)])
            arg = None
            if with_argmax:
                pool = self.branch4[0]
                pooled, arg = ops.maxpool_nhwc_with_argmax(x, pool.kernel_size, pool.stride, pool.padding, pool.ceil_mode)
            else:
                pooled = self.branch4[0](x)
            y = _conv_nobias(pooled, self.branch4[1].conv)
            ops.bias_relu_scatter(y, self.branch4[1].conv.bias, [(out, c1 + c3 + c5, cp)])
            return out, t3, t5, arg

        def forward(self, x):
            folded = all(isinstance(m.bn, torch.nn.Identity) and m.conv.bias is not None for m in self._convs())
            if folded and _fast_input(x):
                return self._fused(x)[0]
            pool = self.branch4[0]
            if (folded and x.requires_grad and torch.is_grad_enabled() and _fast_input(x.detach())
                    and not any(p.requires_grad for p in self.parameters())
                    and isinstance(pool.kernel_size, int) and (pool.kernel_size, pool.stride) == (3, 1)):
                return _InceptionFn.apply(x, self)          # frozen block, gradient with respect to its input only
            return super().forward(x)

    return BasicConv2d, Inception, FusedBasicConv2d, FusedMaxPool2d, FusedInception


def fuse_forward(model):
    """In place: GoogLeNet's BasicConv2d / MaxPool2d / Inception modules get the fused forwards above (class swap; the
    parameters, module names, hooks and state_dict are untouched).  Needs a BatchNorm-folded model to take effect."""
    BasicConv2d, Inception, FusedBasicConv2d, FusedMaxPool2d, FusedInception = _make_fused_classes()
    # GoogLeNet._forward: conv2 -> conv3 -> maxpool2; conv3's bias + relu pass moves behind the pool (see FusedBasicConv2d)
    kids = list(model.named_children())      # (a ModuleList of the leading modules works too: to_attribution_format)
    for (_, a), (_, b) in zip(kids, kids[1:]):
        if type(a) is BasicConv2d and type(b) is torch.nn.MaxPool2d and not a._forward_hooks:
            a._svb_defer_bias = True
    for m in model.modules():
        if type(m) is BasicConv2d:
            m.__class__ = FusedBasicConv2d
        elif type(m) is torch.nn.MaxPool2d:
            m.__class__ = FusedMaxPool2d
        elif type(m) is Inception:
            m.__class__ = FusedInception
    return model


def to_attribution_format(model, device, first_layer=None, dtype=torch.bfloat16, fold_bn=True, nchw_tail=False):
    """The frozen base model for the IE passes (compute_ie.py:365-472), which need its BACKWARD from the loss down to the
    first hooked layer (IE._forward_collect makes that layer's output the autograd leaf): bf16, channels_last, fused
    forward.  The layers in front of the leaf run forward-only on the producer kernels; behind it the convolutions go
    through torch autograd / cuDNN and the max-pools through libsvb's differentiable pair.  nchw_tail=True keeps
    everything behind `first_layer` in NCHW on torchvision's own forward (the round's earlier arrangement: with the
    backward running through the stem to the images, cuDNN's channels_last backward was 4x slower than NCHW; with the
    leaf at the first hooked layer channels_last is the faster one, 6.1 against 7.3 ms per 64 images).
    Feed it channels_last images."""
    if fold_bn:
        model = fold_batchnorm(model.float())
    model = model.to(device=device, dtype=dtype)
    fuse = fold_bn and dtype == torch.bfloat16
    if not nchw_tail:
        model = model.to(memory_format=torch.channels_last)
        return fuse_forward(model) if fuse else model
    first = module_name(first_layer)
    names = [n for n, _ in model.named_children()]
    if first not in names:
        raise ValueError(f"{first_layer} is not a top-level module of the model")
    head = torch.nn.ModuleList([m for n, m in model.named_children()][:names.index(first) + 1])
    head.to(memory_format=torch.channels_last)
    if fuse:
        fuse_forward(head)
    return model
