"""Thin functional layer over the C ABI: allocates outputs with torch, passes raw pointers, raises on error.
These are what the reference-shaped modules in models/, losses/, utils.py, model_pipeline.py and compute_ie.py call.
"""
import ctypes as C

import torch

from . import _lib as L


def _f32c(t):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise ValueError("parameters / optimizer state must be float32 CUDA tensors")
    if not t.is_contiguous():
        raise ValueError("parameters / optimizer state must be contiguous (they are updated in place)")
    return t


def _sae_params(w_enc, b_enc, w_dec, b_dec):
    F, Cc = w_enc.shape
    if tuple(w_dec.shape) != (Cc, F) or b_enc.numel() != F or b_dec.numel() != Cc:
        raise ValueError("inconsistent SAE parameter shapes")
    return L.SaeParams(L.ptr(_f32c(w_enc)), L.ptr(_f32c(b_enc)), L.ptr(_f32c(w_dec)), L.ptr(_f32c(b_dec)), F)


def _gated_params(w_gate, b_gate, b_mag, r_mag, w_dec, b_dec):
    F, Cc = w_gate.shape
    if tuple(w_dec.shape) != (Cc, F):
        raise ValueError("inconsistent Gated-SAE parameter shapes")
    return L.GatedParams(L.ptr(_f32c(w_gate)), L.ptr(_f32c(b_gate)), L.ptr(_f32c(b_mag)), L.ptr(_f32c(r_mag)),
                         L.ptr(_f32c(w_dec)), L.ptr(_f32c(b_dec)), F)


def _adam_state(ms, vs):
    st = L.AdamState()
    for i, (m, v) in enumerate(zip(ms, vs)):
        st.m[i] = _f32c(m).data_ptr()
        st.v[i] = _f32c(v).data_ptr()
    return st


def _opt(optimizer, step, lr, betas, eps, step_dev=None):
    code = {"adam": L.SVB_ADAM, "constrained_adam": L.SVB_CONSTRAINED_ADAM}[optimizer]
    if step_dev is not None and (step_dev.dtype != torch.int32 or not step_dev.is_cuda or step_dev.numel() != 1):
        raise ValueError("step_dev must be a one-element int32 CUDA tensor")
    return L.OptConfig(code, int(step), float(lr), float(betas[0]), float(betas[1]), float(eps), L.ptr(step_dev))


def _tokens(x):
    return x.shape[0] * (x.shape[2] * x.shape[3] if x.dim() == 4 else 1)


# --------------------------------------------------------------------------------------------------- forward
def sae_forward(x, w_enc, b_enc, w_dec, b_dec, want_pre=True, want_dec=True, out_dtype=torch.float32):
    """models/sae_mlp.py:42-53 -> (encoder_output [T,F], decoder_output [T,C], prerelu [T,F] | None), token-major."""
    a, x = L.acts_of(x)
    p = _sae_params(w_enc, b_enc, w_dec, b_dec)
    T, F, Cc = _tokens(x), p.F, a.C
    enc = torch.empty((T, F), device=x.device, dtype=out_dtype)
    pre = torch.empty((T, F), device=x.device, dtype=torch.float32) if want_pre else None
    dec = torch.empty((T, Cc), device=x.device, dtype=out_dtype) if want_dec else None
    out = L.SaeForwardOut(L.ptr(enc), L.dtype_code(enc), L.ptr(pre), L.ptr(dec), L.dtype_code(dec) if want_dec else 0)
    L.check(L.load().svb_sae_forward(L.handle(x.device), L.stream_ptr(x.device), C.byref(a), C.byref(p), C.byref(out)),
            "svb_sae_forward")
    return enc, dec, pre


def gated_forward(x, w_gate, b_gate, b_mag, r_mag, w_dec, b_dec, out_dtype=torch.float32):
    """models/gated_sae.py:28-56 -> (encoder_output, decoder_output, relu_pi_gate, via_gate), token-major."""
    a, x = L.acts_of(x)
    p = _gated_params(w_gate, b_gate, b_mag, r_mag, w_dec, b_dec)
    T, F, Cc = _tokens(x), p.F, a.C
    enc = torch.empty((T, F), device=x.device, dtype=out_dtype)
    rp = torch.empty((T, F), device=x.device, dtype=out_dtype)
    dec = torch.empty((T, Cc), device=x.device, dtype=out_dtype)
    via = torch.empty((T, Cc), device=x.device, dtype=out_dtype)
    code = L.dtype_code(enc)
    out = L.GatedForwardOut(L.ptr(enc), code, L.ptr(dec), code, L.ptr(rp), code, L.ptr(via), code)
    L.check(L.load().svb_gated_forward(L.handle(x.device), L.stream_ptr(x.device), C.byref(a), C.byref(p), C.byref(out)),
            "svb_gated_forward")
    return enc, dec, rp, via


# --------------------------------------------------------------------------------------------------- train step
class StepResult:
    """Device-side results of one training step.  `stats` is a float32[16] CUDA tensor (see _lib.STAT); nothing
    here synchronises with the host until the caller reads a value."""

    def __init__(self, stats, dead, freq, n_active, dec):
        self.stats, self.dead, self.freq, self.n_active, self.dec = stats, dead, freq, n_active, dec

    def scalars(self):
        s = self.stats.tolist()          # ONE device->host copy for all scalars
        return {k: s[i] for k, i in L.STAT.items()}


def _train_out(x, a, F, want_dec, dec_dtype, dec_out=None):
    """Outputs of one step.  The four small results are views of ONE uninitialised allocation (the finalise kernel
    writes every element): a fresh block per step, so a StepResult a caller keeps is never overwritten, but no fill
    kernel and a single allocator call per step."""
    dev = x.device
    n_img = int(a.n_images)
    words = L.STATS_LEN + F + n_img
    block = torch.empty(words * 4 + F, device=dev, dtype=torch.uint8)
    f32 = block[:words * 4].view(torch.float32)
    stats, freq = f32[:L.STATS_LEN], f32[L.STATS_LEN:L.STATS_LEN + F]
    n_active = block[(L.STATS_LEN + F) * 4:words * 4].view(torch.int32)
    dead = block[words * 4:]
    dec = None
    if want_dec:   # the reconstruction goes back in the layout the activations came in (channels_last stays so)
        fmt = torch.channels_last if L.is_channels_last_tokens(x) else torch.contiguous_format
        if dec_out is not None:     # caller-provided destination (e.g. the first half of a larger batch tensor)
            if (dec_out.shape != x.shape or dec_out.device != dev or dec_out.dtype != (dec_dtype or x.dtype)
                    or not dec_out.is_contiguous(memory_format=fmt)):
                raise ValueError("dec_out must have the shape, device, dtype and memory format of the reconstruction")
            dec = dec_out
        else:
            dec = torch.empty(x.shape, device=dev, dtype=dec_dtype or x.dtype, memory_format=fmt)
    out = L.TrainOut(L.ptr(dec), L.dtype_code(dec) if want_dec else 0, a.layout, L.ptr(stats),
                     L.ActivityOut(L.ptr(dead), L.ptr(freq), L.ptr(n_active)))
    return out, StepResult(stats, dead, freq, n_active, dec)


def _single_pixel_fixup(x, res):
    """[B,C,1,1] inputs: the reference's variance_explained (utils.py:2012-2030) takes the unbiased variance over the
    H*W = 1 positions of every (image, channel) and therefore reports nan; the library treats such inputs as B tokens."""
    if x.dim() == 4 and x.shape[2] * x.shape[3] == 1:
        res.stats[L.STAT["var_expl"]] = float("nan")
    return res


def sae_train_step(x, params, adam_m, adam_v, step, lr, lam, expansion_factor, optimizer="constrained_adam",
                   betas=(0.9, 0.999), eps=1e-8, want_dec=True, dec_dtype=None, step_dev=None, dec_out=None):
    """One pass of ModelPipeline.hook's train branch (model_pipeline.py:380-420) for SaeMLP, fully on device.
    params = (encoder.weight, encoder.bias, decoder.weight, decoder.bias); updated in place with the Adam moments.
    step_dev: optional int32 CUDA scalar holding the steps taken so far; the call increments it on the device and
    `step` is ignored (CUDA-graph capture: nothing step-dependent in the launch parameters).
    dec_out: optional destination of the reconstruction (shape / dtype / memory format of x), e.g. a slice of a larger
    batch tensor, instead of a fresh allocation."""
    a, x = L.acts_of(x)
    p = _sae_params(*params)
    out, res = _train_out(x, a, p.F, want_dec, dec_dtype, dec_out)
    st = _adam_state(adam_m, adam_v)
    opt = _opt(optimizer, step, lr, betas, eps, step_dev)
    L.check(L.load().svb_sae_train_step(L.handle(x.device), L.stream_ptr(x.device), C.byref(a), C.byref(p), C.byref(st),
                                        C.byref(opt), float(lam), int(expansion_factor), C.byref(out)),
            "svb_sae_train_step")
    return _single_pixel_fixup(x, res)


def gated_train_step(x, params, adam_m, adam_v, step, lr, lam, expansion_factor, optimizer="constrained_adam",
                     betas=(0.9, 0.999), eps=1e-8, want_dec=True, dec_dtype=None, step_dev=None, dec_out=None):
    """Same for GatedSae; params = (W_gate, b_gate, b_mag, r_mag, decoder.weight, decoder.bias)."""
    a, x = L.acts_of(x)
    p = _gated_params(*params)
    out, res = _train_out(x, a, p.F, want_dec, dec_dtype, dec_out)
    st = _adam_state(adam_m, adam_v)
    opt = _opt(optimizer, step, lr, betas, eps, step_dev)
    L.check(L.load().svb_gated_train_step(L.handle(x.device), L.stream_ptr(x.device), C.byref(a), C.byref(p), C.byref(st),
                                          C.byref(opt), float(lam), int(expansion_factor), C.byref(out)),
            "svb_gated_train_step")
    return _single_pixel_fixup(x, res)


class SplitStep:
    """The two halves of a data-parallel step: grads() -> (all-reduce the flat buffer) -> apply()."""

    def __init__(self, kind, x, params, lam, want_dec=True, dec_dtype=None, dec_out=None):
        self.kind = kind
        self.a, self.x = L.acts_of(x)
        self.params = params
        self.p = _sae_params(*params) if kind == "sae_mlp" else _gated_params(*params)
        self.lam = float(lam)
        self.out, self.res = _train_out(self.x, self.a, self.p.F, want_dec, dec_dtype, dec_out)
        self.lib = L.load()
        self.h = L.handle(self.x.device)

    def grads(self, global_tokens=0):
        fn = self.lib.svb_sae_step_grads if self.kind == "sae_mlp" else self.lib.svb_gated_step_grads
        L.check(fn(self.h, L.stream_ptr(self.x.device), C.byref(self.a), C.byref(self.p), self.lam, int(global_tokens),
                   C.byref(self.out)), "svb_*_step_grads")
        buf, n_sum, n_max = L._vp(), C.c_int64(), C.c_int64()
        L.check(self.lib.svb_sae_grad_buffer(self.h, C.byref(buf), C.byref(n_sum), C.byref(n_max)),
                "svb_sae_grad_buffer")
        return buf.value, n_sum.value, n_max.value

    def peer_allreduce(self):
        """In-place all-reduce of the flat buffer over NVLink peer memory (parallel.connect_peer_memory first)."""
        L.check(self.lib.svb_comm_allreduce(self.h, L.stream_ptr(self.x.device)), "svb_comm_allreduce")

    def early_elems(self):
        """Leading elements of the flat buffer that are final once the communication stream set with
        set_comm_stream() is released (0 when no communication stream is set)."""
        n = C.c_int64()
        L.check(self.lib.svb_grad_early_elems(self.h, C.byref(n)), "svb_grad_early_elems")
        return n.value

    def apply(self, adam_m, adam_v, step, lr, expansion_factor, optimizer, betas, eps=1e-8, global_tokens=0,
              global_images=0, step_dev=None):
        fn = self.lib.svb_sae_step_apply if self.kind == "sae_mlp" else self.lib.svb_gated_step_apply
        st = _adam_state(adam_m, adam_v)
        opt = _opt(optimizer, step, lr, betas, eps, step_dev)
        L.check(fn(self.h, L.stream_ptr(self.x.device), C.byref(self.a), C.byref(self.p), C.byref(st), C.byref(opt), self.lam,
                   int(expansion_factor), int(global_tokens), int(global_images), C.byref(self.out)),
                "svb_*_step_apply")
        return _single_pixel_fixup(self.x, self.res)


def set_comm_stream(device, stream):
    """Data-parallel overlap: `stream` (a torch.cuda.Stream or None) is made to wait for the early gradient bucket of
    every following svb_*_step_grads call on `device` (include/svb.h: svb_set_comm_stream)."""
    L.check(L.load().svb_set_comm_stream(L.handle(device), None if stream is None else stream.cuda_stream),
            "svb_set_comm_stream")


def comm_status(device):
    """0 when every peer-memory all-reduce on `device` completed; 1 + r when rank r never arrived within the timeout
    (include/svb.h: svb_comm_status).  Synchronises with the device: call it per logging interval, not per step."""
    st = C.c_int32()
    L.check(L.load().svb_comm_status(L.handle(device), C.byref(st)), "svb_comm_status")
    return st.value


def wrap_device_buffer(address, n_elems, device):
    """A float32 torch view over library-owned device memory (the flat gradient buffer) for torch.distributed."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(n_elems),), "typestr": "<f4", "data": (int(address), False),
                                  "version": 3}
    return torch.as_tensor(h, device=device)


# --------------------------------------------------------------------------------------------------- optimiser
def adam_step(params, grads, ms, vs, step, lr, betas, eps=1e-8, optimizer="adam", decoder_index=-1):
    """utils.py:50-97 on caller-provided gradients (None grads are skipped like torch.optim.Adam does)."""
    n = len(params)
    dev = params[0].device
    arr = L._vp * n
    rows = (C.c_int64 * n)(*[p.shape[0] if p.dim() == 2 else 1 for p in params])
    cols = (C.c_int64 * n)(*[p.shape[1] if p.dim() == 2 else p.numel() for p in params])
    gs = [None if g is None else g.contiguous() for g in grads]
    L.check(L.load().svb_adam_step(
        L.handle(dev), L.stream_ptr(dev), n, arr(*[_f32c(p).data_ptr() for p in params]),
        arr(*[0 if g is None else _f32c(g).data_ptr() for g in gs]), arr(*[_f32c(m).data_ptr() for m in ms]),
        arr(*[_f32c(v).data_ptr() for v in vs]), rows, cols, int(decoder_index),
        C.byref(_opt(optimizer, step, lr, betas, eps))), "svb_adam_step")
    return gs


def reinit_dead(params, adam_m, adam_v, dead_mask_u8, new_w_enc, new_w_dec, new_b_enc):
    """models/sae_mlp.py:133-176 scatter + column renorm + Adam-moment reset for the dead units."""
    p = _sae_params(*params)
    Cc = params[0].shape[1]
    st = _adam_state(adam_m, adam_v) if adam_m is not None else None
    L.check(L.load().svb_reinit_dead(L.handle(params[0].device), L.stream_ptr(params[0].device), C.byref(p), Cc,
                                     C.byref(st) if st is not None else None, L.ptr(dead_mask_u8.contiguous()),
                                     L.ptr(_f32c(new_w_enc)), L.ptr(_f32c(new_w_dec)), float(new_b_enc)),
            "svb_reinit_dead")


# --------------------------------------------------------------------------------------------------- activity
def measure_inactive(t):
    """utils.py:2032-2069 on a [B,F,H,W] or [N,F] tensor -> (dead uint8 [F], freq [F], n_active int32 [rows])."""
    if not t.is_cuda:
        raise ValueError("CUDA tensor required")
    t = t.contiguous()
    if t.dim() == 4:
        b, f, h, w = t.shape
        layout, n_img, hw, rows = L.SVB_NCHW, b, h * w, b
    elif t.dim() == 2:
        n, f = t.shape
        layout, n_img, hw, rows = L.SVB_TOKENS, n, 1, n
    else:
        raise ValueError(f"Output has unexpected shape {t.dim()}.")
    dead = torch.empty(f, device=t.device, dtype=torch.uint8)
    freq = torch.empty(f, device=t.device, dtype=torch.float32)
    n_active = torch.empty(rows, device=t.device, dtype=torch.int32)
    act = L.ActivityOut(L.ptr(dead), L.ptr(freq), L.ptr(n_active))
    L.check(L.load().svb_measure_inactive(L.handle(t.device), L.stream_ptr(t.device), L.ptr(t), L.dtype_code(t), layout,
                                          n_img, hw, f, C.byref(act)), "svb_measure_inactive")
    return dead, freq, n_active


# --------------------------------------------------------------------------------------------------- eval metrics
def spatial_mean(t, n_images=None):
    """utils.py:1996-2010 average_over_W_H on the device: [B,F,H,W] -> [B,F]; token-major [B*hw, F] with n_images
    given -> [B,F]; a 2-D tensor without n_images is returned as it is (already one value per sample and unit)."""
    if not t.is_cuda:
        raise ValueError("CUDA tensor required")
    if t.dim() == 2 and n_images is None:
        return t
    t = t.contiguous()
    if t.dim() == 4:
        b, f, hh, ww = t.shape
        layout, hw = L.SVB_NCHW, hh * ww
    elif t.dim() == 2:
        b, f, layout = int(n_images), t.shape[1], L.SVB_TOKENS
        hw = t.shape[0] // b
        if hw * b != t.shape[0]:
            raise ValueError("token count is not a multiple of n_images")
    else:
        raise ValueError(f"Output has unexpected shape {t.dim()}.")
    out = torch.empty((b, f), device=t.device, dtype=torch.float32)
    L.check(L.load().svb_spatial_mean(L.handle(t.device), L.stream_ptr(t.device), L.ptr(t), L.dtype_code(t), layout, b,
                                      hw, f, L.ptr(out)), "svb_spatial_mean")
    return out


def image_sum(t, n_images):
    """Sum over the images of a batch, per position (compute_ie.py:146-207): t is [n_images * R, F] token-major or
    [n_images, C, H, W]; returns float32 [R, F] resp. [C, H, W]."""
    if not t.is_cuda:
        raise ValueError("CUDA tensor required")
    t = t.contiguous()
    per = t.numel() // int(n_images)
    if per * int(n_images) != t.numel():
        raise ValueError("element count is not a multiple of n_images")
    shape = tuple(t.shape[1:]) if t.dim() == 4 else (t.shape[0] // int(n_images), t.shape[1])
    out = torch.empty(shape, device=t.device, dtype=torch.float32)
    L.check(L.load().svb_image_sum(L.handle(t.device), L.stream_ptr(t.device), L.ptr(t), L.dtype_code(t), int(n_images),
                                   per, L.ptr(out)), "svb_image_sum")
    return out


def _i64(t):
    if t is None:
        return None
    if t.dtype != torch.int64 or not t.is_cuda:
        raise ValueError("index payloads must be int64 CUDA tensors")
    return t.contiguous()


def topk_columns(values, k, largest=True, indices=None, files=None, values2=None, indices2=None, files2=None):
    """torch.topk(values, k, dim=0, largest=largest) per column on the device; with a second source the candidates are
    the rows of `values` followed by the rows of `values2` (the merge of utils.py:1463-1477).  indices / files are
    optional int64 payloads of the same shape as their values, gathered along; indices=None stands for the row number.
    Returns (values [k,F], indices [k,F] int64, files [k,F] int64 | None)."""
    v0 = values.contiguous().float()
    n0, F = v0.shape
    v1 = values2.contiguous().float() if values2 is not None else None
    n1 = v1.shape[0] if v1 is not None else 0
    i0, f0, i1, f1 = _i64(indices), _i64(files), _i64(indices2), _i64(files2)
    want_files = f0 is not None or f1 is not None
    out_v = torch.empty((k, F), device=v0.device, dtype=torch.float32)
    out_i = torch.empty((k, F), device=v0.device, dtype=torch.int64)
    out_f = torch.empty((k, F), device=v0.device, dtype=torch.int64) if want_files else None
    L.check(L.load().svb_topk_columns(L.handle(v0.device), L.stream_ptr(v0.device), L.ptr(v0), L.ptr(i0), L.ptr(f0), n0,
                                      L.ptr(v1), L.ptr(i1), L.ptr(f1), n1, F, int(k), int(bool(largest)), L.ptr(out_v),
                                      L.ptr(out_i), L.ptr(out_f)), "svb_topk_columns")
    return out_v, out_i, out_f


def histogram_update(hist, values, unit_idx, mins, maxs):
    """hist [bins, U] (float32, updated in place) += per-unit torch.histc of values[:, unit_idx[u]] (utils.py:1956-1960)."""
    values = values.contiguous().float()
    bins, U = hist.shape
    if not hist.is_contiguous() or hist.dtype != torch.float32:
        raise ValueError("hist must be a contiguous float32 [bins, units] tensor")
    idx = _i64(unit_idx)
    L.check(L.load().svb_histogram_update(L.handle(values.device), L.stream_ptr(values.device), L.ptr(values),
                                          values.shape[0], values.shape[1], L.ptr(idx), U,
                                          L.ptr(mins.contiguous().float()), L.ptr(maxs.contiguous().float()), bins,
                                          L.ptr(hist)), "svb_histogram_update")
    return hist


# --------------------------------------------------------------------------------------------------- IE
def ie_channelwise(a, avg, g, batch_size, scale=None):
    """utils.py:2606-2637: a, g [B*H*W, F] (f32/bf16), avg [F,H,W] f32 -> [F]."""
    F, H, W = avg.shape
    if a.shape != g.shape or a.shape[0] != batch_size * H * W or a.shape[1] != F:
        raise ValueError("compute_ie_channel_wise: inconsistent shapes")
    a, g = a.contiguous(), g.contiguous()
    if g.dtype != a.dtype:
        g = g.to(a.dtype)
    avg = avg.contiguous().float()
    out = torch.empty(F, device=a.device, dtype=torch.float32)
    sc = (1.0 / a.shape[0]) if scale is None else scale
    L.check(L.load().svb_ie_channelwise(L.handle(a.device), L.stream_ptr(a.device), L.ptr(a), L.ptr(g), L.dtype_code(a),
                                        L.ptr(avg), batch_size, H * W, F, float(sc), L.ptr(out)),
            "svb_ie_channelwise")
    return out


def ie_allchannels(err, avg, g, batch_size, scale=None):
    """utils.py:2574-2602: err, g [B,C,H,W], avg [C,H,W] -> scalar tensor."""
    B, Cc, H, W = err.shape
    if B != batch_size or tuple(avg.shape) != (Cc, H, W) or g.shape != err.shape:
        raise ValueError("compute_ie_all_channels: inconsistent shapes")
    err, g = err.contiguous(), g.contiguous()
    if g.dtype != err.dtype:
        g = g.to(err.dtype)
    avg = avg.contiguous().float()
    out = torch.empty(1, device=err.device, dtype=torch.float32)
    sc = (1.0 / (B * H * W)) if scale is None else scale
    L.check(L.load().svb_ie_allchannels(L.handle(err.device), L.stream_ptr(err.device), L.ptr(err), L.ptr(g),
                                        L.dtype_code(err), L.ptr(avg), B, Cc, H * W, float(sc), L.ptr(out)),
            "svb_ie_allchannels")
    return out[0]


def node_ie_layer(x, grad, params, enc_avg, err_avg, x_avg, scale=None):
    """compute_ie.py:242-267,442-453 for one layer / one batch -> (ie_sae_features [F], ie_sae_error, ie_neurons [C])."""
    a, x = L.acts_of(x)
    # the gradient is read through the same descriptor as x: same memory format
    grad = grad.contiguous(memory_format=torch.channels_last) if L.is_channels_last_tokens(x) else grad.contiguous()
    if grad.dtype != x.dtype or grad.shape != x.shape:
        raise ValueError("grad must match x in dtype and shape")
    p = _sae_params(*params)
    dev = x.device
    feat = torch.empty(p.F, device=dev, dtype=torch.float32)
    err = torch.empty(1, device=dev, dtype=torch.float32)
    neur = torch.empty(a.C, device=dev, dtype=torch.float32)
    sc = (1.0 / _tokens(x)) if scale is None else scale
    L.check(L.load().svb_node_ie_layer(L.handle(dev), L.stream_ptr(dev), C.byref(a), L.ptr(grad), C.byref(p),
                                       L.ptr(enc_avg.contiguous().float()), L.ptr(err_avg.contiguous().float()),
                                       L.ptr(x_avg.contiguous().float()), float(sc), L.ptr(feat), L.ptr(err),
                                       L.ptr(neur)), "svb_node_ie_layer")
    return feat, err[0], neur


# --------------------------------------------------------------------------------------------------- generic GEMM
def gemm_bf16(A, B, a_mn=False, b_mn=False, out_dtype=torch.float32, alpha=1.0, bias=None, relu=False):
    """D[M,N] = alpha * A[M,K] @ B[N,K]^T on the tcgen05 kernel.  With a_mn / b_mn the operand is given as its
    transpose ([K,M] / [K,N] row-major), which is how the backward GEMMs read token-major tensors without a copy."""
    A = A.contiguous() if A.dtype == torch.bfloat16 else A.to(torch.bfloat16).contiguous()
    B = B.contiguous() if B.dtype == torch.bfloat16 else B.to(torch.bfloat16).contiguous()
    M, K = (A.shape[1], A.shape[0]) if a_mn else (A.shape[0], A.shape[1])
    N, Kb = (B.shape[1], B.shape[0]) if b_mn else (B.shape[0], B.shape[1])
    if K != Kb:
        raise ValueError(f"gemm_bf16: inner dimensions differ ({K} vs {Kb})")
    out = torch.empty((M, N), device=A.device, dtype=out_dtype)
    L.check(L.load().svb_gemm_bf16(L.handle(A.device), L.stream_ptr(A.device), L.ptr(A), int(a_mn), A.shape[1], L.ptr(B),
                                   int(b_mn), B.shape[1], M, N, K, L.ptr(out), L.dtype_code(out), N, float(alpha),
                                   L.ptr(bias.contiguous().float()) if bias is not None else None, int(relu)),
            "svb_gemm_bf16")
    return out


# ---------------------------------------------------------------------------------- activation producer
def _is_nhwc_bf16(x):
    return (x.is_cuda and x.dim() == 4 and x.dtype == torch.bfloat16
            and x.is_contiguous(memory_format=torch.channels_last))


def pool_output_size(n, kernel, stride, pad, ceil_mode):
    """torch's pooling_output_shape for dilation 1 (what nn.MaxPool2d produces)."""
    o = (n + 2 * pad - (kernel - 1) - 1 + (stride - 1 if ceil_mode else 0)) // stride + 1
    if ceil_mode and (o - 1) * stride >= n + pad:
        o -= 1
    return o


def maxpool_nhwc(x, kernel, stride, pad=0, ceil_mode=False):
    """torch.nn.functional.max_pool2d on a bf16 channels_last CUDA tensor (include/svb.h: svb_maxpool_nhwc); exact."""
    if not _is_nhwc_bf16(x):
        raise ValueError("maxpool_nhwc takes a bf16 channels_last CUDA tensor [B,C,H,W]")
    b, c, h, w = x.shape
    oh, ow = pool_output_size(h, kernel, stride, pad, ceil_mode), pool_output_size(w, kernel, stride, pad, ceil_mode)
    out = torch.empty((b, c, oh, ow), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    L.check(L.load().svb_maxpool_nhwc(L.handle(x.device), L.stream_ptr(x.device), L.ptr(x), b, h, w, c, int(kernel),
                                      int(stride), int(pad), int(bool(ceil_mode)), L.ptr(out), oh, ow), "svb_maxpool_nhwc")
    return out


def bias_relu_scatter(src, bias, dests, relu=True):
    """relu(src + bias) of a dense bf16 channels_last convolution output, written in ONE pass into channel ranges of
    other channels_last tensors (include/svb.h: svb_bias_relu_scatter).  `dests` is a list of (tensor, channel_offset,
    channel_count): consecutive source channel ranges in order; each tensor has the batch / spatial size of `src`.
    `dests = [(src, 0, C)]` is the in-place bias + relu.  Bit-identical to `src.add_(bias).relu_()` + torch.cat."""
    if not _is_nhwc_bf16(src):
        raise ValueError("bias_relu_scatter takes a bf16 channels_last CUDA tensor [B,C,H,W]")
    b, c, h, w = src.shape
    if bias.dtype != torch.bfloat16 or bias.numel() != c or not bias.is_contiguous():
        raise ValueError("bias must be a contiguous bf16 vector of C elements")
    if not 1 <= len(dests) <= L.MAX_CHAN_SEGMENTS:
        raise ValueError(f"1..{L.MAX_CHAN_SEGMENTS} destinations")
    segs = (L.ChanSegment * len(dests))()
    begin = 0
    for i, (dst, off, count) in enumerate(dests):
        if not _is_nhwc_bf16(dst) or dst.shape[0] != b or tuple(dst.shape[2:]) != (h, w) or dst.device != src.device:
            raise ValueError("every destination is a bf16 channels_last tensor with the batch and spatial size of src")
        segs[i] = L.ChanSegment(dst.data_ptr(), begin, int(count), dst.shape[1], int(off))
        begin += int(count)
    L.check(L.load().svb_bias_relu_scatter(L.handle(src.device), L.stream_ptr(src.device), L.ptr(src), L.ptr(bias),
                                           b * h * w, c, segs, len(dests), int(bool(relu))), "svb_bias_relu_scatter")


def conv1_pack_weights(weight):
    """[64, 3, 7, 7] bf16 CUDA weights (any memory format) -> the packed B operand of conv1_stem."""
    if not (weight.is_cuda and weight.dtype == torch.bfloat16 and tuple(weight.shape) == (64, 3, 7, 7)):
        raise ValueError("conv1_pack_weights takes the bf16 CUDA weights [64, 3, 7, 7] of GoogLeNet's conv1")
    packed = torch.empty(L.CONV1_PACKED_ELEMS, device=weight.device, dtype=torch.bfloat16)
    so, si, sh, sw = weight.stride()
    L.check(L.load().svb_conv1_pack_weights(L.handle(weight.device), L.stream_ptr(weight.device), L.ptr(weight), so, si,
                                            sh, sw, L.ptr(packed)), "svb_conv1_pack_weights")
    return packed


def conv1_stem(x, packed_weight, bias, relu=True):
    """relu(conv2d(x, w, stride 2, pad 3) + bias) for GoogLeNet's 7x7 stem on bf16 channels_last [B, 3, 224, 224] images
    (include/svb.h: svb_conv1_7x7s2_nhwc) -> bf16 channels_last [B, 64, 112, 112]."""
    if not _is_nhwc_bf16(x) or tuple(x.shape[1:]) != (3, 224, 224):
        raise ValueError("conv1_stem takes bf16 channels_last CUDA images [B, 3, 224, 224]")
    if packed_weight.numel() != L.CONV1_PACKED_ELEMS or packed_weight.dtype != torch.bfloat16:
        raise ValueError("packed_weight comes from conv1_pack_weights")
    if bias.dtype != torch.bfloat16 or bias.numel() != 64 or not bias.is_contiguous():
        raise ValueError("bias must be a contiguous bf16 vector of 64 elements")
    out = torch.empty((x.shape[0], 64, 112, 112), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    L.check(L.load().svb_conv1_7x7s2_nhwc(L.handle(x.device), L.stream_ptr(x.device), L.ptr(x), x.shape[0],
                                          L.ptr(packed_weight), L.ptr(bias), int(bool(relu)), L.ptr(out)),
            "svb_conv1_7x7s2_nhwc")
    return out


class _MaxPoolNhwcFn(torch.autograd.Function):
    """max_pool2d on bf16 channels_last CUDA tensors with a backward (svb_maxpool_nhwc_argmax /
    svb_maxpool_nhwc_backward): what the IE passes differentiate through behind the first hooked layer."""

    @staticmethod
    def forward(ctx, x, kernel, stride, pad, ceil_mode):
        b, c, h, w = x.shape
        oh, ow = pool_output_size(h, kernel, stride, pad, ceil_mode), pool_output_size(w, kernel, stride, pad, ceil_mode)
        out = torch.empty((b, c, oh, ow), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        arg = torch.empty((b, oh, ow, c), device=x.device, dtype=torch.uint8)
        L.check(L.load().svb_maxpool_nhwc_argmax(L.handle(x.device), L.stream_ptr(x.device), L.ptr(x), b, h, w, c,
                                                 int(kernel), int(stride), int(pad), int(bool(ceil_mode)), L.ptr(out),
                                                 L.ptr(arg), oh, ow), "svb_maxpool_nhwc_argmax")
        ctx.save_for_backward(arg)
        ctx.geom = (b, c, h, w, int(kernel), int(stride), int(pad), oh, ow)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (arg,) = ctx.saved_tensors
        b, c, h, w, kernel, stride, pad, oh, ow = ctx.geom
        g = grad_out.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        grad_in = torch.empty((b, c, h, w), device=g.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
        L.check(L.load().svb_maxpool_nhwc_backward(L.handle(g.device), L.stream_ptr(g.device), L.ptr(g), L.ptr(arg), b, h,
                                                   w, c, kernel, stride, pad, oh, ow, L.ptr(grad_in)),
                "svb_maxpool_nhwc_backward")
        return grad_in, None, None, None, None


def maxpool_nhwc_autograd(x, kernel, stride, pad=0, ceil_mode=False):
    """maxpool_nhwc for a tensor that requires grad (indices kept for the backward; ties like torch's)."""
    if not _is_nhwc_bf16(x):
        raise ValueError("maxpool_nhwc_autograd takes a bf16 channels_last CUDA tensor [B,C,H,W]")
    return _MaxPoolNhwcFn.apply(x, kernel, stride, pad, ceil_mode)


def relu_grad_gather(sources, like):
    """Backward of bias_relu_scatter (include/svb.h: svb_relu_grad_gather).  `sources` is a list of
    (grad, grad_offset, y, y_offset, count): channel range [grad_offset, grad_offset + count) of the gradient tensor
    `grad` masked by the forward RESULT `y` at [y_offset, ...); the ranges are laid side by side into a new dense bf16
    channels_last tensor [B, sum(count), H, W] (B, H, W from `like`) = the gradient of the convolution output."""
    b, _, h, w = like.shape
    total = sum(int(s[4]) for s in sources)
    if not 1 <= len(sources) <= L.MAX_CHAN_SEGMENTS:
        raise ValueError(f"1..{L.MAX_CHAN_SEGMENTS} sources")
    dst = torch.empty((b, total, h, w), device=like.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
    segs = (L.GradSegment * len(sources))()
    begin = 0
    for i, (g, g_off, y, y_off, count) in enumerate(sources):
        for t in (g, y):
            if not _is_nhwc_bf16(t) or t.shape[0] != b or tuple(t.shape[2:]) != (h, w) or t.device != like.device:
                raise ValueError("gradients and results are bf16 channels_last tensors of the same batch and spatial size")
        segs[i] = L.GradSegment(g.data_ptr(), y.data_ptr(), begin, int(count), g.shape[1], int(g_off), y.shape[1], int(y_off))
        begin += int(count)
    L.check(L.load().svb_relu_grad_gather(L.handle(like.device), L.stream_ptr(like.device), b * h * w, total, segs,
                                          len(sources), L.ptr(dst)), "svb_relu_grad_gather")
    return dst


def maxpool_nhwc_with_argmax(x, kernel, stride, pad=0, ceil_mode=False):
    """(pooled, argmax bytes [B, OH, OW, C]) -- the forward half of maxpool_nhwc_autograd for callers that write their
    own backward."""
    b, c, h, w = x.shape
    oh, ow = pool_output_size(h, kernel, stride, pad, ceil_mode), pool_output_size(w, kernel, stride, pad, ceil_mode)
    out = torch.empty((b, c, oh, ow), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    arg = torch.empty((b, oh, ow, c), device=x.device, dtype=torch.uint8)
    L.check(L.load().svb_maxpool_nhwc_argmax(L.handle(x.device), L.stream_ptr(x.device), L.ptr(x), b, h, w, c, int(kernel),
                                             int(stride), int(pad), int(bool(ceil_mode)), L.ptr(out), L.ptr(arg), oh, ow),
            "svb_maxpool_nhwc_argmax")
    return out, arg


def maxpool_nhwc_backward(grad_out, arg, in_shape, kernel, stride, pad):
    b, c, h, w = in_shape
    oh, ow = grad_out.shape[2], grad_out.shape[3]
    grad_in = torch.empty((b, c, h, w), device=grad_out.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
    L.check(L.load().svb_maxpool_nhwc_backward(L.handle(grad_out.device), L.stream_ptr(grad_out.device), L.ptr(grad_out),
                                               L.ptr(arg), b, h, w, c, int(kernel), int(stride), int(pad), oh, ow,
                                               L.ptr(grad_in)), "svb_maxpool_nhwc_backward")
    return grad_in
