"""ctypes binding of libsvb.so (include/svb.h).  PyTorch only supplies device memory and streams here; every
entry point receives raw device pointers.  There is NO fallback: if the shared library is missing or a call fails,
an exception is raised.
"""
import ctypes as C
import os
import subprocess
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvb.so")

SVB_F32, SVB_BF16 = 0, 1
SVB_TOKENS, SVB_NCHW = 0, 1
SVB_ADAM, SVB_CONSTRAINED_ADAM = 0, 1
STATS_LEN = 16
STAT = {"loss": 0, "rec": 1, "l1": 2, "nrmse": 3, "rmse": 4, "aux": 5, "var_expl": 6, "sparsity": 7, "n_dead": 8}

_vp = C.c_void_p
_fp = C.c_void_p  # float* passed as raw address


class SvbError(RuntimeError):
    pass


class Acts(C.Structure):
    _fields_ = [("x", _vp), ("dtype", C.c_int32), ("layout", C.c_int32), ("n_images", C.c_int64),
                ("hw", C.c_int32), ("C", C.c_int32)]


class SaeParams(C.Structure):
    _fields_ = [("w_enc", _fp), ("b_enc", _fp), ("w_dec", _fp), ("b_dec", _fp), ("F", C.c_int32)]


class GatedParams(C.Structure):
    _fields_ = [("w_gate", _fp), ("b_gate", _fp), ("b_mag", _fp), ("r_mag", _fp), ("w_dec", _fp), ("b_dec", _fp),
                ("F", C.c_int32)]


class AdamState(C.Structure):
    _fields_ = [("m", _fp * 6), ("v", _fp * 6)]


class OptConfig(C.Structure):
    _fields_ = [("optimizer", C.c_int32), ("step", C.c_int32), ("lr", C.c_double), ("beta1", C.c_double),
                ("beta2", C.c_double), ("eps", C.c_double), ("step_dev", _vp)]


class ActivityOut(C.Structure):
    _fields_ = [("dead", _vp), ("freq", _fp), ("n_active", _vp)]


class TrainOut(C.Structure):
    _fields_ = [("dec_out", _vp), ("dec_dtype", C.c_int32), ("dec_layout", C.c_int32), ("stats", _fp),
                ("activity", ActivityOut)]


class SaeForwardOut(C.Structure):
    _fields_ = [("enc", _vp), ("enc_dtype", C.c_int32), ("pre", _fp), ("dec", _vp), ("dec_dtype", C.c_int32)]


class GatedForwardOut(C.Structure):
    _fields_ = [("enc", _vp), ("enc_dtype", C.c_int32), ("dec", _vp), ("dec_dtype", C.c_int32),
                ("relu_pi", _vp), ("relu_pi_dtype", C.c_int32), ("via", _vp), ("via_dtype", C.c_int32)]


class ChanSegment(C.Structure):
    _fields_ = [("dst", _vp), ("c_begin", C.c_int32), ("c_count", C.c_int32), ("dst_channels", C.c_int32),
                ("dst_offset", C.c_int32)]


class GradSegment(C.Structure):
    _fields_ = [("grad", _vp), ("y", _vp), ("c_begin", C.c_int32), ("c_count", C.c_int32), ("grad_channels", C.c_int32),
                ("grad_offset", C.c_int32), ("y_channels", C.c_int32), ("y_offset", C.c_int32)]


MAX_CHAN_SEGMENTS = 4
CONV1_PACKED_ELEMS = 64 * 168

# every symbol include/svb.h declares: (name, restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "svb_last_error": (C.c_char_p, []),
    "svb_version": (C.c_int, []),
    "svb_create": (C.c_int, [_P(_vp)]),
    "svb_destroy": (C.c_int, [_vp]),
    "svb_workspace_bytes": (C.c_int64, [_vp]),
    "svb_launch_count": (C.c_int64, []),
    "svb_last_step_flags": (C.c_int32, [_vp]),
    "svb_set_tuning": (C.c_int, [C.c_int32, C.c_int32]),
    "svb_get_tuning": (C.c_int32, [C.c_int32]),
    "svb_profile_enable": (C.c_int, [_vp, C.c_int32]),
    "svb_profile_read": (C.c_int, [_vp, C.c_int32, _P(C.c_float), _P(C.c_int32), _P(C.c_int32)]),
    "svb_profile_phase_name": (C.c_char_p, [C.c_int32]),
    "svb_sae_forward": (C.c_int, [_vp, _vp, _P(Acts), _P(SaeParams), _P(SaeForwardOut)]),
    "svb_sae_train_step": (C.c_int, [_vp, _vp, _P(Acts), _P(SaeParams), _P(AdamState), _P(OptConfig), C.c_float,
                                     C.c_int32, _P(TrainOut)]),
    "svb_sae_step_grads": (C.c_int, [_vp, _vp, _P(Acts), _P(SaeParams), C.c_float, C.c_int64, _P(TrainOut)]),
    "svb_sae_step_apply": (C.c_int, [_vp, _vp, _P(Acts), _P(SaeParams), _P(AdamState), _P(OptConfig), C.c_float,
                                     C.c_int32, C.c_int64, C.c_int64, _P(TrainOut)]),
    "svb_sae_grad_buffer": (C.c_int, [_vp, _P(_vp), _P(C.c_int64), _P(C.c_int64)]),
    "svb_set_comm_stream": (C.c_int, [_vp, _vp]),
    "svb_comm_alloc": (C.c_int, [_vp, C.c_int64, _vp]),
    "svb_comm_connect": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp]),
    "svb_comm_region_bytes": (C.c_int, [C.c_int64, _P(C.c_int64), _P(C.c_int64)]),
    "svb_comm_attach": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int64, _P(_vp), _vp]),
    "svb_comm_capacity": (C.c_int, [_vp, _P(C.c_int64)]),
    "svb_comm_allreduce": (C.c_int, [_vp, _vp]),
    "svb_comm_destroy": (C.c_int, [_vp]),
    "svb_comm_set_timeout": (C.c_int, [_vp, C.c_double]),
    "svb_comm_status": (C.c_int, [_vp, _P(C.c_int32)]),
    "svb_grad_early_elems": (C.c_int, [_vp, _P(C.c_int64)]),
    "svb_gated_forward": (C.c_int, [_vp, _vp, _P(Acts), _P(GatedParams), _P(GatedForwardOut)]),
    "svb_gated_train_step": (C.c_int, [_vp, _vp, _P(Acts), _P(GatedParams), _P(AdamState), _P(OptConfig), C.c_float,
                                       C.c_int32, _P(TrainOut)]),
    "svb_gated_step_grads": (C.c_int, [_vp, _vp, _P(Acts), _P(GatedParams), C.c_float, C.c_int64, _P(TrainOut)]),
    "svb_gated_step_apply": (C.c_int, [_vp, _vp, _P(Acts), _P(GatedParams), _P(AdamState), _P(OptConfig), C.c_float,
                                       C.c_int32, C.c_int64, C.c_int64, _P(TrainOut)]),
    "svb_maxpool_nhwc": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_int32, _vp, C.c_int32, C.c_int32]),
    "svb_maxpool_nhwc_argmax": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_int32, _vp, _vp, C.c_int32, C.c_int32]),
    "svb_maxpool_nhwc_backward": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "svb_relu_grad_gather": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, _P(GradSegment), C.c_int32, _vp]),
    "svb_bias_relu_scatter": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, C.c_int32, _P(ChanSegment), C.c_int32, C.c_int32]),
    "svb_conv1_pack_weights": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp]),
    "svb_conv1_7x7s2_nhwc": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, _vp, C.c_int32, _vp]),
    "svb_adam_step": (C.c_int, [_vp, _vp, C.c_int32, _P(_vp), _P(_vp), _P(_vp), _P(_vp), _P(C.c_int64),
                                _P(C.c_int64), C.c_int32, _P(OptConfig)]),
    "svb_reinit_dead": (C.c_int, [_vp, _vp, _P(SaeParams), C.c_int32, _P(AdamState), _vp, _fp, _fp, C.c_float]),
    "svb_measure_inactive": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                       _P(ActivityOut)]),
    "svb_ie_channelwise": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int32, _fp, C.c_int64, C.c_int32, C.c_int32, C.c_float,
                                     _fp]),
    "svb_ie_allchannels": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int32, _fp, C.c_int64, C.c_int32, C.c_int32, C.c_float,
                                     _fp]),
    "svb_node_ie_layer": (C.c_int, [_vp, _vp, _P(Acts), _vp, _P(SaeParams), _fp, _fp, _fp, C.c_float, _fp, _fp, _fp]),
    "svb_gemm_bf16": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int64, _vp, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                C.c_int32, _vp, C.c_int32, C.c_int64, C.c_float, _fp, C.c_int32]),
    "svb_spatial_mean": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, _fp]),
    "svb_image_sum": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int64, C.c_int64, _fp]),
    "svb_topk_columns": (C.c_int, [_vp, _vp, _fp, _vp, _vp, C.c_int32, _fp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, _fp, _vp, _vp]),
    "svb_histogram_update": (C.c_int, [_vp, _vp, _fp, C.c_int64, C.c_int32, _vp, C.c_int32, _fp, _fp, C.c_int32, _fp]),
    "svb_pack_tokens": (C.c_int, [_vp, _vp, _P(Acts), _vp]),
    "svb_unpack_tokens": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int64, C.c_int32, C.c_int32, _vp, C.c_int32]),
}

_lib = None
_lock = threading.Lock()


def build(verbose=False):
    """Compiles libsvb.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    res = subprocess.run(["make", "-C", csrc, "-j4"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise SvbError("building libsvb.so failed")
    return LIB_PATH


def load():
    """Loads libsvb.so and binds every symbol of include/svb.h.  Raises if the library is absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise SvbError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           f"or `make -C sparse_vision_b200/csrc` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc, what):
    if rc != 0:
        msg = load().svb_last_error().decode("utf-8", "replace")
        raise SvbError(f"{what} failed (status {rc}): {msg}")


_handles = {}


def handle(device=None):
    """One svb_handle per CUDA device, created lazily."""
    if not torch.cuda.is_available():
        raise SvbError("sparse_vision_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    if dev is None:
        dev = torch.cuda.current_device()
    if dev not in _handles:
        lib = load()
        h = _vp()
        with torch.cuda.device(dev):
            check(lib.svb_create(C.byref(h)), "svb_create")
        _handles[dev] = h
    return _handles[dev]


def stream_ptr(device=None):
    """The current torch stream OF `device` (the tensors' device, which need not be the current one)."""
    return _vp(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def dtype_code(t):
    if t.dtype == torch.float32:
        return SVB_F32
    if t.dtype == torch.bfloat16:
        return SVB_BF16
    raise ValueError(f"unsupported dtype {t.dtype}: sparse_vision_b200 takes float32 or bfloat16 tensors")


def is_channels_last_tokens(x):
    """True for a 4-D bf16 tensor whose memory is NHWC-dense (what a channels_last cuDNN model emits) and that is not
    also NCHW-dense (H*W == 1 or C == 1 are both at once; those take the ordinary route)."""
    return (x.dim() == 4 and x.dtype == torch.bfloat16 and x.shape[2] * x.shape[3] > 1 and x.shape[1] > 1
            and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last))


def acts_of(x):
    """Builds svb_acts from a [B,C,H,W] or [N,C] CUDA tensor (made contiguous)."""
    if not x.is_cuda:
        raise ValueError("sparse_vision_b200 runs on CUDA tensors only (no CPU fallback)")
    if is_channels_last_tokens(x):
        # a channels_last [B,C,H,W] tensor IS the token matrix [(b h w), C] of sae_mlp.py:44: no layout copy
        b, c, h, w = x.shape
        return Acts(ptr(x), dtype_code(x), SVB_TOKENS, b, h * w, c), x
    x = x.contiguous()
    if x.dim() == 4:
        b, c, h, w = x.shape
        a = Acts(ptr(x), dtype_code(x), SVB_NCHW, b, h * w, c)
    elif x.dim() == 2:
        n, c = x.shape
        a = Acts(ptr(x), dtype_code(x), SVB_TOKENS, n, 1, c)
    else:
        raise ValueError(f"Output has unexpected shape {x.dim()}.")
    return a, x
