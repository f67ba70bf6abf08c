"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/svb.h declares, the host logic
mirrors the reference (initialisation RNG stream, schedule, error behaviour) and the data-parallel plumbing works on
gloo with world_size 2.  No compute entry point is called here."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from sparse_vision_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from sparse_vision_b200 import _lib
    header = open(os.path.join(ROOT, "include", "svb.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(svb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.svb_version() >= 100
    assert isinstance(lib.svb_launch_count(), int)


def test_no_cpu_fallback():
    from sparse_vision_b200 import _lib
    from sparse_vision_b200.models import GatedSae, SaeMLP
    if torch.cuda.is_available():
        pytest.skip("needs a GPU-less machine")
    with pytest.raises(_lib.SvbError):
        _lib.handle()
    with pytest.raises(ValueError):
        SaeMLP(16, 4)(torch.randn(3, 16))
    with pytest.raises(ValueError):
        GatedSae(16, 4)(torch.randn(3, 16))


def test_product_package_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sparse_vision_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), os.path.join(dirpath, f)


def test_module_init_matches_reference_rng_stream(golden_dir):
    from sparse_vision_b200.models import GatedSae, SaeMLP
    g = dict(np.load(os.path.join(golden_dir, "cfg1_mlp_adam.npz")))
    torch.manual_seed(0)
    m = SaeMLP((16,), 4)
    assert [k for k, _ in m.named_parameters()] == ["encoder.weight", "encoder.bias", "decoder.weight", "decoder.bias"]
    for k, v in m.state_dict().items():
        assert np.array_equal(v.numpy(), g["init." + k]), k
    g = dict(np.load(os.path.join(golden_dir, "conv_gated_cadam.npz")))
    torch.manual_seed(0)
    m = GatedSae(32, 4)
    assert [k for k, _ in m.named_parameters()] == ["W_gate", "b_gate", "b_mag", "r_mag", "decoder.weight",
                                                    "decoder.bias"]
    # the planted-dead fixture edits b_gate / b_mag after construction; weights must match bit for bit
    for k in ("W_gate", "decoder.weight", "r_mag", "decoder.bias"):
        assert np.array_equal(m.state_dict()[k].numpy(), g["init." + k]), k


def test_sae_conv_shell_matches_reference(golden_dir):
    from sparse_vision_b200.models import SaeConv
    g = dict(np.load(os.path.join(golden_dir, "sae_conv.npz")))
    m = SaeConv((8, 6, 6), 2)
    m.load_state_dict({k[len("init."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("init.")})
    with torch.no_grad():
        enc, dec = m(torch.from_numpy(g["x"]))
    np.testing.assert_allclose(enc.numpy(), g["enc"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dec.numpy(), g["dec"], rtol=1e-5, atol=1e-6)


def test_dead_neuron_schedule_matches_reference(golden_dir):
    from sparse_vision_b200.model_pipeline import dead_neuron_action
    g = dict(np.load(os.path.join(golden_dir, "schedule.npz")))
    for n, upto in ((9912, 100000), (8, 70)):
        assert [i for i in range(1, upto + 1) if dead_neuron_action(i, n) == "reinit"] == list(g[f"reinit_{n}"])
        assert [i for i in range(1, upto + 1) if dead_neuron_action(i, n) == "clear"] == list(g[f"wait_{n}"])


def test_optimizer_and_criterion_factories():
    from sparse_vision_b200 import utils as U
    from sparse_vision_b200.models import SaeMLP
    m = SaeMLP(16, 2)
    opt, sched = U.get_optimizer("constrained_adam", m, 1e-3)
    assert opt.__class__.__name__ == "ConstrainedAdam" and sched is None
    assert opt.param_groups[0]["betas"] == (0.9, 0.999) and opt.p is m.decoder.weight
    opt, _ = U.get_optimizer("adam", m, 1e-3)
    assert opt.__class__.__name__ == "Adam" and opt.param_groups[0]["betas"] == (0.9, 0.9999)
    with pytest.raises(ValueError, match="Unsupported optimizer"):
        U.get_optimizer("lion", m, 1e-3)
    assert U.get_criterion("sae_loss").__class__.__name__ == "SparseLoss"
    assert U.get_criterion("gated_sae_loss").__class__.__name__ == "GatedSAELoss"
    with pytest.raises(ValueError, match="Unsupported criterion"):
        U.get_criterion("nope")
    with pytest.raises(ValueError, match="Unknown SAE model name"):
        U.sae_inference_and_loss("sae_conv", m, "sae_loss", torch.randn(2, 16), None, 0.1)


def test_loss_modules_match_reference_golden(golden_dir):
    from oracle import sae_oracle as O
    from sparse_vision_b200.losses import GatedSAELoss, SparseLoss
    g = dict(np.load(os.path.join(golden_dir, "conv_mlp_cadam.npz")))
    enc = torch.from_numpy(g["step0.enc"])
    dec = O.to_tokens(torch.from_numpy(g["step0.dec"]))[0]
    x = O.to_tokens(torch.from_numpy(g["x"][0]))[0]
    rec, l1, nrmse, rmse = SparseLoss()(enc, dec, x)
    ref = g["step0.scalars"]
    np.testing.assert_allclose([rec.item(), l1.item(), nrmse.item(), rmse.item()], ref[1:5], rtol=1e-5)
    with pytest.raises(AssertionError):
        SparseLoss()(enc, dec[:, :-1], x)
    assert len(GatedSAELoss()(enc, dec, dec, x)) == 5


def test_shard_images():
    from sparse_vision_b200.parallel import shard_images
    for n, w in ((256, 8), (10, 4), (3, 8)):
        spans = [shard_images(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


_DP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from sparse_vision_b200.parallel import all_reduce_flat, global_counts, shard_images
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = shard_images(10, r, w)
g_img, g_tok = global_counts(hi - lo, 49)
assert (g_img, g_tok) == (10, 490), (g_img, g_tok)
# flat buffer: 6 SUM elements then 4 MAX elements
flat = torch.cat([torch.arange(6, dtype=torch.float32) * (r + 1), torch.tensor([1.0, -2.0, 3.0, 4.0]) * (1 if r == 0 else -1)])
all_reduce_flat(flat, 6, 4)
want_sum = torch.arange(6, dtype=torch.float32) * sum(range(1, w + 1))
assert torch.equal(flat[:6], want_sum), flat
assert torch.equal(flat[6:], torch.tensor([1.0, 2.0, 3.0, 4.0])), flat
# sharded means equal the global mean: the identity the IE / loss reductions rely on
x = torch.arange(10, dtype=torch.float64)
part = torch.tensor([x[lo:hi].sum()])
dist.all_reduce(part)
assert abs(part.item() / g_img - x.mean().item()) < 1e-12
dist.destroy_process_group()
print("ok", r)
'''


def test_data_parallel_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", str(script), ROOT],
                         capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2


def test_bench_reference_arm_prints_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-images", "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    import json
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "act-vec/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_sae_params_table_and_file_paths_match_reference(golden_dir, tmp_path):
    """SURVEY.md section 8 row a17 against outputs of the reference's own get_specific_sae_params / get_file_path
    (tests/golden/sae_params_table.json, made by oracle/gen_golden_params.py)."""
    import json
    from sparse_vision_b200 import utils as U
    g = json.load(open(os.path.join(golden_dir, "sae_params_table.json")))
    for key, want in g["layers"].items():
        layer, sae, opt = key.split("|")
        got = U.get_specific_sae_params(layer, sae, {k: v for k, v in g["model_params_temp"]}, opt)
        assert list(got) == want, key
    for c in g["paths"]:
        folder = None if c["folder"] is None else os.path.join(str(tmp_path), c["folder"])
        fp = U.get_file_path(folder, c["layer"], c["params"], c["file_name"], c["params2"])
        rel = fp if folder is None else os.path.relpath(fp, str(tmp_path))
        assert rel == c["result"], c
        if folder is not None:
            assert os.path.isdir(folder)


def test_get_specific_sae_model_loads_reference_format_checkpoint(tmp_path):
    """A checkpoint written in the reference's format (model_pipeline.py:1268-1273 keys, utils.py:151-185 file name)
    is found, loaded and frozen (utils.py:2745-2767)."""
    from sparse_vision_b200 import utils as U
    mp = {"model_name": "inceptionv1", "epochs": "0", "lr": "0.001", "bs": "512", "opt": "sgd"}
    ps, _, k, _, _ = U.get_specific_sae_params("mixed4c", "sae_mlp", mp, "constrained_adam")
    torch.manual_seed(9)
    src = U.load_model("sae_mlp", img_size=512, expansion_factor=k)
    torch.save({"epoch": 5, "model_state_dict": src.state_dict(), "optimizer_state_dict": {}, "training_step": 7},
               U.get_file_path(str(tmp_path), "mixed4c", params=ps, file_name=".pth"))
    sae, ps2, k2 = U.get_specific_sae_model("mixed4c", 512, "sae_mlp", str(tmp_path), mp, "cpu", "constrained_adam")
    assert ps2 == ps and k2 == 4 and not sae.training
    assert all(not p.requires_grad for p in sae.parameters())
    for a, b in zip(sae.state_dict().values(), src.state_dict().values()):
        assert torch.equal(a, b)


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors in _lib.py must have the size and field offsets gcc gives the structs of include/svb.h (the
    header is plain C: this also checks that it compiles as C, which is what a cgo / FFI binding would do)."""
    import ctypes
    import shutil
    import subprocess
    from sparse_vision_b200 import _lib as L
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = {"svb_acts": L.Acts, "svb_sae_params": L.SaeParams, "svb_gated_params": L.GatedParams,
             "svb_adam_state": L.AdamState, "svb_opt_config": L.OptConfig, "svb_activity_out": L.ActivityOut,
             "svb_train_out": L.TrainOut, "svb_sae_forward_out": L.SaeForwardOut,
             "svb_gated_forward_out": L.GatedForwardOut, "svb_chan_segment": L.ChanSegment,
             "svb_grad_segment": L.GradSegment}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "svb.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  printf("SVB_STATS_LEN %d\\n", (int)SVB_STATS_LEN);',
              '  printf("SVB_MAX_CHAN_SEGMENTS %d\\n", (int)SVB_MAX_CHAN_SEGMENTS);', '  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    assert int(out["SVB_STATS_LEN"]) == L.STATS_LEN
    assert int(out["SVB_MAX_CHAN_SEGMENTS"]) == L.MAX_CHAN_SEGMENTS


def test_small_host_helpers_match_reference_golden(golden_dir):
    """reshape_tensor, reshape_encoder_output_average, average_over_W_H, variance_explained, get_top_k_samples and
    CustomCrossEntropyLoss against outputs of the REAL reference functions (oracle/gen_golden_utils.py)."""
    import numpy as np
    from sparse_vision_b200 import utils as U
    g = np.load(os.path.join(golden_dir, "utils_small.npz"))
    T = lambda k: torch.from_numpy(g[k])   # noqa: E731
    t, flag = U.reshape_tensor(T("x4"))
    assert torch.equal(t, T("reshape4")) and bool(flag) == bool(g["reshape4_flag"])
    t, flag = U.reshape_tensor(T("x2"))
    assert torch.equal(t, T("reshape2")) and bool(flag) == bool(g["reshape2_flag"])
    r = U.reshape_encoder_output_average(T("avg"), 3)
    assert tuple(r.shape) == tuple(g["avg_reshaped_b3"].shape) and torch.equal(r, T("avg_reshaped_b3"))
    a, b = U.average_over_W_H(T("x4"), T("d4"))
    assert torch.allclose(a, T("avgwh_a"), rtol=0, atol=1e-7) and torch.allclose(b, T("avgwh_b"), rtol=0, atol=1e-7)
    a, b = U.average_over_W_H(T("x2"), None)
    assert torch.equal(a, T("avgwh2_a")) and b is None
    assert abs(float(U.variance_explained(T("x4"), T("d4"))) - float(g["var_expl4"])) <= 1e-6
    assert abs(float(U.variance_explained(T("x2"), T("d2"))) - float(g["var_expl2"])) <= 1e-6
    k, bs, F = 4, 5, 6
    for largest in (True, False):
        state = (torch.empty(0, F), torch.empty(0, F, dtype=torch.long), bs, torch.empty(0, F, dtype=torch.long))
        for batch in (1, 2, 3):
            tag = f"topk_{int(largest)}_{batch}"
            state = U.get_top_k_samples(state, T(tag + "_v").clone(), T(tag + "_i").clone(), T(tag + "_f").clone(),
                                        batch, largest, k)
            assert torch.equal(state[0], T(tag + "_out_v")), tag
            assert torch.equal(state[1], T(tag + "_out_i")), tag        # top-k indices: bit-exact
            assert torch.equal(state[3], T(tag + "_out_f")), tag
            assert state[2] == bs
    nll = U.CustomCrossEntropyLoss()(T("nll_probs"), T("nll_targets"))
    assert abs(float(nll) - float(g["nll"])) <= 1e-6


_IE_DP_WORKER = r'''
import collections, os, sys, torch, torch.distributed as dist
from torch import nn
sys.path.insert(0, sys.argv[1])
from oracle import sae_oracle as O
import sparse_vision_b200.ops as ops
import sparse_vision_b200.compute_ie as cie
from sparse_vision_b200.models.sae_mlp import SaeMLP
from sparse_vision_b200.parallel import shard_images

# host orchestration only: the CUDA entry points are replaced by the CPU oracle (tests may use it; the product may not)
KEYS = O.SAE_MLP_KEYS
def fake_sae_forward(x, w_enc, b_enc, w_dec, b_dec, want_pre=True, **kw):
    enc, dec, pre = O.sae_mlp_forward(dict(zip(KEYS, (w_enc, b_enc, w_dec, b_dec))), x.float())
    return enc, dec, (pre if want_pre else None)
def fake_node_ie_layer(x, grad, params, enc_avg, err_avg, x_avg, scale=None):
    f, e, n = O.node_ie_layer(dict(zip(KEYS, params)), x.float(), grad.float(), enc_avg, err_avg, x_avg)
    T = x.shape[0] * x.shape[2] * x.shape[3]
    sc = (1.0 / T) if scale is None else scale
    return f * T * sc, torch.as_tensor(e * T * sc).reshape(()), n * T * sc
def fake_image_sum(t, n_images):
    t = t.float()
    return t.sum(0) if t.dim() == 4 else t.reshape(n_images, -1, t.shape[1]).sum(0)
ops.sae_forward, ops.node_ie_layer, ops.image_sum = fake_sae_forward, fake_node_ie_layer, fake_image_sum
def fake_measure_inactive_units_device(out, k):
    dead, sparsity, freq = O.measure_inactive_units(out, k)
    return dead, torch.as_tensor(sparsity, dtype=torch.float64), freq
cie.measure_inactive_units_device = fake_measure_inactive_units_device

def build():
    torch.manual_seed(3)
    net = nn.Sequential(collections.OrderedDict(
        c1=nn.Conv2d(3, 8, 3, padding=1), r1=nn.ReLU(), p1=nn.MaxPool2d(2),
        c2=nn.Conv2d(8, 16, 3, padding=1), r2=nn.ReLU(),
        gap=nn.AdaptiveAvgPool2d(1), fl=nn.Flatten(), fc=nn.Linear(16, 5))).eval()
    names = {"r1": (8, 2), "r2": (16, 2)}
    torch.manual_seed(5)
    saes = {n: SaeMLP(c, k) for n, (c, k) in names.items()}
    mods = dict(net.named_modules())
    return cie.IE(net, {n: mods[n] for n in names}, saes, {n: k for n, (_, k) in names.items()},
                  device=torch.device("cpu")), names

batches = [(torch.randn(6 - i, 3, 8, 8, generator=torch.Generator().manual_seed(40 + i)),
            torch.randint(0, 5, (6 - i,), generator=torch.Generator().manual_seed(50 + i))) for i in range(2)]
ie, names = build()                                   # one process, whole batches, BEFORE the group exists
avg1 = ie.compute_average([x for x, _ in batches])
f1, e1, n1 = ie.compute_node_ie(batches, avg1)
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
ie, names = build()
local = []
for x, y in batches:
    lo, hi = shard_images(x.shape[0], r, w)           # 3 + 3 and 3 + 2 images
    local.append((x[lo:hi], y[lo:hi]))
avg2 = ie.compute_average([x for x, _ in local])
f2, e2, n2 = ie.compute_node_ie(local, avg2)
for n in names:
    for key in ("encoder_output_average", "sae_error_average", "original_layer_output_average"):
        assert torch.allclose(avg2[key][n], avg1[key][n], rtol=1e-5, atol=1e-6), (n, key)
    assert torch.equal(avg2["dead_units"][n], avg1["dead_units"][n]), n
    assert abs(avg2["sparsity"][n] - avg1["sparsity"][n]) < 1e-6, n
    assert torch.allclose(f2[n], f1[n], rtol=1e-4, atol=1e-7), (n, "features")
    assert abs(float(e2[n]) - float(e1[n])) <= 1e-4 * abs(float(e1[n])) + 1e-9, (n, "error")
    assert torch.allclose(n2[n], n1[n], rtol=1e-4, atol=1e-7), (n, "neurons")
dist.destroy_process_group()
print("ok", r)
'''


def test_attribution_pass_sharded_gloo_world2(tmp_path):
    """SURVEY.md 8e on the CPU: the host side of the sharded attribution pass (per-batch gradient-scale compensation,
    count / sum all-reduces, uneven shards) gives one process's result; device work is stood in for by the oracle."""
    script = tmp_path / "ie_dp_worker.py"
    script.write_text(_IE_DP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script), ROOT],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == 2


def test_tuning_keys_round_trip():
    """svb_set_tuning / svb_get_tuning (include/svb.h): host-side state only, no device needed."""
    from sparse_vision_b200 import _lib as L
    lib = L.load()
    assert lib.svb_get_tuning(99) == -1 and lib.svb_set_tuning(99, 1) != 0
    for key in range(7):
        old = lib.svb_get_tuning(key)
        assert lib.svb_set_tuning(key, old + 1) == 0 and lib.svb_get_tuning(key) == old + 1
        lib.svb_set_tuning(key, old)


def test_fuse_forward_keeps_the_module_tree_and_the_cpu_function():
    """producer.fuse_forward swaps module classes only: names, state_dict and hooks are those of torchvision's model,
    and a tensor the fused kernels do not take (CPU, fp32) goes through torchvision's own forward, bit for bit."""
    import copy
    from sparse_vision_b200.producer import fold_batchnorm, fuse_forward, synthetic_googlenet
    base = fold_batchnorm(synthetic_googlenet(seed=3, calibration_images=2, image_size=96))
    fused = fuse_forward(copy.deepcopy(base))
    assert list(fused.state_dict().keys()) == list(base.state_dict().keys())
    assert [n for n, _ in fused.named_modules()] == [n for n, _ in base.named_modules()]
    assert type(fused.inception3a).__name__ == "FusedInception" and type(fused.maxpool1).__name__ == "FusedMaxPool2d"
    fired = []
    fused.inception4c.register_forward_hook(lambda m, i, o: fired.append(tuple(o.shape)))
    x = torch.randn(2, 3, 96, 96, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        assert torch.equal(fused(x), base(x))
    assert fired == [(2, 512, 6, 6)]
    again = copy.deepcopy(fused)                      # the swapped classes survive a deepcopy (model_copy in the pipeline)
    with torch.no_grad():
        assert torch.equal(again(x), base(x))
    # the concatenated 1x1 weights of a fused block follow weights that are loaded AFTER fuse_forward
    blk = fused.inception4a
    w_before, _ = blk._merged()
    w_before = w_before.clone()
    state = {k: (v * 0.5 if v.is_floating_point() else v) for k, v in fused.state_dict().items()}
    fused.load_state_dict(state)
    w_after, b_after = blk._merged()
    assert torch.allclose(w_after, w_before * 0.5)
    assert torch.equal(b_after, torch.cat([blk.branch1.conv.bias, blk.branch2[0].conv.bias, blk.branch3[0].conv.bias]))


def test_pool_output_size_is_torchs():
    from sparse_vision_b200.ops import pool_output_size
    for n in range(1, 40):
        for k, s, p in ((3, 2, 0), (3, 1, 1), (2, 2, 0), (3, 2, 1)):
            for ceil in (False, True):
                if n + 2 * p < k:
                    continue
                want = torch.nn.functional.max_pool2d(torch.zeros(1, 1, n, n), k, s, p, ceil_mode=ceil).shape[-1]
                assert pool_output_size(n, k, s, p, ceil) == want, (n, k, s, p, ceil)


def test_to_attribution_format_layouts_on_cpu():
    """producer.to_attribution_format: channels_last + fused classes throughout by default; with nchw_tail the modules
    up to the first hooked layer only.  (On the CPU the fused classes take torchvision's forward: same function.)"""
    import copy
    from sparse_vision_b200.producer import synthetic_googlenet, to_attribution_format
    raw = synthetic_googlenet(seed=2, calibration_images=2, image_size=96)
    full = to_attribution_format(copy.deepcopy(raw), "cpu", dtype=torch.float32)
    w = full.inception5b.branch2[1].conv.weight
    assert w.is_contiguous(memory_format=torch.channels_last) and not w.is_contiguous()
    assert type(full.inception5b).__name__ == "Inception"          # fp32: nothing to fuse
    mixed = to_attribution_format(copy.deepcopy(raw), "cpu", "mixed3a", dtype=torch.bfloat16, nchw_tail=True)
    assert type(mixed.conv1).__name__ == "FusedBasicConv2d" and type(mixed.inception3a).__name__ == "FusedInception"
    assert type(mixed.inception3b).__name__ == "Inception" and type(mixed.maxpool3).__name__ == "MaxPool2d"
    assert mixed.conv3.conv.weight.is_contiguous(memory_format=torch.channels_last)
    assert mixed.inception4a.branch2[1].conv.weight.is_contiguous()
    assert [n for n, _ in mixed.named_modules()] == [n for n, _ in raw.named_modules()]
    with pytest.raises(ValueError):
        to_attribution_format(copy.deepcopy(raw), "cpu", "no_such_layer", nchw_tail=True)
    x = torch.randn(2, 3, 96, 96, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        want = to_attribution_format(copy.deepcopy(raw), "cpu", dtype=torch.float32, fold_bn=True)(x)
        got = full(x)
    assert torch.equal(got, want)


def test_producer_ops_have_no_cpu_fallback():
    """The producer kernels are CUDA-only like the rest of the product path: a CPU tensor (or a wrong format) is refused
    before any library call instead of being routed to a torch implementation."""
    from sparse_vision_b200 import ops
    x = torch.zeros(1, 8, 4, 4, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.zeros(8, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        ops.maxpool_nhwc(x, 3, 2)
    with pytest.raises(ValueError):
        ops.maxpool_nhwc_autograd(x, 3, 2)
    with pytest.raises(ValueError):
        ops.bias_relu_scatter(x, b, [(x, 0, 8)])
    with pytest.raises(ValueError):
        ops.relu_grad_gather([(x, 0, x, 0, 8)], x)
    with pytest.raises(ValueError):
        ops.conv1_stem(torch.zeros(1, 3, 224, 224, dtype=torch.bfloat16), torch.zeros(64 * 168, dtype=torch.bfloat16),
                       torch.zeros(64, dtype=torch.bfloat16))
    with pytest.raises(ValueError):
        ops.conv1_pack_weights(torch.zeros(64, 3, 7, 7, dtype=torch.bfloat16))
