#!/usr/bin/env python
"""bench.py — SAE train activation-vectors/sec (BASELINE.json metric) on configs[1]:
SaeMLP (the reference's pixels-as-tokens "Conv-SAE") on GoogLeNet inception3a-shaped activations, C=256, 28x28,
expansion 8 (F=2048), constrained_adam, lambda=5, batch 256 images = 200,704 tokens per GPU, bf16 compute.

    python bench.py --gpus N --steps K --warmup W            this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference ...                      the reference algorithm's CPU path (oracle port)

A step = one pass of ModelPipeline.hook's train branch (model_pipeline.py:380-420): forward, loss, backward,
optimizer step, activity metrics, decoder output handed back.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C_ACT, HW_SIDE, EXPANSION, LAMBDA, LR = 256, 28, 8, 5.0, 1e-3
METRIC = "sae_train_activation_vectors_per_sec"
UNIT = "act-vec/s"


def _env_int(name, default):
    return int(os.environ.get(name, default))


def _make_params(seed=0, dead_frac=0.05):
    """Reference constructor draws under torch.manual_seed(0) + a planted 5 % dead subset."""
    from sparse_vision_b200.models.sae_mlp import SaeMLP
    torch.manual_seed(seed)
    m = SaeMLP(C_ACT, EXPANSION)
    F = m.hidden_size
    idx = torch.randperm(F, generator=torch.Generator().manual_seed(1))[: int(F * dead_frac)]
    with torch.no_grad():
        m.encoder.bias[idx] = -50.0
    return m


def _synthetic_acts(n_images, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn(n_images, C_ACT, HW_SIDE, HW_SIDE, generator=g))


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML
    in-process every ~5 ms when nvidia-ml-py is importable (the timed region is only tens of milliseconds long), else
    an `nvidia-smi -lms 100` child process."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.samples, self.mask, self.max_mhz, self.run, self.live = None, [], 0, None, False, False

    def prepare(self):
        """Everything slow (NVML init / the nvidia-smi child) happens here, before the barrier that opens the timed
        region; start() only flips a flag."""
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = torch.cuda.get_device_properties(self.index).uuid
            try:
                self.dev = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                self.dev = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
            self.nvml, self.run = pynvml, True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.run:
            if self.live:
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM)))
                    self.mask |= int(reasons(self.dev))
                except Exception:
                    pass
            time.sleep(0.005)

    def start(self):
        self.live = True

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml:
            self.run = False
            self.t.join(timeout=1.0)
            sm = sorted(self.samples)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(k for k, b in self.BITS.items() if self.mask & b), "samples": len(sm),
                    "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [q.strip() for q in ln.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def _bind_to_gpu_numa_node(local):
    """Pins this rank's CPU threads to the NUMA node its GPU hangs off, BEFORE the pinned host buffers are allocated
    (first-touch places them on that node): with 8 ranks streaming 55 GB/s each from host memory, buffers that all sit
    on one socket halve the end-to-end rate.  Returns the node or None (no NVML / single node / no sysfs)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = torch.cuda.get_device_properties(local).uuid
        try:
            dev = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
        except Exception:
            dev = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(dev).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]                               # sysfs uses a 4-digit PCI domain
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"bf16_sustained": p.get("bf16_tflops_sustained"), "bf16_burst": p.get("bf16_tflops"),
                "hbm": p.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def _traffic():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return {k: v for k, v in json.load(fh).items() if not k.startswith("_")}
    return {}


CONFIG_WORKLOAD = ("configs[1]: SaeMLP (pixels-as-tokens) C=256 28x28 k=8 F=2048 constrained_adam lambda=5, 256 images = "
                   "200704 tokens per GPU per step, GoogLeNet inception3a-shaped activations")


def cpu_reference_throughput(n_images, steps, warmup, threads=None):
    """The reference algorithm on the host cores: oracle/sae_oracle.py (a restatement pinned to the real reference by
    tests/golden) running the same step on a bounded sample of the same workload (same C, F, HW, optimizer)."""
    from oracle import sae_oracle as O
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    p = O.init_sae_mlp(C_ACT, EXPANSION)
    st = O.new_adam_state(p, O.SAE_MLP_KEYS)
    x = _synthetic_acts(n_images, 1234)
    tokens = n_images * HW_SIDE * HW_SIDE
    for _ in range(warmup):
        O.train_step("sae_mlp", p, st, x, LAMBDA, "constrained_adam", LR, EXPANSION)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step("sae_mlp", p, st, x, LAMBDA, "constrained_adam", LR, EXPANSION)
    dt = (time.perf_counter() - t0) / steps
    return tokens / dt, dt * 1e3, torch.get_num_threads()


def cpu_reference_pipeline(n_images, steps, warmup):
    """The reference's whole train batch on the host cores (model_pipeline.py:603-708): frozen GoogLeNet forward whose
    inception3a hook runs the SAE training step (oracle) and hands the reconstruction back, the unhooked copy's forward,
    KL divergence / same-classification.  Returns (tokens/s, ms per batch)."""
    import copy
    from oracle import sae_oracle as O
    from sparse_vision_b200.producer import synthetic_googlenet
    base = synthetic_googlenet(seed=0)
    ref_copy = copy.deepcopy(base)
    torch.manual_seed(0)
    p = O.init_sae_mlp(C_ACT, EXPANSION)
    st = O.new_adam_state(p, O.SAE_MLP_KEYS)

    def hook(_m, _i, out):
        with torch.enable_grad():                       # model_pipeline.py:380
            r = O.train_step("sae_mlp", p, st, out.detach(), LAMBDA, "constrained_adam", LR, EXPANSION)
        return r["dec"]

    base.inception3a.register_forward_hook(hook)
    x = torch.randn(n_images, 3, 224, 224, generator=torch.Generator().manual_seed(3))

    def batch():
        with torch.no_grad():
            out = base(x)
            orig = ref_copy(x)
            lp_o, lp_m = torch.log_softmax(orig, 1), torch.log_softmax(out, 1)
            kld = torch.nn.functional.kl_div(lp_o, lp_m, reduction="sum", log_target=True).item() / n_images
            same = (orig.argmax(1) == out.argmax(1)).sum().item() / n_images
        return kld, same

    for _ in range(warmup):
        batch()
    t0 = time.perf_counter()
    for _ in range(steps):
        batch()
    dt = (time.perf_counter() - t0) / steps
    return n_images * HW_SIDE * HW_SIDE / dt, dt * 1e3


def gpu_eager_reference(dev, x_bf16, iters=5):
    """SURVEY.md 8(d) / BASELINE.md 3, "the real kernel to beat": the reference algorithm (oracle/sae_oracle.py, i.e. the
    reference modules' own PyTorch ops) run unchanged in eager mode ON THIS GPU at the full configs[1] size -- fp32 as the
    reference ships it, with TF32 allowed, and under bf16 autocast.  Every FLOP there is a cuBLAS call."""
    from oracle import sae_oracle as O
    out = {}
    x = x_bf16.float()
    T = x.shape[0] * x.shape[2] * x.shape[3]
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for mode in ("fp32", "tf32", "bf16_autocast"):
            torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
            torch.manual_seed(0)
            p = {k: v.to(dev) for k, v in O.init_sae_mlp(C_ACT, EXPANSION).items()}
            st = O.new_adam_state(p, O.SAE_MLP_KEYS)

            def step():
                if mode == "bf16_autocast":
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        return O.train_step("sae_mlp", p, st, x, LAMBDA, "constrained_adam", LR, EXPANSION)
                return O.train_step("sae_mlp", p, st, x, LAMBDA, "constrained_adam", LR, EXPANSION)

            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                r = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            out[mode] = {"ms_per_step": ms, "act_vec_per_s": T / (ms * 1e-3), "loss": r["loss"]}
            del p, st, r
            torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    out["what"] = ("oracle.train_step (the reference's PyTorch ops: 6 cuBLAS GEMMs, ~30 elementwise / reduction kernels, "
                   "6+ host syncs) on cuda, same 200704-token batch")
    return out


def gated_section(dev, peaks, n_images=256, iters=10, world=1):
    """configs[2]: GatedSae on inception4c-shaped activations (C=512, 14x14, expansion 16 -> F=8192), one fused training
    step per call on bf16 NCHW activations resident in HBM; 12*C*F FLOP per token.  world > 1:
    data parallel like the main workload (n_images per GPU, peer-memory all-reduce of the flat gradient buffer; every
    rank calls this, times are the max over ranks)."""
    from sparse_vision_b200 import _lib as L, ops
    from sparse_vision_b200.models.gated_sae import GatedSae
    from sparse_vision_b200.parallel import DataParallelStep
    import ctypes as C
    import torch.distributed as dist
    Cc, side, k = 512, 14, 16
    F, T = Cc * k, n_images * side * side
    rank = dist.get_rank() if world > 1 else 0
    torch.manual_seed(0)
    model = GatedSae(Cc, k)
    params = [p.detach().clone().to(dev) for p in model.param_list()]
    ms_ = [torch.zeros_like(p) for p in params]
    vs_ = [torch.zeros_like(p) for p in params]
    g = torch.Generator(device="cpu").manual_seed(21 + rank)
    xs = [torch.relu(torch.randn(n_images, Cc, side, side, generator=g)).to(torch.bfloat16).to(dev) for _ in range(2)]
    dp = DataParallelStep("gated_sae") if world > 1 else None

    def step(i, no):
        if dp is None:
            return ops.gated_train_step(xs[i % 2], params, ms_, vs_, no, LR, 0.1, k, optimizer="constrained_adam")
        return dp.step(xs[i % 2], params, ms_, vs_, no, LR, 0.1, k, "constrained_adam", (0.9, 0.999),
                       n_images * world, T * world, want_dec=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(3):
        res = step(i, i + 1)
    barrier()
    lib, h = L.load(), L.handle(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        res = step(i, i + 4)
    e1.record()
    barrier()
    # phases from a second pass with the per-phase events on (they cost a few percent; see the main workload)
    L.check(lib.svb_profile_enable(h, 1), "svb_profile_enable")
    for i in range(iters):
        res = step(i, i + 4 + iters)
    barrier()
    phases = _read_phases(lib, h)
    L.check(lib.svb_profile_enable(h, 0), "svb_profile_enable")
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    tf = 12.0 * Cc * F * T / (ms * 1e-3) / 1e12          # per GPU
    sc = res.scalars()
    strong = None
    if dp is not None and n_images % world == 0:
        # strong scaling (cfg3 weak AND strong): the same 256-image batch sharded across the ranks
        ns = n_images // world
        xss = [x[:ns].contiguous() for x in xs]
        no = [4 + 2 * iters]

        def sstep(i):
            no[0] += 1
            return dp.step(xss[i % 2], params, ms_, vs_, no[0], LR, 0.1, k, "constrained_adam", (0.9, 0.999), n_images, T,
                           want_dec=True)
        for i in range(3):
            sstep(i)
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for i in range(iters):
            sstep(i)
        q1.record()
        barrier()
        ts = torch.tensor([q0.elapsed_time(q1) / iters], device=dev, dtype=torch.float64)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        strong = {"scaling": "strong", "global_batch_images": n_images, "images_per_gpu": ns, "ms_per_step": float(ts[0]),
                  "act_vec_per_s": T / (float(ts[0]) * 1e-3)}
    if dp is not None:
        dp.check()
    return {"workload": f"configs[2]: GatedSae C=512 14x14 k=16 F=8192, 256 images = 50176 tokens per GPU, "
                        f"constrained_adam, bf16 NCHW, dp{world}",
            "n_gpus": world, "ms_per_step": ms, "act_vec_per_s": world * T / (ms * 1e-3), "algorithmic_tflops_per_gpu": tf,
            "frac_of_burst_peak": tf / peaks["bf16_burst"] if peaks["bf16_burst"] else None,
            "frac_of_sustained_peak": tf / peaks["bf16_sustained"] if peaks["bf16_sustained"] else None,
            "phases_ms": phases, "strong_scaling": strong,
            "final_step_stats": {kk: sc[kk] for kk in ("loss", "rec", "l1", "aux")}}


def _read_phases(lib, h):
    import ctypes as C
    from sparse_vision_b200 import _lib as L
    phase_ms = (C.c_float * 16)()
    n_ph, n_st = C.c_int32(), C.c_int32()
    L.check(lib.svb_profile_read(h, 16, phase_ms, C.byref(n_ph), C.byref(n_st)), "svb_profile_read")
    ph = {lib.svb_profile_phase_name(i).decode(): float(phase_ms[i]) for i in range(n_ph.value)}
    if lib.svb_last_step_flags(h) & 1 and "dE_gemm" in ph:
        # SVB_STEP_FUSED_BWD: one kernel did the dE GEMM, the ReLU mask and the dW_enc GEMM (two GEMMs of work)
        out = {}
        for k, v in ph.items():
            if k == "dE_gemm":
                out[FUSED_BWD_PHASE] = v + ph.get("dWenc_gemm", 0.0)
            elif k != "dWenc_gemm":
                out[k] = v
        return out
    return ph


FUSED_BWD_PHASE = "dE+dWenc_fused_gemm"


def _gemms_in_phase(name):
    return 2 if name == FUSED_BWD_PHASE else 1


def ie_section(dev, peaks, n_images=64, iters=20, world=1):
    """IE images/sec (second half of the BASELINE.json metric) at cfg5 / mixed3a: F=2048 SAE features on 28x28 maps.
    Times (a) the stand-alone compute_ie_channel_wise reduction (utils.py:2606-2637) on fp32 and bf16 [T,F] inputs —
    HBM roofline, algorithmic bytes 2*T*F*s + F*HW*4 + F*4 — and (b) the fused per-layer node-IE
    (encoder GEMM, decoder GEMM, g W_dec GEMM and the three reductions) in images/s.  world > 1: the pass shards by
    image with no data-path collective, so every rank times its own n_images and the job rate is world x n_images over
    the slowest rank's time."""
    import torch.distributed as dist
    from sparse_vision_b200 import ops
    F, HW, Cc = C_ACT * EXPANSION, HW_SIDE * HW_SIDE, C_ACT
    T = n_images * HW
    rank = dist.get_rank() if world > 1 else 0
    g = torch.Generator(device="cpu").manual_seed(7 + rank)
    out = {}

    def job_ms(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    avg = torch.rand(F, HW_SIDE, HW_SIDE, generator=g).to(dev)
    for name, dt, sz in (("f32", torch.float32, 4), ("bf16", torch.bfloat16, 2)):
        # two rotating input sets so that every timed launch reads from HBM, not L2 (2 x 2 x T*F*sz > 126 MB)
        sets = [(torch.rand(T, F, device=dev, dtype=torch.float32).to(dt), torch.randn(T, F, device=dev).to(dt))
                for _ in range(2)]
        for i in range(3):
            ops.ie_channelwise(*[sets[i % 2][0], avg, sets[i % 2][1]], n_images)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            ops.ie_channelwise(sets[i % 2][0], avg, sets[i % 2][1], n_images)
        e1.record()
        torch.cuda.synchronize()
        ms = job_ms(e0.elapsed_time(e1) / iters)
        nbytes = 2.0 * T * F * sz + F * HW * 4 + F * 4
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "algorithmic_bytes": nbytes, "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s",
                     "frac": gbs / peaks["hbm"] if peaks["hbm"] else None, "per": "GPU",
                     "images_per_s": world * n_images / (ms * 1e-3)}
        del sets
    # fused node-IE of one layer on bf16 NCHW activations / gradients
    model = _make_params()
    params = [p.detach().clone().to(dev) for p in model.param_list()]
    x = [_synthetic_acts(n_images, 900 + i + 10 * rank).to(torch.bfloat16).to(dev) for i in range(2)]
    gr = [torch.randn(n_images, Cc, HW_SIDE, HW_SIDE, generator=g).to(torch.bfloat16).to(dev) for _ in range(2)]
    err_avg = torch.zeros(Cc, HW_SIDE, HW_SIDE, device=dev)
    x_avg = x[0].float().mean(0)
    for i in range(3):
        ops.node_ie_layer(x[i % 2], gr[i % 2], params, avg, err_avg, x_avg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.node_ie_layer(x[i % 2], gr[i % 2], params, avg, err_avg, x_avg)
    e1.record()
    torch.cuda.synchronize()
    ms = job_ms(e0.elapsed_time(e1) / iters)
    out["node_ie_layer"] = {"ms": ms, "images_per_s": world * n_images / (ms * 1e-3),
                            "what": "svb_node_ie_layer, one layer: ONE fused kernel keeps a = relu(x W_enc^T + b) and G = g W_dec in TMEM and reduces |G (avg - a)| and sum_f a G on the spot (no [T,F] tensor is written, no decoder GEMM), + the neuron / error passes over [T,C]"}
    out["n_gpus"] = world
    out["config"] = {"workload": "configs[4] / mixed3a: C=256, F=2048, 28x28, %d images per call per GPU" % n_images,
                     "peak_source": peaks["source"] + ", hbm copy"}
    return out


IE_LAYERS = {"mixed3a": 8, "mixed4c": 4, "mixed5b": 4}   # cfg5: three GoogLeNet layers, expansion per utils.py:2671-2724


def ie_pipeline_section(dev, base, n_images=64, n_batches=3, world=1, graph=True):
    """configs[4] end to end: IE.compute_average then IE.compute_node_ie (compute_ie.py:95-226, :365-472) over three
    GoogLeNet layers on 224x224 images -- ONE frozen forward + backward of the base model per batch (cuDNN) and, per
    layer, the SAE encoder / decoder / g W_dec GEMMs and the three reductions of libsvb.  Images are sharded across the
    ranks; the only exchange is the final all-reduce of the per-layer sums."""
    import torch.distributed as dist
    from sparse_vision_b200.compute_ie import IE
    from sparse_vision_b200.models.sae_mlp import SaeMLP
    from sparse_vision_b200.producer import GOOGLENET_LAYERS, hooked_layers
    rank = dist.get_rank() if world > 1 else 0
    saes = {}
    for j, (n, k) in enumerate(IE_LAYERS.items()):
        torch.manual_seed(5 + j)
        saes[n] = SaeMLP(GOOGLENET_LAYERS[n][1], k).to(dev)
    dt = next(base.parameters()).dtype
    g = torch.Generator().manual_seed(40 + rank)
    fmt = (torch.channels_last if base.conv1.conv.weight.is_contiguous(memory_format=torch.channels_last)
           and not base.conv1.conv.weight.is_contiguous() else torch.contiguous_format)
    batches = [(torch.randn(n_images, 3, 224, 224, generator=g).to(dev, dt).contiguous(memory_format=fmt),
                torch.randint(0, 1000, (n_images,), generator=g).to(dev)) for _ in range(n_batches)]
    ie = IE(base, hooked_layers(base, list(IE_LAYERS)), saes, dict(IE_LAYERS), device=dev, cuda_graph=graph)

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return r, float(t[0])

    # The reference's order: the averages over all batches first, then the attribution passes with them.
    timed(lambda: ie.compute_average([b[0] for b in batches]))          # warm-up (cuDNN heuristics, arena, allocator)
    timed(lambda: ie.compute_average([b[0] for b in batches]))
    avg, ms_avg = timed(lambda: ie.compute_average([b[0] for b in batches]))
    # Untimed attribution passes: on a fresh box the first one pays cuDNN's lazy kernel loading for the backward of every
    # layer shape, the allocator's first cudaMallocs and (cuda_graph) the capture of the batch graph -- torch.cuda.graph
    # also empties the caching allocator, which is why the averages are timed BEFORE the first capture (timed after it,
    # compute_average paid 100 ms per batch of fresh cudaMallocs).  Measured: 100-260 ms per batch for the first pass,
    # steady from the second on.
    warm = [timed(lambda: ie.compute_node_ie(batches, avg))[1] for _ in range(3)]
    ms_ie_first = warm[0]
    passes = [timed(lambda: ie.compute_node_ie(batches, avg)) for _ in range(3)]
    (feat, err, neur), _ = passes[-1]
    ms_ie = sum(p[1] for p in passes) / len(passes)      # mean of three timed passes
    n_job = world * n_images * n_batches
    top = {n: [int(i) for i in torch.topk(feat[n], 5).indices.tolist()] for n in feat}
    return {"workload": f"configs[4]: node IE over {list(IE_LAYERS)} of GoogLeNet ({dt}), {n_images} images x {n_batches} "
                        f"batches per GPU, 224x224", "n_gpus": world,
            "base_model": ("NCHW, torchvision's eager forward" if fmt != torch.channels_last else
                           "channels_last; forward-only on the fused producer kernels up to the first hooked layer; " +
                           ("cuDNN autograd + libsvb differentiable max-pool behind it"
                            if base.inception5b.branch2[1].conv.weight.is_contiguous(memory_format=torch.channels_last)
                            and not base.inception5b.branch2[1].conv.weight.is_contiguous() else "NCHW behind it")),
            "compute_average_images_per_s": n_job / (ms_avg * 1e-3), "compute_node_ie_images_per_s": n_job / (ms_ie * 1e-3),
            "ms_per_batch_node_ie": ms_ie / n_batches, "ms_per_batch_average": ms_avg / n_batches,
            "ms_per_batch_node_ie_passes": [p[1] / n_batches for p in passes],
            "ms_per_batch_node_ie_first_pass": ms_ie_first / n_batches,
            "ms_per_batch_node_ie_untimed_passes": [w / n_batches for w in warm],
            "cuda_graph": bool(ie._graphs),
            "top5_features": top}


def dp_parity_check(dev, world, rank):
    """N > 1 (the driver's GPU-test box has one GPU): two data-parallel steps of each SAE kind on a small shape, every
    rank on its image shard, against ONE rank stepping on the whole batch with the same library; replicas must stay
    bit-identical across ranks.  Runs before the timed region; returns a dict for the JSON line."""
    import torch.distributed as dist
    from sparse_vision_b200 import ops
    from sparse_vision_b200.models.gated_sae import GatedSae
    from sparse_vision_b200.models.sae_mlp import SaeMLP
    from sparse_vision_b200.parallel import DataParallelStep, shard_images
    worst, identical, ok = 0.0, True, True
    for kind, cls in (("sae_mlp", SaeMLP), ("gated_sae", GatedSae)):
        Cc, k, B, H = 64, 4, 3 * world, 7
        torch.manual_seed(0)
        m = cls(Cc, k)
        init = [p.detach().clone() for p in m.param_list()]
        x = torch.relu(torch.randn(B, Cc, H, H, generator=torch.Generator().manual_seed(11))).bfloat16()
        lo, hi = shard_images(B, rank, world)
        params = [p.clone().to(dev) for p in init]
        ms_ = [torch.zeros_like(q) for q in params]
        vs_ = [torch.zeros_like(q) for q in params]
        dp = DataParallelStep(kind)
        for step in (1, 2):
            res = dp.step(x[lo:hi].to(dev), params, ms_, vs_, step, 1e-3, 0.5, k, "constrained_adam", (0.9, 0.999), B, B * H * H)
        dp.check()
        ref_params = [p.clone().to(dev) for p in init]
        rm = [torch.zeros_like(q) for q in ref_params]
        rv = [torch.zeros_like(q) for q in ref_params]
        fn = ops.sae_train_step if kind == "sae_mlp" else ops.gated_train_step
        for step in (1, 2):
            ref = fn(x.to(dev), ref_params, rm, rv, step, 1e-3, 0.5, k, optimizer="constrained_adam")
        got, want = res.scalars(), ref.scalars()
        for key in ("loss", "rec", "l1", "var_expl", "sparsity", "n_dead"):
            ok &= abs(got[key] - want[key]) <= 1e-4 * max(abs(want[key]), 1e-3)
        ok &= bool(torch.equal(res.dead, ref.dead))
        for a, b in zip(params, ref_params):
            d = (a - b).abs()
            worst = max(worst, d.max().item())
            ok &= d.max().item() <= 4.2e-3 and d.mean().item() <= 2e-5   # a sign flip of a ~0 gradient: 2*lr per step
            other = a.clone()
            dist.broadcast(other, src=0)
            identical &= bool(torch.equal(other, a))
    flag = torch.tensor([1 if (ok and identical) else 0], device=dev, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"ok": bool(flag.item()), "replicas_bit_identical": identical, "max_param_diff_vs_single_rank": worst,
            "what": "2 DP steps (SaeMLP + GatedSae, C=64 k=4 7x7, 3 images per rank) vs one rank on the whole batch"}


def run_reference(args):
    rank = _env_int("RANK", 0)
    if rank != 0:
        return 0
    n_img = args.cpu_images
    warm = max(args.warmup, 1)
    value, ms, cores = cpu_reference_throughput(n_img, args.steps, warm, os.cpu_count())
    e2e_value, e2e_ms = cpu_reference_pipeline(n_img, max(args.steps // 4, 2), 1)
    sample = (f"{n_img} images = {n_img * HW_SIDE * HW_SIDE} tokens per step of the same workload (fp32, torch CPU, "
              f"{cores} threads); per-token cost is batch-independent")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CONFIG_WORKLOAD, "sample_images_per_step": n_img,
                   "sample_tokens_per_step": n_img * HW_SIDE * HW_SIDE},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "ms_per_step": e2e_ms,
                "note": "ModelPipeline-equivalent train batch on the CPU: GoogLeNet forward with the SAE training step in "
                        "the inception3a hook + unhooked copy forward + KLD, same sample size"},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


def run_svb(args):
    import copy
    import math
    import torch.distributed as dist
    from sparse_vision_b200 import _lib as L, ops
    from sparse_vision_b200.model_pipeline import ModelPipeline
    from sparse_vision_b200.parallel import DataParallelStep
    from sparse_vision_b200.producer import synthetic_googlenet, to_producer_format

    world, rank, local = _env_int("WORLD_SIZE", 1), _env_int("RANK", 0), _env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    skip = set(args.skip.split(",")) if args.skip else set()
    B = args.images
    T = B * HW_SIDE * HW_SIDE
    F = C_ACT * EXPANSION
    peaks = _peaks()
    dp_parity = None
    if world > 1 and "dp_parity" not in skip:
        try:
            dp_parity = dp_parity_check(dev, world, rank)
        except Exception as exc:
            dp_parity = {"ok": False, "error": f"{type(exc).__name__}: {exc}"}
    model = _make_params()
    params = [p.detach().clone().to(dev) for p in model.param_list()]
    ms_ = [torch.zeros_like(p) for p in params]
    vs_ = [torch.zeros_like(p) for p in params]
    numa_node = _bind_to_gpu_numa_node(local) if world > 1 and not args.no_numa else None
    # two distinct resident batches (2 x 103 MB > 126 MB L2), bf16 NCHW as the base model would emit them
    xdev = [_synthetic_acts(B, 1234 + 17 * rank + i).to(torch.bfloat16).to(dev) for i in range(2)]
    if args.acts_format == "channels_last":     # the same values as a channels_last producer emits them: NHWC = tokens
        xdev = [t.contiguous(memory_format=torch.channels_last) for t in xdev]
    lib, h = L.load(), L.handle(dev)
    dp = DataParallelStep("sae_mlp") if world > 1 else None
    g_images, g_tokens = B * world, T * world
    step_no = [0]

    def one_step(x):
        step_no[0] += 1
        if dp is None:
            return ops.sae_train_step(x, params, ms_, vs_, step_no[0], LR, LAMBDA, EXPANSION,
                                      optimizer="constrained_adam", betas=(0.9, 0.999), want_dec=True)
        return dp.step(x, params, ms_, vs_, step_no[0], LR, LAMBDA, EXPANSION, "constrained_adam", (0.9, 0.999),
                       g_images, g_tokens, want_dec=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def job_max(*vals):
        t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    # the same step on the OTHER memory format of the same activations (NCHW needs the pack pass; channels_last is read
    # in place), for the record
    other_fmt = None
    if "other_format" not in skip:
        alt = [t.contiguous() if args.acts_format == "channels_last" else t.contiguous(memory_format=torch.channels_last)
               for t in xdev]
        for i in range(3):
            one_step(alt[i % 2])
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(args.steps):
            one_step(alt[i % 2])
        a1.record()
        barrier()
        (ms_alt,) = job_max(a0.elapsed_time(a1))
        other_fmt = {"format": "nchw" if args.acts_format == "channels_last" else "channels_last",
                     "ms_per_step": ms_alt / args.steps, "value": g_tokens / (ms_alt / args.steps * 1e-3),
                     "note": "timed BEFORE the main region (3 warm-up steps)"}
        del alt

    for i in range(args.warmup):
        res = one_step(xdev[i % 2])
    barrier()

    # ---------------------------------------------------------------- timed region: inputs resident in HBM
    # The per-phase CUDA events of the library's profiler cost 0.04-0.1 ms per step (a record between every two kernels
    # of the step; tools/prof_overhead.py: 0.98 ms without, 1.03 ms with them), so the headline region runs WITHOUT them
    # and the phases come from a second pass of the same K steps on the same inputs right after it.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.prepare()
    launches0 = lib.svb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record()
    for i in range(args.steps):
        res = one_step(xdev[i % 2])
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.svb_launch_count() - launches0
    last_stats = res.scalars()
    (ms_total,) = job_max(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    value = g_tokens / (ms_step * 1e-3)
    # the same K steps again with the per-phase events on (roofline / phases_ms)
    L.check(lib.svb_profile_enable(h, 1), "svb_profile_enable")
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    pv0.record()
    for i in range(args.steps):
        res = one_step(xdev[i % 2])
    pv1.record()
    barrier()
    phases = _read_phases(lib, h)
    L.check(lib.svb_profile_enable(h, 0), "svb_profile_enable")
    (ms_prof_total,) = job_max(pv0.elapsed_time(pv1))
    ms_step_profiled = ms_prof_total / args.steps

    # ---------------------------------------------------------------- strong scaling (N > 1): the SAME global batch of B images
    # split across the ranks (report both); small shards under-fill the GPUs, so this is the harder number
    strong = None
    if world > 1 and "strong" not in skip and B % world == 0:
        Bs = B // world
        xs_strong = [t[:Bs].contiguous(memory_format=torch.channels_last) if args.acts_format == "channels_last"
                     else t[:Bs].contiguous() for t in xdev]
        gi, gt = B, T

        def strong_step(x):
            step_no[0] += 1
            return dp.step(x, params, ms_, vs_, step_no[0], LR, LAMBDA, EXPANSION, "constrained_adam", (0.9, 0.999), gi, gt,
                           want_dec=True)
        for i in range(max(args.warmup, 3)):
            strong_step(xs_strong[i % 2])
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for i in range(args.steps):
            strong_step(xs_strong[i % 2])
        q1.record()
        barrier()
        (ms_strong,) = job_max(q0.elapsed_time(q1))
        strong = {"scaling": "strong", "global_batch_images": B, "images_per_gpu": Bs, "ms_per_step": ms_strong / args.steps,
                  "value": T / (ms_strong / args.steps * 1e-3), "unit": UNIT,
                  "note": "same global batch as the N = 1 run, sharded by image; value = global tokens / max-over-ranks step time"}
        del xs_strong

    # ---------------------------------------------------------------- sustained leg: seconds of back-to-back steps
    sustained = None
    if "sustained" not in skip and args.sustain_s > 0:
        n_sus = max(args.steps, int(math.ceil(args.sustain_s * 1e3 / ms_step)))
        sampler2 = ClockSampler(local)
        if rank == 0:
            sampler2.prepare()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler2.start()
        s0.record()
        for i in range(n_sus):
            res = one_step(xdev[i % 2])
        s1.record()
        barrier()
        clocks2 = sampler2.stop() if rank == 0 else None
        (ms_sus,) = job_max(s0.elapsed_time(s1))
        # 128 more steps, still in the power-capped clock state, with the per-phase events on
        L.check(lib.svb_profile_enable(h, 1), "svb_profile_enable")
        for i in range(128):
            res = one_step(xdev[i % 2])
        barrier()
        phases_sus = _read_phases(lib, h)
        L.check(lib.svb_profile_enable(h, 0), "svb_profile_enable")
        sustained = {"steps": n_sus, "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / n_sus,
                     "value": g_tokens / (ms_sus / n_sus * 1e-3), "clocks": clocks2, "phases_ms_last_128_steps": phases_sus,
                     "phases_note": "128 steps run right after the leg with the per-phase events on"}
    if dp is not None:
        dp.check()

    # ---------------------------------------------------------------- e2e: host images in -> ModelPipeline.hook -> stats out
    e2e = None
    base = None
    if "e2e" not in skip:
        fuse_producer = args.channels_last and not args.no_fold_bn and not args.no_fuse_producer
        base = to_producer_format(synthetic_googlenet(seed=0), dev, torch.bfloat16, channels_last=args.channels_last,
                                  fold_bn=not args.no_fold_bn, fuse=fuse_producer)
        base_copy = copy.deepcopy(base) if args.two_pass else None
        sae = _make_params().to(dev)
        pipe = ModelPipeline(base, sae, "sae_mlp", "inception3a", "constrained_adam", LR, LAMBDA, EXPANSION,
                             data_parallel=world > 1, global_batch_images=g_images if world > 1 else None,
                             model_copy=base_copy, compare_in_one_pass=not args.two_pass,
                             cuda_graph=not args.no_graph)
        pipe.register_hooks(train_sae=True)
        gi = torch.Generator().manual_seed(77 + rank)
        host = [torch.randn(B, 3, 224, 224, generator=gi).to(torch.bfloat16).pin_memory() for _ in range(2)]
        tgt = torch.randint(0, 1000, (B,), generator=gi).to(dev)
        copy_stream = torch.cuda.Stream(device=dev)
        stage = [torch.empty(host[0].shape, device=dev, dtype=torch.bfloat16) for _ in range(2)]
        if args.channels_last:
            stage = [s.contiguous(memory_format=torch.channels_last) for s in stage]
            host = [hh.contiguous(memory_format=torch.channels_last).pin_memory() for hh in host]
        stats_host = torch.empty(L.STATS_LEN + 3, dtype=torch.float32).pin_memory()
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        main = torch.cuda.current_stream()
        for i in range(5):                                    # warm-up: cuDNN heuristics, arena growth, graph capture
            stage[i % 2].copy_(host[i % 2], non_blocking=True)
            pipe.train_batch(stage[i % 2], targets=tgt)
        barrier()
        launches_e0 = lib.svb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        copy_stream.wait_event(e0)
        n_e2e = args.steps
        with torch.cuda.stream(copy_stream):
            stage[0].copy_(host[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(n_e2e):
            cur, nxt = i % 2, (i + 1) % 2
            if i + 1 < n_e2e:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(freed[nxt])
                    stage[nxt].copy_(host[nxt], non_blocking=True)     # H2D of batch i+1 overlaps the compute of batch i
                    ready[nxt].record(copy_stream)
            main.wait_event(ready[cur])
            pipe.train_batch(stage[cur], targets=tgt)
            freed[cur].record(main)
            stats_host[:L.STATS_LEN].copy_(pipe._last.stats, non_blocking=True)       # D2H of the batch's results
            stats_host[L.STATS_LEN:].copy_(pipe.batch_model_stats, non_blocking=True)
        e1.record()
        barrier()
        (ms_e2e,) = job_max(e0.elapsed_time(e1))
        replicas_identical = None
        if pipe.dp is not None:
            pipe.dp.check()
            same = True          # every rank stepped the same SAE: the replicas must be bit-identical after the timed batches
            for prm in sae.param_list():
                other = prm.detach().clone()
                dist.broadcast(other, src=0)
                same &= bool(torch.equal(other, prm.detach()))
            flag = torch.tensor([1 if same else 0], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            replicas_identical = bool(flag.item())
        pipe.remove_hooks()
        sh = stats_host.tolist()
        e2e = {"value": g_tokens / (ms_e2e / n_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": host[0].numel() * 2 * world, "d2h_bytes_per_step": (L.STATS_LEN + 3) * 4 * world,
               "ms_per_step": ms_e2e / n_e2e, "images_per_s": g_images / (ms_e2e / n_e2e * 1e-3),
               "svb_launches_per_step": (pipe.graph_svb_launches if pipe._graph is not None
                                         else (lib.svb_launch_count() - launches_e0) / n_e2e),
               "last_batch": {"loss": sh[0], "rec": sh[1], "kld": sh[L.STATS_LEN], "same_classification": sh[L.STATS_LEN + 1],
                              "loss_diff": sh[L.STATS_LEN + 2]},
               "numa_node_rank0": numa_node,
               "cuda_graph": pipe._graph is not None, "replicas_bit_identical": replicas_identical,
               "producer": {"channels_last": bool(args.channels_last), "batchnorm_folded": not args.no_fold_bn,
                            "fused_forward": bool(fuse_producer),
                            "original_model": "second forward of an unhooked copy" if args.two_pass else
                            "same pass: the hook hands [reconstruction; original activation] (2B) to the rest of the net"},
               "note": "ModelPipeline.train_batch on pinned host images (bf16 3x224x224): H2D (overlapped on a copy stream) "
                       "-> frozen GoogLeNet forward (bf16, cuDNN) whose inception3a hook runs the fused SAE training step "
                       "and hands the reconstruction back -> rest of the network for the modified AND the original "
                       "activations -> KLD / same-classification / loss difference -> stats + comparison scalars D2H"}
        del pipe, base_copy, host, stage

    # ---------------------------------------------------------------- side sections (collective at N > 1: every rank runs them)
    def guarded(fn):
        try:
            return fn()
        except Exception as exc:       # a failure of a side section must never cost the main line
            return {"error": f"{type(exc).__name__}: {exc}"}

    gated = guarded(lambda: gated_section(dev, peaks, world=world)) if "gated" not in skip else None
    ie = guarded(lambda: ie_section(dev, peaks, world=world)) if "ie" not in skip else None
    ie_pipe = None
    if "ie_pipeline" not in skip:
        # The attribution pass needs the BACKWARD of the base model from the loss down to the first hooked layer
        # (IE._forward_collect cuts the graph there).  bf16 channels_last throughout: forward-only in front of the leaf on
        # the fused producer kernels; behind it cuDNN through torch autograd and libsvb's differentiable max-pool.
        # --ie-nchw-tail / --ie-nchw are the earlier arrangements (NCHW behind the leaf / everywhere).
        base = None
        torch.cuda.empty_cache()
        from sparse_vision_b200.producer import to_attribution_format
        if args.ie_nchw:
            ie_base = to_producer_format(synthetic_googlenet(seed=0), dev, torch.bfloat16, channels_last=False, fold_bn=True)
        else:
            ie_base = to_attribution_format(synthetic_googlenet(seed=0), dev, next(iter(IE_LAYERS)), torch.bfloat16,
                                            nchw_tail=args.ie_nchw_tail)
        ie_pipe = guarded(lambda: ie_pipeline_section(dev, ie_base, world=world, graph=not args.no_graph))

    if rank == 0:
        gemm_phases = {k: v for k, v in phases.items() if k.endswith("_gemm")}
        dom = max(gemm_phases, key=gemm_phases.get) if gemm_phases else None
        flops_per_gemm = 2.0 * T * C_ACT * F * (_gemms_in_phase(dom) if dom else 1)   # algorithmic FLOP of the dominant launch
        step_flops = 10.0 * C_ACT * F * T
        achieved = flops_per_gemm / (gemm_phases[dom] * 1e-3) / 1e12 if dom else None
        traffic = _traffic()
        roofline = {
            "bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
            "frac": (achieved / peaks["bf16_burst"]) if achieved and peaks["bf16_burst"] else None,
            "frac_burst": (achieved / peaks["bf16_burst"]) if achieved and peaks["bf16_burst"] else None,
            "traffic": traffic.get(dom),
            "peak_source": peaks["source"] + ": burst bf16 for the %d-step (%.0f ms) timed region; the sustained figure "
                                              "is used for the seconds-long leg only" % (args.steps, ms_total),
            "algorithmic_flops_per_launch": flops_per_gemm,
            "step_tflops": step_flops / (ms_step * 1e-3) / 1e12,
            "step_frac_burst": step_flops / (ms_step * 1e-3) / 1e12 / peaks["bf16_burst"] if peaks["bf16_burst"] else None,
            "phases_ms": phases,
            "phases_from": "a second pass of the same %d steps right after the timed region, with the library's per-phase CUDA "
                           "events on (they cost 0.04-0.1 ms per step, so `value` is timed without them)" % args.steps,
            "ms_per_step_with_phase_events": ms_step_profiled,
        }
        if sustained:
            sus_ph = {k: v for k, v in sustained["phases_ms_last_128_steps"].items() if k.endswith("_gemm")}
            if dom in sus_ph and sus_ph[dom] > 0 and peaks["bf16_sustained"]:
                roofline["achieved_sustained"] = flops_per_gemm / (sus_ph[dom] * 1e-3) / 1e12
                roofline["peak_sustained"] = peaks["bf16_sustained"]
                roofline["frac_sustained"] = roofline["achieved_sustained"] / peaks["bf16_sustained"]
            st_tf = step_flops / (sustained["ms_per_step"] * 1e-3) / 1e12
            sustained["step_tflops"] = st_tf
            sustained["step_frac_sustained"] = st_tf / peaks["bf16_sustained"] if peaks["bf16_sustained"] else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": CONFIG_WORKLOAD, "images_per_gpu": B, "tokens_per_gpu": T, "parallelism": f"dp{world}",
                       "activations": "bf16 %s resident in HBM for `value`; produced by the frozen GoogLeNet from host "
                                      "images for `e2e`" % ("channels_last ([B,C,H,W] whose memory is the token matrix: read in "
                                                            "place, no layout copy)" if args.acts_format == "channels_last" else "NCHW"),
                       "l2": "two rotating 103 MB input batches + ~2.2 GB per-step working set, both > 126 MB L2"},
            "roofline": roofline,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "final_step_stats": {k: last_stats[k] for k in ("loss", "rec", "l1", "n_dead")},
        }
        if strong is not None:
            line["strong_scaling"] = strong
        if dp is not None:
            line["exchange"] = {"mode": dp.mode, "what": "one kernel of libsvb on memory every rank maps; 'nvls': torch symmetric memory, "
                                "SUM section reduced by the NVSwitch (multimem.ld_reduce / multimem.st); 'ipc': CUDA-IPC peer "
                                "memory, peer loads / stores"}
        if e2e is not None:
            line["e2e"] = e2e
        if other_fmt is not None:
            line["other_activation_format"] = other_fmt
        if sustained is not None:
            line["sustained"] = sustained
        if dp_parity is not None:
            line["dp_parity"] = dp_parity
        if ie is not None:
            line["ie"] = ie
        if ie_pipe is not None:
            line["ie_pipeline"] = ie_pipe
        if gated is not None:
            line["gated"] = gated
        if world == 1 and "gpu_eager" not in skip:
            line["gpu_eager_reference"] = guarded(lambda: gpu_eager_reference(dev, xdev[0]))
            ge = line["gpu_eager_reference"]
            if "error" not in ge:
                for mode in ("fp32", "tf32", "bf16_autocast"):
                    ge[mode]["speedup_of_this_repo"] = ge[mode]["ms_per_step"] / ms_step
        if world == 1 and "cpu" not in skip:
            cpu_v, cpu_ms, cores = cpu_reference_throughput(args.cpu_images, 3, 1, os.cpu_count())
            line["cpu_baseline"] = {
                "value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{args.cpu_images} images = {args.cpu_images * HW_SIDE * HW_SIDE} tokens per step, "
                          f"1 warm-up + 3 timed steps of oracle/sae_oracle.py (fp32 torch CPU, {cores} threads)"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL prints its version banner on stdout) must not share the channel of the ONE JSON line: fd 1 is
    pointed at stderr for the whole run and the line is written to the saved descriptor at the end."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="svb", choices=["svb", "reference"])
    ap.add_argument("--images", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--cpu-images", type=int, default=8, help="images per step of the CPU reference sample")
    ap.add_argument("--no-ie", action="store_true", help="skip the indirect-effect and GatedSae sections")
    ap.add_argument("--no-numa", action="store_true", help="N > 1: do not bind ranks to their GPU's NUMA node")
    ap.add_argument("--sustain-s", type=float, default=3.0, help="length of the sustained leg in seconds (0: skip)")
    ap.add_argument("--skip", default="", help="comma-separated sections to skip: sustained,e2e,gated,ie,ie_pipeline,"
                                               "gpu_eager,cpu,dp_parity,other_format,strong")
    ap.add_argument("--acts-format", default="channels_last", choices=["nchw", "channels_last"],
                    help="memory format of the resident activations of the `value` leg")
    ap.add_argument("--no-graph", action="store_true",
                    help="e2e: launch every batch eagerly instead of replaying one captured CUDA graph (N = 1)")
    ap.add_argument("--no-fuse-producer", action="store_true",
                    help="e2e: torchvision's eager forward (ATen max-pool / add_ / relu_ / cat) instead of "
                         "producer.fuse_forward (libsvb max-pool and bias+relu+concat kernels between the cuDNN convolutions)")
    ap.add_argument("--ie-nchw-tail", action="store_true",
                    help="ie_pipeline: NCHW (torchvision's forward, ATen's max-pools) behind the first hooked layer")
    ap.add_argument("--ie-nchw", action="store_true",
                    help="ie_pipeline: the whole base model in NCHW on torchvision's eager forward (no fused head)")
    ap.add_argument("--no-fold-bn", action="store_true", help="e2e: keep the producer's BatchNorm layers un-folded")
    ap.add_argument("--two-pass", action="store_true",
                    help="e2e: compare with a second forward of an unhooked copy (the reference's structure) instead of "
                         "carrying the original activations through the same pass")
    ap.add_argument("--producer-nchw", dest="channels_last", action="store_false",
                    help="e2e: run the producer GoogLeNet in NCHW instead of channels_last (whose NHWC activations are "
                         "the zero-copy token matrix)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "svb":
        args.warmup = 3
    if args.no_ie:
        args.skip = ",".join(filter(None, [args.skip, "gated,ie,ie_pipeline"]))
    _quiet_stdout()
    return run_reference(args) if args.impl == "reference" else run_svb(args)


if __name__ == "__main__":
    sys.exit(main())
