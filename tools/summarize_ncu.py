"""Turns the ncu outputs of a round into the tracked summaries under profiles/.
  python tools/summarize_ncu.py <tag> <launches.csv> <full.ncu-rep>
writes profiles/<tag>_launches.csv, <tag>_launches_summary.md, <tag>_gemms_full_summary.md and refreshes traffic.json."""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PHASE = {"EpiEncT": "enc_gemm", "EpiDecNchw": "dec_gemm", "EpiDec": "dec_gemm", "EpiDPreT": "dE_gemm",
         "fused_bwd": "dE+dWenc_fused_gemm", "EpiPartialOnes": "dWenc_gemm", "EpiPartial": "dWdec_gemm"}


def main():
    tag, launches, rep = sys.argv[1:4]
    prof = os.path.join(ROOT, "profiles")
    txt = open(launches).read()
    txt = txt[txt.index('"ID"'):]
    shutil.copy(launches, os.path.join(prof, f"{tag}_launches.csv"))
    rows = list(csv.DictReader(io.StringIO(txt)))
    names = [r["Kernel Name"] for r in rows]
    # the first kernel of a step: the pack pass (NCHW input) or the x-statistics pass (channels_last input, read in place)
    starts = [i for i, n in enumerate(names) if "pack_nchw" in n or "x_stats_tokens" in n]
    s, e = starts[-3], starts[-2]          # one complete step near the end
    step = rows[s:e]
    total = sum(float(r["Metric Value"]) for r in step) / 1000
    cmd = os.environ.get("SVB_NCU_CMD", "python bench.py --steps 3 --warmup 3")
    out = [f"# {tag} — ncu launch list of one training step (`{cmd}`)", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and serialised — compare SHARES.",
           "", "| # | kernel | grid | us | share |", "|---|---|---|---|---|"]
    for i, r in enumerate(step):
        us = float(r["Metric Value"]) / 1000
        out.append(f"| {i} | `{r['Kernel Name'][:96]}` | {r['Grid Size']} | {us:.1f} | {100 * us / total:.1f}% |")
    out += ["", f"Sum over the step's {len(step)} launches: {total / 1000:.3f} ms."]
    open(os.path.join(prof, f"{tag}_launches_summary.md"), "w").write("\n".join(out) + "\n")

    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rr[0], rr[1], rr[2:]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "lts__t_sector_op_read_hit_rate.pct", "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed",
            "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]
    kn = hdr.index("Kernel Name")
    labels = []
    seen = collections.Counter()
    for d in data:
        lab = next((v for k, v in PHASE.items() if k in d[kn]), None)
        if lab is None:
            seen["dW"] += 1
            lab = "dWenc_gemm" if seen["dW"] == 1 else "dWdec_gemm"   # the encoder-side weight gradient runs first
        labels.append(lab)
    md = [f"# {tag} — `ncu --set full --clock-control none` of the tensor-core launches of one training step", "",
          f"Command: `ncu --set full --clock-control none --import-source on -k regex:\"gemm|fused_bwd\" --launch-skip <3 steps> --launch-count <1 step> {os.environ.get('SVB_NCU_CMD', 'python bench.py --steps 3 --warmup 3')}` (report: `{os.path.basename(rep)}`).",
          "", "| metric | " + " | ".join(labels) + " |", "|---|" + "---|" * len(labels)]
    traffic = {}
    for w in want:
        if w not in hdr:
            continue
        i = hdr.index(w)
        md.append(f"| `{w}` [{units[i]}] | " + " | ".join(d[i] for d in data) + " |")
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    for lab, d in zip(labels, data):
        traffic[lab] = int(float(d[ir]) * scale[units[ir]] + float(d[iw]) * scale[units[iw]])
    traffic["_source"] = f"profiles/{tag}_gemms_full_summary.md: dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, cfg2 step (bytes)"
    open(os.path.join(prof, f"{tag}_gemms_full_summary.md"), "w").write("\n".join(md) + "\n")
    json.dump(traffic, open(os.path.join(prof, "traffic.json"), "w"), indent=1)
    print("\n".join(md[-14:]))


if __name__ == "__main__":
    main()
