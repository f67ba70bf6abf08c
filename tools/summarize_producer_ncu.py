"""profiles/<tag>_producer_full_summary.md from an `ncu --set full` capture of tools/producer_kernels_run.py:
  python tools/summarize_producer_ncu.py <tag> <report.ncu-rep>"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "DRAM read MB"),
        ("dram__bytes_write.sum", "DRAM write MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]


def main():
    tag, rep = sys.argv[1:3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    out = [f"# {tag}: `ncu --set full --clock-control none` of the producer kernels (tools/producer_kernels_run.py, 256 images)",
           "", "Per launch; times under ncu are cold-cache and serialised (never a bench value).  Algorithmic bytes: a max-pool",
           "reads its input and writes its output once; bias+relu reads and writes the convolution output once; the stem",
           "reads 77 MB of images and writes 411 MB.", "",
           "| kernel | " + " | ".join(c[1] for c in COLS) + " |", "|---|" + "---|" * len(COLS)]
    for r in body:
        cells = []
        for key, _ in COLS:
            if key not in hdr:
                cells.append("-")
                continue
            i = hdr.index(key)
            v = r[i].replace(",", "")
            try:
                f = float(v)
                if units[i] in ("byte", "Kbyte", "Mbyte", "Gbyte"):
                    f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[units[i]]
                if units[i] in ("ns", "ms"):
                    f *= {"ns": 1e-3, "ms": 1e3}[units[i]]
                cells.append(f"{f:.1f}" if f < 1e5 else f"{f:.0f}")
            except ValueError:
                cells.append(v)
        name = r[name_i].replace("<unnamed>::", "").split("(")[0][:48]
        out.append(f"| `{name}` | " + " | ".join(cells) + " |")
    path = os.path.join(ROOT, "profiles", f"{tag}_producer_full_summary.md")
    open(path, "w").write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main()
