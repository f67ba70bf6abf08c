"""SASS opcode histogram of the shipped libsvb.so (what the GPU box loads): proves which hardware paths the kernels use.
  python tools/sass_histogram.py [tag]  ->  profiles/<tag>_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = {"UTCHMMA": "tcgen05.mma (5th-gen tensor cores, accumulators in TMEM)", "UTCBAR": "tcgen05.commit -> mbarrier",
       "LDTM": "tcgen05.ld (TMEM -> registers)", "UTCATOMSWS": "tcgen05.alloc / dealloc (TMEM columns)",
       "LDGSTS": "cp.async (stem convolution: input rows of the next tile; fused node-IE kernel: staged averages)", "LDSM": "ldmatrix", "STSM": "stmatrix", "UTMALDG": "cp.async.bulk.tensor load (TMA)", "UTMASTG": "cp.async.bulk.tensor store (TMA)",
       "UTMAPF": "TMA L2 prefetch", "UTMACMDFLUSH": "bulk-group commit", "SYNCS": "mbarrier arrive / try_wait",
       "HMMA": "legacy mma.sync (only in the producer's im2col-free stem convolution, conv1_7x7s2_kernel; absent from the SAE path)", "SHFL": "warp shuffles", "ATOMS": "shared atomics",
       "ATOMG": "global atomics", "RED": "global reductions"}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    lib = os.path.join(ROOT, "sparse_vision_b200", "libsvb.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per_kernel = collections.defaultdict(collections.Counter)
    total = collections.Counter()
    fn = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:110]
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            op = m.group(1)
            total[op] += 1
            per_kernel[fn][op] += 1
    arch = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    out = [f"# {tag} — SASS opcode histogram of `sparse_vision_b200/libsvb.so`", "",
           f"`cuobjdump -sass` over {len(per_kernel)} kernels; ELF images: "
           + ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))) + ".", "",
           "| opcode | count | meaning |", "|---|---|---|"]
    for op, what in KEY.items():
        out.append(f"| `{op}` | {total.get(op, 0)} | {what} |")
    out += ["", f"All opcodes: {sum(total.values())} instructions, {len(total)} distinct.", "",
            "## Kernels that issue tensor-core / TMA instructions", "", "| kernel | UTCHMMA | LDTM | UTMALDG | UTMASTG |",
            "|---|---|---|---|---|"]
    for fn, c in sorted(per_kernel.items()):
        if c.get("UTCHMMA") or c.get("UTMALDG") or c.get("UTMASTG"):
            out.append(f"| `{fn}` | {c.get('UTCHMMA', 0)} | {c.get('LDTM', 0)} | {c.get('UTMALDG', 0)} | {c.get('UTMASTG', 0)} |")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_histogram.md")
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:22]))


if __name__ == "__main__":
    main()
