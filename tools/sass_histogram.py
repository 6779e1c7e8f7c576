#!/usr/bin/env python
"""SASS opcode histogram of every kernel in brutefir_b200/csrc/*.o (cuobjdump -sass), written to
profiles/<round>_sass_histogram.txt.  Evidence for which hardware paths the shipped objects use: packed FP32 pairs
(FFMA2 / FADD2), cp.async (LDGSTS), bulk copies + mbarriers (UBLKCP / SYNCS), and the absence of tensor-core opcodes
(UTC*MMA / HMMA) -- nothing on this path is a contraction."""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FMUL", "FADD", "DFMA", "DMUL", "DADD", "LDGSTS", "UBLKCP", "SYNCS", "LDG", "STG", "LDS",
       "STS", "BAR", "SHFL", "UTCHMMA", "UTCQMMA", "HMMA", "UTMALDG", "LDTM"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except Exception:
        return name


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_histogram.txt")
    lines = []
    for obj in sorted(glob.glob(os.path.join(ROOT, "brutefir_b200", "csrc", "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        fn, hist = None, None
        kernels = []
        for ln in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", ln)
            if m:
                fn, hist = m.group(1), collections.Counter()
                kernels.append((fn, hist))
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
            if m and hist is not None:
                hist[m.group(1)] += 1
        lines.append(f"== {os.path.basename(obj)}: {len(kernels)} kernels")
        for fn, hist in kernels:
            name = demangle(fn)
            name = re.sub(r"\(.*$", "", name).replace("void bf::", "")
            total = sum(hist.values())
            key = " ".join(f"{k}={hist[k]}" for k in KEY if hist[k])
            top = " ".join(f"{k}:{v}" for k, v in hist.most_common(8))
            lines.append(f"{name}\n    {total} instructions | {key}\n    top: {top}")
    with open(out_path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"wrote {out_path}: {len(lines)} lines")


if __name__ == "__main__":
    main()
