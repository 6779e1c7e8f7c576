"""Summarise an ncu report's source page per kernel: stall reasons, opcode mix, hottest lines.
Usage: python tools/ncu_src_summary.py report.ncu-rep kernel_regex [n_hot]"""
import csv, subprocess, sys, io
from collections import Counter
rep, rx = sys.argv[1], sys.argv[2]
nhot = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several kernels may follow each other: split on "Kernel Name" rows
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
        data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if rows[j] and rows[j][0].startswith("0x"):
                data.append(rows[j])
            j += 1
        ix = {h: k for k, h in enumerate(hdr)}
        tot = sum(int(r[ix["# Samples"]]) for r in data)
        ex = sum(int(r[ix["Instructions Executed"]]) for r in data)
        print("==", name[:110]); print("  sass instrs", len(data), "samples", tot, "warp-instr executed", ex)
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = sorted(((sum(int(r[ix[s]]) for r in data), s) for s in stalls), reverse=True)[:7]
        print("  stalls:", ", ".join(f"{s[6:]} {v}" for v, s in agg))
        c, cs = Counter(), Counter()
        for r in data:
            t = r[ix["Source"]].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            c[op] += int(r[ix["Instructions Executed"]]); cs[op] += int(r[ix["# Samples"]])
        print("  ops:", ", ".join(f"{k} {v//1000}k/{cs[k]}" for k, v in c.most_common(16)))
        hot = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:nhot]
        for r in hot:
            top = max(stalls, key=lambda s: int(r[ix[s]]))
            print(f"   {r[ix['# Samples']]:>5} {top[6:]:<12} {r[ix['Source']][:90]}")
        i = j
    else:
        i += 1
