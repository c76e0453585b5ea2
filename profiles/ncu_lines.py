#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel in an ncu report.

    python profiles/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [top_n] [launch_skip]
"""
import csv, io, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
skip = sys.argv[4] if len(sys.argv) > 4 else "0"  # launches of the kernel to skip inside the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      f"regex:{kern}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
items = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        ix = {}
        for j, k in enumerate(hdr):
            ix.setdefault(k, j)
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue  # keep only the per-CUDA-line aggregate rows (Address == '-')
    try:
        n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    except ValueError:
        continue
    items.append((n, s, fname, r[0], r[1].strip()[:100]))
tot = sum(i[0] for i in items) or 1
stot = sum(i[1] for i in items) or 1
print(f"# {kern}: {tot} warp instructions, {stot} stall samples")
for n, s, f, l, src in sorted(items, reverse=True)[:top]:
    print(f"{100*n/tot:5.1f}% inst {100*s/stot:5.1f}% samp  {f}:{l}: {src}")
