#!/usr/bin/env python
"""Per-source-line samples / executed instructions of one kernel from an .ncu-rep (captured with --import-source on,
code built with -lineinfo):   python scripts/ncu_lines.py gpurun_out/x.ncu-rep [top N] [file filter]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
flt = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
out, cur_file, hdr = [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0].isdigit() and len(r) > 8 and r[2] == "-":        # a CUDA source line row (aggregated)
        d = dict(zip(hdr, r))
        out.append((cur_file.split("/")[-1], int(r[0]), r[1].strip(), int(d["# Samples"]), int(d["Instructions Executed"])))
seen, uniq = set(), []
for o in out:
    if (o[0], o[1]) not in seen:
        seen.add((o[0], o[1]))
        uniq.append(o)
tot_s = sum(o[3] for o in uniq) or 1
tot_i = sum(o[4] for o in uniq) or 1
print(f"total samples {tot_s}, total warp instructions {tot_i}")
for o in sorted((u for u in uniq if flt in u[0]), key=lambda x: -x[3])[:top]:
    print(f"{o[0]}:{o[1]:<5d} samples {100.0 * o[3] / tot_s:5.1f}%  inst {100.0 * o[4] / tot_i:5.1f}%  {o[2][:110]}")
