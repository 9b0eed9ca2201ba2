"""Bucket the warp-stall samples of an ncu report along the SASS of one kernel launch (developer aid):
python scripts/ncu_sass_buckets.py report.ncu-rep [chunk] [units]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
CH = int(sys.argv[2]) if len(sys.argv) > 2 else 60
UNITS = float(sys.argv[3]) if len(sys.argv) > 3 else 10000.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, r = rows[0], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max"]
for h, u, v in zip(hdr, rows[1], r):
    if h in want:
        print(h, v, u)
out = []
for h, v in zip(hdr, r):
    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        try:
            out.append((float(v), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError:
            pass
print(sorted(out, reverse=True)[:9])
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
starts = [i for i, x in enumerate(rows) if x and x[0] == "Kernel Name"]
s = starts[0]
e = starts[1] if len(starts) > 1 else len(rows)
hdr = rows[s + 1]
ix = {h: i for i, h in enumerate(hdr)}
data = [x for x in rows[s + 2:e] if len(x) == len(hdr)]


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


print("total samples", sum(f(x[ix["# Samples"]]) for x in data), "instr/unit",
      sum(f(x[ix["Instructions Executed"]]) for x in data) / UNITS)
for i in range(0, len(data), CH):
    ch = data[i:i + CH]
    smp = sum(f(x[ix["# Samples"]]) for x in ch)
    if smp == 0:
        continue
    ex = sum(f(x[ix["Instructions Executed"]]) for x in ch) / len(ch)
    ops = collections.Counter(re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", x[1].strip()).group(2) for x in ch)
    g = lambda k: sum(f(x[ix[k]]) for x in ch)
    print(f"{i:5d} smp={smp:5.0f} avgex={ex:8.0f} ssb={g('stall_short_sb'):4.0f} lsb={g('stall_long_sb'):4.0f} "
          f"mio={g('stall_mio'):4.0f} wait={g('stall_wait'):4.0f} bar={g('stall_barrier'):4.0f} sel={g('stall_selected'):3.0f} "
          f"nsel={g('stall_not_selected'):3.0f} {dict(ops.most_common(4))}")
top = sorted(data, key=lambda x: -f(x[ix["# Samples"]]))[:14]
for x in top:
    print(f"{f(x[ix['# Samples']]):6.0f} ex={f(x[ix['Instructions Executed']]):9.0f} {x[0][-5:]} {x[1][:80]}")
