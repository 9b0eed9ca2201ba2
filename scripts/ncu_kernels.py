#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` capture covering many kernels (front-end, solver):
    python scripts/ncu_kernels.py gpurun_out/x.ncu-rep profiles/out.json "note"
For every kernel name: launches, average duration, DRAM bytes read + written per launch, achieved DRAM GB/s and its
fraction of the measured HBM peak (MEASURED_PEAKS.json), ncu's DRAM / SM throughput percentages, registers, grid."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    peak = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v, u = r[col[name]], units[col[name]]
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "ns": 1e-3, "us": 1, "ms": 1e3, "msecond": 1e3, "usecond": 1, "nsecond": 1e-3}.get(u, 1)
    return x * scale


agg = collections.OrderedDict()
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")).replace("void ", "")
    name = re.sub(r"<.*", "", name).split("::")[-1][:60]
    a = agg.setdefault(name, {"launches": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "dram_pct": 0.0, "sm_pct": 0.0, "regs": 0, "grid": 0, "block": 0})
    a["launches"] += 1
    a["us"] += val(r, "gpu__time_duration.sum") or 0
    a["rd"] += val(r, "dram__bytes_read.sum") or 0
    a["wr"] += val(r, "dram__bytes_write.sum") or 0
    a["dram_pct"] += val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") or 0
    a["sm_pct"] += val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed") or 0
    a["regs"] = int(val(r, "launch__registers_per_thread") or 0)
    a["grid"] = int(val(r, "launch__grid_size") or 0)
    a["block"] = int(val(r, "launch__block_size") or 0)
ks = []
tot = sum(a["us"] for a in agg.values()) or 1
for name, a in agg.items():
    n = a["launches"]
    us, by = a["us"] / n, (a["rd"] + a["wr"]) / n
    ks.append({"kernel": name, "launches": n, "avg_us": round(us, 2), "share_of_capture": round(a["us"] / tot, 4),
               "dram_bytes_per_launch": int(by), "dram_GBps": round(by / (us * 1e-6) / 1e9, 1) if us else None,
               "frac_of_measured_hbm_peak": round(by / (us * 1e-6) / 1e9 / peak, 4) if us else None,
               "ncu_dram_throughput_pct": round(a["dram_pct"] / n, 2), "ncu_sm_throughput_pct": round(a["sm_pct"] / n, 2),
               "registers": a["regs"], "grid": a["grid"], "block": a["block"]})
ks.sort(key=lambda k: -k["avg_us"] * k["launches"])
json.dump({"source": rep, "note": note, "hbm_peak_GBps": peak, "capture_total_us": round(tot, 1), "kernels": ks}, open(out, "w"), indent=1)
print(f"{len(ks)} kernels, {tot:.1f} us in the capture")
for k in ks[:40]:
    print(f"{k['kernel']:44s} x{k['launches']:<3d} {k['avg_us']:8.2f} us  {k['dram_bytes_per_launch'] / 1e6:8.3f} MB  {k['dram_GBps'] or 0:8.1f} GB/s  frac {k['frac_of_measured_hbm_peak'] or 0:6.3f}  sm {k['ncu_sm_throughput_pct']:5.1f}%")
