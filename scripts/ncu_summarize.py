#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep / launch-list CSVs into the small JSON summaries kept under profiles/.
    python scripts/ncu_summarize.py rep  gpurun_out/sc_v2b.ncu-rep profiles/r01_sc_search_v2_ncu_full.json "note"
    python scripts/ncu_summarize.py list gpurun_out/launches_r01.csv profiles/r01_launch_list.json "note"
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max', 'smsp__thread_inst_executed_per_inst_executed.ratio']


def rep(path, out, note):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    launches = []
    for r in data:
        launches.append({h: [r[i], units[i]] for i, h in enumerate(hdr) if h in KEEP})
    for l in launches:
        rd = float(l['dram__bytes_read.sum'][0]) * (1e6 if l['dram__bytes_read.sum'][1] == 'Mbyte' else 1e3 if l['dram__bytes_read.sum'][1] == 'Kbyte' else 1)
        wr = float(l['dram__bytes_write.sum'][0]) * (1e6 if l['dram__bytes_write.sum'][1] == 'Mbyte' else 1e3 if l['dram__bytes_write.sum'][1] == 'Kbyte' else 1)
        l['dram_bytes_per_launch'] = rd + wr
    json.dump({"source": path, "note": note, "launches": launches,
               "dram_bytes_per_launch": launches[0]['dram_bytes_per_launch'] if launches else None}, open(out, "w"), indent=1)


def lst(path, out, note):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = re.sub(r'\(.*', '', r[ki]).replace('<unnamed>::', '').replace('void ', '')[:80]
        v = float(r[vi].replace(',', ''))
        v = v / 1000 if r[ui] == 'ns' else v * 1000 if r[ui] == 'ms' else v
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    ks = [{"kernel": k, "launches": v[0], "total_us": round(v[1], 1), "avg_us": round(v[1] / v[0], 2),
           "share": round(v[1] / tot, 4)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    json.dump({"source": path, "note": note, "total_us": round(tot, 1), "kernels": ks}, open(out, "w"), indent=1)


if __name__ == "__main__":
    {"rep": rep, "list": lst}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
