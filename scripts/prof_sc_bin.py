#!/usr/bin/env python
"""Workload for ncu captures of the ScanContext search kernels (run under gpurun):
    ncu --set full --clock-control none --import-source on -k regex:sc_search_bin -s 12 -c 2 -o gpurun_out/x python scripts/prof_sc_bin.py single
modes: single (one query per launch), batch (256 queries per launch), general (f32-filter kernel, one query per launch)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("navtech-radar-slam_b200")
synth = importlib.import_module("navtech-radar-slam_b200.synth")
mode = sys.argv[1] if len(sys.argv) > 1 else "single"
n_kf = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
ctx = pkg.Context(0)
db, rev = synth.sc_database(10000, 500, seed=42, binary=True)
db = db[:n_kf]
d = pkg.scancontext.SCDatabase(ctx, capacity_hint=4 * n_kf)
for _ in range(4):
    d.append(db)
dev = torch.device("cuda", 0)
qs = torch.from_numpy(np.stack([db[j % n_kf] for (j, _, _) in rev[:256]])).to(dev)
res = torch.zeros(256 * 16, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
if mode == "general":
    d.set_search_mode(1)
if mode in ("single", "general"):
    for i in range(24):
        r = i % 4
        d.query_dev(qs[i % 64].data_ptr(), r * n_kf, (r + 1) * n_kf, 0, res.data_ptr())
else:
    for i in range(8):
        r = i % 4
        d.query_batch_dev(qs.data_ptr(), 256, r * n_kf, (r + 1) * n_kf, 0, res.data_ptr())
ctx.sync()
print("done", mode)
