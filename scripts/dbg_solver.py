"""Developer probe: phase timings of the ORORA solver kernels on synthetic frames (GPU box)."""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("navtech-radar-slam_b200")
synth = importlib.import_module("navtech-radar-slam_b200.synth")
ctx = pkg.Context(0)
orr = pkg.orora
world = synth.radar_world(seed=1234)
poses = [(2.5 * k, 0.1 * k, np.deg2rad(1.0) * k) for k in range(8)]
loader = orr.RadarLoader("mulran", "cen2019", 0.6, True, False, ctx=ctx)
feats = [loader.extractFeatureFromRows(synth.radar_frame(world, poses[k], seed=1234 + k), k) for k in range(8)]
p = orr.default_params()
vel = 0.0
names = ["load", "greedy", "core", "search", "(gap)", "TIMs", "GNC", "COTEx", "COTEy"]
for rep in range(2):
    for k in range(1, 8):
        pm, cm = loader.matchFeatures(feats[k - 1], feats[k])
        t0 = time.perf_counter()
        r = orr.solve(ctx, pm.feat_in_cart_, cm.feat_in_cart_, p, vel)
        dt = time.perf_counter() - t0
        t = (C.c_uint64 * 12)()
        pkg.lib().nrs_orora_last_timings(t)
        t = [int(x) for x in t]
        i = r["info"]
        ph = " ".join(f"{n}={(t[j + 1] - t[j]) / 1e3:.1f}" for j, n in enumerate(names))
        print(k, "N", pm.feat_in_cart_.shape[1], "M", i.M, "nodes", i.clique_nodes, "iters", i.gnc_iters, f"{dt * 1e3:.2f} ms |", ph, "greedy_steps", t[10], flush=True)
        vel = r["T"][0, 3] / 0.25
