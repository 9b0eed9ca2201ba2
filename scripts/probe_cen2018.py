"""Developer probe: a few cen2018 extractions (run under ncu for per-kernel times)."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("navtech-radar-slam_b200")
synth = importlib.import_module("navtech-radar-slam_b200.synth")
ctx = pkg.Context(0)
world = synth.radar_world(seed=1234)
loader = pkg.orora.RadarLoader("mulran", "cen2018", 0.6, True, False, ctx=ctx)
raws = [synth.radar_frame(world, (2.5 * k, 0.1 * k, 0.01 * k), seed=1234 + k) for k in range(3)]
for rep in range(4):
    for r in raws:
        f = loader.extractFeatureFromRows(r, 1)
print("K", f.feat_in_cart_.shape, f.feat_in_polar_.shape)
