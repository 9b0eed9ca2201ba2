"""Developer probe: rate of the reference-faithful detect flow (ring-key 3-NN + 3 candidates) on the GPU."""
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("navtech-radar-slam_b200")
synth = importlib.import_module("navtech-radar-slam_b200.synth")
ctx = pkg.Context(0)
db, rev = synth.sc_database(10000, 500, seed=42)
d = pkg.scancontext.SCDatabase(ctx)
d.append(db)
for w in range(20):
    d.detect_knn(db[w], 0, 10000, 3)
t0 = time.perf_counter()
for k in range(1000):
    d.detect_knn(db[k % 64], 0, 10000, 3)
print("detect_knn q/s", 1000 / (time.perf_counter() - t0))
