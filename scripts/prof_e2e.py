#!/usr/bin/env python
"""Where the host-to-host time of a 256-query ScanContext batch goes (GPU box):
    python scripts/prof_e2e.py
prints, per 256-query step against the 10k database: the device-resident launch, the upload alone, the host-side
classification alone, the C entry point with pinned input and output, and the Python wrapper."""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("navtech-radar-slam_b200")
synth = importlib.import_module("navtech-radar-slam_b200.synth")
sc = pkg.scancontext
L = pkg.lib()
ctx = pkg.Context(0)
db, rev = synth.sc_database(10000, 500, seed=42, binary=True)
d = sc.SCDatabase(ctx)
for _ in range(8):
    d.append(db)                                   # 8 copies: rotate so that L2 does not hold the records
Q = 256
qs = np.stack([db[j] for (j, _, _) in rev[:Q]])
q_pin = torch.from_numpy(qs).pin_memory()
q_dev = q_pin.cuda()
r_dev = torch.zeros(Q * 16, dtype=torch.uint8, device="cuda")
r_pin = torch.zeros(Q * 16, dtype=torch.uint8).pin_memory()
REPS = 40


def timeit(fn, reps=REPS):
    for k in range(5):
        fn(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(reps):
        fn(k)
    torch.cuda.synchronize()
    return 1e6 * (time.perf_counter() - t0) / reps


def dev(k):
    c = k % 8
    d.query_batch_dev(q_dev.data_ptr(), Q, c * 10000, (c + 1) * 10000, 0, r_dev.data_ptr())


def dev_sync(k):
    dev(k)
    torch.cuda.synchronize()


def upload(k):
    q_dev.copy_(q_pin, non_blocking=True)
    torch.cuda.synchronize()


def classify(k):
    p = q_pin.numpy()
    for i in range(Q):
        L.nrs_sc_descriptor_is_binary(p[i].ctypes.data_as(C.c_void_p))


def c_entry(k):
    c = k % 8
    pkg._lib.check(L.nrs_sc_query_batch(d._h, C.c_void_p(q_pin.data_ptr()), Q, c * 10000, (c + 1) * 10000, C.c_void_p(r_pin.data_ptr())))


def wrapper(k):
    c = k % 8
    d.query_batch(q_pin.numpy(), c * 10000, (c + 1) * 10000, as_array=True)


print("device-resident launches back to back   %8.1f us per step" % timeit(dev))
print("device-resident launch + synchronise    %8.1f us" % timeit(dev_sync))
print("upload of 2.46 MB alone + synchronise    %8.1f us" % timeit(upload))
print("256 x nrs_sc_descriptor_is_binary (ctypes) %6.1f us" % timeit(classify, 10))
print("nrs_sc_query_batch, pinned in / pinned out %6.1f us" % timeit(c_entry))
print("SCDatabase.query_batch (numpy out)       %8.1f us" % timeit(wrapper))
