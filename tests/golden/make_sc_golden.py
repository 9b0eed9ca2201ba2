"""Regenerates tests/golden/sc_golden.npz.

The reference ships no golden vectors for ScanContext and its SCManager cannot
be compiled here, so these vectors come from the CPU restatement oracle/sc_ref.c on seeded
synthetic descriptors; the kNN candidate ids additionally come from the REFERENCE's own
nanoflann (oracle/_ref/libref_kdtree.so, built from /root/reference) when it is available.
Run from the repo root:  python tests/golden/make_sc_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402

synth = importlib.import_module("navtech-radar-slam_b200.synth")


def main():
    O.build()
    out = {}
    for tag, binary in (("bin", True), ("gen", False)):
        db, rev = synth.sc_database(96, 12, seed=7 if binary else 8, binary=binary)
        q = db[-1].copy()
        dist, shift, best = O.sc_search(q, db, 0, 95)
        out[f"{tag}_db"] = db
        out[f"{tag}_query"] = q
        out[f"{tag}_dist"] = dist
        out[f"{tag}_shift"] = shift
        out[f"{tag}_best"] = np.array(best, dtype=np.float64)
        keys = np.stack([O.sc_ringkey_f32(d) for d in db])
        out[f"{tag}_ringkeys_f32"] = keys
        out[f"{tag}_sectorkey0"] = O.sc_sectorkey(db[0])
        ref = O.ref_knn(keys[:66], keys[-1], 3)
        if ref is not None:
            out[f"{tag}_ref_knn_idx"] = ref[0]
            out[f"{tag}_ref_knn_d2"] = ref[1]
    cloud = synth.feature_cloud(1500, seed=3)
    out["cloud"] = cloud
    out["cloud_desc"] = O.sc_make(cloud)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "sc_golden.npz"), **out)
    print("wrote sc_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
