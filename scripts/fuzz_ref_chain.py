"""Randomised comparison of the oracle's extract_feature / match_frames chain with the reference's compiled RadarLoader
(oracle/_ref/libref_frontend.so: extractFeature + matchFeatures; cv::imread / cv::ORB / BFMatcher by the real cv2):
python scripts/fuzz_ref_chain.py SEED N_WORLDS -- random worlds and motions, both datasets, both extractors, voxelisation
on / off with three leaf sizes.  cen2019 runs on the reference's libstdc++ tie order (the oracle is handed those
keypoints, see tests/test_ref_frontend.py).  With voxelisation on, the reference side runs the stand-in VoxelGrid."""
import importlib
import os
import shutil
import sys
import tempfile

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import frontend_ref as F, refbridge as R  # noqa: E402

synth = importlib.import_module("navtech-radar-slam_b200.synth")
rng = np.random.default_rng(int(sys.argv[1]))
bad = n_frames = n_pairs = 0
for w in range(int(sys.argv[2])):
    dataset = ("mulran", "oxford")[w % 2]
    extraction = ("cen2018", "cen2019")[(w // 2) % 2]
    vox = bool(rng.integers(0, 2))
    leaf = float(rng.choice([0.4, 0.6, 1.0]))
    world = synth.radar_world(seed=int(rng.integers(0, 1 << 30)))
    d = tempfile.mkdtemp()
    raws, names = [], []
    x = y = yaw = 0.0
    for k in range(3):
        raws.append(synth.radar_frame(world, (x, y, yaw), seed=int(rng.integers(0, 1 << 30)), dataset=dataset))
        names.append(f"{1560000000000000000 + k * 250000000}.png")
        cv2.imwrite(os.path.join(d, names[-1]), raws[-1])
        x, y, yaw = x + rng.uniform(0, 4), y + rng.uniform(-0.5, 0.5), yaw + rng.uniform(-0.05, 0.05)
    L = R.RadarLoader(dataset, extraction, leaf, vox)
    P = F.DatasetParams(dataset)
    fr = [L.extractFeature(d, nm) for nm in names]
    fo = []
    for raw in raws:
        over = None
        if extraction == "cen2019":
            over = F.cen2019features(F.load_radar(raw, P.navtech_version)[3], P.max_points, P.min_range, order="libstdc++")
        fo.append(F.extract_feature(raw, dataset, leaf, vox, extraction, polar_override=over))
    for k in range(3):
        n_frames += 1
        same = all(fr[k][a].shape == fo[k][b].shape and np.array_equal(fr[k][a], fo[k][b])
                   for a, b in (("polar", "polar"), ("img", "img"), ("feat_in_cart", "feat_in_cart"), ("t", "t"), ("kp", "kp"), ("desc", "desc")))
        if not same:
            bad += 1
            print("MISMATCH extract", w, dataset, extraction, vox, leaf, k, file=sys.stderr)
    for k in range(1, 3):
        n_pairs += 1
        pm, cm = L.matchFeatures(fr[k - 1], fr[k])
        po, co, q, t = F.match_frames(fo[k - 1], fo[k])
        if not (np.array_equal(pm["feat_in_cart"], po) and np.array_equal(cm["feat_in_cart"], co)
                and np.array_equal(pm["t"], fo[k - 1]["t"][q]) and np.array_equal(cm["t"], fo[k]["t"][t])):
            bad += 1
            print("MISMATCH match", w, dataset, extraction, vox, leaf, k, file=sys.stderr)
    L.close()
    shutil.rmtree(d, ignore_errors=True)
print(f"done, frames: {n_frames}, pairs: {n_pairs}, mismatches: {bad}", file=sys.stderr)
