"""PINNING the evaluation files on the reference itself: oracle/_ref/libref_posemanager.so is the reference's own
outlier-robust-radar-odometry/include/SE2posemanager.hpp compiled unmodified (with src/radar_utils.cpp for
get_groundtruth_odometry) against the stand-in headers of oracle/refshim/; what the node does per scan pair
(odometry.cpp:253-285) is replayed through it and through this repo's writers -- posemanager.py (Python drivers) and,
via tests/test_cpp_driver.py's byte-for-byte twin test, cpp/se2_posemanager.hpp (C++ node).  Files must be identical:
names (<seq>/outputs/<dataset>_<algorithm>_<extraction>_*.txt), ofstream formatting, every column's definition, the
ground-truth readers (MulRan: loadGTPoses with y read from the yaw column; Oxford: the row matched by its two stamps).
Runs on CPU; skipped where the library is absent."""
import ctypes as C
import importlib
import math
import os

import numpy as np
import pytest

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_posemanager.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_posemanager.so not built (needs /root/reference)")


def se2(x, y, yaw):
    T = np.eye(4)
    c, s = math.cos(yaw), math.sin(yaw)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1], T[0, 3], T[1, 3] = c, -s, s, c, x, y
    return T


@pytest.mark.parametrize("dataset", ["mulran", "oxford"])
def test_files_equal_the_compiled_reference(tmp_path, dataset):
    pmod = importlib.import_module("navtech-radar-slam_b200.posemanager")
    L = C.CDLL(LIB)
    L.ref_pm_create.restype = C.c_void_p
    L.ref_pm_create.argtypes = [C.c_char_p] * 4 + [C.c_double] * 4
    L.ref_pm_pair.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_double]
    L.ref_pm_destroy.argtypes = [C.c_void_p]
    rng = np.random.default_rng(17)
    n = 200
    den = 1e9 if dataset == "mulran" else 1e6
    step = 250000000 if dataset == "mulran" else 250000
    stamps = [1560000000000000000 // (1 if dataset == "mulran" else 1000) + step * k for k in range(n + 2)]
    gts = []
    for k in range(n + 1):
        g = [float(rng.normal(2.5, 1.0)), float(rng.normal(0, 0.2)), float(rng.normal(0, 0.03)), float(rng.normal(0, 0.01))]
        if k % 19 == 0:
            g = [0.0, 0.0, 0.0, 0.0]                         # a stopped pair
        if k % 23 == 1:
            g = [0.001, 0.0, 0.0005, 0.002]                  # nearly stopped
        if k % 31 == 2:
            g = [1e-5, 3e-7, -2e-6, 2.0]                     # values that print in exponent form; a large turn
        gts.append(g)
    seqs = []
    for who in ("ref", "ours"):
        seq = tmp_path / who
        (seq / "gt").mkdir(parents=True)
        with open(seq / "gt" / "radar_odometry.csv", "w") as f:
            if dataset == "mulran":
                for k, g in enumerate(gts):                  # (Time, x, y, yaw)
                    f.write(f"{stamps[k]},{g[0]:.9g},{g[1]:.9g},{g[2]:.9g}\n")
            else:
                f.write("source_timestamp,destination_timestamp,x,y,z,roll,pitch,yaw,source_radar_timestamp,destination_radar_timestamp\n")
                for k, g in enumerate(gts):
                    f.write(f"{k},{k + 1},{g[0]:.9g},{g[1]:.9g},0,0,0,{g[3]:.9g},{stamps[k + 1]},{stamps[k]}\n")
        seqs.append(seq)
    h = L.ref_pm_create(str(seqs[0]).encode(), dataset.encode(), b"ORORA", b"cen2018", 0.6, 0.75, 0.1, math.radians(9.0))
    assert h
    pm = pmod.SE2PoseManager.from_seq_dir(str(seqs[1]), dataset, "ORORA", "cen2018")
    assert os.path.basename(pm.output_pose_path_) == f"{dataset}_ORORA_cen2018_eval_odom.txt"
    curr_gt, curr = np.eye(4), np.eye(4)
    for k in range(n):
        g = gts[k]
        yaw_true = (-g[2] if dataset == "mulran" else g[3])
        T = se2(g[0] + rng.normal(0, 0.05), (g[2] if dataset == "mulran" else g[1]) + rng.normal(0, 0.05), yaw_true + rng.normal(0, 0.004))
        if k % 19 == 0:
            T = np.eye(4)
        t1, t2, us = stamps[k], stamps[k + 1], 300 + 7 * k
        Tf = np.asfortranarray(T)
        assert L.ref_pm_pair(h, t1, t2, Tf.ctypes.data_as(C.c_void_p), us, den) == 0
        xyyaw = pm.getXYYawPose(t1, t2) if dataset == "oxford" else pm.getXYYawPose(t1)
        Tgt = pmod.xyyaw2se2(xyyaw, dataset)
        curr_gt, curr = pmod.mul4(curr_gt, Tgt), pmod.mul4(curr, T)
        pm.writePoseError(Tgt, T)
        pm.writePose(Tgt, T, t1, t2, den)
        pm.writeTraj(curr, pm.est_traj_path_)
        pm.writeTraj(curr_gt, pm.gt_traj_path_)
        pm.writeTimeTaken(us, Tgt)
    assert L.ref_pm_pair(h, 12345, 67890, np.asfortranarray(np.eye(4)).ctypes.data_as(C.c_void_p), 1, den) == 1     # no such stamp: the reference throws
    with pytest.raises(ValueError):
        pm.getXYYawPose(12345, 67890) if dataset == "oxford" else pm.getXYYawPose(12345)
    L.ref_pm_destroy(h)
    for name in ("rel_error_odom", "eval_odom", "time_taken", "est_traj", "gt_traj"):
        a = (seqs[0] / "outputs" / f"{dataset}_ORORA_cen2018_{name}.txt").read_text().splitlines()
        b = (seqs[1] / "outputs" / f"{dataset}_ORORA_cen2018_{name}.txt").read_text().splitlines()
        assert len(a) == len(b) == n, (name, len(a), len(b))
        diff = [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y]
        assert not diff, (name, len(diff), diff[:3])
    # the MulRan reader's own indexing: tsy_gt (column 1) is the csv's yaw column, not its y column
    ev = np.loadtxt(seqs[0] / "outputs" / f"{dataset}_ORORA_cen2018_eval_odom.txt", delimiter=",")
    if dataset == "mulran":
        assert np.allclose(ev[:, 1], [g[2] for g in gts[:n]], atol=1e-6) and not np.allclose(ev[:, 1], [g[1] for g in gts[:n]], atol=1e-3)
