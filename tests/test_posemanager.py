"""The evaluation files of the reference's orora_odom node (SE2posemanager.hpp:151-245) as written by
navtech-radar-slam_b200/posemanager.py: formats, the values' definitions, and parseability the way
ORORA/scripts/evaluate_odometry.py reads eval_odom.txt (np.loadtxt(delimiter=","), 8 columns, first row skipped)."""
import importlib
import math

import numpy as np


def se2(x, y, yaw):
    T = np.eye(4)
    c, s = math.cos(yaw), math.sin(yaw)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1], T[0, 3], T[1, 3] = c, -s, s, c, x, y
    return T


def test_output_files(tmp_path):
    pmod = importlib.import_module("navtech-radar-slam_b200.posemanager")
    pm = pmod.SE2PoseManager(str(tmp_path), "mulran", "ORORA", "cen2018")
    assert pm.output_pose_path_.endswith("mulran_ORORA_cen2018_eval_odom.txt")
    gts = [(2.5, 0.1, 0.0174533), (0.0, 0.0, 0.0), (1.25, -0.05, -0.03)]
    ests = [se2(2.51234567, 0.0987654, -0.0174), se2(0.0, 0.0, 0.0), se2(1.2, -0.04, 0.031)]
    curr_gt, curr = np.eye(4), np.eye(4)
    for k, (g, Test) in enumerate(zip(gts, ests)):
        Tgt = pmod.xyyaw2se2(g, "mulran")
        assert abs(math.atan2(Tgt[1, 0], Tgt[0, 0]) + g[2]) < 1e-6           # MulRan ground truth enters with -yaw (:181-183)
        curr_gt, curr = curr_gt @ Tgt, curr @ Test
        err = pm.writePoseError(Tgt, Test)
        pm.writePose(Tgt, Test, 1560000000000000000 + k * 250000000, 1560000000250000000 + k * 250000000, 1e9)
        pm.writeTraj(curr, pm.est_traj_path_)
        pm.writeTraj(curr_gt, pm.gt_traj_path_)
        pm.writeTimeTaken(1234 + k, Tgt)
        assert err[0] >= 0 and 0 <= err[1] < 5
    rows = open(pm.output_pose_path_).read().splitlines()
    assert len(rows) == 3 and all(r.count(",") == 7 for r in rows)
    # C++ default ostream formatting: 6 significant digits, %g
    assert rows[0].split(",")[3] == "2.51235" and rows[0].split(",")[6] == "1.56e+09"
    assert rows[1] == "0,0,-0,0,0,0,1.56e+09,1.56e+09"               # -asin(+0) prints "-0" like C++ does
    data = np.loadtxt(pm.output_pose_path_, delimiter=",")                   # evaluate_odometry.py:191
    assert data.shape == (3, 8)
    # yaw columns are -asin(T(0,1)) (:220-224): the estimate's yaw comes back, the MulRan ground truth's negated
    assert abs(data[0, 5] - (-0.0174)) < 1e-6 and abs(data[0, 2] + 0.0174533) < 1e-6
    err_rows = np.loadtxt(pm.output_pose_error_path_, delimiter=",")
    assert err_rows.shape == (3, 5) and err_rows[1, 2] == 0 and err_rows[1, 3] == 0     # ts_abs < EPSILON_TRANS (:205-206)
    tt = open(pm.output_time_path_).read().splitlines()
    assert tt == ["1234,0", "1235,1", "1236,0"]                              # is_stopped only for the zero ground-truth motion
    traj = np.loadtxt(pm.est_traj_path_, delimiter=",")
    assert traj.shape == (3, 3) and abs(traj[-1, 0] - curr[0, 3]) < 1e-4
