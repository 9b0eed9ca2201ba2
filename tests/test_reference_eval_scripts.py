"""The evaluation files this repo's drivers write (posemanager.SE2PoseManager / cpp/se2_posemanager.hpp, the formats of
SE2posemanager.hpp:151-259) scored by the REFERENCE'S OWN evaluation scripts, run unmodified from /root/reference:
outlier-robust-radar-odometry/scripts/evaluate_odometry.py (KITTI-style drift + ATE from eval_odom.txt),
calc_mean_errors.py (rel_error_odom.txt) and calc_mean_time.py (time_taken.txt).  matplotlib is not in this image and the
scripts import it at the top for their plots: an empty stand-in package on PYTHONPATH lets the numeric part run.
CPU only; skipped where /root/reference is absent."""
import importlib
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

SCRIPTS = "/root/reference/outlier-robust-radar-odometry/scripts"
pytestmark = pytest.mark.skipif(not os.path.isdir(SCRIPTS), reason="reference scripts not present")


def se2(x, y, yaw):
    T = np.eye(4)
    c, s = math.cos(yaw), math.sin(yaw)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1], T[0, 3], T[1, 3] = c, -s, s, c, x, y
    return T


def run(script, path, cwd, env):
    r = subprocess.run([sys.executable, os.path.join(SCRIPTS, script), "-f", str(path)], cwd=cwd, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    return r.stdout


@pytest.mark.parametrize("drift", [0.0, 0.01])
def test_reference_scripts_score_our_files(tmp_path, drift):
    pmod = importlib.import_module("navtech-radar-slam_b200.posemanager")
    stub = tmp_path / "stub" / "matplotlib"
    stub.mkdir(parents=True)
    (stub / "__init__.py").write_text("")
    (stub / "pyplot.py").write_text("")
    env = dict(os.environ, PYTHONPATH=str(tmp_path / "stub"))
    out = tmp_path / "outputs"
    pm = pmod.SE2PoseManager(str(out), "mulran", "ORORA", "cen2018")
    rng = np.random.default_rng(1)
    curr_gt, curr = np.eye(4), np.eye(4)
    n = 420                                               # 2.5 m per pair -> > 1 km: every KITTI segment length is used
    for k in range(n):
        g = (2.5, 0.0, -0.01)                             # MulRan ground truth enters with -yaw (SE2posemanager.hpp:181-183)
        Tgt = pmod.xyyaw2se2(g, "mulran")
        # the estimate: the true motion, scaled by (1 + drift) along x
        Test = se2(2.5 * (1.0 + drift), 0.0, 0.01)
        curr_gt, curr = pmod.mul4(curr_gt, Tgt), pmod.mul4(curr, Test)
        pm.writePoseError(Tgt, Test)
        pm.writePose(Tgt, Test, 1560000000000000000 + 250000000 * k, 1560000000250000000 + 250000000 * k, 1e9)
        pm.writeTraj(curr, pm.est_traj_path_)
        pm.writeTraj(curr_gt, pm.gt_traj_path_)
        pm.writeTimeTaken(400 + (k % 7), Tgt if k % 50 else np.eye(4))     # every 50th pair "stopped"
    txt = run("evaluate_odometry.py", pm.output_pose_path_, tmp_path, env)
    m = re.search(r"(\d+\.\d+)/(\d+\.\d+)", re.sub(r"\x1b\[[0-9;]*m", "", txt))
    assert m, txt
    t_err, r_err = float(m.group(1)), float(m.group(2))
    ate = float(re.search(r"ATE: ([0-9.eE+-]+) m", txt).group(1))
    if drift == 0.0:
        assert t_err < 0.01 and r_err < 0.01 and ate < 1e-3          # printed with 6 significant digits: tiny, not zero
    else:
        assert 0.8 < t_err < 1.1 and r_err < 0.01                    # a 1 % scale error along a gently curving path: ~1 % drift
        assert ate > 0.1
    assert os.path.exists(tmp_path / "pose_error_rigid_icra3.csv")   # the script's own by-product
    txt = run("calc_mean_errors.py", pm.output_pose_error_path_, tmp_path, env)
    mean_ts = float(re.search(r"Mean traslation error: ([0-9.eE+-]+) m", txt).group(1))
    assert abs(mean_ts - 2.5 * drift) < 1e-5
    perc = float(re.search(r"perc\. error: ([0-9.eE+-]+) %", txt).group(1))
    assert abs(perc - 100 * drift) < 1e-3
    txt = run("calc_mean_time.py", pm.output_time_path_, tmp_path, env)
    assert f"({n}, 2)" in txt and f"({n - 9}, 2)" in txt             # 9 of the 420 rows carry is_stopped = 1
    total = float(re.search(r"Mean total time: ([0-9.eE+-]+) ms", txt).group(1))
    assert 0.400 <= total <= 0.406
