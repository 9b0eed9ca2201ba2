"""ROS-free driver over a directory of "oxford-form" polar PNGs: the loop of
outlier-robust-radar-odometry/src/odometry.cpp:153-393 (extractFeature -> matchFeatures -> ORORA::calcTransform,
poses chained as curr_odom = curr_odom * T) plus, optionally, what the SC-A-LOAM back end does with the clouds
that node publishes (one ScanContext keyframe per frame, detectLoopClosureID against the growing database;
pgo/SC-A-LOAM/src/laserPosegraphOptimization.cpp:486-495,553-577).  Ground-truth handling, rviz topics and the
pose-graph optimisation itself are not part of this (SURVEY.md 8f rank 3 / out of scope)."""
import os
import time

import numpy as np

from . import orora, scancontext


def radar_files(datadir):
    """get_file_names (radar_utils.cpp:88-102): the PNGs of the directory, sorted by name (= timestamp)."""
    return sorted(f for f in os.listdir(datadir) if f.endswith(".png"))


def run_odometry(datadir, dataset="mulran", keypoint_extraction="cen2018", voxel_size=0.6, use_voxelization=True,
                 frame_rate=4.0, loop_detection=False, sc_dist_thres=0.2, ctx=None, out_dir=None, end_idx=-1,
                 noise_bound=0.75, noise_bound_radial=0.1, noise_bound_tangential_deg=9.0):
    """Returns a dict with the relative poses (n-1, 4, 4), the chained trajectory (n, 4, 4), per-pair wall times
    and, with loop_detection, (frame, loop_id, yaw) triples.  Writes est_traj.txt / time_taken.txt / loops.txt
    into out_dir when given (one row per frame: stamp and the row-major 3x4 pose, like writeTraj)."""
    files = radar_files(datadir)
    if end_idx >= 0:
        files = files[:end_idx + 1]
    loader = orora.RadarLoader(dataset, keypoint_extraction, voxel_size, use_voxelization, False, ctx=ctx)
    est = orora.ORORA(dataset, "rot", True, False, True, 700, 0.049, noise_bound_radial,
                      np.deg2rad(noise_bound_tangential_deg), ctx=loader.ctx)
    params = orora.setParams(noise_bound, 1.0, False, 100, 1.4, 1e-6)
    mgr = None
    if loop_detection:
        mgr = scancontext.SCManager(loader.ctx)
        mgr.setSCdistThres(sc_dist_thres)
    rel, traj, times, loops = [], [np.eye(4)], [], []
    prev = None
    for i, name in enumerate(files):
        t0 = time.perf_counter()
        cur = loader.extractFeature(datadir, name)
        if prev is not None:
            pm, cm = loader.matchFeatures(prev, cur)
            T = est.calcTransform(params, 1.0 / frame_rate, pm, cm)
            rel.append(T)
            traj.append(traj[-1] @ T)
            times.append(time.perf_counter() - t0)
        prev = cur
        if mgr is not None:
            # the node publishes the features with z = feat_in_cart_ row 2 (+1.0): odometry.cpp:305-316
            z = cur.feat_in_cart_[2] + 1.0
            cloud = np.column_stack([cur.feat_in_cart_[0], cur.feat_in_cart_[1], z]).astype(np.float32)
            mgr.makeAndSaveScancontextAndKeys(cloud)
            loop_id, yaw = mgr.detectLoopClosureID()
            if loop_id >= 0:
                loops.append((i, loop_id, yaw))
    out = {"files": files, "relative": np.array(rel), "trajectory": np.array(traj), "seconds": np.array(times),
           "loops": loops}
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "est_traj.txt"), "w") as f:
            for name, P in zip(files, traj):
                f.write(name.split(".")[0] + " " + " ".join(f"{v:.9g}" for v in P[:3].reshape(-1)) + "\n")
        with open(os.path.join(out_dir, "time_taken.txt"), "w") as f:
            for s in times:
                f.write(f"{int(s * 1e6)}\n")
        with open(os.path.join(out_dir, "loops.txt"), "w") as f:
            for i, l, y in loops:
                f.write(f"{i} {l} {y:.9g}\n")
    return out
