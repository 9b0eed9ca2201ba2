"""ROS-free driver over a directory of "oxford-form" polar PNGs: the loop of
outlier-robust-radar-odometry/src/odometry.cpp:153-393 (extractFeature -> matchFeatures -> ORORA::calcTransform,
poses chained as curr_odom = curr_odom * T) plus, optionally, what the SC-A-LOAM back end does with the clouds
that node publishes (one ScanContext keyframe per frame, detectLoopClosureID against the growing database;
pgo/SC-A-LOAM/src/laserPosegraphOptimization.cpp:486-495,553-577).  Ground-truth handling, rviz topics and the
pose-graph optimisation itself are not part of this (SURVEY.md 8f rank 3 / out of scope)."""
import os
import time

import numpy as np

from . import orora, posemanager, scancontext


def stamp_of(name):
    """getTimestamp / less_than_img (radar_utils.cpp:70-83): the integer before the first '.'"""
    return int(name.split(".")[0])


def radar_files(datadir):
    """get_file_names (radar_utils.cpp:85-102): the directory's regular files whose last '.'-part is "png", in
    ascending order of the time stamp in the name."""
    return sorted((f for f in os.listdir(datadir) if f.split(".")[-1] == "png" and os.path.isfile(os.path.join(datadir, f))),
                  key=stamp_of)


def run_odometry(datadir, dataset="mulran", keypoint_extraction="cen2018", voxel_size=0.6, use_voxelization=True,
                 frame_rate=4.0, loop_detection=False, sc_dist_thres=0.2, ctx=None, out_dir=None, end_idx=-1,
                 noise_bound=0.75, noise_bound_radial=0.1, noise_bound_tangential_deg=9.0, gt_rel=None,
                 check_stop_motion=True, seq_dir=None):
    """n files give n - 2 scan pairs, like the node (below).  Returns a dict with the relative poses (n-2, 4, 4),
    the chained trajectory (n-1, 4, 4: Identity first), per-pair wall times
    and, with loop_detection, (frame, loop_id, yaw) triples.  With out_dir, writes the reference node's own
    evaluation files (posemanager.SE2PoseManager: <dataset>_ORORA_<extraction>_{eval_odom, rel_error_odom,
    time_taken, est_traj, gt_traj}.txt, scorable by ORORA/scripts/evaluate_odometry.py) plus loops.txt.
    gt_rel: per pair (x, y, yaw) relative ground-truth motion as the reference's getXYYawPose returns it
    (odometry.cpp:253-259); None writes identity ground truth.
    seq_dir (datadir = None): the node's own layout -- frames from <seq_dir>/polar_oxford_form/, ground truth from
    <seq_dir>/gt/radar_odometry.csv through SE2PoseManager.getXYYawPose, files into <seq_dir>/outputs/."""
    if seq_dir is not None:
        if datadir is not None or out_dir is not None or gt_rel is not None:
            raise ValueError("seq_dir fixes the frame directory, the ground truth and the output directory")
        datadir, out_dir = os.path.join(seq_dir, "polar_oxford_form"), os.path.join(seq_dir, "outputs")
    files = radar_files(datadir)
    # the node stops one file short of the list -- pair i is written with the stamps of files i and i + 1
    # (odometry.cpp:157-158: `if (end_idx < 0) end_idx = radar_files.size() - 1; for (i = start_idx; i < end_idx; ++i)`,
    # :180-185) -- so n files give n - 2 scan pairs
    end_idx = len(files) - 1 if end_idx < 0 else min(end_idx, len(files) - 1)
    next_stamp = [stamp_of(f) for f in files[1:end_idx + 1]]
    files = files[:max(end_idx, 0)]
    loader = orora.RadarLoader(dataset, keypoint_extraction, voxel_size, use_voxelization, False, ctx=ctx)
    est = orora.ORORA(dataset, "rot", True, False, check_stop_motion, 700, 0.049, noise_bound_radial,
                      np.deg2rad(noise_bound_tangential_deg), ctx=loader.ctx)
    params = orora.setParams(noise_bound, 1.0, False, 100, 1.4, 1e-6)
    mgr = None
    if loop_detection:
        mgr = scancontext.SCManager(loader.ctx)
        mgr.setSCdistThres(sc_dist_thres)
    rel, traj, times, loops, skipped = [], [np.eye(4)], [], [], []
    if seq_dir is not None:
        pm_out = posemanager.SE2PoseManager.from_seq_dir(seq_dir, dataset, "ORORA", keypoint_extraction)      # odometry.cpp:143-144
    else:
        pm_out = posemanager.SE2PoseManager(out_dir, dataset, "ORORA", keypoint_extraction) if out_dir else None
    curr_gt = np.eye(4)
    prev = None
    for i, name in enumerate(files):
        cur = loader.extractFeature(datadir, name)
        if prev is not None:
            pm, cm = loader.matchFeatures(prev, cur)
            t0 = time.perf_counter()                                   # the node times the estimator only (odometry.cpp:197-251)
            try:
                T = est.calcTransform(params, 1.0 / frame_rate, pm, cm)
            except orora._lib.NrsError as e:
                # more correspondences than the solver holds (1024; only reachable with check_stop_motion off): the
                # pair is reported and skipped with an Identity motion, like the reference's stop-motion gate
                if e.code != orora._lib.NRS_ECAPACITY:
                    raise
                T = np.eye(4)
                skipped.append(i)
            dt = time.perf_counter() - t0
            rel.append(T)
            traj.append(posemanager.mul4(traj[-1], T))                 # curr_odom = curr_odom * Test (odometry.cpp:264)
            times.append(dt)
            if pm_out is not None:
                if seq_dir is not None:                                # odometry.cpp:253-259
                    xyyaw = (pm_out.getXYYawPose(stamp_of(name), next_stamp[i]) if dataset == "oxford"
                             else pm_out.getXYYawPose(stamp_of(name)))
                    Tgt = posemanager.xyyaw2se2(xyyaw, dataset)
                else:
                    Tgt = posemanager.xyyaw2se2(gt_rel[i - 1], dataset) if gt_rel is not None else np.eye(4)
                curr_gt = posemanager.mul4(curr_gt, Tgt)
                pm_out.writePoseError(Tgt, T)                          # odometry.cpp:281-285
                pm_out.writePose(Tgt, T, stamp_of(name), next_stamp[i], loader.time_unit_denominator_)   # time1, time2 (:183-185)
                pm_out.writeTraj(traj[-1], pm_out.est_traj_path_)
                pm_out.writeTraj(curr_gt, pm_out.gt_traj_path_)
                pm_out.writeTimeTaken(int(dt * 1e6), Tgt)
        prev = cur
        if mgr is not None:
            # the node zeroes z and publishes the local cloud with z = 1.0 whatever feat_in_cart_'s third row holds
            # (eigen2pcl(..., true) then p.z = 1.0: odometry.cpp:305-316)
            cloud = np.column_stack([cur.feat_in_cart_[0], cur.feat_in_cart_[1], np.ones(cur.feat_in_cart_.shape[1])]).astype(np.float32)
            mgr.makeAndSaveScancontextAndKeys(cloud)
            loop_id, yaw = mgr.detectLoopClosureID()
            if loop_id >= 0:
                loops.append((i, loop_id, yaw))
    out = {"files": files, "relative": np.array(rel), "trajectory": np.array(traj), "seconds": np.array(times),
           "loops": loops, "skipped_pairs": skipped}
    if out_dir:
        with open(os.path.join(out_dir, "loops.txt"), "w") as f:
            for i, l, y in loops:
                f.write(f"{i} {l} {y:.9g}\n")
        out["eval_odom"] = pm_out.output_pose_path_
    return out
