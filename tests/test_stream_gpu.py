"""End-to-end stream on one GPU (the shape of BASELINE configs[4]): a trajectory that drives out and comes
back over its own start; every frame goes through RadarLoader::extractFeature (cen2018, what the shipped launch
runs) -> matchFeatures -> ORORA::calcTransform with the Doppler velocity carried from pair to pair, and every
frame's feature cloud becomes a ScanContext keyframe that is queried against the growing database
(makeAndSaveScancontextAndKeys + detectLoopClosureID).  Compared step by step with the oracle chain: keypoints,
matches and loop decisions exactly, poses within north_star's 1e-4 m / 1e-5 rad."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_odometry_and_loop_stream(nrs, ctx, oracle, synth):
    F = importlib.import_module("oracle.frontend_ref")
    sc = nrs.scancontext
    world = synth.radar_world(seed=77)
    n_frames = 44
    # out along x, a U-turn, back over the start: frames 36.. revisit the places of frames 0..8
    poses = []
    for k in range(n_frames):
        if k < 18:
            poses.append((3.0 * k, 0.0, 0.0))
        elif k < 22:
            a = (k - 17) * np.pi / 5
            poses.append((51.0 + 4.0 * np.sin(a), 4.0 * (1 - np.cos(a)), a))
        else:
            poses.append((51.0 - 3.0 * (k - 21), 8.0 * 0 + 0.3, np.pi))
    loader = nrs.orora.RadarLoader("mulran", "cen2018", 0.6, True, False, ctx=ctx)
    est = nrs.orora.ORORA("mulran", "rot", True, False, True, 700, 0.049, 0.1, np.deg2rad(9.0), ctx=ctx)
    params = nrs.orora.setParams(0.75, 1.0, False, 100, 1.4, 1e-6)
    mgr = sc.SCManager(ctx)
    mgr.setSCdistThres(0.45)
    o_descs, counter, search_end, o_vel = [], 0, 0, 0.0
    prev_g = prev_o = None
    loops = []
    for k in range(n_frames):
        raw = synth.radar_frame(world, poses[k], seed=500 + k, dataset="mulran")
        o = F.extract_feature(raw, "mulran", keypoint_extraction="cen2018")
        g = loader.extractFeatureFromRows(raw, k)
        assert np.array_equal(g.feat_in_polar_, o["polar"])
        assert np.array_equal(g.feat_in_cart_, o["feat_in_cart"]) and np.array_equal(g.desc_, o["desc"])
        if prev_g is not None:
            pm_o, cm_o, q_o, t_o = F.match_frames(prev_o, o)
            pm, cm = loader.matchFeatures(prev_g, g)
            assert np.array_equal(pm.query_idx, q_o) and np.array_equal(pm.train_idx, t_o)
            T = est.calcTransform(params, 0.25, pm, cm)
            ro, o_vel = oracle.orora_calc_transform(pm_o, cm_o, o_vel)
            dth = np.arctan2(T[1, 0], T[0, 0]) - np.arctan2(ro["T"][1, 0], ro["T"][0, 0])
            dth = (dth + np.pi) % (2 * np.pi) - np.pi
            assert np.hypot(T[0, 3] - ro["T"][0, 3], T[1, 3] - ro["T"][1, 3]) <= 1e-4 and abs(dth) <= 1e-5, k
        prev_g, prev_o = g, o
        # the odometry node publishes the features as a cloud with z = 1.0 (ORORA/src/odometry.cpp:305-316)
        cloud = np.column_stack([g.feat_in_cart_[0], g.feat_in_cart_[1], np.ones(g.feat_in_cart_.shape[1])]).astype(np.float32)
        mgr.makeAndSaveScancontextAndKeys(cloud)
        o_descs.append(oracle.sc_make(cloud))
        assert np.array_equal(sc.flat(mgr.polarcontexts_[-1]), o_descs[-1])
        loop, yaw = mgr.detectLoopClosureID()
        if len(o_descs) < 31:
            assert (loop, yaw) == (-1, 0.0)
            continue
        if counter % 30 == 0:
            search_end = len(o_descs) - 30
        counter += 1
        _, _, (bd, bi, bs) = oracle.sc_search(o_descs[-1], np.stack(o_descs), 0, search_end)
        assert mgr.last_min_dist == bd and mgr.last_nn_idx == bi
        assert loop == (bi if bd < 0.45 else -1)
        loops.append((k, loop, bd))
    # the drive back passes the places of the first frames: the detector must say so
    assert any(l >= 0 for _, l, _ in loops), loops


def test_png_directory_driver(nrs, ctx, synth, tmp_path):
    """navtech-radar-slam_b200/stream.py::run_odometry on a directory of PNGs written by cv2 (what the node reads
    with cv::imread): the chained trajectory follows the planted motion and the outputs are written."""
    import cv2
    stream = importlib.import_module("navtech-radar-slam_b200.stream")
    world = synth.radar_world(seed=1234)
    for k in range(7):                                 # n files give n - 2 scan pairs (the node stops one file short)
        raw = synth.radar_frame(world, (2.5 * k, 0.1 * k, np.deg2rad(1.0) * k), seed=900 + k, dataset="mulran")
        assert cv2.imwrite(str(tmp_path / f"{1560000000000000000 + k * 250000000}.png"), raw)
    gt = [(2.5, 0.1, -np.deg2rad(1.0))] * 5          # MulRan-style ground truth: the reference negates its yaw (SE2posemanager.hpp:181-183)
    r = stream.run_odometry(str(tmp_path), "mulran", "cen2019", loop_detection=True, ctx=ctx, out_dir=str(tmp_path / "out"), gt_rel=gt)
    assert r["relative"].shape == (5, 4, 4) and r["trajectory"].shape == (6, 4, 4)
    for T in r["relative"]:                                   # each step: ~2.5 m forward, ~1 degree
        assert abs(T[0, 3] - 2.5) < 0.6 and abs(np.arctan2(T[1, 0], T[0, 0]) - np.deg2rad(1.0)) < 2e-2
    assert abs(r["trajectory"][-1][0, 3] - 12.5) < 2.5
    # the node's own evaluation files (SE2posemanager.hpp:151-245), one row per scan pair
    out = tmp_path / "out"
    ev = np.loadtxt(out / "mulran_ORORA_cen2019_eval_odom.txt", delimiter=",")      # as evaluate_odometry.py reads it
    assert ev.shape == (5, 8) and np.allclose(ev[:, 0], 2.5) and np.all(np.abs(ev[:, 3] - 2.5) < 0.6)
    assert np.all(np.abs(ev[:, 5] - np.deg2rad(1.0)) < 2e-2) and np.all(np.abs(ev[:, 2] - np.deg2rad(1.0)) < 1e-5)
    rel = np.loadtxt(out / "mulran_ORORA_cen2019_rel_error_odom.txt", delimiter=",")
    assert rel.shape == (5, 5) and np.all(rel[:, 0] < 0.6) and np.all(rel[:, 1] < 1.5)
    assert (out / "mulran_ORORA_cen2019_time_taken.txt").read_text().count("\n") == 5
    assert (out / "mulran_ORORA_cen2019_est_traj.txt").read_text().count("\n") == 5
    assert (out / "mulran_ORORA_cen2019_gt_traj.txt").read_text().count("\n") == 5
