"""GPU stage-parity of the ORORA front-end against the cv2-driven oracle (oracle/frontend_ref.py):
every stage up to the matcher is compared exactly (keypoint list, image bytes, descriptor bytes,
match list), the pose within north_star's 1e-4 m / 1e-5 rad."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F(oracle):
    return importlib.import_module("oracle.frontend_ref")


@pytest.fixture(scope="module")
def world(synth):
    return synth.radar_world(seed=1234)


def check_frame(nrs, F, raw, dataset, loader):
    o = F.extract_feature(raw, dataset)
    g = loader.extractFeatureFromRows(raw, 1234)
    info = g.handle.info
    st = o["stages"]
    R, C = o["fft"].shape
    gs = g.handle.stages(R, C)
    assert np.array_equal(gs["fft"], o["fft"])                                  # load_radar
    assert info.maxg == np.float32(st["maxg"]) and info.mean == st["mean"]
    assert info.mean_h == st["mean_h"]
    assert info.n_candidates == st["n_cand"]
    assert info.j_consumed == st["j_consumed"] and info.l_regions == st["l_regions"]
    assert np.array_equal(gs["R"], st["R"])                                     # marked matrix
    assert np.array_equal(g.feat_in_polar_, o["polar"])                         # cen2019 keypoints
    cart_o = F.polar_to_cartesian_radar_img(o["azimuths"], o["fft"], loader.params.radar_resolution,
                                            loader.params.cart_resolution, 964, True,
                                            loader.params.navtech_version, as_u8=False)
    assert np.array_equal(gs["cart_f32"], cart_o)                               # cv::remap
    assert np.array_equal(g.img_, o["img"])                                     # convertTo u8
    assert np.array_equal(gs["cart_pre"], o["cart_pre_voxel"][:2])              # polar_to_cartesian_points
    assert info.K == o["feat_in_cart"].shape[1]
    assert np.array_equal(g.feat_in_cart_, o["feat_in_cart"])                   # voxelize + convert_to_bev
    assert np.array_equal(g.kp_, o["kp"]) and info.K_desc == len(o["kp"])     # kp_ after cv::ORB's border filter
    assert np.array_equal(g.t_, o["t"])
    assert np.array_equal(g.desc_, o["desc"])                                   # ORB
    return o, g


@pytest.mark.parametrize("dataset", ["mulran", "oxford"])
def test_extract_match_solve_parity(nrs, ctx, oracle, synth, F, world, dataset):
    """BASELINE configs[0] (MulRan 400x3360) and configs[2] (Oxford-format 400x3768): two frames of the
    same world related by a known SE(2); every stage vs the oracle, then the pose."""
    loader = nrs.orora.RadarLoader(dataset, "cen2019", 0.6, True, False, ctx=ctx)
    raw0 = synth.radar_frame(world, (0, 0, 0), seed=1234, dataset=dataset)
    raw1 = synth.radar_frame(world, (2.5, 0.1, np.deg2rad(1.0)), seed=1235, dataset=dataset)
    o0, g0 = check_frame(nrs, F, raw0, dataset, loader)
    o1, g1 = check_frame(nrs, F, raw1, dataset, loader)
    assert g0.handle.info.K > 300
    pm_o, cm_o, q_o, t_o = F.match_frames(o0, o1)
    pm, cm = loader.matchFeatures(g0, g1)
    assert np.array_equal(pm.query_idx, q_o) and np.array_equal(pm.train_idx, t_o)     # match list
    assert np.array_equal(pm.feat_in_cart_, pm_o) and np.array_equal(cm.feat_in_cart_, cm_o)
    est = nrs.orora.ORORA(dataset, "rot", True, False, True, 700, 0.049, 0.1, np.deg2rad(9.0), ctx=ctx)
    params = nrs.orora.setParams(0.75, 1.0, False, 100, 1.4, 1e-6)
    T = est.calcTransform(params, 0.25, pm, cm)
    ro, _ = oracle.orora_calc_transform(pm_o, cm_o, 0.0)
    assert np.array_equal(est.last_result["clique"], ro["clique"])
    dth = np.arctan2(T[1, 0], T[0, 0]) - np.arctan2(ro["T"][1, 0], ro["T"][0, 0])
    assert np.hypot(T[0, 3] - ro["T"][0, 3], T[1, 3] - ro["T"][1, 3]) <= 1e-4 and abs(dth) <= 1e-5
    # and the answer is close to the planted motion (loose: on the oxford frame cv::ORB drops one
    # border keypoint, and the reference's kp_/feat_in_cart_ misalignment that follows -- reproduced
    # here on purpose -- turns the later matches into 0.6 m-level outliers)
    assert abs(np.arctan2(T[1, 0], T[0, 0]) - np.deg2rad(1.0)) < 2e-2
    assert abs(T[0, 3] - 2.5) < 0.5 and abs(T[1, 3] - 0.1) < 0.5
    if dataset == "oxford":
        assert g0.handle.info.K_desc < g0.handle.info.K or g1.handle.info.K_desc < g1.handle.info.K


def test_remap_get_azimuth_index_branch(nrs, ctx, synth, F, world):
    """RadarLoader::extractFeature never hands polar_to_cartesian_radar_img a navtech_version (radar_loader.hpp:105-107),
    so the node's image uses the uniform-step mapping for MulRan too -- the default of nrs_extract_default_params and what
    every other test here runs.  A caller may still ask for the CIR204 branch (get_azimuth_index over the frame's own,
    non-uniform azimuths, radar_utils.cpp:227-250,293): same kernel, the other mapping, against the oracle."""
    loader = nrs.orora.RadarLoader("mulran", "cen2019", 0.6, True, False, ctx=ctx)
    assert loader.params.navtech_version == F.CTS350 and loader.params.range_bins == 3360
    raw = synth.radar_frame(world, (1.0, 0.2, 0.03), seed=4242, dataset="mulran")
    _, az, _, fft = F.load_radar(raw, F.CIR204)
    uniform = loader.extractFeatureFromRows(raw, 1).handle.stages(*fft.shape)["cart_f32"]
    assert np.array_equal(uniform, F.polar_to_cartesian_radar_img(az, fft, loader.params.radar_resolution,
                                                                   loader.params.cart_resolution, 964, True, F.CTS350, as_u8=False))
    loader.params.navtech_version = F.CIR204
    g = loader.extractFeatureFromRows(raw, 2)
    indexed = g.handle.stages(*fft.shape)["cart_f32"]
    want = F.polar_to_cartesian_radar_img(az, fft, loader.params.radar_resolution, loader.params.cart_resolution, 964, True,
                                          F.CIR204, as_u8=False)
    assert np.array_equal(indexed, want) and not np.array_equal(indexed, uniform)
    assert np.array_equal(g.img_, F.polar_to_cartesian_radar_img(az, fft, loader.params.radar_resolution,
                                                                 loader.params.cart_resolution, 964, True, F.CIR204, as_u8=True))


def test_cen2019_operator_small_and_adversarial(nrs, ctx, F):
    """cen2019features as a stand-alone operator on small frames: random, heavily tied (few grey
    levels), nearly empty, and one where fewer than max_points regions exist."""
    rng = np.random.default_rng(7)
    cases = []
    cases.append(((rng.integers(0, 256, (40, 500)) / 255.0).astype(np.float32), 300, 58))
    cases.append(((rng.integers(0, 4, (64, 700)) * 60 / 255.0).astype(np.float32), 500, 58))      # ties
    f = np.zeros((32, 450), dtype=np.float32)
    f[5, 100:104] = 0.9
    f[6, 101:103] = 0.8
    f[20, 449] = 1.0
    cases.append((f, 10000, 58))
    cases.append(((rng.random((400, 600)) ** 3).astype(np.float32), 2000, 10))                      # cols > rows quirk
    cases.append(((rng.integers(0, 256, (16, 3360)) / 255.0).astype(np.float32), 10000, 58))
    for fft, mp, mr in cases:
        fft[:, :42] = 0
        want = F.cen2019features(fft, mp, mr)
        got = nrs.orora.cen2019features(ctx, fft, mp, mr)
        assert got.shape == want.shape and np.array_equal(got, want)


def test_matcher_ties_and_edges(nrs, ctx, F, synth, world):
    loader = nrs.orora.RadarLoader("mulran", "cen2019", 0.6, False, False, ctx=ctx)     # no voxelisation
    raw0 = synth.radar_frame(world, (0, 0, 0), seed=99)
    o = F.extract_feature(raw0, "mulran", use_voxelization=False)
    g = loader.extractFeatureFromRows(raw0, 7)
    assert np.array_equal(g.feat_in_cart_, o["feat_in_cart"]) and g.feat_in_cart_[2].min() == 1.0
    assert np.array_equal(g.desc_, o["desc"])
    # a frame matched against itself: every keypoint's best is itself (distance 0); duplicates tie -> lowest index
    pm, cm = loader.matchFeatures(g, g)
    q_o, t_o = F.match_features(o["desc"], o["desc"])
    assert np.array_equal(pm.query_idx, q_o) and np.array_equal(pm.train_idx, t_o)
    with pytest.raises(ValueError):
        nrs.orora.RadarLoader("kitti", "cen2019")
    with pytest.raises(ValueError):
        nrs.orora.RadarLoader("mulran", "orb")


def test_cen2018_operator_stagewise(nrs, ctx, F, synth, world):
    """cen2018features (features.cpp:44-131, the extractor run_orora.launch selects) as a stand-alone operator:
    the per-pixel response y and the keypoint list against the oracle (reference loops in C, the Gaussian
    filter2D by the real cv2).  y is compared bit for bit, every pixel: exp() -- the one libm call on the
    per-pixel path, not correctly rounded by either glibc or CUDA -- is narrowed to float through
    exp_to_float on both sides (double-double re-evaluation wherever the platform's last bit could reach
    the float), so the response, the threshold decisions and the keypoints do not depend on the libm."""
    rng = np.random.default_rng(5)
    frames = []
    raw = synth.radar_frame(world, (0, 0, 0), seed=1234, dataset="mulran")
    frames.append((F.load_radar(raw, F.CIR204)[3], 3.0, 17, 58))
    raw = synth.radar_frame(world, (1.0, 0.5, 0.1), seed=99, dataset="oxford")
    frames.append((F.load_radar(raw, F.CTS350)[3], 3.0, 17, 58))
    f = (rng.integers(0, 24, (37, 701)) / 255.0).astype(np.float32)
    hot = rng.random(f.shape) < 0.03
    f[hot] = (rng.integers(120, 256, int(hot.sum())) / 255.0).astype(np.float32)
    f[:, :42] = 0
    frames.append((f, 2.0, 5, 10))                                   # odd sizes, short filter, ragged tiles, scalar filter tail
    f = np.zeros((9, 300), dtype=np.float32)
    f[3, 100:160] = 1.0
    f[4, 290:] = 0.7                                                 # run open at the last column
    f[5] = 0.25                                                      # constant row: q == 0, sigma_q = 0.034
    frames.append((f, 1.0, 17, 58))
    n_total = 0
    for fft, zq, sg, mr in frames:
        st = F.cen2018_stages(fft, zq, sg, mr)
        got, y = nrs.orora.cen2018features(ctx, fft, zq, sg, mr, return_response=True)
        assert np.array_equal(y.view(np.uint32)[:, mr:], st["y"].view(np.uint32)[:, mr:])
        want = st["targets"]
        assert got.shape == want.shape and np.array_equal(got, want)
        n_total += want.shape[1]
    assert n_total > 3000


@pytest.mark.parametrize("dataset", ["mulran"])
def test_extract_with_cen2018_parity(nrs, ctx, oracle, synth, F, world, dataset):
    """RadarLoader("...", "cen2018"): the whole extractFeature chain with the cen2018 keypoints, then match and
    solve, against the oracle chain."""
    loader = nrs.orora.RadarLoader(dataset, "cen2018", 0.6, True, False, ctx=ctx)
    res = []
    for pose, seed in (((0, 0, 0), 1234), ((2.5, 0.1, np.deg2rad(1.0)), 1235)):
        raw = synth.radar_frame(world, pose, seed=seed, dataset=dataset)
        o = F.extract_feature(raw, dataset, keypoint_extraction="cen2018")
        g = loader.extractFeatureFromRows(raw, 77)
        assert np.array_equal(g.feat_in_polar_, o["polar"])
        assert np.array_equal(g.img_, o["img"])
        assert np.array_equal(g.feat_in_cart_, o["feat_in_cart"])
        assert np.array_equal(g.kp_, o["kp"]) and np.array_equal(g.t_, o["t"])
        assert np.array_equal(g.desc_, o["desc"])
        res.append((o, g))
    (o0, g0), (o1, g1) = res
    pm_o, cm_o, q_o, t_o = F.match_frames(o0, o1)
    pm, cm = loader.matchFeatures(g0, g1)
    assert np.array_equal(pm.query_idx, q_o) and np.array_equal(pm.train_idx, t_o)
    est = nrs.orora.ORORA(dataset, "rot", True, False, True, 700, 0.049, 0.1, np.deg2rad(9.0), ctx=ctx)
    params = nrs.orora.setParams(0.75, 1.0, False, 100, 1.4, 1e-6)
    T = est.calcTransform(params, 0.25, pm, cm)
    ro, _ = oracle.orora_calc_transform(pm_o, cm_o, 0.0)
    dth = np.arctan2(T[1, 0], T[0, 0]) - np.arctan2(ro["T"][1, 0], ro["T"][0, 0])
    assert np.hypot(T[0, 3] - ro["T"][0, 3], T[1, 3] - ro["T"][1, 3]) <= 1e-4 and abs(dth) <= 1e-5


def test_cen2018_against_committed_golden(nrs, ctx):
    """The CUDA path against the fixtures generated in the build container (tests/golden/make_cen2018_golden.py):
    keypoints exactly, the response within the same <= 4-pixels-one-ulp allowance as above."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cen2018_golden.npz"))
    for tag in ("a", "b"):
        zq, sg, mr = g[f"{tag}_par"]
        got, y = nrs.orora.cen2018features(ctx, g[f"{tag}_fft"], float(zq), int(sg), int(mr), return_response=True)
        diff = y.view(np.uint32)[:, int(mr):] != g[f"{tag}_y"].view(np.uint32)[:, int(mr):]
        assert diff.sum() <= 4
        assert np.array_equal(got, g[f"{tag}_targets"]) and got.shape[1] > 50


def test_cen2019_subset_threshold_sequence(nrs, ctx, F, synth, world):
    """cen2019's top-candidate subset (threshold on h carried from frame to frame): a sequence whose contrast
    changes abruptly -- the carried threshold is far too high (dark frame: the subset cannot vouch, the full sort
    runs), then far too low (bright frame: more than the subset's capacity) -- must give the oracle's keypoints on
    every frame, whatever the history."""
    rng = np.random.default_rng(3)
    raw = synth.radar_frame(world, (0, 0, 0), seed=77)
    base = (raw[:, 11:].astype(np.float32) / 255.0).astype(np.float32)
    base[:, :42] = 0
    frames = [base, (base * 0.25).astype(np.float32), base, np.clip(base * 3.0, 0, 1).astype(np.float32),
              (rng.integers(0, 256, base.shape) / 255.0).astype(np.float32), base]
    frames[4][:, :42] = 0
    for k, fft in enumerate(frames):
        for mp in ((10000, 300) if k in (0, 5) else (10000,)):
            want = F.cen2019features(fft, mp, 58)
            got = nrs.orora.cen2019features(ctx, fft, mp, 58)
            assert got.shape == want.shape and np.array_equal(got, want), (k, mp)


@pytest.mark.parametrize("extraction", ["cen2019", "cen2018"])
def test_fused_pair_equals_separate_calls(nrs, ctx, synth, world, extraction):
    """nrs_orora_pair (extract + match + solve in one call, device-side chaining) against the three separate
    entry points on a 5-frame sequence with the Doppler velocity carried from pair to pair: frames, match
    counts, cliques and poses identical."""
    orr = nrs.orora
    raws = [synth.radar_frame(world, (2.5 * k, 0.1 * k, np.deg2rad(1.0) * k), seed=500 + k) for k in range(5)]
    loader = orr.RadarLoader("mulran", extraction, 0.6, True, False, ctx=ctx)
    sp = orr.default_params()
    # separate calls
    feats = [loader.extractFeatureFromRows(r, k) for k, r in enumerate(raws)]
    sep, vel = [], 0.0
    for k in range(1, 5):
        pm, cm = loader.matchFeatures(feats[k - 1], feats[k])
        r = orr.solve(ctx, pm.feat_in_cart_, cm.feat_in_cart_, sp, vel)
        sep.append((pm.feat_in_cart_.shape[1], r))
        vel = r["T"][0, 3] / 0.25
    # fused
    prev = feats[0].handle
    vel = 0.0
    alive = []
    for k in range(1, 5):
        fr, T, n, info = orr.pair(ctx, prev, raws[k], loader.params, sp, vel, t_criteria=k)
        alive.append(fr)
        n_sep, r = sep[k - 1]
        assert n == n_sep and info.M == r["info"].M and info.n_edges == r["info"].n_edges
        assert np.array_equal(T, r["T"]), k
        d = fr.get(want_img=True)
        assert np.array_equal(d["desc"], feats[k].desc_) and np.array_equal(d["feat_in_cart"], feats[k].feat_in_cart_)
        assert np.array_equal(d["img"], feats[k].img_) and np.array_equal(d["polar"], feats[k].feat_in_polar_)
        vel = T[0, 3] / 0.25
        prev = fr


def test_fused_pair_size_rules(nrs, ctx, synth, world):
    """The fused call enqueues the solver before it knows the match count: the host-side rules must come out the
    same as through nrs_orora_solve -- the stop-motion gate (identical frames: ~1900 matches > 700), the capacity
    error above 1024 matches without the gate, the host-sized second solve for counts between 512 and 1024 (coarse
    voxels on identical frames), and a pair with too few matches (unrelated frame)."""
    orr = nrs.orora
    raw = synth.radar_frame(world, (0.0, 0.0, 0.0), seed=900)
    # 1. stop-motion gate and capacity
    loader = orr.RadarLoader("mulran", "cen2019", 0.6, True, False, ctx=ctx)
    f0 = loader.extractFeatureFromRows(raw, 0)
    pm, cm = loader.matchFeatures(f0, f0)
    n_same = pm.feat_in_cart_.shape[1]
    assert n_same > 1024
    sp = orr.default_params()
    fr, T, n, info = orr.pair(ctx, f0.handle, raw, loader.params, sp, 0.0, t_criteria=1)
    assert n == n_same and info.stop_motion == 1 and np.array_equal(T, np.eye(4))
    sp.check_stop_motion = 0
    with pytest.raises(nrs._lib.NrsError) as e:
        orr.pair(ctx, f0.handle, raw, loader.params, sp, 0.0, t_criteria=1)
    assert e.value.code == nrs._lib.NRS_ECAPACITY
    # 2. between the speculative launch's 512 and the solver's 1024: solved again with the host-sized launch
    for leaf in (2.5, 3.0, 3.5, 4.0, 5.0):
        lc = orr.RadarLoader("mulran", "cen2019", leaf, True, False, ctx=ctx)
        g0 = lc.extractFeatureFromRows(raw, 0)
        pm, cm = lc.matchFeatures(g0, g0)
        n_mid = pm.feat_in_cart_.shape[1]
        if 512 < n_mid <= 1024:
            break
    assert 512 < n_mid <= 1024, n_mid
    r = orr.solve(ctx, pm.feat_in_cart_, cm.feat_in_cart_, sp, 0.0)
    fr, T, n, info = orr.pair(ctx, g0.handle, raw, lc.params, sp, 0.0, t_criteria=1)
    assert n == n_mid and info.M == r["info"].M and info.n_edges == r["info"].n_edges and np.array_equal(T, r["T"])
    # 3. (almost) no matches: a frame of another world
    other = synth.radar_frame(synth.radar_world(seed=4321), (0.0, 0.0, 0.0), seed=901)
    h0 = loader.extractFeatureFromRows(other, 0)
    pm, cm = loader.matchFeatures(f0, h0)
    r = orr.solve(ctx, pm.feat_in_cart_, cm.feat_in_cart_, sp, 0.0)
    fr, T, n, info = orr.pair(ctx, f0.handle, other, loader.params, sp, 0.0, t_criteria=1)
    assert n == pm.feat_in_cart_.shape[1] and info.M == r["info"].M and np.array_equal(T, r["T"])


@pytest.mark.parametrize("extraction", ["cen2019", "cen2018"])
def test_extract_degenerate_frames(nrs, ctx, synth, world, extraction):
    """The fused kernels on frames with nothing (or almost nothing) in them: an all-zero frame gives an empty feature
    set without tripping the staging rows, the in-shared-memory sorts or the cluster sort (zero items, zero passes);
    a frame with one bright blob gives a handful of keypoints; and a normal frame right after is unaffected by the
    threshold the empty frames carried over (the extraction never depends on it)."""
    orr = nrs.orora
    loader = orr.RadarLoader("mulran", extraction, 0.6, True, False, ctx=ctx)
    normal = synth.radar_frame(world, (0.0, 0.0, 0.0), seed=77)
    ref = loader.extractFeatureFromRows(normal, 0)
    blank = normal.copy()
    blank[:, 11:] = 0
    for rep in range(2):                                              # twice: the second run replays the cached graph
        f = loader.extractFeatureFromRows(blank, 1 + rep)
        assert f.feat_in_cart_.shape[1] == 0 and f.desc_.shape[0] == 0
    blob = blank.copy()
    blob[100:104, 11 + 800:11 + 806] = 255
    blob[300:303, 11 + 1500:11 + 1504] = 200
    f = loader.extractFeatureFromRows(blob, 3)
    assert f.feat_in_polar_.shape[1] <= 200 and f.feat_in_cart_.shape[1] <= f.feat_in_polar_.shape[1]
    again = loader.extractFeatureFromRows(normal, 4)
    assert np.array_equal(again.feat_in_polar_, ref.feat_in_polar_) and np.array_equal(again.desc_, ref.desc_)
    assert np.array_equal(again.feat_in_cart_, ref.feat_in_cart_)
