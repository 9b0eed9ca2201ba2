"""PINNING the ScanContext oracle on the reference itself: oracle/_ref/libref_sc.so is the reference's own
pgo/SC-A-LOAM/include/scancontext/Scancontext.cpp (+ its vendored nanoflann), compiled unmodified from /root/reference
against the stand-in headers of oracle/refshim/ (recipe: oracle/Makefile `ref`).  Every public SCManager function is
compared with oracle/sc_ref.c, bit for bit:

    makeScancontext                      Scancontext.cpp:151-195     (B1)
    makeRingkey/SectorkeyFromScancontext :198-227                    (B2)
    circshift / fastAlignUsingVkey       :39-59, :93-113             (B3, B4)
    distDirectSC                         :69-90                      (B5)
    distanceBtnScanContext               :116-148                    (B6)
    makeAndSave / save / detectLoopClosureID, with the tree rebuilt every 30th call and the newest 30 excluded  :236-260, :331-422   (B7)
    detectLoopClosureIDBetweenSession, one-shot batch tree   :267-328                                                            (B8)

What is compiled is the reference's logic and its nanoflann; Eigen is not in this image, so the stand-in MatrixXd
evaluates norm / dot / mean in the order Eigen 3.3's SSE2 redux uses (recalled; oracle/refshim/Eigen/Dense states it
independently of oracle/sc_ref.c).  Runs on CPU; skipped where the library is absent."""
import importlib

import numpy as np
import pytest

RS = importlib.import_module("oracle.refsc")

pytestmark = pytest.mark.skipif(not RS.available(), reason="oracle/_ref/libref_sc.so not built (needs /root/reference)")


@pytest.fixture(scope="module")
def ref():
    m = RS.SCManager()
    yield m
    m.close()


def clouds(seed, n, pts=900):
    """feature clouds like the odometry node publishes them (z = 1.0), plus LiDAR-like ones with real heights, points on
    bin borders, on the axes, at the origin and beyond the 80 m radius"""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        xy = rng.uniform(-95, 95, (pts, 2)).astype(np.float32)
        z = np.ones(pts, dtype=np.float32) if k % 2 == 0 else rng.uniform(-1.5, 6.0, pts).astype(np.float32)
        c = np.column_stack([xy, z]).astype(np.float32)
        c[:8, :2] = [[0, 0], [4, 0], [0, 4], [-4, 0], [0, -4], [80, 0], [56.568542, 56.568542], [8.0, 8.0]]
        out.append(c)
    return out


def test_make_and_keys(ref, oracle):
    for c in clouds(1, 12) + [np.zeros((0, 3), np.float32), np.array([[500.0, 0, 1]], np.float32)]:
        d = ref.makeScancontext(c)
        assert np.array_equal(d, oracle.sc_make(c))
        r, s = ref.keys(d)
        assert np.array_equal(r, oracle.sc_ringkey(d)) and np.array_equal(s, oracle.sc_sectorkey(d))


def test_pairwise_functions(ref, oracle, synth):
    rng = np.random.default_rng(2)
    for binary in (True, False):
        db, rev = synth.sc_database(60, 12, seed=7, binary=binary)
        pairs = [tuple(rng.integers(0, 60, 2)) for _ in range(120)] + [(r[0], r[1]) for r in rev] + [(5, 5)]
        for i, j in pairs:
            a, b = db[i], db[j]
            assert ref.distDirectSC(a, b) == oracle.sc_dist_direct(a, b)
            assert ref.fastAlignUsingVkey(oracle.sc_sectorkey(a), oracle.sc_sectorkey(b)) == oracle.sc_fast_align(
                oracle.sc_sectorkey(a), oracle.sc_sectorkey(b))
            d, s = ref.distanceBtnScanContext(a, b)
            od, os_ = oracle.sc_distance(a, b)
            assert (d, s) == (od, os_), (binary, i, j)
    # empty descriptors: no effective column -> 0/0 = NaN never wins, the sentinel distance and shift 0 come back
    zero = np.zeros(1200)
    assert np.isnan(ref.distDirectSC(zero, db[0])) and np.isnan(oracle.sc_dist_direct(zero, db[0]))
    assert ref.distanceBtnScanContext(zero, db[0]) == oracle.sc_distance(zero, db[0]) == (1e7, 0)
    assert ref.distanceBtnScanContext(zero, zero) == oracle.sc_distance(zero, zero)
    for s in (0, 1, 17, 59):
        assert np.array_equal(RS.circshift(db[3], 20, 60, s), synth.shift_desc(db[3], s))


def exhaustive_set(oracle, descs, query, first, last, thres):
    _, _, (bd, bi, bs) = oracle.sc_search(query, np.stack(descs), first, last)
    return bd, bi, bs


def test_detect_loop_closure_flow(oracle, synth):
    """200 keyframes, a third of them revisits: the reference's own detectLoopClosureID call by call (tree rebuilt when
    its counter says so, newest 30 keys excluded, 3 nanoflann candidates, strict <) against the oracle's candidate
    evaluation on the candidates of the reference's real nanoflann (oracle/_ref/libref_kdtree.so)."""
    db, rev = synth.sc_database(200, 70, seed=21)
    m = RS.SCManager()
    m.setSCdistThres(0.3)
    keys, counter, searchable = [], 0, 0
    found = 0
    for k in range(200):
        m.saveScancontextAndKeys(db[k])
        d, key = m.get(k)
        assert np.array_equal(d, db[k]) and np.array_equal(key, oracle.sc_ringkey_f32(db[k]))
        keys.append(key)
        loop, yaw = m.detectLoopClosureID()
        if len(keys) < 31:
            assert (loop, yaw) == (-1, 0.0)
            continue
        if counter % 30 == 0:                         # Scancontext.cpp:348-359: the tree is rebuilt on every 30th call
            searchable = len(keys) - 30
        counter += 1
        cand = oracle.ref_knn(np.stack(keys[:searchable]), keys[-1], 3)[0]
        o_loop, o_yaw, o_md, o_nn = oracle.sc_detect_from_candidates(db[k], db[:k + 1], cand, 0.3)
        assert (loop, np.float32(yaw)) == (o_loop, np.float32(o_yaw)), k
        found += loop >= 0
    assert found > 10
    m.close()


def test_detect_between_sessions(oracle, synth):
    """detectLoopClosureIDBetweenSession: the batch tree is built at the FIRST call and never again (keyframes saved
    later are invisible to it), unfilled candidate slots stay 0 when fewer than 3 keys exist."""
    db, rev = synth.sc_database(90, 30, seed=33)
    m = RS.SCManager()
    m.setSCdistThres(0.4)
    for k in range(60):
        m.saveScancontextAndKeys(db[k])
    keys = np.stack([oracle.sc_ringkey_f32(db[k]) for k in range(60)])
    for k in range(60, 90):
        q = db[k]
        loop, yaw = m.detectLoopClosureIDBetweenSession(oracle.sc_ringkey_f32(q), q)
        cand = oracle.ref_knn(keys, oracle.sc_ringkey_f32(q), 3)[0]
        o_loop, o_yaw, _, _ = oracle.sc_detect_from_candidates(q, db[:60], cand, 0.4)
        assert (loop, np.float32(yaw)) == (o_loop, np.float32(o_yaw)), k
        m.saveScancontextAndKeys(q)                   # saved after the first call: never searched
        assert loop < 60
    m.close()
    m = RS.SCManager()
    m.setSCdistThres(0.4)
    m.saveScancontextAndKeys(db[0])
    m.saveScancontextAndKeys(db[1])
    loop, yaw = m.detectLoopClosureIDBetweenSession(oracle.sc_ringkey_f32(db[1]), db[1])
    assert loop == 1 and yaw == 0.0
    m.close()
