"""GPU parity tests of the ScanContext path: CUDA (through the C ABI) vs the CPU oracle on the
same seeded inputs, vs the committed golden vectors, and size-independent properties at the
full 10k-keyframe size.  Bar: shift / index exact, distance bit-exact (binary descriptors: all
partial sums exact; general values: same summation order, no FMA), tolerance 1e-12 stated."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "sc_golden.npz")
TOL = 1e-12


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.fixture(scope="module")
def sc(nrs, ctx):
    return nrs.scancontext


def test_pair_parity(sc, nrs, ctx, oracle, synth):
    m = sc.SCManager(ctx)
    for binary in (True, False):
        db, _ = synth.sc_database(120, 20, seed=21, binary=binary)
        rng = np.random.default_rng(1)
        for _ in range(120):
            i, j = rng.integers(0, 120, 2)
            d, s = m.distanceBtnScanContext(sc.as_desc(db[i]), sc.as_desc(db[j]))
            od, os_ = oracle.sc_distance(db[i], db[j])
            assert s == os_
            assert abs(d - od) <= TOL
            assert bits([d])[0] == bits([od])[0]


def test_operators_parity(sc, ctx, oracle, synth):
    m = sc.SCManager(ctx)
    db, _ = synth.sc_database(12, 0, seed=22, binary=False)
    for i in range(0, 12, 2):
        a, b = db[i], db[i + 1]
        rk, sk = m.makeRingkeyFromScancontext(sc.as_desc(a)), m.makeSectorkeyFromScancontext(sc.as_desc(a))
        assert rk.shape == (20, 1) and sk.shape == (1, 60)
        assert np.array_equal(bits(rk.reshape(-1)), bits(oracle.sc_ringkey(a)))
        assert np.array_equal(bits(sk.reshape(-1)), bits(oracle.sc_sectorkey(a)))
        vb = oracle.sc_sectorkey(b)
        assert m.fastAlignUsingVkey(sk, vb) == oracle.sc_fast_align(sk.reshape(-1), vb)
        assert bits([m.distDirectSC(sc.as_desc(a), sc.as_desc(b))])[0] == bits([oracle.sc_dist_direct(a, b)])[0]
    zero = np.zeros(1200)
    assert np.isnan(m.distDirectSC(sc.as_desc(zero), sc.as_desc(db[0])))
    assert m.distanceBtnScanContext(sc.as_desc(zero), sc.as_desc(db[0])) == (1e7, 0)


@pytest.mark.parametrize("binary", [True, False])
def test_query_all_1k(sc, ctx, oracle, synth, binary):
    """BASELINE config 2: 1 query vs 1k-keyframe DB."""
    db, _ = synth.sc_database(1000, 50, seed=42, binary=binary)
    d = sc.SCDatabase(ctx)
    d.append(db)
    assert len(d) == 1000
    q = db[-1]
    dist, shift, best = d.query_all(q, 0, 999)
    odist, oshift, obest = oracle.sc_search(q, db, 0, 999)
    assert np.array_equal(shift, oshift)
    assert np.max(np.abs(dist - odist)) <= TOL
    assert np.array_equal(bits(dist), bits(odist))
    assert best == obest
    assert d.query(q, 0, 999) == obest
    # sub-ranges, including empty and single
    for a, b in ((0, 0), (5, 6), (100, 731), (998, 999), (0, 1000)):
        _, _, ob = oracle.sc_search(q, db, a, b)
        assert d.query(q, a, b) == ob
    assert np.array_equal(bits(d.read(10, 3)), bits(db[10:13]))
    d.close()


def test_candidates_mode_and_golden(sc, ctx, oracle):
    g = np.load(GOLD)
    for tag in ("bin", "gen"):
        db, q = g[f"{tag}_db"], g[f"{tag}_query"]
        d = sc.SCDatabase(ctx, capacity_hint=4)      # forces growth
        for k in range(0, 96, 7):
            d.append(db[k:k + 7])
        dist, shift, best = d.query_all(q, 0, 95)
        assert np.array_equal(shift, g[f"{tag}_shift"])
        assert np.array_equal(bits(dist), bits(g[f"{tag}_dist"]))
        assert tuple(best) == tuple(g[f"{tag}_best"])
        if f"{tag}_ref_knn_idx" in g:
            cand = g[f"{tag}_ref_knn_idx"]
            loop, yaw, md, nn = oracle.sc_detect_from_candidates(q, db, cand, 0.45)
            gd, gi, gs = d.query(q, candidates=cand)
            assert (gd, gi) == (md, nn)
        # candidate order decides ties: duplicates of the same keyframe
        gd, gi, gs = d.query(q, candidates=[40, 12, 40, 12])
        od = [oracle.sc_distance(q, db[c]) for c in (40, 12)]
        exp = 40 if od[0][0] <= od[1][0] else 12
        assert gi == exp
        d.close()


def test_edge_cases(sc, ctx, oracle, synth):
    db, _ = synth.sc_database(40, 0, seed=23)
    zero = np.zeros(1200)
    d = sc.SCDatabase(ctx)
    d.append(np.stack([zero, zero, db[3], zero]))
    assert d.query(db[3]) == (oracle.sc_distance(db[3], db[3])[0], 2, 0)
    assert d.query(zero) == (1e7, 0, 0)                 # NaN distances never win
    assert d.query(db[3], 0, 2) == (1e7, 0, 0)
    dist, shift, _ = d.query_all(db[3])
    assert dist[0] == 1e7 and shift[0] == 0
    # ties: lowest index wins
    d.clear()
    d.append(np.stack([db[5], db[7], db[5], db[7]]))
    assert d.query(db[7])[1] == 1 and d.query(db[5])[1] == 0
    # sparse descriptors with many empty columns
    sp = db[9].copy().reshape(60, 20)
    sp[::2] = 0
    sp = sp.reshape(-1)
    d.clear()
    d.append(db)
    assert d.query(sp) == oracle.sc_search(sp, db)[2]
    with pytest.raises(Exception):
        d.query(db[0], 0, 41)
    # a candidate list far longer than the database (repeats): first occurrence wins
    cand = np.tile(np.arange(40, dtype=np.int32)[::-1], 200)
    assert d.query(db[3], candidates=cand) == (oracle.sc_distance(db[3], db[3])[0], 3, 0)
    d.close()


def same_result(a, b):
    """(dist, idx, shift) equality where NaN never appears (the search returns 1e7 instead)."""
    return bits([a[0]])[0] == bits([b[0]])[0] and a[1] == b[1] and a[2] == b[2]


def test_filter_cannot_separate_near_duplicates(sc, ctx, oracle, synth):
    """Hundreds of keyframes within 1e-14 .. 1e-7 of the query: far below the f32 filter's
    resolution, so every one of them must reach the f64 certification and the exact arg-min
    (lowest index on ties) must come out."""
    rng = np.random.default_rng(7)
    base, _ = synth.sc_database(64, 0, seed=11, binary=False)
    q = base[5]
    rel = 10.0 ** rng.uniform(-14, -7, size=400)
    near = q[None, :] * (1.0 + rel[:, None] * rng.standard_normal((400, 1200)))
    near[37] = near[311]                                   # exact tie: the lower index must win
    db = np.concatenate([base, near, base[:16]])
    d = sc.SCDatabase(ctx)
    d.append(db)
    odist, oshift, obest = oracle.sc_search(q, db)
    assert same_result(d.query(q), obest)
    dist, shift, best = d.query_all(q)
    assert np.array_equal(bits(dist), bits(odist)) and np.array_equal(shift, oshift) and best == obest
    # the same with the query itself absent from the range (all survivors are near-ties)
    _, _, ob = oracle.sc_search(q, db, 64, 464)
    assert same_result(d.query(q, 64, 464), ob)
    # rotated near-duplicates: the alignment arg-min is decided in f64 as well
    rot = np.stack([synth.shift_desc(near[k], int(s)) for k, s in enumerate(rng.integers(0, 60, 400))])
    d.clear()
    d.append(rot)
    odist, oshift, obest = oracle.sc_search(q, rot)
    assert same_result(d.query(q), obest)
    dist, shift, _ = d.query_all(q)
    assert np.array_equal(bits(dist), bits(odist)) and np.array_equal(shift, oshift)
    d.close()


@pytest.mark.parametrize("scale_db,scale_q", [(1e-200, 1e-200), (1e-160, 1.0), (1e-30, 1e-30), (1e-12, 1e9),
                                               (1e20, 1e20), (1e30, 1e-30), (1e160, 1e160), (1e200, 1.0)])
def test_filter_magnitude_hazards(sc, ctx, oracle, synth, scale_db, scale_q):
    """Magnitudes at which the f32 filter's error bounds do not hold (or f32 over/underflows) must
    fall back to the f64 evaluation: results stay identical to the oracle's."""
    db, rev = synth.sc_database(300, 40, seed=3, binary=False)
    db = db * scale_db
    q = db[rev[0][0]] / scale_db * scale_q
    d = sc.SCDatabase(ctx)
    d.append(db)
    with np.errstate(all="ignore"):
        odist, oshift, obest = oracle.sc_search(q, db)
    dist, shift, best = d.query_all(q)
    assert np.array_equal(dist, odist, equal_nan=True) and np.array_equal(shift, oshift)
    assert same_result(best, obest)
    assert same_result(d.query(q), obest)
    d.close()


def test_filter_mixed_hazards(sc, ctx, oracle, synth):
    """A few poisoned keyframes (NaN / inf entries, one column 1e-170, one column 1e170, tiny sector
    means) inside an otherwise ordinary database must neither win wrongly nor mask the winner."""
    db, rev = synth.sc_database(400, 60, seed=9, binary=False)
    j, i, s = rev[1]
    db = db.copy()
    db[2, 7] = np.nan
    db[3, 100] = np.inf
    db[4].reshape(60, 20)[11] *= 1e-170
    db[6].reshape(60, 20)[40] *= 1e170
    db[8] *= 1e-18
    db[9] = db[j] * (1 + 1e-13)                            # near-duplicate of the query keyframe at a low index
    d = sc.SCDatabase(ctx)
    d.append(db)
    for q in (db[j], db[4], db[6], db[8], db[2]):
        with np.errstate(all="ignore"):
            odist, oshift, obest = oracle.sc_search(q, db, 0, 390)
        dist, shift, best = d.query_all(q, 0, 390)
        assert np.array_equal(dist, odist, equal_nan=True) and np.array_equal(shift, oshift)
        assert same_result(best, obest)
        assert same_result(d.query(q, 0, 390), obest)
    d.close()


def test_make_scancontext_parity(sc, ctx, oracle, synth):
    m = sc.SCManager(ctx)
    g = np.load(GOLD)
    assert np.array_equal(sc.flat(m.makeScancontext(g["cloud"])), g["cloud_desc"])
    for seed in range(6):
        cloud = synth.feature_cloud(500 + 700 * seed, seed=100 + seed, max_range=60 + 20 * seed)
        if seed == 5:
            cloud[:, 2] = np.random.default_rng(0).uniform(-3, 3, cloud.shape[0]).astype(np.float32)
        assert np.array_equal(sc.flat(m.makeScancontext(cloud)), oracle.sc_make(cloud)), seed
    axes = np.array([[1, 0, 1], [0, 1, 1], [-1, 0, 1], [0, -1, 1], [0, 0, 1], [80, 0, 1], [56.5685, 56.5685, 1],
                     [100, 0, 1], [3, 3, -5000]], dtype=np.float32)
    assert np.array_equal(sc.flat(m.makeScancontext(axes)), oracle.sc_make(axes))
    assert np.array_equal(sc.flat(m.makeScancontext(np.zeros((0, 3), np.float32))), np.zeros(1200))


def test_manager_flow_matches_reference_flow(sc, ctx, oracle, synth):
    """makeAndSaveScancontextAndKeys + detectLoopClosureID, keyframe by keyframe, against the
    reference flow restated with the oracle (kd-tree candidates from the reference's own
    nanoflann when oracle/_ref is present)."""
    knn = (lambda keys, q, k: oracle.ref_knn(keys, q, min(k, len(keys)))[0]) if oracle.ref_kdtree() \
        else (lambda keys, q, k: oracle.sc_knn_bruteforce(keys, q, min(k, len(keys)))[0])
    m_ex = sc.SCManager(ctx)                      # exhaustive (north_star)
    m_cd = sc.SCManager(ctx, candidate_fn=knn)    # reference candidate set
    m_ex.setSCdistThres(0.45)
    m_cd.setSCdistThres(0.45)
    rng = np.random.default_rng(5)
    clouds = [synth.feature_cloud(900, seed=200 + i, max_range=75) for i in range(20)]
    seq = [clouds[i % 20] if i < 20 or i % 3 else
           (clouds[(i * 7) % 20] + rng.normal(0, 0.05, clouds[0].shape).astype(np.float32) * [1, 1, 0])
           for i in range(75)]
    o_descs, o_keys, counter, search_end = [], [], 0, 0
    for i, cl in enumerate(seq):
        cl = cl.astype(np.float32)
        m_ex.makeAndSaveScancontextAndKeys(cl)
        m_cd.makeAndSaveScancontextAndKeys(cl)
        o_descs.append(oracle.sc_make(cl))
        o_keys.append(oracle.sc_ringkey_f32(o_descs[-1]))
        assert np.array_equal(sc.flat(m_ex.polarcontexts_[-1]), o_descs[-1])
        assert np.array_equal(m_ex.polarcontext_invkeys_mat_[-1], o_keys[-1])
        r_ex, r_cd = m_ex.detectLoopClosureID(), m_cd.detectLoopClosureID()
        if len(o_descs) < 31:
            assert r_ex == (-1, 0.0) and r_cd == (-1, 0.0)
            continue
        if counter % 30 == 0:
            search_end = len(o_descs) - 30
        counter += 1
        db = np.stack(o_descs)
        cand = knn(np.stack(o_keys[:search_end]), o_keys[-1], 3)
        loop, yaw, md, nn = oracle.sc_detect_from_candidates(o_descs[-1], db, cand, 0.45)
        assert r_cd == (loop, yaw)
        _, _, (bd, bi, bs) = oracle.sc_search(o_descs[-1], db, 0, search_end)
        assert r_ex[0] == (bi if bd < 0.45 else -1)
        assert m_ex.last_min_dist == bd and m_ex.last_nn_idx == bi
        assert bd <= md          # exhaustive is never worse than the 3 candidates


def test_full_size_properties_10k(sc, ctx, oracle, synth):
    """BASELINE config 4 size (10k keyframes) through size-independent properties: planted
    revisits are recovered with the planted shift, self-queries return distance ~0 at the
    lowest duplicate index, a split search merges to the same answer, and a sample of
    per-keyframe results is bit-identical to the oracle."""
    db, rev = synth.sc_database(10000, 500, seed=42)
    d = sc.SCDatabase(ctx)
    d.append(db)
    hits = 0
    for j, i, s in rev[::25]:
        gd, gi, gs = d.query(db[j], 0, j - 30)
        od, os_ = oracle.sc_distance(db[j], db[i])
        assert gd <= od
        if gi == i:
            assert gs == os_ == s
            hits += 1
    assert hits >= 15
    for i in (0, 4999, 9999):
        gd, gi, gs = d.query(db[i])
        assert gi <= i and gd < 1e-12 and (gi != i or gs == 0)
    q = db[rev[3][0]]
    whole = d.query(q, 0, 9000)
    parts = [d.query(q, a, b) for a, b in ((0, 1234), (1234, 5000), (5000, 9000))]
    assert sc.best_of(parts) == whole
    dist, shift, best = d.query_all(q, 0, 10000)
    assert best == d.query(q, 0, 10000)
    idx = np.random.default_rng(0).choice(10000, 300, replace=False)
    for k in idx:
        od, os_ = oracle.sc_distance(q, db[k])
        assert shift[k] == os_ and bits([dist[k]])[0] == bits([od])[0]
    assert np.argmin(np.where(np.isnan(dist), 1e9, dist)) == best[1]
    d.close()


def test_cpp_adapters_end_to_end():
    """SCManager / ORORA through the C++ adapters (navtech-radar-slam_b200/cpp/nrs_adapters.hpp)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "navtech-radar-slam_b200", "cpp", "adapter_check")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "adapters ok" in r.stdout


def test_ringkey_knn_matches_reference_nanoflann(sc, ctx, oracle, synth):
    """The candidate step of detectLoopClosureID on the device (nrs_sc_knn) against (a) the oracle's
    brute force and (b) the reference's own vendored nanoflann compiled into oracle/_ref: same ids in the
    same order, bit-identical squared distances (no exact ties in general-valued keys)."""
    db, _ = synth.sc_database(9000, 300, seed=5, binary=False)
    d = sc.SCDatabase(ctx, capacity_hint=16)                  # grows several times
    d.append(db)
    keys = np.stack([oracle.sc_ringkey_f32(x) for x in db[:9000]])
    ref = oracle.ref_kdtree()
    for qi, (first, last, k) in enumerate([(0, 9000, 3), (0, 8970, 10), (100, 4197, 16), (5, 6, 3), (0, 2, 3), (7, 7, 3)]):
        q = oracle.sc_ringkey_f32(db[(qi * 997 + 13) % 9000])
        idx, d2 = d.knn(q, first, last, k)
        oi, od = oracle.sc_knn_bruteforce(keys[first:last], q, k)
        oi = np.where(oi >= 0, oi + first, -1)
        assert np.array_equal(idx, oi), (first, last, k)
        n = int(np.sum(oi >= 0))
        assert np.array_equal(d2[:n].view(np.uint32), od[:n].view(np.uint32))
        if ref is not None and last - first >= k:
            ri, rd = oracle.ref_knn(keys[first:last], q, k)
            assert np.array_equal(idx, ri + first) and np.array_equal(d2.view(np.uint32), rd.view(np.uint32))
    # exact ties (duplicated keyframes): ascending index among equal distances
    dup = np.concatenate([db[:50], db[:50]])
    d.clear()
    d.append(dup)
    idx, d2 = d.knn(oracle.sc_ringkey_f32(db[7]), 0, 100, 4)
    assert list(idx[:2]) == [7, 57] and d2[0] == d2[1] == 0.0
    d.close()


def test_detect_flow_kdtree_mode(sc, ctx, oracle, synth):
    """SCManager in the reference's own mode (ring-key k-NN + 3 candidates, every step on the device)
    against the oracle fed with the reference nanoflann's candidates over the same searchable set."""
    db, rev = synth.sc_database(400, 120, seed=21, binary=False)
    m = sc.SCManager(ctx, candidate_fn="kdtree")
    ref = oracle.ref_kdtree()
    checked = 0
    for i in range(400):
        m.saveScancontextAndKeys(sc.as_desc(db[i]))
        loop, yaw = m.detectLoopClosureID()
        n = i + 1
        if n < m.NUM_EXCLUDE_RECENT + 1:
            assert loop == -1
            continue
        end = m._searchable_end
        keys = np.stack([oracle.sc_ringkey_f32(x) for x in db[:end]])
        qk = oracle.sc_ringkey_f32(db[i])
        if end >= 3:
            cand = oracle.ref_knn(keys, qk, 3)[0] if ref is not None else oracle.sc_knn_bruteforce(keys, qk, 3)[0]
            assert np.array_equal(m.last_candidates, cand)
            eloop, eyaw, emd, enn = oracle.sc_detect_from_candidates(db[i], db, cand, m.SC_DIST_THRES)
            assert (loop, yaw) == (eloop, eyaw) and m.last_min_dist == emd and m.last_nn_idx == enn
            checked += 1
    assert checked > 300


@pytest.mark.parametrize("binary", [True, False])
def test_detect_knn_batch_equals_single_calls(sc, ctx, oracle, synth, binary):
    """nrs_sc_detect_knn_batch (query = second grid dimension of the k-NN kernel, per-query candidate lists in
    the search kernel) against Q single calls and against the oracle fed with the reference nanoflann's ids;
    {0,c} and general-valued databases, ragged ranges, fewer keyframes than k."""
    db, rev = synth.sc_database(700, 150, seed=33, binary=binary)
    d = sc.SCDatabase(ctx, 1024)
    d.append(db)
    qs = np.ascontiguousarray(db[[rev[i][0] for i in range(40)]] if len(rev) >= 40 else db[:40])
    ref = oracle.ref_kdtree()
    for first, last, k in ((0, 700, 3), (13, 650, 3), (0, 700, 10), (5, 7, 3), (100, 100, 3)):
        resb, candb = d.detect_knn_batch(qs, first, last, k)
        for i in range(len(qs)):
            res1, cand1 = d.detect_knn(qs[i], first, last, k)
            assert resb[i] == res1 and np.array_equal(candb[i], cand1), (first, last, k, i)
        if last - first >= k:
            keys = np.stack([oracle.sc_ringkey_f32(x) for x in db[first:last]])
            for i in range(0, len(qs), 7):
                qk = oracle.sc_ringkey_f32(qs[i])
                cand = (oracle.ref_knn(keys, qk, k)[0] if ref is not None else oracle.sc_knn_bruteforce(keys, qk, k)[0]) + first
                assert np.array_equal(candb[i], cand)
                eloop, eyaw, emd, enn = oracle.sc_detect_from_candidates(qs[i], db, cand, 0.2)
                assert (resb[i][0], resb[i][1]) == (emd, enn)
    d.close()


def test_between_session_flow(sc, ctx, oracle, synth):
    """detectLoopClosureIDBetweenSession (Scancontext.cpp:267-328): the batch tree is made once, over the keys
    stored at the FIRST call; later calls search exactly those keyframes.  Both modes against the oracle: the
    reference's 3 tree candidates (its own nanoflann when oracle/_ref is present) and the exhaustive search."""
    db, rev = synth.sc_database(260, 80, seed=33, binary=True)
    ref = oracle.ref_kdtree()
    for mode in ("kdtree", None):
        m = sc.SCManager(ctx, candidate_fn=mode)
        m.setSCdistThres(0.45)
        for i in range(200):
            m.saveScancontextAndKeys(sc.as_desc(db[i]))
        keys = np.stack([oracle.sc_ringkey_f32(x) for x in db[:200]])
        for step, qi in enumerate([210, 220, 5, 250, 231]):
            if step == 2:                                   # saved AFTER the batch tree was made: never searched
                for i in range(200, 240):
                    m.saveScancontextAndKeys(sc.as_desc(db[i]))
            q = db[qi]
            qk = oracle.sc_ringkey_f32(q)
            got = m.detectLoopClosureIDBetweenSession(qk, sc.as_desc(q))
            if mode == "kdtree":
                cand = oracle.ref_knn(keys, qk, 3)[0] if ref is not None else oracle.sc_knn_bruteforce(keys, qk, 3)[0]
                assert np.array_equal(m.last_candidates, cand)
                loop, yaw, md, nn = oracle.sc_detect_from_candidates(q, db, cand, 0.45)
                assert got == (loop, yaw)
            else:
                _, _, (bd, bi, bs) = oracle.sc_search(q, db, 0, 200, want_all=False)
                assert got == ((bi if bd < 0.45 else -1), m._yaw(bs))
            assert got[0] < 200
    # fewer stored keyframes than candidates: unfilled ids stay 0 like the reference's zero-initialised vector
    m = sc.SCManager(ctx, candidate_fn="kdtree")
    m.setSCdistThres(0.45)
    m.saveScancontextAndKeys(sc.as_desc(db[3]))
    assert m.detectLoopClosureIDBetweenSession(oracle.sc_ringkey_f32(db[3]), sc.as_desc(db[3]))[0] == 0
    assert list(m.last_candidates) == [0, 0, 0]
