"""CPU tests of the ScanContext oracle (oracle/sc_ref.c): against an independent numpy
re-derivation of the reference formulas, against the reference's own nanoflann (oracle/_ref)
and against the committed golden vectors."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sc_golden.npz")
R, S = 20, 60


def np_sectorkey(d):
    return d.reshape(S, R).mean(axis=1)


def np_distance(q, k):
    """distanceBtnScanContext written straight from SC/Scancontext.cpp:69-148 with numpy
    (naive summation order) -- independent of the C restatement's Eigen-order bookkeeping."""
    Q, K = q.reshape(S, R), k.reshape(S, R)
    vq, vk = Q.mean(axis=1), K.mean(axis=1)
    norms = [np.linalg.norm(vq - np.roll(vk, s)) for s in range(S)]
    s0 = int(np.argmin(norms))
    best, arg = 1e7, 0
    for s in sorted({(s0 + d) % S for d in range(-3, 4)}):
        Ks = np.roll(K, s, axis=0)
        nq, nk = np.linalg.norm(Q, axis=1), np.linalg.norm(Ks, axis=1)
        eff = (nq != 0) & (nk != 0)
        if eff.sum() == 0:
            continue
        sim = (Q[eff] * Ks[eff]).sum(axis=1) / (nq[eff] * nk[eff])
        d = 1.0 - sim.sum() / eff.sum()
        if d < best:
            best, arg = d, s
    return best, arg, norms


def test_distance_vs_numpy(oracle, synth):
    for binary in (True, False):
        db, _ = synth.sc_database(80, 10, seed=11, binary=binary)
        rng = np.random.default_rng(0)
        n_checked = 0
        for _ in range(60):
            i, j = rng.integers(0, 80, 2)
            d, s = oracle.sc_distance(db[i], db[j])
            nd, ns, norms = np_distance(db[i], db[j])
            # the numpy argmin can legitimately differ on (near-)ties of the key norm
            srt = np.sort(norms)
            if srt[1] - srt[0] < 1e-9:
                continue
            assert s == ns
            assert abs(d - nd) <= 1e-12
            n_checked += 1
        assert n_checked > 30


def test_shifted_copy_is_found(oracle, synth):
    db, _ = synth.sc_database(40, 0, seed=5)
    for i, s in ((0, 0), (3, 7), (10, 59), (20, 31)):
        q = synth.shift_desc(db[i], s)
        d, sh = oracle.sc_distance(q, db[i])
        assert sh == s and d < 1e-12


def test_nan_and_empty_semantics(oracle, synth):
    db, _ = synth.sc_database(8, 0, seed=1)
    zero = np.zeros(1200)
    d, s = oracle.sc_distance(zero, db[0])           # no effective column -> NaN -> (1e7, 0)
    assert d == 1e7 and s == 0
    assert np.isnan(oracle.sc_dist_direct(zero, db[0]))
    _, _, best = oracle.sc_search(db[0], np.stack([zero, zero]), 0, 2)
    assert best == (1e7, 0, 0)
    _, _, best = oracle.sc_search(db[0], db, 3, 3)     # empty range
    assert best == (1e7, 0, 0)


def test_search_lowest_index_on_ties(oracle, synth):
    db, _ = synth.sc_database(16, 0, seed=2)
    dup = np.stack([db[3], db[5], db[3], db[5], db[3]])
    for th in (1, 4):
        _, _, best = oracle.sc_search(db[3], dup, 0, 5, threads=th)
        assert best[1] == 0
        _, _, best = oracle.sc_search(db[5], dup, 0, 5, threads=th)
        assert best[1] == 1


def test_make_scancontext_vs_numpy(oracle, synth):
    cloud = synth.feature_cloud(2000, seed=9)
    desc = oracle.sc_make(cloud).reshape(S, R)
    x, y = cloud[:, 0].astype(np.float64), cloud[:, 1].astype(np.float64)
    rng_ = np.sqrt(x * x + y * y)
    th = np.degrees(np.arctan2(y, x)) % 360.0
    keep = rng_ <= 80.0
    ring = np.clip(np.ceil(rng_ / 80.0 * R), 1, R).astype(int) - 1
    sect = np.clip(np.ceil(th / 360.0 * S), 1, S).astype(int) - 1
    ref = np.zeros((S, R))
    ref[sect[keep], ring[keep]] = 3.0
    # float32 rounding may move a point sitting on a bin edge: allow a handful of bins
    assert (desc != ref).sum() <= 4
    assert set(np.unique(desc)) <= {0.0, 3.0}


def test_keys(oracle, synth):
    db, _ = synth.sc_database(4, 0, seed=3, binary=False)
    d = db[1]
    assert np.allclose(oracle.sc_sectorkey(d), np_sectorkey(d), rtol=0, atol=1e-15)
    assert np.allclose(oracle.sc_ringkey(d), d.reshape(S, R).mean(axis=0), rtol=0, atol=1e-15)
    assert oracle.sc_ringkey_f32(d).dtype == np.float32


def test_knn_metric_matches_reference_nanoflann(oracle):
    """Pins the candidate step on the REFERENCE's own kd-tree (vendored nanoflann compiled from
    /root/reference into oracle/_ref): same ids, bit-identical squared distances."""
    if oracle.ref_kdtree() is None:
        pytest.skip("oracle/_ref/libref_kdtree.so not built (no /root/reference here)")
    rng = np.random.default_rng(4)
    for n in (1, 2, 3, 10, 11, 200, 1500):
        keys = (rng.random((n, 20)) * 1.2).astype(np.float32)
        for _ in range(5):
            q = (rng.random(20) * 1.2).astype(np.float32)
            k = min(3, n)
            ridx, rd2 = oracle.ref_knn(keys, q, k)
            bidx, bd2 = oracle.sc_knn_bruteforce(keys, q, k)
            assert np.array_equal(ridx, bidx)
            assert np.array_equal(rd2.view(np.uint32), bd2.view(np.uint32))


def test_golden(oracle):
    g = np.load(GOLD)
    for tag in ("bin", "gen"):
        dist, shift, best = oracle.sc_search(g[f"{tag}_query"], g[f"{tag}_db"], 0, 95)
        assert np.array_equal(shift, g[f"{tag}_shift"])
        assert np.array_equal(dist.view(np.uint64), g[f"{tag}_dist"].view(np.uint64))
        assert tuple(best) == tuple(g[f"{tag}_best"])
        keys = np.stack([oracle.sc_ringkey_f32(d) for d in g[f"{tag}_db"]])
        assert np.array_equal(keys, g[f"{tag}_ringkeys_f32"])
        if f"{tag}_ref_knn_idx" in g:
            bidx, bd2 = oracle.sc_knn_bruteforce(keys[:66], keys[-1], 3)
            assert np.array_equal(bidx, g[f"{tag}_ref_knn_idx"])
            assert np.array_equal(bd2, g[f"{tag}_ref_knn_d2"])
    assert np.array_equal(oracle.sc_make(g["cloud"]), g["cloud_desc"])
