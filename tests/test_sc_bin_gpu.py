"""GPU parity tests of the bit-mask ScanContext search (sc_search_bin.cu): {0,c} descriptors -- what this
pipeline produces (ORORA publishes z = 1.0, odometry.cpp:305-316; makeScancontext adds LIDAR_HEIGHT) --
searched from 256-byte mask records.  Everything is compared with the CPU oracle (dist bit-exact, idx and
shift exact) and with the general-valued f32-filter kernel on the same database."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def same(a, b):
    return bits([a[0]])[0] == bits([b[0]])[0] and a[1] == b[1] and a[2] == b[2]


@pytest.fixture(scope="module")
def sc(nrs, ctx):
    return nrs.scancontext


def rand_binary(rng, n, c=3.0, p_lo=0.05, p_hi=0.5):
    p = np.linspace(p_hi, p_lo, 20)
    occ = rng.random((n, 60, 20)) < p[None, None, :]
    return np.where(occ, c, 0.0).reshape(n, 1200)


def test_bin_vs_general_vs_oracle(sc, ctx, oracle, synth):
    db, rev = synth.sc_database(3000, 200, seed=77, binary=True)
    rng = np.random.default_rng(3)
    d = sc.SCDatabase(ctx, capacity_hint=64)
    d.append(db)
    assert d.is_binary()
    queries = [db[j] for (j, _, _) in rev[:12]] + list(rand_binary(rng, 6)) + [db[0], db[2999]]
    sparse = db[10].copy().reshape(60, 20)
    sparse[::3] = 0
    queries.append(sparse.reshape(-1))
    for q in queries:
        _, _, ob = oracle.sc_search(q, db, want_all=False)
        got = d.query(q)
        assert same(got, ob), (got, ob)
        d.set_search_mode(1)
        assert same(d.query(q), ob)
        d.set_search_mode(0)
        for a, b in ((0, 0), (17, 18), (100, 1731), (2999, 3000), (5, 2990)):
            _, _, ob = oracle.sc_search(q, db, a, b, want_all=False)
            assert same(d.query(q, a, b), ob), (a, b)
    cand = rng.integers(0, 3000, 64).astype(np.int32)
    q = queries[0]
    od = [oracle.sc_distance(q, db[c]) for c in cand]
    k = min(range(64), key=lambda i: (od[i][0], i))
    assert same(d.query(q, candidates=cand), (od[k][0], int(cand[k]), od[k][1]))
    d.close()


def test_bin_per_keyframe_parity(sc, ctx, oracle, synth):
    """Single-keyframe ranges make the search return that keyframe's own (dist, shift): per-pair parity of the
    bit-mask path (filter + mask-based f64 certification) against the oracle's distanceBtnScanContext."""
    rng = np.random.default_rng(11)
    db = rand_binary(rng, 400)
    # column-count patterns with many exact ties in the integer alignment score
    for k, period in enumerate((2, 3, 4, 5, 6, 10, 12, 15, 20, 30)):
        m = np.zeros((60, 20))
        for c in range(60):
            m[c, : 1 + (c % period) % 7] = 3.0
        db[k] = m.reshape(-1)
        db[20 + k] = synth.shift_desc(db[k], 7 * k + 1)
    db[50] = 0.0                                             # empty keyframe
    db[51].reshape(60, 20)[::2] = 0                          # half the columns empty
    d = sc.SCDatabase(ctx)
    d.append(db)
    assert d.is_binary()
    for qi in (0, 3, 5, 9, 21, 27, 51, 100, 399):
        q = db[qi]
        res = d.query_batch(np.repeat(q[None], 1, 0))        # Q = 1 through the batch entry
        _, _, ob = oracle.sc_search(q, db, want_all=False)
        assert same(res[0], ob)
        for k in list(range(0, 60)) + list(rng.integers(60, 400, 40)):
            od, os_ = oracle.sc_distance(q, db[k])
            got = d.query(q, int(k), int(k) + 1)
            exp = (od, int(k), os_) if od == od and od < 1e7 else (1e7, 0, 0)
            assert same(got, exp), (qi, k, got, exp)
    d.close()


def test_bin_batch(sc, ctx, oracle, synth):
    db, rev = synth.sc_database(2500, 150, seed=5, binary=True)
    d = sc.SCDatabase(ctx)
    d.append(db)
    qs = np.stack([db[j] for (j, _, _) in rev[:70]])
    exp = [oracle.sc_search(q, db, 0, 2400, want_all=False)[2] for q in qs]
    got = d.query_batch(qs, 0, 2400)
    assert all(same(g, e) for g, e in zip(got, exp))
    assert all(same(d.query(q, 0, 2400), e) for q, e in zip(qs[:10], exp[:10]))
    # more queries than one launch takes (NRS_SC_MAX_BATCH = 1024): split transparently
    small = db[:40]
    d.clear()
    d.append(small)
    rng = np.random.default_rng(2)
    many = small[rng.integers(0, 40, 1100)]
    got = d.query_batch(many)
    for g, q in zip(got[::37], many[::37]):
        assert same(g, oracle.sc_search(q, small, want_all=False)[2])
    assert same(got[1099], oracle.sc_search(many[1099], small, want_all=False)[2])
    # host batches are routed by a test of the first twelve columns only: a batch with visibly general-valued queries
    # takes the general kernel; a query that is general-valued only further on reaches the bit-mask kernel, which
    # classifies every query itself and answers such a one exactly from the canonical descriptors
    mixed = small[rng.integers(0, 40, 300)].copy()
    for k in (5, 100, 299):
        mixed[k] = mixed[k] * rng.uniform(0.5, 1.5, 1200)
    got = d.query_batch(mixed)
    for k in list(range(0, 300, 23)) + [5, 100, 299]:
        assert same(got[k], oracle.sc_search(mixed[k], small, want_all=False)[2]), k
    late = small[rng.integers(0, 40, 200)].copy()
    for k in (0, 77, 199):
        late[k, 700:] = late[k, 700:] * rng.uniform(0.5, 1.5, 500)      # columns 35..59 only
        late[k, :240] = small[3, :240]
    got = d.query_batch(late)
    for k in list(range(0, 200, 19)) + [0, 77, 199]:
        assert same(got[k], oracle.sc_search(late[k], small, want_all=False)[2]), k
    d.close()


def test_bin_scales_and_mixed_databases(sc, ctx, oracle, synth):
    rng = np.random.default_rng(8)
    base = rand_binary(rng, 300, c=1.0)
    cs = np.array([3.0, 1.5, 0.75, 7.0, 3.0 * 2 ** 20, 3.0 * 2 ** -30])
    db = base * cs[rng.integers(0, len(cs), 300)][:, None]
    d = sc.SCDatabase(ctx)
    d.append(db)
    assert d.is_binary()
    for c in (3.0, 1.5, 2.0 ** -12, 5.0 * 2 ** 15):
        for qi in (4, 77, 250):
            q = base[qi] * c
            _, _, ob = oracle.sc_search(q, db, want_all=False)
            assert same(d.query(q), ob), (c, qi)
            got = d.query_batch(np.stack([q, base[qi + 1] * c]))
            assert same(got[0], ob)
    # a value with more than 24 significant bits is not {0,c} for the mask path: the database goes general
    d.append((base[0] * 3.1)[None])
    assert not d.is_binary()
    q = base[5] * 3.0
    db2 = np.concatenate([db, (base[0] * 3.1)[None]])
    assert same(d.query(q), oracle.sc_search(q, db2, want_all=False)[2])
    assert all(same(g, oracle.sc_search(x, db2, want_all=False)[2]) for g, x in zip(d.query_batch(db2[:5]), db2[:5]))
    d.clear()
    assert d.is_binary()
    d.append(db[:50])
    assert same(d.query(q), oracle.sc_search(q, db[:50], want_all=False)[2])
    d.close()


def test_bin_many_survivors_and_ties(sc, ctx, oracle, synth):
    """Hundreds of keyframes within a few bins of the query (most survive the f32 filter), exact duplicates at
    several indices (lowest index wins), rotated copies (the alignment differs per keyframe)."""
    rng = np.random.default_rng(13)
    q = rand_binary(rng, 1)[0]
    rows = []
    for k in range(600):
        x = q.copy()
        flips = rng.integers(0, 1200, k % 4)
        x[flips] = np.where(x[flips] > 0, 0.0, 3.0)
        if k % 3 == 1:
            x = synth.shift_desc(x, int(rng.integers(0, 60)))
        rows.append(x)
    db = np.stack(rows)
    d = sc.SCDatabase(ctx)
    d.append(db)
    for a, b in ((0, 600), (1, 600), (5, 333), (301, 302)):
        _, _, ob = oracle.sc_search(q, db, a, b, want_all=False)
        assert same(d.query(q, a, b), ob), (a, b)
    got = d.query_batch(np.stack([q, db[7], db[100], db[599]]))
    for g, x in zip(got, (q, db[7], db[100], db[599])):
        assert same(g, oracle.sc_search(x, db, want_all=False)[2])
    d.close()


def test_bin_general_query_on_device_path(sc, ctx, oracle, synth):
    """A general-valued query that reaches the bit-mask kernel through a device-pointer entry point (the host
    cannot classify it there) is answered exactly on the kernel's f64 path."""
    import ctypes as C
    import torch
    db, rev = synth.sc_database(500, 60, seed=31, binary=True)
    gen, _ = synth.sc_database(8, 0, seed=32, binary=False)
    d = sc.SCDatabase(ctx)
    d.append(db)
    dev = torch.device("cuda", ctx.device)
    qs = np.stack([gen[0], db[rev[0][0]], gen[3], db[5]])
    q_dev = torch.from_numpy(qs).to(dev)
    res = torch.zeros(16 * 4, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)
    d.query_batch_dev(q_dev.data_ptr(), 4, 0, 490, 1000, res.data_ptr())
    ctx.sync()
    r = res.cpu().numpy().view([("dist", "<f8"), ("idx", "<i4"), ("shift", "<i4")])
    for k in range(4):
        _, _, ob = oracle.sc_search(qs[k], db, 0, 490, want_all=False)
        assert same((float(r["dist"][k]), int(r["idx"][k]), int(r["shift"][k])), (ob[0], ob[1] + 1000, ob[2])), k
    # the host entry point routes the same general query to the general kernel: same answer
    assert same(d.query(gen[0], 0, 490), oracle.sc_search(gen[0], db, 0, 490, want_all=False)[2])
    d.close()


def test_bin_full_size_10k_batch(sc, ctx, oracle, synth):
    """BASELINE size: 10k keyframes, a 64-query batch against the oracle, and batch == single for all."""
    db, rev = synth.sc_database(10000, 500, seed=42, binary=True)
    d = sc.SCDatabase(ctx)
    d.append(db)
    qs = np.stack([db[j] for (j, _, _) in rev[:64]])
    got = d.query_batch(qs)
    for k in range(0, 64, 4):
        assert same(got[k], oracle.sc_search(qs[k], db, want_all=False)[2])
    for k in range(64):
        assert same(got[k], d.query(qs[k]))
    # planted revisits are found with the planted shift when the revisit itself is excluded from the range
    hits = 0
    for (j, i, s) in rev[:40]:
        gd, gi, gs = d.query(db[j], 0, j - 30)
        if gi == i:
            assert gs == s
            hits += 1
    assert hits >= 30
    d.close()
