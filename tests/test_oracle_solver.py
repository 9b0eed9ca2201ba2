"""CPU tests of the solver oracle (oracle/solver_ref.c) against independent numpy derivations."""
import itertools

import numpy as np


def test_svd_rot_matches_numpy_kabsch(oracle):
    rng = np.random.default_rng(0)
    for trial in range(50):
        M = int(rng.integers(3, 60))
        X = rng.normal(size=(2, M)) * 20
        th = rng.uniform(-np.pi, np.pi)
        Rt = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        Y = Rt @ X + rng.normal(size=(2, M)) * 0.3
        w = rng.random(M)
        R = oracle.orora_svd_rot2d(X, Y, w)
        H = (X * w) @ Y.T
        U, S, Vt = np.linalg.svd(H)
        V = Vt.T
        if np.linalg.det(U) * np.linalg.det(V) < 0:
            V[:, 1] *= -1
        Rn = V @ U.T
        assert np.allclose(R, Rn, atol=1e-12)
        assert abs(np.linalg.det(R) - 1) < 1e-12
        dth = np.arctan2(R[1, 0], R[0, 0]) - th
        assert abs((dth + np.pi) % (2 * np.pi) - np.pi) < 0.1


def brute_clique(adjm):
    n = adjm.shape[0]
    best = ()
    for r in range(n, 0, -1):
        for comb in itertools.combinations(range(n), r):
            if all(adjm[a, b] for a, b in itertools.combinations(comb, 2)):
                return comb
    return best


def bits_from_matrix(adjm):
    n = adjm.shape[0]
    W = (n + 63) // 64
    adj = np.zeros((n, W), dtype=np.uint64)
    for i in range(n):
        for j in range(n):
            if adjm[i, j]:
                adj[i, j >> 6] |= np.uint64(1) << np.uint64(j & 63)
    return adj


def test_max_clique_small_exact(oracle):
    rng = np.random.default_rng(1)
    for trial in range(40):
        n = int(rng.integers(4, 15))
        m = rng.random((n, n)) < rng.uniform(0.2, 0.8)
        m = np.triu(m, 1)
        m = m | m.T
        cl, nodes = oracle.orora_max_clique(bits_from_matrix(m))
        bc = brute_clique(m)
        assert len(cl) == len(bc)
        assert all(m[a, b] for a, b in itertools.combinations(cl, 2))
        assert list(cl) == sorted(cl)


def test_max_clique_planted(oracle):
    rng = np.random.default_rng(2)
    for n, k in ((80, 30), (200, 90), (128, 64), (65, 20)):
        m = rng.random((n, n)) < 0.15
        m = np.triu(m, 1)
        m = m | m.T
        members = np.sort(rng.choice(n, k, replace=False))
        m[np.ix_(members, members)] = True
        np.fill_diagonal(m, False)
        cl, nodes = oracle.orora_max_clique(bits_from_matrix(m))
        assert np.array_equal(cl, members)


def test_cote_axis_vs_bruteforce(oracle):
    """The sweep minimises sum_{in}(x - xhat)^2 + sum_{out} alpha over consensus sets reachable
    at an endpoint; check the estimate against a direct evaluation of the same cost at every
    endpoint."""
    rng = np.random.default_rng(3)
    for trial in range(30):
        N = int(rng.integers(5, 60))
        X = np.concatenate([rng.normal(2.5, 0.05, N // 2 + 1), rng.uniform(-30, 30, N - N // 2 - 1)])
        al = rng.uniform(0.1, 0.6, N)
        est, inl = oracle.orora_cote_axis(X, al)
        ends = np.sort(np.concatenate([X - al, X + al]))
        best = None
        for e in ends:
            S = (X - al <= e) & (e < X + al)
            if S.sum() == 0:
                continue
            w = 1 / al[S] ** 2
            xh = (w * X[S]).sum() / w.sum()
            cost = ((X[S] - xh) ** 2).sum() + al[~S].sum()
            if best is None or cost < best[0] - 1e-12:
                best = (cost, xh)
        assert abs(est - best[1]) < 1e-9
        assert np.array_equal(inl, np.abs(X - est) <= al)
        assert abs(est - 2.5) < 0.2


def test_solver_recovers_pose(oracle, synth):
    for N, rho, seed in ((60, 0.7, 1), (250, 0.5, 2), (500, 0.8, 3), (700, 0.45, 4)):
        prev, curr, inl, T = synth.correspondences(N, rho, seed)
        r = oracle.orora_solve(curr, prev)
        assert r["valid"]
        assert set(r["clique"]) <= set(np.nonzero(inl)[0])
        assert len(r["clique"]) >= 0.9 * inl.sum()
        th = np.arctan2(r["T"][1, 0], r["T"][0, 0])
        assert abs(th - np.deg2rad(1.0)) < 2e-3
        assert abs(r["T"][0, 3] - 2.5) < 0.05 and abs(r["T"][1, 3] - 0.1) < 0.05
        assert r["T"][2, 3] == 0 and r["T"][2, 2] == 1 and r["T"][3, 3] == 1


def test_doppler_and_gates(oracle, synth):
    prev, curr, inl, T = synth.correspondences(100, 0.8, 5)
    d = oracle.orora_doppler(prev, 10.0)
    x, y = prev
    rsq = x * x + y * y
    nx = x + 0.049 * 10 * x * x / rsq
    ny = y + 0.049 * 10 * nx * y / rsq
    assert np.allclose(d, np.stack([nx, ny]), atol=1e-12)
    r, v = oracle.orora_calc_transform(prev, curr, 0.0, stop_thr=50)        # stop-motion gate
    assert np.array_equal(r["T"], np.eye(4)) and v == 0.0
    r, v = oracle.orora_calc_transform(prev, curr, 0.0)
    assert r["valid"] and abs(v - r["T"][0, 3] / 0.25) < 1e-12
    # fewer than two mutually consistent pairs -> Identity, valid = false
    r = oracle.orora_solve(np.array([[0.0, 50.0], [0.0, 0.0]]), np.array([[0.0, 5.0], [0.0, 0.0]]))
    assert (not r["valid"]) and np.array_equal(r["T"], np.eye(4))


def test_orora_golden(oracle):
    import importlib
    import os
    F = importlib.import_module("oracle.frontend_ref")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "orora_golden.npz"))
    for tag in ("a", "b"):
        r = oracle.orora_solve(g[f"sol_{tag}_curr"], g[f"sol_{tag}_prev"])
        assert np.array_equal(r["clique"], g[f"sol_{tag}_clique"])
        assert np.array_equal(r["final_inliers"], g[f"sol_{tag}_final"])
        assert np.allclose(r["T"], g[f"sol_{tag}_T"], rtol=0, atol=1e-12)
        i = r["info"]
        assert [i.n_edges, i.M, i.gnc_iters, i.n_rot_inliers] == list(g[f"sol_{tag}_meta"])
    fft = (g["cen_fft_u8"].astype(np.float64) / 255.0).astype(np.float32)
    assert np.array_equal(F.cen2019features(fft, 400, 58), g["cen_targets"])
