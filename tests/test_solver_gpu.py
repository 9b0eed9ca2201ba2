"""GPU parity of the ORORA solver (nrs_orora_solve) against oracle/solver_ref.c: adjacency /
clique / inlier sets exact, pose within north_star's 1e-4 m / 1e-5 rad."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL_M, TOL_RAD = 1e-4, 1e-5


def pose_err(Ta, Tb):
    dth = np.arctan2(Ta[1, 0], Ta[0, 0]) - np.arctan2(Tb[1, 0], Tb[0, 0])
    return np.hypot(Ta[0, 3] - Tb[0, 3], Ta[1, 3] - Tb[1, 3]), abs((dth + np.pi) % (2 * np.pi) - np.pi)


def to_oracle_params(O, p):
    return O.OroraParams(p.noise_bound, p.cbar2, p.rotation_max_iterations, p.rotation_gnc_factor,
                         p.rotation_cost_threshold, p.noise_bound_radial, p.noise_bound_tangential)


@pytest.mark.parametrize("N,rho,seed", [(2, 1.0, 9), (3, 1.0, 10), (40, 0.6, 11), (64, 0.9, 12), (65, 0.5, 13),
                                         (200, 0.5, 14), (333, 0.7, 15), (512, 0.4, 16), (700, 0.8, 17),
                                         (700, 0.3, 18), (1024, 0.6, 19)])
def test_solve_parity(nrs, ctx, oracle, synth, N, rho, seed):
    prev, curr, inl, T = synth.correspondences(N, rho, seed)
    p = nrs.orora.default_params()
    p.check_stop_motion = 0
    p.use_doppler_compensation = 0
    g = nrs.orora.solve(ctx, prev, curr, p)
    o = oracle.orora_solve(curr, prev, to_oracle_params(oracle, p))
    gi, oi = g["info"], o["info"]
    assert gi.valid == oi.valid
    assert gi.n_edges == oi.n_edges
    assert gi.M == oi.M and np.array_equal(g["clique"], o["clique"])
    assert gi.clique_exact == 1
    if not gi.valid:
        assert np.array_equal(g["T"], np.eye(4))
        return
    assert gi.gnc_early_return == oi.gnc_early_return
    assert gi.gnc_iters == oi.gnc_iters
    assert gi.n_rot_inliers == oi.n_rot_inliers
    assert np.array_equal(g["final_inliers"], o["final_inliers"])
    dm, dr = pose_err(g["T"], o["T"])
    assert dm <= TOL_M and dr <= TOL_RAD
    assert dm < 1e-9 and dr < 1e-11            # in practice far inside the tolerance
    assert g["T"][2, 3] == 0 and g["T"][2, 2] == 1 and g["T"][3, 3] == 1


def test_doppler_stopmotion_and_state(nrs, ctx, oracle, synth):
    """ORORA.calcTransform over a short sequence: velocity state, Doppler compensation with the
    previous velocity, stop-motion gate -- step by step against the oracle."""
    est = nrs.orora.ORORA("mulran", "rot", True, False, True, 700, 0.049, 0.1, np.deg2rad(9.0), ctx=ctx)
    params = nrs.orora.setParams(0.75, 1.0, False, 100, 1.4, 1e-6)
    vel = 0.0
    for k, (N, rho) in enumerate([(300, 0.7), (350, 0.6), (800, 0.9), (280, 0.5), (120, 0.8)]):
        prev, curr, inl, T = synth.correspondences(N, rho, 100 + k, tx=2.5 + 0.2 * k, theta=np.deg2rad(0.5 * k))
        pm = nrs.orora.RadarFeature(feat_in_cart_=prev)
        cm = nrs.orora.RadarFeature(feat_in_cart_=curr)
        Tg = est.calcTransform(params, 0.25, pm, cm)
        o, vel = oracle.orora_calc_transform(prev, curr, vel, to_oracle_params(oracle, params))
        if N > 700:
            assert np.array_equal(Tg, np.eye(4)) and est.last_result["info"].stop_motion == 1
        dm, dr = pose_err(Tg, o["T"])
        assert dm <= TOL_M and dr <= TOL_RAD
        assert abs(est.curr_vel_se3_[0] - vel) <= TOL_M / 0.25
        if N <= 700:
            assert np.array_equal(est.last_result["clique"], o["clique"])
            src_mc, tgt_mc = est.getMaxCliques()
            assert src_mc.shape == (3, len(o["clique"])) and np.array_equal(src_mc[:2], curr[:, o["clique"]])
            src_in, tgt_in = est.getFinalInliers()
            assert np.array_equal(tgt_in[:2], prev[:, o["final_inliers"]])


def test_degenerate_inputs(nrs, ctx, oracle):
    p = nrs.orora.default_params()
    p.check_stop_motion = 0
    # no two mutually consistent pairs -> Identity, valid = 0 (nonminimal_solver.hpp:715-720)
    prev = np.array([[0.0, 5.0], [0.0, 0.0]])
    curr = np.array([[0.0, 50.0], [0.0, 0.0]])
    g = nrs.orora.solve(ctx, prev, curr, p)
    assert (not g["valid"]) and np.array_equal(g["T"], np.eye(4)) and g["info"].M == 1
    for n in (0, 1):
        g = nrs.orora.solve(ctx, np.zeros((2, n)), np.zeros((2, n)), p)
        assert (not g["valid"]) and np.array_equal(g["T"], np.eye(4))
    with pytest.raises(nrs.NrsError):
        nrs.orora.solve(ctx, np.zeros((2, 1025)), np.zeros((2, 1025)), p)
    # pure identity motion, no noise: GNC terminates at initialisation (mu <= 0) like the reference
    rng = np.random.default_rng(0)
    pts = rng.uniform(-50, 50, (2, 100))
    g = nrs.orora.solve(ctx, pts, pts, p)
    o = oracle.orora_solve(pts, pts)
    assert g["info"].gnc_early_return == o["info"].gnc_early_return
    assert np.allclose(g["T"], o["T"], atol=1e-9)


def test_golden_vectors(nrs, ctx):
    """CUDA path against the committed golden vectors (tests/golden/orora_golden.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "orora_golden.npz"))
    p = nrs.orora.default_params()
    p.check_stop_motion = 0
    p.use_doppler_compensation = 0
    for tag in ("a", "b"):
        r = nrs.orora.solve(ctx, g[f"sol_{tag}_prev"], g[f"sol_{tag}_curr"], p)
        assert np.array_equal(r["clique"], g[f"sol_{tag}_clique"])
        assert np.array_equal(r["final_inliers"], g[f"sol_{tag}_final"])
        dm, dr = pose_err(r["T"], g[f"sol_{tag}_T"])
        assert dm <= TOL_M and dr <= TOL_RAD
        i = r["info"]
        assert [i.n_edges, i.M, i.gnc_iters, i.n_rot_inliers] == list(g[f"sol_{tag}_meta"])
    fft = (g["cen_fft_u8"].astype(np.float64) / 255.0).astype(np.float32)
    assert np.array_equal(nrs.orora.cen2019features(ctx, fft, 400, 58), g["cen_targets"])
