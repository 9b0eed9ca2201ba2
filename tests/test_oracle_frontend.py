"""CPU tests of the front-end oracle: the OpenCV arithmetic the CUDA kernels restate is pinned
against the real cv2 calls, the exact parallel form of the cen2019 marking loop is checked against
the sequential loop, and the whole chain recovers a planted motion."""
import importlib
import math

import cv2
import numpy as np
import pytest


@pytest.fixture(scope="module")
def F(oracle):
    return importlib.import_module("oracle.frontend_ref")


def test_remap_restatement_matches_cv2():
    rng = np.random.default_rng(0)
    H, W = 50, 300
    src = rng.random((H, W)).astype(np.float32)
    mx = (rng.random((64, 64)) * (W + 20) - 10).astype(np.float32)
    my = (rng.random((64, 64)) * (H + 6) - 3).astype(np.float32)
    ref = cv2.remap(src, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    sx = np.rint(mx * np.float32(32)).astype(np.int64)
    sy = np.rint(my * np.float32(32)).astype(np.int64)
    ix, iy = sx >> 5, sy >> 5
    fx = (sx & 31).astype(np.float32) * np.float32(1 / 32)
    fy = (sy & 31).astype(np.float32) * np.float32(1 / 32)

    def tap(y, x):
        ok = (x >= 0) & (x < W) & (y >= 0) & (y < H)
        return np.where(ok, src[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)], np.float32(0))
    one = np.float32(1)
    out = ((tap(iy, ix) * ((one - fy) * (one - fx)) + tap(iy, ix + 1) * ((one - fy) * fx))
           + tap(iy + 1, ix) * (fy * (one - fx))) + tap(iy + 1, ix + 1) * (fy * fx)
    assert np.array_equal(out, ref)


def test_u8_conversion_and_prewitt_and_mean():
    rng = np.random.default_rng(1)
    v = (rng.random((200, 200)) * 3).astype(np.float32)
    mxv = float(v.max())
    assert np.array_equal(cv2.convertScaleAbs(v, alpha=255.0 / mxv),
                          np.clip(np.rint(v * np.float32(255.0 / mxv)), 0, 255).astype(np.uint8))
    f = (rng.integers(0, 256, (40, 100)) / 255.0).astype(np.float32)
    pw = np.zeros((1, 3), np.float32)
    pw[0, 0], pw[0, 2] = -1, 1
    g = cv2.filter2D(f, -1, pw, borderType=cv2.BORDER_REFLECT101)
    fp = np.pad(f, ((0, 0), (1, 1)), mode="reflect")
    assert np.array_equal(g, fp[:, 2:] - fp[:, :-2])
    s = float(f.astype(np.float64).sum())            # exact for k/255 data in any order
    assert cv2.mean(f)[0] == s * (1.0 / f.size)


GAUSS_BITS = np.array([1032826801, 1040595070, 1044597305, 1046301408, 1044597305, 1040595070, 1032826801],
                      dtype=np.uint32)


def blur_restated(img):
    """ORB's blur: sepFilter2D float path.  row: k0*S0 then FMA chain; column: k3*c then FMA of
    symmetric pairs; rounded to u8.  (FMA emulated in float64: product exact, one rounding.)"""
    k = GAUSS_BITS.view(np.float32)
    H, W = img.shape

    def fma(a, b, c):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)
    pad = np.pad(img, ((0, 0), (3, 3)), mode="reflect").astype(np.float32)
    s = (k[0] * pad[:, 0:W]).astype(np.float32)
    for t in range(1, 7):
        s = fma(np.full_like(s, k[t]), pad[:, t:t + W], s)
    pv = np.pad(s, ((3, 3), (0, 0)), mode="reflect")
    o = (k[3] * pv[3:3 + H]).astype(np.float32)
    for t in range(1, 4):
        o = fma(np.full_like(o, k[3 + t]), (pv[3 + t:3 + t + H] + pv[3 - t:3 - t + H]).astype(np.float32), o)
    return np.clip(np.rint(o), 0, 255).astype(np.uint8)


def brief_pattern(patch=21):
    state = 0x34985739

    def nxt():
        nonlocal state
        state = ((state & 0xffffffff) * 4164903690 + (state >> 32)) & 0xffffffffffffffff
        return state & 0xffffffff
    half = patch // 2
    angle = np.float32(-1.0) * np.float32(math.pi / 180.0)
    a, b = np.float32(math.cos(float(angle))), np.float32(math.sin(float(angle)))
    out = []
    for _ in range(512):
        px = np.float32(-half + nxt() % (2 * half + 1))
        py = np.float32(-half + nxt() % (2 * half + 1))
        out.append((int(np.rint(px * a - py * b)), int(np.rint(px * b + py * a))))
    return out


def test_orb_restatement_matches_cv2():
    """The kernel constants (Gaussian taps), the blur arithmetic, the RNG pattern at angle -1 degree and
    the bit order, all against cv2.ORB.compute on 600 keypoints incl. x.5 coordinates."""
    assert np.array_equal(cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel().view(np.uint32), GAUSS_BITS)
    rng = np.random.default_rng(2)
    for smooth in (False, True):
        img = rng.integers(0, 256, (320, 320), dtype=np.uint8)
        if smooth:
            img = cv2.GaussianBlur(img, (9, 9), 3)
        blur = blur_restated(img)
        ref8 = cv2.sepFilter2D(img, cv2.CV_8U, cv2.getGaussianKernel(7, 2, cv2.CV_32F),
                               cv2.getGaussianKernel(7, 2, cv2.CV_32F), borderType=cv2.BORDER_REFLECT_101)
        assert (blur[:, :-4] != ref8[:, :-4]).sum() == 0      # last W%8 columns take cv2's scalar tail
        pts = rng.uniform(22, 297, (300, 2))
        pts[:60] = np.floor(pts[:60]) + 0.5
        orb = cv2.ORB_create()
        orb.setPatchSize(21)
        orb.setEdgeThreshold(21)
        kps, desc = orb.compute(img, [cv2.KeyPoint(float(x), float(y), 21) for x, y in pts])
        assert len(kps) == 300
        pat = brief_pattern()
        for i, kp in enumerate(kps):
            cx, cy = int(np.rint(np.float32(kp.pt[0]))), int(np.rint(np.float32(kp.pt[1])))
            d = np.zeros(32, dtype=np.uint8)
            for n in range(32):
                byte = 0
                for kb in range(8):
                    (x0, y0), (x1, y1) = pat[2 * (8 * n + kb)], pat[2 * (8 * n + kb) + 1]
                    byte |= int(blur[cy + y0, cx + x0] < blur[cy + y1, cx + x1]) << kb
                d[n] = byte
            assert np.array_equal(d, desc[i]), i


def parallel_mark(s, h, mean_h, rows, cols, max_points):
    """The claim formulation the CUDA kernels use (csrc/frontend_cen2019.cu), in numpy."""
    ca, cr = np.nonzero(h > mean_h)
    order = np.argsort(-h[ca, cr], kind="stable")
    ca, cr = ca[order], cr[order]
    claim = np.full((rows, cols), np.iinfo(np.int64).max)
    regions = []
    for k, (a, r) in enumerate(zip(ca, cr)):
        lo = r
        while lo - 1 >= 0 and s[a, lo - 1] < 0:
            lo -= 1
        hi = r
        if r < rows - 1:
            while hi + 1 < cols and s[a, hi + 1] < 0:
                hi += 1
        regions.append((a, lo, hi))
        claim[a, lo:hi + 1] = np.minimum(claim[a, lo:hi + 1], k)
    incr = np.array([int(np.all(claim[a, lo:hi + 1] == k)) for k, (a, lo, hi) in enumerate(regions)], dtype=np.int64)
    cums = np.cumsum(incr)
    hit = np.nonzero((cums == max_points) & (incr == 1))[0]
    J = int(hit[0]) if len(hit) else len(regions) - 1
    return (claim <= J).astype(np.uint8), J + 1, int(min(cums[-1], max_points)) if len(cums) else 0


def test_parallel_marking_equals_sequential_loop(F):
    """Exactness of the claim/atomicMin formulation against features.cpp:226-249, incl. heavy ties,
    the `r < s.rows - 1` column quirk regime (cols > rows) and cut-offs in the middle of ties."""
    rng = np.random.default_rng(3)
    for trial in range(40):
        rows, cols = int(rng.integers(4, 24)), int(rng.integers(30, 90))
        levels = int(rng.choice([3, 8, 256]))
        fft = (rng.integers(0, levels, (rows, cols)) * (255 // max(levels - 1, 1)) / 255.0).astype(np.float32)
        mp = int(rng.choice([3, 10, 40, 10000]))
        st = F.cen2019_stages(fft, mp, min_range=int(rng.integers(0, 8)))
        Rm, consumed, l = parallel_mark(st["s"], st["h"], st["mean_h"], rows, cols, mp)
        assert np.array_equal(Rm, st["R"]), trial
        assert l == st["l_regions"]
        if st["l_regions"] == mp:
            assert consumed == st["j_consumed"]


def test_voxelize_and_bev(F):
    rng = np.random.default_rng(4)
    pts = np.vstack([rng.uniform(-150, 150, (2, 3000)), np.ones((1, 3000))])
    v = F.voxelize(0.6, pts)
    assert v.shape[0] == 3 and v.shape[1] < 3000 and np.all(v[2] == 0)
    # every centroid lies in its own voxel, voxels are distinct and ordered y-major
    inv = np.float32(1) / np.float32(0.6)
    ij = np.floor(v[:2].astype(np.float32) * inv).astype(np.int64)
    assert len({(a, b) for a, b in ij.T}) == v.shape[1]
    key = ij[1] * 100000 + ij[0]
    assert np.all(np.diff(key) > 0)
    cart, kp, t = F.convert_to_bev(v, 0.2592, 964, 21, np.arange(3000, dtype=np.int64))
    assert np.all(kp > 21) and np.all(kp < 943) and cart.shape[1] == len(kp) == len(t)


def test_chain_recovers_planted_motion(F, oracle, synth):
    w = synth.radar_world(seed=1234)
    f0 = F.extract_feature(synth.radar_frame(w, (0, 0, 0), seed=1234), "mulran")
    f1 = F.extract_feature(synth.radar_frame(w, (2.5, 0.1, np.deg2rad(1.0)), seed=1235), "mulran")
    assert f0["stages"]["l_regions"] == 10000 and f0["feat_in_cart"].shape[1] > 500
    pm, cm, q, t = F.match_frames(f0, f1)
    assert pm.shape[1] > 100
    r, _ = oracle.orora_calc_transform(pm, cm, 0.0)
    assert r["valid"] and r["info"].M > 50
    assert abs(np.arctan2(r["T"][1, 0], r["T"][0, 0]) - np.deg2rad(1.0)) < 5e-3
    assert abs(r["T"][0, 3] - 2.5) < 0.3 and abs(r["T"][1, 3] - 0.1) < 0.3


def test_cen2018_filter_rule_and_run_centres(F):
    """cen2018features (features.cpp:44-131): (1) the CUDA path's restatement of cv::filter2D with the
    1 x 51 Gaussian -- forward FMA chain from 0, BORDER_REFLECT101 -- equals the real cv2 bit for bit on
    whole rows including both borders; (2) the run-centre rule (middle element, closing column or end of row)
    against a direct Python transcription of :102-112."""
    rng = np.random.default_rng(11)
    n_total = 0
    for trial, (rows, cols) in enumerate([(6, 3360), (5, 700), (3, 130), (4, 64)]):
        fft = (rng.integers(0, 24, (rows, cols)) / 255.0).astype(np.float32)           # receiver noise ...
        hot = rng.random((rows, cols)) < 0.02
        fft[hot] = (rng.integers(120, 256, int(hot.sum())) / 255.0).astype(np.float32)   # ... and sparse strong returns
        fft[:, : min(42, cols // 3)] = 0
        if trial == 1:
            fft[2, 200:260] = 1.0                       # a long plateau: even / odd run lengths
            fft[3, cols - 7:] = 1.0                     # a run still open at the last column is emitted (:111-112)
        st = F.cen2018_stages(fft, 3.0, 17, 58 if cols > 200 else 5)
        q, taps = st["q"], st["taps"]
        pad = np.pad(q, ((0, 0), (25, 25)), mode="reflect")
        acc = np.zeros_like(q)
        tail = np.zeros_like(q)
        for k in range(51):
            x = pad[:, k:k + cols]
            acc = (acc.astype(np.float64) + x.astype(np.float64) * np.float64(taps[k])).astype(np.float32)   # fma: one rounding
            tail = (tail + (x * taps[k]).astype(np.float32)).astype(np.float32)                              # multiply, then add
        body = cols & ~7                                # OpenCV's 8-wide vector body; the rest is its scalar tail
        acc[:, body:] = tail[:, body:]
        assert np.array_equal(acc.view(np.uint32), st["p"].view(np.uint32))
        mr = 58 if cols > 200 else 5
        exp = []
        for i in range(rows):
            thres = np.float32(3.0) * st["sigma_q"][i]
            run = []
            for j in range(mr, cols):
                if st["y"][i, j] > thres:
                    run.append(j)
                elif run:
                    exp.append((i, run[len(run) // 2])); run = []
            if run:
                exp.append((i, run[len(run) // 2]))
        got = [(int(a), int(r)) for a, r in zip(st["targets"][0], st["targets"][1])]
        assert got == exp and np.all(st["targets"][2] == 1)
        n_total += len(exp)
    assert n_total > 20


def test_cen2018_oracle_reproduces_committed_golden(F):
    """the oracle on this machine (its glibc, its cv2) against the committed fixture"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cen2018_golden.npz"))
    for tag in ("a", "b"):
        zq, sg, mr = g[f"{tag}_par"]
        st = F.cen2018_stages(g[f"{tag}_fft"], float(zq), int(sg), int(mr))
        assert np.array_equal(st["targets"], g[f"{tag}_targets"])
        assert np.array_equal(st["sigma_q"], g[f"{tag}_sigma"])
        assert (st["y"].view(np.uint32) != g[f"{tag}_y"].view(np.uint32)).sum() <= 4


def test_cen2018_c_loops_against_numpy_rederivation(F):
    """The oracle's C loops for cen2018 (bias, noise level, response) re-derived independently in numpy:
    np.cumsum on float32 is a sequential float32 accumulation, i.e. the reference's summation order, so bias and
    noise level must agree bit for bit; the response goes through numpy's own exp (not glibc's), so it is
    compared to one float ulp."""
    rng = np.random.default_rng(21)
    fft = (rng.integers(0, 24, (12, 900)) / 255.0).astype(np.float32)
    hot = rng.random(fft.shape) < 0.03
    fft[hot] = (rng.integers(120, 256, int(hot.sum())) / 255.0).astype(np.float32)
    fft[:, :42] = 0
    fft[7] = 0.5                                              # constant row: no negative q, sigma = 0.034
    st = F.cen2018_stages(fft, 3.0, 17, 58)
    rows, cols = fft.shape
    mean = np.cumsum(fft, axis=1, dtype=np.float32)[:, -1] / np.float32(cols)
    q = fft - mean[:, None]
    assert np.array_equal(q.view(np.uint32), st["q"].view(np.uint32))
    terms = np.where(q < 0, np.float32(2) * (q * q), np.float32(0)).astype(np.float32)
    acc = np.cumsum(terms, axis=1, dtype=np.float32)[:, -1]
    nz = (q < 0).sum(axis=1)
    with np.errstate(all="ignore"):
        sigma = np.where(nz > 0, np.sqrt(acc / np.maximum(nz, 1).astype(np.float32)), np.float32(0.034)).astype(np.float32)
    assert np.array_equal(sigma.view(np.uint32), st["sigma_q"].view(np.uint32))
    p = st["p"]
    a = ((q - p) / sigma[:, None]).astype(np.float32).astype(np.float64)
    b = (p / sigma[:, None]).astype(np.float32).astype(np.float64)
    nqp = np.exp(-0.5 * a * a).astype(np.float32)
    npp = np.exp(-0.5 * b * b).astype(np.float32)
    y = (q * (np.float32(1) - nqp)).astype(np.float32) + (p * (nqp - npp).astype(np.float32)).astype(np.float32)
    d = np.abs(y[:, 58:].astype(np.float64) - st["y"][:, 58:])
    assert d.max() <= 2e-7 and (y[:, 58:].view(np.uint32) != st["y"][:, 58:].view(np.uint32)).mean() < 0.01


def test_exp_to_float_is_libm_independent():
    """cen2018's exp() narrowed to float (features.cpp:98-99): the double-double evaluation is the correctly
    rounded double exp (against decimal's 60-digit exp), and exp_to_float equals float(correctly rounded exp)
    on random, tiny, float-subnormal and squared-float arguments."""
    import ctypes as C
    import decimal
    from oracle import pyoracle as O
    O.build()
    L = O.lib()
    L.oracle_exp_dd.restype, L.oracle_exp_dd.argtypes = C.c_double, [C.c_double]
    L.oracle_exp_to_float.restype, L.oracle_exp_to_float.argtypes = C.c_float, [C.c_double]
    decimal.getcontext().prec = 60
    rng = np.random.default_rng(1)
    xs = np.concatenate([-rng.uniform(0, 110, 600), -10.0 ** rng.uniform(-12, 2, 300),
                         -0.5 * (rng.standard_normal(600).astype(np.float32).astype(np.float64) ** 2),
                         -rng.uniform(87.0, 104.0, 200), [0.0, -0.0, -745.2, -800.0, -1e-300]])
    for x in xs:
        cr = float(decimal.Decimal(float(x)).exp())
        assert L.oracle_exp_dd(float(x)) == cr, x
        assert np.float32(L.oracle_exp_to_float(float(x))) == np.float32(cr), x
