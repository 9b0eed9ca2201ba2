"""Seeded synthetic workloads.  numpy only; shared by tests and bench.py.

There is no network and the reference ships no data, so every benchmark and parity input is
generated here from a seed and described in bench.py's ``data`` field.
"""
import numpy as np

RINGS, SECTORS, DESC = 20, 60, 1200


def shift_desc(desc_flat, s):
    """circshift (SC/Scancontext.cpp:39-59) on a flat column-major descriptor:
    out.col((c + s) % 60) = in.col(c)."""
    m = np.asarray(desc_flat).reshape(SECTORS, RINGS)      # row c of this view == column c
    return np.roll(m, s, axis=0).reshape(-1)


def sc_database(D, n_revisits, seed=42, binary=True, flip=0.05):
    """D descriptors, 20x60 f64 column-major, flat (D,1200).

    binary=True : occupancy Bernoulli(p_ring), p decaying 0.45 -> 0.10 with the ring index,
                  occupied bins = 3.0 -- the {0, 3.0} descriptors this pipeline produces
                  (ORORA publishes z = 1.0, SCManager adds LIDAR_HEIGHT = 2.0).
    binary=False: occupied bins get a height U(0.5, 4) (exercises inexact f64 arithmetic).
    n_revisits planted loop closures: entry j is an earlier entry i circularly shifted by a
    random s with `flip` of its bins toggled.  Returns (db, revisits[(j, i, s)]).
    """
    rng = np.random.default_rng(seed)
    p = np.linspace(0.45, 0.10, RINGS)
    occ = rng.random((D, SECTORS, RINGS)) < p[None, None, :]
    if binary:
        db = np.where(occ, 3.0, 0.0)
    else:
        db = np.where(occ, rng.uniform(0.5, 4.0, size=occ.shape), 0.0)
    db = db.reshape(D, DESC)
    revisits = []
    if n_revisits and D > 64:
        js = rng.choice(np.arange(D // 2, D), size=min(n_revisits, D // 2), replace=False)
        for j in sorted(js.tolist()):
            i = int(rng.integers(0, max(1, j - 40)))
            s = int(rng.integers(0, SECTORS))
            d = shift_desc(db[i], s).copy()
            tog = rng.random(DESC) < flip
            if binary:
                d = np.where(tog, np.where(d > 0, 0.0, 3.0), d)
            else:
                d = np.where(tog, np.where(d > 0, 0.0, rng.uniform(0.5, 4.0, DESC)), d)
            db[j] = d
            revisits.append((j, i, s))
    return np.ascontiguousarray(db), revisits


def feature_cloud(n_points, seed, max_range=100.0, z=1.0):
    """A 2-D radar feature cloud as /orora/cloud_local carries it: (x, y, z = 1.0) float32
    (ORORA/src/odometry.cpp:305-316).  Points cluster along a few walls plus clutter."""
    rng = np.random.default_rng(seed)
    n_wall = n_points // 2
    segs = rng.uniform(-max_range * 0.7, max_range * 0.7, size=(12, 4))
    which = rng.integers(0, 12, n_wall)
    t = rng.random(n_wall)
    wx = segs[which, 0] + t * (segs[which, 2] - segs[which, 0]) + rng.normal(0, 0.2, n_wall)
    wy = segs[which, 1] + t * (segs[which, 3] - segs[which, 1]) + rng.normal(0, 0.2, n_wall)
    r = max_range * np.sqrt(rng.random(n_points - n_wall))
    a = rng.uniform(0, 2 * np.pi, n_points - n_wall)
    x = np.concatenate([wx, r * np.cos(a)])
    y = np.concatenate([wy, r * np.sin(a)])
    return np.stack([x, y, np.full(n_points, z)], axis=1).astype(np.float32)


def se2(theta, tx, ty):
    c, s = np.cos(theta), np.sin(theta)
    T = np.eye(4)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1], T[0, 3], T[1, 3] = c, -s, s, c, tx, ty
    return T


def correspondences(N, inlier_ratio, seed, theta=np.deg2rad(1.0), tx=2.5, ty=0.1, noise=0.05,
                    max_range=100.0):
    """Matched point pairs as RadarLoader::matchFeatures hands them to ORORA::calcTransform:
    prev (2xN) and curr (2xN) with prev ~= R * curr + t for inliers (so the solver, called
    with src = curr and tgt = prev, should recover (theta, tx, ty)); outliers are random.
    Returns (prev, curr, inlier_mask, T_true)."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(6.0, max_range, N)
    a = rng.uniform(0, 2 * np.pi, N)
    curr = np.stack([r * np.cos(a), r * np.sin(a)])
    c, s = np.cos(theta), np.sin(theta)
    R = np.array([[c, -s], [s, c]])
    prev = R @ curr + np.array([[tx], [ty]]) + rng.normal(0, noise, (2, N))
    inl = rng.random(N) < inlier_ratio
    n_out = int((~inl).sum())
    ro = rng.uniform(6.0, max_range, n_out)
    ao = rng.uniform(0, 2 * np.pi, n_out)
    prev[:, ~inl] = np.stack([ro * np.cos(ao), ro * np.sin(ao)])
    return np.ascontiguousarray(prev), np.ascontiguousarray(curr), inl, se2(theta, tx, ty)


# -------------------------------------------------------------------------------------------------
# synthetic Navtech polar frames in the "oxford-form" PNG row layout load_radar decodes
# (ORORA/src/radar_utils.cpp:104-129): bytes 0-7 int64 stamp, 8-9 uint16 encoder, 10 valid,
# 11.. one byte of power per range bin.
# -------------------------------------------------------------------------------------------------
RADAR = {"mulran": dict(range_bins=3360, res=0.05952, stamp_step=625000, jitter=True),
         "oxford": dict(range_bins=3768, res=0.0432, stamp_step=625, jitter=False)}


def radar_world(seed=1234, n_points=300, n_walls=20, extent=200.0):
    """Static scene: point scatterers + wall segments (sampled every 0.25 m) with a peak power."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-extent / 2, extent / 2, size=(n_points, 2))
    pw = rng.uniform(120, 250, n_points)
    xs, ps = [pts], [pw]
    for _ in range(n_walls):
        a = rng.uniform(-extent / 2, extent / 2, 2)
        d = rng.uniform(0, 2 * np.pi)
        ln = rng.uniform(8, 40)
        n = int(ln / 0.25)
        seg = a[None, :] + np.linspace(0, ln, n)[:, None] * np.array([np.cos(d), np.sin(d)])[None, :]
        xs.append(seg)
        ps.append(np.full(n, rng.uniform(140, 240)))
    return np.concatenate(xs), np.concatenate(ps)


def radar_frame(world, pose=(0.0, 0.0, 0.0), seed=1234, dataset="mulran", t0=1_560_000_000_000_000_000,
                n_azimuths=400):
    """Render the world as seen from pose (x, y, yaw) into uint8 rows (n_azimuths, 11 + bins).
    A scatterer at sensor-frame polar (range, azimuth) lights a 3-5 bin range blob over 1-3
    azimuths; background is clip(N(30, 8))."""
    cfg = RADAR[dataset]
    bins, res = cfg["range_bins"], cfg["res"]
    rng = np.random.default_rng(seed)
    pts, pw = world
    x0, y0, yaw = pose
    c, s = np.cos(yaw), np.sin(yaw)
    dx, dy = pts[:, 0] - x0, pts[:, 1] - y0
    lx, ly = c * dx + s * dy, -s * dx + c * dy                      # world -> sensor frame
    rngm = np.hypot(lx, ly)
    az = np.arctan2(ly, lx) % (2 * np.pi)
    jit = rng.integers(0, 2, n_azimuths) if cfg["jitter"] else np.zeros(n_azimuths, dtype=np.int64)
    # the first two azimuths keep the nominal 14-tick step: the node derives the azimuth step of its Cartesian image from
    # exactly these two (radar_utils.cpp:279, reached with the default CTS350 for MulRan too: radar_loader.hpp:105-107),
    # so a jittered pair would warp the whole image by 7 % -- real scans start regular
    jit[:2] = 0
    enc = ((np.arange(n_azimuths) * 14 + jit) % 5600).astype(np.uint16)
    az_table = enc.astype(np.float64) * 2 * np.pi / 5600.0
    img = np.clip(rng.normal(30, 8, size=(n_azimuths, bins)), 0, 255)
    rbin = np.round(rngm / res - 0.5).astype(np.int64)
    arow = np.round(az / (2 * np.pi) * n_azimuths).astype(np.int64) % n_azimuths
    half_w = rng.integers(1, 3, len(pts))                           # range blob half width 1..2 -> 3..5 bins
    half_a = rng.integers(0, 2, len(pts))                           # azimuth half width 0..1 -> 1..3 rows
    for k in range(len(pts)):
        if rbin[k] < 60 or rbin[k] >= bins - 3:
            continue
        for da in range(-half_a[k], half_a[k] + 1):
            r_ = (arow[k] + da) % n_azimuths
            for dr in range(-half_w[k], half_w[k] + 1):
                v = pw[k] * np.exp(-0.5 * (dr / 1.2) ** 2) * (1.0 if da == 0 else 0.8)
                b = rbin[k] + dr
                if img[r_, b] < v:
                    img[r_, b] = v
    raw = np.zeros((n_azimuths, 11 + bins), dtype=np.uint8)
    stamps = (t0 + np.arange(n_azimuths, dtype=np.int64) * cfg["stamp_step"]).astype("<i8")
    raw[:, 0:8] = stamps.view(np.uint8).reshape(n_azimuths, 8)
    raw[:, 8:10] = enc.astype("<u2").view(np.uint8).reshape(n_azimuths, 2)
    raw[:, 10] = 255
    raw[:, 11:] = np.round(img).astype(np.uint8)
    return raw
