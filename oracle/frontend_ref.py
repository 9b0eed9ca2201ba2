"""CPU restatement of the ORORA front-end (TEST INFRASTRUCTURE ONLY; see oracle/__init__.py).

Follows /root/reference/outlier-robust-radar-odometry/
    src/radar_utils.cpp:104-129   load_radar
    src/features.cpp:197-299      cen2019features
    src/features.cpp:44-131       cen2018features
    src/radar_utils.cpp:252-316   polar_to_cartesian_radar_img
    src/radar_utils.cpp:329-341   polar_to_cartesian_points
    include/radar_loader.hpp:207-222  voxelize  (pcl::VoxelGrid<PointXYZ>, recalled semantics)
    src/radar_utils.cpp:382-403   convert_to_bev
    include/radar_loader.hpp:69-71,126   cv::ORB compute (patchSize 21)
    include/radar_loader.hpp:143-205     matchFeatures (BF Hamming knn(2) + ratio 0.8)
    include/radar_loader.hpp:74-141      extractFeature (the chain of the above)

Every OpenCV stage is executed by the REAL cv2 (4.13, the library family the reference links):
filter2D, mean, minMaxLoc, remap, convertScaleAbs (== convertTo(CV_8U, a) on non-negative
data), ORB.compute, BFMatcher.knnMatch.  Loops that are the reference's own code live in
oracle/frontend_ref.c (libm-exact maps, the sequential marking loop, the azimuth scan).

PARITY PIN STATUS: the reference ships no fixtures for this path; OpenCV's version is not pinned
by the reference and PCL is absent, so this oracle is "parity unpinned" for the PCL VoxelGrid
restatement (recalled from PCL's voxel_grid.hpp) and pinned-by-library for the cv2 stages.
Unstable std::sort calls are restated as stable sorts (ties keep input order).
"""
import ctypes as C
import math
import os

import cv2
import numpy as np

from . import pyoracle

CTS350, CIR204 = 0, 1      # radar_utils.hpp:25-26 navtech versions (MulRan uses CIR204)
ENCODER_SIZE = 5600


class DatasetParams:
    """radar_loader.hpp:48-62"""

    def __init__(self, dataset):
        if dataset == "oxford":
            self.radar_resolution = np.float32(0.0432)
            self.navtech_version = CTS350
            self.range_bins = 3768
        elif dataset == "mulran":
            self.radar_resolution = np.float32(0.05952)
            self.navtech_version = CIR204
            self.range_bins = 3360
        else:
            raise ValueError("Note implemented!")
        self.min_range = 58
        self.cart_resolution = np.float32(0.2592)
        self.cart_pixel_width = 964
        self.max_points = 10000          # radar_loader.hpp:287
        self.patch_size = 21             # :290
        self.nndr = np.float32(0.8)      # :162


def _lib():
    L = pyoracle.lib()
    if not hasattr(L, "_fe_declared"):
        vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
        L.fe_build_maps.restype = None
        L.fe_build_maps.argtypes = [vp, i32, f32, f32, i32, i32, vp, vp, vp]
        L.fe_cen2019_mark.restype = i32
        L.fe_cen2019_mark.argtypes = [vp, i32, i32, vp, vp, i32, i32, vp, vp]
        L.fe_cen2019_scan.restype = i32
        L.fe_cen2019_scan.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32]
        L.fe_cen2018_bias.restype = None
        L.fe_cen2018_bias.argtypes = [vp, i32, i32, vp]
        L.fe_cen2018_taps.restype = f32
        L.fe_cen2018_taps.argtypes = [i32, vp]
        L.fe_cen2018_sigma.restype = None
        L.fe_cen2018_sigma.argtypes = [vp, i32, i32, vp]
        L.fe_cen2018_scan.restype = i32
        L.fe_cen2018_scan.argtypes = [vp, vp, vp, i32, i32, f32, i32, vp, vp, vp, i32]
        L._fe_declared = True
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- radar_utils.cpp:104-129 --------------------------------------------------------------------
def load_radar(raw, navtech_version=CTS350):
    """raw = decoded 8-bit PNG rows, shape (N, 11 + range_bins[+...]).  Returns
    (timestamps int64[N], azimuths f64[N], valid bool[N], fft f32[N, range_bins])."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    N = raw.shape[0]
    range_bins = 3360 if navtech_version == CIR204 else 3768
    timestamps = raw[:, 0:8].copy().view(np.int64).reshape(N)
    enc = raw[:, 8:10].copy().view(np.uint16).reshape(N)
    azimuths = np.array([int(e) * 2 * math.pi / float(ENCODER_SIZE) for e in enc], dtype=np.float64)
    valid = raw[:, 10] == 255
    fft = np.zeros((N, range_bins), dtype=np.float32)
    fft[:, 42:] = (raw[:, 11 + 42: 11 + range_bins].astype(np.float64) / 255.0).astype(np.float32)
    return timestamps, azimuths, valid, fft


# ---- features.cpp:197-299 -------------------------------------------------------------------------
_stdsort = None


def _std_sort_desc(hv, ca, cr):
    """The candidate order this toolchain's std::sort (libstdc++ introsort) gives the reference (oracle/stdsort.cpp)."""
    global _stdsort
    if _stdsort is None:
        _stdsort = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libstdsort.so"))
    hv = np.ascontiguousarray(hv, dtype=np.float32)
    ca = np.ascontiguousarray(ca, dtype=np.int32).copy()
    cr = np.ascontiguousarray(cr, dtype=np.int32).copy()
    _stdsort.oracle_std_sort_desc(_p(hv), _p(ca), _p(cr), len(hv))
    return ca, cr


def cen2019_stages(fft, max_points=10000, min_range=58, order="stable"):
    """Returns a dict with every intermediate the GPU stages are compared against.
    order: how candidates of EQUAL response are ordered by the reference's unstable std::sort (features.cpp:224):
    "stable" = row-major among ties (the oracle's definition, what the GPU path implements); "libstdc++" = the
    permutation this toolchain's introsort produces, i.e. what the reference compiled here does (pinned exactly against
    oracle/_ref/libref_frontend.so in tests/test_ref_frontend.py)."""
    fft = np.ascontiguousarray(fft, dtype=np.float32)
    rows, cols = fft.shape
    prewitt = np.zeros((1, 3), dtype=np.float32)
    prewitt[0, 0], prewitt[0, 2] = -1, 1
    g = cv2.filter2D(fft, -1, prewitt, anchor=(-1, -1), delta=0, borderType=cv2.BORDER_REFLECT101)
    g = np.abs(g)
    ming, maxg, _, _ = cv2.minMaxLoc(g)
    with np.errstate(all="ignore"):                   # a constant frame has maxg = 0: 1./0 = inf and 0 * inf = NaN, like in C++
        g = g * np.float32(np.float64(1.0) / np.float64(maxg))      # Mat /= double == convertTo(-1, 1./s)
    mean = np.float32(cv2.mean(fft)[0])
    s = fft - mean
    h = s * (np.float32(1) - g)
    mean_h = np.float32(cv2.mean(h)[0])
    # candidates in row-major order, then stable descending sort on h
    ca, cr = np.nonzero(h > mean_h)
    hv = h[ca, cr]
    if order == "libstdc++":
        ca, cr = _std_sort_desc(hv, ca, cr)
    else:
        assert order == "stable"
        perm = np.argsort(-hv, kind="stable")
        ca = np.ascontiguousarray(ca[perm], dtype=np.int32)
        cr = np.ascontiguousarray(cr[perm], dtype=np.int32)
    R = np.zeros((rows, cols), dtype=np.uint8)
    j = C.c_int()
    L = _lib()
    s = np.ascontiguousarray(s)
    h = np.ascontiguousarray(h)
    l = L.fe_cen2019_mark(_p(s), rows, cols, _p(ca), _p(cr), len(ca), max_points, _p(R), C.byref(j))
    cap = rows * (cols // 2 + 1)
    oa = np.zeros(cap, dtype=np.int32)
    orr = np.zeros(cap, dtype=np.int32)
    k = L.fe_cen2019_scan(_p(R), _p(h), rows, cols, min_range, _p(oa), _p(orr), cap)
    assert k <= cap
    targets = np.ones((3, k), dtype=np.float64)
    targets[0] = oa[:k]
    targets[1] = orr[:k]
    return {"g": g, "maxg": maxg, "mean": mean, "s": s, "h": h, "mean_h": mean_h, "n_cand": len(ca),
            "cand_a": ca, "cand_r": cr, "R": R, "j_consumed": j.value, "l_regions": l, "targets": targets}


def cen2019features(fft, max_points=10000, min_range=58, order="stable"):
    return cen2019_stages(fft, max_points, min_range, order)["targets"]


def cen2018_stages(fft, zq=3.0, sigma_gauss=17, min_range=58):
    """features.cpp:44-131 (what run_orora.launch selects): the reference's loops in C, its one OpenCV call --
    filter2D with the 1 x 3*sigma Gaussian, BORDER_REFLECT101 -- by the real cv2."""
    fft = np.ascontiguousarray(fft, dtype=np.float32)
    rows, cols = fft.shape
    L = _lib()
    q = np.empty_like(fft)
    L.fe_cen2018_bias(_p(fft), rows, cols, _p(q))
    taps = np.zeros((1, 3 * sigma_gauss), dtype=np.float32)
    s = L.fe_cen2018_taps(sigma_gauss, _p(taps))
    filt = taps * np.float32(1.0 / float(s))                # filter /= s (:73): Mat /= double == convertTo(-1, 1./s)
    p = cv2.filter2D(q, -1, filt, anchor=(-1, -1), delta=0, borderType=cv2.BORDER_REFLECT101)
    sigma_q = np.zeros(rows, dtype=np.float32)
    L.fe_cen2018_sigma(_p(q), rows, cols, _p(sigma_q))
    cap = rows * (cols // 2 + 1)
    oa = np.zeros(cap, dtype=np.int32)
    orr = np.zeros(cap, dtype=np.int32)
    y = np.zeros_like(fft)
    p = np.ascontiguousarray(p)
    with np.errstate(all="ignore"):
        k = L.fe_cen2018_scan(_p(q), _p(p), _p(sigma_q), rows, cols, np.float32(zq), min_range, _p(y), _p(oa), _p(orr), cap)
    assert k <= cap
    targets = np.ones((3, k), dtype=np.float64)
    targets[0] = oa[:k]
    targets[1] = orr[:k]
    return {"q": q, "taps": filt.reshape(-1), "p": p, "sigma_q": sigma_q, "y": y, "targets": targets}


def cen2018features(fft, zq=3.0, sigma_gauss=17, min_range=58):
    return cen2018_stages(fft, zq, sigma_gauss, min_range)["targets"]


def get_azimuth_index(azimuths, azimuth):
    """radar_utils.cpp:227-250"""
    az = np.ascontiguousarray(azimuths, dtype=np.float64)
    fn = _lib().fe_get_azimuth_index
    fn.restype = C.c_double
    fn.argtypes = [C.c_void_p, C.c_int, C.c_double]
    return fn(_p(az), len(az), float(azimuth))


# ---- radar_utils.cpp:252-316 -------------------------------------------------------------------------
def build_maps(azimuths, radar_resolution, cart_resolution, width, navtech_version):
    az = np.ascontiguousarray(azimuths, dtype=np.float64)
    rng = np.empty((width, width), dtype=np.float32)
    ang = np.empty((width, width), dtype=np.float32)
    th = np.empty((width, width), dtype=np.float32)
    _lib().fe_build_maps(_p(az), len(az), float(radar_resolution), float(cart_resolution), width,
                         1 if navtech_version == CIR204 else 0, _p(rng), _p(ang), _p(th))
    return rng, ang, th


def polar_to_cartesian_radar_img(azimuths, fft, radar_resolution, cart_resolution, width,
                                 interpolate_crossover=True, navtech_version=CTS350, as_u8=True):
    rng, ang, _ = build_maps(azimuths, radar_resolution, cart_resolution, width, navtech_version)
    src = np.ascontiguousarray(fft, dtype=np.float32)
    if interpolate_crossover:                                        # :299-309
        src = np.vstack([src[-1:], src, src[:1]])
        ang = ang + np.float32(1)
    cart = cv2.remap(src, rng, ang, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    if not as_u8:
        return cart
    _, mx, _, _ = cv2.minMaxLoc(cart)                                # :311-315
    return cv2.convertScaleAbs(cart, alpha=255.0 / mx)


# ---- radar_utils.cpp:329-341 ---------------------------------------------------------------------------
def polar_to_cartesian_points(azimuths, times, polar_points, radar_resolution):
    res = np.float32(radar_resolution)
    half = float(np.float32(res / np.float32(2)))                    # float / int -> float
    K = polar_points.shape[1]
    cart = polar_points.astype(np.float64).copy()
    t = np.zeros(K, dtype=np.int64)
    for i in range(K):
        a = int(polar_points[0, i])
        azimuth = azimuths[a]
        r = polar_points[1, i] * float(res) + half
        cart[0, i] = r * math.cos(azimuth)
        cart[1, i] = r * math.sin(azimuth)
        t[i] = times[a]
    return cart, t


# ---- radar_loader.hpp:207-222 (pcl::VoxelGrid<PointXYZ>, leaf 0.6) -----------------------------------
def voxelize(voxel_size, cart_targets):
    """PCL voxel_grid.hpp applyFilter, restated: float32 arithmetic, voxel index
    ijk0 + ijk1*dx + ijk2*dx*dy with ijk = floor(p * inv_leaf) - min_b, points sorted by index
    (stable here), centroid = float32 running sum / float(count), output in index order.
    Returns the 3 x K' matrix the reference builds (row 2 = 0)."""
    K = cart_targets.shape[1]
    if K == 0:
        return np.zeros((3, 0))
    pts = cart_targets.astype(np.float32)                            # eigen2pcl: double -> float
    x, y, z = pts[0], pts[1], pts[2]
    leaf = np.float32(voxel_size)
    inv = np.float32(1.0) / leaf
    mn = np.array([x.min(), y.min(), z.min()], dtype=np.float32)
    mx = np.array([x.max(), y.max(), z.max()], dtype=np.float32)
    min_b = np.floor(mn * inv).astype(np.int32)
    max_b = np.floor(mx * inv).astype(np.int32)
    div_b = max_b - min_b + 1
    mul = np.array([1, div_b[0], div_b[0] * div_b[1]], dtype=np.int64)
    ijk0 = (np.floor(x * inv) - np.float32(min_b[0])).astype(np.int32)
    ijk1 = (np.floor(y * inv) - np.float32(min_b[1])).astype(np.int32)
    ijk2 = (np.floor(z * inv) - np.float32(min_b[2])).astype(np.int32)
    idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2]
    order = np.argsort(idx, kind="stable")
    sidx = idx[order]
    starts = np.flatnonzero(np.r_[True, sidx[1:] != sidx[:-1]])
    ends = np.r_[starts[1:], len(sidx)]
    out = np.zeros((3, len(starts)), dtype=np.float64)
    for v, (a, b) in enumerate(zip(starts, ends)):
        sx, sy = np.float32(0), np.float32(0)
        for q in order[a:b]:
            sx = np.float32(sx + x[q])
            sy = np.float32(sy + y[q])
        n = np.float32(b - a)
        out[0, v] = np.float32(sx / n)
        out[1, v] = np.float32(sy / n)
    return out


# ---- radar_utils.cpp:382-403 ----------------------------------------------------------------------------
def convert_to_bev(cart_points, cart_resolution, width, patch_size, point_times):
    res = np.float32(cart_resolution)
    cmr = np.float32((width // 2) * res)
    if width % 2 == 0:
        cmr = np.float32((width // 2 - 0.5) * float(res))            # int - double -> double * float -> float
    cmr, res = float(cmr), float(res)
    kp, keep = [], []
    for i in range(cart_points.shape[1]):
        u = (cmr + cart_points[1, i]) / res
        v = (cmr - cart_points[0, i]) / res
        if 0 < u - patch_size and u + patch_size < width and 0 < v - patch_size and v + patch_size < width:
            kp.append((np.float32(u), np.float32(v)))
            keep.append(i)
    keep = np.array(keep, dtype=np.int64)
    cart = np.zeros((3, len(keep)))
    if len(keep):
        cart[:2] = cart_points[:2, keep]
        cart[2] = cart_points[2, keep]
    # quirk kept: point_times[j] = point_times[i] walks the PRE-voxelisation array (radar_loader.hpp:123)
    t = np.array([point_times[i] for i in keep], dtype=np.int64) if len(keep) else np.zeros(0, np.int64)
    return cart, np.array(kp, dtype=np.float32).reshape(-1, 2), t


# ---- ORB / matcher ---------------------------------------------------------------------------------------
def orb_compute(img_u8, kp_uv, patch_size=21):
    orb = cv2.ORB_create()
    orb.setPatchSize(patch_size)
    orb.setEdgeThreshold(patch_size)
    kps = [cv2.KeyPoint(float(u), float(v), float(patch_size)) for u, v in kp_uv]
    kps2, desc = orb.compute(img_u8, kps)
    if desc is None:
        desc = np.zeros((0, 32), dtype=np.uint8)
    kept = np.array([k.pt for k in kps2], dtype=np.float32).reshape(-1, 2)
    return desc, kept


def match_features(desc_prev, desc_curr, nndr=np.float32(0.8)):
    """radar_loader.hpp:154-167: query = prev, train = curr.  Returns (query_idx, train_idx)."""
    if len(desc_prev) == 0 or len(desc_curr) < 2:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    knn = bf.knnMatch(desc_prev, desc_curr, k=2)
    q, t = [], []
    for m in knn:
        if not len(m):
            continue
        if np.float32(m[0].distance) < np.float32(nndr) * np.float32(m[1].distance):
            q.append(m[0].queryIdx)
            t.append(m[0].trainIdx)
    return np.array(q, dtype=np.int32), np.array(t, dtype=np.int32)


# ---- radar_loader.hpp:74-141 ---------------------------------------------------------------------------------
def extract_feature(raw, dataset="mulran", voxel_size=0.6, use_voxelization=True, keypoint_extraction="cen2019",
                    polar_override=None):
    """polar_override: keypoints to run the rest of the chain on (tests/test_ref_frontend.py hands in cen2019's
    keypoints in the libstdc++ tie order when comparing with the compiled RadarLoader)."""
    P = DatasetParams(dataset)
    times, azimuths, valid, fft = load_radar(raw, P.navtech_version)
    if keypoint_extraction == "cen2018":                  # radar_loader.hpp:95-99, zq_ / sigma_gauss_ :284-285
        st = cen2018_stages(fft, 3.0, 17, P.min_range)
    else:
        st = cen2019_stages(fft, P.max_points, P.min_range)
    polar = st["targets"] if polar_override is None else polar_override
    # radar_loader.hpp:105-107 passes no navtech_version: polar_to_cartesian_radar_img runs with its default CTS350
    # (uniform azimuth step) for MulRan too -- pinned against the compiled RadarLoader in tests/test_ref_frontend.py
    img = polar_to_cartesian_radar_img(azimuths, fft, P.radar_resolution, P.cart_resolution,
                                       P.cart_pixel_width, True, CTS350)
    cart, t = polar_to_cartesian_points(azimuths, times, polar, P.radar_resolution)
    cart_pre_voxel = cart
    if use_voxelization:
        cart = voxelize(voxel_size, cart)
    cart, kp, t = convert_to_bev(cart, P.cart_resolution, P.cart_pixel_width, P.patch_size, t)
    # cv::ORB::compute filters kp_ in place (runByImageBorder on rounded coordinates): kp_/desc_ can
    # end up shorter than feat_in_cart_/t_, and matchFeatures indexes the latter with descriptor rows
    # (radar_loader.hpp:181-187).  Kept as is.
    desc, kp_orb = orb_compute(img, kp, P.patch_size)
    return {"fft": fft, "azimuths": azimuths, "times": times, "polar": polar, "img": img,
            "cart_pre_voxel": cart_pre_voxel, "feat_in_cart": cart, "kp_bev": kp, "kp": kp_orb, "t": t, "desc": desc,
            "stages": st}


def match_frames(prev, curr):
    """matchFeatures on two extract_feature() results -> (prev_2xN, curr_2xN, q, t)."""
    q, t = match_features(prev["desc"], curr["desc"])
    return prev["feat_in_cart"][:2, q], curr["feat_in_cart"][:2, t], q, t
