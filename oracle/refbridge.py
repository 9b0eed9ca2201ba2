"""Bridge to oracle/_ref/libref_frontend.so -- the REFERENCE's own features.cpp / radar_utils.cpp compiled in place
(oracle/Makefile `ref`, stand-in headers in oracle/refshim/).  The OpenCV *algorithms* those files call are served by
the real cv2 of this process through the callback table registered here, so a call runs: the reference's loops
(this box's g++ / libstdc++ / libm) + the real OpenCV.  Used to PIN the oracle (oracle/frontend_ref.*):
tests/test_ref_frontend.py.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Available only where the library was built, i.e. where
/root/reference exists (this container) or where the prebuilt file travelled (the GPU box).
"""
import ctypes as C
import os

import cv2
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref_frontend.so")

CTS350, CIR204 = 0, 1

_FILTER2D = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_float))
_MEAN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_float), C.c_int, C.c_int)
_MINMAX = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double))
_REMAP = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int,
                     C.POINTER(C.c_float))
_TO_U8 = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_double, C.POINTER(C.c_uint8))
_IMREAD = C.CFUNCTYPE(C.c_int, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint8), C.c_long)
_ORB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_float,
                   C.POINTER(C.c_uint8))
_KNN = C.CFUNCTYPE(None, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                   C.POINTER(C.c_float))


class _Callbacks(C.Structure):
    _fields_ = [("filter2D", _FILTER2D), ("mean", _MEAN), ("minmax", _MINMAX), ("remap", _REMAP), ("to_u8", _TO_U8),
                ("imread_gray", _IMREAD), ("orb_compute", _ORB), ("knn_match", _KNN)]


def _arr(p, rows, cols, dtype=np.float32):
    return np.ctypeslib.as_array(p, shape=(rows, cols)) if dtype == np.float32 else np.ctypeslib.as_array(p, shape=(rows, cols))


def _cb_filter2d(src, rows, cols, kernel, krows, kcols, dst):
    out = cv2.filter2D(_arr(src, rows, cols), -1, _arr(kernel, krows, kcols).copy(), anchor=(-1, -1), delta=0,
                       borderType=cv2.BORDER_REFLECT101)
    _arr(dst, rows, cols)[:] = out


def _cb_mean(src, rows, cols):
    return float(cv2.mean(_arr(src, rows, cols))[0])


def _cb_minmax(src, rows, cols, mn, mx):
    lo, hi, _, _ = cv2.minMaxLoc(_arr(src, rows, cols))
    mn[0], mx[0] = lo, hi


def _cb_remap(src, rows, cols, mapx, mapy, drows, dcols, dst):
    out = cv2.remap(_arr(src, rows, cols), _arr(mapx, drows, dcols), _arr(mapy, drows, dcols), cv2.INTER_LINEAR,
                    borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    _arr(dst, drows, dcols)[:] = out


def _cb_to_u8(src, rows, cols, alpha, dst):
    # Mat::convertTo(CV_8UC1, alpha) on a non-negative CV_32F image: the Python binding reaches the same
    # cvtScale kernel through convertScaleAbs (|x * alpha| saturated to u8; the remapped image is >= 0)
    a = _arr(src, rows, cols)
    assert not (a < 0).any()
    np.ctypeslib.as_array(dst, shape=(rows, cols))[:] = cv2.convertScaleAbs(a, alpha=alpha)


def _cb_imread(path, rows, cols, buf, cap):
    img = cv2.imread(path.decode(), cv2.IMREAD_GRAYSCALE)
    if img is None:
        return 1
    rows[0], cols[0] = img.shape
    if buf:
        if cap < img.size:
            return 2
        np.ctypeslib.as_array(buf, shape=img.shape)[:] = img
    return 0


def _cb_orb(img, rows, cols, patch_size, edge_threshold, kp_xy, n, kp_size, desc):
    orb = cv2.ORB_create()
    orb.setPatchSize(patch_size)
    orb.setEdgeThreshold(edge_threshold)
    xy = np.ctypeslib.as_array(kp_xy, shape=(max(n, 1), 2))
    kps = [cv2.KeyPoint(float(xy[i, 0]), float(xy[i, 1]), float(kp_size)) for i in range(n)]
    kps2, d = orb.compute(np.ctypeslib.as_array(img, shape=(rows, cols)), kps)
    m = len(kps2)
    for i, k in enumerate(kps2):
        xy[i, 0], xy[i, 1] = k.pt
    if m:
        np.ctypeslib.as_array(desc, shape=(max(n, 1), 32))[:m] = d
    return m


def _cb_knn(q, nq, t, nt, counts, idx, dist):
    cnt = np.ctypeslib.as_array(counts, shape=(max(nq, 1),))
    cnt[:] = 0
    if nq == 0 or nt == 0:
        return
    ii = np.ctypeslib.as_array(idx, shape=(nq, 2))
    dd = np.ctypeslib.as_array(dist, shape=(nq, 2))
    knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(np.ctypeslib.as_array(q, shape=(nq, 32)), np.ctypeslib.as_array(t, shape=(nt, 32)), k=2)
    for j, m in enumerate(knn):
        cnt[j] = len(m)
        for c, mm in enumerate(m):
            assert mm.queryIdx == j
            ii[j, c], dd[j, c] = mm.trainIdx, mm.distance


_lib = None
_keep = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib, _keep
    if _lib is None:
        L = C.CDLL(_PATH)
        _keep = _Callbacks(_FILTER2D(_cb_filter2d), _MEAN(_cb_mean), _MINMAX(_cb_minmax), _REMAP(_cb_remap), _TO_U8(_cb_to_u8),
                           _IMREAD(_cb_imread), _ORB(_cb_orb), _KNN(_cb_knn))
        L.ref_set_callbacks(C.byref(_keep))
        L.ref_get_azimuth_index.restype = C.c_double
        L.ref_get_azimuth_index.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.ref_loader_create.restype = C.c_void_p
        L.ref_loader_create.argtypes = [C.c_char_p, C.c_char_p, C.c_double, C.c_int]
        L.ref_extract_feature.restype = C.c_void_p
        L.ref_extract_feature.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        for name in ("ref_loader_destroy", "ref_feature_destroy"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.ref_feature_sizes.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_feature_get.argtypes = [C.c_void_p] * 10
        L.ref_match_features.argtypes = [C.c_void_p] * 3 + [C.POINTER(C.c_void_p)] * 2
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _targets(fn, fft, *args):
    fft = np.ascontiguousarray(fft, dtype=np.float32)
    cap = 1 << 16
    while True:
        out = np.zeros(3 * cap)
        k = fn(_p(fft), fft.shape[0], fft.shape[1], *args, _p(out), cap)
        if k >= 0:
            return out[:3 * k].reshape(k, 3).T.copy()
        cap = -k


def cen2019features(fft, max_points=10000, min_range=58):
    """features.cpp:197-299, the reference's compiled code -> 3 x K targets"""
    return _targets(lib().ref_cen2019features, fft, int(max_points), int(min_range))


def cen2018features(fft, zq=3.0, sigma_gauss=17, min_range=58):
    """features.cpp:44-131"""
    return _targets(lib().ref_cen2018features, fft, C.c_float(zq), int(sigma_gauss), int(min_range))


def load_radar(path, navtech_version=CTS350, max_rows=4096):
    """radar_utils.cpp:104-129 -> (timestamps int64, azimuths f64, valid bool, fft f32 N x range_bins)"""
    bins = 3360 if navtech_version == CIR204 else 3768
    t = np.zeros(max_rows, dtype=np.int64)
    az = np.zeros(max_rows)
    v = np.zeros(max_rows, dtype=np.uint8)
    fft = np.zeros((max_rows, bins), dtype=np.float32)
    r, c = C.c_int(), C.c_int()
    rc = lib().ref_load_radar(str(path).encode(), int(navtech_version), max_rows, _p(t), _p(az), _p(v), _p(fft), C.byref(r), C.byref(c))
    assert rc == 0 and c.value == bins, (rc, r.value, c.value)
    n = r.value
    return t[:n].copy(), az[:n].copy(), v[:n].astype(bool), fft[:n].copy()


def get_azimuth_index(azimuths, azimuth):
    """radar_utils.cpp:227-250"""
    az = np.ascontiguousarray(azimuths, dtype=np.float64)
    return lib().ref_get_azimuth_index(_p(az), len(az), float(azimuth))


def polar_to_cartesian_radar_img(azimuths, fft, radar_resolution, cart_resolution, width, interpolate_crossover=True,
                                 navtech_version=CTS350, as_u8=True):
    """radar_utils.cpp:252-316 -> (image, rows the caller's fft_data has afterwards)"""
    az = np.ascontiguousarray(azimuths, dtype=np.float64)
    fft = np.ascontiguousarray(fft, dtype=np.float32)
    out = np.zeros((width, width), dtype=np.uint8 if as_u8 else np.float32)
    rows_after = C.c_int()
    lib().ref_polar_to_cartesian_radar_img(_p(az), len(az), _p(fft), fft.shape[0], fft.shape[1], C.c_float(radar_resolution),
                                           C.c_float(cart_resolution), int(width), int(bool(interpolate_crossover)),
                                           int(navtech_version), int(bool(as_u8)), _p(out), C.byref(rows_after))
    return out, rows_after.value


def polar_to_cartesian_points(azimuths, times, polar, radar_resolution):
    """radar_utils.cpp:329-341 -> (cart 3 x N, point_times)"""
    az = np.ascontiguousarray(azimuths, dtype=np.float64)
    t = np.ascontiguousarray(times, dtype=np.int64)
    pol = np.asfortranarray(np.asarray(polar, dtype=np.float64))
    n = pol.shape[1]
    cart = np.zeros((3, n), order="F")
    pt = np.zeros(max(n, 1), dtype=np.int64)
    lib().ref_polar_to_cartesian_points(_p(az), len(az), _p(t), _p(pol), n, C.c_float(radar_resolution), _p(cart), _p(pt))
    return np.ascontiguousarray(cart), pt[:n].copy()


def convert_to_bev(cart, cart_resolution, width, patch_size, point_times):
    """radar_utils.cpp:382-403 -> (cart 3 x K, keypoints K x 2 (u, v) float32, point_times K)"""
    c = np.asfortranarray(np.asarray(cart, dtype=np.float64)).copy(order="F")
    n = c.shape[1]
    kp = np.zeros((max(n, 1), 2), dtype=np.float32)
    pt = np.ascontiguousarray(point_times, dtype=np.int64).copy()
    if n == 0:
        pt = np.zeros(1, dtype=np.int64)
    k = lib().ref_convert_to_bev(_p(c), n, C.c_float(cart_resolution), int(width), int(patch_size), _p(kp), _p(pt))
    return np.ascontiguousarray(c[:, :k]), kp[:k].copy(), pt[:k].copy()


class RadarLoader:
    """The reference's RadarLoader (radar_loader.hpp:35-308), compiled: extractFeature(datadir, file) and matchFeatures
    return dicts of plain arrays.  pcl::VoxelGrid inside is the stand-in (oracle/refshim/pcl/filters/voxel_grid.h: the
    oracle's own recalled restatement, NOT a pin of A5); everything else is the reference's code + the real cv2."""

    def __init__(self, dataset, keypoint_extraction, voxel_size=0.6, use_voxelization=True):
        self.h = lib().ref_loader_create(dataset.encode(), keypoint_extraction.encode(), float(voxel_size), int(bool(use_voxelization)))
        if not self.h:
            raise ValueError("RadarLoader: the reference threw std::invalid_argument")

    def close(self):
        if self.h:
            lib().ref_loader_destroy(self.h)
            self.h = None

    @staticmethod
    def _unpack(f):
        L = lib()
        sz = np.zeros(9, dtype=np.int32)
        L.ref_feature_sizes(f, _p(sz))
        kc, rc, kp0, nkp, nd, nt, ntn, ir, ic = (int(v) for v in sz)
        cart = np.zeros((rc, kc), order="F")
        polar = np.zeros((3, kp0), order="F")
        kp = np.zeros((max(nkp, 1), 2), dtype=np.float32)
        desc = np.zeros((max(nd, 1), 32), dtype=np.uint8)
        t = np.zeros(max(nt, 1), dtype=np.int64)
        tn = np.zeros(max(ntn, 1))
        img = np.zeros((max(ir, 1), max(ic, 1)), dtype=np.uint8)
        tc, den = C.c_int64(), C.c_double()
        L.ref_feature_get(f, _p(cart), _p(polar), _p(kp), _p(desc), _p(t), _p(tn), _p(img),
                          C.cast(C.byref(tc), C.c_void_p), C.cast(C.byref(den), C.c_void_p))
        return {"feat_in_cart": np.ascontiguousarray(cart), "polar": np.ascontiguousarray(polar), "kp": kp[:nkp].copy(),
                "desc": desc[:nd].copy(), "t": t[:nt].copy(), "t_norm": tn[:ntn].copy(), "img": img if ir else img[:0],
                "t_criteria": tc.value, "time_unit_denominator": den.value, "_handle": f}

    def extractFeature(self, datadir, radar_file):
        f = lib().ref_extract_feature(self.h, str(datadir).encode(), str(radar_file).encode())
        return self._unpack(f)

    def matchFeatures(self, prev, curr):
        pm, cm = C.c_void_p(), C.c_void_p()
        lib().ref_match_features(self.h, prev["_handle"], curr["_handle"], C.byref(pm), C.byref(cm))
        return self._unpack(pm.value), self._unpack(cm.value)
