"""ctypes bindings of oracle/_build/liboracle.so (our CPU restatement of the reference) and,
when present, oracle/_ref/libref_kdtree.so (the reference's own vendored nanoflann).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libref_kdtree.so")

RINGS, SECTORS, DESC = 20, 60, 1200


def _cpu_id():
    """model name + flags of this host's CPU: the library is built -march=native, so a copy that travelled from
    another machine is rebuilt (seconds)."""
    try:
        txt = open("/proc/cpuinfo").read()
        model = next((l.split(":", 1)[1].strip() for l in txt.splitlines() if l.startswith("model name")), "?")
        flags = next((l.split(":", 1)[1].strip() for l in txt.splitlines() if l.startswith("flags")), "?")
        import hashlib
        return model + " " + hashlib.sha1(flags.encode()).hexdigest()[:12]
    except Exception:
        return "unknown"


def build(force=False):
    """Compile the C restatement (gcc, seconds).  Also builds oracle/_ref (the reference's own nanoflann, and its
    features.cpp / radar_utils.cpp against oracle/refshim/) when /root/reference exists (this container); on the GPU
    box the prebuilt _ref files are used as they travelled."""
    stamp = os.path.join(_HERE, "_build", ".cpu")
    same_cpu = os.path.exists(stamp) and open(stamp).read() == _cpu_id()
    if force or not same_cpu or not os.path.exists(_LIB) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
            for f in os.listdir(_HERE) if f.endswith(".c") or f == "stdsort.cpp" or f == "Makefile") or not os.path.exists(
                os.path.join(_HERE, "_build", "libstdsort.so")):
        if not same_cpu:
            subprocess.call(["make", "-s", "-C", _HERE, "clean"])
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
        with open(stamp, "w") as f:
            f.write(_cpu_id())
    if os.path.isdir("/root/reference/pgo/SC-A-LOAM/include/scancontext"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _declare(_lib)
    return _lib


def ref_kdtree():
    """The reference's real nanoflann kd-tree, or None when oracle/_ref was not built."""
    global _ref
    if _ref is None and os.path.exists(_REF):
        _ref = C.CDLL(_REF)
        _ref.ref_kdtree_knn.restype = C.c_int
        _ref.ref_kdtree_knn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_void_p]
    return _ref


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _declare(L):
    vp, i32, f64, f32 = C.c_void_p, C.c_int, C.c_double, C.c_float
    L.sc_xy2theta.restype = f32
    L.sc_xy2theta.argtypes = [f32, f32]
    L.sc_make.restype = None
    L.sc_make.argtypes = [vp, i32, f64, f64, vp]
    L.sc_ringkey.restype = None
    L.sc_ringkey.argtypes = [vp, vp]
    L.sc_sectorkey.restype = None
    L.sc_sectorkey.argtypes = [vp, vp]
    L.sc_ringkey_f32.restype = None
    L.sc_ringkey_f32.argtypes = [vp, vp]
    L.sc_fast_align.restype = i32
    L.sc_fast_align.argtypes = [vp, vp]
    L.sc_dist_direct.restype = f64
    L.sc_dist_direct.argtypes = [vp, vp]
    L.sc_distance.restype = None
    L.sc_distance.argtypes = [vp, vp, vp, vp]
    L.sc_search_range.restype = None
    L.sc_search_range.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp]
    L.sc_detect_from_candidates.restype = None
    L.sc_detect_from_candidates.argtypes = [vp, vp, vp, i32, f64, vp, vp, vp, vp]
    L.sc_knn_bruteforce.restype = None
    L.sc_knn_bruteforce.argtypes = [vp, i32, i32, vp, i32, vp, vp]


# ---------------------------------------------------------------------------------------------
# ScanContext
# ---------------------------------------------------------------------------------------------
def sc_make(xyz, lidar_height=2.0, max_radius=80.0):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    desc = np.empty(DESC, dtype=np.float64)
    lib().sc_make(_p(xyz), xyz.shape[0], lidar_height, max_radius, _p(desc))
    return desc


def sc_ringkey(desc):
    desc = np.ascontiguousarray(desc, dtype=np.float64)
    out = np.empty(RINGS, dtype=np.float64)
    lib().sc_ringkey(_p(desc), _p(out))
    return out


def sc_ringkey_f32(desc):
    desc = np.ascontiguousarray(desc, dtype=np.float64)
    out = np.empty(RINGS, dtype=np.float32)
    lib().sc_ringkey_f32(_p(desc), _p(out))
    return out


def sc_sectorkey(desc):
    desc = np.ascontiguousarray(desc, dtype=np.float64)
    out = np.empty(SECTORS, dtype=np.float64)
    lib().sc_sectorkey(_p(desc), _p(out))
    return out


def sc_fast_align(vk1, vk2):
    vk1 = np.ascontiguousarray(vk1, dtype=np.float64)
    vk2 = np.ascontiguousarray(vk2, dtype=np.float64)
    return int(lib().sc_fast_align(_p(vk1), _p(vk2)))


def sc_dist_direct(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return float(lib().sc_dist_direct(_p(a), _p(b)))


def sc_distance(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    d = C.c_double()
    s = C.c_int()
    lib().sc_distance(_p(a), _p(b), C.byref(d), C.byref(s))
    return d.value, s.value


def sc_search(query, db, first=0, last=None, want_all=True, threads=None):
    """Exhaustive distanceBtnScanContext(query, db[i]) for i in [first,last).
    Returns (dist_all, shift_all, (best_dist, best_idx, best_shift))."""
    query = np.ascontiguousarray(query, dtype=np.float64)
    db = np.ascontiguousarray(db, dtype=np.float64).reshape(-1, DESC)
    last = db.shape[0] if last is None else last
    dist = np.full(db.shape[0], np.nan) if want_all else None
    shift = np.zeros(db.shape[0], dtype=np.int32) if want_all else None
    bd, bi, bs = C.c_double(), C.c_int(), C.c_int()
    if threads is not None:
        lib().oracle_set_threads(int(threads))
    lib().sc_search_range(_p(query), _p(db), first, last,
                          _p(dist) if want_all else None, _p(shift) if want_all else None,
                          C.byref(bd), C.byref(bi), C.byref(bs))
    if want_all:
        return dist[first:last], shift[first:last], (bd.value, bi.value, bs.value)
    return None, None, (bd.value, bi.value, bs.value)


def sc_detect_from_candidates(query, db, cand, thres):
    query = np.ascontiguousarray(query, dtype=np.float64)
    db = np.ascontiguousarray(db, dtype=np.float64).reshape(-1, DESC)
    cand = np.ascontiguousarray(cand, dtype=np.int32)
    loop, yaw, md, nn = C.c_int(), C.c_float(), C.c_double(), C.c_int()
    lib().sc_detect_from_candidates(_p(query), _p(db), _p(cand), len(cand), thres,
                                    C.byref(loop), C.byref(yaw), C.byref(md), C.byref(nn))
    return loop.value, yaw.value, md.value, nn.value


def sc_knn_bruteforce(keys, query, k):
    keys = np.ascontiguousarray(keys, dtype=np.float32)
    query = np.ascontiguousarray(query, dtype=np.float32)
    idx = np.empty(k, dtype=np.int32)
    d2 = np.empty(k, dtype=np.float32)
    lib().sc_knn_bruteforce(_p(keys), keys.shape[0], keys.shape[1], _p(query), k, _p(idx), _p(d2))
    return idx, d2


def ref_knn(keys, query, k):
    """kNN through the reference's real nanoflann tree (SC/Scancontext.cpp:356,371-373)."""
    R = ref_kdtree()
    if R is None:
        return None
    keys = np.ascontiguousarray(keys, dtype=np.float32)
    query = np.ascontiguousarray(query, dtype=np.float32)
    idx = np.empty(k, dtype=np.int32)
    d2 = np.empty(k, dtype=np.float32)
    R.ref_kdtree_knn(_p(keys), keys.shape[0], keys.shape[1], _p(query), k, _p(idx), _p(d2))
    return idx, d2


# ---------------------------------------------------------------------------------------------
# ORORA solver (oracle/solver_ref.c)
# ---------------------------------------------------------------------------------------------
class OroraParams(C.Structure):
    _fields_ = [("noise_bound", C.c_double), ("cbar2", C.c_double), ("max_iter", C.c_int),
                ("gnc_factor", C.c_double), ("cost_thr", C.c_double), ("nb_radial", C.c_double),
                ("nb_tangential", C.c_double)]


class OroraInfo(C.Structure):
    _fields_ = [("valid", C.c_int), ("n_edges", C.c_int), ("M", C.c_int), ("gnc_iters", C.c_int),
                ("gnc_early_return", C.c_int), ("R", C.c_double * 4), ("t", C.c_double * 2),
                ("cost", C.c_double), ("n_rot_inliers", C.c_int), ("n_final_inliers", C.c_int)]


def default_params():
    """ORORA/config/orora_params.yaml + odometry.cpp:79-135 effective values."""
    return OroraParams(0.75, 1.0, 100, 1.4, 1e-6, 0.1, np.deg2rad(9.0))


def _cols(a):
    """2xN matrix (numpy (2,N) or (N,2)-as-columns) -> flat column-major doubles."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 2 and a.shape[0] == 2:
        return np.ascontiguousarray(a.T).reshape(-1)
    return np.ascontiguousarray(a).reshape(-1)


def orora_doppler(p2d, vel_x, beta=0.049):
    p = _cols(p2d)
    out = np.empty_like(p)
    lib().orora_doppler.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
    lib().orora_doppler.restype = None
    lib().orora_doppler(_p(p), p.size // 2, vel_x, beta, _p(out))
    return out.reshape(-1, 2).T


def orora_adjacency(src, tgt, noise_bound=0.75, cbar2=1.0):
    s, t = _cols(src), _cols(tgt)
    N = s.size // 2
    W = (N + 63) // 64
    adj = np.zeros((N, W), dtype=np.uint64)
    L = lib()
    L.orora_build_adjacency.restype = C.c_int
    L.orora_build_adjacency.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_int]
    e = L.orora_build_adjacency(_p(s), _p(t), N, noise_bound, cbar2, _p(adj), W)
    return adj, e


def orora_max_clique(adj):
    adj = np.ascontiguousarray(adj, dtype=np.uint64)
    N, W = adj.shape
    out = np.zeros(max(N, 1), dtype=np.int32)
    nodes = C.c_longlong()
    L = lib()
    L.orora_max_clique.restype = C.c_int
    L.orora_max_clique.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    n = L.orora_max_clique(_p(adj), N, W, _p(out), C.byref(nodes))
    return out[:n].copy(), nodes.value


def orora_svd_rot2d(X, Y, Wt):
    x, y = _cols(X), _cols(Y)
    w = np.ascontiguousarray(Wt, dtype=np.float64)
    R = np.zeros(4)
    L = lib()
    L.orora_svd_rot2d.restype = None
    L.orora_svd_rot2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.orora_svd_rot2d(_p(x), _p(y), _p(w), w.size, _p(R))
    return R.reshape(2, 2).T     # column-major -> numpy


def orora_cote_axis(X, ranges):
    x = np.ascontiguousarray(X, dtype=np.float64)
    r = np.ascontiguousarray(ranges, dtype=np.float64)
    est = C.c_double()
    inl = np.zeros(x.size, dtype=np.uint8)
    L = lib()
    L.orora_cote_axis.restype = None
    L.orora_cote_axis.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.orora_cote_axis(_p(x), _p(r), x.size, C.byref(est), _p(inl))
    return est.value, inl.astype(bool)


def orora_solve(src, tgt, params=None):
    """NonMinimalSolver::computeTransformation: src = curr matches, tgt = prev matches (2xN).
    Returns dict(T (4,4), clique, final_inliers, info)."""
    params = params or default_params()
    s, t = _cols(src), _cols(tgt)
    N = s.size // 2
    T = np.zeros(16)
    clique = np.zeros(max(N, 1), dtype=np.int32)
    fin = np.zeros(max(N, 1), dtype=np.int32)
    info = OroraInfo()
    L = lib()
    L.orora_solve.restype = C.c_int
    L.orora_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orora_solve(_p(s), _p(t), N, C.byref(params), _p(T), _p(clique), _p(fin), C.byref(info))
    return {"T": T.reshape(4, 4).T.copy(), "clique": clique[:info.M].copy(),
            "final_inliers": fin[:info.n_final_inliers].copy(), "info": info, "valid": bool(info.valid)}


def orora_calc_transform(prev, curr, vel_x, params=None, delta_t=0.25, beta=0.049, use_doppler=True,
                         check_stop_motion=True, stop_thr=700):
    """ORORA::calcTransform.  Returns (result dict, new vel_x)."""
    params = params or default_params()
    pv, cv = _cols(prev), _cols(curr)
    N = pv.size // 2
    T = np.zeros(16)
    clique = np.zeros(max(N, 1), dtype=np.int32)
    fin = np.zeros(max(N, 1), dtype=np.int32)
    info = OroraInfo()
    v = C.c_double(vel_x)
    L = lib()
    L.orora_calc_transform.restype = C.c_int
    L.orora_calc_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double,
                                       C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
    L.orora_calc_transform(_p(pv), _p(cv), N, C.byref(params), delta_t, beta, int(use_doppler),
                           int(check_stop_motion), stop_thr, C.byref(v), _p(T), _p(clique), _p(fin), C.byref(info))
    return ({"T": T.reshape(4, 4).T.copy(), "clique": clique[:info.M].copy(),
             "final_inliers": fin[:info.n_final_inliers].copy(), "info": info, "valid": bool(info.valid)}, v.value)


# -------------------------------------------------------------------------------------------------
# loop verification: pcl::IterativeClosestPoint as doICPVirtualRelative configures it
# (pgo/SC-A-LOAM/src/laserPosegraphOptimization.cpp:355-406).  PCL is not vendored in the reference and not
# installed here: this restates its published algorithm (registration/impl/icp.hpp,
# default_convergence_criteria.hpp, transformation_estimation_svd.hpp) -- PARITY UNPINNED against PCL itself.
# Float points / distances, double sums; nearest neighbours exhaustively, lowest index on ties.
# -------------------------------------------------------------------------------------------------
def _nn(src, tgt):
    best_d = np.full(src.shape[0], np.inf, dtype=np.float32)
    best_i = np.zeros(src.shape[0], dtype=np.int64)
    for a in range(0, tgt.shape[0], 4096):
        t = tgt[a:a + 4096]
        dx = src[:, None, 0] - t[None, :, 0]
        dy = src[:, None, 1] - t[None, :, 1]
        dz = src[:, None, 2] - t[None, :, 2]
        d2 = (dx * dx + dy * dy + dz * dz).astype(np.float32)
        j = np.argmin(d2, axis=1)
        dj = d2[np.arange(src.shape[0]), j]
        upd = dj < best_d
        best_d[upd], best_i[upd] = dj[upd], j[upd] + a
    return best_i, best_d


def _xform32(T, p):
    R, t = T[:3, :3].astype(np.float32), T[:3, 3].astype(np.float32)
    x = R[0, 0] * p[:, 0] + R[0, 1] * p[:, 1] + R[0, 2] * p[:, 2] + t[0]
    y = R[1, 0] * p[:, 0] + R[1, 1] * p[:, 1] + R[1, 2] * p[:, 2] + t[1]
    z = R[2, 0] * p[:, 0] + R[2, 1] * p[:, 1] + R[2, 2] * p[:, 2] + t[2]
    return np.stack([x, y, z], axis=1).astype(np.float32)


def icp_align(src_xyz, tgt_xyz, max_corr_dist=150.0, max_iterations=100, transformation_epsilon=1e-6,
              euclidean_fitness_epsilon=1e-6, fitness_threshold=0.3, min_correspondences=3):
    """-> dict(T (4,4) source->target, converged, iterations, state, fitness, accepted)"""
    src = np.ascontiguousarray(src_xyz, dtype=np.float32).reshape(-1, 3).copy()
    tgt = np.ascontiguousarray(tgt_xyz, dtype=np.float32).reshape(-1, 3)
    final = np.eye(4)
    out = {"T": final, "converged": False, "iterations": 0, "state": 5, "fitness": float("inf"), "accepted": False, "n_corr": 0}
    if src.shape[0] == 0 or tgt.shape[0] == 0:
        return out
    prev_mse, it, state, conv = np.finfo(np.float64).max, 0, 0, False
    while True:
        idx, d2 = _nn(src, tgt)
        keep = d2.astype(np.float64) <= max_corr_dist * max_corr_dist
        n = int(keep.sum())
        out["n_corr"] = n
        if n < min_correspondences:
            state, conv = 5, False
            break
        s, t = src[keep].astype(np.float64), tgt[idx[keep]].astype(np.float64)
        cs, ct = s.sum(0) / n, t.sum(0) / n
        H = s.T @ t - n * np.outer(cs, ct)
        U, _, Vt = np.linalg.svd(H)
        V = Vt.T
        D = np.diag([1.0, 1.0, -1.0 if np.linalg.det(U) * np.linalg.det(V) < 0 else 1.0])
        R = V @ D @ U.T
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, ct - R @ cs
        src = _xform32(T, src)
        final = T @ final
        it += 1
        if it >= max_iterations:
            state, conv = 1, True
            break
        cos_angle = 0.5 * (T[0, 0] + T[1, 1] + T[2, 2] - 1.0)
        tr2 = float(T[0, 3] ** 2 + T[1, 3] ** 2 + T[2, 3] ** 2)
        if cos_angle >= 1.0 - transformation_epsilon and tr2 <= transformation_epsilon:
            state, conv = 2, True
            break
        mse = float(d2[keep].astype(np.float64).sum() / n)
        if abs(mse - prev_mse) < 1e-12:
            state, conv = 3, True
            break
        if abs(mse - prev_mse) / prev_mse < euclidean_fitness_epsilon:
            state, conv = 4, True
            break
        prev_mse = mse
    _, d2 = _nn(src, tgt)
    fitness = float(d2.astype(np.float64).sum() / src.shape[0])
    out.update(T=final, converged=conv, iterations=it, state=state, fitness=fitness,
               accepted=bool(conv and fitness <= fitness_threshold))
    return out
