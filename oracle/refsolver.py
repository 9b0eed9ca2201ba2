"""Bridge to oracle/_ref/libref_solver.so -- the REFERENCE's own estimator (ORORA.hpp, solver/nonminimal_solver.hpp,
solver/graph.h, src/graph.cc compiled in place; oracle/Makefile `ref`; stand-in headers in oracle/refshim/).  Pins the
reference's LOGIC against oracle/solver_ref.c (identical index sets, poses to north_star's tolerance); Eigen's evaluation
order and PMC are not reproduced by the stand-ins (see oracle/ref_solver_wrap.cpp).  The exact clique search behind the
PMC stand-in is the oracle's own routine, handed in here as a callback.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

from . import pyoracle

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_solver.so")
_lib = None
_keep = None
_CLIQUE = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_int))


def available():
    return os.path.exists(_PATH)


def _cb_clique(adj, n, out):
    if n == 0:
        return 0
    a = np.ctypeslib.as_array(adj, shape=(n, n)).astype(bool)
    W = (n + 63) // 64
    bits = np.zeros((n, W), dtype=np.uint64)
    for w in range(W):
        blk = a[:, 64 * w:64 * (w + 1)]
        bits[:, w] = (blk.astype(np.uint64) << np.arange(blk.shape[1], dtype=np.uint64)).sum(axis=1, dtype=np.uint64)
    clique, _ = pyoracle.orora_max_clique(bits)
    for i, v in enumerate(clique):
        out[i] = int(v)
    return len(clique)


def lib():
    global _lib, _keep
    if _lib is None:
        L = C.CDLL(_PATH)
        _keep = _CLIQUE(_cb_clique)
        L.ref_solver_set_max_clique(_keep)
        L.ref_solver_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                     C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_orora_create.restype = C.c_void_p
        L.ref_orora_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.ref_orora_destroy.argtypes = [C.c_void_p]
        L.ref_orora_calc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double,
                                     C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def solve(src, tgt, params=None):
    """NonMinimalSolver::reset / setInputSource / setInputTarget / computeTransformation on 2 x N points."""
    P = params or pyoracle.default_params()
    s, t = pyoracle._cols(src), pyoracle._cols(tgt)
    n = s.size // 2
    T = np.zeros(16)
    clique = np.zeros(max(n, 1), dtype=np.int32)
    fin = np.zeros(max(n, 1), dtype=np.int32)
    counts = np.zeros(6, dtype=np.int32)
    lib().ref_solver_run(_p(s), _p(t), n, P.noise_bound, P.cbar2, P.max_iter, P.gnc_factor, P.cost_thr, P.nb_radial, P.nb_tangential,
                         _p(T), _p(clique), _p(fin), _p(counts))
    return {"T": T.reshape(4, 4).T.copy(), "valid": bool(counts[0]), "n_edges": int(counts[1]), "clique": clique[:counts[2]].copy(),
            "n_rot_inliers": int(counts[3]), "final_inliers": fin[:counts[4]].copy()}


class ORORA:
    """orora::ORORA -- keeps the Doppler velocity from pair to pair like the node's object does."""

    def __init__(self, dataset="mulran", deskewing_target="rot", use_doppler=True, use_deskewing=False, check_stop_motion=True,
                 num_feat_thr=700, beta=0.049, noise_bound_radial=0.1, noise_bound_tangential=np.deg2rad(9.0)):
        self.h = lib().ref_orora_create(dataset.encode(), deskewing_target.encode(), int(use_doppler), int(use_deskewing),
                                        int(check_stop_motion), int(num_feat_thr), beta, noise_bound_radial, noise_bound_tangential)

    def close(self):
        if self.h:
            lib().ref_orora_destroy(self.h)
            self.h = None

    def calcTransform(self, prev, curr, delta_t=0.25, params=None):
        P = params or pyoracle.default_params()
        pv, cv = pyoracle._cols(prev), pyoracle._cols(curr)
        n = pv.size // 2
        T = np.zeros(16)
        pc, cc = np.zeros(max(2 * n, 1)), np.zeros(max(2 * n, 1))
        lib().ref_orora_calc(self.h, _p(pv), _p(cv), n, P.noise_bound, P.cbar2, P.max_iter, P.gnc_factor, P.cost_thr, delta_t,
                             _p(T), _p(pc), _p(cc))
        return T.reshape(4, 4).T.copy(), pc[:2 * n].reshape(-1, 2).T.copy(), cc[:2 * n].reshape(-1, 2).T.copy()
