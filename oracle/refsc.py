"""Bridge to oracle/_ref/libref_sc.so -- the REFERENCE's own SCManager: pgo/SC-A-LOAM/include/scancontext/Scancontext.cpp
compiled in place with its vendored nanoflann (oracle/Makefile `ref`, stand-in headers in oracle/refshim/; the Eigen
stand-in evaluates norm / dot / mean in Eigen 3.3's SSE2 redux order, recalled).  Used to PIN oracle/sc_ref.c:
tests/test_ref_sc.py.  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py."""
import ctypes as C
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_sc.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        L.ref_sc_create.restype = C.c_void_p
        for name, res, args in (("ref_sc_destroy", None, [C.c_void_p]), ("ref_sc_set_thres", None, [C.c_void_p, C.c_double]),
                                ("ref_sc_size", C.c_int, [C.c_void_p]),
                                ("ref_sc_make", None, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
                                ("ref_sc_keys", None, [C.c_void_p] * 4), ("ref_sc_fast_align", C.c_int, [C.c_void_p] * 3),
                                ("ref_sc_dist_direct", C.c_double, [C.c_void_p] * 3),
                                ("ref_sc_distance", C.c_double, [C.c_void_p] * 3 + [C.POINTER(C.c_int)]),
                                ("ref_sc_circshift", None, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
                                ("ref_sc_make_and_save", None, [C.c_void_p, C.c_void_p, C.c_int]),
                                ("ref_sc_save", None, [C.c_void_p, C.c_void_p]),
                                ("ref_sc_get", None, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
                                ("ref_sc_detect", C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
                                ("ref_sc_detect_between", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)])):
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1)


class SCManager:
    """Scancontext.h:57-122.  Descriptors travel as 1200 doubles, column-major (ring fastest), like oracle.pyoracle."""

    def __init__(self):
        self.h = lib().ref_sc_create()

    def close(self):
        if self.h:
            lib().ref_sc_destroy(self.h)
            self.h = None

    def setSCdistThres(self, t):
        lib().ref_sc_set_thres(self.h, float(t))

    def size(self):
        return lib().ref_sc_size(self.h)

    def makeScancontext(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        out = np.empty(1200)
        lib().ref_sc_make(self.h, _p(xyz), len(xyz), _p(out))
        return out

    def keys(self, desc):
        desc = _d(desc)
        r, s = np.empty(20), np.empty(60)
        lib().ref_sc_keys(self.h, _p(desc), _p(r), _p(s))
        return r, s

    def fastAlignUsingVkey(self, v1, v2):
        v1, v2 = _d(v1), _d(v2)
        return lib().ref_sc_fast_align(self.h, _p(v1), _p(v2))

    def distDirectSC(self, a, b):
        a, b = _d(a), _d(b)
        return lib().ref_sc_dist_direct(self.h, _p(a), _p(b))

    def distanceBtnScanContext(self, a, b):
        a, b = _d(a), _d(b)
        s = C.c_int()
        d = lib().ref_sc_distance(self.h, _p(a), _p(b), C.byref(s))
        return d, s.value

    def makeAndSaveScancontextAndKeys(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        lib().ref_sc_make_and_save(self.h, _p(xyz), len(xyz))

    def saveScancontextAndKeys(self, desc):
        desc = _d(desc)
        lib().ref_sc_save(self.h, _p(desc))

    def get(self, idx):
        d, k = np.empty(1200), np.empty(20, dtype=np.float32)
        lib().ref_sc_get(self.h, int(idx), _p(d), _p(k))
        return d, k

    def detectLoopClosureID(self):
        yaw = C.c_float()
        i = lib().ref_sc_detect(self.h, C.byref(yaw))
        return i, yaw.value

    def detectLoopClosureIDBetweenSession(self, key20, desc):
        key20 = np.ascontiguousarray(key20, dtype=np.float32)
        desc = _d(desc)
        yaw = C.c_float()
        i = lib().ref_sc_detect_between(self.h, _p(key20), _p(desc), C.byref(yaw))
        return i, yaw.value


def circshift(mat_colmajor, rows, cols, shift):
    m = _d(mat_colmajor)
    out = np.empty_like(m)
    lib().ref_sc_circshift(_p(m), rows, cols, int(shift), _p(out))
    return out
