"""ctypes binding of libnrs_b200.so -- fails loudly when the extension is missing."""
import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "libnrs_b200.so")
_HEADER = os.path.join(_ROOT, "include", "nrs_b200.h")

NRS_OK, NRS_EINVAL, NRS_ENODEVICE, NRS_ECUDA, NRS_ENOMEM, NRS_ECAPACITY, NRS_ENCCL = 0, -1, -2, -3, -4, -5, -6


class NrsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nrs error {code}: {msg}")
        self.code = code


class ScResult(C.Structure):
    _fields_ = [("dist", C.c_double), ("idx", C.c_int32), ("shift", C.c_int32)]


class SolverParams(C.Structure):
    """nrs_solver_params (include/nrs_b200.h)."""
    _fields_ = [("noise_bound", C.c_double), ("cbar2", C.c_double), ("rotation_max_iterations", C.c_int32),
                ("rotation_gnc_factor", C.c_double), ("rotation_cost_threshold", C.c_double),
                ("noise_bound_radial", C.c_double), ("noise_bound_tangential", C.c_double),
                ("use_doppler_compensation", C.c_int32), ("doppler_beta", C.c_double),
                ("check_stop_motion", C.c_int32), ("num_feat_thr_for_stop_motion", C.c_int32),
                ("clique_node_limit", C.c_int64)]


class SolverInfo(C.Structure):
    """nrs_solver_info (include/nrs_b200.h)."""
    _fields_ = [("valid", C.c_int32), ("stop_motion", C.c_int32), ("n_edges", C.c_int32), ("M", C.c_int32),
                ("gnc_iters", C.c_int32), ("gnc_early_return", C.c_int32), ("n_rot_inliers", C.c_int32),
                ("n_final_inliers", C.c_int32), ("clique_exact", C.c_int32), ("pad_", C.c_int32),
                ("clique_nodes", C.c_int64), ("R", C.c_double * 4), ("t", C.c_double * 2), ("cost", C.c_double)]


class ExtractParams(C.Structure):
    """nrs_extract_params (include/nrs_b200.h)."""
    _fields_ = [("navtech_version", C.c_int32), ("range_bins", C.c_int32), ("radar_resolution", C.c_float),
                ("cart_resolution", C.c_float), ("cart_pixel_width", C.c_int32), ("min_range", C.c_int32),
                ("max_points", C.c_int32), ("patch_size", C.c_int32), ("use_voxelization", C.c_int32),
                ("interpolate_crossover", C.c_int32), ("voxel_size", C.c_double), ("nndr", C.c_float),
                ("keypoint_extraction", C.c_int32), ("zq", C.c_float), ("sigma_gauss", C.c_int32)]


class IcpParams(C.Structure):
    """nrs_icp_params (include/nrs_b200.h)."""
    _fields_ = [("max_correspondence_distance", C.c_double), ("max_iterations", C.c_int32),
                ("transformation_epsilon", C.c_double), ("euclidean_fitness_epsilon", C.c_double),
                ("fitness_threshold", C.c_double), ("min_correspondences", C.c_int32)]


class IcpInfo(C.Structure):
    """nrs_icp_info (include/nrs_b200.h)."""
    _fields_ = [("converged", C.c_int32), ("iterations", C.c_int32), ("state", C.c_int32), ("n_correspondences", C.c_int32),
                ("fitness", C.c_double), ("accepted", C.c_int32), ("pad_", C.c_int32)]


class FrameInfo(C.Structure):
    """nrs_frame_info (include/nrs_b200.h)."""
    _fields_ = [("n_candidates", C.c_int32), ("j_consumed", C.c_int32), ("l_regions", C.c_int32),
                ("K_polar", C.c_int32), ("K_voxel", C.c_int32), ("K", C.c_int32), ("K_desc", C.c_int32),
                ("n_sorted", C.c_int32), ("maxg", C.c_float),
                ("mean", C.c_float), ("mean_h", C.c_float), ("cart_max", C.c_float), ("t_criteria", C.c_int64)]


def library_path():
    return _SO


def declared_symbols():
    """Every function name include/nrs_b200.h declares (used by the CPU-side export test)."""
    text = open(_HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nrs_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib():
    """Load the CUDA library.  There is deliberately no fallback: a missing .so is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise NrsError(NRS_ENODEVICE,
                           f"{_SO} not built -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "this package has no CPU fallback")
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


def _declare(L):
    vp, i32, f64 = C.c_void_p, C.c_int, C.c_double
    L.nrs_version.restype = i32
    L.nrs_last_error.restype = C.c_char_p
    L.nrs_kernel_launches.restype = C.c_uint64
    L.nrs_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.nrs_ctx_destroy.argtypes = [vp]
    L.nrs_ctx_sync.argtypes = [vp]
    L.nrs_ctx_stream.restype = vp
    L.nrs_ctx_stream.argtypes = [vp]
    L.nrs_ctx_device.argtypes = [vp]
    L.nrs_ctx_sm_count.argtypes = [vp]
    L.nrs_sc_make.argtypes = [vp, vp, i32, f64, f64, vp]
    L.nrs_sc_keys.argtypes = [vp, vp, vp, vp]
    L.nrs_sc_pair.argtypes = [vp, vp, vp, vp, vp]
    L.nrs_sc_fast_align.argtypes = [vp, vp, vp, vp]
    L.nrs_sc_dist_direct.argtypes = [vp, vp, vp, vp]
    L.nrs_version.argtypes = []
    L.nrs_last_error.argtypes = []
    L.nrs_kernel_launches.argtypes = []
    L.nrs_sc_db_create.argtypes = [vp, i32, C.POINTER(vp)]
    L.nrs_sc_db_destroy.argtypes = [vp]
    L.nrs_sc_db_size.argtypes = [vp]
    L.nrs_sc_db_clear.argtypes = [vp]
    L.nrs_sc_db_append.argtypes = [vp, vp, i32]
    L.nrs_sc_db_append_cloud.argtypes = [vp, vp, i32, f64, f64, vp]
    L.nrs_sc_db_read.argtypes = [vp, i32, i32, vp]
    L.nrs_sc_db_device_ptr.restype = vp
    L.nrs_sc_db_device_ptr.argtypes = [vp]
    L.nrs_sc_filter_record_bytes.argtypes = []
    L.nrs_orora_last_timings.argtypes = [vp]
    L.nrs_sc_knn.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.nrs_sc_detect_knn.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.nrs_sc_detect_knn_batch.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.nrs_sc_descriptor_is_binary.argtypes = [vp]
    L.nrs_sc_xchg_create.argtypes = [vp, i32, i32, vp]
    L.nrs_sc_xchg_open.argtypes = [vp, vp]
    L.nrs_sc_query_dev_sharded.argtypes = [vp, vp, i32, i32, vp]
    L.nrs_sc_db_tail_times.argtypes = [vp, vp]
    L.nrs_sc_query.argtypes = [vp, vp, i32, i32, vp, i32, vp]
    L.nrs_sc_query_all.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    L.nrs_sc_query_dev.argtypes = [vp, vp, i32, i32, i32, vp]
    L.nrs_sc_query_batch.argtypes = [vp, vp, i32, i32, i32, vp]
    L.nrs_sc_query_batch_dev.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    L.nrs_sc_query_batch_dev_sharded.argtypes = [vp, vp, i32, i32, i32, vp]
    L.nrs_sc_xchg_set_timeout_ms.argtypes = [vp, i32]
    L.nrs_nccl_available.argtypes = []
    L.nrs_nccl_get_unique_id.argtypes = [vp]
    L.nrs_sc_nccl_init.argtypes = [vp, i32, i32, vp]
    L.nrs_sc_nccl_attach.argtypes = [vp, vp, i32, i32]
    L.nrs_sc_query_batch_dev_nccl.argtypes = [vp, vp, i32, i32, i32, vp]
    L.nrs_sc_query_batch_sharded.argtypes = [vp, vp, i32, i32, i32, vp]
    L.nrs_sc_db_is_binary.argtypes = [vp]
    L.nrs_sc_db_set_search_mode.argtypes = [vp, i32]
    L.nrs_sc_bin_record_bytes.argtypes = []
    L.nrs_solver_default_params.restype = None
    L.nrs_solver_default_params.argtypes = [vp]
    L.nrs_orora_solve.argtypes = [vp, vp, vp, i32, vp, f64, vp, vp, vp, vp]
    L.nrs_extract_default_params.argtypes = [C.c_char_p, vp]
    L.nrs_cen2019.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, vp]
    L.nrs_cen2018.argtypes = [vp, vp, i32, i32, C.c_float, i32, i32, vp, i32, vp, vp]
    L.nrs_frame_extract.argtypes = [vp, vp, i32, i32, vp, C.c_int64, C.POINTER(vp)]
    L.nrs_frame_destroy.argtypes = [vp]
    L.nrs_frame_info_get.argtypes = [vp, vp]
    L.nrs_frontend_tail_times.argtypes = [vp, vp]
    L.nrs_frame_get.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.nrs_frame_get_stages.argtypes = [vp, vp, vp, vp, vp]
    L.nrs_match.argtypes = [vp, vp, vp, C.c_float, i32, vp, vp, vp, vp, vp, vp, vp]
    L.nrs_icp_default_params.restype = None
    L.nrs_icp_default_params.argtypes = [vp]
    L.nrs_icp_align.argtypes = [vp, vp, i32, vp, i32, vp, vp, vp]
    L.nrs_orora_pair.argtypes = [vp, vp, vp, i32, i32, vp, C.c_int64, vp, f64, C.POINTER(vp), vp, vp, vp]
    for name in declared_symbols():
        fn = getattr(L, name, None)
        if fn is not None and fn.restype is C.c_int and name not in ("nrs_version",):
            pass


def check(rc):
    if rc < 0:
        raise NrsError(rc, lib().nrs_last_error().decode(errors="replace"))
    return rc


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} doubles, got {a.size}")
    return a


class Context:
    """One CUDA device + stream (nrs_ctx)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self._children = []          # weakrefs of objects whose native handles hang off this ctx
        check(lib().nrs_ctx_create(device, C.byref(self._h)))

    def adopt(self, obj):
        """Register a dependent (database, frame, ...) so close() tears it down first."""
        import weakref
        if len(self._children) >= 64:                 # a long stream adopts one frame per scan: drop the dead ones
            self._children = [r for r in self._children if r() is not None]
        self._children.append(weakref.ref(obj))

    @property
    def handle(self):
        return self._h

    @property
    def stream(self):
        return lib().nrs_ctx_stream(self._h)

    @property
    def sm_count(self):
        return lib().nrs_ctx_sm_count(self._h)

    @property
    def device(self):
        return lib().nrs_ctx_device(self._h)

    def sync(self):
        check(lib().nrs_ctx_sync(self._h))

    def close(self):
        if self._h:
            for ref in self._children:
                obj = ref()
                if obj is not None:
                    obj.close()
            self._children = []
            lib().nrs_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
