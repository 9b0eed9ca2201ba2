"""Host-side mirror of the reference's ORORA operator API over the C ABI.

* ``RadarFeature``  <-> ORORA/include/radar_loader.hpp:16-33
* ``ORORA``         <-> ORORA/include/ORORA.hpp:13-181 (calcTransform, getMaxCliques, getFinalInliers)
* ``setParams``     <-> ORORA/include/solver/nonminimal_solver.hpp:959-973

Matrices follow the reference's Eigen layout: ``feat_in_cart_`` is a (2|3, N) numpy float64
array (column j = point j).  All arithmetic runs in libnrs_b200.so on the GPU.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import SolverInfo, SolverParams, check, lib, ptr


@dataclass
class RadarFeature:
    """radar_loader.hpp:16-33.  All per-point members have the same length."""
    img_: np.ndarray = None                 # 964x964 uint8 Cartesian image
    t_: np.ndarray = None                   # int64 per-point stamps
    time_unit_denominator_: float = 1.0
    desc_: np.ndarray = None                # K x 32 uint8 rBRIEF descriptors
    kp_: np.ndarray = None                  # K x 2 float32 keypoint (u, v)
    feat_in_cart_: np.ndarray = None        # 3 x K (extract) or 2 x N (matched) float64
    feat_in_polar_: np.ndarray = None       # 3 x K0 (azimuth, range, 1)
    feat_compensated_: np.ndarray = None
    t_crietria_: int = 0
    t_norm_: np.ndarray = None
    handle: object = field(default=None, repr=False)   # device-resident frame (nrs_frame), if any


def setParams(noise_bound, cbar2, estimating_scale, num_max_iter, gnc_factor, rot_cost_thr, params=None):
    """orora::setParams (nonminimal_solver.hpp:959-973).  estimating_scale is accepted and, as
    in the reference (solveForScale forces *scale = 1, :239), has no effect."""
    p = params if params is not None else default_params()
    p.noise_bound = noise_bound
    p.cbar2 = cbar2
    p.rotation_max_iterations = int(num_max_iter)
    p.rotation_gnc_factor = gnc_factor
    p.rotation_cost_threshold = rot_cost_thr
    return p


def default_params():
    p = SolverParams()
    lib().nrs_solver_default_params(C.byref(p))
    return p


def _cols2(a):
    """(2|3, N) matrix -> flat column-major 2 x N doubles (top two rows)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[0] < 2:
        raise ValueError("expected a (2|3, N) point matrix")
    return np.ascontiguousarray(a[:2].T).reshape(-1)


def solve(ctx, prev, curr, params, vel_x=0.0):
    """nrs_orora_solve.  Returns dict(T (4,4), clique, final_inliers, info)."""
    pv, cv = _cols2(prev), _cols2(curr)
    if pv.size != cv.size:
        raise ValueError("prev and curr must hold the same number of points")
    N = pv.size // 2
    T = np.zeros(16)
    clique = np.zeros(max(N, 1), dtype=np.int32)
    fin = np.zeros(max(N, 1), dtype=np.int32)
    info = SolverInfo()
    check(lib().nrs_orora_solve(ctx.handle, ptr(pv), ptr(cv), N, C.byref(params), float(vel_x), ptr(T),
                                ptr(clique), ptr(fin), C.byref(info)))
    return {"T": T.reshape(4, 4).T.copy(), "clique": clique[:info.M].copy(),
            "final_inliers": fin[:info.n_final_inliers].copy(), "info": info, "valid": bool(info.valid)}


class ORORA:
    """orora::ORORA (ORORA.hpp:13-181)."""

    def __init__(self, dataset, deskewing_target, use_doppler_compenstation, use_deskewing,
                 check_stop_motion, num_feat_thr_for_stop_motion, beta, noise_bound_radial,
                 noise_bound_tangential, ctx=None):
        if use_deskewing:
            # motionDeskewing (ORORA.hpp:95-137) needs the rotational part of SE3tose3, whose sign
            # comes out of Eigen's ComplexEigenSolver (association.cpp:182-197); it is off in the
            # shipped configuration (orora_params.yaml:4) and out of this build's scope.
            raise ValueError("use_deskewing=True is not supported (off in the reference's config)")
        self.ctx = ctx if ctx is not None else _lib.Context(0)
        self.dataset_, self.deskewing_target_ = dataset, deskewing_target
        self.use_doppler_compensation_ = bool(use_doppler_compenstation)
        self.check_stop_motion_ = bool(check_stop_motion)
        self.num_feat_thr_for_stop_motion_ = int(num_feat_thr_for_stop_motion)
        self.beta_ = beta
        self.noise_bound_radial_, self.noise_bound_tangential_ = noise_bound_radial, noise_bound_tangential
        self.prev_vel_se3_ = np.zeros(6)
        self.curr_vel_se3_ = np.zeros(6)
        self.T_prev_ = np.eye(4)
        self.T_curr_ = np.eye(4)
        self._last = None

    def calcTransform(self, params, delta_t, prev_matched, curr_matched):
        """ORORA.hpp:37-93.  prev_matched / curr_matched: RadarFeature with 2xN feat_in_cart_."""
        self.prev_vel_se3_ = self.curr_vel_se3_.copy()                       # :42
        p = SolverParams.from_buffer_copy(params)
        p.noise_bound_radial, p.noise_bound_tangential = self.noise_bound_radial_, self.noise_bound_tangential_
        p.use_doppler_compensation = int(self.use_doppler_compensation_)
        p.doppler_beta = self.beta_
        p.check_stop_motion = int(self.check_stop_motion_)
        p.num_feat_thr_for_stop_motion = self.num_feat_thr_for_stop_motion_
        r = solve(self.ctx, prev_matched.feat_in_cart_, curr_matched.feat_in_cart_, p, self.prev_vel_se3_[0])
        self._last = (r, prev_matched, curr_matched)
        output = r["T"]
        self.T_prev_ = output                                                # :88
        vel = np.zeros(6)
        vel[:3] = output[:3, 3] / delta_t                                    # :89, association.cpp:215
        self.curr_vel_se3_ = vel
        self.T_curr_ = output
        return self.T_curr_

    def _pairs(self, ids):
        r, prev_m, curr_m = self._last
        src = np.asarray(curr_m.feat_in_cart_)[:2, ids]
        tgt = np.asarray(prev_m.feat_in_cart_)[:2, ids]
        one = np.ones((1, len(ids)))
        return np.vstack([src, one]), np.vstack([tgt, one])

    def getMaxCliques(self):
        """ORORA.hpp:154-156 -> (source_max_clique, target_max_clique) as 3xM arrays."""
        return self._pairs(self._last[0]["clique"])

    def getFinalInliers(self):
        return self._pairs(self._last[0]["final_inliers"])

    @property
    def last_result(self):
        return self._last[0] if self._last else None
