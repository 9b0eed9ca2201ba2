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


# =================================================================================================
# front-end: RadarLoader (ORORA/include/radar_loader.hpp:35-308)
# =================================================================================================
from ._lib import ExtractParams, FrameInfo  # noqa: E402


def extract_params(dataset):
    p = ExtractParams()
    check(lib().nrs_extract_default_params(dataset.encode(), C.byref(p)))
    return p


def cen2019features(ctx, fft_data, max_points, min_range):
    """features.hpp:36 cen2019features(fft, max_points, min_range, targets) -> targets (3, K)."""
    fft = np.ascontiguousarray(fft_data, dtype=np.float32)
    R, Cc = fft.shape
    cap = 16384
    out = np.zeros(3 * cap, dtype=np.float64)
    K = C.c_int()
    check(lib().nrs_cen2019(ctx.handle, ptr(fft), R, Cc, int(max_points), int(min_range), ptr(out), cap, C.byref(K)))
    return out[: 3 * K.value].reshape(-1, 3).T.copy()


def cen2018features(ctx, fft_data, zq, sigma_gauss, min_range, return_response=False):
    """features.hpp:27 cen2018features(fft, zq, sigma_gauss, min_range, targets) -> targets (3, K)
    (and the per-pixel response y, R x C, when asked -- a stage output for the parity tests)."""
    fft = np.ascontiguousarray(fft_data, dtype=np.float32)
    R, Cc = fft.shape
    cap = 16384
    out = np.zeros(3 * cap, dtype=np.float64)
    y = np.zeros((R, Cc), dtype=np.float32) if return_response else None
    K = C.c_int()
    check(lib().nrs_cen2018(ctx.handle, ptr(fft), R, Cc, float(zq), int(sigma_gauss), int(min_range), ptr(out), cap,
                            C.byref(K), ptr(y) if y is not None else None))
    t = out[: 3 * K.value].reshape(-1, 3).T.copy()
    return (t, y) if return_response else t


class Frame:
    """nrs_frame: the device-resident products of one extractFeature call."""

    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle
        ctx.adopt(self)
        self.info = FrameInfo()
        check(lib().nrs_frame_info_get(self._h, C.byref(self.info)))

    @property
    def handle(self):
        return self._h

    def get(self, want_img=True):
        i = self.info
        W = 964
        polar = np.zeros(3 * max(i.K_polar, 1))
        feat = np.zeros(3 * max(i.K, 1))
        kp = np.zeros(2 * max(i.K_desc, 1), dtype=np.float32)
        t = np.zeros(max(i.K, 1), dtype=np.int64)
        tn = np.zeros(max(i.K, 1))
        desc = np.zeros(32 * max(i.K_desc, 1), dtype=np.uint8)
        img = np.zeros(W * W, dtype=np.uint8) if want_img else None
        check(lib().nrs_frame_get(self._h, ptr(polar), ptr(feat), ptr(kp), ptr(t), ptr(tn), ptr(desc), ptr(img)))
        return {"polar": polar[: 3 * i.K_polar].reshape(-1, 3).T.copy(),
                "feat_in_cart": feat[: 3 * i.K].reshape(-1, 3).T.copy(),
                "kp": kp[: 2 * i.K_desc].reshape(-1, 2).copy(), "t": t[: i.K].copy(), "t_norm": tn[: i.K].copy(),
                "desc": desc[: 32 * i.K_desc].reshape(-1, 32).copy(),
                "img": img.reshape(W, W) if want_img else None}

    def stages(self, R, Cc, W=964):
        fft = np.zeros((R, Cc), dtype=np.float32)
        marked = np.zeros((R, Cc), dtype=np.uint8)
        pre = np.zeros(2 * max(self.info.K_polar, 1))
        cart = np.zeros((W, W), dtype=np.float32)
        check(lib().nrs_frame_get_stages(self._h, ptr(fft), ptr(marked), ptr(pre), ptr(cart)))
        return {"fft": fft, "R": marked, "cart_pre": pre[: 2 * self.info.K_polar].reshape(-1, 2).T.copy(), "cart_f32": cart}

    def close(self):
        if self._h:
            lib().nrs_frame_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RadarLoader:
    """RadarLoader (radar_loader.hpp:35-308): extractFeature + matchFeatures on the GPU."""

    def __init__(self, dataset, keypoint_extraction="cen2019", voxel_size=0.6, use_voxelization=True,
                 viz_extraction=False, ctx=None):
        if dataset not in ("oxford", "mulran"):
            raise ValueError("Note implemented!")                                  # radar_loader.hpp:62
        if keypoint_extraction not in ("cen2018", "cen2019"):
            raise ValueError("Keypoint extraction method seems to be wrong!")      # :64-66
        self.ctx = ctx if ctx is not None else _lib.Context(0)
        self.dataset_ = dataset
        self.keypoint_extraction_ = keypoint_extraction
        self.params = extract_params(dataset)
        self.params.keypoint_extraction = 1 if keypoint_extraction == "cen2018" else 0      # radar_loader.hpp:95-99
        self.params.voxel_size = voxel_size
        self.params.use_voxelization = int(bool(use_voxelization))
        self.time_unit_denominator_ = 1000000.0 if dataset == "oxford" else 1000000000.0

    def extractFeatureFromRows(self, raw_rows, t_criteria=0):
        """raw_rows: the decoded 8-bit PNG (what cv::imread gives load_radar)."""
        raw = np.ascontiguousarray(raw_rows, dtype=np.uint8)
        h = C.c_void_p()
        check(lib().nrs_frame_extract(self.ctx.handle, ptr(raw), raw.shape[0], raw.shape[1], C.byref(self.params),
                                      int(t_criteria), C.byref(h)))
        fr = Frame(self.ctx, h)
        d = fr.get()
        return RadarFeature(img_=d["img"], t_=d["t"], time_unit_denominator_=self.time_unit_denominator_,
                            desc_=d["desc"], kp_=d["kp"], feat_in_cart_=d["feat_in_cart"],
                            feat_in_polar_=d["polar"], t_crietria_=int(t_criteria), t_norm_=d["t_norm"], handle=fr)

    def extractFeature(self, datadir, radar_file_path):
        """radar_loader.hpp:74: reads datadir/radar_file_path (8-bit PNG) and extracts."""
        import cv2
        import os
        raw = cv2.imread(os.path.join(datadir, radar_file_path), cv2.IMREAD_GRAYSCALE)
        if raw is None:
            raise FileNotFoundError(os.path.join(datadir, radar_file_path))
        return self.extractFeatureFromRows(raw, int(radar_file_path.split(".")[0]))

    def matchFeatures(self, prev_feat, curr_feat):
        """radar_loader.hpp:143-205 -> (prev_matched, curr_matched) with 2xN feat_in_cart_."""
        cap = max(prev_feat.handle.info.K_desc, 1)
        q = np.zeros(cap, dtype=np.int32)
        t = np.zeros(cap, dtype=np.int32)
        pxy = np.zeros(2 * cap)
        cxy = np.zeros(2 * cap)
        pt = np.zeros(cap, dtype=np.int64)
        ct = np.zeros(cap, dtype=np.int64)
        n = C.c_int()
        check(lib().nrs_match(self.ctx.handle, prev_feat.handle.handle, curr_feat.handle.handle,
                              float(self.params.nndr), cap, ptr(q), ptr(t), ptr(pxy), ptr(cxy), ptr(pt), ptr(ct),
                              C.byref(n)))
        N = n.value
        q, t = q[:N].copy(), t[:N].copy()
        pm = RadarFeature(img_=prev_feat.img_, t_=pt[:N].copy(), time_unit_denominator_=prev_feat.time_unit_denominator_,
                          desc_=prev_feat.desc_, kp_=prev_feat.kp_[q], feat_in_cart_=pxy[: 2 * N].reshape(-1, 2).T.copy(),
                          feat_in_polar_=prev_feat.feat_in_polar_, t_crietria_=prev_feat.t_crietria_,
                          t_norm_=prev_feat.t_norm_[q], handle=prev_feat.handle)   # :187: indexed by descriptor row, like the reference
        cm = RadarFeature(img_=curr_feat.img_, t_=ct[:N].copy(), time_unit_denominator_=curr_feat.time_unit_denominator_,
                          desc_=curr_feat.desc_, kp_=prev_feat.kp_[np.minimum(t, len(prev_feat.kp_) - 1)] if len(prev_feat.kp_) else curr_feat.kp_[t],
                          feat_in_cart_=cxy[: 2 * N].reshape(-1, 2).T.copy(),
                          feat_in_polar_=curr_feat.feat_in_polar_, t_crietria_=curr_feat.t_crietria_,
                          t_norm_=curr_feat.t_norm_[t], handle=curr_feat.handle)
        pm.query_idx, pm.train_idx = q, t
        return pm, cm


def pair(ctx, prev_frame, raw_rows, extract_p, solver_p, vel_x=0.0, t_criteria=0):
    """nrs_orora_pair: extractFeature(new frame) + matchFeatures(prev, curr) + calcTransform in one call
    (odometry.cpp:153-284).  prev_frame: Frame of the previous scan.  -> (Frame of the new scan, T (4,4),
    number of matches, SolverInfo)."""
    raw = np.ascontiguousarray(raw_rows, dtype=np.uint8)
    h = C.c_void_p()
    T = np.zeros(16)
    n = C.c_int()
    info = SolverInfo()
    check(lib().nrs_orora_pair(ctx.handle, prev_frame.handle, ptr(raw), raw.shape[0], raw.shape[1], C.byref(extract_p),
                               int(t_criteria), C.byref(solver_p), float(vel_x), C.byref(h), ptr(T), C.byref(n),
                               C.byref(info)))
    return Frame(ctx, h), T.reshape(4, 4).T.copy(), n.value, info
