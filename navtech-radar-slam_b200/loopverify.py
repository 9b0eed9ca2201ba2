"""Loop verification behind the ScanContext detector: host mirror of doICPVirtualRelative
(pgo/SC-A-LOAM/src/laserPosegraphOptimization.cpp:355-406) over the C ABI.  The ICP itself (nearest neighbours,
rigid fits, convergence, fitness) runs in libnrs_b200.so on the GPU (icp_kernels.cu); assembling the submap
(loopFindNearKeyframesCloud :332-352: +-25 keyframes moved by the loop keyframe's pose, VoxelGrid 0.4 m) is
host bookkeeping like in the node."""
import ctypes as C

import numpy as np

from ._lib import IcpInfo, IcpParams, check, lib, ptr


def default_params():
    p = IcpParams()
    lib().nrs_icp_default_params(C.byref(p))
    return p


def icp_align(ctx, src_xyz, tgt_xyz, params=None):
    """pcl::IterativeClosestPoint::align as the node configures it -> dict(T (4,4) source->target, converged,
    iterations, state, fitness, accepted)."""
    src = np.ascontiguousarray(src_xyz, dtype=np.float32).reshape(-1, 3)
    tgt = np.ascontiguousarray(tgt_xyz, dtype=np.float32).reshape(-1, 3)
    p = params if params is not None else default_params()
    T = np.zeros(16)
    info = IcpInfo()
    check(lib().nrs_icp_align(ctx.handle, ptr(src), src.shape[0], ptr(tgt), tgt.shape[0], C.byref(p), ptr(T), C.byref(info)))
    return {"T": T.reshape(4, 4).T.copy(), "converged": bool(info.converged), "iterations": info.iterations,
            "state": info.state, "fitness": info.fitness, "accepted": bool(info.accepted), "n_corr": info.n_correspondences}


def voxel_downsample(xyz, leaf):
    """pcl::VoxelGrid (downSizeFilterICP, leaf 0.4: laserPosegraphOptimization.cpp:687-688): centroid per occupied
    voxel in ascending voxel index, float like PCL (host restatement, the same semantics as the front-end's voxelize)."""
    p = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    if p.shape[0] == 0:
        return p
    inv = np.float32(1.0) / np.float32(leaf)
    mn = np.floor(p.min(0) * inv).astype(np.int64)
    mx = np.floor(p.max(0) * inv).astype(np.int64)
    div = mx - mn + 1
    ijk = np.floor(p * inv).astype(np.int64) - mn
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(key, kind="stable")
    key, p = key[order], p[order]
    starts = np.flatnonzero(np.concatenate([[True], key[1:] != key[:-1]]))
    out = np.empty((len(starts), 3), dtype=np.float32)
    ends = list(starts[1:]) + [len(key)]
    for n, (a, b) in enumerate(zip(starts, ends)):
        acc = np.zeros(3, dtype=np.float32)
        for q in p[a:b]:
            acc = acc + q                        # float running sums in point order, like PCL
        out[n] = acc / np.float32(b - a)
    return out


def find_near_keyframes_cloud(keyframe_clouds, keyframe_poses, key, submap_size, root_idx, leaf=0.4):
    """loopFindNearKeyframesCloud (:332-352): keyframes key-submap_size..key+submap_size, each moved by the ROOT
    keyframe's pose (local2global with keyframePosesUpdated[root_idx]), stacked, then downsampled."""
    T = np.asarray(keyframe_poses[root_idx], dtype=np.float32)
    parts = []
    for i in range(-submap_size, submap_size + 1):
        k = key + i
        if k < 0 or k >= len(keyframe_clouds):
            continue
        c = np.asarray(keyframe_clouds[k], dtype=np.float32).reshape(-1, 3)
        parts.append((c @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
    if not parts:
        return np.zeros((0, 3), dtype=np.float32)
    return voxel_downsample(np.concatenate(parts), leaf)


def doICPVirtualRelative(ctx, keyframe_clouds, keyframe_poses, loop_kf_idx, curr_kf_idx, params=None):
    """:355-406 -> None when the fitness test fails, else the 4x4 the node turns into poseFrom.between(poseTo)
    (the inverse of ICP's final transformation)."""
    cur = find_near_keyframes_cloud(keyframe_clouds, keyframe_poses, curr_kf_idx, 0, loop_kf_idx)
    tgt = find_near_keyframes_cloud(keyframe_clouds, keyframe_poses, loop_kf_idx, 25, loop_kf_idx)
    r = icp_align(ctx, cur, tgt, params)
    if not r["accepted"]:
        return None, r
    return np.linalg.inv(r["T"]), r
