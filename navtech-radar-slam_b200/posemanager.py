"""The text outputs of the reference's orora_odom node, byte-compatible with what its evaluation scripts read
(outlier-robust-radar-odometry/include/SE2posemanager.hpp:151-245, written from odometry.cpp:281-285):

    <prefix>rel_error_odom.txt   ts_error,rot_error,ts_perc_error,rot_deg_per_m_error,ts_abs        (writePose :194-214)
    <prefix>eval_odom.txt        tsx_gt,tsy_gt,yaw_gt,tsx_test,tsy_test,yaw_test,time1/den,time2/den (writePose :216-229)
    <prefix>time_taken.txt       time_taken[us],is_stopped                                           (writeTimeTaken :231-245)
    <prefix>est_traj.txt / gt_traj.txt   x,y,yaw                                                     (writeTraj :247-259)

Numbers are formatted like C++'s default ostream << double (6 significant digits, %g); eval_odom.txt is what
scripts/evaluate_odometry.py parses (np.loadtxt(delimiter=","), first row skipped).  Host-side bookkeeping only --
no arithmetic of the hot path lives here."""
import math
import os

import numpy as np

EPSILON_ROT = 0.00000000001      # SE2posemanager.hpp:13
EPSILON_TRANS = 0.001            # :14


def _g(v):
    """C++ default stream formatting of a double / int"""
    if isinstance(v, (int, np.integer)) and not isinstance(v, bool):
        return str(int(v))
    return format(float(v), ".6g")


def euler_yaw(T):
    """rot.eulerAngles(0, 1, 2)(2) of a planar rotation (Eigen's convention: first angle in [0, pi])"""
    return math.atan2(T[1, 0], T[1, 1])


def xyyaw2se2(xyyaw_rel, dataset):
    """XYYaw2SE2 (:176-188): the MulRan ground truth enters with its yaw negated, like the reference.  The yaw goes
    through tf::createQuaternionFromYaw (double: z = sin(yaw/2), w = cos(yaw/2), x = y = 0), is narrowed to an
    Eigen::Quaternionf and becomes a float rotation matrix (Quaternion::toRotationMatrix) that is widened again."""
    f32 = np.float32
    T = np.eye(4)
    T[0, 3], T[1, 3] = float(f32(xyyaw_rel[0])), float(f32(xyyaw_rel[1]))
    yaw = float(f32(xyyaw_rel[2]))
    if dataset == "mulran":
        yaw = -yaw
    z, w = f32(math.sin(yaw * 0.5)), f32(math.cos(yaw * 0.5))
    tz = f32(2.0) * z
    twz, tzz = tz * w, tz * z
    T[0, 0] = T[1, 1] = float(f32(1.0) - (f32(0.0) + tzz))
    T[0, 1], T[1, 0] = float(f32(0.0) - twz), float(f32(0.0) + twz)
    return T


class SE2PoseManager:
    """Writer half of the reference's SE2PoseManager (ground-truth loading is the caller's business: pass the
    relative ground-truth motions in)."""

    def __init__(self, out_dir, dataset_name="mulran", algorithm_name="ORORA", keypoint_extraction="cen2018"):
        os.makedirs(out_dir, exist_ok=True)
        common = os.path.join(out_dir, f"{dataset_name}_{algorithm_name}_{keypoint_extraction}_")     # :107-112
        self.dataset_name_ = dataset_name
        self.output_pose_error_path_ = common + "rel_error_odom.txt"      # setFilenames :151-157
        self.output_pose_path_ = common + "eval_odom.txt"
        self.output_time_path_ = common + "time_taken.txt"
        self.est_traj_path_ = common + "est_traj.txt"
        self.gt_traj_path_ = common + "gt_traj.txt"
        for p in (self.output_pose_error_path_, self.output_pose_path_, self.output_time_path_, self.est_traj_path_,
                  self.gt_traj_path_):
            open(p, "w").close()                                           # clearOutputFile :190-194

    def writePoseError(self, Tgt, Test):
        """writePose(Tgt, Test) :194-214 -> [ts_error, rot_error]"""
        Trel = np.linalg.inv(Tgt) @ Test
        ts_error = math.sqrt(Trel[0, 3] ** 2 + Trel[1, 3] ** 2)
        x = (np.trace(Trel[:3, :3]) - 1) / 2 + EPSILON_ROT
        rot_error = math.degrees(math.acos(x)) if -1.0 <= x <= 1.0 else 0.0      # isnan -> 0 (:200-203)
        ts_abs = math.sqrt(Tgt[0, 3] ** 2 + Tgt[1, 3] ** 2)
        ts_perc = 0 if ts_abs < EPSILON_TRANS else ts_error / ts_abs * 100.0
        rot_per_m = 0 if ts_abs < EPSILON_TRANS else rot_error / ts_abs
        with open(self.output_pose_error_path_, "a") as f:
            f.write(",".join(_g(v) for v in (ts_error, rot_error, ts_perc, rot_per_m, ts_abs)) + "\n")
        return [ts_error, rot_error]

    def writePose(self, Tgt, Test, time1, time2, time_unit_denominator):
        """writePose(Tgt, Test, time1, time2, den) :216-229"""
        row = (Tgt[0, 3], Tgt[1, 3], -1 * math.asin(Tgt[0, 1]), Test[0, 3], Test[1, 3], -1 * math.asin(max(-1.0, min(1.0, Test[0, 1]))),
               time1 / time_unit_denominator, time2 / time_unit_denominator)
        with open(self.output_pose_path_, "a") as f:
            f.write(",".join(_g(v) for v in row) + "\n")

    def writeTimeTaken(self, time_taken_us, Tgt):
        """writeTimeTaken :231-245"""
        rel_ts = math.sqrt(Tgt[0, 3] ** 2 + Tgt[1, 3] ** 2)
        rel_rot = euler_yaw(Tgt)
        is_stopped = rel_ts < 0.003 and rel_rot < 0.001
        with open(self.output_time_path_, "a") as f:
            f.write(f"{int(time_taken_us)},{1 if is_stopped else 0}\n")

    def writeTraj(self, Tf4x4, path):
        """writeTraj :247-259"""
        with open(path, "a") as f:
            f.write(",".join(_g(v) for v in (Tf4x4[0, 3], Tf4x4[1, 3], euler_yaw(Tf4x4))) + "\n")
