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
    """rot.eulerAngles(0, 1, 2)(2) of a rotation about z, as Eigen/src/Geometry/EulerAngles.h evaluates it with
    i, j, k = 0, 1, 2: the first angle atan2(R(1,2), R(2,2)) is 0, so s1 = 0, c1 = 1 and the third angle is
    -atan2(s1 R(2,0) - c1 R(1,0), c1 R(1,1) - s1 R(2,1)) = -atan2(0 - R(1,0), R(1,1)) -- written with the subtraction so
    that an exact R(1,0) = +0 gives -0 like there (the node's est_traj.txt starts "0,0,-0"; pinned against the compiled
    reference in tests/test_ref_posemanager.py)"""
    return -math.atan2(0.0 - float(T[1, 0]), float(T[1, 1]))


def mul4(a, b):
    """4 x 4 product in plain index order (k ascending, no fused multiply-add), the arithmetic of the C++ twin's Mat4 --
    numpy's matmul goes through BLAS kernels whose last bit depends on the host"""
    r = np.zeros((4, 4))
    for i in range(4):
        for j in range(4):
            s = 0.0
            for k in range(4):
                s += float(a[i, k]) * float(b[k, j])
            r[i, j] = s
    return r


def inverse_affine(a):
    """inverse of [A t; 0 0 0 1] by cofactors, the same expression order as the C++ twin's Mat4::inverse_affine"""
    f = lambda i, j: float(a[i, j])
    c00 = f(1, 1) * f(2, 2) - f(1, 2) * f(2, 1)
    c01 = f(1, 2) * f(2, 0) - f(1, 0) * f(2, 2)
    c02 = f(1, 0) * f(2, 1) - f(1, 1) * f(2, 0)
    det = f(0, 0) * c00 + f(0, 1) * c01 + f(0, 2) * c02
    r = np.eye(4)
    r[0, 0], r[1, 0], r[2, 0] = c00 / det, c01 / det, c02 / det
    r[0, 1] = (f(0, 2) * f(2, 1) - f(0, 1) * f(2, 2)) / det
    r[1, 1] = (f(0, 0) * f(2, 2) - f(0, 2) * f(2, 0)) / det
    r[2, 1] = (f(0, 1) * f(2, 0) - f(0, 0) * f(2, 1)) / det
    r[0, 2] = (f(0, 1) * f(1, 2) - f(0, 2) * f(1, 1)) / det
    r[1, 2] = (f(0, 2) * f(1, 0) - f(0, 0) * f(1, 2)) / det
    r[2, 2] = (f(0, 0) * f(1, 1) - f(0, 1) * f(1, 0)) / det
    for i in range(3):
        r[i, 3] = -(r[i, 0] * f(0, 3) + r[i, 1] * f(1, 3) + r[i, 2] * f(2, 3))
    return r


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


def loadGTPoses(gt_path):
    """loadGTPoses (:17-36): "(Time, x, y, yaw)" rows without a header -> {stamp: [x, y, yaw]} as float32.  Kept as
    written there: the y entry is read from column 3 like the yaw (`{stof(words[1]), stof(words[3]), stof(words[3])}`),
    so the relative ground truth the node writes and scores against has y == yaw."""
    gt_rel = {}
    with open(gt_path) as f:
        for line in f.read().splitlines():
            words = line.split(",")
            if len(words) < 4:
                raise ValueError(f"loadGTPoses: a row of {gt_path} has fewer than 4 columns")
            gt_rel[int(words[0])] = [float(np.float32(words[1])), float(np.float32(words[3])), float(np.float32(words[3]))]
    if not gt_rel:
        raise ValueError("Some parameters seem to be wrong! Load GT failed!")
    return gt_rel


def get_groundtruth_odometry(gtfile, t1, t2):
    """get_groundtruth_odometry (radar_utils.cpp:483-501): Oxford's radar_odometry.csv, first line = header; the row with
    parts[9] == t1 and parts[8] == t2 gives parts[2..7] = x, y, z, roll, pitch, yaw (float32), else None."""
    with open(gtfile) as f:
        f.readline()
        for line in f:
            parts = line.rstrip("\n").split(",")
            if len(parts) < 10:
                continue
            if int(parts[9]) == t1 and int(parts[8]) == t2:
                return [float(np.float32(v)) for v in parts[2:8]]
    return None


class SE2PoseManager:
    """The reference's SE2PoseManager: the five output files and, with from_seq_dir, the ground-truth reader the node
    uses for their ground-truth columns (otherwise the caller passes the relative ground-truth motions in)."""
    gt_path_ = None
    gt_rel_ = None

    @classmethod
    def from_seq_dir(cls, seq_dir, dataset_name="mulran", algorithm_name="ORORA", keypoint_extraction="cen2018"):
        """The node's form (:73-135 with ANALYZE_PARAM off): ground truth from <seq_dir>/gt/radar_odometry.csv (loaded
        up front for MulRan, searched per pair for Oxford), files under <seq_dir>/outputs/, truncated."""
        pm = cls(os.path.join(seq_dir, "outputs"), dataset_name, algorithm_name, keypoint_extraction)
        pm.gt_path_ = os.path.join(seq_dir, "gt", "radar_odometry.csv")
        if dataset_name == "mulran":
            pm.gt_rel_ = loadGTPoses(pm.gt_path_)
        return pm

    def getXYYawPose(self, time1, time2=None):
        """MulRan: getXYYawPose(t) (:137-142); Oxford: getXYYawPose(time1, time2) (:159-166)"""
        if time2 is None:
            if time1 not in self.gt_rel_:
                raise ValueError("The corresponding time is not inserted!... exiting.")
            return self.gt_rel_[time1]
        gtvec = get_groundtruth_odometry(self.gt_path_, time1, time2)
        if gtvec is None:
            raise ValueError("Stop!!")
        return [gtvec[0], gtvec[1], gtvec[5]]

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
        Trel = mul4(inverse_affine(Tgt), Test)
        ts_error = math.sqrt(Trel[0, 3] ** 2 + Trel[1, 3] ** 2)
        x = (Trel[0, 0] + Trel[1, 1] + Trel[2, 2] - 1) / 2 + EPSILON_ROT
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
