// ref_posemanager_wrap.cpp -- extern "C" doors into the REFERENCE's own SE2PoseManager, compiled from the header where it
// lies (outlier-robust-radar-odometry/include/SE2posemanager.hpp, unmodified; get_groundtruth_odometry comes from the
// reference's src/radar_utils.cpp) against the stand-in headers of oracle/refshim/ (Eigen, tf, boost::format; the real
// libstdc++ <experimental/filesystem> and ofstream formatting).  ref_pm_pair replays what the node does per scan pair
// (odometry.cpp:253-285).  Pins the evaluation FILES (names, formats, the values' definitions, the MulRan ground-truth
// reader with its column indexing) that posemanager.py / cpp/se2_posemanager.hpp write: tests/test_ref_posemanager.py.
// Eigen's eulerAngles / Quaternionf::toRotationMatrix / inverse are stand-in restatements (recalled).
// Recipe: oracle/Makefile `ref`; output oracle/_ref/libref_posemanager.so.  TEST INFRASTRUCTURE ONLY.
#include <Eigen/Dense>
#include <opencv2/core.hpp>
#include <tf/tf.h>

#include <cstring>
#include <fstream>
#include <iostream>
#include <unordered_map>

#include "radar_utils.hpp"
#include "SE2posemanager.hpp"

namespace {
struct Session {
    SE2PoseManager pm;
    Eigen::Matrix4d curr_gt = Eigen::Matrix4d::Identity(), curr_odom = Eigen::Matrix4d::Identity();
    std::string dataset;
};
}  // namespace

extern "C" {

RefCvCallbacks g_ref_cv;

// SE2PoseManager(seq_dir, dataset, algorithm, keypoint_extraction, voxel_size, noise_bound, noise_bound_radial,
// noise_bound_tangential) as odometry.cpp:143-144 constructs it; NULL when it throws (e.g. no ground truth for MulRan)
void *ref_pm_create(const char *seq_dir, const char *dataset, const char *algorithm, const char *keypoint_extraction, double voxel_size,
                    double noise_bound, double noise_bound_radial, double noise_bound_tangential)
{
    try {
        Session *s = new Session{SE2PoseManager(seq_dir, dataset, algorithm, keypoint_extraction, voxel_size, noise_bound, noise_bound_radial,
                                                noise_bound_tangential)};
        s->dataset = dataset;
        return s;
    } catch (const std::exception &) {
        return nullptr;
    }
}
void ref_pm_destroy(void *h) { delete (Session *)h; }

// odometry.cpp:253-285 for one scan pair: ground truth by time stamp, XYYaw2SE2, the two chained poses, the five writes.
// Returns 0, or 1 when the ground-truth lookup throws.
int ref_pm_pair(void *h, int64_t time1, int64_t time2, const double *Test16, int time_taken, double time_unit_denominator)
{
    Session &s = *(Session *)h;
    Eigen::Matrix4d Test, Tgt = Eigen::Matrix4d::Identity();
    std::memcpy(Test.data(), Test16, sizeof(double) * 16);
    std::vector<float> xyyaw_rel;
    try {
        if (s.dataset == "oxford") xyyaw_rel = s.pm.getXYYawPose(time1, time2);
        else if (s.dataset == "mulran") xyyaw_rel = s.pm.getXYYawPose(time1);
    } catch (const std::exception &) {
        return 1;
    }
    Eigen::Matrix4d TgtTmp = Eigen::Matrix4d::Identity();
    s.pm.XYYaw2SE2(xyyaw_rel, TgtTmp);
    Tgt = TgtTmp;
    s.curr_gt = s.curr_gt * Tgt;
    s.curr_odom = s.curr_odom * Test;
    s.pm.writePose(Tgt, Test);
    s.pm.writePose(Tgt, Test, time1, time2, time_unit_denominator);
    s.pm.writeTraj(s.curr_odom, s.pm.est_traj_path_);
    s.pm.writeTraj(s.curr_gt, s.pm.gt_traj_path_);
    s.pm.writeTimeTaken(time_taken, Tgt);
    return 0;
}

}  // extern "C"
