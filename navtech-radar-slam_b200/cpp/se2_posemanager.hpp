// se2_posemanager.hpp -- writer half of the reference's SE2PoseManager
// (outlier-robust-radar-odometry/include/SE2posemanager.hpp:151-259): the five text files the orora_odom node
// appends one row per scan pair to (odometry.cpp:281-285), in the reference's formatting (ofstream << double), so
// ORORA/scripts/evaluate_odometry.py scores them unchanged:
//     <prefix>rel_error_odom.txt   ts_error,rot_error,ts_perc_error,rot_deg_per_m_error,ts_abs      (writePose :194-214)
//     <prefix>eval_odom.txt        tsx_gt,tsy_gt,yaw_gt,tsx_test,tsy_test,yaw_test,time1/den,time2/den (:216-229)
//     <prefix>time_taken.txt       time_taken[us],is_stopped                                         (:231-245)
//     <prefix>est_traj.txt / gt_traj.txt   x,y,yaw                                                   (:247-259)
// and the reader half the node uses for the ground-truth columns: loadGTPoses (:17-36, MulRan:
// <seq_dir>/gt/radar_odometry.csv keyed by time stamp) and get_groundtruth_odometry (radar_utils.cpp:483-501, Oxford:
// the row whose two stamps match), behind getXYYawPose (:137-140,159-166) -- quirks kept (see loadGTPoses).
// The Python twin is navtech-radar-slam_b200/posemanager.py; tests/test_cpp_driver.py compares the two byte for byte.
// Host-side bookkeeping only.
#ifndef NRS_SE2_POSEMANAGER_HPP
#define NRS_SE2_POSEMANAGER_HPP
#include <cmath>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <unordered_map>
#include <vector>

namespace nrs {

namespace gt_detail {
inline std::vector<std::string> split_commas(const std::string &line)        // boost::split(.., is_any_of(","))
{
    std::vector<std::string> parts;
    size_t a = 0;
    for (;;) {
        const size_t b = line.find(',', a);
        parts.push_back(line.substr(a, b == std::string::npos ? std::string::npos : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    return parts;
}
}  // namespace gt_detail

// loadGTPoses (SE2posemanager.hpp:17-36): "(Time, x, y, yaw)" rows without a header line.  Kept as written there:
// the y entry is read from column 3 like the yaw (`{stof(words[1]), stof(words[3]), stof(words[3])}`), so the
// relative ground truth the node writes and scores against has y == yaw.
inline void loadGTPoses(const std::string &gt_path, std::unordered_map<int64_t, std::vector<float>> &gt_rel)
{
    gt_rel.clear();
    std::ifstream ifs(gt_path);
    std::string line;
    while (std::getline(ifs, line)) {
        const std::vector<std::string> words = gt_detail::split_commas(line);
        if (words.size() < 4) throw std::invalid_argument("loadGTPoses: a row of " + gt_path + " has fewer than 4 columns");
        gt_rel[std::stoll(words[0])] = {std::stof(words[1]), std::stof(words[3]), std::stof(words[3])};
    }
    if (gt_rel.empty()) throw std::invalid_argument("Some parameters seem to be wrong! Load GT failed!");
}

// get_groundtruth_odometry (radar_utils.cpp:483-501): Oxford's radar_odometry.csv, first line = header; the row with
// parts[9] == t1 and parts[8] == t2 gives parts[2..7] = x, y, z, roll, pitch, yaw
inline bool get_groundtruth_odometry(const std::string &gtfile, int64_t t1, int64_t t2, std::vector<float> &gt)
{
    std::ifstream ifs(gtfile);
    std::string line;
    std::getline(ifs, line);
    gt.clear();
    while (std::getline(ifs, line)) {
        const std::vector<std::string> parts = gt_detail::split_commas(line);
        if (parts.size() < 10) continue;
        if (std::stoll(parts[9]) == t1 && std::stoll(parts[8]) == t2) {
            for (int i = 2; i < 8; ++i) gt.push_back(std::stof(parts[i]));
            return true;
        }
    }
    return false;
}

// 4x4 double, column-major like Eigen::Matrix4d (only what the node's bookkeeping needs)
struct Mat4 {
    double m[16];
    Mat4() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
    double &operator()(int i, int j) { return m[4 * j + i]; }
    double operator()(int i, int j) const { return m[4 * j + i]; }
    double *data() { return m; }
    const double *data() const { return m; }
    Mat4 operator*(const Mat4 &b) const
    {
        Mat4 r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double s = 0;
                for (int k = 0; k < 4; ++k) s += (*this)(i, k) * b(k, j);
                r(i, j) = s;
            }
        return r;
    }
    // inverse of an affine transform [A t; 0 1] (every matrix that reaches the pose manager has that bottom row)
    Mat4 inverse_affine() const
    {
        const Mat4 &a = *this;
        const double c00 = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1), c01 = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2),
                     c02 = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
        const double det = a(0, 0) * c00 + a(0, 1) * c01 + a(0, 2) * c02;
        Mat4 r;
        r(0, 0) = c00 / det;
        r(1, 0) = c01 / det;
        r(2, 0) = c02 / det;
        r(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) / det;
        r(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) / det;
        r(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) / det;
        r(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) / det;
        r(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) / det;
        r(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) / det;
        for (int i = 0; i < 3; ++i) r(i, 3) = -(r(i, 0) * a(0, 3) + r(i, 1) * a(1, 3) + r(i, 2) * a(2, 3));
        return r;
    }
};

// rot.eulerAngles(0, 1, 2)(2) of a rotation about z (Eigen/src/Geometry/EulerAngles.h with i,j,k = 0,1,2: the first
// angle atan2(R(1,2), R(2,2)) is 0, so s1 = 0, c1 = 1 and the third is -atan2(s1 R(2,0) - c1 R(1,0), c1 R(1,1) - s1 R(2,1))
// = -atan2(0 - R(1,0), R(1,1)); the subtraction is kept so that an exact R(1,0) = +0 gives -0 like there)
inline double euler_yaw(const Mat4 &T) { return -std::atan2(0.0 - T(1, 0), T(1, 1)); }

class SE2PoseManager {
public:
    static constexpr double EPSILON_ROT = 0.00000000001;       // SE2posemanager.hpp:13
    static constexpr double EPSILON_TRANS = 0.001;             // :14

    SE2PoseManager() {}
    // out_dir plays the role of "<seq_dir>/outputs" (:107-112); the files are truncated like clearOutputFile (:190-194)
    SE2PoseManager(const std::string &out_dir, const std::string &dataset_name, const std::string &algorithm_name,
                   const std::string &keypoint_extraction)
        : dataset_name_(dataset_name)
    {
        setFilenames(out_dir + "/" + dataset_name + "_" + algorithm_name + "_" + keypoint_extraction + "_");
        for (const std::string *p : {&output_pose_error_path_, &output_pose_path_, &output_time_path_, &est_traj_path_,
                                     &gt_traj_path_})
            clearOutputFile(*p);
    }

    // the node's form (:73-135 with ANALYZE_PARAM off): ground truth from <seq_dir>/gt/radar_odometry.csv (loaded up
    // front for MulRan, searched per pair for Oxford), files under <seq_dir>/outputs/, truncated (is_overwrite_available)
    static SE2PoseManager from_seq_dir(const std::string &seq_dir, const std::string &dataset_name,
                                       const std::string &algorithm_name, const std::string &keypoint_extraction)
    {
        const std::string output_dir = seq_dir + "/outputs";
        struct stat st;
        if (stat(output_dir.c_str(), &st) != 0) mkdir(output_dir.c_str(), 0777);
        SE2PoseManager pm(output_dir, dataset_name, algorithm_name, keypoint_extraction);
        pm.gt_path_ = seq_dir + "/gt/radar_odometry.csv";
        if (dataset_name == "mulran") loadGTPoses(pm.gt_path_, pm.gt_rel_);
        return pm;
    }

    std::vector<float> getXYYawPose(const int64_t t)             // MulRan (:137-142)
    {
        const auto it = gt_rel_.find(t);
        if (it == gt_rel_.end()) throw std::invalid_argument("The corresponding time is not inserted!... exiting.");
        return it->second;
    }
    std::vector<float> getXYYawPose(const int64_t time1, const int64_t time2)     // Oxford (:159-166)
    {
        std::vector<float> gtvec;
        if (!get_groundtruth_odometry(gt_path_, time1, time2, gtvec)) throw std::invalid_argument("Stop!!");
        return {gtvec[0], gtvec[1], gtvec[5]};
    }

    void setFilenames(std::string common_dir)                   // :151-157
    {
        output_pose_error_path_ = common_dir + "rel_error_odom.txt";
        output_pose_path_ = common_dir + "eval_odom.txt";
        output_time_path_ = common_dir + "time_taken.txt";
        est_traj_path_ = common_dir + "est_traj.txt";
        gt_traj_path_ = common_dir + "gt_traj.txt";
    }

    // XYYaw2SE2 (:176-188): translation, and the yaw through tf::createQuaternionFromYaw (double) narrowed to an
    // Eigen::Quaternionf whose float rotation matrix is widened again; MulRan's ground truth enters with -yaw
    void XYYaw2SE2(const std::vector<float> &xyyaw_rel, Mat4 &Tgt) const
    {
        Tgt(0, 3) = xyyaw_rel[0];
        Tgt(1, 3) = xyyaw_rel[1];
        const double yaw = dataset_name_ == "mulran" ? -(double)xyyaw_rel[2] : (double)xyyaw_rel[2];
        const float z = (float)std::sin(yaw * 0.5), w = (float)std::cos(yaw * 0.5);    // tf setRPY(0, 0, yaw): x = y = 0
        const float tz = 2.0f * z, twz = tz * w, tzz = tz * z;
        Tgt(0, 0) = 1.0f - (0.0f + tzz); Tgt(0, 1) = 0.0f - twz; Tgt(0, 2) = 0.0;
        Tgt(1, 0) = 0.0f + twz;          Tgt(1, 1) = 1.0f - (0.0f + tzz); Tgt(1, 2) = 0.0;
        Tgt(2, 0) = 0.0; Tgt(2, 1) = 0.0; Tgt(2, 2) = 1.0;
    }

    void clearOutputFile(const std::string &path) { std::ofstream f(path); }

    std::vector<double> writePose(const Mat4 &Tgt, const Mat4 &Test)             // :194-214
    {
        const Mat4 Trel = Tgt.inverse_affine() * Test;
        const double ts_error = std::sqrt(Trel(0, 3) * Trel(0, 3) + Trel(1, 3) * Trel(1, 3));
        double rot_error = std::acos((Trel(0, 0) + Trel(1, 1) + Trel(2, 2) - 1) / 2 + EPSILON_ROT) * 180.0 / M_PI;
        if (std::isnan(rot_error)) rot_error = 0.0;
        const double ts_abs = std::sqrt(Tgt(0, 3) * Tgt(0, 3) + Tgt(1, 3) * Tgt(1, 3));
        const double ts_perc_error = ts_abs < EPSILON_TRANS ? 0 : ts_error / ts_abs * 100.0;
        const double rot_deg_per_m_error = ts_abs < EPSILON_TRANS ? 0 : rot_error / ts_abs;
        std::ofstream f(output_pose_error_path_, std::ios::app);
        f << ts_error << "," << rot_error << "," << ts_perc_error << "," << rot_deg_per_m_error << "," << ts_abs << "\n";
        return {ts_error, rot_error};
    }

    void writePose(const Mat4 &Tgt, const Mat4 &Test, int64_t time1, int64_t time2, const double time_unit_denominator)
    {                                                                             // :216-229
        std::ofstream f(output_pose_path_, std::ios::app);
        f << Tgt(0, 3) << "," << Tgt(1, 3) << "," << -1 * std::asin(Tgt(0, 1)) << "," << Test(0, 3) << "," << Test(1, 3)
          << "," << -1 * std::asin(Test(0, 1)) << "," << time1 / time_unit_denominator << ","
          << time2 / time_unit_denominator << "\n";
    }

    void writeTimeTaken(int time_taken, const Mat4 &Tgt)                          // :231-245
    {
        const double rel_ts = std::sqrt(std::pow(Tgt(0, 3), 2) + std::pow(Tgt(1, 3), 2));
        const double rel_rot = euler_yaw(Tgt);
        const bool is_stopped = rel_ts < 0.003 && rel_rot < 0.001;
        std::ofstream f(output_time_path_, std::ios::app);
        f << time_taken << "," << is_stopped << "\n";
    }

    void writeTraj(const Mat4 &Tf4x4, const std::string &path)                    // :247-259
    {
        std::ofstream f(path, std::ios::app);
        f << Tf4x4(0, 3) << "," << Tf4x4(1, 3) << "," << euler_yaw(Tf4x4) << "\n";
    }

    std::string dataset_name_, gt_path_;
    std::unordered_map<int64_t, std::vector<float>> gt_rel_;
    std::string output_pose_error_path_, output_pose_path_, output_time_path_, est_traj_path_, gt_traj_path_;
};

}  // namespace nrs
#endif
