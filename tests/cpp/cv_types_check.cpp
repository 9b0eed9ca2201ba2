// cv_types_check.cpp -- the C++ adapters instantiated with the NODE'S OWN TYPE SHAPES: Eigen::MatrixXd, cv::Mat,
// cv::KeyPoint, pcl::PointCloud<pcl::PointXYZ / PointXYZI> -- i.e. `RadarFeature` with the reference's exact members
// (radar_loader.hpp:16-33), so the node's visualisation branch (odometry.cpp:335-339,362) compiles against it.  On a
// machine with the real libraries this builds against them; in this image (no Eigen / OpenCV / PCL) the test compiles it
// against the interface stand-ins of oracle/refshim (tests/test_cpp_driver.py::test_adapters_with_cv_and_eigen_types).
// Exit codes: 0 ok (GPU present: one frame extracted, matched, solved through these types), 77 no GPU (compiled, linked,
// instantiated), 1 failure.
#include <Eigen/Dense>
#include <opencv2/core.hpp>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include <cstdio>
#include <random>

#include "nrs_adapters.hpp"

using RadarFeature = nrs::RadarFeatureT<Eigen::MatrixXd, cv::Mat, cv::KeyPoint>;
using RadarLoader = nrs::RadarLoaderT<Eigen::MatrixXd, cv::Mat, cv::KeyPoint>;
using ORORA = nrs::ORORAT<Eigen::MatrixXd, pcl::PointCloud<pcl::PointXYZ>>;
using SCManager = nrs::SCManagerT<Eigen::MatrixXd, pcl::PointCloud<pcl::PointXYZI>>;

extern "C" {
RefCvCallbacks g_ref_cv;          // only the stand-in headers know this symbol; harmless elsewhere
}

static std::vector<uint8_t> synthetic_frame(unsigned seed, float shift_bins)
{
    const int R = 400, C = 3360 + 11;
    std::vector<uint8_t> raw((size_t)R * C, 0);
    std::mt19937 rng(seed), world(99);
    std::normal_distribution<float> noise(30.f, 8.f);
    for (int i = 0; i < R; ++i) {
        uint8_t *row = raw.data() + (size_t)i * C;
        const int64_t stamp = 1560000000000000000LL + (int64_t)i * 625000;
        std::memcpy(row, &stamp, 8);
        const uint16_t enc = (uint16_t)((i * 14) % 5600);
        std::memcpy(row + 8, &enc, 2);
        row[10] = 255;
        for (int j = 0; j < 3360; ++j) row[11 + j] = (uint8_t)std::min(255.f, std::max(0.f, noise(rng)));
    }
    std::uniform_int_distribution<int> az(0, R - 1), rg(200, 3000), pw(140, 250);
    for (int k = 0; k < 500; ++k) {
        const int a = az(world), r = rg(world) + (int)shift_bins, p = pw(world);
        for (int da = 0; da < 2; ++da)
            for (int dr = -2; dr <= 2; ++dr) raw[(size_t)((a + da) % R) * C + 11 + r + dr] = (uint8_t)p;
    }
    return raw;
}

int main()
{
    nrs_ctx *ctx = nullptr;
    if (nrs_ctx_create(0, &ctx) != NRS_OK) {
        std::printf("no GPU: %s (adapters compiled, linked and instantiated with cv::Mat / cv::KeyPoint / Eigen / pcl types)\n", nrs_last_error());
        return 77;
    }
    RadarLoader loader("mulran", "cen2018", 0.6, true, false, ctx);
    const std::vector<uint8_t> r0 = synthetic_frame(1, 0.f), r1 = synthetic_frame(2, 3.f);
    RadarFeature f0 = loader.extractFeature(r0.data(), 400, 3371, 1560000000000000000LL);
    RadarFeature f1 = loader.extractFeature(r1.data(), 400, 3371, 1560000000250000000LL);
    if (f0.img_.rows != 964 || f0.img_.cols != 964 || f0.desc_.cols != 32 || f0.desc_.rows != (int)f0.kp_.size() || f0.kp_.empty()) {
        std::printf("unexpected feature shapes: img %d x %d, desc %d x %d, kp %zu\n", f0.img_.rows, f0.img_.cols, f0.desc_.rows, f0.desc_.cols, f0.kp_.size());
        return 1;
    }
    if (f0.kp_[0].size != 21.0f || f0.kp_[0].pt.x <= 0 || f0.kp_[0].pt.y <= 0) { std::printf("bad keypoint\n"); return 1; }
    auto [pm, cm] = loader.matchFeatures(f0, f1);
    if (pm.kp_.size() != (size_t)pm.feat_in_cart_.cols() || pm.feat_in_cart_.rows() != 2) { std::printf("bad match shapes\n"); return 1; }
    ORORA est("mulran", "rot", true, false, true, 700, 0.049, 0.1, 9.0 * M_PI / 180.0, ctx);
    nrs_solver_params params;
    nrs::setParams(0.75, 1.0, false, 100, 1.4, 1e-6, params);
    const Eigen::MatrixXd T = est.calcTransform(params, 0.25, pm, cm);
    if (T.rows() != 4 || T.cols() != 4 || pm.feat_compensated_.cols() != pm.feat_in_cart_.cols()) { std::printf("bad solve shapes\n"); return 1; }
    SCManager sc(ctx);
    pcl::PointCloud<pcl::PointXYZI> cloud;
    for (long j = 0; j < f0.feat_in_cart_.cols(); ++j) {
        pcl::PointXYZI p;
        p.x = (float)f0.feat_in_cart_(0, j); p.y = (float)f0.feat_in_cart_(1, j); p.z = 1.0f;
        cloud.points.push_back(p);
    }
    sc.makeAndSaveScancontextAndKeys(cloud);
    std::printf("cv/Eigen/pcl-typed adapters: %zu keypoints, %ld matches, x = %.3f m, %zu keyframe(s)\n", f0.kp_.size(), (long)pm.feat_in_cart_.cols(),
                T(0, 3), sc.polarcontexts_.size());
    return 0;
}
