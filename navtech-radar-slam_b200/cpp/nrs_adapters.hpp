// nrs_adapters.hpp -- header-only C++ host side above the C ABI (include/nrs_b200.h).
//
// The reference is compiled C++; these adapters keep the class and method names, argument order
// and error behaviour of the reference's operator API so its two ROS nodes can call the GPU path
// without source changes beyond the include:
//
//   nrs::SCManagerT<MatrixXd, Cloud>   <-> pgo/SC-A-LOAM/include/scancontext/Scancontext.h:57-122
//   nrs::cen2019features               <-> outlier-robust-radar-odometry/include/features.hpp:36
//   nrs::cen2018features               <-> outlier-robust-radar-odometry/include/features.hpp:27
//   nrs::RadarLoaderT<MatrixXd>        <-> include/radar_loader.hpp:35-308 (extractFeature, matchFeatures)
//   nrs::ORORAT<MatrixXd>              <-> include/ORORA.hpp:13-181 (calcTransform)
//   nrs::setParams                     <-> include/solver/nonminimal_solver.hpp:959-973
//
// Eigen / PCL / OpenCV are not available in the build container, so the adapters are templates over
// the matrix and cloud types: anything with Eigen::MatrixXd's interface (column-major data(),
// rows(), cols(), resize(), operator()(i,j)) and pcl::PointCloud's (points[i].x/y/z) fits.  With
// the real libraries:  using SCManager = nrs::SCManagerT<Eigen::MatrixXd, pcl::PointCloud<pcl::PointXYZI>>;
// navtech-radar-slam_b200/cpp/adapter_check.cpp instantiates them with minimal stand-ins.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "nrs_b200.h"

namespace nrs {

inline void check(int rc)
{
    if (rc == NRS_EINVAL) throw std::invalid_argument(nrs_last_error());
    if (rc < 0) throw std::runtime_error(std::string("nrs_b200: ") + nrs_last_error());
}

// one context per process (the reference is single-GPU, single-process per node)
inline nrs_ctx *default_ctx(int device = 0)
{
    static nrs_ctx *ctx = nullptr;
    if (!ctx) check(nrs_ctx_create(device, &ctx));
    return ctx;
}

// ================================================================================================
// SCManager (Scancontext.h:57-122)
// ================================================================================================
template <class MatrixXd, class Cloud>
class SCManagerT {
public:
    using KeyMat = std::vector<std::vector<float>>;

    explicit SCManagerT(nrs_ctx *ctx = nullptr) : ctx_(ctx ? ctx : default_ctx()) { check(nrs_sc_db_create(ctx_, 1024, &db_)); }
    ~SCManagerT() { nrs_sc_db_destroy(db_); }
    SCManagerT(const SCManagerT &) = delete;
    SCManagerT &operator=(const SCManagerT &) = delete;

    MatrixXd makeScancontext(Cloud &_scan_down)
    {
        std::vector<float> xyz = flatten(_scan_down);
        MatrixXd desc(PC_NUM_RING, PC_NUM_SECTOR);
        check(nrs_sc_make(ctx_, xyz.data(), (int)(xyz.size() / 3), LIDAR_HEIGHT, PC_MAX_RADIUS, desc.data()));
        return desc;
    }
    MatrixXd makeRingkeyFromScancontext(MatrixXd &_desc)
    {
        MatrixXd key(_desc.rows(), 1);
        check(nrs_sc_keys(ctx_, _desc.data(), key.data(), nullptr));
        return key;
    }
    MatrixXd makeSectorkeyFromScancontext(MatrixXd &_desc)
    {
        MatrixXd key(1, _desc.cols());
        check(nrs_sc_keys(ctx_, _desc.data(), nullptr, key.data()));
        return key;
    }
    int fastAlignUsingVkey(MatrixXd &_vkey1, MatrixXd &_vkey2)
    {
        int32_t s = 0;
        check(nrs_sc_fast_align(ctx_, _vkey1.data(), _vkey2.data(), &s));
        return s;
    }
    double distDirectSC(MatrixXd &_sc1, MatrixXd &_sc2)
    {
        double d = 0;
        check(nrs_sc_dist_direct(ctx_, _sc1.data(), _sc2.data(), &d));
        return d;
    }
    std::pair<double, int> distanceBtnScanContext(MatrixXd &_sc1, MatrixXd &_sc2)
    {
        double d = 0; int32_t s = 0;
        check(nrs_sc_pair(ctx_, _sc1.data(), _sc2.data(), &d, &s));
        return std::make_pair(d, (int)s);
    }

    // User-side API
    void makeAndSaveScancontextAndKeys(Cloud &_scan_down)
    {
        std::vector<float> xyz = flatten(_scan_down);
        MatrixXd sc(PC_NUM_RING, PC_NUM_SECTOR);
        check(nrs_sc_db_append_cloud(db_, xyz.data(), (int)(xyz.size() / 3), LIDAR_HEIGHT, PC_MAX_RADIUS, sc.data()));
        save_host(sc);
    }
    void saveScancontextAndKeys(MatrixXd _scd)
    {
        check(nrs_sc_db_append(db_, _scd.data(), 1));
        save_host(_scd);
    }
    // Scancontext.cpp:331-422; exhaustive over the set the reference's tree would hold
    std::pair<int, float> detectLoopClosureID(void)
    {
        int loop_id = -1;
        if ((int)polarcontext_invkeys_mat_.size() < NUM_EXCLUDE_RECENT + 1) return {loop_id, 0.0f};
        if (tree_making_period_conter % TREE_MAKING_PERIOD_ == 0)
            searchable_end_ = (int)polarcontext_invkeys_mat_.size() - NUM_EXCLUDE_RECENT;
        tree_making_period_conter = tree_making_period_conter + 1;
        nrs_sc_result r;
        if (use_tree_candidates)      // the reference's own flow: ring-key k-NN, then its 3 candidates (:361-396)
            check(nrs_sc_detect_knn(db_, polarcontexts_.back().data(), 0, searchable_end_, NUM_CANDIDATES_FROM_TREE,
                                    nullptr, &r));
        else                          // exhaustive over the set the tree would hold (north_star)
            check(nrs_sc_query(db_, polarcontexts_.back().data(), 0, searchable_end_, nullptr, 0, &r));
        if (r.dist < SC_DIST_THRES) loop_id = r.idx;
        return {loop_id, yaw(r.shift)};
    }
    std::pair<int, float> detectLoopClosureIDBetweenSession(std::vector<float> &curr_key, MatrixXd &curr_desc)
    {
        is_tree_batch_made = true;
        nrs_sc_result r;
        const int n = (int)polarcontexts_.size();
        if (use_tree_candidates && (int)curr_key.size() == PC_NUM_RING) {
            int32_t cand[16];
            check(nrs_sc_knn(db_, curr_key.data(), 0, n, NUM_CANDIDATES_FROM_TREE, cand, nullptr));
            int m = 0;
            while (m < NUM_CANDIDATES_FROM_TREE && cand[m] >= 0) ++m;
            check(nrs_sc_query(db_, curr_desc.data(), 0, 0, cand, m, &r));
        } else
            check(nrs_sc_query(db_, curr_desc.data(), 0, n, nullptr, 0, &r));
        return {r.dist < SC_DIST_THRES ? r.idx : -1, yaw(r.shift)};
    }
    const MatrixXd &getConstRefRecentSCD(void) { return polarcontexts_.back(); }
    void setSCdistThres(double _new_thres) { SC_DIST_THRES = _new_thres; }

    // hyper parameters (Scancontext.h:83-104)
    const double LIDAR_HEIGHT = 2.0;
    const int PC_NUM_RING = 20;
    const int PC_NUM_SECTOR = 60;
    const double PC_MAX_RADIUS = 80.0;
    const double PC_UNIT_SECTORANGLE = 360.0 / double(PC_NUM_SECTOR);
    const double PC_UNIT_RINGGAP = PC_MAX_RADIUS / double(PC_NUM_RING);
    const int NUM_EXCLUDE_RECENT = 30;
    const int NUM_CANDIDATES_FROM_TREE = 3;
    const double SEARCH_RATIO = 0.1;
    double SC_DIST_THRES = 0.2;
    const int TREE_MAKING_PERIOD_ = 30;
    int tree_making_period_conter = 0;

    // data (Scancontext.h:110-116): kept populated on the host because they are public
    std::vector<double> polarcontexts_timestamp_;
    std::vector<MatrixXd> polarcontexts_;
    std::vector<MatrixXd> polarcontext_invkeys_;
    std::vector<MatrixXd> polarcontext_vkeys_;
    KeyMat polarcontext_invkeys_mat_;
    bool is_tree_batch_made = false;
    // false (default): exhaustive search; true: ring-key nearest neighbours + 3 candidates like the reference
    bool use_tree_candidates = false;

private:
    static std::vector<float> flatten(Cloud &c)
    {
        std::vector<float> xyz(3 * c.points.size());
        for (size_t i = 0; i < c.points.size(); ++i) {
            xyz[3 * i] = c.points[i].x; xyz[3 * i + 1] = c.points[i].y; xyz[3 * i + 2] = c.points[i].z;
        }
        return xyz;
    }
    void save_host(MatrixXd &sc)
    {
        MatrixXd ringkey = makeRingkeyFromScancontext(sc);
        MatrixXd sectorkey = makeSectorkeyFromScancontext(sc);
        std::vector<float> vec(ringkey.data(), ringkey.data() + ringkey.rows());      // eig2stdvec :62-66
        polarcontexts_.push_back(sc);
        polarcontext_invkeys_.push_back(ringkey);
        polarcontext_vkeys_.push_back(sectorkey);
        polarcontext_invkeys_mat_.push_back(vec);
    }
    float yaw(int nn_align) const
    {
        float degrees = (float)(nn_align * PC_UNIT_SECTORANGLE);                     // deg2rad(float) :17-20
        return (float)(degrees * M_PI / 180.0);
    }
    nrs_ctx *ctx_;
    nrs_sc_db *db_ = nullptr;
    int searchable_end_ = 0;
};

// ================================================================================================
// cen2019features (features.hpp:36): fft_data = R x C float32 row-major (cv::Mat CV_32F data)
// ================================================================================================
template <class MatrixXd>
double cen2019features(const float *fft_data, int rows, int cols, int max_points, int min_range, MatrixXd &targets,
                       nrs_ctx *ctx = nullptr)
{
    std::vector<double> buf(3 * 16384);
    int K = 0;
    check(nrs_cen2019(ctx ? ctx : default_ctx(), fft_data, rows, cols, max_points, min_range, buf.data(), 16384, &K));
    targets.resize(3, K);
    std::memcpy(targets.data(), buf.data(), sizeof(double) * 3 * (size_t)K);
    return 0.0;
}

// cen2018features (features.hpp:27): same conventions
template <class MatrixXd>
double cen2018features(const float *fft_data, int rows, int cols, float zq, int sigma_gauss, int min_range,
                       MatrixXd &targets, nrs_ctx *ctx = nullptr)
{
    std::vector<double> buf(3 * 16384);
    int K = 0;
    check(nrs_cen2018(ctx ? ctx : default_ctx(), fft_data, rows, cols, zq, sigma_gauss, min_range, buf.data(), 16384, &K,
                      nullptr));
    targets.resize(3, K);
    std::memcpy(targets.data(), buf.data(), sizeof(double) * 3 * (size_t)K);
    return 0.0;
}

// ================================================================================================
// RadarFeature / RadarLoader (radar_loader.hpp:16-33, 35-308)
// ================================================================================================
template <class MatrixXd>
struct RadarFeatureT {
    std::vector<uint8_t> img_;                 // 964 x 964 (cv::Mat CV_8UC1 data)
    std::vector<int64_t> t_;
    double time_unit_denominator_ = 1.0;
    std::vector<uint8_t> desc_;                // K_desc x 32
    std::vector<float> kp_;                    // K_desc x 2 (u, v)
    MatrixXd feat_in_cart_;
    MatrixXd feat_in_polar_;
    MatrixXd feat_compensated_;
    int64_t t_crietria_ = 0;
    std::vector<double> t_norm_;
    std::shared_ptr<nrs_frame> frame_;         // device-resident copy used by matchFeatures
};

template <class MatrixXd>
class RadarLoaderT {
public:
    using RadarFeature = RadarFeatureT<MatrixXd>;
    using FeatureMatching = std::tuple<RadarFeature, RadarFeature>;

    RadarLoaderT(const std::string &dataset, const std::string keypoint_extraction, const double voxel_size,
                 const bool use_voxelization, const bool /*viz_extraction*/, nrs_ctx *ctx = nullptr)
        : ctx_(ctx ? ctx : default_ctx())
    {
        check(nrs_extract_default_params(dataset.c_str(), &p_));                   // throws invalid_argument like :62
        if (keypoint_extraction != "cen2018" && keypoint_extraction != "cen2019")
            throw std::invalid_argument("Keypoint extraction method seems to be wrong!");   // :64-66
        p_.keypoint_extraction = keypoint_extraction == "cen2018" ? 1 : 0;         // :95-99
        p_.voxel_size = voxel_size;
        p_.use_voxelization = use_voxelization ? 1 : 0;
        time_unit_denominator_ = dataset == "oxford" ? 1000000.0 : 1000000000.0;
    }

    // the decoded 8-bit PNG rows (what cv::imread(path, IMREAD_GRAYSCALE) returns: data, rows, step)
    RadarFeature extractFeature(const uint8_t *png_rows, int rows, int step, int64_t t_criteria)
    {
        nrs_frame *f = nullptr;
        check(nrs_frame_extract(ctx_, png_rows, rows, step, &p_, t_criteria, &f));
        RadarFeature r;
        r.frame_ = std::shared_ptr<nrs_frame>(f, [](nrs_frame *x) { nrs_frame_destroy(x); });
        nrs_frame_info info;
        check(nrs_frame_info_get(f, &info));
        r.feat_in_polar_.resize(3, info.K_polar);
        r.feat_in_cart_.resize(3, info.K);
        r.kp_.resize(2 * (size_t)info.K_desc);
        r.t_.resize(info.K);
        r.t_norm_.resize(info.K);
        r.desc_.resize(32 * (size_t)info.K_desc);
        r.img_.resize((size_t)p_.cart_pixel_width * p_.cart_pixel_width);
        check(nrs_frame_get(f, r.feat_in_polar_.data(), r.feat_in_cart_.data(), r.kp_.data(), r.t_.data(),
                            r.t_norm_.data(), r.desc_.data(), r.img_.data()));
        r.t_crietria_ = t_criteria;
        r.time_unit_denominator_ = time_unit_denominator_;
        return r;
    }

    FeatureMatching matchFeatures(const RadarFeature &prev_feat, const RadarFeature &curr_feat)
    {
        RadarFeature prev_matched = prev_feat, curr_matched = curr_feat;           // :144-146
        const int cap = (int)(prev_feat.kp_.size() / 2) + 1;
        std::vector<int32_t> q(cap), t(cap);
        std::vector<double> pxy(2 * (size_t)cap), cxy(2 * (size_t)cap);
        std::vector<int64_t> pt(cap), ct(cap);
        int n = 0;
        check(nrs_match(ctx_, prev_feat.frame_.get(), curr_feat.frame_.get(), p_.nndr, cap, q.data(), t.data(),
                        pxy.data(), cxy.data(), pt.data(), ct.data(), &n));
        prev_matched.feat_in_cart_.resize(2, n);
        curr_matched.feat_in_cart_.resize(2, n);
        std::memcpy(prev_matched.feat_in_cart_.data(), pxy.data(), sizeof(double) * 2 * (size_t)n);
        std::memcpy(curr_matched.feat_in_cart_.data(), cxy.data(), sizeof(double) * 2 * (size_t)n);
        prev_matched.t_.assign(pt.begin(), pt.begin() + n);
        curr_matched.t_.assign(ct.begin(), ct.begin() + n);
        prev_matched.t_norm_.resize(n);
        curr_matched.t_norm_.resize(n);
        for (int j = 0; j < n; ++j) {
            prev_matched.t_norm_[j] = prev_feat.t_norm_[q[j]];
            curr_matched.t_norm_[j] = curr_feat.t_norm_[t[j]];
        }
        return {prev_matched, curr_matched};
    }

    double getTimeDenominator() { return time_unit_denominator_; }

private:
    nrs_ctx *ctx_;
    nrs_extract_params p_;
    double time_unit_denominator_ = 1.0;
};

// ================================================================================================
// ORORA (ORORA.hpp:13-181) and setParams (nonminimal_solver.hpp:959-973)
// ================================================================================================
inline void setParams(double noise_bound_of_each_measurement, double cbar2, double /*estimating_scale*/,
                      int num_max_iter, double control_parameter_for_gnc, double rot_cost_thr,
                      nrs_solver_params &params)
{
    nrs_solver_default_params(&params);
    params.noise_bound = noise_bound_of_each_measurement;
    params.cbar2 = cbar2;
    params.rotation_max_iterations = num_max_iter;
    params.rotation_gnc_factor = control_parameter_for_gnc;
    params.rotation_cost_threshold = rot_cost_thr;
}

template <class MatrixXd>
class ORORAT {
public:
    using RadarFeature = RadarFeatureT<MatrixXd>;

    ORORAT(const std::string dataset, const std::string deskewing_target, const bool use_doppler_compenstation,
           const bool use_deskewing, const bool check_stop_motion, const int num_feat_thr_for_stop_motion,
           const double beta, const double noise_bound_radial, const double noise_bound_tangential,
           nrs_ctx *ctx = nullptr)
        : ctx_(ctx ? ctx : default_ctx()), dataset_(dataset), deskewing_target_(deskewing_target),
          use_doppler_(use_doppler_compenstation), check_stop_motion_(check_stop_motion),
          stop_thr_(num_feat_thr_for_stop_motion), beta_(beta), nbr_(noise_bound_radial), nbt_(noise_bound_tangential)
    {
        if (use_deskewing) throw std::invalid_argument("use_deskewing is not supported by the GPU path (off in orora_params.yaml)");
    }

    // returns the 4x4 transform, column-major like Eigen (ORORA.hpp:37-93)
    MatrixXd calcTransform(const nrs_solver_params params, const double delta_t, RadarFeature &prev_matched,
                           RadarFeature &curr_matched)
    {
        nrs_solver_params p = params;
        p.noise_bound_radial = nbr_; p.noise_bound_tangential = nbt_;
        p.use_doppler_compensation = use_doppler_ ? 1 : 0; p.doppler_beta = beta_;
        p.check_stop_motion = check_stop_motion_ ? 1 : 0; p.num_feat_thr_for_stop_motion = stop_thr_;
        const int N = (int)prev_matched.feat_in_cart_.cols();
        std::vector<double> pv(2 * (size_t)N), cv(2 * (size_t)N);
        for (int j = 0; j < N; ++j) {
            pv[2 * j] = prev_matched.feat_in_cart_(0, j); pv[2 * j + 1] = prev_matched.feat_in_cart_(1, j);
            cv[2 * j] = curr_matched.feat_in_cart_(0, j); cv[2 * j + 1] = curr_matched.feat_in_cart_(1, j);
        }
        clique_.assign(N > 0 ? N : 1, 0);
        final_.assign(N > 0 ? N : 1, 0);
        double T[16];
        check(nrs_orora_solve(ctx_, pv.data(), cv.data(), N, &p, curr_vel_x_, T, clique_.data(), final_.data(), &info_));
        curr_vel_x_ = T[12] / delta_t;                                              // :89 + association.cpp:215
        MatrixXd out(4, 4);
        std::memcpy(out.data(), T, sizeof(T));
        return out;
    }
    int getNumMaxCliqueInliers() const { return info_.M; }
    int getNumRotaionInliers() const { return info_.n_rot_inliers; }
    std::vector<int> getFinalInliersIndices() const { return std::vector<int>(final_.begin(), final_.begin() + info_.n_final_inliers); }
    std::vector<int> getMaxCliqueIndices() const { return std::vector<int>(clique_.begin(), clique_.begin() + info_.M); }
    const nrs_solver_info &info() const { return info_; }

private:
    nrs_ctx *ctx_;
    std::string dataset_, deskewing_target_;
    bool use_doppler_, check_stop_motion_;
    int stop_thr_;
    double beta_, nbr_, nbt_;
    double curr_vel_x_ = 0.0;
    std::vector<int32_t> clique_, final_;
    nrs_solver_info info_{};
};

}  // namespace nrs
