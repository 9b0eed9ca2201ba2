// ref_frontend_wrap.cpp -- extern "C" doors into the REFERENCE's own front-end functions, compiled from the sources
// where they lie (outlier-robust-radar-odometry/src/features.cpp, src/radar_utils.cpp and the header-only
// include/radar_loader.hpp, unmodified) against the
// stand-in headers of oracle/refshim/ (Eigen / OpenCV / Boost / PCL are not in this image; every OpenCV algorithm call
// is forwarded to the real cv2 through g_ref_cv -- see oracle/refshim/README.md).  Recipe: oracle/Makefile `ref`;
// output oracle/_ref/libref_frontend.so.  TEST INFRASTRUCTURE ONLY: the oracle (oracle/frontend_ref.*) is checked
// against this in tests/test_ref_frontend.py; nothing in the product links or loads it.
#include <cstring>

#include "features.hpp"          // the reference's header (declares cen2018features / cen2019features, pulls radar_utils.hpp)
#include "radar_loader.hpp"      // the reference's RadarLoader (header-only): extractFeature / matchFeatures as the node calls them

extern "C" {

RefCvCallbacks g_ref_cv;

void ref_set_callbacks(const RefCvCallbacks *cb) { g_ref_cv = *cb; }

static cv::Mat mat_from(const float *p, int rows, int cols)
{
    cv::Mat m(rows, cols, CV_32F);
    std::memcpy(m.data, p, sizeof(float) * (size_t)rows * cols);
    return m;
}

static int put_targets(const Eigen::MatrixXd &t, double *out_3xcap, int cap)
{
    const int K = (int)t.cols();
    if (K > cap) return -K;
    if (K) std::memcpy(out_3xcap, t.data(), sizeof(double) * 3 * (size_t)K);
    return K;
}

// features.cpp:197-299 -> K (column-major 3 x K in out), or -K when cap is too small
int ref_cen2019features(const float *fft, int rows, int cols, int max_points, int min_range, double *out_3xcap, int cap)
{
    Eigen::MatrixXd targets;
    cen2019features(mat_from(fft, rows, cols), max_points, min_range, targets);
    return put_targets(targets, out_3xcap, cap);
}

// features.cpp:44-131
int ref_cen2018features(const float *fft, int rows, int cols, float zq, int sigma_gauss, int min_range, double *out_3xcap, int cap)
{
    Eigen::MatrixXd targets;
    cen2018features(mat_from(fft, rows, cols), zq, sigma_gauss, min_range, targets);
    return put_targets(targets, out_3xcap, cap);
}

// radar_utils.cpp:104-129 (cv::imread through the bridge): fft must hold N x range_bins floats for the version
int ref_load_radar(const char *path, int navtech_version, int cap_rows, int64_t *timestamps, double *azimuths, uint8_t *valid,
                   float *fft, int *rows, int *cols)
{
    std::vector<int64_t> t;
    std::vector<double> az;
    std::vector<bool> v;
    cv::Mat f;
    load_radar(path, t, az, v, f, navtech_version);
    *rows = f.rows; *cols = f.cols;
    if (f.rows > cap_rows) return -1;
    for (int i = 0; i < f.rows; ++i) { timestamps[i] = t[i]; azimuths[i] = az[i]; valid[i] = v[i] ? 1 : 0; }
    std::memcpy(fft, f.data, sizeof(float) * (size_t)f.rows * f.cols);
    return 0;
}

// radar_utils.cpp:227-250
double ref_get_azimuth_index(const double *azimuths, int n, double azimuth)
{
    std::vector<double> az(azimuths, azimuths + n);
    return get_azimuth_index(az, azimuth);
}

// radar_utils.cpp:252-316.  out: width x width floats (as_u8 = 0) or bytes (as_u8 = 1); *rows_after = the caller's
// fft_data.rows after the call (the function stacks the wrap-around rows into ITS ARGUMENT)
int ref_polar_to_cartesian_radar_img(const double *azimuths, int n_az, const float *fft, int rows, int cols, float radar_resolution,
                                     float cart_resolution, int cart_pixel_width, int interpolate_crossover, int navtech_version,
                                     int as_u8, void *out, int *rows_after)
{
    std::vector<double> az(azimuths, azimuths + n_az);
    cv::Mat f = mat_from(fft, rows, cols), img;
    polar_to_cartesian_radar_img(az, f, radar_resolution, cart_resolution, cart_pixel_width, interpolate_crossover != 0, img,
                                 as_u8 ? CV_8UC1 : CV_32F, navtech_version);
    *rows_after = f.rows;
    std::memcpy(out, img.data, (size_t)img.rows * img.cols * img.elemSize());
    return 0;
}

// radar_utils.cpp:329-341 (the overload extractFeature calls, radar_loader.hpp:107-109)
int ref_polar_to_cartesian_points(const double *azimuths, int n_az, const int64_t *times, const double *polar_3xN, int N,
                                  float radar_resolution, double *cart_3xN, int64_t *point_times)
{
    std::vector<double> az(azimuths, azimuths + n_az);
    std::vector<int64_t> t(times, times + n_az), pt;
    Eigen::MatrixXd polar(3, N), cart;
    if (N) std::memcpy(polar.data(), polar_3xN, sizeof(double) * 3 * (size_t)N);
    polar_to_cartesian_points(az, t, polar, radar_resolution, cart, pt);
    if (N) std::memcpy(cart_3xN, cart.data(), sizeof(double) * 3 * (size_t)N);
    for (int i = 0; i < N; ++i) point_times[i] = pt[i];
    return 0;
}

// radar_utils.cpp:382-403 (the KeyPoint overload, radar_loader.hpp:123-124): in place, returns the surviving count
int ref_convert_to_bev(double *cart_3xN, int N, float cart_resolution, int cart_pixel_width, int patch_size, float *kp_uv_Nx2,
                       int64_t *point_times)
{
    Eigen::MatrixXd cart(3, N);
    if (N) std::memcpy(cart.data(), cart_3xN, sizeof(double) * 3 * (size_t)N);
    std::vector<int64_t> pt(point_times, point_times + N);
    std::vector<cv::KeyPoint> kp;
    convert_to_bev(cart, cart_resolution, cart_pixel_width, patch_size, kp, pt);
    const int K = (int)kp.size();
    if (K) std::memcpy(cart_3xN, cart.data(), sizeof(double) * 3 * (size_t)K);
    for (int i = 0; i < K; ++i) { kp_uv_Nx2[2 * i] = kp[i].pt.x; kp_uv_Nx2[2 * i + 1] = kp[i].pt.y; point_times[i] = pt[i]; }
    return K;
}

// ---- RadarLoader (radar_loader.hpp:35-308): the chain and the matcher exactly as the node drives them ----
void *ref_loader_create(const char *dataset, const char *keypoint_extraction, double voxel_size, int use_voxelization)
{
    try {
        return new RadarLoader(dataset, keypoint_extraction, voxel_size, use_voxelization != 0, false);
    } catch (const std::exception &) {
        return nullptr;
    }
}
void ref_loader_destroy(void *l) { delete (RadarLoader *)l; }

// radar_loader.hpp:74-141 (cv::imread, cv::ORB through the bridge; pcl::VoxelGrid is the stand-in, see refshim)
void *ref_extract_feature(void *loader, const char *datadir, const char *radar_file)
{
    return new RadarFeature(((RadarLoader *)loader)->extractFeature(datadir, radar_file));
}
void ref_feature_destroy(void *f) { delete (RadarFeature *)f; }

// sizes: [0] feat_in_cart_ cols, [1] feat_in_cart_ rows, [2] feat_in_polar_ cols, [3] kp_, [4] desc_ rows, [5] t_, [6] t_norm_,
// [7] img_ rows, [8] img_ cols
void ref_feature_sizes(void *fp, int *sizes9)
{
    const RadarFeature &f = *(RadarFeature *)fp;
    sizes9[0] = (int)f.feat_in_cart_.cols(); sizes9[1] = (int)f.feat_in_cart_.rows(); sizes9[2] = (int)f.feat_in_polar_.cols();
    sizes9[3] = (int)f.kp_.size(); sizes9[4] = f.desc_.rows; sizes9[5] = (int)f.t_.size(); sizes9[6] = (int)f.t_norm_.size();
    sizes9[7] = f.img_.rows; sizes9[8] = f.img_.cols;
}

// any pointer may be NULL; cart / polar column-major
void ref_feature_get(void *fp, double *cart, double *polar, float *kp_xy, uint8_t *desc, int64_t *t, double *t_norm, uint8_t *img,
                     int64_t *t_criteria, double *time_unit_denominator)
{
    const RadarFeature &f = *(RadarFeature *)fp;
    if (cart && f.feat_in_cart_.size()) std::memcpy(cart, f.feat_in_cart_.data(), sizeof(double) * (size_t)f.feat_in_cart_.size());
    if (polar && f.feat_in_polar_.size()) std::memcpy(polar, f.feat_in_polar_.data(), sizeof(double) * (size_t)f.feat_in_polar_.size());
    if (kp_xy) for (size_t i = 0; i < f.kp_.size(); ++i) { kp_xy[2 * i] = f.kp_[i].pt.x; kp_xy[2 * i + 1] = f.kp_[i].pt.y; }
    if (desc && f.desc_.rows) std::memcpy(desc, f.desc_.data, 32 * (size_t)f.desc_.rows);
    if (t) for (size_t i = 0; i < f.t_.size(); ++i) t[i] = f.t_[i];
    if (t_norm) for (size_t i = 0; i < f.t_norm_.size(); ++i) t_norm[i] = f.t_norm_[i];
    if (img && f.img_.rows) std::memcpy(img, f.img_.data, (size_t)f.img_.rows * f.img_.cols);
    if (t_criteria) *t_criteria = f.t_crietria_;
    if (time_unit_denominator) *time_unit_denominator = f.time_unit_denominator_;
}

// radar_loader.hpp:143-205 -> two new features (prev_matched, curr_matched)
void ref_match_features(void *loader, void *prev, void *curr, void **prev_matched, void **curr_matched)
{
    auto [pm, cm] = ((RadarLoader *)loader)->matchFeatures(*(RadarFeature *)prev, *(RadarFeature *)curr);
    *prev_matched = new RadarFeature(pm);
    *curr_matched = new RadarFeature(cm);
}

}  // extern "C"
