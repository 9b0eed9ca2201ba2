// orora_odom.cpp -- ROS-free counterpart of the reference's orora_odom node
// (outlier-robust-radar-odometry/src/odometry.cpp:139-393): the node's own loop body
//     curr_feat = Loader.extractFeature(datadir, file);  [prev, curr] = Loader.matchFeatures(prev_feat, curr_feat);
//     Test = orora_estimator.calcTransform(params, delta_t, prev_matched, curr_matched);
//     curr_odom = curr_odom * Test;  PoseManager.writePose / writeTraj / writeTimeTaken
// written against the C++ adapters of nrs_adapters.hpp (RadarLoaderT / ORORAT / SCManagerT over the C ABI), with
// stand-ins for Eigen::MatrixXd and pcl::PointCloud (neither library is in this image).  What ROS gave the node
// (parameters, rviz topics) becomes command-line flags with orora_params.yaml's values as defaults; the PNG is decoded
// on the host by png_gray.hpp where the node calls cv::imread.  With --loops every frame's feature cloud also
// becomes a ScanContext keyframe and is searched like SC-A-LOAM's process_lcd does
// (pgo/SC-A-LOAM/src/laserPosegraphOptimization.cpp:486-495,553-577).
//
//   orora_odom <dir of polar_oxford_form PNGs> [--dataset mulran|oxford] [--extractor cen2018|cen2019] [--out DIR]
//              [--gt FILE] [--voxel-size 0.6] [--no-voxelization] [--frame-rate 4] [--end-idx N] [--loops]
//              [--sc-dist-thres 0.2] [--noise-bound 0.75] [--noise-bound-radial 0.1] [--noise-bound-tangential 9]
//              [--no-stop-motion] [--no-doppler] [--device 0]
//   orora_odom --seq-dir SEQ ...        the node's own layout (run_orora.launch's seq_dir): frames from
//              SEQ/polar_oxford_form/, ground truth from SEQ/gt/radar_odometry.csv, files into SEQ/outputs/
//   orora_odom --decode-only in.png out.raw            (no GPU: rows, cols as two int32, then the bytes)
//   orora_odom --gt-lookup SEQ DATASET T1 T2           (no GPU: prints getXYYawPose's x y yaw with %.9g)
//   orora_odom --posefile-selftest in.bin OUTDIR DATASET EXTRACTION   (no GPU: see tests/test_cpp_driver.py)
//
// --gt FILE: one "x,y,yaw" line per scan pair, the relative ground-truth motion getXYYawPose would return
// (odometry.cpp:253-259); without it the ground truth is Identity.  Exit codes: 0 ok, 77 no GPU, 1 error.
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "nrs_adapters.hpp"
#include "png_gray.hpp"
#include "se2_posemanager.hpp"

namespace {

struct MiniMatrix {                      // the slice of Eigen::MatrixXd's interface the adapters use
    MiniMatrix() = default;
    MiniMatrix(long r, long c) : r_(r), c_(c), d_((size_t)r * c, 0.0) {}
    double *data() { return d_.data(); }
    const double *data() const { return d_.data(); }
    long rows() const { return r_; }
    long cols() const { return c_; }
    void resize(long r, long c) { r_ = r; c_ = c; d_.assign((size_t)r * c, 0.0); }
    double &operator()(long i, long j) { return d_[(size_t)j * r_ + i]; }      // column-major
    double operator()(long i, long j) const { return d_[(size_t)j * r_ + i]; }
    long r_ = 0, c_ = 0;
    std::vector<double> d_;
};
struct MiniPoint { float x, y, z, intensity; };
struct MiniCloud { std::vector<MiniPoint> points; };

using RadarLoader = nrs::RadarLoaderT<MiniMatrix>;
using RadarFeature = RadarLoader::RadarFeature;
using ORORA = nrs::ORORAT<MiniMatrix>;
using SCManager = nrs::SCManagerT<MiniMatrix, MiniCloud>;

// getTimestamp / less_than_img (radar_utils.cpp:70-83): the integer before the first '.'
int64_t stamp_of(const std::string &name) { return std::stoll(name.substr(0, name.find('.'))); }

// get_file_names (radar_utils.cpp:85-102): regular files whose last '.'-part is the extension, ascending stamps
std::vector<std::string> get_file_names(const std::string &path, const std::string &extension = "png")
{
    std::vector<std::string> files;
    DIR *dirp = opendir(path.c_str());
    if (!dirp) throw std::invalid_argument("cannot open directory " + path);
    while (struct dirent *dp = readdir(dirp)) {
        const std::string name = dp->d_name;
        struct stat st;
        if (stat((path + "/" + name).c_str(), &st) != 0 || !S_ISREG(st.st_mode)) continue;
        const size_t dot = name.rfind('.');
        if (dot == std::string::npos || name.substr(dot + 1) != extension) continue;
        files.push_back(name);
    }
    closedir(dirp);
    std::sort(files.begin(), files.end(), [](const std::string &a, const std::string &b) { return stamp_of(a) < stamp_of(b); });
    return files;
}

nrs::Mat4 to_mat4(const MiniMatrix &T)
{
    nrs::Mat4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r(i, j) = T(i, j);
    return r;
}

int decode_only(const char *in, const char *out)
{
    const nrs::GrayImage im = nrs::read_png_gray(in);
    std::FILE *f = std::fopen(out, "wb");
    if (!f) { std::fprintf(stderr, "cannot write %s\n", out); return 1; }
    const int32_t hdr[2] = {im.rows, im.cols};
    std::fwrite(hdr, sizeof(hdr), 1, f);
    std::fwrite(im.data.data(), 1, im.data.size(), f);
    std::fclose(f);
    return 0;
}

// in.bin: int32 n, then per pair: 3 float (x, y, yaw of the ground truth), 16 double (Test, column-major), 2 int64
// (time1, time2), double denominator, int32 time_taken.  Runs the node's five writes per pair.
int posefile_selftest(const char *in, const std::string &out_dir, const std::string &dataset, const std::string &extraction)
{
    std::ifstream f(in, std::ios::binary);
    int32_t n = 0;
    f.read((char *)&n, 4);
    nrs::SE2PoseManager pm(out_dir, dataset, "ORORA", extraction);
    nrs::Mat4 curr_gt, curr_odom;
    for (int k = 0; k < n; ++k) {
        float g[3]; nrs::Mat4 Test; int64_t t[2]; double den; int32_t us;
        f.read((char *)g, 12); f.read((char *)Test.data(), 128); f.read((char *)t, 16); f.read((char *)&den, 8); f.read((char *)&us, 4);
        if (!f) { std::fprintf(stderr, "short read in %s\n", in); return 1; }
        nrs::Mat4 Tgt;
        pm.XYYaw2SE2({g[0], g[1], g[2]}, Tgt);
        curr_gt = curr_gt * Tgt;
        curr_odom = curr_odom * Test;
        pm.writePose(Tgt, Test);
        pm.writePose(Tgt, Test, t[0], t[1], den);
        pm.writeTraj(curr_odom, pm.est_traj_path_);
        pm.writeTraj(curr_gt, pm.gt_traj_path_);
        pm.writeTimeTaken(us, Tgt);
    }
    return 0;
}

struct Options {
    std::string datadir, seq_dir, dataset = "mulran", extractor = "cen2018", out_dir, gt_path;
    double voxel_size = 0.6, frame_rate = 4.0, sc_dist_thres = 0.2;
    double noise_bound = 0.75, noise_bound_radial = 0.1, noise_bound_tangential_deg = 9.0;
    bool use_voxelization = true, loops = false, check_stop_motion = true, use_doppler = true;
    int end_idx = -1, device = 0;
};

bool parse(int argc, char **argv, Options &o)
{
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&]() -> const char * { if (i + 1 >= argc) throw std::invalid_argument("missing value after " + a); return argv[++i]; };
        if (a == "--dataset") o.dataset = val();
        else if (a == "--extractor") o.extractor = val();
        else if (a == "--out") o.out_dir = val();
        else if (a == "--gt") o.gt_path = val();
        else if (a == "--seq-dir") o.seq_dir = val();
        else if (a == "--voxel-size") o.voxel_size = std::atof(val());
        else if (a == "--no-voxelization") o.use_voxelization = false;
        else if (a == "--frame-rate") o.frame_rate = std::atof(val());
        else if (a == "--end-idx") o.end_idx = std::atoi(val());
        else if (a == "--loops") o.loops = true;
        else if (a == "--sc-dist-thres") o.sc_dist_thres = std::atof(val());
        else if (a == "--noise-bound") o.noise_bound = std::atof(val());
        else if (a == "--noise-bound-radial") o.noise_bound_radial = std::atof(val());
        else if (a == "--noise-bound-tangential") o.noise_bound_tangential_deg = std::atof(val());
        else if (a == "--no-stop-motion") o.check_stop_motion = false;
        else if (a == "--no-doppler") o.use_doppler = false;
        else if (a == "--device") o.device = std::atoi(val());
        else if (!a.empty() && a[0] != '-' && o.datadir.empty()) o.datadir = a;
        else throw std::invalid_argument("unknown argument " + a);
    }
    if (!o.seq_dir.empty()) {
        if (!o.datadir.empty() || !o.out_dir.empty() || !o.gt_path.empty())
            throw std::invalid_argument("--seq-dir fixes the frame directory, the ground truth and the output directory");
        o.datadir = o.seq_dir + "/polar_oxford_form/";                               // odometry.cpp:141
    }
    return !o.datadir.empty();
}

std::vector<std::vector<float>> read_gt(const std::string &path)
{
    std::vector<std::vector<float>> rows;
    std::ifstream f(path);
    if (!f) throw std::invalid_argument("cannot open " + path);
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        std::replace(line.begin(), line.end(), ',', ' ');
        std::istringstream s(line);
        float x, y, yaw;
        if (s >> x >> y >> yaw) rows.push_back({x, y, yaw});
    }
    return rows;
}

int run(const Options &o)
{
    nrs_ctx *ctx = nullptr;
    if (nrs_ctx_create(o.device, &ctx) != NRS_OK) {
        std::printf("no GPU: %s\n", nrs_last_error());
        return 77;
    }
    // odometry.cpp:139-151 (values: config/orora_params.yaml)
    nrs_solver_params params;
    nrs::setParams(o.noise_bound, 1.0, false, 100, 1.4, 1e-6, params);
    ORORA orora_estimator(o.dataset, "rot", o.use_doppler, false, o.check_stop_motion, 700, 0.049, o.noise_bound_radial,
                          o.noise_bound_tangential_deg * (M_PI / 180.0), ctx);     // DEG2RAD
    std::vector<std::string> radar_files = get_file_names(o.datadir);
    const bool from_seq = !o.seq_dir.empty();
    const bool write = from_seq || !o.out_dir.empty();
    const std::string out_dir = from_seq ? o.seq_dir + "/outputs" : o.out_dir;
    if (write && !from_seq) mkdir(o.out_dir.c_str(), 0777);
    nrs::SE2PoseManager PoseManager;
    if (from_seq) PoseManager = nrs::SE2PoseManager::from_seq_dir(o.seq_dir, o.dataset, "ORORA", o.extractor);   // odometry.cpp:143-144
    else if (write) PoseManager = nrs::SE2PoseManager(o.out_dir, o.dataset, "ORORA", o.extractor);
    RadarLoader Loader(o.dataset, o.extractor, o.voxel_size, o.use_voxelization, false, ctx);
    const double time_unit_denominator = Loader.getTimeDenominator();
    std::vector<std::vector<float>> gt;
    if (!o.gt_path.empty()) gt = read_gt(o.gt_path);
    std::unique_ptr<SCManager> sc;
    if (o.loops) { sc.reset(new SCManager(ctx)); sc->setSCdistThres(o.sc_dist_thres); }
    std::ofstream loops_file;
    if (write && o.loops) loops_file.open(out_dir + "/loops.txt");

    RadarFeature prev_feat, curr_feat;
    nrs::Mat4 curr_gt, curr_odom;
    // the node stops one file short of the list: pair i needs the stamp of file i + 1 (odometry.cpp:157-158,184-185)
    int end_idx = o.end_idx < 0 ? (int)radar_files.size() - 1 : std::min(o.end_idx, (int)radar_files.size() - 1);
    int n_pairs = 0, n_loops = 0, n_skipped = 0;
    double solve_s = 0.0;
    const auto wall0 = std::chrono::steady_clock::now();
    for (int i = 0; i < end_idx; ++i) {
        if (i > 0) prev_feat = curr_feat;
        const nrs::GrayImage png = nrs::read_png_gray(o.datadir + "/" + radar_files[i]);       // cv::imread, radar_utils.cpp:107
        curr_feat = Loader.extractFeature(png.data.data(), png.rows, png.cols, stamp_of(radar_files[i]));
        if (sc) {
            // the cloud the node publishes: x, y of feat_in_cart_ with z = 1.0 (odometry.cpp:305-316)
            MiniCloud cloud;
            for (long j = 0; j < curr_feat.feat_in_cart_.cols(); ++j)
                cloud.points.push_back({(float)curr_feat.feat_in_cart_(0, j), (float)curr_feat.feat_in_cart_(1, j), 1.0f, 0.f});
            sc->makeAndSaveScancontextAndKeys(cloud);
            const std::pair<int, float> loop = sc->detectLoopClosureID();
            if (loop.first >= 0) {
                ++n_loops;
                if (loops_file.is_open()) {
                    char line[96];
                    std::snprintf(line, sizeof(line), "%d %d %.9g\n", i, loop.first, (double)loop.second);
                    loops_file << line;
                }
            }
        }
        if (i == 0) continue;
        auto [prev_matched, curr_matched] = Loader.matchFeatures(prev_feat, curr_feat);
        const int64_t time1 = stamp_of(radar_files[i]), time2 = stamp_of(radar_files[i + 1]);
        const double delta_t = 1.0 / o.frame_rate;
        nrs::Mat4 Test, Tgt;
        const auto start = std::chrono::steady_clock::now();
        try {
            Test = to_mat4(orora_estimator.calcTransform(params, delta_t, prev_matched, curr_matched));
        } catch (const nrs::CapacityError &e) {
            // more correspondences than the solver holds (1024; only reachable with --no-stop-motion): reported and
            // skipped as an Identity motion, like the reference's stop-motion gate
            std::fprintf(stderr, "pair %d skipped: %s\n", i, e.what());
            ++n_skipped;
        }
        const auto end = std::chrono::steady_clock::now();
        const int time_taken = (int)std::chrono::duration_cast<std::chrono::microseconds>(end - start).count();
        solve_s += time_taken * 1e-6;
        PoseManager.dataset_name_ = o.dataset;
        if (from_seq) {                                                            // odometry.cpp:253-259
            const std::vector<float> xyyaw_rel = o.dataset == "oxford" ? PoseManager.getXYYawPose(time1, time2)
                                                                       : PoseManager.getXYYawPose(time1);
            PoseManager.XYYaw2SE2(xyyaw_rel, Tgt);
        } else if (!gt.empty()) {
            if ((size_t)(i - 1) >= gt.size()) throw std::invalid_argument("The corresponding time is not inserted!... exiting.");
            PoseManager.XYYaw2SE2(gt[i - 1], Tgt);
        }
        curr_gt = curr_gt * Tgt;
        curr_odom = curr_odom * Test;
        ++n_pairs;
        if (write) {
            PoseManager.writePose(Tgt, Test);
            PoseManager.writePose(Tgt, Test, time1, time2, time_unit_denominator);
            PoseManager.writeTraj(curr_odom, PoseManager.est_traj_path_);
            PoseManager.writeTraj(curr_gt, PoseManager.gt_traj_path_);
            PoseManager.writeTimeTaken(time_taken, Tgt);
        }
    }
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
    std::printf("%d pairs (%d skipped), %.1f pairs/s incl. PNG decode, estimator %.3f ms per pair; final pose x=%.6g y=%.6g yaw=%.6g; "
                "%d loop closures\n", n_pairs, n_skipped, n_pairs / std::max(wall, 1e-9), n_pairs ? 1e3 * solve_s / n_pairs : 0.0,
                curr_odom(0, 3), curr_odom(1, 3), nrs::euler_yaw(curr_odom), n_loops);
    return 0;
}

}  // namespace

int main(int argc, char **argv)
{
    try {
        if (argc == 4 && !std::strcmp(argv[1], "--decode-only")) return decode_only(argv[2], argv[3]);
        if (argc == 6 && !std::strcmp(argv[1], "--gt-lookup")) {
            nrs::SE2PoseManager pm;
            pm.dataset_name_ = argv[3];
            pm.gt_path_ = std::string(argv[2]) + "/gt/radar_odometry.csv";
            if (pm.dataset_name_ == "mulran") nrs::loadGTPoses(pm.gt_path_, pm.gt_rel_);
            const std::vector<float> g = pm.dataset_name_ == "oxford" ? pm.getXYYawPose(std::atoll(argv[4]), std::atoll(argv[5]))
                                                                      : pm.getXYYawPose(std::atoll(argv[4]));
            std::printf("%.9g %.9g %.9g\n", g[0], g[1], g[2]);
            return 0;
        }
        if (argc == 6 && !std::strcmp(argv[1], "--posefile-selftest")) return posefile_selftest(argv[2], argv[3], argv[4], argv[5]);
        Options o;
        if (!parse(argc, argv, o)) {
            std::fprintf(stderr, "usage: orora_odom <dir of PNGs> [--dataset mulran|oxford] [--extractor cen2018|cen2019] [--out DIR] "
                                 "[--gt FILE] [--loops] ... (see the head of orora_odom.cpp)\n");
            return 1;
        }
        return run(o);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "orora_odom: %s\n", e.what());
        return 1;
    }
}
