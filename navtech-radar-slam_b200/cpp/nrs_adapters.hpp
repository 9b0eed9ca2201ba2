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
//   nrs::ORORAT<MatrixXd>              <-> include/ORORA.hpp:13-181 (calcTransform, getMaxCliques, getFinalInliers)
//   nrs::NonMinimalSolverT<MatrixXd>   <-> include/solver/nonminimal_solver.hpp:52-957 (reset, setInputSource/Target,
//                                          computeTransformation, getters)
//   nrs::ShardedSCManagerT             <-> SCManager over a keyframe-sharded database, one process per GPU
//   nrs::setParams                     <-> include/solver/nonminimal_solver.hpp:959-973
//
// Eigen / PCL / OpenCV are not available in the build container, so the adapters are templates over
// the matrix and cloud types: anything with Eigen::MatrixXd's interface (column-major data(),
// rows(), cols(), resize(), operator()(i,j)) and pcl::PointCloud's (points[i].x/y/z) fits.  With
// the real libraries:  using SCManager = nrs::SCManagerT<Eigen::MatrixXd, pcl::PointCloud<pcl::PointXYZI>>;
// navtech-radar-slam_b200/cpp/adapter_check.cpp instantiates them with minimal stand-ins.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <functional>
#include <thread>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "nrs_b200.h"

namespace nrs {

// NRS_ECAPACITY as its own type: a caller's buffer or a fixed device-side capacity was too small (the one the
// reference does not have: more than 1024 correspondences in one solve, reachable only with the stop-motion gate off)
struct CapacityError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void check(int rc)
{
    if (rc == NRS_EINVAL) throw std::invalid_argument(nrs_last_error());
    if (rc == NRS_ECAPACITY) throw CapacityError(std::string("nrs_b200: ") + nrs_last_error());
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
    // Scancontext.cpp:267-328.  The reference builds its batch tree ONCE, on the first call, over the keys
    // present then (is_tree_batch_made): later calls search exactly those keyframes, whatever was appended
    // since.  Candidates the k-NN search cannot fill stay 0 in the reference (zero-initialised vector, :292).
    std::pair<int, float> detectLoopClosureIDBetweenSession(std::vector<float> &curr_key, MatrixXd &curr_desc)
    {
        if (!is_tree_batch_made) {
            batch_end_ = (int)polarcontext_invkeys_mat_.size();
            is_tree_batch_made = true;
        }
        nrs_sc_result r;
        const int n = batch_end_;
        if (use_tree_candidates && (int)curr_key.size() == PC_NUM_RING && n > 0) {
            int32_t cand[16];
            check(nrs_sc_knn(db_, curr_key.data(), 0, n, NUM_CANDIDATES_FROM_TREE, cand, nullptr));
            for (int m = 0; m < NUM_CANDIDATES_FROM_TREE; ++m) if (cand[m] < 0) cand[m] = 0;
            check(nrs_sc_query(db_, curr_desc.data(), 0, 0, cand, NUM_CANDIDATES_FROM_TREE, &r));
        } else
            check(nrs_sc_query(db_, curr_desc.data(), 0, n, nullptr, 0, &r));
        return {r.dist < SC_DIST_THRES ? r.idx : -1, yaw(r.shift)};
    }
    const MatrixXd &getConstRefRecentSCD(void) { return polarcontexts_.back(); }
    void setSCdistThres(double _new_thres) { SC_DIST_THRES = _new_thres; }
    // (extension) the raw search loop of :380-396 over keyframes [first, last): (min_dist, nn_idx, nn_align)
    nrs_sc_result searchRange(MatrixXd &query, int first, int last)
    {
        nrs_sc_result r;
        check(nrs_sc_query(db_, query.data(), first, last, nullptr, 0, &r));
        return r;
    }

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
    int batch_end_ = 0;
};


// ================================================================================================
// Keyframe-sharded SCManager: one process per GPU, keyframe g lives on rank g % world.
// Same method names as SCManager; every rank calls every method with the same arguments (SPMD) and gets
// the same answers.  The per-shard (dist, idx, shift) records are merged inside the search kernel over
// peer memory (CUDA IPC handles travel once through `allgather`), or -- when peers cannot be mapped or
// prefer_nccl is set -- by two ncclAllReduce(min) calls (sc_nccl.cu).  `allgather(mine, all, bytes)` is any
// host channel that concatenates `bytes` bytes of every rank in rank order (MPI_Allgather, sockets, files).
// ================================================================================================
template <class MatrixXd, class Cloud>
class ShardedSCManagerT {
public:
    using AllGather = std::function<void(const void *mine, void *all, size_t bytes)>;

    ShardedSCManagerT(int world, int rank, AllGather allgather, nrs_ctx *ctx = nullptr, bool prefer_nccl = false)
        : world_(world), rank_(rank), ctx_(ctx ? ctx : default_ctx())
    {
        if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("ShardedSCManager: bad world / rank");
        check(nrs_sc_db_create(ctx_, 1024, &db_));
        if (world == 1) return;
        int ok = 0;
        if (!prefer_nccl) {
            std::vector<unsigned char> mine(NRS_SC_XCHG_HANDLE_BYTES), all((size_t)world * NRS_SC_XCHG_HANDLE_BYTES);
            const int rc = nrs_sc_xchg_create(db_, world, rank, mine.data());
            allgather(mine.data(), all.data(), mine.size());
            ok = (rc == NRS_OK && nrs_sc_xchg_open(db_, all.data()) == NRS_OK) ? 1 : 0;
        }
        std::vector<int> oks(world);
        allgather(&ok, oks.data(), sizeof(int));
        bool peer = !prefer_nccl;
        for (int v : oks) peer = peer && v;
        exchange_ = peer ? "peer" : "nccl";
        if (!peer) {
            std::vector<unsigned char> id(NRS_NCCL_UNIQUE_ID_BYTES), ids((size_t)world * NRS_NCCL_UNIQUE_ID_BYTES);
            if (rank == 0) check(nrs_nccl_get_unique_id(id.data()));
            allgather(id.data(), ids.data(), id.size());
            check(nrs_sc_nccl_init(db_, world, rank, ids.data()));          // rank 0's id
        }
    }
    ~ShardedSCManagerT() { nrs_sc_db_destroy(db_); }
    ShardedSCManagerT(const ShardedSCManagerT &) = delete;
    ShardedSCManagerT &operator=(const ShardedSCManagerT &) = delete;

    void makeAndSaveScancontextAndKeys(Cloud &_scan_down)
    {
        std::vector<float> xyz(3 * _scan_down.points.size());
        for (size_t i = 0; i < _scan_down.points.size(); ++i) {
            xyz[3 * i] = _scan_down.points[i].x; xyz[3 * i + 1] = _scan_down.points[i].y; xyz[3 * i + 2] = _scan_down.points[i].z;
        }
        MatrixXd sc(NRS_SC_RINGS, NRS_SC_SECTORS);
        check(nrs_sc_make(ctx_, xyz.data(), (int)(xyz.size() / 3), LIDAR_HEIGHT, PC_MAX_RADIUS, sc.data()));
        saveScancontextAndKeys(sc);
    }
    void saveScancontextAndKeys(MatrixXd _scd)
    {
        if ((int)(polarcontexts_.size() % (size_t)world_) == rank_) check(nrs_sc_db_append(db_, _scd.data(), 1));   // the owner stores it
        MatrixXd ringkey(NRS_SC_RINGS, 1);
        check(nrs_sc_keys(ctx_, _scd.data(), ringkey.data(), nullptr));
        polarcontext_invkeys_mat_.push_back(std::vector<float>(ringkey.data(), ringkey.data() + NRS_SC_RINGS));
        polarcontexts_.push_back(_scd);
    }
    // Scancontext.cpp:331-422 with the exhaustive search of north_star over the set the reference's tree would hold
    std::pair<int, float> detectLoopClosureID(void)
    {
        if ((int)polarcontexts_.size() < NUM_EXCLUDE_RECENT + 1) return {-1, 0.0f};
        if (tree_making_period_conter % TREE_MAKING_PERIOD_ == 0) searchable_end_ = (int)polarcontexts_.size() - NUM_EXCLUDE_RECENT;
        tree_making_period_conter = tree_making_period_conter + 1;
        const nrs_sc_result r = search(polarcontexts_.back(), 0, searchable_end_);
        return {r.dist < SC_DIST_THRES ? r.idx : -1, yaw(r.shift)};
    }
    // one sharded search of global keyframes [first, last): identical (dist, global idx, shift) on every rank
    nrs_sc_result search(MatrixXd &query, int first, int last)
    {
        nrs_sc_result r;
        if (world_ == 1) { check(nrs_sc_query(db_, query.data(), first, last, nullptr, 0, &r)); return r; }
        int lo, hi;
        local_range(first, last, lo, hi);
        check(nrs_sc_query_batch_sharded(db_, query.data(), 1, lo, hi, &r));
        if (r.idx == -2) throw std::runtime_error("ShardedSCManager: a peer rank did not answer within the exchange timeout");
        return r;
    }
    // Q queries (Q x 1200 doubles, query-major), one launch and ONE exchange
    std::vector<nrs_sc_result> searchBatch(const std::vector<double> &queries, int first, int last)
    {
        const int Q = (int)(queries.size() / NRS_SC_DESC);
        std::vector<nrs_sc_result> out((size_t)Q);
        if (Q == 0) return out;
        if (world_ == 1) { check(nrs_sc_query_batch(db_, queries.data(), Q, first, last, out.data())); return out; }
        int lo, hi;
        local_range(first, last, lo, hi);
        check(nrs_sc_query_batch_sharded(db_, queries.data(), Q, lo, hi, out.data()));
        return out;
    }
    void setSCdistThres(double _new_thres) { SC_DIST_THRES = _new_thres; }
    const MatrixXd &getConstRefRecentSCD(void) { return polarcontexts_.back(); }
    const char *exchange() const { return exchange_; }
    int localSize() { return nrs_sc_db_size(db_); }

    const double LIDAR_HEIGHT = 2.0, PC_MAX_RADIUS = 80.0;
    const int NUM_EXCLUDE_RECENT = 30, TREE_MAKING_PERIOD_ = 30;
    double SC_DIST_THRES = 0.2;
    int tree_making_period_conter = 0;
    std::vector<MatrixXd> polarcontexts_;                       // every rank keeps the public vectors (they are host data)
    std::vector<std::vector<float>> polarcontext_invkeys_mat_;

private:
    // local slots [lo, hi) of this rank whose global index slot * world + rank falls in [first, last)
    void local_range(int first, int last, int &lo, int &hi) const
    {
        auto ceil_div = [](int a, int b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); };
        lo = std::max(0, ceil_div(first - rank_, world_));
        hi = std::max(lo, ceil_div(last - rank_, world_));
        const int n = nrs_sc_db_size(db_);
        if (hi > n) hi = n;
        if (lo > hi) lo = hi;
    }
    float yaw(int nn_align) const
    {
        float degrees = (float)(nn_align * (360.0 / double(NRS_SC_SECTORS)));
        return (float)(degrees * M_PI / 180.0);
    }
    int world_, rank_;
    nrs_ctx *ctx_;
    nrs_sc_db *db_ = nullptr;
    int searchable_end_ = 0;
    const char *exchange_ = "none";
};

// A file-based all-gather for hosts without MPI (tests, single-node launchers): every call writes
// <dir>/ag_<seq>_<rank> and waits for the world's files of the same sequence number.
class FileAllGather {
public:
    FileAllGather(std::string dir, int world, int rank) : dir_(std::move(dir)), world_(world), rank_(rank) {}
    void operator()(const void *mine, void *all, size_t bytes)
    {
        const std::string base = dir_ + "/ag_" + std::to_string(seq_++) + "_";
        {
            const std::string tmp = base + std::to_string(rank_) + ".tmp", fin = base + std::to_string(rank_);
            FILE *f = std::fopen(tmp.c_str(), "wb");
            if (!f || std::fwrite(mine, 1, bytes, f) != bytes) throw std::runtime_error("FileAllGather: cannot write " + tmp);
            std::fclose(f);
            if (std::rename(tmp.c_str(), fin.c_str()) != 0) throw std::runtime_error("FileAllGather: rename failed");
        }
        for (int r = 0; r < world_; ++r) {
            const std::string fin = base + std::to_string(r);
            for (int tries = 0;; ++tries) {
                FILE *f = std::fopen(fin.c_str(), "rb");
                if (f) {
                    const size_t got = std::fread(static_cast<unsigned char *>(all) + (size_t)r * bytes, 1, bytes, f);
                    std::fclose(f);
                    if (got == bytes) break;
                }
                if (tries > 60000) throw std::runtime_error("FileAllGather: rank " + std::to_string(r) + " never wrote " + fin);
                std::this_thread::sleep_for(std::chrono::milliseconds(2));
            }
        }
    }
private:
    std::string dir_;
    int world_, rank_, seq_ = 0;
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
// Mat / KeyPoint: the node's own types where OpenCV is around -- RadarFeatureT<Eigen::MatrixXd, cv::Mat, cv::KeyPoint> has
// the reference's exact members (cv::Mat img_ / desc_, std::vector<cv::KeyPoint> kp_), so its visualisation branch
// (odometry.cpp:335-339,362) compiles against it unchanged.  The defaults are plain buffers: img_ = W*W bytes,
// desc_ = K_desc*32 bytes, kp_ = K_desc (u, v) float pairs.
struct RawKeyPoint { float u, v; };

namespace detail {
// 8-bit single-channel matrix <- bytes: std::vector<uint8_t>, or anything with cv::Mat's create(rows, cols, type) / data
inline void fill_bytes(std::vector<uint8_t> &dst, const uint8_t *src, int rows, int cols) { dst.assign(src, src + (size_t)rows * cols); }
template <class Mat>
auto fill_bytes(Mat &dst, const uint8_t *src, int rows, int cols) -> decltype(dst.create(rows, cols, 0), void())
{
    dst.create(rows, cols, 0 /* CV_8UC1 */);
    if (rows && cols) std::memcpy(dst.data, src, (size_t)rows * cols);
}
// keypoints <- K (u, v) pairs: RawKeyPoint keeps the pair, anything else is built like cv::KeyPoint(x, y, size)
inline void fill_keypoints(std::vector<RawKeyPoint> &dst, const float *uv, int n, float)
{
    dst.resize((size_t)n);
    for (int i = 0; i < n; ++i) dst[(size_t)i] = RawKeyPoint{uv[2 * i], uv[2 * i + 1]};
}
template <class KeyPoint>
void fill_keypoints(std::vector<KeyPoint> &dst, const float *uv, int n, float size)
{
    dst.clear();
    dst.reserve((size_t)n);
    for (int i = 0; i < n; ++i) dst.emplace_back(uv[2 * i], uv[2 * i + 1], size);
}
}  // namespace detail

template <class MatrixXd, class Mat = std::vector<uint8_t>, class KeyPoint = RawKeyPoint>
struct RadarFeatureT {
    Mat img_;                                  // 964 x 964 (CV_8UC1)
    std::vector<int64_t> t_;
    double time_unit_denominator_ = 1.0;
    Mat desc_;                                 // K_desc x 32 (CV_8UC1)
    std::vector<KeyPoint> kp_;                 // K_desc keypoints (u, v), size = patch_size
    MatrixXd feat_in_cart_;
    MatrixXd feat_in_polar_;
    MatrixXd feat_compensated_;
    int64_t t_crietria_ = 0;
    std::vector<double> t_norm_;
    std::shared_ptr<nrs_frame> frame_;         // device-resident copy used by matchFeatures
};

template <class MatrixXd, class Mat = std::vector<uint8_t>, class KeyPoint = RawKeyPoint>
class RadarLoaderT {
public:
    using RadarFeature = RadarFeatureT<MatrixXd, Mat, KeyPoint>;
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
        r.t_.resize(info.K);
        r.t_norm_.resize(info.K);
        const int W = p_.cart_pixel_width;
        std::vector<float> kp(2 * (size_t)info.K_desc + 2);
        std::vector<uint8_t> desc(32 * (size_t)info.K_desc + 32), img((size_t)W * W);
        check(nrs_frame_get(f, r.feat_in_polar_.data(), r.feat_in_cart_.data(), kp.data(), r.t_.data(),
                            r.t_norm_.data(), desc.data(), img.data()));
        detail::fill_bytes(r.img_, img.data(), W, W);
        detail::fill_bytes(r.desc_, desc.data(), info.K_desc, 32);
        detail::fill_keypoints(r.kp_, kp.data(), info.K_desc, (float)p_.patch_size);    // cv::KeyPoint(u, v, patch_size): radar_utils.cpp:393
        r.t_crietria_ = t_criteria;
        r.time_unit_denominator_ = time_unit_denominator_;
        return r;
    }

    FeatureMatching matchFeatures(const RadarFeature &prev_feat, const RadarFeature &curr_feat)
    {
        RadarFeature prev_matched = prev_feat, curr_matched = curr_feat;           // :144-146
        const int cap = (int)prev_feat.kp_.size() + 1;
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
        prev_matched.kp_.resize(n);
        curr_matched.kp_.resize(n);
        for (int j = 0; j < n; ++j) {
            prev_matched.t_norm_[j] = prev_feat.t_norm_[q[j]];
            curr_matched.t_norm_[j] = curr_feat.t_norm_[t[j]];
            // "For visualization" (:189-191): BOTH come from prev_feat.kp_ in the reference, the second indexed with the train
            // index -- out of range there when the current frame has more keypoints; clamped here
            prev_matched.kp_[j] = prev_feat.kp_[q[j]];
            curr_matched.kp_[j] = prev_feat.kp_[std::min<size_t>((size_t)t[j], prev_feat.kp_.size() - 1)];
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

// pcl::PointCloud<pcl::PointXYZ> stand-in for InlierSetPairs when PCL is not around
struct SimplePoint { float x, y, z; };
struct SimpleCloud { std::vector<SimplePoint> points; };

// ================================================================================================
// NonMinimalSolver (solver/nonminimal_solver.hpp:52-957): reset / setInputSource / setInputTarget /
// computeTransformation and the getters the node and ORORA.hpp call.  Matrix4d = anything with data()
// (column-major 16 doubles), e.g. Eigen::Matrix4d.
// ================================================================================================
template <class MatrixXd, class Matrix4d = MatrixXd, class Cloud = SimpleCloud>
class NonMinimalSolverT {
public:
    using Params = nrs_solver_params;                                    // nonminimal_solver.hpp:104-162
    using InlierSetPairs = std::tuple<Cloud, Cloud>;                     // :55
    struct RegistrationSolution {                                        // :63-70
        bool valid = true;
        double scale = 1.0;
        double translation[3] = {0, 0, 0};
        double rotation[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};                // 3x3 column-major
    };
    RegistrationSolution solution_;

    explicit NonMinimalSolverT(nrs_ctx *ctx = nullptr) : ctx_(ctx ? ctx : default_ctx()) { nrs_solver_default_params(&params_); }

    Params getParams() { return params_; }
    void setParams(Params params) { params_ = params; }

    void setInputSource(const MatrixXd &cloudEigen) { take(cloudEigen, src_); }                  // :171-176
    void setInputTarget(const MatrixXd &cloudEigen) { take(cloudEigen, tgt_); }                  // :178-183

    void reset(const Params &params, const double noise_bound_radial, const double noise_bound_tangential)   // :665-676
    {
        params_ = params;
        params_.noise_bound_radial = noise_bound_radial;
        params_.noise_bound_tangential = noise_bound_tangential;
        max_clique_.clear(); final_inliers_.clear();
        info_ = nrs_solver_info{};
        solution_ = RegistrationSolution{};
    }

    // :678-825.  initial_guess is accepted and, as in the reference (:332-333), unused.  On failure the output
    // stays Identity and solution_.valid = false (:715-720) -- without the reference's blocking cin.ignore().
    void computeTransformation(const Matrix4d & /*initial_guess*/, Matrix4d &output)
    {
        const int N = (int)(src_.size() / 2);
        if ((int)(tgt_.size() / 2) != N) throw std::invalid_argument("source and target must hold the same number of points");
        Params p = params_;
        p.use_doppler_compensation = 0;                  // the solver class never compensates; ORORA does (ORORA.hpp:56-65)
        p.check_stop_motion = 0;
        std::vector<int32_t> clique(N > 0 ? N : 1), fin(N > 0 ? N : 1);
        double T[16];
        // nrs_orora_solve(prev = target, curr = source): ORORA.hpp:83-85
        check(nrs_orora_solve(ctx_, tgt_.data(), src_.data(), N, &p, 0.0, T, clique.data(), fin.data(), &info_));
        max_clique_.assign(clique.begin(), clique.begin() + info_.M);
        final_inliers_.assign(fin.begin(), fin.begin() + info_.n_final_inliers);
        solution_.valid = info_.valid != 0;
        solution_.translation[0] = T[12]; solution_.translation[1] = T[13]; solution_.translation[2] = 0.0;
        solution_.rotation[0] = T[0]; solution_.rotation[1] = T[1]; solution_.rotation[3] = T[4]; solution_.rotation[4] = T[5];
        std::memcpy(output.data(), T, sizeof(T));
    }

    InlierSetPairs getMaxCliques() { return {setInliers(src_, max_clique_), setInliers(tgt_, max_clique_)}; }          // :872-876
    InlierSetPairs getFinalInliers() { return {setInliers(src_, final_inliers_), setInliers(tgt_, final_inliers_)}; }  // :878-882
    std::vector<int> getFinalInliersIndices() { return final_inliers_; }                                                // :884
    int getNumRotaionInliers() { return info_.n_rot_inliers; }                                                          // :888
    int getNumMaxCliqueInliers() { return info_.M; }                                                                    // :892
    const nrs_solver_info &info() const { return info_; }

private:
    static void take(const MatrixXd &m, std::vector<double> &xy)        // top two rows, homogeneous 1 implied (:173-175)
    {
        const int N = (int)m.cols();
        xy.resize(2 * (size_t)N);
        for (int j = 0; j < N; ++j) { xy[2 * j] = m(0, j); xy[2 * j + 1] = m(1, j); }
    }
    static Cloud setInliers(const std::vector<double> &xy, const std::vector<int> &idx)           // :827-845, z = the homogeneous 1
    {
        Cloud c;
        c.points.reserve(idx.size());
        for (const int i : idx) {
            typename std::remove_reference<decltype(c.points[0])>::type p{};
            p.x = (float)xy[2 * (size_t)i]; p.y = (float)xy[2 * (size_t)i + 1]; p.z = 1.0f;
            c.points.push_back(p);
        }
        return c;
    }
    nrs_ctx *ctx_;
    Params params_;
    nrs_solver_info info_{};
    std::vector<double> src_, tgt_;
    std::vector<int> max_clique_, final_inliers_;
};

template <class MatrixXd, class Cloud = SimpleCloud>
class ORORAT {
public:
    using RadarFeature = RadarFeatureT<MatrixXd>;              // calcTransform takes any RadarFeatureT<MatrixXd, Mat, KeyPoint>
    using Solver = NonMinimalSolverT<MatrixXd, MatrixXd, Cloud>;
    using InlierSetPairs = typename Solver::InlierSetPairs;

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
    template <class Feature>
    MatrixXd calcTransform(const nrs_solver_params params, const double delta_t, Feature &prev_matched, Feature &curr_matched)
    {
        const double prev_vel_x = curr_vel_x_;                                      // prev_vel_se3_ = curr_vel_se3_ (:42)
        // feat_compensated_ is a public product of the call (the node draws it, odometry.cpp:344-345)
        if (use_doppler_) {
            prev_matched.feat_compensated_ = compensateDopplerDistortion(prev_vel_x, beta_, prev_matched.feat_in_cart_);
            curr_matched.feat_compensated_ = compensateDopplerDistortion(prev_vel_x, beta_, curr_matched.feat_in_cart_);
        } else {
            prev_matched.feat_compensated_ = prev_matched.feat_in_cart_;            // :63-64
            curr_matched.feat_compensated_ = curr_matched.feat_in_cart_;
        }
        MatrixXd out(4, 4);
        for (int i = 0; i < 4; ++i) out(i, i) = 1.0;
        const int N = (int)curr_matched.feat_compensated_.cols();
        if (check_stop_motion_ && N > stop_thr_) {
            stop_motion_ = true;                                                    // :74-79: Identity, solver untouched
        } else {
            stop_motion_ = false;
            solver_.reset(new Solver(ctx_));                                        // :81-85
            solver_->reset(params, nbr_, nbt_);
            solver_->setInputSource(curr_matched.feat_compensated_);
            solver_->setInputTarget(prev_matched.feat_compensated_);
            solver_->computeTransformation(out, out);
        }
        curr_vel_x_ = out(0, 3) / delta_t;                                          // :89 + association.cpp:215
        return out;
    }

    // ORORA.hpp:139-151 (the updated x enters the y update, :146-147)
    MatrixXd compensateDopplerDistortion(const double vel_x, const double beta, const MatrixXd &p2d)
    {
        const int N = (int)p2d.cols();
        MatrixXd compensated(2, N);
        const double v = std::fabs(vel_x);
        for (int i = 0; i < N; ++i) {
            double p0 = p2d(0, i), p1 = p2d(1, i);
            volatile double xx = p0 * p0, yy = p1 * p1;           // products rounded before the sum (no FMA contraction)
            const double rsq = xx + yy;
            p0 += beta * v * p0 * p0 / rsq;
            p1 += beta * v * p0 * p1 / rsq;
            compensated(0, i) = p0; compensated(1, i) = p1;
        }
        return compensated;
    }

    // Use it after calculating relative pose! (ORORA.hpp:153-161)
    InlierSetPairs getMaxCliques() { need_solver(); return solver_->getMaxCliques(); }
    InlierSetPairs getFinalInliers() { need_solver(); return solver_->getFinalInliers(); }

    int getNumMaxCliqueInliers() const { return solver_ ? solver_->info().M : 0; }
    int getNumRotaionInliers() const { return solver_ ? solver_->info().n_rot_inliers : 0; }
    std::vector<int> getFinalInliersIndices() const { return solver_ ? solver_->getFinalInliersIndices() : std::vector<int>(); }
    bool stopMotionDetected() const { return stop_motion_; }
    nrs_solver_info info() const { return solver_ ? solver_->info() : nrs_solver_info{}; }

private:
    void need_solver() const { if (!solver_) throw std::runtime_error("ORORA: no pose has been calculated yet"); }
    nrs_ctx *ctx_;
    std::string dataset_, deskewing_target_;
    bool use_doppler_, check_stop_motion_;
    int stop_thr_;
    double beta_, nbr_, nbt_;
    double curr_vel_x_ = 0.0;
    bool stop_motion_ = false;
    std::unique_ptr<Solver> solver_;
};

}  // namespace nrs
