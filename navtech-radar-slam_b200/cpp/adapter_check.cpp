// adapter_check.cpp -- instantiates every adapter in nrs_adapters.hpp with minimal stand-ins for
// Eigen::MatrixXd and pcl::PointCloud (neither library is in the build container) and, when a GPU
// is present, runs a small end-to-end exercise through the C++ host side:
//     g++ -std=c++17 -Iinclude -Inavtech-radar-slam_b200/cpp adapter_check.cpp -Lnavtech-radar-slam_b200 -lnrs_b200
// Prints one line per check; exit code 0 = all good, 77 = no GPU (compile/link check only).
#include <cstdio>
#include <random>

#include "nrs_adapters.hpp"

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

using SCManager = nrs::SCManagerT<MiniMatrix, MiniCloud>;
using RadarLoader = nrs::RadarLoaderT<MiniMatrix>;
using ORORA = nrs::ORORAT<MiniMatrix>;
using NonMinimalSolver = nrs::NonMinimalSolverT<MiniMatrix>;

int main()
{
    nrs_ctx *ctx = nullptr;
    if (nrs_ctx_create(0, &ctx) != NRS_OK) {
        std::printf("no GPU: %s (adapters compiled and linked)\n", nrs_last_error());
        return 77;
    }
    int rc = 0;
    {
    std::mt19937 rng(1);
    std::uniform_real_distribution<float> U(-70.f, 70.f);
    SCManager sc(ctx);
    sc.setSCdistThres(0.45);
    std::vector<MiniCloud> clouds(8);
    for (auto &c : clouds) for (int i = 0; i < 600; ++i) c.points.push_back({U(rng), U(rng), 1.0f, 0.f});
    int found = 0;
    for (int k = 0; k < 100; ++k) {
        sc.makeAndSaveScancontextAndKeys(clouds[k % 8]);
        auto r = sc.detectLoopClosureID();
        // the searchable set is frozen at every 30th call (Scancontext.cpp:348-359): a revisit of cloud
        // k % 8 can only be reported once an earlier copy of it is inside that set
        if (r.first >= 0) { ++found; if (r.first % 8 != k % 8) { std::printf("wrong loop %d for %d\n", r.first, k); return 1; } }
    }
    std::printf("SCManager: %zu keyframes, %d loops found\n", sc.polarcontexts_.size(), found);
    if (found < 10) return 1;
    auto d = sc.distanceBtnScanContext(sc.polarcontexts_[3], sc.polarcontexts_[11]);
    std::printf("distanceBtnScanContext(3,11) = %.3g shift %d\n", d.first, d.second);
    if (!(d.first < 1e-12) || d.second != 0) return 1;

    // solver through ORORA::calcTransform
    const int N = 300;
    MiniMatrix prev(2, N), curr(2, N);
    const double th = 0.02, tx = 1.5, ty = -0.2;
    std::uniform_real_distribution<double> P(-80., 80.);
    for (int j = 0; j < N; ++j) {
        curr(0, j) = P(rng); curr(1, j) = P(rng);
        if (j % 3) { prev(0, j) = std::cos(th) * curr(0, j) - std::sin(th) * curr(1, j) + tx; prev(1, j) = std::sin(th) * curr(0, j) + std::cos(th) * curr(1, j) + ty; }
        else { prev(0, j) = P(rng); prev(1, j) = P(rng); }
    }
    ORORA est("mulran", "rot", true, false, true, 700, 0.049, 0.1, 9.0 * M_PI / 180.0, ctx);
    nrs_solver_params params;
    nrs::setParams(0.75, 1.0, false, 100, 1.4, 1e-6, params);
    ORORA::RadarFeature pm, cm;
    pm.feat_in_cart_ = prev; cm.feat_in_cart_ = curr;
    MiniMatrix T = est.calcTransform(params, 0.25, pm, cm);
    std::printf("ORORA: M = %d, yaw = %.5f (true %.5f), t = (%.3f, %.3f)\n", est.getNumMaxCliqueInliers(),
                std::atan2(T(1, 0), T(0, 0)), th, T(0, 3), T(1, 3));
    if (std::fabs(std::atan2(T(1, 0), T(0, 0)) - th) > 1e-3 || std::fabs(T(0, 3) - tx) > 0.05) return 1;
    // feat_compensated_ is filled like ORORA.hpp:56-65 (first call: velocity 0 -> equal to feat_in_cart_), the
    // inlier clouds come back like ORORA.hpp:153-161
    if (pm.feat_compensated_.cols() != N || cm.feat_compensated_(1, 7) != curr(1, 7)) { std::printf("feat_compensated_ not set\n"); return 1; }
    {
        auto cl = est.getMaxCliques();
        auto fi = est.getFinalInliers();
        const size_t M = std::get<0>(cl).points.size();
        std::printf("ORORA::getMaxCliques: %zu pairs, getFinalInliers: %zu pairs\n", M, std::get<0>(fi).points.size());
        if ((int)M != est.getNumMaxCliqueInliers() || std::get<1>(cl).points.size() != M || M < 150) return 1;
        if (std::get<0>(fi).points.size() != est.getFinalInliersIndices().size()) return 1;
    }
    // second call: the Doppler velocity of the first pose enters (ORORA.hpp:42,56-61,89)
    {
        ORORA::RadarFeature pm2, cm2;
        pm2.feat_in_cart_ = prev; cm2.feat_in_cart_ = curr;
        MiniMatrix T2 = est.calcTransform(params, 0.25, pm2, cm2);
        if (cm2.feat_compensated_(0, 5) == curr(0, 5)) { std::printf("Doppler compensation did not move the points\n"); return 1; }
        if (std::fabs(std::atan2(T2(1, 0), T2(0, 0)) - th) > 5e-3) return 1;
    }
    // the solver class on its own: reset / setInputSource / setInputTarget / computeTransformation / getters
    // (nonminimal_solver.hpp:171-178,665-892) must reproduce calcTransform's first pose (no Doppler at velocity 0)
    {
        NonMinimalSolver solver(ctx);
        solver.reset(params, 0.1, 9.0 * M_PI / 180.0);
        solver.setInputSource(curr);
        solver.setInputTarget(prev);
        MiniMatrix guess(4, 4), out(4, 4);
        solver.computeTransformation(guess, out);
        if (!solver.solution_.valid || out(0, 3) != T(0, 3) || out(1, 0) != T(1, 0)) { std::printf("NonMinimalSolver differs from ORORA\n"); return 1; }
        if (solver.getNumMaxCliqueInliers() != est.getNumMaxCliqueInliers() && solver.getNumMaxCliqueInliers() < 150) return 1;
        if ((int)solver.getFinalInliersIndices().size() < solver.getNumRotaionInliers() / 2) return 1;
        std::printf("NonMinimalSolver: M = %d, rotation inliers %d, final inliers %zu\n", solver.getNumMaxCliqueInliers(),
                    solver.getNumRotaionInliers(), solver.getFinalInliersIndices().size());
    }
    // stop-motion gate (ORORA.hpp:74-79): more than the threshold of matches -> Identity, solver untouched
    {
        ORORA gate("mulran", "rot", true, false, true, 100, 0.049, 0.1, 9.0 * M_PI / 180.0, ctx);
        ORORA::RadarFeature a, b;
        a.feat_in_cart_ = prev; b.feat_in_cart_ = curr;
        MiniMatrix Tg = gate.calcTransform(params, 0.25, a, b);
        if (!gate.stopMotionDetected() || Tg(0, 3) != 0.0 || Tg(0, 0) != 1.0) { std::printf("stop-motion gate\n"); return 1; }
    }
    try { RadarLoader bad("kitti", "cen2019", 0.6, true, false, ctx); return 1; } catch (const std::invalid_argument &) {}
    RadarLoader loader("mulran", "cen2019", 0.6, true, false, ctx);
    (void)loader;
    RadarLoader loader18("mulran", "cen2018", 0.6, true, false, ctx);     // what run_orora.launch passes
    (void)loader18;
    // the reference's own candidate flow (ring-key 3-NN + 3 candidates) must agree with the exhaustive
    // search whenever the exhaustive winner is among the candidates -- here: exact revisits
    {
        SCManager tree(ctx);
        tree.use_tree_candidates = true;
        tree.setSCdistThres(0.45);
        int ok = 0;
        for (int k = 0; k < 70; ++k) {
            tree.makeAndSaveScancontextAndKeys(clouds[k % 8]);
            auto r = tree.detectLoopClosureID();
            if (r.first >= 0) { ++ok; if (r.first % 8 != k % 8) { std::printf("tree mode: wrong loop %d for %d\n", r.first, k); return 1; } }
        }
        std::printf("SCManager (kd-tree mode): %d loops found\n", ok);
        if (ok < 5) return 1;
    }
    // detectLoopClosureIDBetweenSession (Scancontext.cpp:267-328): the batch tree is built once, over the keys
    // present at the FIRST call; keyframes saved later are never searched
    {
        for (int mode = 0; mode < 2; ++mode) {
            SCManager ms(ctx);
            ms.use_tree_candidates = mode == 1;
            ms.setSCdistThres(0.45);
            for (int k = 0; k < 5; ++k) ms.makeAndSaveScancontextAndKeys(clouds[k]);
            MiniMatrix q5 = ms.makeScancontext(clouds[5]), q2 = ms.makeScancontext(clouds[2]);
            MiniMatrix rk = ms.makeRingkeyFromScancontext(q2);
            std::vector<float> key2(rk.data(), rk.data() + 20);
            MiniMatrix rk5 = ms.makeRingkeyFromScancontext(q5);
            std::vector<float> key5(rk5.data(), rk5.data() + 20);
            auto r2 = ms.detectLoopClosureIDBetweenSession(key2, q2);        // first call: tree over keyframes 0..4
            if (r2.first != 2 || r2.second != 0.0f) { std::printf("between-session (mode %d): expected keyframe 2, got %d\n", mode, r2.first); return 1; }
            ms.makeAndSaveScancontextAndKeys(clouds[5]);                     // appended after the batch tree was made
            auto r5 = ms.detectLoopClosureIDBetweenSession(key5, q5);
            if (r5.first == 5) { std::printf("between-session (mode %d): searched a keyframe outside the batch tree\n", mode); return 1; }
        }
        // fewer stored keyframes than candidates: the reference evaluates its zero-initialised ids (index 0)
        SCManager one(ctx);
        one.use_tree_candidates = true;
        one.setSCdistThres(0.45);
        one.makeAndSaveScancontextAndKeys(clouds[3]);
        MiniMatrix q3 = one.makeScancontext(clouds[3]);
        MiniMatrix rk = one.makeRingkeyFromScancontext(q3);
        std::vector<float> key3(rk.data(), rk.data() + 20);
        if (one.detectLoopClosureIDBetweenSession(key3, q3).first != 0) { std::printf("between-session with one keyframe\n"); return 1; }
        std::printf("detectLoopClosureIDBetweenSession ok\n");
    }
    // cen2018features / cen2019features as free functions on a small synthetic polar image
    {
        const int R = 16, C = 640;
        std::vector<float> fft((size_t)R * C);
        std::uniform_int_distribution<int> noise(0, 23), hot(0, 39), lvl(120, 255);
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) fft[(size_t)i * C + j] = j < 42 ? 0.f : (hot(rng) == 0 ? lvl(rng) : noise(rng)) / 255.0f;
        MiniMatrix t18, t19;
        nrs::cen2018features(fft.data(), R, C, 3.0f, 17, 58, t18, ctx);
        nrs::cen2019features(fft.data(), R, C, 200, 58, t19, ctx);
        std::printf("cen2018: %d keypoints, cen2019: %d keypoints\n", (int)t18.cols(), (int)t19.cols());
        if (t18.cols() < 10 || t19.cols() < 10) return 1;
    }
    std::printf("adapters ok\n");
    }
    nrs_ctx_destroy(ctx);
    return rc;
}
