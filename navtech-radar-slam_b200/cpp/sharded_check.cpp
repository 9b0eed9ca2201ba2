// sharded_check.cpp -- the C++ multi-GPU host: one process per GPU, ShardedSCManagerT over a keyframe-sharded
// database (nrs_adapters.hpp), against an unsharded SCManagerT on the same keyframes.
//     sharded_check <rank> <world> <rendezvous dir> [nccl]
// Every rank feeds the same 120 clouds; detectLoopClosureID and searchBatch must return, on every rank, exactly
// what the single-GPU manager returns.  "nccl" forces the two-all-reduce exchange (sc_nccl.cu) instead of the
// in-kernel peer-memory one.  Exit code 0 = ok, 77 = no GPU / fewer GPUs than ranks.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "nrs_adapters.hpp"

struct MiniMatrix {
    MiniMatrix() = default;
    MiniMatrix(long r, long c) : r_(r), c_(c), d_((size_t)r * c, 0.0) {}
    double *data() { return d_.data(); }
    const double *data() const { return d_.data(); }
    long rows() const { return r_; }
    long cols() const { return c_; }
    void resize(long r, long c) { r_ = r; c_ = c; d_.assign((size_t)r * c, 0.0); }
    double &operator()(long i, long j) { return d_[(size_t)j * r_ + i]; }
    double operator()(long i, long j) const { return d_[(size_t)j * r_ + i]; }
    long r_ = 0, c_ = 0;
    std::vector<double> d_;
};
struct MiniPoint { float x, y, z, intensity; };
struct MiniCloud { std::vector<MiniPoint> points; };

int main(int argc, char **argv)
{
    if (argc < 4) { std::printf("usage: sharded_check <rank> <world> <dir> [nccl]\n"); return 2; }
    const int rank = std::atoi(argv[1]), world = std::atoi(argv[2]);
    const bool force_nccl = argc > 4 && !std::strcmp(argv[4], "nccl");
    nrs_ctx *ctx = nullptr;
    if (nrs_ctx_create(rank, &ctx) != NRS_OK) { std::printf("rank %d: no GPU %d: %s\n", rank, rank, nrs_last_error()); return 77; }
    int rc = 0;
    try {
        nrs::FileAllGather ag(argv[3], world, rank);
        nrs::ShardedSCManagerT<MiniMatrix, MiniCloud> sharded(world, rank, std::ref(ag), ctx, force_nccl);
        nrs::SCManagerT<MiniMatrix, MiniCloud> single(ctx);
        sharded.setSCdistThres(0.45);
        single.setSCdistThres(0.45);
        std::mt19937 rng(7);
        std::uniform_real_distribution<float> U(-70.f, 70.f);
        std::vector<MiniCloud> clouds(12);
        for (auto &c : clouds) for (int i = 0; i < 500; ++i) c.points.push_back({U(rng), U(rng), 1.0f, 0.f});
        int loops = 0;
        for (int k = 0; k < 120; ++k) {
            MiniCloud c = clouds[(k * 5) % 12];
            sharded.makeAndSaveScancontextAndKeys(c);
            single.makeAndSaveScancontextAndKeys(c);
            const auto a = sharded.detectLoopClosureID(), b = single.detectLoopClosureID();
            if (a != b) { std::printf("rank %d: keyframe %d: sharded (%d, %g) != single (%d, %g)\n", rank, k, a.first, a.second, b.first, b.second); rc = 1; break; }
            loops += a.first >= 0;
        }
        if (!rc) {
            std::vector<double> qs;
            for (int k = 0; k < 37; ++k) qs.insert(qs.end(), single.polarcontexts_[(k * 3) % 120].data(), single.polarcontexts_[(k * 3) % 120].data() + NRS_SC_DESC);
            const auto got = sharded.searchBatch(qs, 3, 117);
            for (int k = 0; k < 37 && !rc; ++k) {
                MiniMatrix q = single.polarcontexts_[(k * 3) % 120];
                const nrs_sc_result e = single.searchRange(q, 3, 117);        // the single-GPU answer, exhaustive over [3, 117)
                if (std::memcmp(&e.dist, &got[k].dist, 8) || e.idx != got[k].idx || e.shift != got[k].shift) {
                    std::printf("rank %d: batch query %d: (%.17g, %d, %d) != (%.17g, %d, %d)\n", rank, k, got[k].dist, got[k].idx, got[k].shift, e.dist, e.idx, e.shift);
                    rc = 1;
                }
            }
        }
        if (!rc) std::printf("rank %d/%d: sharded manager ok (%s exchange, %d local keyframes, %d loops)\n", rank, world, sharded.exchange(), sharded.localSize(), loops);
        if (!rc && loops < 20) { std::printf("rank %d: too few loops\n", rank); rc = 1; }
    } catch (const std::exception &e) {
        std::printf("rank %d: exception: %s\n", rank, e.what());
        rc = 1;
    }
    nrs_ctx_destroy(ctx);
    return rc;
}
