// frontend.cuh -- shared state of the ORORA front-end kernels (cen2019, polar->Cartesian, ORB,
// matcher).  One FrontendState hangs off a context per dataset parameter set; one nrs_frame holds
// the device-resident products of RadarLoader::extractFeature for one scan.
#pragma once
#include <vector>

#include "common.cuh"

namespace nrs {

constexpr int FE_MAX_AZ = 1024;          // azimuth rows supported (Navtech: 400)
constexpr int FE_KP_CAP = 16384;         // >= max_points (10000): one keypoint per marked region at most
constexpr int FE_PATTERN = 512;          // rBRIEF sample points (256 pairs)
constexpr int FE_SUB_CAP = 65536;        // cen2019: capacity of the top-candidate subset (>= 4 x max_points regions)

struct FrameCounts {
    int   n_cand;        // pixels with h > mean_h
    int   j_star;        // last candidate rank consumed by the marking loop
    int   l_regions;     // regions counted (<= max_points)
    int   K0;            // cen2019 keypoints (polar)
    int   Kv;            // after voxelisation
    int   K;             // after the BEV border filter (feat_in_cart_ columns)
    int   Kd;            // keypoints cv::ORB keeps (runByImageBorder on rounded coords) == descriptor rows
    float maxg;
    float mean;
    float mean_h;
    float cart_max;
    int   overflow;      // a capacity was exceeded (reported as NRS_ECAPACITY)
    int   n_sub;         // top-candidate subset: pixels with h >= the carried threshold (0 in full-sort mode)
    int   n_sorted;      // candidates in the sorted list the marking loop walked (== n_cand in full-sort mode)
    int   need_full;     // the subset could not prove the marking loop's outcome: redo with the full sort
    unsigned long long t_tail[8];   // %globaltimer stamps of fe_tail_kernel's phases (diagnostics: nrs_frontend_tail_times)
    unsigned int ticket[3];   // last-block tickets of fe_stats_kernel / fe_h_kernel / fe_incr_kernel (zeroed with the counts)
};

// ---- bit rows: the per-azimuth run scans (cen2019's marked regions, cen2018's peaks) first turn the azimuth's
// columns into one bit per column in shared memory (coalesced, independent loads) and then find run starts, run
// ends and neighbour tests with word operations instead of walking global memory column by column ----
constexpr int FE_ROW_WORDS = 256;        // => range_bins <= 8192
// first clear bit at or after position p (32 * nwords when there is none)
__device__ __forceinline__ int fe_first_zero(const unsigned int *mw, int nwords, int p)
{
    int w = p >> 5;
    unsigned int z = ~mw[w] & (0xffffffffu << (p & 31));
    while (!z) {
        if (++w >= nwords) return nwords * 32;
        z = ~mw[w];
    }
    return w * 32 + __ffs(z) - 1;
}
// any set bit in [lo, hi]
__device__ __forceinline__ bool fe_any_set(const unsigned int *aw, int lo, int hi)
{
    const int w0 = lo >> 5, w1 = hi >> 5;
    const unsigned int first = 0xffffffffu << (lo & 31), last = 0xffffffffu >> (31 - (hi & 31));
    if (w0 == w1) return (aw[w0] & first & last) != 0u;
    if (aw[w0] & first) return true;
    for (int w = w0 + 1; w < w1; ++w)
        if (aw[w]) return true;
    return (aw[w1] & last) != 0u;
}
// exclusive prefix of one count per thread over a 256-thread block (total in `total`); wsum: 8 ints of shared memory
__device__ __forceinline__ int fe_block256_exscan(int c, int *wsum, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int before = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const int v = wsum[w]; if (w < warp) before += v; total += v; }
    return before + inc - c;
}

// exclusive prefix of one value per thread over a 1024-thread block, total in `total`: warp scans by shuffles, the 32
// warp totals through `wtot` (32 entries of shared memory), every warp scans those itself -- one barrier, ~30
// instructions per thread (cub::BlockScan's warp-scans variant walks the 32 warp aggregates one by one in every
// thread).  The caller keeps a barrier between two calls that share `wtot` (every call site has one anyway).
template <class T>
__device__ __forceinline__ T fe_block1024_exscan(T v, T *wtot, T &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    const T wt = wtot[lane];
    T winc = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T up = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += up;
    }
    total = __shfl_sync(0xffffffffu, winc, 31);
    return __shfl_sync(0xffffffffu, winc - wt, warp) + inc - v;
}

struct SubPair { unsigned int e; float h; };   // cen2019 top-candidate subset: (pixel, h) as the compaction emits them

// per-frame azimuth tables, device layout == pinned host layout: ONE host-to-device copy per frame
struct AzTables {
    double az[FE_MAX_AZ];                // azimuths (R)
    double azs[FE_MAX_AZ];               // sorted distinct azimuth values
    double cossin[2 * FE_MAX_AZ];        // host libm cos / sin of each azimuth
    long long times[FE_MAX_AZ];          // row stamps
    int azi[FE_MAX_AZ];                  // lowest original index of each sorted distinct value
};

struct FrontendState {
    nrs_extract_params p;
    // static maps of polar_to_cartesian_radar_img (radar_utils.cpp:257-291), host libm, uploaded once
    float *d_theta = nullptr;            // W*W, atan2f(y, x) wrapped to [0, 2pi)
    int   *d_sx = nullptr;               // W*W, cvRound(range * 32)
    // rBRIEF sample offsets for angle = -1 deg (cv::ORB with supplied keypoints), 512 x (dx, dy)
    signed char *d_pattern = nullptr;
    // big per-frame scratch (reused by every frame of this context)
    unsigned char *d_raw = nullptr;      // R * pitch
    float *d_fft = nullptr;              // R * C
    float *d_h = nullptr;                // R * C
    float *d_keys[2] = {nullptr, nullptr};       // sort keys (double buffer)
    unsigned int *d_vals[2] = {nullptr, nullptr};// sort values
    unsigned int *d_claim = nullptr;     // R * C
    unsigned int *d_incr = nullptr;      // R * C (flags then inclusive scan)
    double *d_partial = nullptr;         // block partial sums
    float *d_cart = nullptr;             // W * W float image
    float *d_rowbuf = nullptr;           // W * W blur row pass
    unsigned char *d_blur = nullptr;     // W * W blurred u8
    void *d_cub = nullptr;               // CUB temp storage
    size_t cub_bytes = 0;
    // per-frame small tables
    AzTables *d_tables = nullptr, *h_tables = nullptr;   // h_tables pinned; the five pointers below alias d_tables
    double *d_az = nullptr;              // azimuths (R)
    double *d_azsort = nullptr;          // sorted distinct azimuth values
    int    *d_azidx = nullptr;           // their (lowest) original index
    double *d_cossin = nullptr;          // 2 * R
    long long *d_times = nullptr;        // R
    FrameCounts *d_counts = nullptr;
    FrameCounts *h_counts = nullptr;     // pinned
    int *h_n = nullptr;                  // pinned: the fused pair call's match count
    // cen2019 top-candidate subset: threshold on h carried from frame to frame (device), valid once a frame ran
    float *d_thresh = nullptr;
    bool   have_thresh = false;
    // the extraction pipeline as CUDA graphs: ~35 launches on two streams replayed by one cudaGraphLaunch.  One
    // executable graph per (frame block, extractor mode, azimuth count); node parameters are all device pointers of
    // this state or of the frame block, so a graph stays valid while both live
    struct ExtractGraph { unsigned char *block; int mode; int n_sorted; nrs_extract_params p; cudaGraphExec_t exec; int nodes; };
    std::vector<ExtractGraph> graphs;
    bool graphs_off = false;             // set when a capture failed once: plain stream launches from then on
    // cen2018's Gaussian taps (host libm) live here, not on the stack: the upload is a node of the captured graph and
    // reads this buffer at every replay
    float h_taps[256] = {};
    int   taps_sigma = -1;
    // voxel scratch
    float *d_pts32 = nullptr;            // 2 * KP_CAP float x,y
    SubPair *d_sel = nullptr;            // R x C: the compacted candidate subset (at most every pixel)
    float *d_lut = nullptr;              // 256: (float)(b / 255.0), load_radar's sample values (radar_utils.cpp:117-126)
    int *d_stage = nullptr;              // R x C: one staging row of keypoint range bins per azimuth
    int   *d_rowcount = nullptr;         // per-row keypoint counts / offsets
    int    R = 0, C = 0, pitch = 0;
    // the Cartesian-image branch of extractFeature (remap, 8-bit conversion, ORB's blur) does not depend on
    // the keypoints: it runs on a second stream beside cen2019 and joins in front of the descriptor kernel
    cudaStream_t st_img = nullptr;
    cudaEvent_t  ev_fork = nullptr, ev_join = nullptr;
};

}  // namespace nrs

struct nrs_frame {
    nrs_ctx *ctx = nullptr;
    unsigned char *d_block = nullptr;    // one allocation holding everything below
    size_t block_bytes = 0;
    // products of extractFeature (radar_loader.hpp:74-141)
    int    *d_polar = nullptr;           // 2 * KP_CAP (azimuth row, range bin)  feat_in_polar_
    double *d_cart_pre = nullptr;        // 2 * KP_CAP x,y before voxelisation
    long long *d_t_pre = nullptr;        // KP_CAP stamps before voxelisation
    double *d_feat = nullptr;            // 2 * KP_CAP feat_in_cart_ (x, y) after voxel + BEV
    long long *d_t = nullptr;            // KP_CAP  t_
    float  *d_kp = nullptr;              // 2 * KP_CAP keypoints (u, v) from convert_to_bev
    float  *d_kp_orb = nullptr;          // 2 * KP_CAP keypoints cv::ORB::compute keeps (kp_ after :126)
    unsigned char *d_desc = nullptr;     // KP_CAP * 32
    unsigned char *d_img = nullptr;      // W * W u8 Cartesian image  img_
    nrs::FrameCounts counts;
    int W = 0;
    long long t_criteria = 0;
    long long t_min = 0, t_max = 0;
};
