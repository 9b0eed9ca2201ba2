// frontend_img.cu -- polar -> Cartesian image (cv::remap restated), u8 conversion, ORB-style rBRIEF
// descriptors and the brute-force Hamming matcher on sm_100a, plus the front-end entry points of
// include/nrs_b200.h.
//
// Reference: outlier-robust-radar-odometry/src/radar_utils.cpp:227-316
// (get_azimuth_index, polar_to_cartesian_radar_img), include/radar_loader.hpp:69-71,126
// (cv::ORB::compute with supplied keypoints, patchSize 21), :143-205 (matchFeatures).
//
// Third-party arithmetic restated from OpenCV (verified bit-exact against cv2 4.13 by the oracle
// tests, tests/test_oracle_frontend.py):
//   remap INTER_LINEAR/f32   coordinates quantised to 1/32 (cvRound(map*32)), weights (1-fy)(1-fx)..
//                            in f32, taps summed left to right without FMA, constant-0 border.
//   convertTo(CV_8U, a)      saturate(rint(v * (float)a)).
//   ORB's GaussianBlur       7x7 sigma 2 on a sub-matrix => sepFilter2D float path: row pass
//                            k0*S0 then FMA chain, column pass k3*c then FMA of symmetric pairs,
//                            rounded to u8 (NOT the fixed-point path a full-image GaussianBlur takes).
//   rBRIEF                   patchSize 21 => 512 random sample points from cv::RNG(0x34985739),
//                            keypoint angle -1 degree (supplied keypoints carry no orientation).
//   BFMatcher(HAMMING) knn2  ties -> lowest train index.
// This file is compiled with --fmad=false; every fused multiply-add OpenCV issues is written as an
// explicit fmaf().
#include <cub/cub.cuh>
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>

#include "frontend.cuh"

namespace nrs {

int fe_cen2019_device(nrs_ctx *ctx, FrontendState &S, int R, int C, int max_points, int min_range, int *d_polar,
                      bool reset_counts = true, bool force_full = false, bool stats_done = false);
int fe_cen2018_taps_host(int sigma_gauss, float *taps);
int fe_cen2018_device(nrs_ctx *ctx, FrontendState &S, int R, int C, float zq, int sigma_gauss, int min_range,
                      int *d_polar, bool reset_counts, bool keep_y);
int fe_points_voxel_bev_device(nrs_ctx *ctx, FrontendState &S, nrs_frame *f);
int fe_decode_device(nrs_ctx *ctx, FrontendState &S, int R, int C, int pitch, bool with_stats);
size_t fe_cub_bytes(size_t N);

// cv::getGaussianKernel(7, 2, CV_32F) of OpenCV 4.13 (bit patterns; bit-exact softdouble evaluation)
__constant__ unsigned int c_gauss_bits[7] = {1032826801u, 1040595070u, 1044597305u, 1046301408u,
                                             1044597305u, 1040595070u, 1032826801u};

// ------------------------------------------------------------------------------------------------
// K4: polar_to_cartesian_radar_img
// ------------------------------------------------------------------------------------------------
// get_azimuth_index (radar_utils.cpp:227-250) by binary search over the distinct sorted azimuths;
// identical to the linear scan: nearest value, first index on ties (strict '<' while scanning up).
__device__ __forceinline__ double azimuth_index(const double *__restrict__ az, const double *__restrict__ azs,
                                                const int *__restrict__ azi, int n_sorted, int M,
                                                double azimuth)
{
    int lo = 0, hi = n_sorted;                    // upper_bound: first sorted value > azimuth
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (azs[mid] > azimuth) hi = mid; else lo = mid + 1;
    }
    int closest_i = 0;
    double mind = 1000;
    bool have = false;
    if (lo - 1 >= 0) {
        const double d = fabs(azs[lo - 1] - azimuth);
        if (d < mind) { mind = d; closest_i = azi[lo - 1]; have = true; }
    }
    if (lo < n_sorted) {
        const double d = fabs(azs[lo] - azimuth);
        if (d < mind || (have && d == mind && azi[lo] < closest_i)) { mind = d; closest_i = azi[lo]; }
    }
    double closest = closest_i;
    if (az[closest_i] < azimuth) {
        double delta = 0;
        if (closest < M - 1) delta = (azimuth - az[closest_i]) / (az[closest_i + 1] - az[closest_i]);
        closest += delta;
    } else if (az[closest_i] > azimuth) {
        double delta = 0;
        if (closest > 0) delta = (az[closest_i] - azimuth) / (az[closest_i] - az[closest_i - 1]);
        closest -= delta;
    }
    return closest;
}

__global__ void __launch_bounds__(256)
fe_remap_kernel(const float *__restrict__ fft, int R, int C, const float *__restrict__ theta,
                const int *__restrict__ sxq, const double *__restrict__ az, const double *__restrict__ azs,
                const int *__restrict__ azi, int n_sorted, int cir204, int interpolate, int W,
                float *__restrict__ cart, int *__restrict__ max_bits)
{
    __shared__ float smax[8];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    float out = 0.0f;
    if (e < W * W) {
        const float th = theta[e];
        float angle;
        if (cir204) angle = (float)azimuth_index(az, azs, azi, n_sorted, R, (double)th);     // :293
        else        angle = (float)(((double)th - az[0]) / (az[1] - az[0]));                  // :295
        if (interpolate) angle = angle + 1.0f;                                                  // :308
        const int sx = sxq[e];
        const int sy = __float2int_rn(angle * 32.0f);
        const int ix = sx >> 5, iy = sy >> 5;
        const float fx = (float)(sx & 31) * 0.03125f, fy = (float)(sy & 31) * 0.03125f;
        const int rows = interpolate ? R + 2 : R;
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int x = ix + (t & 1), y = iy + (t >> 1);
            float s = 0.0f;                                                                    // BORDER_CONSTANT 0
            if (x >= 0 && x < C && y >= 0 && y < rows) {
                int fr = y;
                if (interpolate) fr = (y == 0) ? R - 1 : (y == R + 1) ? 0 : y - 1;             // :299-307
                s = fft[(size_t)fr * C + x];
            }
            v[t] = s;
        }
        const float w0 = (1.0f - fy) * (1.0f - fx), w1 = (1.0f - fy) * fx, w2 = fy * (1.0f - fx), w3 = fy * fx;
        out = ((v[0] * w0 + v[1] * w1) + v[2] * w2) + v[3] * w3;
        cart[e] = out;
    }
    float m = out;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmaxf(m, smax[w]);
        atomicMax(max_bits, __float_as_int(m));
    }
}

// cart_img.convertTo(CV_8UC1, 255.0 / max) (:311-315)
__global__ void fe_to_u8_kernel(const float *__restrict__ cart, int n, const FrameCounts *__restrict__ cnt,
                                unsigned char *__restrict__ img)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const float alpha = (float)(255.0 / (double)cnt->cart_max);
    const float v = rintf(cart[e] * alpha);
    img[e] = (unsigned char)(v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v));   // NaN (max == 0) -> 0
}

// ------------------------------------------------------------------------------------------------
// K5: ORB descriptors
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int i, int n)
{
    if (i < 0) return -i;
    if (i >= n) return 2 * n - 2 - i;
    return i;
}

__global__ void fe_blur_row_kernel(const unsigned char *__restrict__ img, int W, float *__restrict__ rowbuf)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const unsigned char *row = img + (size_t)y * W;
    float s = __uint_as_float(c_gauss_bits[0]) * (float)row[reflect101(x - 3, W)];
#pragma unroll
    for (int t = 1; t < 7; ++t) s = fmaf(__uint_as_float(c_gauss_bits[t]), (float)row[reflect101(x - 3 + t, W)], s);
    rowbuf[(size_t)y * W + x] = s;
}

__global__ void fe_blur_col_kernel(const float *__restrict__ rowbuf, int W, unsigned char *__restrict__ blur)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    float s = __uint_as_float(c_gauss_bits[3]) * rowbuf[(size_t)y * W + x];
#pragma unroll
    for (int t = 1; t <= 3; ++t) {
        const float pair = rowbuf[(size_t)reflect101(y + t, W) * W + x] + rowbuf[(size_t)reflect101(y - t, W) * W + x];
        s = fmaf(__uint_as_float(c_gauss_bits[3 + t]), pair, s);
    }
    const float v = rintf(s);
    blur[(size_t)y * W + x] = (unsigned char)(v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v));
}

// one thread per (keypoint, descriptor byte): 8 binary tests (orb.cpp computeOrbDescriptors, WTA_K 2)
__global__ void fe_brief_kernel(const unsigned char *__restrict__ blur, int W, const float *__restrict__ kp,
                                const FrameCounts *__restrict__ cnt, const signed char *__restrict__ pattern,
                                unsigned char *__restrict__ desc)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = gid >> 5, byte = gid & 31;
    if (k >= cnt->Kd) return;
    const int cx = __float2int_rn(kp[2 * k]), cy = __float2int_rn(kp[2 * k + 1]);
    const unsigned char *center = blur + (size_t)cy * W + cx;
    unsigned int val = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const int i0 = 2 * (8 * byte + b), i1 = i0 + 1;
        const int t0 = center[(int)pattern[2 * i0 + 1] * W + (int)pattern[2 * i0]];
        const int t1 = center[(int)pattern[2 * i1 + 1] * W + (int)pattern[2 * i1]];
        val |= (unsigned int)(t0 < t1) << b;
    }
    desc[(size_t)k * 32 + byte] = (unsigned char)val;
}

// ------------------------------------------------------------------------------------------------
// K6: BFMatcher(NORM_HAMMING).knnMatch(prev, curr, 2) + ratio test, one warp per query descriptor
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool lex_less(int d1, int i1, int d2, int i2) { return d1 < d2 || (d1 == d2 && i1 < i2); }

constexpr int MATCH_PARTS = 4;           // warps that share one query descriptor (each scans a quarter of the train set)
constexpr int MATCH_QPB = 8 / MATCH_PARTS;   // queries per 256-thread block

__global__ void __launch_bounds__(256)
fe_match_kernel(const unsigned char *__restrict__ qdesc, int nq, const unsigned char *__restrict__ tdesc, int nt_host,
                const FrameCounts *__restrict__ tcnt, float nndr, int *__restrict__ best_train,
                unsigned char *__restrict__ good)
{
    // BFMatcher(NORM_HAMMING).knnMatch(k = 2) of one query row: the two smallest (distance, train index) pairs.
    // Four warps scan a quarter of the train descriptors each (the scan is a chain of dependent loads: shorter
    // chains, four times the warps in flight), lane 0..3 of the first warp then merge the four partial answers.
    __shared__ int4 part_res[MATCH_QPB][MATCH_PARTS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ql = warp / MATCH_PARTS, part = warp % MATCH_PARTS;
    const int q = blockIdx.x * MATCH_QPB + ql;
    const int nt = tcnt ? tcnt->Kd : nt_host;            // fused pair call: the train frame's counts are still on the device
    if (q < nq) {
        const uint4 *qp = reinterpret_cast<const uint4 *>(qdesc + (size_t)q * 32);
        const uint4 qa = qp[0], qb = qp[1];
        const int chunk = (((nt + MATCH_PARTS - 1) / MATCH_PARTS) + 31) & ~31;
        const int t0 = part * chunk, t1 = min(nt, t0 + chunk);
        int d0 = 0x7fffffff, i0 = 0x7fffffff, d1 = 0x7fffffff, i1 = 0x7fffffff;
        for (int t = t0 + lane; t < t1; t += 32) {
            const uint4 *tp = reinterpret_cast<const uint4 *>(tdesc + (size_t)t * 32);
            const uint4 ta = tp[0], tb = tp[1];
            const int d = __popc(qa.x ^ ta.x) + __popc(qa.y ^ ta.y) + __popc(qa.z ^ ta.z) + __popc(qa.w ^ ta.w) +
                          __popc(qb.x ^ tb.x) + __popc(qb.y ^ tb.y) + __popc(qb.z ^ tb.z) + __popc(qb.w ^ tb.w);
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = t; }
            else if (d < d1) { d1 = d; i1 = t; }
        }
        // the part's best, then its second = best of (every lane's best except the winner's) U (the winner's second)
        int bd = d0, bi = i0;
        for (int o = 16; o > 0; o >>= 1) {
            const int od = __shfl_xor_sync(0xffffffffu, bd, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (lex_less(od, oi, bd, bi)) { bd = od; bi = oi; }
        }
        const bool winner = (i0 == bi);
        int sd = winner ? d1 : d0, si = winner ? i1 : i0;
        for (int o = 16; o > 0; o >>= 1) {
            const int od = __shfl_xor_sync(0xffffffffu, sd, o), oi = __shfl_xor_sync(0xffffffffu, si, o);
            if (lex_less(od, oi, sd, si)) { sd = od; si = oi; }
        }
        if (lane == 0) part_res[ql][part] = make_int4(bd, bi, sd, si);
    }
    __syncthreads();
    if (q < nq && part == 0 && lane < MATCH_PARTS) {
        const int4 r = part_res[ql][lane];
        const unsigned int m = (1u << MATCH_PARTS) - 1u;
        int bd = r.x, bi = r.y;
        for (int o = MATCH_PARTS / 2; o > 0; o >>= 1) {
            const int od = __shfl_xor_sync(m, bd, o), oi = __shfl_xor_sync(m, bi, o);
            if (lex_less(od, oi, bd, bi)) { bd = od; bi = oi; }
        }
        const bool winner = (r.y == bi);
        int sd = winner ? r.z : r.x, si = winner ? r.w : r.y;
        for (int o = MATCH_PARTS / 2; o > 0; o >>= 1) {
            const int od = __shfl_xor_sync(m, sd, o), oi = __shfl_xor_sync(m, si, o);
            if (lex_less(od, oi, sd, si)) { sd = od; si = oi; }
        }
        if (lane == 0) {
            best_train[q] = bi;
            // radar_loader.hpp:159-167; rows with fewer than two neighbours are skipped (the reference
            // dereferences knn_matches[j][1] -- undefined behaviour -- when curr has a single keypoint)
            good[q] = (nt >= 2) && ((float)bd < nndr * (float)sd);
        }
    }
}

__global__ void __launch_bounds__(1024)
fe_match_gather_kernel(const int *__restrict__ best_train, const unsigned char *__restrict__ good, int nq,
                       const double *__restrict__ pfeat, const double *__restrict__ cfeat,
                       const long long *__restrict__ pt, const long long *__restrict__ ct, int cap,
                       int *__restrict__ qidx, int *__restrict__ tidx, double *__restrict__ prev_xy,
                       double *__restrict__ curr_xy, long long *__restrict__ prev_t,
                       long long *__restrict__ curr_t, int *__restrict__ n_out)
{
    typedef cub::BlockScan<int, 1024> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ int carry;
    const int t = threadIdx.x;
    if (t == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nq; base += 1024) {
        const int q = base + t;
        const bool g = q < nq && good[q];
        int ex, total;
        BS(tmp).ExclusiveSum(g ? 1 : 0, ex, total);
        const int c0 = carry;
        if (g && c0 + ex < cap) {
            const int j = c0 + ex, tr = best_train[q];
            qidx[j] = q; tidx[j] = tr;
            prev_xy[2 * j] = pfeat[2 * q];  prev_xy[2 * j + 1] = pfeat[2 * q + 1];      // :181
            curr_xy[2 * j] = cfeat[2 * tr]; curr_xy[2 * j + 1] = cfeat[2 * tr + 1];    // :182
            prev_t[j] = pt[q];                                                          // :184-185
            curr_t[j] = ct[tr];
        }
        __syncthreads();
        if (t == 0) carry = c0 + total;
        __syncthreads();
    }
    if (t == 0) *n_out = carry;
}

// ------------------------------------------------------------------------------------------------
// static state: maps (host libm, like the reference's own build), rBRIEF pattern, scratch
// ------------------------------------------------------------------------------------------------
struct CvRNG {          // cv::RNG: multiply-with-carry (core/operations.hpp RNG::next, uniform(int,int))
    uint64_t state;
    explicit CvRNG(uint64_t s) : state(s) {}
    unsigned next() { state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32); return (unsigned)state; }
    int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

static void frontend_free(void *p)
{
    FrontendState *S = static_cast<FrontendState *>(p);
    void *dev[] = {S->d_theta, S->d_sx, S->d_pattern, S->d_raw, S->d_fft, S->d_h, S->d_keys[0], S->d_keys[1],
                   S->d_vals[0], S->d_vals[1], S->d_claim, S->d_incr, S->d_partial, S->d_cart, S->d_rowbuf,
                   S->d_blur, S->d_cub, S->d_tables, S->d_counts,
                   S->d_pts32, S->d_rowcount, S->d_thresh, S->d_stage, S->d_sel, S->d_lut};
    for (void *d : dev) if (d) cudaFree(d);
    for (auto &g : S->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    S->graphs.clear();
    if (S->h_counts) cudaFreeHost(S->h_counts);
    if (S->h_n) cudaFreeHost(S->h_n);
    if (S->h_tables) cudaFreeHost(S->h_tables);
    if (S->st_img) { cudaStreamSynchronize(S->st_img); cudaStreamDestroy(S->st_img); }
    if (S->ev_fork) cudaEventDestroy(S->ev_fork);
    if (S->ev_join) cudaEventDestroy(S->ev_join);
    delete S;
}

template <class T> static int dmalloc(T **p, size_t n)
{
    NRS_CUDA(cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)));
    return NRS_OK;
}

static bool same_params(const nrs_extract_params &a, const nrs_extract_params &b)
{
    return a.navtech_version == b.navtech_version && a.range_bins == b.range_bins &&
           a.radar_resolution == b.radar_resolution && a.cart_resolution == b.cart_resolution &&
           a.cart_pixel_width == b.cart_pixel_width && a.patch_size == b.patch_size;
}

// (re)build the static state for parameter set p, image geometry R x C x pitch
static int frontend_state(nrs_ctx *ctx, const nrs_extract_params *p, int R, int C, int pitch, FrontendState **out)
{
    FrontendState *S = static_cast<FrontendState *>(ctx->frontend);
    if (S && same_params(S->p, *p) && S->R == R && S->C == C && S->pitch >= pitch) {
        S->p = *p;
        *out = S;
        return NRS_OK;
    }
    if (S) { cudaStreamSynchronize(ctx->stream); frontend_free(S); ctx->frontend = nullptr; }
    S = new FrontendState();
    ctx->frontend = S;
    ctx->frontend_free = frontend_free;
    S->p = *p; S->R = R; S->C = C; S->pitch = pitch;
    NRS_CUDA(cudaStreamCreateWithFlags(&S->st_img, cudaStreamNonBlocking));
    NRS_CUDA(cudaEventCreateWithFlags(&S->ev_fork, cudaEventDisableTiming));
    NRS_CUDA(cudaEventCreateWithFlags(&S->ev_join, cudaEventDisableTiming));
    const int W = p->cart_pixel_width;
    const size_t N = (size_t)R * C, WW = (size_t)W * W;
    int rc = 0;
    rc |= dmalloc(&S->d_theta, WW); rc |= dmalloc(&S->d_sx, WW); rc |= dmalloc(&S->d_pattern, FE_PATTERN * 2);
    rc |= dmalloc(&S->d_raw, (size_t)R * pitch); rc |= dmalloc(&S->d_fft, N); rc |= dmalloc(&S->d_h, N);
    rc |= dmalloc(&S->d_keys[0], N); rc |= dmalloc(&S->d_keys[1], N); rc |= dmalloc(&S->d_vals[0], N);
    rc |= dmalloc(&S->d_vals[1], N); rc |= dmalloc(&S->d_claim, N); rc |= dmalloc(&S->d_incr, N);
    rc |= dmalloc(&S->d_partial, std::max((size_t)ctx->sm_count * 8, (size_t)R) + 8);
    rc |= dmalloc(&S->d_cart, WW); rc |= dmalloc(&S->d_rowbuf, WW); rc |= dmalloc(&S->d_blur, WW);
    S->cub_bytes = fe_cub_bytes(N);
    rc |= dmalloc((unsigned char **)&S->d_cub, S->cub_bytes);
    rc |= dmalloc(&S->d_tables, 1); rc |= dmalloc(&S->d_counts, 1);
    rc |= dmalloc(&S->d_pts32, 2 * FE_KP_CAP); rc |= dmalloc(&S->d_rowcount, FE_MAX_AZ);
    rc |= dmalloc(&S->d_thresh, 1); rc |= dmalloc(&S->d_stage, N); rc |= dmalloc(&S->d_sel, N); rc |= dmalloc(&S->d_lut, 256);
    if (rc) return rc;
    NRS_CUDA(cudaMallocHost((void **)&S->h_counts, sizeof(FrameCounts)));
    NRS_CUDA(cudaMallocHost((void **)&S->h_n, sizeof(int)));
    NRS_CUDA(cudaMallocHost((void **)&S->h_tables, sizeof(AzTables)));
    {
        float lut[256];
        for (int b = 0; b < 256; ++b) lut[b] = (float)((double)b / 255.0);
        NRS_CUDA(cudaMemcpy(S->d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
    }
    S->d_az = S->d_tables->az; S->d_azsort = S->d_tables->azs; S->d_azidx = S->d_tables->azi;      // address arithmetic only
    S->d_cossin = S->d_tables->cossin; S->d_times = S->d_tables->times;

    // ---- static maps (radar_utils.cpp:257-291), computed with the host's libm like the reference ----
    std::vector<float> theta(WW);
    std::vector<int> sxq(WW);
    const float cart_resolution = p->cart_resolution, radar_resolution = p->radar_resolution;
    float cart_min_range = (W / 2) * cart_resolution;
    if (W % 2 == 0) cart_min_range = (W / 2 - 0.5) * cart_resolution;
    for (int i = 0; i < W; ++i)
        for (int j = 0; j < W; ++j) {
            const float x = cart_min_range - i * cart_resolution;
            const float y = -1 * cart_min_range + j * cart_resolution;
            float r = (sqrt(pow((double)x, 2) + pow((double)y, 2)) - radar_resolution / 2) / radar_resolution;
            if (r < 0) r = 0;
            float th = atan2f(y, x);
            if (th < 0) th += 2 * M_PI;
            theta[(size_t)i * W + j] = th;
            sxq[(size_t)i * W + j] = (int)lrintf(r * 32.0f);        // cv::remap: cvRound(mapx * INTER_TAB_SIZE)
        }
    NRS_CUDA(cudaMemcpyAsync(S->d_theta, theta.data(), WW * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    NRS_CUDA(cudaMemcpyAsync(S->d_sx, sxq.data(), WW * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));

    // ---- rBRIEF sample offsets: ORB_Impl::detectAndCompute with patchSize != 31 draws 512 points
    //      from RNG(0x34985739) (makeRandomPattern); supplied keypoints keep angle = -1 degree ------
    signed char pat[FE_PATTERN * 2];
    {
        CvRNG rng(0x34985739);
        const int half = p->patch_size / 2;
        float angle = -1.0f;
        angle *= (float)(3.1415926535897932384626433832795 / 180.f);
        const float a = (float)cos((double)angle), b = (float)sin((double)angle);
        for (int i = 0; i < FE_PATTERN; ++i) {
            const int px = rng.uniform(-half, half + 1);
            const int py = rng.uniform(-half, half + 1);
            volatile float xa = (float)px * a, yb = (float)py * b, xb = (float)px * b, ya = (float)py * a;
            const float x = xa - yb, y = xb + ya;                    // no FMA contraction (volatile products)
            pat[2 * i] = (signed char)lrintf(x);
            pat[2 * i + 1] = (signed char)lrintf(y);
        }
    }
    NRS_CUDA(cudaMemcpyAsync(S->d_pattern, pat, sizeof(pat), cudaMemcpyHostToDevice, ctx->stream));
    NRS_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = S;
    return NRS_OK;
}

static int alloc_frame(nrs_ctx *ctx, int W, nrs_frame **out)
{
    nrs_frame *f = new nrs_frame();
    f->ctx = ctx;
    f->W = W;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_polar = take(2 * FE_KP_CAP * sizeof(int)), o_pre = take(2 * FE_KP_CAP * sizeof(double));
    const size_t o_tpre = take(FE_KP_CAP * sizeof(long long)), o_feat = take(2 * FE_KP_CAP * sizeof(double));
    const size_t o_t = take(FE_KP_CAP * sizeof(long long)), o_kp = take(2 * FE_KP_CAP * sizeof(float));
    const size_t o_kpo = take(2 * FE_KP_CAP * sizeof(float));
    const size_t o_desc = take((size_t)FE_KP_CAP * 32), o_img = take((size_t)W * W);
    // recycle a block of a destroyed frame: cudaMalloc / cudaFree synchronise and cost milliseconds
    for (size_t i = 0; i < ctx->frame_pool.size(); ++i)
        if (ctx->frame_pool[i].second >= off) {
            f->d_block = (unsigned char *)ctx->frame_pool[i].first;
            f->block_bytes = ctx->frame_pool[i].second;
            ctx->frame_pool.erase(ctx->frame_pool.begin() + i);
            break;
        }
    if (!f->d_block) {
        cudaError_t e = cudaMalloc((void **)&f->d_block, off);
        if (e != cudaSuccess) { delete f; set_error("frame allocation: %s", cudaGetErrorString(e)); return NRS_ENOMEM; }
        f->block_bytes = off;
    }
    unsigned char *b = f->d_block;
    f->d_polar = (int *)(b + o_polar); f->d_cart_pre = (double *)(b + o_pre); f->d_t_pre = (long long *)(b + o_tpre);
    f->d_feat = (double *)(b + o_feat); f->d_t = (long long *)(b + o_t); f->d_kp = (float *)(b + o_kp);
    f->d_desc = b + o_desc; f->d_img = b + o_img; f->d_kp_orb = (float *)(b + o_kpo);
    *out = f;
    return NRS_OK;
}

// called with ctx->mu held
static void frame_release_locked(nrs_frame *f)
{
    if (f->d_block) {
        if (f->ctx->frame_pool.size() < 8) f->ctx->frame_pool.emplace_back(f->d_block, f->block_bytes);
        else cudaFree(f->d_block);
    }
    delete f;
}

}  // namespace nrs

using namespace nrs;

extern "C" {

int nrs_extract_default_params(const char *dataset, nrs_extract_params *p)
{
    NRS_CHECK_ARG(dataset && p, "bad argument");
    memset(p, 0, sizeof(*p));
    if (!strcmp(dataset, "oxford")) {                 // radar_loader.hpp:48-54
        p->navtech_version = NRS_NAVTECH_CTS350; p->range_bins = 3768; p->radar_resolution = 0.0432f;
    } else if (!strcmp(dataset, "mulran")) {          // :55-61
        // load_radar gets CIR204 here (radar_loader.hpp:82-84: 3360 range bins), but extractFeature hands
        // polar_to_cartesian_radar_img no version at all (:105-107), i.e. its default CTS350: the node's Cartesian image
        // uses the uniform-step azimuth mapping (radar_utils.cpp:295) for MulRan too; get_azimuth_index (:293) is only
        // reached by a caller that sets NRS_NAVTECH_CIR204 here itself
        p->navtech_version = NRS_NAVTECH_CTS350; p->range_bins = 3360; p->radar_resolution = 0.05952f;
    } else {
        set_error("nrs_extract_default_params: dataset '%s' not implemented (radar_loader.hpp:62)", dataset);
        return NRS_EINVAL;
    }
    p->cart_resolution = 0.2592f; p->cart_pixel_width = 964; p->min_range = 58;
    p->max_points = 10000; p->patch_size = 21; p->use_voxelization = 1; p->interpolate_crossover = 1;
    p->voxel_size = 0.6; p->nndr = 0.8f;
    p->keypoint_extraction = 0; p->zq = 3.0f; p->sigma_gauss = 17;     // radar_loader.hpp:284-285
    return NRS_OK;
}

int nrs_cen2018(nrs_ctx *ctx, const float *fft, int R, int C, float zq, int sigma_gauss, int min_range,
                double *targets, int Kcap, int *K, float *y_out)
{
    NRS_CHECK_ARG(ctx && fft && K && R >= 2 && R <= FE_MAX_AZ && C >= 3 && C <= 32 * FE_ROW_WORDS && min_range >= 0 &&
                  (targets || Kcap == 0), "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    NRS_CUDA(cudaSetDevice(ctx->device));
    nrs_extract_params p;
    nrs_extract_default_params("mulran", &p);
    p.range_bins = C; p.min_range = min_range;
    FrontendState *S = nullptr;
    int rc = frontend_state(ctx, &p, R, C, C + 11, &S);
    if (rc) return rc;
    rc = ctx->scratch[4].reserve(2 * FE_KP_CAP * sizeof(int));
    if (rc) return rc;
    int *d_polar = ctx->scratch[4].as<int>();
    NRS_CUDA(cudaMemcpyAsync(S->d_fft, fft, (size_t)R * C * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    rc = fe_cen2018_device(ctx, *S, R, C, zq, sigma_gauss, min_range, d_polar, true, y_out != nullptr);
    if (rc) return rc;
    NRS_CUDA(cudaMemcpyAsync(S->h_counts, S->d_counts, sizeof(FrameCounts), cudaMemcpyDeviceToHost, ctx->stream));
    if (y_out) NRS_CUDA(cudaMemcpyAsync(y_out, S->d_keys[1], (size_t)R * C * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    NRS_CUDA(cudaStreamSynchronize(ctx->stream));
    const int k = S->h_counts->K0;
    *K = k;
    if (S->h_counts->overflow) { set_error("nrs_cen2018: more than %d keypoints", FE_KP_CAP); return NRS_ECAPACITY; }
    if (k > Kcap) { set_error("nrs_cen2018: K = %d exceeds Kcap = %d", k, Kcap); return NRS_ECAPACITY; }
    std::vector<int> hp(2 * (size_t)std::max(k, 1));
    if (k) NRS_CUDA(cudaMemcpy(hp.data(), d_polar, 2 * (size_t)k * sizeof(int), cudaMemcpyDeviceToHost));
    for (int c = 0; c < k; ++c) {
        targets[3 * c] = hp[2 * c]; targets[3 * c + 1] = hp[2 * c + 1]; targets[3 * c + 2] = 1.0;
    }
    return NRS_OK;
}

int nrs_cen2019(nrs_ctx *ctx, const float *fft, int R, int C, int max_points, int min_range,
                double *targets, int Kcap, int *K)
{
    NRS_CHECK_ARG(ctx && fft && K && R >= 2 && R <= FE_MAX_AZ && C >= 3 && C <= 32 * FE_ROW_WORDS && max_points > 0 && min_range >= 0 &&
                  (targets || Kcap == 0), "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    NRS_CUDA(cudaSetDevice(ctx->device));
    nrs_extract_params p;
    nrs_extract_default_params("mulran", &p);
    p.range_bins = C; p.max_points = max_points; p.min_range = min_range;
    FrontendState *S = nullptr;
    int rc = frontend_state(ctx, &p, R, C, C + 11, &S);
    if (rc) return rc;
    rc = ctx->scratch[4].reserve(2 * FE_KP_CAP * sizeof(int));
    if (rc) return rc;
    int *d_polar = ctx->scratch[4].as<int>();
    NRS_CUDA(cudaMemcpyAsync(S->d_fft, fft, (size_t)R * C * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    for (int attempt = 0; attempt < 2; ++attempt) {
        rc = fe_cen2019_device(ctx, *S, R, C, max_points, min_range, d_polar, true, attempt == 1);
        if (rc) return rc;
        NRS_CUDA(cudaMemcpyAsync(S->h_counts, S->d_counts, sizeof(FrameCounts), cudaMemcpyDeviceToHost, ctx->stream));
        NRS_CUDA(cudaStreamSynchronize(ctx->stream));
        if (!S->h_counts->need_full) break;          // the top-candidate subset vouched for the result (or the full sort ran)
    }
    const int k = S->h_counts->K0;
    *K = k;
    if (S->h_counts->overflow) { set_error("nrs_cen2019: more than %d keypoints", FE_KP_CAP); return NRS_ECAPACITY; }
    if (k > Kcap) { set_error("nrs_cen2019: K = %d exceeds Kcap = %d", k, Kcap); return NRS_ECAPACITY; }
    std::vector<int> hp(2 * (size_t)std::max(k, 1));
    if (k) NRS_CUDA(cudaMemcpy(hp.data(), d_polar, 2 * (size_t)k * sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < k; ++i) {                    // features.cpp:284-292: (azimuth, range, 1)
        targets[3 * i] = hp[2 * i]; targets[3 * i + 1] = hp[2 * i + 1]; targets[3 * i + 2] = 1.0;
    }
    return NRS_OK;
}

namespace {

// row headers (radar_utils.cpp:121-123) and the libm tables derived from them, uploaded with the raw rows
struct HostTables {
    int n_sorted = 0;
    long long tmin = 0, tmax = 0;
};

// cos / sin of the 5600 azimuths a Navtech encoder can report (radar_utils.cpp:117-123: enc * 2 pi / 5600), host
// libm like the reference, evaluated once
struct EncoderTable { double az[5600], c[5600], s[5600]; };
const EncoderTable &encoder_table()
{
    static const EncoderTable *tab = [] {
        EncoderTable *t = new EncoderTable();
        for (int e = 0; e < 5600; ++e) {
            t->az[e] = e * 2 * M_PI / double(5600);
            t->c[e] = cos(t->az[e]);
            t->s[e] = sin(t->az[e]);
        }
        return t;
    }();
    return *tab;
}

int extract_upload(nrs_ctx *ctx, FrontendState *S, const uint8_t *raw, int R, int row_pitch, HostTables &H)
{
    cudaStream_t st = ctx->stream;
    // the frame goes first: its 1.35 MB are on the wire while the host walks the row headers (one cache-cold line per
    // azimuth) and builds the tables
    NRS_CUDA(cudaMemcpyAsync(S->d_raw, raw, (size_t)R * row_pitch, cudaMemcpyHostToDevice, st));
    const EncoderTable &E = encoder_table();
    AzTables &T = *S->h_tables;          // pinned; the previous frame's copy finished before its call returned
    bool increasing = true;
    for (int i = 0; i < R; ++i) {
        const uint8_t *row = raw + (size_t)i * row_pitch;
        if (i + 8 < R) __builtin_prefetch(row + (size_t)8 * row_pitch);
        int64_t stamp; uint16_t enc;
        memcpy(&stamp, row, 8);
        memcpy(&enc, row + 8, 2);
        T.times[i] = stamp;
        if (enc < 5600) { T.az[i] = E.az[enc]; T.cossin[2 * i] = E.c[enc]; T.cossin[2 * i + 1] = E.s[enc]; }
        else {                           // not an encoder value the sensor produces; same expression all the same
            T.az[i] = enc * 2 * M_PI / double(5600);
            T.cossin[2 * i] = cos(T.az[i]);
            T.cossin[2 * i + 1] = sin(T.az[i]);
        }
        if (i && !(T.az[i] > T.az[i - 1])) increasing = false;
        if (i == 0 || stamp < H.tmin) H.tmin = stamp;
        if (i == 0 || stamp > H.tmax) H.tmax = stamp;
    }
    H.n_sorted = 0;
    if (increasing) {                    // one sweep without a wrap: already sorted and distinct
        for (int i = 0; i < R; ++i) { T.azs[i] = T.az[i]; T.azi[i] = i; }
        H.n_sorted = R;
    } else {
        int order[FE_MAX_AZ];
        for (int i = 0; i < R; ++i) order[i] = i;
        std::stable_sort(order, order + R, [&](int a, int b) { return T.az[a] < T.az[b]; });
        for (int k = 0; k < R; ++k) {
            const int i = order[k];
            if (H.n_sorted && T.azs[H.n_sorted - 1] == T.az[i]) continue;     // duplicates: keep the lowest index
            T.azs[H.n_sorted] = T.az[i]; T.azi[H.n_sorted] = i; ++H.n_sorted;
        }
    }
    NRS_CUDA(cudaMemcpyAsync(S->d_tables, S->h_tables, sizeof(AzTables), cudaMemcpyHostToDevice, st));
    return NRS_OK;
}

// the whole device pipeline of extractFeature for the uploaded frame, enqueued on the context stream (image
// branch on the second stream, joined before the descriptors); no host synchronisation
int extract_enqueue(nrs_ctx *ctx, FrontendState *S, nrs_frame *f, int R, int C, int row_pitch, const nrs_extract_params *p,
                    int n_sorted, bool force_full)
{
    cudaStream_t st = ctx->stream;
    const int W = p->cart_pixel_width;
    int rc;
    NRS_CUDA(cudaMemsetAsync(S->d_counts, 0, sizeof(FrameCounts), st));
    const bool cen2019 = p->keypoint_extraction != 1;
    if ((rc = fe_decode_device(ctx, *S, R, C, row_pitch, cen2019))) return rc;      // cen2019: sum(f), max(g) in the same pass
    // fork: Cartesian image branch (radar_utils.cpp:227-316 + ORB's blur) on the second stream
    cudaStream_t si = S->st_img;
    NRS_CUDA(cudaEventRecord(S->ev_fork, st));
    NRS_CUDA(cudaStreamWaitEvent(si, S->ev_fork, 0));
    fe_remap_kernel<<<(W * W + 255) / 256, 256, 0, si>>>(S->d_fft, R, C, S->d_theta, S->d_sx, S->d_az, S->d_azsort,
                                                        S->d_azidx, n_sorted, p->navtech_version == NRS_NAVTECH_CIR204,
                                                        p->interpolate_crossover, W, S->d_cart,
                                                        (int *)&S->d_counts->cart_max);
    NRS_LAUNCHED();
    fe_to_u8_kernel<<<(W * W + 255) / 256, 256, 0, si>>>(S->d_cart, W * W, S->d_counts, f->d_img);
    NRS_LAUNCHED();
    dim3 bg((W + 255) / 256, W);
    fe_blur_row_kernel<<<bg, 256, 0, si>>>(f->d_img, W, S->d_rowbuf);
    NRS_LAUNCHED();
    fe_blur_col_kernel<<<bg, 256, 0, si>>>(S->d_rowbuf, W, S->d_blur);
    NRS_LAUNCHED();
    NRS_CUDA(cudaEventRecord(S->ev_join, si));
    // keypoint branch on the context stream
    rc = (p->keypoint_extraction == 1)
             ? fe_cen2018_device(ctx, *S, R, C, p->zq, p->sigma_gauss, p->min_range, nullptr, false, false)
             : fe_cen2019_device(ctx, *S, R, C, p->max_points, p->min_range, nullptr, false, force_full, true);
    if (rc) { cudaStreamSynchronize(si); return rc; }
    if ((rc = fe_points_voxel_bev_device(ctx, *S, f))) { cudaStreamSynchronize(si); return rc; }
    NRS_CUDA(cudaStreamWaitEvent(st, S->ev_join, 0));           // join
    fe_brief_kernel<<<(FE_KP_CAP * 32 + 255) / 256, 256, 0, st>>>(S->d_blur, W, f->d_kp_orb, S->d_counts, S->d_pattern, f->d_desc);
    NRS_LAUNCHED();
    return NRS_OK;
}

// extract_enqueue through a cached CUDA graph (captured from the very same enqueue code on first use)
int extract_run(nrs_ctx *ctx, FrontendState *S, nrs_frame *f, int R, int C, int row_pitch, const nrs_extract_params *p,
                int n_sorted, bool force_full)
{
    static const bool env_off = getenv("NRS_NO_GRAPHS") != nullptr;
    if (S->graphs_off || env_off) return extract_enqueue(ctx, S, f, R, C, row_pitch, p, n_sorted, force_full);
    cudaStream_t st = ctx->stream;
    const int mode = p->keypoint_extraction == 1 ? 2 : ((S->have_thresh && !force_full) ? 1 : 0);
    if (mode == 2 && S->taps_sigma != p->sigma_gauss && p->sigma_gauss >= 1 && p->sigma_gauss * 3 <= 256) {
        // a cached cen2018 graph uploads S->h_taps at every replay: bring the taps back to this sigma first
        fe_cen2018_taps_host(p->sigma_gauss, S->h_taps);
        S->taps_sigma = p->sigma_gauss;
    }
    for (auto &g : S->graphs)
        if (g.block == f->d_block && g.mode == mode && g.n_sorted == n_sorted && !memcmp(&g.p, p, sizeof(*p))) {   // every by-value kernel argument
            NRS_CUDA(cudaGraphLaunch(g.exec, st));
            g_launches.fetch_add((uint64_t)g.nodes, std::memory_order_relaxed);
            if (mode != 2) S->have_thresh = true;
            return NRS_OK;
        }
    // capture: the launches below are recorded, not run
    const uint64_t launches_before = g_launches.load();
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        S->graphs_off = true;
        return extract_enqueue(ctx, S, f, R, C, row_pitch, p, n_sorted, force_full);
    }
    const bool had_thresh = S->have_thresh;
    const int rc = extract_enqueue(ctx, S, f, R, C, row_pitch, p, n_sorted, force_full);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    const int nodes = (int)(g_launches.load() - launches_before);      // this thread's launches were recorded, not run
    g_launches.fetch_sub((uint64_t)nodes, std::memory_order_relaxed);
    cudaGraphExec_t exec = nullptr;
    if (rc != NRS_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        S->graphs_off = true;                        // plain launches from now on (and for this frame)
        S->have_thresh = had_thresh;
        return extract_enqueue(ctx, S, f, R, C, row_pitch, p, n_sorted, force_full);
    }
    cudaGraphDestroy(graph);
    if (S->graphs.size() >= 12) {                    // frames recycle a handful of blocks; keep the cache small
        cudaGraphExecDestroy(S->graphs.front().exec);
        S->graphs.erase(S->graphs.begin());
    }
    S->graphs.push_back(FrontendState::ExtractGraph{f->d_block, mode, n_sorted, *p, exec, nodes});
    NRS_CUDA(cudaGraphLaunch(exec, st));
    g_launches.fetch_add((uint64_t)nodes, std::memory_order_relaxed);
    return NRS_OK;
}

int check_extract_args(nrs_ctx *ctx, const uint8_t *raw, int R, int row_pitch, const nrs_extract_params *p)
{
    NRS_CHECK_ARG(ctx && raw && p, "bad argument");
    NRS_CHECK_ARG(R >= 2 && R <= FE_MAX_AZ, "number of azimuths out of range");
    NRS_CHECK_ARG(p->range_bins >= 3 && p->range_bins <= 32 * FE_ROW_WORDS, "range_bins out of range (3..8192)");
    NRS_CHECK_ARG(row_pitch >= p->range_bins + 11, "row pitch smaller than 11 + range_bins");
    NRS_CHECK_ARG(p->cart_pixel_width >= 64 && p->patch_size >= 3 && p->patch_size <= 63 && p->max_points <= FE_KP_CAP, "bad parameters");
    NRS_CHECK_ARG(p->keypoint_extraction == 0 || p->keypoint_extraction == 1, "keypoint_extraction must be 0 (cen2019) or 1 (cen2018)");
    return NRS_OK;
}

// scratch of the matcher (ctx->scratch[5])
struct MatchBufs {
    int *best, *q, *t, *n;
    unsigned char *good;
    double *p, *c;
    long long *pt, *ct;
};
int match_carve(nrs_ctx *ctx, int nq, int cap, MatchBufs &m)
{
    const size_t c = (size_t)std::max(cap, 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_best = take((size_t)std::max(nq, 1) * 4), o_good = take(std::max(nq, 1)), o_q = take(c * 4), o_t = take(c * 4);
    const size_t o_p = take(2 * c * 8), o_c = take(2 * c * 8), o_pt = take(c * 8), o_ct = take(c * 8), o_n = take(4);
    int rc = ctx->scratch[5].reserve(off);
    if (rc) return rc;
    unsigned char *b = ctx->scratch[5].as<unsigned char>();
    m.best = (int *)(b + o_best); m.good = b + o_good; m.q = (int *)(b + o_q); m.t = (int *)(b + o_t);
    m.p = (double *)(b + o_p); m.c = (double *)(b + o_c); m.pt = (long long *)(b + o_pt); m.ct = (long long *)(b + o_ct);
    m.n = (int *)(b + o_n);
    return NRS_OK;
}

}  // namespace

int nrs_frame_extract(nrs_ctx *ctx, const uint8_t *raw, int R, int row_pitch, const nrs_extract_params *p,
                      int64_t t_criteria, nrs_frame **out)
{
    NRS_CHECK_ARG(out, "bad argument");
    *out = nullptr;
    int rc = check_extract_args(ctx, raw, R, row_pitch, p);
    if (rc) return rc;
    const int C = p->range_bins;
    std::lock_guard<std::mutex> lk(ctx->mu);
    NRS_CUDA(cudaSetDevice(ctx->device));
    FrontendState *S = nullptr;
    rc = frontend_state(ctx, p, R, C, row_pitch, &S);
    if (rc) return rc;
    nrs_frame *f = nullptr;
    rc = alloc_frame(ctx, p->cart_pixel_width, &f);
    if (rc) return rc;
    f->t_criteria = t_criteria;
    cudaStream_t st = ctx->stream;
    HostTables H;
    if ((rc = extract_upload(ctx, S, raw, R, row_pitch, H))) { frame_release_locked(f); return rc; }
    f->t_min = H.tmin; f->t_max = H.tmax;
    // a second pass with the full-frame sort only when cen2019's top-candidate subset could not vouch for the
    // marking loop's outcome (FrameCounts::need_full)
    for (int attempt = 0; attempt < 2; ++attempt) {
        if ((rc = extract_run(ctx, S, f, R, C, row_pitch, p, H.n_sorted, attempt == 1))) { frame_release_locked(f); return rc; }
        NRS_CUDA(cudaMemcpyAsync(S->h_counts, S->d_counts, sizeof(FrameCounts), cudaMemcpyDeviceToHost, st));
        NRS_CUDA(cudaStreamSynchronize(st));
        if (!S->h_counts->need_full) break;
    }
    f->counts = *S->h_counts;
    if (f->counts.overflow) {
        frame_release_locked(f);
        set_error("nrs_frame_extract: more than %d keypoints", FE_KP_CAP);
        return NRS_ECAPACITY;
    }
    *out = f;
    return NRS_OK;
}

// One iteration of the node's loop (odometry.cpp:153-284) in ONE call: extractFeature(new frame),
// matchFeatures(prev, curr), ORORA::calcTransform -- the matcher is enqueued behind the extraction without
// waiting for it (the new frame's keypoint count stays on the device), the matched points go to the solver
// without a round trip through host memory, and it is enqueued before the match count is known to the host (the
// kernels read it on the device and gate themselves): ONE host synchronisation per scan pair instead of four.
int nrs_orora_pair(nrs_ctx *ctx, const nrs_frame *prev, const uint8_t *raw, int R, int row_pitch,
                   const nrs_extract_params *p, int64_t t_criteria, const nrs_solver_params *sp, double vel_x,
                   nrs_frame **curr_out, double T_out[16], int *n_matches, nrs_solver_info *info_out)
{
    NRS_CHECK_ARG(prev && sp && curr_out && T_out && info_out, "bad argument");
    *curr_out = nullptr;
    int rc = check_extract_args(ctx, raw, R, row_pitch, p);
    if (rc) return rc;
    NRS_CHECK_ARG(prev->ctx == ctx, "prev belongs to another context");
    const int C = p->range_bins;
    std::lock_guard<std::mutex> lk(ctx->mu);
    NRS_CUDA(cudaSetDevice(ctx->device));
    FrontendState *S = nullptr;
    rc = frontend_state(ctx, p, R, C, row_pitch, &S);
    if (rc) return rc;
    nrs_frame *f = nullptr;
    rc = alloc_frame(ctx, p->cart_pixel_width, &f);
    if (rc) return rc;
    f->t_criteria = t_criteria;
    cudaStream_t st = ctx->stream;
    HostTables H;
    if ((rc = extract_upload(ctx, S, raw, R, row_pitch, H))) { frame_release_locked(f); return rc; }
    f->t_min = H.tmin; f->t_max = H.tmax;
    const int nq = prev->counts.Kd, cap = std::max(nq, 1);
    MatchBufs m;
    if ((rc = match_carve(ctx, nq, cap, m))) { frame_release_locked(f); return rc; }
    int h_n = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if ((rc = extract_run(ctx, S, f, R, C, row_pitch, p, H.n_sorted, attempt == 1))) { frame_release_locked(f); return rc; }
        NRS_CUDA(cudaMemsetAsync(m.n, 0, 4, st));
        if (nq > 0) {
            fe_match_kernel<<<(nq + MATCH_QPB - 1) / MATCH_QPB, 256, 0, st>>>(prev->d_desc, nq, f->d_desc, 0, S->d_counts, p->nndr, m.best, m.good);
            NRS_LAUNCHED();
            fe_match_gather_kernel<<<1, 1024, 0, st>>>(m.best, m.good, nq, prev->d_feat, f->d_feat, prev->d_t, f->d_t, cap, m.q, m.t,
                                                       m.p, m.c, m.pt, m.ct, m.n);
            NRS_LAUNCHED();
        }
        // the solver goes in behind the matcher without waiting for the match count (it stays on the device): ONE host
        // synchronisation per scan pair
        if ((rc = nrs::orora_pair_enqueue(ctx, m.p, m.c, m.n, sp, vel_x))) { frame_release_locked(f); return rc; }
        NRS_CUDA(cudaMemcpyAsync(S->h_counts, S->d_counts, sizeof(FrameCounts), cudaMemcpyDeviceToHost, st));
        NRS_CUDA(cudaMemcpyAsync(S->h_n, m.n, 4, cudaMemcpyDeviceToHost, st));
        NRS_CUDA(cudaStreamSynchronize(st));
        h_n = *S->h_n;
        if (!S->h_counts->need_full) break;
    }
    f->counts = *S->h_counts;
    if (f->counts.overflow) {
        frame_release_locked(f);
        set_error("nrs_orora_pair: more than %d keypoints", FE_KP_CAP);
        return NRS_ECAPACITY;
    }
    if (f->counts.Kd == 0) h_n = 0;                  // the matcher saw an empty train set
    if (n_matches) *n_matches = h_n;
    rc = nrs::orora_pair_finish(ctx, m.p, m.c, h_n, sp, vel_x, T_out, info_out);
    if (rc) { frame_release_locked(f); return rc; }
    *curr_out = f;
    return NRS_OK;
}

int nrs_frame_destroy(nrs_frame *f)
{
    if (!f) return NRS_OK;
    // every consumer of a frame's buffers runs on the context stream and the entry points that
    // enqueue them synchronise before returning, so the block can be recycled immediately
    if (f->d_block) {
        std::lock_guard<std::mutex> lk(f->ctx->mu);
        if (f->ctx->frame_pool.size() < 8) f->ctx->frame_pool.emplace_back(f->d_block, f->block_bytes);
        else { cudaSetDevice(f->ctx->device); cudaFree(f->d_block); }
    }
    delete f;
    return NRS_OK;
}

int nrs_frontend_tail_times(nrs_ctx *ctx, uint64_t *t_ns8)
{
    NRS_CHECK_ARG(ctx && t_ns8, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    FrontendState *S = static_cast<FrontendState *>(ctx->frontend);
    NRS_CHECK_ARG(S && S->h_counts, "no extraction ran on this context yet");
    for (int i = 0; i < 8; ++i) t_ns8[i] = S->h_counts->t_tail[i];
    return NRS_OK;
}

int nrs_frame_info_get(const nrs_frame *f, nrs_frame_info *info)
{
    NRS_CHECK_ARG(f && info, "bad argument");
    info->n_candidates = f->counts.n_cand;
    info->j_consumed = f->counts.j_star + 1;
    info->l_regions = f->counts.l_regions;
    info->K_polar = f->counts.K0;
    info->K_voxel = f->counts.Kv;
    info->K = f->counts.K;
    info->K_desc = f->counts.Kd;
    info->n_sorted = f->counts.n_sorted;
    info->maxg = f->counts.maxg; info->mean = f->counts.mean; info->mean_h = f->counts.mean_h;
    info->cart_max = f->counts.cart_max;
    info->t_criteria = f->t_criteria;
    return NRS_OK;
}

int nrs_frame_get(nrs_frame *f, double *polar, double *feat, float *kp, int64_t *t, double *t_norm,
                  uint8_t *desc, uint8_t *img)
{
    NRS_CHECK_ARG(f, "frame is NULL");
    nrs_ctx *ctx = f->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    NRS_CUDA(cudaSetDevice(ctx->device));
    const int K0 = f->counts.K0, K = f->counts.K, Kd = f->counts.Kd;
    std::vector<int> hp;
    std::vector<double> hf;
    std::vector<long long> ht;
    if (polar && K0) { hp.resize(2 * (size_t)K0); NRS_CUDA(cudaMemcpyAsync(hp.data(), f->d_polar, hp.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)); }
    if (feat && K) { hf.resize(2 * (size_t)K); NRS_CUDA(cudaMemcpyAsync(hf.data(), f->d_feat, hf.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)); }
    if ((t || t_norm) && K) { ht.resize(K); NRS_CUDA(cudaMemcpyAsync(ht.data(), f->d_t, (size_t)K * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream)); }
    if (kp && Kd) NRS_CUDA(cudaMemcpyAsync(kp, f->d_kp_orb, 2 * (size_t)Kd * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (desc && Kd) NRS_CUDA(cudaMemcpyAsync(desc, f->d_desc, (size_t)Kd * 32, cudaMemcpyDeviceToHost, ctx->stream));
    if (img) NRS_CUDA(cudaMemcpyAsync(img, f->d_img, (size_t)f->W * f->W, cudaMemcpyDeviceToHost, ctx->stream));
    NRS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (polar) for (int i = 0; i < K0; ++i) { polar[3 * i] = hp[2 * i]; polar[3 * i + 1] = hp[2 * i + 1]; polar[3 * i + 2] = 1.0; }
    const FrontendState *S = static_cast<FrontendState *>(ctx->frontend);
    const double z = (S && S->p.use_voxelization) ? 0.0 : 1.0;          // radar_loader.hpp:216 Zero(3, n)
    if (feat) for (int i = 0; i < K; ++i) { feat[3 * i] = hf[2 * i]; feat[3 * i + 1] = hf[2 * i + 1]; feat[3 * i + 2] = z; }
    if (t) for (int i = 0; i < K; ++i) t[i] = ht[i];
    if (t_norm) {                                                        // normalizeTimestamps :244-252
        const double den = (double)(f->t_max - f->t_min);
        for (int i = 0; i < K; ++i) t_norm[i] = (double)(ht[i] - f->t_min) / den;
    }
    return NRS_OK;
}

// marked matrix R = { claim <= J } materialised for parity tests
__global__ void fe_marked_kernel(const unsigned int *__restrict__ claim, size_t N, const FrameCounts *__restrict__ cnt,
                                 unsigned char *__restrict__ out)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < N) out[e] = (cnt->j_star >= 0 && claim[e] <= (unsigned int)cnt->j_star) ? 1 : 0;
}

int nrs_frame_get_stages(nrs_frame *f, float *fft, uint8_t *marked, double *cart_pre, float *cart_f32)
{
    NRS_CHECK_ARG(f, "frame is NULL");
    nrs_ctx *ctx = f->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    FrontendState *S = static_cast<FrontendState *>(ctx->frontend);
    NRS_CHECK_ARG(S != nullptr, "no extraction has run on this context");
    NRS_CUDA(cudaSetDevice(ctx->device));
    const size_t N = (size_t)S->R * S->C;
    if (fft) NRS_CUDA(cudaMemcpyAsync(fft, S->d_fft, N * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (marked) {
        unsigned char *tmp = reinterpret_cast<unsigned char *>(S->d_keys[1]);
        NRS_CUDA(cudaMemcpyAsync(S->d_counts, &f->counts, sizeof(FrameCounts), cudaMemcpyHostToDevice, ctx->stream));
        fe_marked_kernel<<<(unsigned int)((N + 255) / 256), 256, 0, ctx->stream>>>(S->d_claim, N, S->d_counts, tmp);
        NRS_LAUNCHED();
        NRS_CUDA(cudaMemcpyAsync(marked, tmp, N, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (cart_pre && f->counts.K0)
        NRS_CUDA(cudaMemcpyAsync(cart_pre, f->d_cart_pre, 2 * (size_t)f->counts.K0 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (cart_f32) NRS_CUDA(cudaMemcpyAsync(cart_f32, S->d_cart, (size_t)f->W * f->W * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    NRS_CUDA(cudaStreamSynchronize(ctx->stream));
    return NRS_OK;
}

int nrs_match(nrs_ctx *ctx, const nrs_frame *prev, const nrs_frame *curr, float nndr, int cap,
              int32_t *query_idx, int32_t *train_idx, double *prev_2xN, double *curr_2xN,
              int64_t *prev_t, int64_t *curr_t, int *N)
{
    NRS_CHECK_ARG(ctx && prev && curr && N && cap >= 0, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    NRS_CUDA(cudaSetDevice(ctx->device));
    const int nq = prev->counts.Kd, nt = curr->counts.Kd;   // desc_ rows; indices address feat_in_cart_ as-is
    *N = 0;
    if (nq == 0 || nt == 0) return NRS_OK;
    const size_t c = (size_t)std::max(cap, 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_best = take((size_t)nq * 4), o_good = take(nq), o_q = take(c * 4), o_t = take(c * 4);
    const size_t o_p = take(2 * c * 8), o_c = take(2 * c * 8), o_pt = take(c * 8), o_ct = take(c * 8), o_n = take(4);
    int rc = ctx->scratch[5].reserve(off);
    if (rc) return rc;
    unsigned char *b = ctx->scratch[5].as<unsigned char>();
    cudaStream_t st = ctx->stream;
    fe_match_kernel<<<(nq + MATCH_QPB - 1) / MATCH_QPB, 256, 0, st>>>(prev->d_desc, nq, curr->d_desc, nt, nullptr, nndr,
                                                          (int *)(b + o_best), b + o_good);
    NRS_LAUNCHED();
    fe_match_gather_kernel<<<1, 1024, 0, st>>>((int *)(b + o_best), b + o_good, nq, prev->d_feat, curr->d_feat,
                                               prev->d_t, curr->d_t, cap, (int *)(b + o_q), (int *)(b + o_t),
                                               (double *)(b + o_p), (double *)(b + o_c), (long long *)(b + o_pt),
                                               (long long *)(b + o_ct), (int *)(b + o_n));
    NRS_LAUNCHED();
    int n = 0;
    NRS_CUDA(cudaMemcpyAsync(&n, b + o_n, 4, cudaMemcpyDeviceToHost, st));
    NRS_CUDA(cudaStreamSynchronize(st));
    *N = n;
    if (n > cap) { set_error("nrs_match: %d matches exceed cap = %d", n, cap); return NRS_ECAPACITY; }
    if (n) {
        if (query_idx) NRS_CUDA(cudaMemcpyAsync(query_idx, b + o_q, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        if (train_idx) NRS_CUDA(cudaMemcpyAsync(train_idx, b + o_t, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        if (prev_2xN) NRS_CUDA(cudaMemcpyAsync(prev_2xN, b + o_p, 2 * (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        if (curr_2xN) NRS_CUDA(cudaMemcpyAsync(curr_2xN, b + o_c, 2 * (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        if (prev_t) NRS_CUDA(cudaMemcpyAsync(prev_t, b + o_pt, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        if (curr_t) NRS_CUDA(cudaMemcpyAsync(curr_t, b + o_ct, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        NRS_CUDA(cudaStreamSynchronize(st));
    }
    return NRS_OK;
}

}  // extern "C"
