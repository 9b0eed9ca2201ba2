// frontend_cen2019.cu -- load_radar decode + cen2019 keypoint extraction + polar->Cartesian
// points + voxelisation + BEV filter on sm_100a.
//
// Reference: outlier-robust-radar-odometry/src/radar_utils.cpp:104-129 (load_radar),
// src/features.cpp:133-299 (cen2019features and its helpers), src/radar_utils.cpp:329-341
// (polar_to_cartesian_points), include/radar_loader.hpp:207-222 (voxelize -> pcl::VoxelGrid),
// src/radar_utils.cpp:382-403 (convert_to_bev).
//
// cen2019 is NOT purely per-azimuth in the reference: it has three global reductions, a global
// descending sort of every pixel with h > mean(h) and a sequential region-marking loop that
// stops after max_points fresh regions.  The loop is parallelised EXACTLY:
//   K1  fe_decode / fe_stats / fe_h     elementwise f32 passes in the reference's operation order
//                                       (Prewitt |f[j+1]-f[j-1]|, /max, s = f-mean, h = s*(1-g)),
//                                       f64 block sums combined in a fixed order, max by atomics.
//   K2  sort + claim                    all R*C (key = h or -FLT_MAX, value = pixel) pairs are radix
//                                       sorted descending (stable -> ties keep pixel order, which is
//                                       how the oracle restates std::sort); candidate k claims its
//                                       region [rlow, rhigh] (findRangeBoundaries incl. the s.rows
//                                       quirk) with atomicMin(rank); a region counts as fresh iff it
//                                       owns all of its pixels; an inclusive scan finds the rank J at
//                                       which the max_points-th fresh region appears; the marked matrix
//                                       R is { pixel : claim <= J }.
//   K3  fe_scan_rows                    one warp per azimuth: maximal marked runs (ballot), adjacency
//                                       test against the neighbouring azimuths (wrap-around),
//                                       getMaxInRegion's int-truncated running max, ordered emission.
// Bound: HBM (R*C*4 = 5.4 MB per frame plus the sort's passes); DESIGN.md gives the byte counts.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <float.h>
#include <math.h>
#include <string.h>

#include "frontend.cuh"

namespace nrs {

// ------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------
// load_radar (radar_utils.cpp:117-126): bins 0..41 stay 0, fft = (float)byte / 255.0 (double division)
__global__ void fe_decode_kernel(const unsigned char *__restrict__ raw, int pitch, int R, int C,
                                 float *__restrict__ fft)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= C) return;
    float v = 0.0f;
    if (j >= 42) v = (float)((double)raw[(size_t)i * pitch + 11 + j] / 255.0);
    fft[(size_t)i * C + j] = v;
}

__device__ __forceinline__ float prewitt_abs(const float *__restrict__ row, int j, int C)
{
    // filter2D [-1 0 1], BORDER_REFLECT101 (features.cpp:200-205): edge columns give 0
    const int jl = (j == 0) ? 1 : j - 1;
    const int jr = (j == C - 1) ? C - 2 : j + 1;
    return fabsf(row[jr] - row[jl]);
}

constexpr int RED_THREADS = 256;

// sum(f) in f64 (exact for k/255 data: every partial sum is a multiple of 2^-31 below 2^21) and max(g)
__global__ void __launch_bounds__(RED_THREADS)
fe_stats_kernel(const float *__restrict__ fft, int R, int C, double *__restrict__ partial,
                int *__restrict__ maxg_bits)
{
    const size_t N = (size_t)R * C;
    double s = 0.0;
    float mg = 0.0f;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < N; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / C), j = (int)(e - (size_t)i * C);
        const float *row = fft + (size_t)i * C;
        s += (double)row[j];
        mg = fmaxf(mg, prewitt_abs(row, j, C));
    }
    typedef cub::BlockReduce<double, RED_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    __shared__ float smax[RED_THREADS / 32];
    const double bs = BR(tmp).Sum(s);
    for (int o = 16; o > 0; o >>= 1) mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, o));
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mg;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.0f;
        for (int w = 0; w < RED_THREADS / 32; ++w) m = fmaxf(m, smax[w]);
        partial[blockIdx.x] = bs;
        atomicMax(maxg_bits, __float_as_int(m));      // g >= 0: int order == float order
    }
}

// cv::mean = sum * (1.0 / N) in f64, truncated to float by the reference (features.cpp:211,214)
__global__ void fe_mean_kernel(const double *__restrict__ partial, int nblocks, double inv_n,
                               float *__restrict__ mean_out)
{
    // one warp, fixed order: lane l sums partial[l], partial[l+32], ... then a fixed shuffle tree
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 32) s += partial[b];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *mean_out = (float)(s * inv_n);
}

// h = s .* (1 - g / maxg) (features.cpp:206-213), all f32; Mat /= double multiplies by (float)(1/maxg)
__global__ void __launch_bounds__(RED_THREADS)
fe_h_kernel(const float *__restrict__ fft, int R, int C, const FrameCounts *__restrict__ cnt,
            float *__restrict__ h, double *__restrict__ partial)
{
    const size_t N = (size_t)R * C;
    const float mean = cnt->mean;
    const float inv = (float)(1.0 / (double)cnt->maxg);
    double s = 0.0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < N; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / C), j = (int)(e - (size_t)i * C);
        const float *row = fft + (size_t)i * C;
        const float g = prewitt_abs(row, j, C) * inv;
        const float sv = row[j] - mean;
        const float hv = sv * (1.0f - g);
        h[e] = hv;
        s += (double)hv;
    }
    typedef cub::BlockReduce<double, RED_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    const double bs = BR(tmp).Sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = bs;
}

// ------------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS)
fe_keys_kernel(const float *__restrict__ h, size_t N, FrameCounts *__restrict__ cnt,
               float *__restrict__ keys, unsigned int *__restrict__ vals,
               unsigned int *__restrict__ claim)
{
    const float mean_h = cnt->mean_h;
    int c = 0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < N; e += (size_t)gridDim.x * blockDim.x) {
        const float hv = h[e];
        const bool cand = hv > mean_h;                       // features.cpp:220
        keys[e] = cand ? (hv == 0.0f ? 0.0f : hv) : -FLT_MAX; // -0.0 and +0.0 compare equal in the reference
        vals[e] = (unsigned int)e;
        claim[e] = 0xffffffffu;
        c += cand;
    }
    typedef cub::BlockReduce<int, RED_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    const int bc = BR(tmp).Sum(c);
    if (threadIdx.x == 0 && bc) atomicAdd(&cnt->n_cand, bc);
}

// findRangeBoundaries (features.cpp:151-170); s < 0  <=>  f < mean (f32 subtraction keeps the sign)
__device__ __forceinline__ void range_bounds(const float *__restrict__ row, float mean, int R, int C,
                                             int r, int &rlow, int &rhigh)
{
    rlow = r;
    rhigh = r;
    for (int i = r - 1; i >= 0; --i) {
        if (row[i] - mean < 0.0f) rlow = i; else break;
    }
    if (r < R - 1) {                                         // quirk kept: s.rows, not s.cols (:162)
        for (int i = r + 1; i < C; ++i) {
            if (row[i] - mean < 0.0f) rhigh = i; else break;
        }
    }
}

__global__ void fe_claim_kernel(const unsigned int *__restrict__ sorted_vals,
                                const float *__restrict__ fft, int R, int C,
                                const FrameCounts *__restrict__ cnt, unsigned int *__restrict__ claim)
{
    const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (unsigned int)cnt->n_sorted) return;
    const unsigned int e = sorted_vals[k];
    const int a = e / C, r = e - a * C;
    const float *row = fft + (size_t)a * C;
    int lo, hi;
    range_bounds(row, cnt->mean, R, C, r, lo, hi);
    unsigned int *crow = claim + (size_t)a * C;
    for (int i = lo; i <= hi; ++i) atomicMin(&crow[i], k);
}

__global__ void fe_incr_kernel(const unsigned int *__restrict__ sorted_vals,
                               const float *__restrict__ fft, int R, int C,
                               const FrameCounts *__restrict__ cnt,
                               const unsigned int *__restrict__ claim, unsigned int *__restrict__ incr,
                               size_t N)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    unsigned int fresh = 0;
    if (k < (size_t)cnt->n_sorted) {
        const unsigned int e = sorted_vals[k];
        const int a = e / C, r = e - a * C;
        const float *row = fft + (size_t)a * C;
        int lo, hi;
        range_bounds(row, cnt->mean, R, C, r, lo, hi);
        const unsigned int *crow = claim + (size_t)a * C;
        fresh = 1;
        for (int i = lo; i <= hi; ++i)
            if (crow[i] != (unsigned int)k) { fresh = 0; break; }
    }
    incr[k] = fresh;
}

// the marking loop exits right after the increment that makes l == max_points (features.cpp:231)
__global__ void fe_jstar_kernel(const unsigned int *__restrict__ incr, const unsigned int *__restrict__ cums,
                                size_t N, int max_points, FrameCounts *__restrict__ cnt)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    if (k < (size_t)cnt->n_sorted && incr[k] && cums[k] == (unsigned int)max_points) cnt->j_star = (int)k;
    if (k == N - 1) cnt->l_regions = min((int)cums[k], max_points);
}
// the loop ran out of candidates before max_points regions: every candidate was consumed -- which the
// top-candidate subset can only vouch for when it held them all.  Also carries the threshold to the next frame:
// the h value at rank ~4x the candidates this frame consumed (at least 16k, at most 48k of the 64k capacity).
__global__ void fe_jstar_fix_kernel(FrameCounts *__restrict__ cnt, const float *__restrict__ sorted_keys,
                                    float *__restrict__ thresh)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (cnt->j_star < 0) {
        if (cnt->n_sorted == cnt->n_cand) cnt->j_star = cnt->n_cand - 1;
        else cnt->need_full = 1;
    }
    if (cnt->need_full || cnt->n_sorted <= 0) return;
    const int want = min(max(4 * (cnt->j_star + 1), 16384), 49152);
    if (cnt->n_sorted > want) *thresh = sorted_keys[want];
    else if (cnt->n_sorted == cnt->n_cand) *thresh = sorted_keys[cnt->n_sorted - 1];           // everything fits: keep it all
    else if (cnt->n_sorted >= 2 * (cnt->j_star + 1)) *thresh = sorted_keys[cnt->n_sorted - 1]; // comfortable margin: keep
    else *thresh = cnt->mean_h + 0.5f * (sorted_keys[cnt->n_sorted - 1] - cnt->mean_h);        // thin margin: loosen
}

// ---- top-candidate subset (replaces the full-frame sort when a threshold is known) -----------------------------
// The marking loop (features.cpp:227-249) walks the candidates in descending h and stops after max_points fresh
// regions: only the first j_star + 1 of the ~6e5 candidates matter.  With a threshold T carried from the previous
// frame, the pixels with h >= T are compacted IN PIXEL ORDER (cub::DeviceSelect), sorted (64k pairs instead of
// 1.3M; stable, so ties keep pixel order exactly like the full sort) and walked.  Their ranks equal their ranks in
// the full order, so claims, regions and j_star are identical -- provided the loop terminates inside the subset
// (or the subset holds every candidate).  Otherwise need_full is raised and the host redoes the frame with the
// full sort; the result never depends on T.
struct SubSelect {
    const float *h;
    const float *thresh;
    const FrameCounts *cnt;
    __device__ __forceinline__ bool operator()(const unsigned int &e) const
    {
        const float hv = h[e];
        return hv > cnt->mean_h && hv >= *thresh;
    }
};

__global__ void __launch_bounds__(RED_THREADS)
fe_prep_sub_kernel(const float *__restrict__ h, size_t N, FrameCounts *__restrict__ cnt, unsigned int *__restrict__ claim)
{
    const float mean_h = cnt->mean_h;
    int c = 0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < N; e += (size_t)gridDim.x * blockDim.x) {
        c += h[e] > mean_h;                                  // features.cpp:220
        claim[e] = 0xffffffffu;
    }
    typedef cub::BlockReduce<int, RED_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    const int bc = BR(tmp).Sum(c);
    if (threadIdx.x == 0 && bc) atomicAdd(&cnt->n_cand, bc);
}

__global__ void fe_sub_keys_kernel(const float *__restrict__ h, FrameCounts *__restrict__ cnt, float *__restrict__ keys,
                                   unsigned int *__restrict__ vals)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= FE_SUB_CAP) return;
    const int n_sub = cnt->n_sub;
    if (k == 0) {
        cnt->n_sorted = min(n_sub, FE_SUB_CAP);
        if (n_sub > FE_SUB_CAP) cnt->need_full = 1;         // the first 64k in pixel order are not the top 64k
    }
    if (k < n_sub) {
        const float hv = h[vals[k]];
        keys[k] = (hv == 0.0f) ? 0.0f : hv;
    } else {
        keys[k] = -FLT_MAX;
        vals[k] = 0xffffffffu;
    }
}
__global__ void fe_set_sorted_full_kernel(FrameCounts *__restrict__ cnt)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) cnt->n_sorted = cnt->n_cand;
}

// ------------------------------------------------------------------------------------------------
// K3: per-azimuth scan (features.cpp:251-278), one warp per azimuth
// ------------------------------------------------------------------------------------------------
template <bool EMIT>
__global__ void __launch_bounds__(256)
fe_scan_rows_kernel(const unsigned int *__restrict__ claim, const float *__restrict__ h, int R, int C,
                    int min_range, const FrameCounts *__restrict__ cnt, int *__restrict__ rowcount,
                    int *__restrict__ polar, int cap)
{
    // one 256-thread block per azimuth, one column per thread per step; emission stays ordered by
    // column through per-warp ballots and a running base
    __shared__ int wsum[8];
    __shared__ int running;
    const int i = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int J = (unsigned int)cnt->j_star;
    const bool none = cnt->j_star < 0;
    const unsigned int *crow = claim + (size_t)i * C;
    const int below = (i == 0) ? R - 1 : i - 1, above = (i + 1 >= R) ? 0 : i + 1;     // :173-178
    const unsigned int *cb = claim + (size_t)below * C, *ca = claim + (size_t)above * C;
    const float *hrow = h + (size_t)i * C;
    if (threadIdx.x == 0) running = EMIT ? rowcount[i] : 0;
    __syncthreads();
    for (int j0 = min_range; j0 < C; j0 += 256) {
        const int j = j0 + threadIdx.x;
        bool emit = false;
        int max_r = 0;
        if (j < C && !none) {
            const bool m = crow[j] <= J;
            const bool prev = (j > min_range) && (crow[j - 1] <= J);
            if (m && !prev) {                      // a run starts here
                int end = j;
                while (end + 1 < C && crow[end + 1] <= J) ++end;
                if (end + 1 < C) {                 // closed by an unmarked pixel; else never emitted
                    bool adj = false;              // checkAdjacentMarked :172-184
                    for (int r = j; r <= end; ++r)
                        if (cb[r] <= J || ca[r] <= J) { adj = true; break; }
                    if (adj) {
                        // getMaxInRegion :186-194 with its int-typed running max
                        int mx = -1000;
                        max_r = j;
                        for (int r = j; r <= end; ++r) {
                            const float hv = hrow[r];
                            if (hv > (float)mx) { mx = (int)hv; max_r = r; }
                        }
                        emit = true;
                    }
                }
            }
        }
        const unsigned int mask = __ballot_sync(0xffffffffu, emit);
        if (lane == 0) wsum[warp] = __popc(mask);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const int c = wsum[w]; if (w < warp) before += c; total += c; }
        const int base = running;
        if (EMIT && emit) {
            const int pos = base + before + __popc(mask & ((1u << lane) - 1u));
            if (pos < cap) { polar[2 * pos] = i; polar[2 * pos + 1] = max_r; }
        }
        __syncthreads();
        if (threadIdx.x == 0) running = base + total;
        __syncthreads();
    }
    if (!EMIT && threadIdx.x == 0) rowcount[i] = running;
}

// exclusive scan of the per-azimuth counts (R <= 1024), total -> K0
__global__ void __launch_bounds__(1024)
fe_rowscan_kernel(int *__restrict__ rowcount, int R, FrameCounts *__restrict__ cnt, int cap)
{
    typedef cub::BlockScan<int, 1024> BS;
    __shared__ typename BS::TempStorage tmp;
    const int t = threadIdx.x;
    int v = (t < R) ? rowcount[t] : 0, ex, total;
    BS(tmp).ExclusiveSum(v, ex, total);
    if (t < R) rowcount[t] = ex;
    if (t == 0) {
        cnt->K0 = min(total, cap);
        if (total > cap) cnt->overflow = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// polar_to_cartesian_points (radar_utils.cpp:329-341); cos/sin come from the host's libm table
// ------------------------------------------------------------------------------------------------
__global__ void fe_points_kernel(const int *__restrict__ polar, const FrameCounts *__restrict__ cnt,
                                 const double *__restrict__ cossin, const long long *__restrict__ times,
                                 float radar_resolution, double *__restrict__ cart_pre,
                                 long long *__restrict__ t_pre)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt->K0) return;
    const int a = polar[2 * k], rb = polar[2 * k + 1];
    const double r = (double)rb * (double)radar_resolution + (double)(radar_resolution / 2);
    cart_pre[2 * k] = r * cossin[2 * a];
    cart_pre[2 * k + 1] = r * cossin[2 * a + 1];
    t_pre[k] = times[a];
}

// ------------------------------------------------------------------------------------------------
// voxelize (radar_loader.hpp:207-222 -> pcl::VoxelGrid<PointXYZ>::applyFilter), float32 like PCL
// ------------------------------------------------------------------------------------------------
struct VoxGrid { int min_bx, min_by, div_x; float inv; };

__global__ void __launch_bounds__(1024)
fe_vox_keys_kernel(const double *__restrict__ cart_pre, const FrameCounts *__restrict__ cnt,
                   float leaf, int cap, unsigned int *__restrict__ keys, unsigned int *__restrict__ vals,
                   float *__restrict__ pts32)
{
    __shared__ float smin[2][32], smax[2][32];
    __shared__ VoxGrid g;
    const int t = threadIdx.x, K0 = cnt->K0;
    float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;
    for (int k = t; k < K0; k += blockDim.x) {
        const float x = (float)cart_pre[2 * k], y = (float)cart_pre[2 * k + 1];   // eigen2pcl: double -> float
        pts32[2 * k] = x; pts32[2 * k + 1] = y;
        mnx = fminf(mnx, x); mxx = fmaxf(mxx, x);
        mny = fminf(mny, y); mxy = fmaxf(mxy, y);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((t & 31) == 0) { smin[0][t >> 5] = mnx; smin[1][t >> 5] = mny; smax[0][t >> 5] = mxx; smax[1][t >> 5] = mxy; }
    __syncthreads();
    if (t == 0) {
        for (int w = 1; w < 32; ++w) {
            mnx = fminf(mnx, smin[0][w]); mny = fminf(mny, smin[1][w]);
            mxx = fmaxf(mxx, smax[0][w]); mxy = fmaxf(mxy, smax[1][w]);
        }
        const float inv = 1.0f / leaf;                          // inverse_leaf_size_
        g.inv = inv;
        g.min_bx = (int)floorf(mnx * inv);
        g.min_by = (int)floorf(mny * inv);
        g.div_x = (int)floorf(mxx * inv) - g.min_bx + 1;        // div_b_[0]; z is constant -> ijk2 = 0
    }
    __syncthreads();
    for (int k = t; k < cap; k += blockDim.x) {
        unsigned int key = 0xffffffffu;
        if (k < K0) {
            const int i0 = (int)(floorf(pts32[2 * k] * g.inv) - (float)g.min_bx);
            const int i1 = (int)(floorf(pts32[2 * k + 1] * g.inv) - (float)g.min_by);
            key = (unsigned int)(i0 + i1 * g.div_x);
        }
        keys[k] = key;
        vals[k] = (unsigned int)k;
    }
}

// centroids in ascending voxel index, f32 running sums in (stable) point order, / float(count)
__global__ void __launch_bounds__(1024)
fe_vox_centroid_kernel(const unsigned int *__restrict__ skeys, const unsigned int *__restrict__ svals,
                       const float *__restrict__ pts32, FrameCounts *__restrict__ cnt, int cap,
                       double *__restrict__ vox_xy)
{
    typedef cub::BlockScan<int, 1024> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ int carry;
    const int t = threadIdx.x, K0 = cnt->K0;
    if (t == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < cap; base += 1024) {
        const int k = base + t;
        const bool head = (k < K0) && (k == 0 || skeys[k] != skeys[k - 1]);
        int ex, total;
        BS(tmp).ExclusiveSum(head ? 1 : 0, ex, total);
        const int c0 = carry;
        if (head) {
            float sx = 0.0f, sy = 0.0f;
            int n = 0;
            for (int q = k; q < K0 && skeys[q] == skeys[k]; ++q) {
                const unsigned int p = svals[q];
                sx += pts32[2 * p];
                sy += pts32[2 * p + 1];
                ++n;
            }
            const float fn = (float)n;
            vox_xy[2 * (c0 + ex)] = (double)(sx / fn);
            vox_xy[2 * (c0 + ex) + 1] = (double)(sy / fn);
        }
        __syncthreads();
        if (t == 0) carry = c0 + total;
        __syncthreads();
    }
    if (t == 0) cnt->Kv = carry;
}

// convert_to_bev (radar_utils.cpp:382-403): ordered compaction, keypoints (u, v), stamps with the
// reference's pre-voxelisation indexing quirk (radar_loader.hpp:123-124)
__global__ void __launch_bounds__(1024)
fe_bev_kernel(const double *__restrict__ pts_xy, const long long *__restrict__ t_pre,
              FrameCounts *__restrict__ cnt, int use_vox, float cart_resolution, int width, int patch,
              int cap, double *__restrict__ feat, float *__restrict__ kp, long long *__restrict__ t_out,
              float *__restrict__ kp_orb)
{
    typedef cub::BlockScan<int, 1024> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ int carry;
    const int t = threadIdx.x;
    const int n = use_vox ? cnt->Kv : cnt->K0;
    float cmr = (width / 2) * cart_resolution;
    if (width % 2 == 0) cmr = (float)((width / 2 - 0.5) * cart_resolution);
    if (t == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < cap; base += 1024) {
        const int i = base + t;
        bool keep = false;
        double x = 0, y = 0, u = 0, v = 0;
        if (i < n) {
            x = pts_xy[2 * i]; y = pts_xy[2 * i + 1];
            u = ((double)cmr + y) / (double)cart_resolution;
            v = ((double)cmr - x) / (double)cart_resolution;
            keep = (0 < u - patch) && (u + patch < width) && (0 < v - patch) && (v + patch < width);
        }
        int ex, total;
        BS(tmp).ExclusiveSum(keep ? 1 : 0, ex, total);
        const int c0 = carry;
        if (keep) {
            const int j = c0 + ex;
            feat[2 * j] = x; feat[2 * j + 1] = y;
            kp[2 * j] = (float)u; kp[2 * j + 1] = (float)v;
            t_out[j] = t_pre[i];
        }
        __syncthreads();
        if (t == 0) carry = c0 + total;
        __syncthreads();
    }
    __syncthreads();
    const int K = carry;
    __syncthreads();
    if (t == 0) { cnt->K = K; carry = 0; }
    __syncthreads();
    // cv::ORB::compute (radar_loader.hpp:126) first runs KeyPointsFilter::runByImageBorder(kps, size,
    // edgeThreshold = patch): Rect(Point(b,b), Point(W-b,W-b)).contains(Point(cvRound(u), cvRound(v))).
    // Keypoints with cvRound(u|v) == W - patch are dropped from kp_/desc_ while feat_in_cart_ keeps its
    // columns -- the reference's index misalignment is reproduced, not repaired.
    for (int base = 0; base < K; base += 1024) {
        const int j = base + t;
        bool keep = false;
        float u = 0.f, v = 0.f;
        if (j < K) {
            u = kp[2 * j]; v = kp[2 * j + 1];
            const int ru = __float2int_rn(u), rv = __float2int_rn(v);
            keep = ru >= patch && ru < width - patch && rv >= patch && rv < width - patch;
        }
        int ex, total;
        BS(tmp).ExclusiveSum(keep ? 1 : 0, ex, total);
        const int c0 = carry;
        if (keep) { kp_orb[2 * (c0 + ex)] = u; kp_orb[2 * (c0 + ex) + 1] = v; }
        __syncthreads();
        if (t == 0) carry = c0 + total;
        __syncthreads();
    }
    if (t == 0) cnt->Kd = carry;
}

// ------------------------------------------------------------------------------------------------
// host-side orchestration of cen2019 on a device-resident f32 frame
// ------------------------------------------------------------------------------------------------
static int stats_blocks(const nrs_ctx *ctx) { return ctx->sm_count * 8; }

int fe_launch_rowscan(cudaStream_t st, int *rowcount, int R, FrameCounts *cnt, int cap)
{
    fe_rowscan_kernel<<<1, 1024, 0, st>>>(rowcount, R, cnt, cap);
    NRS_LAUNCHED();
    return NRS_OK;
}

int fe_cen2019_device(nrs_ctx *ctx, FrontendState &S, int R, int C, int max_points, int min_range,
                      int *d_polar, bool reset_counts, bool force_full)
{
    cudaStream_t st = ctx->stream;
    const size_t N = (size_t)R * C;
    const int nb = stats_blocks(ctx);
    FrameCounts *cnt = S.d_counts;
    if (reset_counts) NRS_CUDA(cudaMemsetAsync(cnt, 0, sizeof(FrameCounts), st));
    fe_stats_kernel<<<nb, RED_THREADS, 0, st>>>(S.d_fft, R, C, S.d_partial, (int *)&cnt->maxg);
    NRS_LAUNCHED();
    fe_mean_kernel<<<1, 32, 0, st>>>(S.d_partial, nb, 1.0 / (double)N, &cnt->mean);
    NRS_LAUNCHED();
    fe_h_kernel<<<nb, RED_THREADS, 0, st>>>(S.d_fft, R, C, cnt, S.d_h, S.d_partial);
    NRS_LAUNCHED();
    fe_mean_kernel<<<1, 32, 0, st>>>(S.d_partial, nb, 1.0 / (double)N, &cnt->mean_h);
    NRS_LAUNCHED();
    const bool sub = S.have_thresh && !force_full;
    const size_t n_walk = sub ? (size_t)FE_SUB_CAP : N;       // length of the sorted list the marking kernels cover
    size_t need = S.cub_bytes;
    if (sub) {
        fe_prep_sub_kernel<<<nb, RED_THREADS, 0, st>>>(S.d_h, N, cnt, S.d_claim);
        NRS_LAUNCHED();
        thrust::counting_iterator<unsigned int> ids(0u);
        NRS_CUDA(cub::DeviceSelect::If(S.d_cub, need, ids, S.d_vals[0], &cnt->n_sub, (int)N, SubSelect{S.d_h, S.d_thresh, cnt}, st));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        fe_sub_keys_kernel<<<FE_SUB_CAP / 256, 256, 0, st>>>(S.d_h, cnt, S.d_keys[0], S.d_vals[0]);
        NRS_LAUNCHED();
    } else {
        fe_keys_kernel<<<nb, RED_THREADS, 0, st>>>(S.d_h, N, cnt, S.d_keys[0], S.d_vals[0], S.d_claim);
        NRS_LAUNCHED();
        fe_set_sorted_full_kernel<<<1, 32, 0, st>>>(cnt);
        NRS_LAUNCHED();
    }
    need = S.cub_bytes;
    NRS_CUDA(cub::DeviceRadixSort::SortPairsDescending(S.d_cub, need, S.d_keys[0], S.d_keys[1], S.d_vals[0],
                                                       S.d_vals[1], (int)n_walk, 0, 32, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const int tb = 256;
    const unsigned int gb = (unsigned int)((n_walk + tb - 1) / tb);
    fe_claim_kernel<<<gb, tb, 0, st>>>(S.d_vals[1], S.d_fft, R, C, cnt, S.d_claim);
    NRS_LAUNCHED();
    fe_incr_kernel<<<gb, tb, 0, st>>>(S.d_vals[1], S.d_fft, R, C, cnt, S.d_claim, S.d_incr, n_walk);
    NRS_LAUNCHED();
    need = S.cub_bytes;
    NRS_CUDA(cub::DeviceScan::InclusiveSum(S.d_cub, need, S.d_incr, S.d_vals[0], (int)n_walk, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    NRS_CUDA(cudaMemsetAsync(&cnt->j_star, 0xff, sizeof(int), st));      // -1
    fe_jstar_kernel<<<gb, tb, 0, st>>>(S.d_incr, S.d_vals[0], n_walk, max_points, cnt);
    NRS_LAUNCHED();
    fe_jstar_fix_kernel<<<1, 32, 0, st>>>(cnt, S.d_keys[1], S.d_thresh);
    NRS_LAUNCHED();
    S.have_thresh = true;
    const int sb = R;
    fe_scan_rows_kernel<false><<<sb, 256, 0, st>>>(S.d_claim, S.d_h, R, C, min_range, cnt, S.d_rowcount, nullptr, 0);
    NRS_LAUNCHED();
    fe_rowscan_kernel<<<1, 1024, 0, st>>>(S.d_rowcount, R, cnt, FE_KP_CAP);
    NRS_LAUNCHED();
    fe_scan_rows_kernel<true><<<sb, 256, 0, st>>>(S.d_claim, S.d_h, R, C, min_range, cnt, S.d_rowcount, d_polar, FE_KP_CAP);
    NRS_LAUNCHED();
    return NRS_OK;
}

int fe_points_voxel_bev_device(nrs_ctx *ctx, FrontendState &S, nrs_frame *f)
{
    cudaStream_t st = ctx->stream;
    FrameCounts *cnt = S.d_counts;
    const nrs_extract_params &p = S.p;
    fe_points_kernel<<<(FE_KP_CAP + 255) / 256, 256, 0, st>>>(f->d_polar, cnt, S.d_cossin, S.d_times,
                                                             p.radar_resolution, f->d_cart_pre, f->d_t_pre);
    NRS_LAUNCHED();
    const double *pts = f->d_cart_pre;
    if (p.use_voxelization) {
        fe_vox_keys_kernel<<<1, 1024, 0, st>>>(f->d_cart_pre, cnt, (float)p.voxel_size, FE_KP_CAP, S.d_vkey[0],
                                               S.d_vval[0], S.d_pts32);
        NRS_LAUNCHED();
        size_t need = S.cub_bytes;
        // voxel keys are i0 + i1 * div_x with both indices below extent / leaf + 2 (the points lie within the
        // sensor range): when that product stays under 2^24 three radix passes suffice.  The padding key
        // 0xffffffff has its low 24 bits set, so it still sorts behind every real key.
        const double cells = 2.0 * S.C * p.radar_resolution / p.voxel_size + 4.0;
        const int end_bit = (cells * cells < 16777214.0) ? 24 : 32;
        NRS_CUDA(cub::DeviceRadixSort::SortPairs(S.d_cub, need, S.d_vkey[0], S.d_vkey[1], S.d_vval[0], S.d_vval[1],
                                                 FE_KP_CAP, 0, end_bit, st));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        // voxel centroids are staged in d_partial's neighbour: reuse d_keys[0] (free after the sort)
        double *vox = reinterpret_cast<double *>(S.d_keys[0]);
        fe_vox_centroid_kernel<<<1, 1024, 0, st>>>(S.d_vkey[1], S.d_vval[1], S.d_pts32, cnt, FE_KP_CAP, vox);
        NRS_LAUNCHED();
        pts = vox;
    }
    fe_bev_kernel<<<1, 1024, 0, st>>>(pts, f->d_t_pre, cnt, p.use_voxelization, p.cart_resolution,
                                      p.cart_pixel_width, p.patch_size, FE_KP_CAP, f->d_feat, f->d_kp, f->d_t, f->d_kp_orb);
    NRS_LAUNCHED();
    return NRS_OK;
}

int fe_decode_device(nrs_ctx *ctx, FrontendState &S, int R, int C, int pitch)
{
    dim3 grid((C + 255) / 256, R);
    fe_decode_kernel<<<grid, 256, 0, ctx->stream>>>(S.d_raw, pitch, R, C, S.d_fft);
    NRS_LAUNCHED();
    return NRS_OK;
}

size_t fe_cub_bytes(size_t N)
{
    size_t a = 0, b = 0, c = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, a, (float *)nullptr, (float *)nullptr, (unsigned int *)nullptr,
                                              (unsigned int *)nullptr, (int)N);
    cub::DeviceScan::InclusiveSum(nullptr, b, (unsigned int *)nullptr, (unsigned int *)nullptr, (int)N);
    cub::DeviceRadixSort::SortPairs(nullptr, c, (unsigned int *)nullptr, (unsigned int *)nullptr, (unsigned int *)nullptr,
                                    (unsigned int *)nullptr, FE_KP_CAP);
    size_t d = 0;
    thrust::counting_iterator<unsigned int> ids(0u);
    cub::DeviceSelect::If(nullptr, d, ids, (unsigned int *)nullptr, (int *)nullptr, (int)N, SubSelect{nullptr, nullptr, nullptr});
    return std::max(std::max(a, d), std::max(b, c)) + 256;
}

}  // namespace nrs
