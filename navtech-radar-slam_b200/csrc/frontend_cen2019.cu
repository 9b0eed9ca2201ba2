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
//   K1  fe_decode_stats / fe_h          the decode and the reference's elementwise f32 passes in its operation order
//                                       (Prewitt |f[j+1]-f[j-1]|, /max, s = f-mean, h = s*(1-g)), one CTA per azimuth;
//                                       f64 block sums combined in a fixed order by the CTA that draws the last
//                                       ticket, max by atomics.  (fe_decode / fe_stats: the standalone operators.)
//   K2  sort + claim                    (key = h, value = pixel) pairs sorted descending, stable -> ties keep pixel
//                                       order, which is how the oracle restates std::sort.  Steady state: only the
//                                       pixels above a threshold carried from the previous frame are compacted
//                                       (cub::DeviceSelect) and sorted by fe_subsort_kernel -- a thread-block cluster
//                                       holding the pairs in distributed shared memory; first frame / fallback: all
//                                       R*C pairs through cub::DeviceRadixSort.  Candidate k claims its region
//                                       [rlow, rhigh] (findRangeBoundaries incl. the s.rows quirk) with
//                                       atomicMin(rank); a region counts as fresh iff it owns all of its pixels;
//                                       the last CTA of fe_incr_kernel finds the rank J at which the max_points-th
//                                       fresh region appears; the marked matrix R is { pixel : claim <= J }.
//   K3  fe_scan_rows / fe_tail          one CTA per azimuth: the row as one bit per column, maximal marked runs,
//                                       adjacency test against the neighbouring azimuths (wrap-around),
//                                       getMaxInRegion's int-truncated running max, into the azimuth's staging row;
//                                       fe_tail_kernel (one CTA) strings the rows together and does everything
//                                       from polar_to_cartesian_points to the BEV / ORB border filters, incl. the
//                                       voxel grid's sort in shared memory.
// Bound: a frame is 5.4 MB (1-2 us of HBM time per pass): every kernel here is launch / latency-bound, which is why the
// work of this file went into fewer launches and shorter serial chains; DESIGN.md gives the byte counts and times.
#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
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

// cv::mean = sum * (1.0 / N) in f64, truncated to float by the reference (features.cpp:211,214).  Called by every
// thread of every block after thread 0 stored the block's partial sum; the block that draws the last ticket reduces all partials in one
// fixed order (lane l sums partial[l], partial[l + 32], ... then a fixed shuffle tree), whichever block it is.
__device__ __forceinline__ void last_block_mean(unsigned int *ticket, const double *partial, double inv_n,
                                                float *mean_out, bool *is_last)
{
    const unsigned int nblocks = gridDim.x * gridDim.y;
    if (threadIdx.x == 0) {
        __threadfence();
        *is_last = atomicAdd(ticket, 1u) == nblocks - 1;
    }
    __syncthreads();
    if (*is_last && threadIdx.x < 32) {
        __threadfence();
        double s = 0.0;
        for (int b = threadIdx.x; b < (int)nblocks; b += 32) s += __ldcg(partial + b);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) *mean_out = (float)(s * inv_n);
    }
}

// sum(f) in f64 (exact for k/255 data: every partial sum is a multiple of 2^-31 below 2^21) and max(g)
__global__ void __launch_bounds__(RED_THREADS)
fe_stats_kernel(const float *__restrict__ fft, int R, int C, double *__restrict__ partial,
                FrameCounts *__restrict__ cnt)
{
    __shared__ bool is_last;
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
        atomicMax(reinterpret_cast<int *>(&cnt->maxg), __float_as_int(m));      // g >= 0: int order == float order
    }
    last_block_mean(&cnt->ticket[0], partial, 1.0 / (double)N, &cnt->mean, &is_last);
}

// K1 of the extraction pipeline in one pass over the raw PNG rows: load_radar's decode (fft written for the image
// branch and fe_h_kernel) and fe_stats_kernel's sum(f) / max(g) taken from the bytes directly.  The 256 possible
// sample values (float)(b / 255.0) come from a table computed once on the host (IEEE f64 division and narrowing:
// the same bits as the device's), so no division is left in the kernel.  One block per azimuth; the partial sums
// are exact (see fe_stats_kernel), so the mean does not depend on how the frame is cut into blocks.
__global__ void __launch_bounds__(RED_THREADS)
fe_decode_stats_kernel(const unsigned char *__restrict__ raw, int pitch, int R, int C, const float *__restrict__ lut_g,
                       float *__restrict__ fft, double *__restrict__ partial, FrameCounts *__restrict__ cnt)
{
    __shared__ float lut[256];
    __shared__ float smax[RED_THREADS / 32];
    __shared__ bool is_last;
    typedef cub::BlockReduce<double, RED_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    lut[threadIdx.x] = lut_g[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x;
    const unsigned char *row = raw + (size_t)i * pitch + 11;
    float *frow = fft + (size_t)i * C;
    double s = 0.0;
    float mg = 0.0f;
#pragma unroll 4
    for (int j = threadIdx.x; j < C; j += RED_THREADS) {
        const int jl = (j == 0) ? 1 : j - 1, jr = (j == C - 1) ? C - 2 : j + 1;       // BORDER_REFLECT101
        const float v = (j >= 42) ? lut[row[j]] : 0.0f;
        const float vl = (jl >= 42) ? lut[row[jl]] : 0.0f, vr = (jr >= 42) ? lut[row[jr]] : 0.0f;
        frow[j] = v;
        s += (double)v;
        mg = fmaxf(mg, fabsf(vr - vl));
    }
    const double bs = BR(tmp).Sum(s);
    for (int o = 16; o > 0; o >>= 1) mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, o));
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mg;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.0f;
        for (int w = 0; w < RED_THREADS / 32; ++w) m = fmaxf(m, smax[w]);
        partial[blockIdx.x] = bs;
        atomicMax(reinterpret_cast<int *>(&cnt->maxg), __float_as_int(m));      // g >= 0: int order == float order
    }
    last_block_mean(&cnt->ticket[0], partial, 1.0 / (double)((size_t)R * C), &cnt->mean, &is_last);
}

// h = s .* (1 - g / maxg) (features.cpp:206-213), all f32; Mat /= double multiplies by (float)(1/maxg).  One CTA per
// azimuth (no index division, four independent columns in flight per thread); the pass also resets the claim map the
// marking kernels use, so no later pass has to write 5 MB for that.
__global__ void __launch_bounds__(RED_THREADS)
fe_h_kernel(const float *__restrict__ fft, int R, int C, FrameCounts *__restrict__ cnt,
            float *__restrict__ h, double *__restrict__ partial, unsigned int *__restrict__ claim)
{
    __shared__ bool is_last;
    const float mean = cnt->mean;
    const float inv = (float)(1.0 / (double)cnt->maxg);
    const int i = blockIdx.x;
    const float *row = fft + (size_t)i * C;
    float *hrow = h + (size_t)i * C;
    unsigned int *crow = claim + (size_t)i * C;
    double s = 0.0;
#pragma unroll 4
    for (int j = threadIdx.x; j < C; j += RED_THREADS) {
        const float g = prewitt_abs(row, j, C) * inv;
        const float sv = row[j] - mean;
        const float hv = sv * (1.0f - g);
        hrow[j] = hv;
        crow[j] = 0xffffffffu;
        s += (double)hv;
    }
    typedef cub::BlockReduce<double, RED_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    const double bs = BR(tmp).Sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = bs;
    last_block_mean(&cnt->ticket[1], partial, 1.0 / (double)((size_t)R * C), &cnt->mean_h, &is_last);
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

// Second half of the marking loop (features.cpp:227-249), after every candidate claimed its range: candidate k
// opens a fresh region iff it still owns its whole range.  The flags go out as one bit per candidate (warp
// ballots); the block that draws the last ticket then replays the loop's bookkeeping on them -- the loop exits
// right after the increment that makes l == max_points (:231), so j_star is the position of the max_points-th
// set bit -- and settles what used to be four more launches (an inclusive scan, the j_star search, its fix-up):
//   * fewer than max_points regions in the whole list: every candidate was consumed, which the top-candidate
//     subset can only vouch for when it held them all (else need_full);
//   * the threshold carried to the next frame: the h value at rank ~1.5x the candidates this frame consumed (at
//     least 8k, at most 48k of the 64k capacity) -- small enough for the single-CTA sort in the usual case; a
//     frame that needs more than the subset holds raises need_full and is redone with the full sort.
constexpr int INCR_T = 256;
__global__ void __launch_bounds__(INCR_T)
fe_incr_kernel(const unsigned int *__restrict__ sorted_vals, const float *__restrict__ fft, int R, int C,
               FrameCounts *__restrict__ cnt, const unsigned int *__restrict__ claim,
               unsigned int *__restrict__ bits, int max_points, const float *__restrict__ sorted_keys,
               float *__restrict__ thresh)
{
    typedef cub::BlockScan<int, INCR_T> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ bool is_last;
    __shared__ int s_carry, s_jstar;
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n_sorted = cnt->n_sorted;
    bool fresh = false;
    if (k < (size_t)n_sorted) {
        const unsigned int e = sorted_vals[k];
        const int a = e / C, r = e - a * C;
        const float *row = fft + (size_t)a * C;
        int lo, hi;
        range_bounds(row, cnt->mean, R, C, r, lo, hi);
        const unsigned int *crow = claim + (size_t)a * C;
        fresh = true;
        for (int i = lo; i <= hi; ++i)
            if (crow[i] != (unsigned int)k) { fresh = false; break; }
    }
    const unsigned int word = __ballot_sync(0xffffffffu, fresh);
    if ((threadIdx.x & 31) == 0) bits[k >> 5] = word;
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(&cnt->ticket[2], 1u) == gridDim.x - 1;
        s_carry = 0;
        s_jstar = -1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int nwords = (n_sorted + 31) >> 5;                 // every word below was written (the grid covers n_sorted)
    for (int base = 0; base < nwords; base += INCR_T) {
        const int w = base + threadIdx.x;
        const unsigned int x = (w < nwords) ? __ldcg(bits + w) : 0u;
        const int c = __popc(x);
        int ex, total;
        BS(tmp).ExclusiveSum(c, ex, total);
        const int before = s_carry + ex;                     // fresh regions ahead of this word
        if (before < max_points && before + c >= max_points) {
            unsigned int y = x;                              // the (max_points - before)-th set bit
            for (int q = max_points - before - 1; q > 0; --q) y &= y - 1;
            s_jstar = w * 32 + __ffs(y) - 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
        if (s_jstar >= 0) break;                             // uniform: written before the barrier above
    }
    if (threadIdx.x != 0) return;
    int j_star = s_jstar;
    cnt->l_regions = (j_star >= 0) ? max_points : s_carry;
    if (j_star < 0) {
        if (n_sorted == cnt->n_cand) j_star = cnt->n_cand - 1;
        else cnt->need_full = 1;
    }
    cnt->j_star = j_star;
    if (cnt->need_full || n_sorted <= 0) return;
    const int want = min(max(3 * (j_star + 1) / 2, 8192), 49152);
    if (n_sorted > want) *thresh = sorted_keys[want];                                 // shrink to the target
    else if (n_sorted == cnt->n_cand) *thresh = sorted_keys[n_sorted - 1];           // everything fits: keep it all
    else if (4 * n_sorted >= 5 * (j_star + 1)) *thresh = sorted_keys[n_sorted - 1];  // 1.25x margin or better: keep
    else *thresh = cnt->mean_h + 0.5f * (sorted_keys[n_sorted - 1] - cnt->mean_h);   // thin margin: loosen
}

// ---- top-candidate subset (replaces the full-frame sort when a threshold is known) -----------------------------
// The marking loop (features.cpp:227-249) walks the candidates in descending h and stops after max_points fresh
// regions: only the first j_star + 1 of the ~6e5 candidates matter.  With a threshold T carried from the previous
// frame, the pixels with h >= T are compacted IN PIXEL ORDER (cub::DeviceSelect), sorted (64k pairs instead of
// 1.3M; stable, so ties keep pixel order exactly like the full sort) and walked.  Their ranks equal their ranks in
// the full order, so claims, regions and j_star are identical -- provided the loop terminates inside the subset
// (or the subset holds every candidate).  Otherwise need_full is raised and the host redoes the frame with the
// full sort; the result never depends on T.
// The compaction carries (pixel, h) pairs: the sort that follows reads its keys in one coalesced sweep instead of
// gathering h[pixel] pixel by pixel.
struct SubLoad {
    const float *h;
    __device__ __forceinline__ SubPair operator()(unsigned int e) const { return SubPair{e, h[e]}; }
};
struct SubSelect {
    const float *thresh;
    const FrameCounts *cnt;
    __device__ __forceinline__ bool operator()(const SubPair &p) const { return p.h > cnt->mean_h && p.h >= *thresh; }
};
typedef thrust::transform_iterator<SubLoad, thrust::counting_iterator<unsigned int>> SubPairIn;

__global__ void __launch_bounds__(RED_THREADS)
fe_prep_sub_kernel(const float *__restrict__ h, size_t N, FrameCounts *__restrict__ cnt)
{
    // the candidate count (features.cpp:220) -- read-only: fe_h_kernel has reset the claim map
    const float mean_h = cnt->mean_h;
    int c = 0;
    const size_t n4 = N >> 2;
    const float4 *h4 = reinterpret_cast<const float4 *>(h);
#pragma unroll 2
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (size_t)gridDim.x * blockDim.x) {
        const float4 v = h4[e];
        c += (v.x > mean_h) + (v.y > mean_h) + (v.z > mean_h) + (v.w > mean_h);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(N & 3)) c += h[(n4 << 2) + threadIdx.x] > mean_h;
    typedef cub::BlockReduce<int, RED_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    const int bc = BR(tmp).Sum(c);
    if (threadIdx.x == 0 && bc) atomicAdd(&cnt->n_cand, bc);
}

// Stable LSD radix sort of n <= 16384 (key, 16-bit payload) pairs held in shared memory by ONE 1024-thread CTA:
// 8-bit digits, `passes` of them from the least significant; every warp owns a contiguous segment of whole
// 32-groups and a private histogram (hist: [32 warps][256] u16), ranks inside a 32-group come from match.any, the
// (digit, warp) offsets from one block scan.  Returns the buffer (0 / 1) that holds the result.
constexpr int SORT_T = 1024;
__device__ __forceinline__ int fe_block_sort16(unsigned int *k0, unsigned int *k1, unsigned short *v0, unsigned short *v1,
                                               unsigned short *hist, int n, int passes, int *wtot)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned int *kbuf[2] = {k0, k1};
    unsigned short *vbuf[2] = {v0, v1};
    const int seg = ((n + SORT_T - 1) / SORT_T) * 32;             // whole 32-groups per warp, contiguous in order
    const int g0 = min(warp * seg, n), g1 = min(g0 + seg, n);
    unsigned short *hw = hist + warp * 256;
    const unsigned int lt = (1u << lane) - 1u;
    int cur = 0;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = 8 * pass;
        const unsigned int *sk = kbuf[cur];
        const unsigned short *sv = vbuf[cur];
        unsigned int *dk = kbuf[cur ^ 1];
        unsigned short *dv = vbuf[cur ^ 1];
        for (int i = t; i < 32 * 256 / 2; i += SORT_T) reinterpret_cast<unsigned int *>(hist)[i] = 0u;
        __syncthreads();
        for (int gb = g0; gb < g1; gb += 32) {
            const int k = gb + lane;
            const bool ok = k < g1;
            const unsigned int dig = ok ? ((sk[k] >> shift) & 255u) : 0xffffffffu;
            const unsigned int peers = __match_any_sync(0xffffffffu, dig);
            if (ok && !(peers & lt)) hw[dig] += (unsigned short)__popc(peers);
            __syncwarp();
        }
        __syncthreads();
        {   // exclusive offsets in (digit, warp) order: thread t owns digit t / 4, warps 8 * (t % 4) .. + 7
            const int d = t >> 2, w0 = (t & 3) * 8;
            int c[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j] = hist[(w0 + j) * 256 + d]; sum += c[j]; }
            int total;
            int ex = fe_block1024_exscan(sum, wtot, total);
#pragma unroll
            for (int j = 0; j < 8; ++j) { hist[(w0 + j) * 256 + d] = (unsigned short)ex; ex += c[j]; }
        }
        __syncthreads();
        for (int gb = g0; gb < g1; gb += 32) {
            const int k = gb + lane;
            const bool ok = k < g1;
            const unsigned int key = ok ? sk[k] : 0u;
            const unsigned int dig = ok ? ((key >> shift) & 255u) : 0xffffffffu;
            const unsigned int peers = __match_any_sync(0xffffffffu, dig);
            const int rank = __popc(peers & lt);
            const int base = ok ? hw[dig] : 0;
            __syncwarp();
            if (ok && rank == 0) hw[dig] = (unsigned short)(base + __popc(peers));
            __syncwarp();
            if (ok) { dk[base + rank] = key; dv[base + rank] = sv[k]; }
        }
        __syncthreads();
        cur ^= 1;
    }
    return cur;
}

// The subset's sort as ONE launch: a thread-block cluster of 8 CTAs holds the <= 64k (key, pixel) pairs in its
// distributed shared memory -- dealt EVENLY, ceil(n / 8) consecutive ranks per CTA, so that eight SMs' shared-memory
// pipes work on a pass whatever n is -- and runs a stable LSD radix sort over them: 9-bit digits, per-warp private
// histograms over contiguous segments (ranks inside a 32-group from match.any, like fe_block_sort16), per-digit CTA
// totals exchanged through DSMEM, the scatter written straight into the owning CTA's shared memory
// (st.shared::cluster).  Keys: h mapped to an ascending-sortable integer u (-0.0 == +0.0 like the reference's
// compare), then re-based to umax - u over the whole subset: descending h = ascending key, and the passes cover only
// the digits the subset's value range needs (3 passes for a range below 2^27 ulps, i.e. max(h) < ~16 x the threshold).  Equal h keep pixel order exactly as
// cub::DeviceRadixSort::SortPairsDescending over the same list would.  Two cluster barriers per pass plus one for
// the range.  Replaces fe_sub_keys_kernel + six CUB launches (histogram, scan, four onesweep passes).
namespace cg = cooperative_groups;
constexpr int SS_CL = 8, SS_T = 1024, SS_PER = FE_SUB_CAP / SS_CL;
constexpr int SS_DIG = 9, SS_NB = 1 << SS_DIG;   // digit width / bins: 512 bins x 32 warps of u16 counters = 32 KB
struct SubSortSmem {
    unsigned int key[2][SS_PER];
    unsigned int val[2][SS_PER];
    unsigned short hist[32][SS_NB];      // [warp][digit]: counts, then exclusive offsets among this CTA's warps
    unsigned int tot[SS_NB];             // this CTA's items per digit (read by the whole cluster)
    unsigned int base[SS_NB];            // global rank of this CTA's first item with the digit
    unsigned int umax, umin;             // this CTA's key range (read by the whole cluster)
};
constexpr size_t SS_SMEM = sizeof(SubSortSmem);

__global__ void __cluster_dims__(SS_CL, 1, 1) __launch_bounds__(SS_T)
fe_subsort_kernel(const SubPair *__restrict__ sel, FrameCounts *__restrict__ cnt,
                  float *__restrict__ keys_out, unsigned int *__restrict__ vals_out)
{
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ unsigned int wtot[32];
    SubSortSmem &sm = *reinterpret_cast<SubSortSmem *>(dyn);
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int n_sub = cnt->n_sub;
    const int n = min(n_sub, FE_SUB_CAP);
    if (c == 0 && t == 0) {
        cnt->n_sorted = n;
        if (n_sub > FE_SUB_CAP) cnt->need_full = 1;          // the first 64k in pixel order are not the top 64k
    }
    const int per = max(32, ((n + SS_CL * 32 - 1) / (SS_CL * 32)) * 32);      // ranks per CTA (whole 32-groups, <= SS_PER)
    const int my_n = max(0, min(per, n - c * per));
    if (t == 0) { sm.umax = 0u; sm.umin = 0xffffffffu; }
    __syncthreads();
    {
        unsigned int mx = 0u, mn = 0xffffffffu;
        for (int i = t; i < my_n; i += SS_T) {
            const SubPair p = sel[c * per + i];
            const unsigned int b = (p.h == 0.0f) ? 0u : __float_as_uint(p.h);
            const unsigned int u = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);  // ascending-sortable
            sm.key[0][i] = u;
            sm.val[0][i] = p.e;
            mx = max(mx, u); mn = min(mn, u);
        }
        mx = __reduce_max_sync(0xffffffffu, mx);
        mn = __reduce_min_sync(0xffffffffu, mn);
        if (lane == 0 && my_n > 0) { atomicMax(&sm.umax, mx); atomicMin(&sm.umin, mn); }
    }
    cluster.sync();
    unsigned int umax = 0u, umin = 0xffffffffu;
#pragma unroll
    for (int cc = 0; cc < SS_CL; ++cc) {
        umax = max(umax, *cluster.map_shared_rank(&sm.umax, cc));
        umin = min(umin, *cluster.map_shared_rank(&sm.umin, cc));
    }
    for (int i = t; i < my_n; i += SS_T) sm.key[0][i] = umax - sm.key[0][i];      // own items: no barrier needed
    const int passes = (n > 0) ? ((32 - __clz(umax - umin)) + SS_DIG - 1) / SS_DIG : 0;
    if (c == 0 && t == 0) cnt->t_tail[6] = (unsigned long long)passes;     // diagnostics (nrs_frontend_tail_times)
    if (passes == 0) cluster.sync();                          // nobody leaves while its range is still being read
    const int seg = ((my_n + SS_T - 1) / SS_T) * 32;          // whole 32-groups per warp, contiguous in order
    const int g0 = min(warp * seg, my_n), g1 = min(g0 + seg, my_n);
    unsigned short *hw = sm.hist[warp];
    const unsigned int lt = (1u << lane) - 1u;
    int cur = 0;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = SS_DIG * pass;
        const unsigned int *sk = sm.key[cur];
        const unsigned int *sv = sm.val[cur];
        for (int i = t; i < 32 * SS_NB / 2; i += SS_T) reinterpret_cast<unsigned int *>(&sm.hist[0][0])[i] = 0u;
        __syncthreads();
        for (int gb = g0; gb < g1; gb += 32) {
            const int k = gb + lane;
            const bool ok = k < g1;
            const unsigned int dig = ok ? ((sk[k] >> shift) & (SS_NB - 1)) : 0xffffffffu;
            const unsigned int peers = __match_any_sync(0xffffffffu, dig);
            if (ok && !(peers & lt)) hw[dig] += (unsigned short)__popc(peers);
            __syncwarp();
        }
        __syncthreads();
        for (int d = t >> 2; d < SS_NB; d += SS_T / 4) {
            // exclusive offsets among this CTA's warps, digit totals: four threads per digit, warps 8 * (t % 4) .. + 7
            const int w0 = (t & 3) * 8;
            unsigned int cw[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { cw[j] = sm.hist[w0 + j][d]; sum += cw[j]; }
            unsigned int inc = sum;                            // inclusive scan over the digit's four threads
            unsigned int up = __shfl_up_sync(0xffffffffu, inc, 1);
            if ((t & 3) >= 1) inc += up;
            up = __shfl_up_sync(0xffffffffu, inc, 2);
            if ((t & 3) >= 2) inc += up;
            unsigned int ex = inc - sum;
#pragma unroll
            for (int j = 0; j < 8; ++j) { sm.hist[w0 + j][d] = (unsigned short)ex; ex += cw[j]; }
            if ((t & 3) == 3) sm.tot[d] = inc;
        }
        cluster.sync();
        unsigned int all = 0, before = 0;
        if (t < SS_NB) {
#pragma unroll
            for (int cc = 0; cc < SS_CL; ++cc) {
                const unsigned int v = *cluster.map_shared_rank(&sm.tot[t], cc);
                all += v;
                if (cc < c) before += v;
            }
        }
        unsigned int all_total;
        const unsigned int ex = fe_block1024_exscan(all, wtot, all_total);
        if (t < SS_NB) sm.base[t] = ex + before;
        __syncthreads();
        unsigned int *dk = sm.key[cur ^ 1], *dv = sm.val[cur ^ 1];
        for (int gb = g0; gb < g1; gb += 32) {
            const int k = gb + lane;
            const bool ok = k < g1;
            const unsigned int key = ok ? sk[k] : 0u;
            const unsigned int dig = ok ? ((key >> shift) & (SS_NB - 1)) : 0xffffffffu;
            const unsigned int peers = __match_any_sync(0xffffffffu, dig);
            const int rank = __popc(peers & lt);
            const unsigned int loc = ok ? hw[dig] : 0u;
            __syncwarp();
            if (ok && rank == 0) hw[dig] = (unsigned short)(loc + __popc(peers));
            __syncwarp();
            if (ok) {
                const unsigned int dst = sm.base[dig] + loc + rank;            // rank in the whole list
                const unsigned int dc = dst / (unsigned int)per, di = dst - dc * (unsigned int)per;
                *cluster.map_shared_rank(&dk[di], dc) = key;
                *cluster.map_shared_rank(&dv[di], dc) = sv[k];
            }
        }
        cur ^= 1;
        cluster.sync();                                        // remote stores landed; tot / base free for the next pass
    }
    for (int i = t; i < my_n; i += SS_T) {
        const unsigned int u = umax - sm.key[cur][i];
        const unsigned int b = (u >> 31) ? (u ^ 0x80000000u) : ~u;
        keys_out[c * per + i] = __uint_as_float(b);
        vals_out[c * per + i] = sm.val[cur][i];
    }
}

__global__ void fe_set_sorted_full_kernel(FrameCounts *__restrict__ cnt)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) cnt->n_sorted = cnt->n_cand;
}

// ------------------------------------------------------------------------------------------------
// K3: per-azimuth scan (features.cpp:251-278), one warp per azimuth
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fe_scan_rows_kernel(const unsigned int *__restrict__ claim, const float *__restrict__ h, int R, int C,
                    int min_range, const FrameCounts *__restrict__ cnt, int *__restrict__ rowcount,
                    int *__restrict__ stage)
{
    // one block per azimuth.  Pass 1 (streaming, no dependent loads): one bit per column for "marked"
    // (claim <= j_star, from min_range on) and one for "a vertical neighbour is marked" (checkAdjacentMarked
    // :172-184, rows wrap).  Pass 2: thread t owns word t of the row: a run starts at a marked bit whose left
    // neighbour is clear, ends before the next clear bit (a run that reaches the last column is never closed and
    // never emitted), counts if any neighbour bit is set inside it; the survivors take getMaxInRegion's bin
    // (:186-194, int-typed running max) and go to the azimuth's staging row in column order.
    __shared__ unsigned int mw[FE_ROW_WORDS], aw[FE_ROW_WORDS];
    __shared__ int wsum[8];
    const int i = blockIdx.x, t = threadIdx.x;
    const int lane = t & 31;
    const unsigned int J = (unsigned int)cnt->j_star;
    const bool none = cnt->j_star < 0;
    const unsigned int *crow = claim + (size_t)i * C;
    const int below = (i == 0) ? R - 1 : i - 1, above = (i + 1 >= R) ? 0 : i + 1;     // :173-178
    const unsigned int *cb = claim + (size_t)below * C, *ca = claim + (size_t)above * C;
    const int nwords = (C + 31) >> 5;
#pragma unroll 4
    for (int j0 = (t >> 5) * 32; j0 < nwords * 32; j0 += 256) {
        const int j = j0 + lane;
        bool m = false, a = false;
        if (j < C && !none) {
            m = (j >= min_range) && (crow[j] <= J);
            a = (cb[j] <= J) || (ca[j] <= J);
        }
        const unsigned int mb = __ballot_sync(0xffffffffu, m), ab = __ballot_sync(0xffffffffu, a);
        if (lane == 0) { mw[j0 >> 5] = mb; aw[j0 >> 5] = ab; }
    }
    __syncthreads();
    unsigned int emit = 0u;
    if (t < nwords) {
        const unsigned int m = mw[t];
        const unsigned int starts = m & ~((m << 1) | (t ? mw[t - 1] >> 31 : 0u));
        for (unsigned int x = starts; x; x &= x - 1) {
            const int b = __ffs(x) - 1, j = 32 * t + b;
            const int p = fe_first_zero(mw, nwords, j);                // the unmarked pixel that closes the run
            if (p < C && fe_any_set(aw, j, p - 1)) emit |= 1u << b;
        }
    }
    int total;
    int pos = fe_block256_exscan(__popc(emit), wsum, total);
    const float *hrow = h + (size_t)i * C;
    int *srow = stage + (size_t)i * C;
    for (unsigned int x = emit; x; x &= x - 1) {
        const int j = 32 * t + __ffs(x) - 1;
        const int end = fe_first_zero(mw, nwords, j) - 1;
        int mx = -1000, max_r = j;
        for (int r = j; r <= end; ++r) {
            const float hv = hrow[r];
            if (hv > (float)mx) { mx = (int)hv; max_r = r; }
        }
        srow[pos++] = max_r;
    }
    if (t == 0) rowcount[i] = total;
}

// The standalone detector entry points (nrs_cen2019 / nrs_cen2018) have no tail kernel behind them: the same
// stringing of the staging rows on its own, (azimuth, range bin) pairs -> polar, total -> K0.
__global__ void __launch_bounds__(1024)
fe_gather_rows_kernel(const int *__restrict__ rowcount, const int *__restrict__ stage, int R, int C,
                      FrameCounts *__restrict__ cnt, int *__restrict__ polar, int cap)
{
    typedef cub::BlockScan<int, 1024> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ int rowoff[FE_MAX_AZ];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int v = (t < R) ? rowcount[t] : 0;
    int ex, total;
    BS(tmp).ExclusiveSum(v, ex, total);
    if (t < R) rowoff[t] = ex;
    if (t == 0) {
        cnt->K0 = min(total, cap);
        if (total > cap) cnt->overflow = 1;
    }
    __syncthreads();
    for (int a = warp; a < R; a += 32) {
        const int c = rowcount[a], off = rowoff[a];
        for (int e = lane; e < c && off + e < cap; e += 32) {
            polar[2 * (off + e)] = a;
            polar[2 * (off + e) + 1] = stage[(size_t)a * C + e];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3 tail: polar_to_cartesian_points (radar_utils.cpp:329-341; cos/sin from the host's libm table), voxelize
// (radar_loader.hpp:207-222 -> pcl::VoxelGrid<PointXYZ>::applyFilter, float32 like PCL), convert_to_bev
// (radar_utils.cpp:382-403) and ORB's border filter in ONE single-CTA launch.  At most FE_KP_CAP points: the
// voxel sort is a stable LSD radix sort held entirely in shared memory (8-bit digits, one private histogram per
// warp over a contiguous segment, ranks inside a 32-group from match.any), as many passes as the largest voxel
// index has bytes.  The keypoints arrive as one staging row of range bins per azimuth (fe_scan_rows_kernel /
// fe18_runs_kernel): the kernel scans the row counts and strings the rows together (K0, d_polar) itself.
// Replaces eight launches plus a three-pass cub::DeviceRadixSort.
// ------------------------------------------------------------------------------------------------
struct VoxGrid { int min_bx, min_by, div_x; float inv; };
constexpr int TAIL_T = 1024;
constexpr size_t FE_TAIL_SMEM = (size_t)FE_KP_CAP * (2 * 4 + 2 * 2) + 32 * 256 * 2;

__global__ void __launch_bounds__(TAIL_T)
fe_tail_kernel(const int *__restrict__ rowcount, const int *__restrict__ stage, int R, int C,
               int *__restrict__ polar, FrameCounts *__restrict__ cnt, const double *__restrict__ cossin,
               const long long *__restrict__ times, float radar_resolution, int use_vox, float leaf, int cap,
               double *__restrict__ cart_pre, long long *__restrict__ t_pre, float *__restrict__ pts32,
               float cart_resolution, int width, int patch,
               double *__restrict__ feat, float *__restrict__ kp, long long *__restrict__ t_out,
               float *__restrict__ kp_orb)
{
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ int wtot[32];
    __shared__ unsigned long long wtot3[32];
    __shared__ float smin[2][32], smax[2][32];
    __shared__ VoxGrid g;
    __shared__ unsigned int maxkey;
    __shared__ int rowoff[FE_MAX_AZ];
    unsigned int *kbuf[2] = {reinterpret_cast<unsigned int *>(dyn), reinterpret_cast<unsigned int *>(dyn) + cap};
    unsigned short *vbuf[2] = {reinterpret_cast<unsigned short *>(kbuf[1] + cap), reinterpret_cast<unsigned short *>(kbuf[1] + cap) + cap};
    unsigned short *hist = vbuf[1] + cap;                           // [warp][digit]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
#define TAIL_STAMP(i) do { if (t == 0) { unsigned long long ts_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts_)); cnt->t_tail[i] = ts_; } } while (0)
    TAIL_STAMP(0);
    // ---- the azimuths' staging rows, strung together in azimuth order (features.cpp:251-278 pushes them in
    // this order); more than cap keypoints is reported, not truncated silently ----
    int K0;
    {
        const int v = (t < R) ? rowcount[t] : 0;
        int total;
        const int ex = fe_block1024_exscan(v, wtot, total);
        if (t < R) rowoff[t] = ex;
        K0 = min(total, cap);
        if (t == 0) {
            cnt->K0 = K0;
            if (total > cap) cnt->overflow = 1;
        }
        __syncthreads();
    }

    TAIL_STAMP(1);
    // ---- polar_to_cartesian_points; PCL's float copy and its bounding box ----
    float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;
    for (int k = t; k < K0; k += TAIL_T) {
        // the keypoint's azimuth: the last row whose offset is <= k (empty rows share their successor's offset)
        int lo = 0, hi = R;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (rowoff[mid] <= k) lo = mid + 1; else hi = mid;
        }
        const int a = lo - 1;
        const int rb = stage[(size_t)a * C + (k - rowoff[a])];
        polar[2 * k] = a;
        polar[2 * k + 1] = rb;
        const double r = (double)rb * (double)radar_resolution + (double)(radar_resolution / 2);
        const double cx = r * cossin[2 * a], cy = r * cossin[2 * a + 1];
        cart_pre[2 * k] = cx;
        cart_pre[2 * k + 1] = cy;
        t_pre[k] = times[a];
        if (use_vox) {
            const float x = (float)cx, y = (float)cy;                // eigen2pcl: double -> float
            pts32[2 * k] = x; pts32[2 * k + 1] = y;
            mnx = fminf(mnx, x); mxx = fmaxf(mxx, x);
            mny = fminf(mny, y); mxy = fmaxf(mxy, y);
        }
    }
    __syncthreads();                                                 // pts32 is read by other threads below
    TAIL_STAMP(2);
    int sorted_in = 0;
    if (use_vox) {
        for (int o = 16; o > 0; o >>= 1) {
            mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        }
        if (lane == 0) { smin[0][warp] = mnx; smin[1][warp] = mny; smax[0][warp] = mxx; smax[1][warp] = mxy; }
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
            maxkey = 0u;
        }
        __syncthreads();
        unsigned int mk = 0u;
        for (int k = t; k < K0; k += TAIL_T) {
            const int i0 = (int)(floorf(pts32[2 * k] * g.inv) - (float)g.min_bx);
            const int i1 = (int)(floorf(pts32[2 * k + 1] * g.inv) - (float)g.min_by);
            const unsigned int key = (unsigned int)(i0 + i1 * g.div_x);
            kbuf[0][k] = key;
            vbuf[0][k] = (unsigned short)k;
            mk = max(mk, key);
        }
        mk = __reduce_max_sync(0xffffffffu, mk);
        if (lane == 0 && mk) atomicMax(&maxkey, mk);
        __syncthreads();
        TAIL_STAMP(3);
        const int passes = (32 - __clz(maxkey) + 7) >> 3;
        if (t == 0) cnt->t_tail[7] = (unsigned long long)passes;         // diagnostics
        const int cur = fe_block_sort16(kbuf[0], kbuf[1], vbuf[0], vbuf[1], hist, K0, passes, wtot);
        sorted_in = cur;
    }
    TAIL_STAMP(4);

    // ---- one sweep over the (voxel-sorted) points does the rest:
    //   * voxel centroids in ascending voxel index: the thread at the head of a run of equal keys sums the run's
    //     points in f32, in (stable) point order, and divides by float(count) like PCL;
    //   * convert_to_bev (radar_utils.cpp:382-403): ordered compaction, keypoints (u, v), stamps with the
    //     reference's pre-voxelisation indexing quirk (radar_loader.hpp:123-124: t_pre at the VOXEL's index);
    //   * the keypoints cv::ORB::compute keeps (radar_loader.hpp:126): it first runs KeyPointsFilter::
    //     runByImageBorder(kps, size, edgeThreshold = patch): Rect(Point(b,b), Point(W-b,W-b)).contains(
    //     Point(cvRound(u), cvRound(v))).  Keypoints with cvRound(u|v) == W - patch are dropped from kp_/desc_ while
    //     feat_in_cart_ keeps its columns -- the reference's index misalignment is reproduced, not repaired.
    // The three running counts (voxels, BEV keypoints, ORB keypoints; at most 16384 each) ride in one 64-bit scan.
    __shared__ unsigned long long carry3;
    const unsigned int *skeys = kbuf[sorted_in];
    const unsigned short *svals = vbuf[sorted_in];
    float cmr = (width / 2) * cart_resolution;
    if (width % 2 == 0) cmr = (float)((width / 2 - 0.5) * cart_resolution);
    if (t == 0) carry3 = 0ull;
    __syncthreads();
    for (int base = 0; base < K0; base += TAIL_T) {
        const int k = base + t;
        bool valid = false, keep = false, keep_orb = false;
        double x = 0, y = 0;
        float uf = 0.f, vf = 0.f;
        if (k < K0) {
            if (use_vox) {
                valid = (k == 0) || (skeys[k] != skeys[k - 1]);
                if (valid) {
                    float sx = 0.0f, sy = 0.0f;
                    int m = 0;
                    for (int q = k; q < K0 && skeys[q] == skeys[k]; ++q) {
                        const unsigned int p = svals[q];
                        sx += pts32[2 * p];
                        sy += pts32[2 * p + 1];
                        ++m;
                    }
                    const float fn = (float)m;
                    x = (double)(sx / fn);
                    y = (double)(sy / fn);
                }
            } else {
                valid = true;
                x = cart_pre[2 * k]; y = cart_pre[2 * k + 1];
            }
            if (valid) {
                const double u = ((double)cmr + y) / (double)cart_resolution;
                const double v = ((double)cmr - x) / (double)cart_resolution;
                keep = (0 < u - patch) && (u + patch < width) && (0 < v - patch) && (v + patch < width);
                uf = (float)u; vf = (float)v;
                const int ru = __float2int_rn(uf), rv = __float2int_rn(vf);
                keep_orb = keep && ru >= patch && ru < width - patch && rv >= patch && rv < width - patch;
            }
        }
        unsigned long long total;
        const unsigned long long ex = fe_block1024_exscan((valid ? 1ull : 0ull) | (keep ? 1ull << 21 : 0ull) | (keep_orb ? 1ull << 42 : 0ull),
                                                          wtot3, total);
        const unsigned long long c0 = carry3, at = c0 + ex;
        if (keep) {
            const int i = (int)(at & 0x1fffffull), j = (int)((at >> 21) & 0x1fffffull);
            feat[2 * j] = x; feat[2 * j + 1] = y;
            kp[2 * j] = uf; kp[2 * j + 1] = vf;
            t_out[j] = t_pre[i];
        }
        if (keep_orb) {
            const int j = (int)(at >> 42);
            kp_orb[2 * j] = uf; kp_orb[2 * j + 1] = vf;
        }
        __syncthreads();
        if (t == 0) carry3 = c0 + total;
        __syncthreads();
    }
    TAIL_STAMP(5);
#undef TAIL_STAMP
    if (t == 0) {
        if (use_vox) cnt->Kv = (int)(carry3 & 0x1fffffull);
        cnt->K = (int)((carry3 >> 21) & 0x1fffffull);
        cnt->Kd = (int)(carry3 >> 42);
    }
}

// ------------------------------------------------------------------------------------------------
// host-side orchestration of cen2019 on a device-resident f32 frame
// ------------------------------------------------------------------------------------------------
static int stats_blocks(const nrs_ctx *ctx) { return ctx->sm_count * 8; }

int fe_gather_rows_device(nrs_ctx *ctx, FrontendState &S, int *d_polar)
{
    fe_gather_rows_kernel<<<1, 1024, 0, ctx->stream>>>(S.d_rowcount, S.d_stage, S.R, S.C, S.d_counts, d_polar, FE_KP_CAP);
    NRS_LAUNCHED();
    return NRS_OK;
}

int fe_cen2019_device(nrs_ctx *ctx, FrontendState &S, int R, int C, int max_points, int min_range,
                      int *d_polar, bool reset_counts, bool force_full, bool stats_done)
{
    cudaStream_t st = ctx->stream;
    const size_t N = (size_t)R * C;
    const int nb = stats_blocks(ctx);
    FrameCounts *cnt = S.d_counts;
    if (reset_counts) NRS_CUDA(cudaMemsetAsync(cnt, 0, sizeof(FrameCounts), st));
    if (!stats_done) {                       // the extraction pipeline's decode has them already (fe_decode_stats_kernel)
        fe_stats_kernel<<<nb, RED_THREADS, 0, st>>>(S.d_fft, R, C, S.d_partial, cnt);   // -> maxg, mean (last block)
        NRS_LAUNCHED();
    }
    fe_h_kernel<<<R, RED_THREADS, 0, st>>>(S.d_fft, R, C, cnt, S.d_h, S.d_partial, S.d_claim);   // -> h, mean_h (last block), claim reset
    NRS_LAUNCHED();
    const bool sub = S.have_thresh && !force_full;
    const size_t n_walk = sub ? (size_t)FE_SUB_CAP : N;       // length of the sorted list the marking kernels cover
    size_t need = S.cub_bytes;
    if (sub) {
        fe_prep_sub_kernel<<<nb, RED_THREADS, 0, st>>>(S.d_h, N, cnt);
        NRS_LAUNCHED();
        SubPairIn pairs(thrust::counting_iterator<unsigned int>(0u), SubLoad{S.d_h});
        NRS_CUDA(cub::DeviceSelect::If(S.d_cub, need, pairs, S.d_sel, &cnt->n_sub, (int)N, SubSelect{S.d_thresh, cnt}, st));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        {   // opt in to the cluster sort's 146 KB of dynamic shared memory once per device
            static std::mutex cfg_mu;
            static bool configured[64] = {};
            std::lock_guard<std::mutex> lk(cfg_mu);
            const int dv = (ctx->device >= 0 && ctx->device < 64) ? ctx->device : -1;
            if (dv < 0 || !configured[dv]) {
                NRS_CUDA(cudaFuncSetAttribute(fe_subsort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SS_SMEM));
                if (dv >= 0) configured[dv] = true;
            }
        }
        fe_subsort_kernel<<<SS_CL, SS_T, SS_SMEM, st>>>(S.d_sel, cnt, S.d_keys[1], S.d_vals[1]);
        NRS_LAUNCHED();
    } else {
        fe_keys_kernel<<<nb, RED_THREADS, 0, st>>>(S.d_h, N, cnt, S.d_keys[0], S.d_vals[0], S.d_claim);
        NRS_LAUNCHED();
        fe_set_sorted_full_kernel<<<1, 32, 0, st>>>(cnt);
        NRS_LAUNCHED();
    }
    if (!sub) {
        need = S.cub_bytes;
        NRS_CUDA(cub::DeviceRadixSort::SortPairsDescending(S.d_cub, need, S.d_keys[0], S.d_keys[1], S.d_vals[0],
                                                           S.d_vals[1], (int)n_walk, 0, 32, st));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    const int tb = 256;
    const unsigned int gb = (unsigned int)((n_walk + tb - 1) / tb);
    fe_claim_kernel<<<gb, tb, 0, st>>>(S.d_vals[1], S.d_fft, R, C, cnt, S.d_claim);
    NRS_LAUNCHED();
    fe_incr_kernel<<<gb, INCR_T, 0, st>>>(S.d_vals[1], S.d_fft, R, C, cnt, S.d_claim, S.d_incr, max_points, S.d_keys[1], S.d_thresh);
    NRS_LAUNCHED();
    S.have_thresh = true;
    const int sb = R;
    fe_scan_rows_kernel<<<sb, 256, 0, st>>>(S.d_claim, S.d_h, R, C, min_range, cnt, S.d_rowcount, S.d_stage);
    NRS_LAUNCHED();
    if (d_polar) return fe_gather_rows_device(ctx, S, d_polar);        // standalone detector call; else fe_tail_kernel does it
    return NRS_OK;
}

int fe_points_voxel_bev_device(nrs_ctx *ctx, FrontendState &S, nrs_frame *f)
{
    cudaStream_t st = ctx->stream;
    const nrs_extract_params &p = S.p;
    {   // opt in to the 208 KB of dynamic shared memory once per device (contexts on several GPUs may share the process)
        static std::mutex cfg_mu;
        static bool configured[64] = {};
        std::lock_guard<std::mutex> lk(cfg_mu);
        const int dv = (ctx->device >= 0 && ctx->device < 64) ? ctx->device : -1;
        if (dv < 0 || !configured[dv]) {
            NRS_CUDA(cudaFuncSetAttribute(fe_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FE_TAIL_SMEM));
            if (dv >= 0) configured[dv] = true;
        }
    }
    fe_tail_kernel<<<1, TAIL_T, FE_TAIL_SMEM, st>>>(S.d_rowcount, S.d_stage, S.R, S.C, f->d_polar, S.d_counts, S.d_cossin, S.d_times, p.radar_resolution,
                                                   p.use_voxelization, (float)p.voxel_size, FE_KP_CAP, f->d_cart_pre,
                                                   f->d_t_pre, S.d_pts32, p.cart_resolution, p.cart_pixel_width,
                                                   p.patch_size, f->d_feat, f->d_kp, f->d_t, f->d_kp_orb);
    NRS_LAUNCHED();
    return NRS_OK;
}

int fe_decode_device(nrs_ctx *ctx, FrontendState &S, int R, int C, int pitch, bool with_stats)
{
    dim3 grid((C + 255) / 256, R);
    if (with_stats) fe_decode_stats_kernel<<<R, RED_THREADS, 0, ctx->stream>>>(S.d_raw, pitch, R, C, S.d_lut, S.d_fft, S.d_partial, S.d_counts);
    else fe_decode_kernel<<<grid, 256, 0, ctx->stream>>>(S.d_raw, pitch, R, C, S.d_fft);
    NRS_LAUNCHED();
    return NRS_OK;
}

size_t fe_cub_bytes(size_t N)
{
    size_t a = 0, b = 0, c = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, a, (float *)nullptr, (float *)nullptr, (unsigned int *)nullptr,
                                              (unsigned int *)nullptr, (int)N);
    cub::DeviceScan::InclusiveSum(nullptr, b, (unsigned int *)nullptr, (unsigned int *)nullptr, (int)N);
    size_t d = 0;
    SubPairIn pairs(thrust::counting_iterator<unsigned int>(0u), SubLoad{nullptr});
    cub::DeviceSelect::If(nullptr, d, pairs, (SubPair *)nullptr, (int *)nullptr, (int)N, SubSelect{nullptr, nullptr});
    return std::max(std::max(a, d), std::max(b, c)) + 256;
}

}  // namespace nrs
