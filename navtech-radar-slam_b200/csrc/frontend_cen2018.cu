// frontend_cen2018.cu -- cen2018features (outlier-robust-radar-odometry/src/features.cpp:44-131), the
// keypoint extractor the shipped launch file selects (run_orora.launch:11).
//
// Arithmetic contract (what makes the output identical to the reference's):
//   * per-azimuth bias (:49-59) and noise level (:77-91) are SEQUENTIAL float sums in column order; they
//     are kept sequential here (one thread per azimuth walks its row, staged whole in shared memory by its
//     CTA), because any other order rounds differently;
//   * the Gaussian taps (:62-73) are computed on the host with the same libm exp() the reference links;
//   * cv::filter2D with that 1 x 3*sigma kernel (:74) is OpenCV's direct filter engine: per output pixel one
//     forward chain acc = fma(src[j + k - anchor], tap[k], acc) over k, starting from 0, BORDER_REFLECT101;
//     the last C % 8 columns of a row are its scalar tail, acc = acc + src * tap without fusion (both radar
//     widths, 3360 and 3768, are multiples of 8).  Established against cv2 4.13 (AVX2 build) in
//     tests/test_oracle_frontend.py::test_cen2018_filter_rule_and_run_centres; the oracle runs the real cv2;
//   * the per-pixel response (:97-101) is float arithmetic without contraction (this library is built with
//     --fmad=false); pow(float, 2) promotes to double and squares exactly; exp() is evaluated in double and
//     rounded to float.  exp is the one libm call on the per-pixel path: CUDA's and glibc's double exp are
//     both within 1 ulp but not bit-identical, so a last-bit difference survives the rounding to float with
//     probability ~2^-28 per pixel and then matters only if y lies within one float ulp of the threshold.
//   * peak centres (:102-112): runs of consecutive columns with y > thres, closed by a column that fails the
//     test or by the end of the row, emit their middle element run[len / 2], azimuth-major (:118-128).
#include <math.h>

#include "frontend.cuh"

namespace nrs {

namespace {

constexpr int ROWS_PB = 1;       // azimuths per CTA in the bias / noise kernel (one CTA per azimuth: the staging and the q pass of a row spread over 256 threads, the two add chains are the floor)

// :49-59 + :77-91.  q = f - mean(row), sigma_q(row); both sums are sequential in column order, so the
// floor is two dependent chains of C float adds per azimuth.  Each CTA stages its rows whole in shared
// memory (one coalesced read of the frame), one thread per row walks it with 16-byte loads kept ahead of the
// add chain, all threads form q in place and write it out, the walkers go over q once more.
__global__ void __launch_bounds__(256)
fe18_bias_sigma_kernel(const float *__restrict__ fft, int R, int C, int stride, float *__restrict__ q,
                       float *__restrict__ sigma_q)
{
    extern __shared__ __align__(16) float rows[];            // ROWS_PB x stride (stride = C rounded up to 4, + 4)
    __shared__ float s_mean[ROWS_PB];
    const int t = threadIdx.x, row0 = blockIdx.x * ROWS_PB;
    const int nr = min(ROWS_PB, R - row0);
    for (int r = 0; r < nr; ++r) {
        const float *src = fft + (size_t)(row0 + r) * C;
        float *dst = rows + r * stride;
        for (int c = t; c < stride; c += 256) dst[c] = (c < C) ? src[c] : 0.f;
    }
    __syncthreads();
    const bool walker = t < nr;
    const float4 *mine = reinterpret_cast<const float4 *>(rows + (walker ? t : 0) * stride);
    const int c4 = C >> 2;
    if (walker) {
        float acc = 0.f;
        for (int k = 0; k < c4; ++k) {                          // mean += fft_data.at<float>(i, j)
            const float4 a = mine[k];
            acc += a.x; acc += a.y; acc += a.z; acc += a.w;
        }
        for (int c = c4 * 4; c < C; ++c) acc += rows[t * stride + c];
        s_mean[t] = acc / (float)C;                             // mean /= fft_data.cols
    }
    __syncthreads();
    for (int r = 0; r < nr; ++r) {
        const float m = s_mean[r];
        float *dst = rows + r * stride;
        float *qo = q + (size_t)(row0 + r) * C;
        for (int c = t; c < C; c += 256) {
            const float v = dst[c] - m;
            dst[c] = v;
            qo[c] = v;
        }
    }
    __syncthreads();
    if (walker) {
        float sig = 0.f;
        int nonzero = 0;
        // adding +0.0f to the non-negative running sum changes nothing, so the test of :82 becomes a select
#define NRS_SIG_STEP(v) { const float a_ = 2 * ((v) * (v)); const bool neg_ = (v) < 0.f; sig = sig + (neg_ ? a_ : 0.f); nonzero += neg_; }
        for (int k = 0; k < c4; ++k) {
            const float4 a = mine[k];
            NRS_SIG_STEP(a.x) NRS_SIG_STEP(a.y) NRS_SIG_STEP(a.z) NRS_SIG_STEP(a.w)
        }
        for (int c = c4 * 4; c < C; ++c) { const float v = rows[t * stride + c]; NRS_SIG_STEP(v) }
#undef NRS_SIG_STEP
        sigma_q[row0 + t] = nonzero ? sqrtf(sig / (float)nonzero) : 0.034f;   // :87-90
    }
}

__device__ __forceinline__ int reflect101(int j, int C)
{
    if (C == 1) return 0;
    while (j < 0 || j >= C) j = (j < 0) ? -j : 2 * C - 2 - j;
    return j;
}

// :74 cv::filter2D(q, p, -1, filter, Point(-1,-1), 0, BORDER_REFLECT101): forward FMA chain per pixel.
// Each thread produces four adjacent columns: the 16 taps of a chunk and the 19 inputs they touch come in by
// 16-byte shared-memory loads and serve 64 multiply-adds (the one-column version spent two loads per FMA).
constexpr int FILT_THREADS = 256;
constexpr int FILT_COLS = FILT_THREADS * 4;      // columns per CTA
constexpr int FILT_MAX = 255;                    // taps (3 * sigma_gauss, odd)
constexpr int FILT_KPAD = 256;                   // taps padded to whole chunks of 16
__global__ void __launch_bounds__(FILT_THREADS)
fe18_filter_kernel(const float *__restrict__ q, int R, int C, const float *__restrict__ taps, int fsize,
                   float *__restrict__ p)
{
    __shared__ __align__(16) float row[FILT_COLS + FILT_KPAD + 4];
    __shared__ __align__(16) float kf[FILT_KPAD];
    const int i = blockIdx.y, j0 = blockIdx.x * FILT_COLS, t = threadIdx.x;
    const int anchor = fsize / 2;
    const int nchunk = (fsize + 15) >> 4;
    for (int k = t; k < nchunk * 16; k += FILT_THREADS) kf[k] = (k < fsize) ? taps[k] : 0.f;
    const float *qr = q + (size_t)i * C;
    for (int e = t; e < FILT_COLS + nchunk * 16 + 4; e += FILT_THREADS)
        row[e] = (e < FILT_COLS + fsize - 1) ? qr[reflect101(j0 + e - anchor, C)] : 0.f;
    __syncthreads();
    const int body = C & ~7;                         // OpenCV's 8-wide vector body; beyond it its scalar tail
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float4 *rw = reinterpret_cast<const float4 *>(row + 4 * t);
    const float4 *kw = reinterpret_cast<const float4 *>(kf);
    for (int ch = 0; ch < nchunk; ++ch) {
        float x[20], w[16];
#pragma unroll
        for (int v = 0; v < 5; ++v) { const float4 a = rw[ch * 4 + v]; x[4 * v] = a.x; x[4 * v + 1] = a.y; x[4 * v + 2] = a.z; x[4 * v + 3] = a.w; }
#pragma unroll
        for (int v = 0; v < 4; ++v) { const float4 a = kw[ch * 4 + v]; w[4 * v] = a.x; w[4 * v + 1] = a.y; w[4 * v + 2] = a.z; w[4 * v + 3] = a.w; }
        // a thread's four columns lie on one side of the body / tail border (both are multiples of 4)
        if (j0 + 4 * t < body) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (ch * 16 + k < fsize) {
#pragma unroll
                    for (int o = 0; o < 4; ++o) acc[o] = fmaf(x[k + o], w[k], acc[o]);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (ch * 16 + k < fsize) {
#pragma unroll
                    for (int o = 0; o < 4; ++o) acc[o] = acc[o] + x[k + o] * w[k];
                }
            }
        }
    }
    const int jb = j0 + 4 * t;
    if ((C & 3) == 0 && jb + 3 < C) {
        *reinterpret_cast<float4 *>(p + (size_t)i * C + jb) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (jb + o < C) p[(size_t)i * C + jb + o] = acc[o];
    }
}


/* ---- exp() of a double, narrowed to float, INDEPENDENT OF THE PLATFORM'S libm --------------------------------
 * The reference stores exp(double) into a float (features.cpp:98-99).  libm's exp is not correctly rounded (glibc:
 * < 0.52 ulp, CUDA: <= 1 ulp), so two platforms can disagree in the last bit of the double -- which reaches the
 * float only when the double lands within a couple of ulps of a float rounding boundary (a mid-point between two
 * floats).  Exactly there, and for results in the float-subnormal range, the value is recomputed in double-double
 * arithmetic (~100 bits: Cody-Waite reduction by ln2, x / 512, degree-8 Taylor series, nine squarings of
 * (1 + s)^2 - 1) and rounded once: everywhere the result is float(correctly rounded double exp(x)).  The same
 * operation sequence is compiled on the CPU (-ffp-contract=off) and on the GPU (--fmad=false), so both sides agree
 * bit for bit by construction.  Valid for x <= 0 (the only use: x = -0.5 t^2). */
typedef struct { double hi, lo; } dd_t;
__device__ __forceinline__ dd_t dd_fast_two_sum(double a, double b) { dd_t r; r.hi = a + b; r.lo = b - (r.hi - a); return r; }
__device__ __forceinline__ dd_t dd_two_sum(double a, double b)
{
    dd_t r; r.hi = a + b; const double bb = r.hi - a; r.lo = (a - (r.hi - bb)) + (b - bb); return r;
}
__device__ __forceinline__ dd_t dd_two_prod(double a, double b)
{
    const double ca = 134217729.0 * a, cb = 134217729.0 * b;
    const double ah = ca - (ca - a), al = a - ah, bh = cb - (cb - b), bl = b - bh;
    dd_t r; r.hi = a * b; r.lo = ((ah * bh - r.hi) + ah * bl + al * bh) + al * bl; return r;
}
__device__ __forceinline__ dd_t dd_add(dd_t a, dd_t b) { dd_t s = dd_two_sum(a.hi, b.hi); s.lo += a.lo + b.lo; return dd_fast_two_sum(s.hi, s.lo); }
__device__ __forceinline__ dd_t dd_mul(dd_t a, dd_t b)
{
    dd_t p = dd_two_prod(a.hi, b.hi); p.lo += a.hi * b.lo + a.lo * b.hi; return dd_fast_two_sum(p.hi, p.lo);
}
__device__ __noinline__ double exp_dd(double x)
{
    if (!(x > -800.0)) return 0.0;
    if (x >= 0.0) return 1.0;
    const double k = rint(x * 1.4426950408889634);
    /* r = x - k ln2, ln2 = 0x1.62e42fefa39efp-1 + 0x1.abc9e3b39803fp-56 */
    dd_t kl = dd_two_prod(k, 0.6931471805599453);
    kl.lo += k * 2.3190468138462996e-17;
    kl = dd_fast_two_sum(kl.hi, kl.lo);
    dd_t neg; neg.hi = -kl.hi; neg.lo = -kl.lo;
    dd_t xx; xx.hi = x; xx.lo = 0.0;
    dd_t r = dd_add(xx, neg);
    r.hi *= 0.001953125; r.lo *= 0.001953125;               /* / 512, exact */
    /* s = exp(r) - 1 = r (1 + r/2 (1 + r/3 (... (1 + r/8)))) */
    const double inv[9] = {0, 1.0, 0.5, 0.3333333333333333, 0.25, 0.2, 0.16666666666666666, 0.14285714285714285, 0.125};
    const double invl[9] = {0, 0, 0, 1.850371707708594e-17, 0, -1.1102230246251566e-17, 9.25185853854297e-18, 7.93016446160826e-18, 0};
    dd_t one; one.hi = 1.0; one.lo = 0.0;
    dd_t acc = one;
    for (int n = 8; n >= 2; --n) {
        dd_t c; c.hi = inv[n]; c.lo = invl[n];
        acc = dd_add(one, dd_mul(dd_mul(r, c), acc));
    }
    dd_t s = dd_mul(r, acc);
    for (int i = 0; i < 9; ++i) {                            /* (1 + s)^2 - 1 = 2 s + s^2 */
        dd_t s2 = dd_mul(s, s);
        dd_t t; t.hi = 2.0 * s.hi; t.lo = 2.0 * s.lo;
        s = dd_add(t, s2);
    }
    const dd_t e = dd_add(one, s);
    return ldexp(e.hi, (int)k);                           /* |k| <= 1155: the scaling is exact above the double-subnormal range */
}
__device__ __forceinline__ float exp_to_float(double x)
{
    double e = exp(x);
    unsigned long long b;
    b = (unsigned long long)__double_as_longlong(e);
    const long long low = (long long)(b & 0x1fffffffull) - 0x10000000ll;      /* distance to the float mid-point, in double ulps */
    if ((low >= -2 && low <= 2) || (e < 1.2e-38 && x > -110.0)) e = exp_dd(x);
    return (float)e;
}

// :95-103 the per-pixel response and its threshold test (flag = y > zq * sigma_q[i]); y optional (tests)
__global__ void __launch_bounds__(256)
fe18_flags_kernel(const float *__restrict__ q, const float *__restrict__ p, const float *__restrict__ sigma_q,
                  int R, int C, float zq, int min_range, unsigned char *__restrict__ flags, float *__restrict__ y_out)
{
    const int i = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
    if (j >= C) return;
    const size_t e = (size_t)i * C + j;
    unsigned char f = 0;
    if (j >= min_range) {
        const float sg = sigma_q[i];
        const float thres = zq * sg;
        const float qv = q[e], pv = p[e];
        const double a = (double)((qv - pv) / sg), b2 = (double)(pv / sg);
        const float nqp = exp_to_float(-0.5 * (a * a));
        const float npp = exp_to_float(-0.5 * (b2 * b2));
        const float b = nqp - npp;
        const float y = qv * (1 - nqp) + pv * b;
        if (y_out) y_out[e] = y;
        f = y > thres;
    } else if (y_out) {
        y_out[e] = 0.f;
    }
    flags[e] = f;
}

// :102-112 one CTA per azimuth: a run starts at a flagged column whose left neighbour (from min_range on) is
// not flagged; it emits column start + len / 2.  Emission stays ordered through per-warp ballots.
__global__ void __launch_bounds__(256)
fe18_runs_kernel(const unsigned char *__restrict__ flags, int R, int C, int min_range, int *__restrict__ rowcount,
                 int *__restrict__ stage)
{
    // pass 1: the azimuth's flags (from min_range on) as one bit per column in shared memory; pass 2: thread t owns
    // word t, run starts = set bits with a clear left neighbour, run end = before the next clear bit (or the last
    // column); the staging row takes start + len / 2 in column order (fe_tail_kernel strings the rows)
    __shared__ unsigned int mw[FE_ROW_WORDS];
    __shared__ int wsum[8];
    const int i = blockIdx.x, t = threadIdx.x, lane = t & 31;
    const unsigned char *fr = flags + (size_t)i * C;
    const int nwords = (C + 31) >> 5;
#pragma unroll 4
    for (int j0 = (t >> 5) * 32; j0 < nwords * 32; j0 += 256) {
        const int j = j0 + lane;
        const bool m = (j < C) && (j >= min_range) && fr[j];
        const unsigned int mb = __ballot_sync(0xffffffffu, m);
        if (lane == 0) mw[j0 >> 5] = mb;
    }
    __syncthreads();
    unsigned int starts = 0u;
    if (t < nwords) {
        const unsigned int m = mw[t];
        starts = m & ~((m << 1) | (t ? mw[t - 1] >> 31 : 0u));
    }
    int total;
    int pos = fe_block256_exscan(__popc(starts), wsum, total);
    int *srow = stage + (size_t)i * C;
    for (unsigned int x = starts; x; x &= x - 1) {
        const int j = 32 * t + __ffs(x) - 1;
        const int end = min(fe_first_zero(mw, nwords, j), C) - 1;
        srow[pos++] = j + (end - j + 1) / 2;                        // peak_points[peak_points.size() / 2]
    }
    if (t == 0) rowcount[i] = total;
}

}  // namespace

int fe_gather_rows_device(nrs_ctx *ctx, FrontendState &S, int *d_polar);   // frontend_cen2019.cu

// host libm taps, the reference's expression (:62-73); returns the number of taps
int fe_cen2018_taps_host(int sigma_gauss, float *taps)
{
    const int fsize = sigma_gauss * 3, mu = fsize / 2;
    const float sig_sqr = (float)(sigma_gauss * sigma_gauss);
    float s = 0;
    for (int i = 0; i < fsize; ++i) {
        taps[i] = (float)exp(-0.5 * (i - mu) * (i - mu) / sig_sqr);
        s += taps[i];
    }
    // filter /= s: cv::Mat::operator/=(double) is convertTo(-1, 1. / s), i.e. a multiplication by (float)(1 / s)
    const float inv = (float)(1.0 / (double)s);
    for (int i = 0; i < fsize; ++i) taps[i] = taps[i] * inv;
    return fsize;
}

// d_fft (S.d_fft) -> the azimuths' staging rows (S.d_stage, S.d_rowcount); with d_polar also -> d_polar + cnt->K0.  Scratch reuse: q in S.d_h, p in d_keys[0], flags in d_claim (bytes),
// y (optional, tests) in d_keys[1], sigma_q + taps in d_partial's neighbour d_incr.
int fe_cen2018_device(nrs_ctx *ctx, FrontendState &S, int R, int C, float zq, int sigma_gauss, int min_range,
                      int *d_polar, bool reset_counts, bool keep_y)
{
    cudaStream_t st = ctx->stream;
    FrameCounts *cnt = S.d_counts;
    if (reset_counts) NRS_CUDA(cudaMemsetAsync(cnt, 0, sizeof(FrameCounts), st));
    static_assert(FILT_MAX <= 256, "FrontendState::h_taps");
    float *taps = S.h_taps;                                  // persistent: a captured graph re-reads it at every replay
    if (sigma_gauss < 1 || sigma_gauss * 3 > FILT_MAX || sigma_gauss % 2 != 1) {
        set_error("cen2018: sigma_gauss must be odd and 3 * sigma_gauss <= %d (features.cpp:63 asserts odd)", FILT_MAX);
        return NRS_EINVAL;
    }
    int fsize = sigma_gauss * 3;
    if (S.taps_sigma != sigma_gauss) { fsize = fe_cen2018_taps_host(sigma_gauss, taps); S.taps_sigma = sigma_gauss; }
    float *d_q = S.d_h, *d_p = S.d_keys[0], *d_y = keep_y ? S.d_keys[1] : nullptr;
    unsigned char *d_flags = reinterpret_cast<unsigned char *>(S.d_claim);
    float *d_sigma = reinterpret_cast<float *>(S.d_incr);
    float *d_taps = d_sigma + FE_MAX_AZ;
    NRS_CUDA(cudaMemcpyAsync(d_taps, taps, fsize * sizeof(float), cudaMemcpyHostToDevice, st));
    {
        const int stride = ((C + 3) & ~3) + 4;               // rows 16-byte aligned, walkers on different banks
        const size_t smem = (size_t)ROWS_PB * stride * sizeof(float);
        // per device and thread-safe (contexts on several GPUs may share this process)
        static std::mutex cfg_mu;
        static size_t configured[64] = {};
        {
            std::lock_guard<std::mutex> lk(cfg_mu);
            const int dv = (ctx->device >= 0 && ctx->device < 64) ? ctx->device : 0;
            if (smem > configured[dv] || ctx->device >= 64) {
                NRS_CUDA(cudaFuncSetAttribute(fe18_bias_sigma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                configured[dv] = smem;
            }
        }
        fe18_bias_sigma_kernel<<<(R + ROWS_PB - 1) / ROWS_PB, 256, smem, st>>>(S.d_fft, R, C, stride, d_q, d_sigma);
        NRS_LAUNCHED();
    }
    dim3 gf((C + FILT_COLS - 1) / FILT_COLS, R);
    fe18_filter_kernel<<<gf, FILT_THREADS, 0, st>>>(d_q, R, C, d_taps, fsize, d_p);
    NRS_LAUNCHED();
    dim3 gp((C + 255) / 256, R);
    fe18_flags_kernel<<<gp, 256, 0, st>>>(d_q, d_p, d_sigma, R, C, zq, min_range, d_flags, d_y);
    NRS_LAUNCHED();
    fe18_runs_kernel<<<R, 256, 0, st>>>(d_flags, R, C, min_range, S.d_rowcount, S.d_stage);
    NRS_LAUNCHED();
    if (d_polar) return fe_gather_rows_device(ctx, S, d_polar);        // standalone detector call; else fe_tail_kernel does it
    return NRS_OK;
}

}  // namespace nrs
