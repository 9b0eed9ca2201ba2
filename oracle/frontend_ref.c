/*
 * oracle/frontend_ref.c -- C part of the CPU restatement of the ORORA front-end (TEST
 * INFRASTRUCTURE ONLY; see oracle/__init__.py).  The OpenCV stages (filter2D, mean, remap,
 * convertTo, GaussianBlur, ORB, BFMatcher) are driven from oracle/frontend_ref.py with the
 * REAL cv2 calls; this file holds the loops that are the reference's own code:
 *
 *   outlier-robust-radar-odometry/src/radar_utils.cpp:227-250  get_azimuth_index
 *                                                   :252-298  polar_to_cartesian_radar_img maps
 *   outlier-robust-radar-odometry/src/features.cpp:151-194     findRangeBoundaries,
 *                                                              checkAdjacentMarked, getMaxInRegion
 *                                                 :226-249     sequential region marking loop
 *                                                 :251-292     per-azimuth scan
 *
 * libm calls (atan2f, sqrt, pow) are the same glibc the reference links, which is the point of
 * doing this in C rather than numpy.  Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* radar_utils.cpp:227-250 */
static double get_azimuth_index(const double *azimuths, int M, double azimuth)
{
    double mind = 1000;
    double closest = 0;
    for (int i = 0; i < M; ++i) {
        double d = fabs(azimuths[i] - azimuth);
        if (d < mind) { mind = d; closest = i; }
    }
    int ci = (int)closest;
    if (azimuths[ci] < azimuth) {
        double delta = 0;
        if (closest < M - 1) delta = (azimuth - azimuths[ci]) / (azimuths[ci + 1] - azimuths[ci]);
        closest += delta;
    } else if (azimuths[ci] > azimuth) {
        double delta = 0;
        if (closest > 0) delta = (azimuths[ci] - azimuth) / (azimuths[ci] - azimuths[ci - 1]);
        closest -= delta;
    }
    return closest;
}

/* the restatement above, callable on its own (pinned against the reference's compiled get_azimuth_index) */
double fe_get_azimuth_index(const double *azimuths, int M, double azimuth) { return get_azimuth_index(azimuths, M, azimuth); }

/* radar_utils.cpp:257-298: the static part (x, y, range, theta) and the per-frame angle map.
 * range/angle/theta are width*width float32, row-major.  cir204 selects the MulRan branch
 * (:292-293) vs the uniform-step branch (:295). */
void fe_build_maps(const double *azimuths, int n_az, float radar_resolution, float cart_resolution,
                   int width, int cir204, float *range, float *angle, float *theta_out)
{
    float cart_min_range = (width / 2) * cart_resolution;
    if (width % 2 == 0) cart_min_range = (width / 2 - 0.5) * cart_resolution;
    double azimuth_step = azimuths[1] - azimuths[0];
    for (int i = 0; i < width; ++i) {
        for (int j = 0; j < width; ++j) {
            float x = cart_min_range - i * cart_resolution;           /* map_x :273 */
            float y = -1 * cart_min_range + j * cart_resolution;      /* map_y :267 */
            float r = (sqrt(pow(x, 2) + pow(y, 2)) - radar_resolution / 2) / radar_resolution;
            if (r < 0) r = 0;
            range[(size_t)i * width + j] = r;
            float theta = atan2f(y, x);
            if (theta < 0) theta += 2 * M_PI;
            if (theta_out) theta_out[(size_t)i * width + j] = theta;
            if (cir204) angle[(size_t)i * width + j] = get_azimuth_index(azimuths, n_az, theta);
            else        angle[(size_t)i * width + j] = (theta - azimuths[0]) / azimuth_step;
        }
    }
}

/* features.cpp:151-170.  NOTE the quirk kept from the reference: the upward extension is
 * guarded by s.rows (the number of azimuths), not s.cols. */
static void find_range_boundaries(const float *s, int rows, int cols, int a, int r, int *rlow, int *rhigh)
{
    *rlow = r;
    *rhigh = r;
    if (r > 0) {
        for (int i = r - 1; i >= 0; i--) {
            if (s[(size_t)a * cols + i] < 0) *rlow = i;
            else break;
        }
    }
    if (r < rows - 1) {
        for (int i = r + 1; i < cols; i++) {
            if (s[(size_t)a * cols + i] < 0) *rhigh = i;
            else break;
        }
    }
}

/* features.cpp:226-249.  cand_* = candidates already sorted by descending h (stable).
 * R = rows*cols bytes (0/1).  Returns l (regions counted); *j_out = candidates consumed. */
int fe_cen2019_mark(const float *s, int rows, int cols, const int *cand_a, const int *cand_r,
                    int n_cand, int max_points, unsigned char *R, int *j_out)
{
    int false_count = rows * cols;
    int j = 0, l = 0;
    memset(R, 0, (size_t)rows * cols);
    while (l < max_points && j < n_cand && false_count > 0) {
        int a = cand_a[j], r = cand_r[j];
        if (!R[(size_t)a * cols + r]) {
            int rlow = r, rhigh = r;
            find_range_boundaries(s, rows, cols, a, r, &rlow, &rhigh);
            int already_marked = 0;
            for (int i = rlow; i <= rhigh; i++) {
                if (R[(size_t)a * cols + i]) { already_marked = 1; continue; }
                R[(size_t)a * cols + i] = 1;
                false_count--;
            }
            if (!already_marked) l++;
        }
        j++;
    }
    *j_out = j;
    return l;
}

/* features.cpp:172-194 + :251-292.  Emits (azimuth, range) pairs azimuth-major. */
int fe_cen2019_scan(const unsigned char *R, const float *h, int rows, int cols, int min_range,
                    int *out_a, int *out_r, int cap)
{
    int k = 0;
    for (int i = 0; i < rows; i++) {
        int start = 0, end = 0, counting = 0;
        for (int j = min_range; j < cols; j++) {
            if (R[(size_t)i * cols + j]) {
                if (!counting) { start = j; end = j; counting = 1; }
                else end = j;
            } else if (counting) {
                int below = i - 1, above = i + 1;
                if (below < 0) below = rows - 1;
                if (above >= rows) above = 0;
                int adj = 0;
                for (int r = start; r <= end; r++)
                    if (R[(size_t)below * cols + r] || R[(size_t)above * cols + r]) { adj = 1; break; }
                if (adj) {
                    int max_r = start;
                    int max = -1000;                       /* int running max: features.cpp:187 */
                    for (int r = start; r <= end; r++) {
                        if (h[(size_t)i * cols + r] > max) {
                            max = h[(size_t)i * cols + r];   /* float -> int truncation, kept */
                            max_r = r;
                        }
                    }
                    if (k < cap) { out_a[k] = i; out_r[k] = max_r; }
                    k++;
                }
                counting = 0;
            }
        }
    }
    return k;
}

/* 64-bit Hamming distance helper for the brute-force matcher cross-check */
int fe_hamming256(const unsigned char *a, const unsigned char *b)
{
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount(a[i] ^ b[i]);
    return d;
}

/* ---------------------------------------------------------------------------------------------
 * cen2018features (outlier-robust-radar-odometry/src/features.cpp:44-131), the reference's own
 * loops; the 1 x 3*sigma Gaussian cv::filter2D in the middle (:74) is run by the real cv2 from
 * oracle/frontend_ref.py.  All arithmetic is float unless the C++ promotes it (noted).
 * --------------------------------------------------------------------------------------------- */

/* :49-59 per-azimuth bias: sequential float sum, mean /= cols, q = f - mean */
void fe_cen2018_bias(const float *fft, int rows, int cols, float *q)
{
    for (int i = 0; i < rows; ++i) {
        float mean = 0;
        for (int j = 0; j < cols; ++j) mean += fft[(size_t)i * cols + j];
        mean /= cols;
        for (int j = 0; j < cols; ++j) q[(size_t)i * cols + j] = fft[(size_t)i * cols + j] - mean;
    }
}

/* :62-72 the filter taps: exp() of a double argument stored to float, float running sum.
 * (filter /= s is a cv::Mat operation: done by cv2 in the Python driver.) */
float fe_cen2018_taps(int sigma_gauss, float *taps)
{
    const int fsize = sigma_gauss * 3, mu = fsize / 2;
    const float sig_sqr = sigma_gauss * sigma_gauss;
    float s = 0;
    for (int i = 0; i < fsize; ++i) {
        taps[i] = exp(-0.5 * (i - mu) * (i - mu) / sig_sqr);
        s += taps[i];
    }
    return s;
}

/* :77-91 noise level per azimuth from the negative half of q */
void fe_cen2018_sigma(const float *q, int rows, int cols, float *sigma_q)
{
    for (int i = 0; i < rows; ++i) {
        int nonzero = 0;
        float acc = 0;
        for (int j = 0; j < cols; ++j) {
            const float n = q[(size_t)i * cols + j];
            if (n < 0) { acc += 2 * (n * n); nonzero++; }
        }
        if (nonzero) acc = sqrt(acc / nonzero);
        else acc = 0.034;
        sigma_q[i] = acc;
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
static dd_t dd_fast_two_sum(double a, double b) { dd_t r; r.hi = a + b; r.lo = b - (r.hi - a); return r; }
static dd_t dd_two_sum(double a, double b)
{
    dd_t r; r.hi = a + b; const double bb = r.hi - a; r.lo = (a - (r.hi - bb)) + (b - bb); return r;
}
static dd_t dd_two_prod(double a, double b)
{
    const double ca = 134217729.0 * a, cb = 134217729.0 * b;
    const double ah = ca - (ca - a), al = a - ah, bh = cb - (cb - b), bl = b - bh;
    dd_t r; r.hi = a * b; r.lo = ((ah * bh - r.hi) + ah * bl + al * bh) + al * bl; return r;
}
static dd_t dd_add(dd_t a, dd_t b) { dd_t s = dd_two_sum(a.hi, b.hi); s.lo += a.lo + b.lo; return dd_fast_two_sum(s.hi, s.lo); }
static dd_t dd_mul(dd_t a, dd_t b)
{
    dd_t p = dd_two_prod(a.hi, b.hi); p.lo += a.hi * b.lo + a.lo * b.hi; return dd_fast_two_sum(p.hi, p.lo);
}
static double exp_dd(double x)
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
static float exp_to_float(double x)
{
    double e = exp(x);
    unsigned long long b;
    memcpy(&b, &e, sizeof(b));
    const long long low = (long long)(b & 0x1fffffffull) - 0x10000000ll;      /* distance to the float mid-point, in double ulps */
    if ((low >= -2 && low <= 2) || (e < 1.2e-38 && x > -110.0)) e = exp_dd(x);
    return (float)e;
}

/* :94-113 peak centres per azimuth; y is also written out (nullable) for stage-wise comparison.
 * Returns the number of targets; out_a/out_r hold (azimuth, range bin) azimuth-major like :118-128. */
int fe_cen2018_scan(const float *q, const float *p, const float *sigma_q, int rows, int cols, float zq,
                    int min_range, float *y_out, int *out_a, int *out_r, int cap)
{
    int k = 0;
    int *peak = (int *)malloc(sizeof(int) * (size_t)(cols > 0 ? cols : 1));
    for (int i = 0; i < rows; ++i) {
        int np = 0;
        const float thres = zq * sigma_q[i];
        for (int j = min_range; j < cols; ++j) {
            const float qv = q[(size_t)i * cols + j], pv = p[(size_t)i * cols + j];
            /* pow(float, int) promotes to double in C++11: the float quotient is squared exactly */
            const float nqp = exp_to_float(-0.5 * pow((double)((qv - pv) / sigma_q[i]), 2));
            const float npp = exp_to_float(-0.5 * pow((double)(pv / sigma_q[i]), 2));
            const float b = nqp - npp;
            const float y = qv * (1 - nqp) + pv * b;
            if (y_out) y_out[(size_t)i * cols + j] = y;
            if (y > thres) {
                peak[np++] = j;
            } else if (np > 0) {
                if (k < cap) { out_a[k] = i; out_r[k] = peak[np / 2]; }
                ++k;
                np = 0;
            }
        }
        if (np > 0) {
            if (k < cap) { out_a[k] = i; out_r[k] = peak[np / 2]; }
            ++k;
        }
    }
    free(peak);
    return k;
}

/* test hooks: the double-double exp and the float narrowing, checked against an arbitrary-precision exp */
double oracle_exp_dd(double x) { return exp_dd(x); }
float oracle_exp_to_float(double x) { return exp_to_float(x); }
