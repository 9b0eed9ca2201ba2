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
            const float nqp = exp(-0.5 * pow((double)((qv - pv) / sigma_q[i]), 2));
            const float npp = exp(-0.5 * pow((double)(pv / sigma_q[i]), 2));
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
