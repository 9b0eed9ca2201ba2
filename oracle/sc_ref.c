/*
 * oracle/sc_ref.c -- CPU restatement of the reference ScanContext loop detector.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under navtech-radar-slam_b200/ may include,
 * link or call this file.  It is imported by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs, where it plays the checker
 * and the timed CPU baseline -- never the product.
 *
 * PARITY PIN STATUS: the reference ships no tests, fixtures or golden vectors for
 * this path and its SCManager cannot be compiled here
 * (Eigen, PCL, OpenCV-C++ absent).  The kd-tree candidate step IS pinned against the
 * reference's own vendored nanoflann (oracle/ref_kdtree_wrap.cpp -> oracle/_ref/).
 * Everything else in this file is "parity unpinned": a line-by-line restatement,
 * each function citing the reference lines it follows
 * (SC/ = /root/reference/pgo/SC-A-LOAM/include/scancontext/).
 *
 * Arithmetic contract (what "Eigen" means here): the reference builds without
 * -march flags, so Eigen 3.3 vectorises doubles with SSE2 packets of 2 and no FMA.
 * Dynamic-size reductions (sum/mean/dot/norm) use Redux.h's
 * LinearVectorizedTraversal + NoUnrolling: two packet accumulators walking the data
 * 4 scalars at a time, combined, then a horizontal add, then a scalar tail.
 * redux_sum() below is that order.  Compile with -ffp-contract=off so gcc does not
 * fuse a*b+c.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SC_RINGS   20   /* SC/Scancontext.h:85 PC_NUM_RING   */
#define SC_SECTORS 60   /* SC/Scancontext.h:86 PC_NUM_SECTOR */

/* Eigen 3.3 Redux.h redux_impl<Func, Derived, LinearVectorizedTraversal, NoUnrolling>
 * for a contiguous, 16-byte aligned run of n doubles (alignedStart == 0), packet = 2. */
static double redux_sum(const double *v, int n)
{
    const int aligned2 = (n / 4) * 4;
    const int aligned  = (n / 2) * 2;
    double res;
    if (aligned) {
        double p0a = v[0], p0b = v[1];
        if (aligned > 2) {
            double p1a = v[2], p1b = v[3];
            for (int i = 4; i < aligned2; i += 4) {
                p0a += v[i];     p0b += v[i + 1];
                p1a += v[i + 2]; p1b += v[i + 3];
            }
            p0a += p1a; p0b += p1b;
            if (aligned > aligned2) { p0a += v[aligned2]; p0b += v[aligned2 + 1]; }
        }
        res = p0a + p0b;
        for (int i = aligned; i < n; ++i) res += v[i];
    } else {
        res = v[0];
        for (int i = 1; i < n; ++i) res += v[i];
    }
    return res;
}

static double redux_dot(const double *a, const double *b, int n)
{
    double t[SC_SECTORS > SC_RINGS ? SC_SECTORS : SC_RINGS];
    for (int i = 0; i < n; ++i) t[i] = a[i] * b[i];
    return redux_sum(t, n);
}

/* SC/Scancontext.cpp:23-36 xy2theta.  float in, float out; `atan` on a float argument
 * resolves to the float overload once <math.h> is in the translation unit (it is, via the
 * ROS/PCL headers), so atanf; the (180/M_PI) factor is a double and the product is rounded
 * to float on return. */
float sc_xy2theta(float x, float y)
{
    if ((x >= 0) & (y >= 0)) return (float)((180 / M_PI) * atanf(y / x));
    if ((x < 0) & (y >= 0))  return (float)(180 - ((180 / M_PI) * atanf(y / (-x))));
    if ((x < 0) & (y < 0))   return (float)(180 + ((180 / M_PI) * atanf(y / x)));
    if ((x >= 0) & (y < 0))  return (float)(360 - ((180 / M_PI) * atanf((-y) / x)));
    return 0.0f; /* NaN input: the reference falls off the end (UB) */
}

/* SC/Scancontext.cpp:151-195 makeScancontext.  xyz = P x 3 float (x,y,z per point).
 * desc = 20 x 60 doubles, column-major (ring fastest) like Eigen::MatrixXd. */
void sc_make(const float *xyz, int P, double lidar_height, double max_radius, double *desc)
{
    const double NO_POINT = -1000;
    for (int i = 0; i < SC_RINGS * SC_SECTORS; ++i) desc[i] = NO_POINT;
    for (int p = 0; p < P; ++p) {
        float px = xyz[3 * p], py = xyz[3 * p + 1];
        float pz = (float)(xyz[3 * p + 2] + lidar_height);          /* :168 pt.z is a float field */
        float azim_range = sqrtf(px * px + py * py);                 /* :171 */
        float azim_angle = sc_xy2theta(px, py);                      /* :172 */
        if (azim_range > max_radius) continue;                       /* :175 */
        int ring = (int)ceil((azim_range / max_radius) * SC_RINGS);  /* :178 double math */
        ring = ring < SC_RINGS ? ring : SC_RINGS; ring = ring > 1 ? ring : 1;
        int sect = (int)ceil((azim_angle / 360.0) * SC_SECTORS);     /* :179 */
        sect = sect < SC_SECTORS ? sect : SC_SECTORS; sect = sect > 1 ? sect : 1;
        double *bin = &desc[(sect - 1) * SC_RINGS + (ring - 1)];
        if (*bin < pz) *bin = pz;                                    /* :182-183 */
    }
    for (int i = 0; i < SC_RINGS * SC_SECTORS; ++i)
        if (desc[i] == NO_POINT) desc[i] = 0;                        /* :187-190 */
}

/* SC/Scancontext.cpp:198-211: row-wise mean (copy of the row, then Eigen mean()). */
void sc_ringkey(const double *desc, double *key20)
{
    double row[SC_SECTORS];
    for (int r = 0; r < SC_RINGS; ++r) {
        for (int c = 0; c < SC_SECTORS; ++c) row[c] = desc[c * SC_RINGS + r];
        key20[r] = redux_sum(row, SC_SECTORS) / (double)SC_SECTORS;
    }
}

/* SC/Scancontext.cpp:214-227: column-wise mean. */
void sc_sectorkey(const double *desc, double *key60)
{
    for (int c = 0; c < SC_SECTORS; ++c)
        key60[c] = redux_sum(desc + c * SC_RINGS, SC_RINGS) / (double)SC_RINGS;
}

/* SC/Scancontext.cpp:62-66 eig2stdvec: double -> float narrowing of the ring key. */
void sc_ringkey_f32(const double *desc, float *key20f)
{
    double k[SC_RINGS];
    sc_ringkey(desc, k);
    for (int r = 0; r < SC_RINGS; ++r) key20f[r] = (float)k[r];
}

/* SC/Scancontext.cpp:93-113 fastAlignUsingVkey (circshift :39-59 folded in:
 * shifted[(c+s)%60] = vkey2[c]). Compares norms AFTER sqrt, strict '<'. */
int sc_fast_align(const double *vkey1, const double *vkey2)
{
    int argmin = 0;
    double min_norm = 10000000;
    double d2[SC_SECTORS];
    for (int s = 0; s < SC_SECTORS; ++s) {
        for (int c = 0; c < SC_SECTORS; ++c) {
            double d = vkey1[c] - vkey2[(c - s + SC_SECTORS) % SC_SECTORS];
            d2[c] = d * d;
        }
        double nrm = sqrt(redux_sum(d2, SC_SECTORS));
        if (nrm < min_norm) { argmin = s; min_norm = nrm; }
    }
    return argmin;
}

/* SC/Scancontext.cpp:69-90 distDirectSC, with sc2 read through a column shift
 * (shifted.col(c) == sc2.col((c - shift) mod 60)). */
static double dist_direct_shifted(const double *sc1, const double *sc2, int shift)
{
    int num_eff_cols = 0;
    double sum_sim = 0;
    for (int c = 0; c < SC_SECTORS; ++c) {
        const double *a = sc1 + c * SC_RINGS;
        const double *b = sc2 + ((c - shift + SC_SECTORS) % SC_SECTORS) * SC_RINGS;
        double na = sqrt(redux_dot(a, a, SC_RINGS));
        double nb = sqrt(redux_dot(b, b, SC_RINGS));
        if ((na == 0) | (nb == 0)) continue;
        double sim = redux_dot(a, b, SC_RINGS) / (na * nb);
        sum_sim = sum_sim + sim;
        num_eff_cols = num_eff_cols + 1;
    }
    double sc_sim = sum_sim / num_eff_cols;   /* 0/0 -> NaN, kept */
    return 1.0 - sc_sim;
}

double sc_dist_direct(const double *sc1, const double *sc2)
{
    return dist_direct_shifted(sc1, sc2, 0);
}

static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

/* SC/Scancontext.cpp:116-148 distanceBtnScanContext. */
void sc_distance(const double *sc1, const double *sc2, double *dist_out, int *shift_out)
{
    double vk1[SC_SECTORS], vk2[SC_SECTORS];
    sc_sectorkey(sc1, vk1);
    sc_sectorkey(sc2, vk2);
    int s0 = sc_fast_align(vk1, vk2);
    const int radius = (int)round(0.5 * 0.1 * SC_SECTORS);   /* SEARCH_RATIO 0.1 -> 3 */
    int space[2 * 8 + 1], n = 0;
    space[n++] = s0;
    for (int ii = 1; ii < radius + 1; ++ii) {
        space[n++] = (s0 + ii + SC_SECTORS) % SC_SECTORS;
        space[n++] = (s0 - ii + SC_SECTORS) % SC_SECTORS;
    }
    qsort(space, n, sizeof(int), cmp_int);                   /* :130 ascending */
    int argmin_shift = 0;
    double min_dist = 10000000;
    for (int k = 0; k < n; ++k) {
        double d = dist_direct_shifted(sc1, sc2, space[k]);
        if (d < min_dist) { argmin_shift = space[k]; min_dist = d; }
    }
    *dist_out = min_dist;
    *shift_out = argmin_shift;
}

/* Exhaustive search over db[first, last): the loop of SC/Scancontext.cpp:380-396 run over
 * every stored keyframe instead of 3 kd-tree candidates (what BASELINE.json's north_star
 * asks for).  Strict '<' keeps the lowest index on ties.  Per-keyframe results optional. */
void sc_search_range(const double *query, const double *db, int first, int last,
                     double *dist_all, int *shift_all,
                     double *best_dist, int *best_idx, int *best_shift)
{
    double min_dist = 10000000; int nn_align = 0, nn_idx = 0;
#pragma omp parallel
    {
        double l_min = 10000000; int l_align = 0, l_idx = 0;
#pragma omp for schedule(static) nowait
        for (int i = first; i < last; ++i) {
            double d; int s;
            sc_distance(query, db + (size_t)i * SC_RINGS * SC_SECTORS, &d, &s);
            if (dist_all) dist_all[i] = d;
            if (shift_all) shift_all[i] = s;
            if (d < l_min) { l_min = d; l_align = s; l_idx = i; }
        }
#pragma omp critical
        {
            if (l_min < min_dist || (l_min == min_dist && l_min < 10000000 && l_idx < nn_idx)) {
                min_dist = l_min; nn_align = l_align; nn_idx = l_idx;
            }
        }
    }
    *best_dist = min_dist; *best_idx = nn_idx; *best_shift = nn_align;
}

/* The candidate loop of detectLoopClosureID (SC/Scancontext.cpp:362-420) given candidate ids
 * (from the kd-tree, see oracle/ref_kdtree_wrap.cpp).  Returns loop id or -1 and the yaw
 * (float, deg2rad(shift * 6)). */
void sc_detect_from_candidates(const double *query, const double *db, const int *cand, int n_cand,
                               double thres, int *loop_id, float *yaw, double *min_dist_out,
                               int *nn_idx_out)
{
    double min_dist = 10000000; int nn_align = 0, nn_idx = 0;
    for (int k = 0; k < n_cand; ++k) {
        double d; int s;
        sc_distance(query, db + (size_t)cand[k] * SC_RINGS * SC_SECTORS, &d, &s);
        if (d < min_dist) { min_dist = d; nn_align = s; nn_idx = cand[k]; }
    }
    *loop_id = (min_dist < thres) ? nn_idx : -1;
    float degrees = (float)(nn_align * (360.0 / (double)SC_SECTORS));   /* deg2rad(float) :17-20 */
    *yaw = (float)(degrees * M_PI / 180.0);
    if (min_dist_out) *min_dist_out = min_dist;
    if (nn_idx_out) *nn_idx_out = nn_idx;
}

/* nanoflann L2_Simple?  No: KDTreeVectorOfVectorsAdaptor uses metric_L2 -> L2_Adaptor
 * (nanoflann.hpp:375-412): 4-way unrolled float accumulation
 *   result += d0*d0 + d1*d1 + d2*d2 + d3*d3  per group of 4, then a scalar tail.
 * Brute-force kNN with that metric; equal distances keep the earlier-visited (lower) index
 * first.  Used to cross-check the real nanoflann in oracle/_ref. */
float sc_key_l2(const float *a, const float *b, int dim)
{
    float result = 0;
    int i = 0;
    for (; i + 3 < dim; i += 4) {
        const float d0 = a[i] - b[i], d1 = a[i + 1] - b[i + 1];
        const float d2 = a[i + 2] - b[i + 2], d3 = a[i + 3] - b[i + 3];
        result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    for (; i < dim; ++i) { const float d0 = a[i] - b[i]; result += d0 * d0; }
    return result;
}

void sc_knn_bruteforce(const float *keys, int n, int dim, const float *query, int k,
                       int *out_idx, float *out_d2)
{
    for (int j = 0; j < k; ++j) { out_idx[j] = -1; out_d2[j] = INFINITY; }
    for (int i = 0; i < n; ++i) {
        float d = sc_key_l2(query, keys + (size_t)i * dim, dim);
        int j = k;
        while (j > 0 && out_d2[j - 1] > d) {
            if (j < k) { out_d2[j] = out_d2[j - 1]; out_idx[j] = out_idx[j - 1]; }
            --j;
        }
        if (j < k) { out_d2[j] = d; out_idx[j] = i; }
    }
}

/* torchrun exports OMP_NUM_THREADS=1; the timed CPU baseline asks for its thread count explicitly */
#ifdef _OPENMP
#include <omp.h>
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_max_threads(void) { return omp_get_max_threads(); }
#else
void oracle_set_threads(int n) { (void)n; }
int oracle_max_threads(void) { return 1; }
#endif
