/*
 * oracle/solver_ref.c -- CPU restatement of the ORORA solver (TEST INFRASTRUCTURE ONLY; see
 * oracle/__init__.py).
 *
 * Follows, line by line, /root/reference/outlier-robust-radar-odometry/
 *   include/ORORA.hpp:37-93,139-151           calcTransform, compensateDopplerDistortion
 *   include/solver/nonminimal_solver.hpp:171-264 (inputs, TIMs, scale consistency)
 *                                        :266-440 (GNC-TLS rotation), :844-859 (svdRot2d)
 *                                        :480-529, :532-661 (anisotropic COTE translation)
 *                                        :678-825 (computeTransformation)
 *   src/graph.cc:12-130 + solver/graph.h (max clique: PMC is NOT vendored -- pinned archive
 *       LimHyungTae/pmc tag libpmc, 3rdparty/pmc/pmc.cmake:28-29 -- so the maximum clique is
 *       found here by an exact bitset branch and bound; any exact solver returns the same set
 *       when the maximum clique is unique, which the test scenes guarantee and the tests check)
 *   src/association.cpp:179-222 SE3tose3 (degenerates to rho = T(0:3,3), :215)
 *
 * PARITY PIN STATUS: unpinned by the reference (no tests / golden vectors exist, Eigen and PMC
 * are absent so the solver cannot be compiled here).  Sums follow plain left-to-right order;
 * Eigen's GEMM blocking makes the reference's own last-ulp behaviour machine dependent
 * (kc is derived from cpuid cache sizes), hence the 1e-4 m / 1e-5 rad tolerance of north_star.
 * std::sort (unstable) is restated as a stable sort.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double noise_bound;        /* params.noise_bound                (orora_params.yaml: 0.75) */
    double cbar2;              /* params.cbar2 = noise_bound_coeff  (1.0)                     */
    int    max_iter;           /* rotation_max_iterations           (100)                     */
    double gnc_factor;         /* rotation_gnc_factor               (1.4)                     */
    double cost_thr;           /* rotation_cost_threshold           (1e-6)                    */
    double nb_radial;          /* noise_bound_radial                (0.1)                     */
    double nb_tangential;      /* noise_bound_tangential [rad]      (9 deg)                   */
} orora_params;

typedef struct {
    int    valid;
    int    n_edges;            /* scale-consistent TIMs = edges of the inlier graph           */
    int    M;                  /* max clique size                                              */
    int    gnc_iters;          /* iterations of the main GNC loop actually run                 */
    int    gnc_early_return;   /* 1 when mu <= 0 at initialisation (:348-352)                  */
    double R[4];               /* 2x2 rotation, column-major                                   */
    double t[2];
    double cost;
    int    n_rot_inliers;      /* rotation_inliers_.size()  (:776-790)                         */
    int    n_final_inliers;
} orora_info;

/* ---- ORORA.hpp:139-151 ------------------------------------------------------------------------ */
void orora_doppler(const double *p2d, int N, double vel_x, double beta, double *out)
{
    double v = fabs(vel_x);
    for (int i = 0; i < N; ++i) {
        double px = p2d[2 * i], py = p2d[2 * i + 1];
        double rsq = px * px + py * py;
        px += beta * v * px * px / rsq;
        py += beta * v * px * py / rsq;      /* uses the already updated x (:146-147) */
        out[2 * i] = px;
        out[2 * i + 1] = py;
    }
}

/* ---- nonminimal_solver.hpp:185-264: is TIM (i,j) scale consistent? ---------------------------- */
static int tim_consistent(const double *src, const double *tgt, int i, int j, double beta)
{
    double sx = src[2 * j] - src[2 * i], sy = src[2 * j + 1] - src[2 * i + 1];
    double tx = tgt[2 * j] - tgt[2 * i], ty = tgt[2 * j + 1] - tgt[2 * i + 1];
    double v1 = sqrt(sx * sx + sy * sy + 0.0);     /* z of a TIM is 1 - 1 = 0 */
    double v2 = sqrt(tx * tx + ty * ty + 0.0);
    double alpha_f = beta * (1.0 / v1);
    double raw_f = v2 / v1;
    int in_f = fabs(raw_f - 1.0) <= alpha_f;
    double alpha_r = beta * (1.0 / v2);
    double raw_r = v1 / v2;
    int in_r = fabs(raw_r - 1.0) <= alpha_r;
    return in_f && in_r;
}

/* adjacency as N x W 64-bit words; returns number of edges */
int orora_build_adjacency(const double *src, const double *tgt, int N, double noise_bound,
                          double cbar2, uint64_t *adj, int W)
{
    double beta = 2 * noise_bound * sqrt(cbar2);
    int edges = 0;
    memset(adj, 0, (size_t)N * W * sizeof(uint64_t));
    for (int i = 0; i < N - 1; ++i)
        for (int j = i + 1; j < N; ++j)
            if (tim_consistent(src, tgt, i, j, beta)) {
                adj[(size_t)i * W + (j >> 6)] |= 1ull << (j & 63);
                adj[(size_t)j * W + (i >> 6)] |= 1ull << (i & 63);
                ++edges;
            }
    return edges;
}

/* ---- exact maximum clique: bitset branch and bound with greedy colouring (BBMC-style) --------- */
typedef struct {
    int N, W;
    const uint64_t *adj;   /* in search order */
    int *best, best_n;
    int *cur, cur_n;
    uint64_t *pool;        /* per-depth P / scratch */
    int *ord_pool, *col_pool;
    long long nodes;
} mc_state;

static int bs_count(const uint64_t *a, int W)
{
    int c = 0;
    for (int w = 0; w < W; ++w) c += __builtin_popcountll(a[w]);
    return c;
}

static void mc_expand(mc_state *S, int depth, uint64_t *P)
{
    const int W = S->W;
    S->nodes++;
    int *order = S->ord_pool + (size_t)depth * S->N;
    int *color = S->col_pool + (size_t)depth * S->N;
    uint64_t *Q = S->pool + (size_t)(3 * depth) * W;
    uint64_t *Qc = Q + W;
    uint64_t *Pn = Qc + W;
    int n = 0, k = 0;
    memcpy(Q, P, W * sizeof(uint64_t));
    for (;;) {
        int any = 0;
        for (int w = 0; w < W; ++w) any |= (Q[w] != 0);
        if (!any) break;
        ++k;
        memcpy(Qc, Q, W * sizeof(uint64_t));
        for (;;) {
            int v = -1;
            for (int w = 0; w < W; ++w)
                if (Qc[w]) { v = w * 64 + __builtin_ctzll(Qc[w]); break; }
            if (v < 0) break;
            Qc[v >> 6] &= ~(1ull << (v & 63));
            Q[v >> 6] &= ~(1ull << (v & 63));
            const uint64_t *av = S->adj + (size_t)v * W;
            for (int w = 0; w < W; ++w) Qc[w] &= ~av[w];
            order[n] = v; color[n] = k; ++n;
        }
    }
    if (k == n && n > 0) {
        /* every vertex got its own colour: P is a clique, the best leaf of this subtree is cur + P */
        if (S->cur_n + n > S->best_n) {
            memcpy(S->best, S->cur, S->cur_n * sizeof(int));
            for (int i = 0; i < n; ++i) S->best[S->cur_n + i] = order[i];
            S->best_n = S->cur_n + n;
        }
        return;
    }
    for (int idx = n - 1; idx >= 0; --idx) {
        if (S->cur_n + color[idx] <= S->best_n) return;
        int v = order[idx];
        S->cur[S->cur_n++] = v;
        const uint64_t *av = S->adj + (size_t)v * W;
        int any = 0;
        for (int w = 0; w < W; ++w) { Pn[w] = P[w] & av[w]; any |= (Pn[w] != 0); }
        if (!any) {
            if (S->cur_n > S->best_n) { S->best_n = S->cur_n; memcpy(S->best, S->cur, S->cur_n * sizeof(int)); }
        } else {
            mc_expand(S, depth + 1, Pn);
        }
        S->cur_n--;
        P[v >> 6] &= ~(1ull << (v & 63));
    }
}


/* ---- maximum independent set on the complement of the pruned core ---------------------------
 * Inlier graphs are near-cliques: inside the best_n-core almost every pair is adjacent, so the
 * complement H is sparse and a vertex-cover style search closes in a handful of nodes where the
 * colouring branch and bound needs thousands.  Deterministic, order-free ("Jacobi") reductions so
 * that the CUDA kernel (one thread per vertex) follows exactly the same tree:
 *   round: degrees d(v) = |H(v) & P| from ONE snapshot of P;
 *          d = 0           -> v joins the set;
 *          d = 1, nbr u    -> if d(u) = 1 (isolated edge) the lower index joins and the other leaves,
 *                             else v joins and u leaves;
 *   repeat until nothing changes; bound size + |P| - ceil(m / maxdeg) (m = edges of H[P]);
 *   branch on the vertex of maximum degree (lowest index on ties): include it first, then exclude. */
typedef struct {
    int N, W;
    const uint64_t *adj;      /* search-order adjacency */
    const uint64_t *core;     /* W words */
    int best_n;
    uint64_t *bestC;          /* W words; valid when improved */
    int improved;
    long long nodes, limit;
    int exact;
} mis_state;

static inline int bit_get(const uint64_t *b, int v) { return (int)((b[v >> 6] >> (v & 63)) & 1ull); }

static void mis_rec(mis_state *S, const uint64_t *Pin, const uint64_t *Cin, int size)
{
    const int N = S->N, W = S->W;
    S->nodes++;
    uint64_t *P = (uint64_t *)malloc(4 * (size_t)W * sizeof(uint64_t));
    uint64_t *C = P + W, *take = C + W, *rem = take + W;
    int *d = (int *)malloc((size_t)N * sizeof(int));
    memcpy(P, Pin, W * sizeof(uint64_t));
    memcpy(C, Cin, W * sizeof(uint64_t));
    for (;;) {
        for (int v = 0; v < N; ++v) {
            if (!bit_get(P, v)) { d[v] = -1; continue; }
            int c = 0;
            for (int w = 0; w < W; ++w) c += __builtin_popcountll(~S->adj[(size_t)v * W + w] & S->core[w] & P[w]);
            d[v] = c - 1;                                   /* the complement row includes v itself */
        }
        memset(take, 0, 2 * (size_t)W * sizeof(uint64_t));
        int changed = 0;
        for (int v = 0; v < N; ++v) {
            if (d[v] == 0) { take[v >> 6] |= 1ull << (v & 63); changed = 1; }
            else if (d[v] == 1) {
                int u = -1;
                for (int w = 0; w < W && u < 0; ++w) {
                    uint64_t x = ~S->adj[(size_t)v * W + w] & S->core[w] & P[w];
                    if ((v >> 6) == w) x &= ~(1ull << (v & 63));
                    if (x) u = w * 64 + __builtin_ctzll(x);
                }
                if (d[u] == 1) { if (v < u) { take[v >> 6] |= 1ull << (v & 63); rem[u >> 6] |= 1ull << (u & 63); changed = 1; } }
                else { take[v >> 6] |= 1ull << (v & 63); rem[u >> 6] |= 1ull << (u & 63); changed = 1; }
            }
        }
        if (!changed) break;
        for (int w = 0; w < W; ++w) {
            size += __builtin_popcountll(take[w]);
            C[w] |= take[w];
            P[w] &= ~(take[w] | rem[w]);
        }
    }
    int np = 0, m2 = 0, md = 0, vstar = -1;
    for (int v = 0; v < N; ++v)
        if (d[v] >= 0) { ++np; m2 += d[v]; if (d[v] > md) { md = d[v]; vstar = v; } }
    if (np == 0) {
        if (size > S->best_n) { S->best_n = size; memcpy(S->bestC, C, W * sizeof(uint64_t)); S->improved = 1; }
    } else {
        const int m = m2 / 2;
        const int ub = size + np - (m + md - 1) / md;
        if (ub > S->best_n) {
            if (S->nodes >= S->limit) S->exact = 0;
            else {
                uint64_t *Q = (uint64_t *)malloc(2 * (size_t)W * sizeof(uint64_t)), *QC = Q + W;
                for (int w = 0; w < W; ++w) { Q[w] = P[w] & S->adj[(size_t)vstar * W + w]; QC[w] = C[w]; }   /* P \ H[v] \ {v} */
                QC[vstar >> 6] |= 1ull << (vstar & 63);
                mis_rec(S, Q, QC, size + 1);
                for (int w = 0; w < W; ++w) Q[w] = P[w];
                Q[vstar >> 6] &= ~(1ull << (vstar & 63));
                mis_rec(S, Q, C, size);
                free(Q);
            }
        }
    }
    free(d); free(P);
}

/* Returns clique size; clique[] holds ORIGINAL vertex ids sorted ascending
 * (nonminimal_solver.hpp:712 sorts the solver's output). */
int orora_max_clique(const uint64_t *adj_in, int N, int W, int *clique, long long *nodes_out)
{
    if (N == 0) return 0;
    int *deg = (int *)malloc(N * sizeof(int)), *perm = (int *)malloc(N * sizeof(int)),
        *inv = (int *)malloc(N * sizeof(int));
    for (int i = 0; i < N; ++i) { deg[i] = bs_count(adj_in + (size_t)i * W, W); perm[i] = i; }
    /* search order: non-increasing degree, ties by index (stable insertion sort) */
    for (int i = 1; i < N; ++i) {
        int p = perm[i], j = i - 1;
        while (j >= 0 && deg[perm[j]] < deg[p]) { perm[j + 1] = perm[j]; --j; }
        perm[j + 1] = p;
    }
    for (int i = 0; i < N; ++i) inv[perm[i]] = i;
    uint64_t *adj = (uint64_t *)calloc((size_t)N * W, sizeof(uint64_t));
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j)
            if (adj_in[(size_t)i * W + (j >> 6)] >> (j & 63) & 1ull)
                adj[(size_t)inv[i] * W + (inv[j] >> 6)] |= 1ull << (inv[j] & 63);
    mc_state S;
    S.N = N; S.W = W; S.adj = adj; S.nodes = 0;
    S.best = (int *)malloc(N * sizeof(int)); S.best_n = 0;
    S.cur = (int *)malloc(N * sizeof(int)); S.cur_n = 0;
    S.pool = (uint64_t *)malloc((size_t)3 * (N + 1) * W * sizeof(uint64_t));
    S.ord_pool = (int *)malloc((size_t)(N + 1) * N * sizeof(int));
    S.col_pool = (int *)malloc((size_t)(N + 1) * N * sizeof(int));
    /* greedy lower bound: repeatedly take the vertex with the most neighbours inside the current
     * candidate set (ties: lowest index in search order), shrink the set to its neighbourhood */
    {
        uint64_t *P = (uint64_t *)malloc(W * sizeof(uint64_t));
        for (int w = 0; w < W; ++w) P[w] = 0;
        for (int i = 0; i < N; ++i) P[i >> 6] |= 1ull << (i & 63);
        int n = 0;
        for (;;) {
            int v = -1, bestc = -1;
            for (int u = 0; u < N; ++u) {
                if (!(P[u >> 6] >> (u & 63) & 1ull)) continue;
                int c = 0;
                for (int w = 0; w < W; ++w) c += __builtin_popcountll(adj[(size_t)u * W + w] & P[w]);
                if (c > bestc) { bestc = c; v = u; }
            }
            if (v < 0) break;
            S.best[n++] = v;
            for (int w = 0; w < W; ++w) P[w] &= adj[(size_t)v * W + w];
        }
        S.best_n = n;
        for (int w = 0; w < W; ++w) P[w] = 0;
        for (int i = 0; i < N; ++i) P[i >> 6] |= 1ull << (i & 63);
        /* a clique larger than the greedy one only contains vertices with >= best_n neighbours inside
         * the surviving set: peel everything else off (the best_n-core; unique whatever the order) */
        for (int changed = 1; changed;) {
            changed = 0;
            for (int v = 0; v < N; ++v) {
                if (!(P[v >> 6] >> (v & 63) & 1ull)) continue;
                int c = 0;
                for (int w = 0; w < W; ++w) c += __builtin_popcountll(adj[(size_t)v * W + w] & P[w]);
                if (c < S.best_n) { P[v >> 6] &= ~(1ull << (v & 63)); changed = 1; }
            }
        }
        int any = 0, n_core = 0;
        for (int w = 0; w < W; ++w) { any |= (P[w] != 0); n_core += __builtin_popcountll(P[w]); }
        if (any) {
            /* complement edges inside the core decide the exact method (both are exact) */
            long long mh2 = 0;
            for (int v = 0; v < N; ++v)
                if (P[v >> 6] >> (v & 63) & 1ull) {
                    int c = 0;
                    for (int w = 0; w < W; ++w) c += __builtin_popcountll(~adj[(size_t)v * W + w] & P[w]);
                    mh2 += c - 1;
                }
            if (mh2 / 2 <= 3LL * n_core) {
                mis_state M;
                M.N = N; M.W = W; M.adj = adj; M.core = P; M.best_n = S.best_n; M.improved = 0;
                M.bestC = (uint64_t *)calloc(W, sizeof(uint64_t));
                M.nodes = 0; M.limit = 2000000; M.exact = 1;
                uint64_t *C0 = (uint64_t *)calloc(W, sizeof(uint64_t));
                uint64_t *P0 = (uint64_t *)malloc(W * sizeof(uint64_t));
                memcpy(P0, P, W * sizeof(uint64_t));
                mis_rec(&M, P0, C0, 0);
                if (M.improved) {
                    int k = 0;
                    for (int v = 0; v < N; ++v) if (M.bestC[v >> 6] >> (v & 63) & 1ull) S.best[k++] = v;
                    S.best_n = k;
                }
                S.nodes = M.nodes;
                free(M.bestC); free(C0); free(P0);
            } else {
                mc_expand(&S, 0, P);
            }
        }
        free(P);
    }
    int n = S.best_n;
    for (int i = 0; i < n; ++i) clique[i] = perm[S.best[i]];
    for (int i = 1; i < n; ++i) { int p = clique[i], j = i - 1; while (j >= 0 && clique[j] > p) { clique[j + 1] = clique[j]; --j; } clique[j + 1] = p; }
    if (nodes_out) *nodes_out = S.nodes;
    free(deg); free(perm); free(inv); free(adj); free(S.best); free(S.cur); free(S.pool);
    free(S.ord_pool); free(S.col_pool);
    return n;
}

/* ---- nonminimal_solver.hpp:844-859 svdRot2d ---------------------------------------------------- */
/* 2x2 SVD by the closed-form two-sided Jacobi construction; returns R = V * U^T with the
 * determinant fix of :854-856.  Column-major 2x2. */
static void svd2x2(const double H[4], double U[4], double sv[2], double V[4])
{
    /* H = [a b; c d] with column-major storage H[0]=a H[1]=c H[2]=b H[3]=d */
    double a = H[0], c = H[1], b = H[2], d = H[3];
    /* symmetric eigen-decomposition of H^T H */
    double e = a * a + c * c, f = a * b + c * d, g = b * b + d * d;
    double theta = 0.5 * atan2(2 * f, e - g);
    double ct = cos(theta), st = sin(theta);
    /* V columns: (ct, st), (-st, ct) */
    V[0] = ct; V[1] = st; V[2] = -st; V[3] = ct;
    /* H V */
    double hv00 = a * ct + b * st, hv10 = c * ct + d * st;
    double hv01 = -a * st + b * ct, hv11 = -c * st + d * ct;
    double s0 = sqrt(hv00 * hv00 + hv10 * hv10), s1 = sqrt(hv01 * hv01 + hv11 * hv11);
    if (s0 < s1) {   /* sort singular values descending: swap columns */
        double t;
        t = s0; s0 = s1; s1 = t;
        t = hv00; hv00 = hv01; hv01 = t;
        t = hv10; hv10 = hv11; hv11 = t;
        t = V[0]; V[0] = V[2]; V[2] = t;
        t = V[1]; V[1] = V[3]; V[3] = t;
    }
    sv[0] = s0; sv[1] = s1;
    if (s0 > 0) { U[0] = hv00 / s0; U[1] = hv10 / s0; } else { U[0] = 1; U[1] = 0; }
    /* second left singular vector: +-perp(u1), signed so that H v2 = s1 u2 with s1 >= 0; this stays
     * orthonormal when H is rank deficient (JacobiSVD always returns orthogonal factors) */
    U[2] = -U[1]; U[3] = U[0];
    if (U[2] * hv01 + U[3] * hv11 < 0) { U[2] = -U[2]; U[3] = -U[3]; }
}

void orora_svd_rot2d(const double *X, const double *Y, const double *Wt, int M, double R[4])
{
    double H[4] = {0, 0, 0, 0};    /* H = X * diag(W) * Y^T */
    for (int j = 0; j < M; ++j) {
        double wx0 = X[2 * j] * Wt[j], wx1 = X[2 * j + 1] * Wt[j];
        H[0] += wx0 * Y[2 * j];     H[2] += wx0 * Y[2 * j + 1];
        H[1] += wx1 * Y[2 * j];     H[3] += wx1 * Y[2 * j + 1];
    }
    double U[4], V[4], sv[2];
    svd2x2(H, U, sv, V);
    double detU = U[0] * U[3] - U[2] * U[1], detV = V[0] * V[3] - V[2] * V[1];
    if (detU * detV < 0) { V[2] = -V[2]; V[3] = -V[3]; }
    /* R = V * U^T */
    R[0] = V[0] * U[0] + V[2] * U[2];  R[2] = V[0] * U[1] + V[2] * U[3];
    R[1] = V[1] * U[0] + V[3] * U[2];  R[3] = V[1] * U[1] + V[3] * U[3];
}

/* ---- nonminimal_solver.hpp:292-440 solveForRotation2D (GNC-TLS) -------------------------------- */
void orora_gnc_rotation(const double *src, const double *dst, const double *init_w, int M,
                        const orora_params *p, double R[4], double *weights, unsigned char *inl,
                        int *iters_out, int *early_out, double *cost_out)
{
    double mu = 1, prev_cost = INFINITY, cost = INFINITY;
    double nb2 = pow(p->noise_bound, 2);
    if (nb2 < 1e-16) nb2 = 1e-2;
    double *res = (double *)malloc(M * sizeof(double));
    for (int j = 0; j < M; ++j) weights[j] = 1.0;
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 1;
    *iters_out = 0; *early_out = 0;
    orora_svd_rot2d(src, dst, init_w, M, R);                         /* :333 */
    double max_res = -INFINITY;
    for (int j = 0; j < M; ++j) {
        double dx = dst[2 * j] - (R[0] * src[2 * j] + R[2] * src[2 * j + 1]);
        double dy = dst[2 * j + 1] - (R[1] * src[2 * j] + R[3] * src[2 * j + 1]);
        res[j] = dx * dx + dy * dy;
        if (res[j] > max_res) max_res = res[j];
    }
    const double EPSILON = 0.00000000000001;
    if (max_res < nb2) nb2 = 0.01;                                   /* :342-344 */
    mu = 1 / (2 * max_res / nb2 - 1);
    if (mu <= 0) {                                                   /* :348-352 */
        *early_out = 1;
        goto done;
    }
    {
        double th1 = (mu + 1) / mu * nb2, th2 = mu / (mu + 1) * nb2;
        cost = 0;
        for (int j = 0; j < M; ++j) {
            cost += weights[j] * res[j];
            if (res[j] >= th1) weights[j] = 0;
            else if (res[j] <= th2) weights[j] = 1;
            else weights[j] = sqrt(nb2 * mu * (mu + 1) / res[j] + EPSILON) - mu;   /* :368 */
        }
        mu = mu * p->gnc_factor;
    }
    for (int i = 0; i < p->max_iter; ++i) {                          /* :376-433 */
        orora_svd_rot2d(src, dst, weights, M, R);
        for (int j = 0; j < M; ++j) {
            double dx = dst[2 * j] - (R[0] * src[2 * j] + R[2] * src[2 * j + 1]);
            double dy = dst[2 * j + 1] - (R[1] * src[2 * j] + R[3] * src[2 * j + 1]);
            res[j] = dx * dx + dy * dy;
        }
        double th1 = (mu + 1) / mu * nb2, th2 = mu / (mu + 1) * nb2;
        cost = 0;
        for (int j = 0; j < M; ++j) {
            cost += weights[j] * res[j];
            if (res[j] >= th1) weights[j] = 0;
            else if (res[j] <= th2) weights[j] = 1;
            else weights[j] = sqrt(nb2 * mu * (mu + 1) / res[j]) - mu;              /* :406 */
        }
        double cost_diff = fabs(cost - prev_cost);
        mu = mu * p->gnc_factor;
        prev_cost = cost;
        *iters_out = i + 1;
        if (cost_diff < p->cost_thr) break;
    }
done:
    for (int j = 0; j < M; ++j) inl[j] = weights[j] >= 0.4;          /* :435-439 (also on early return? no) */
    *cost_out = cost;
    free(res);
}

/* ---- nonminimal_solver.hpp:532-661 estimate (COTE on one axis) -------------------------------- */
typedef struct { double v; int tag; int ord; } endpoint;
static int cmp_ep(const void *a, const void *b)
{
    const endpoint *x = (const endpoint *)a, *y = (const endpoint *)b;
    if (x->v < y->v) return -1;
    if (x->v > y->v) return 1;
    return x->ord - y->ord;       /* stable */
}

void orora_cote_axis(const double *X, const double *ranges, int N, double *estimate,
                     unsigned char *inliers)
{
    endpoint *h = (endpoint *)malloc(2 * (size_t)N * sizeof(endpoint));
    for (int i = 0; i < N; ++i) {
        h[2 * i].v = X[i] - ranges[i];     h[2 * i].tag = i + 1;      h[2 * i].ord = 2 * i;
        h[2 * i + 1].v = X[i] + ranges[i]; h[2 * i + 1].tag = -i - 1; h[2 * i + 1].ord = 2 * i + 1;
    }
    qsort(h, 2 * (size_t)N, sizeof(endpoint), cmp_ep);
    double ranges_inverse_sum = 0;
    for (int i = 0; i < N; ++i) ranges_inverse_sum += ranges[i];      /* ranges.sum() */
    double dot_X_weights = 0, dot_weights_consensus = 0, sum_xi = 0, sum_xi_square = 0;
    int consensus = 0;
    double best_cost = 0, best_hat = 0;
    int have = 0;
    for (int i = 0; i < 2 * N; ++i) {
        int idx = abs(h[i].tag) - 1;
        int eps = (h[i].tag > 0) ? 1 : -1;
        double w = 1.0 / (ranges[idx] * ranges[idx]);                  /* :559-560 */
        consensus += eps;
        dot_weights_consensus += eps * w;
        dot_X_weights += eps * w * X[idx];
        ranges_inverse_sum -= eps * ranges[idx];
        sum_xi += eps * X[idx];
        sum_xi_square += eps * X[idx] * X[idx];
        double x_hat = dot_X_weights / dot_weights_consensus;
        double residual = consensus * x_hat * x_hat + sum_xi_square - 2 * sum_xi * x_hat;
        double x_cost = residual + ranges_inverse_sum;
        if (!have || x_cost < best_cost) { best_cost = x_cost; best_hat = x_hat; have = 1; }   /* minCoeff: first */
    }
    *estimate = best_hat;
    for (int i = 0; i < N; ++i) inliers[i] = fabs(X[i] - best_hat) <= ranges[i];
    free(h);
}

/* ---- nonminimal_solver.hpp:480-529 solveForTranslation ------------------------------------------ */
void orora_translation(const double *src_rot, const double *dst, int N, double nb_radial,
                       double nb_tangential, double t[2], unsigned char *inliers)
{
    double var_r = nb_radial * nb_radial, var_t = nb_tangential * nb_tangential;
    double *a0 = (double *)malloc(N * sizeof(double)), *a1 = (double *)malloc(N * sizeof(double));
    double *x0 = (double *)malloc(N * sizeof(double)), *x1 = (double *)malloc(N * sizeof(double));
    unsigned char *tmp = (unsigned char *)malloc(N);
    for (int j = 0; j < N; ++j) {
        double sx = src_rot[2 * j], sy = src_rot[2 * j + 1], dx = dst[2 * j], dy = dst[2 * j + 1];
        double ang_s = atan2(sy, sx), ang_d = atan2(dy, dx);
        double c_s = cos(ang_s), s_s = sin(ang_s), c_d = cos(ang_d), s_d = sin(ang_d);
        double r2s = sx * sx + sy * sy + 1.0;      /* squaredNorm of (x, y, 1): :513-514 */
        double r2d = dx * dx + dy * dy + 1.0;
        double cs11 = c_s * c_s * var_r + s_s * s_s * r2s * var_t;
        double cs22 = s_s * s_s * var_r + c_s * c_s * r2s * var_t;
        double cd11 = c_d * c_d * var_r + s_d * s_d * r2d * var_t;
        double cd22 = s_d * s_d * var_r + c_d * c_d * r2d * var_t;
        a0[j] = sqrt(cs11 + cd11);
        a1[j] = sqrt(cs22 + cd22);
        x0[j] = dx - sx;
        x1[j] = dy - sy;
    }
    for (int j = 0; j < N; ++j) inliers[j] = 1;
    orora_cote_axis(x0, a0, N, &t[0], tmp);
    for (int j = 0; j < N; ++j) inliers[j] = inliers[j] && tmp[j];
    orora_cote_axis(x1, a1, N, &t[1], tmp);
    for (int j = 0; j < N; ++j) inliers[j] = inliers[j] && tmp[j];
    free(a0); free(a1); free(x0); free(x1); free(tmp);
}

/* ---- nonminimal_solver.hpp:678-825 computeTransformation ---------------------------------------
 * src = curr frame matches (setInputSource), tgt = prev frame matches (setInputTarget), 2xN
 * column-major.  T_out = 4x4 column-major.  clique_out has room for N ints; final_inliers N. */
int orora_solve(const double *src, const double *tgt, int N, const orora_params *p, double *T_out,
                int *clique_out, int *final_inliers_out, orora_info *info)
{
    for (int i = 0; i < 16; ++i) T_out[i] = (i % 5 == 0) ? 1.0 : 0.0;
    memset(info, 0, sizeof(*info));
    info->R[0] = 1; info->R[3] = 1;
    if (N < 2) { info->valid = 0; return 0; }
    int W = (N + 63) / 64;
    uint64_t *adj = (uint64_t *)malloc((size_t)N * W * sizeof(uint64_t));
    info->n_edges = orora_build_adjacency(src, tgt, N, p->noise_bound, p->cbar2, adj, W);
    int M = orora_max_clique(adj, N, W, clique_out, NULL);
    free(adj);
    info->M = M;
    if (M <= 1) { info->valid = 0; return 0; }                         /* :715-720 */

    double *ps = (double *)malloc(2 * (size_t)M * sizeof(double)), *pd = (double *)malloc(2 * (size_t)M * sizeof(double));
    double *w0 = (double *)malloc(M * sizeof(double)), *wt = (double *)malloc(M * sizeof(double));
    unsigned char *rot_inl = (unsigned char *)malloc(M), *tr_inl = (unsigned char *)malloc(M);
    double max_w = 0.0;
    for (int i = 0; i < M; ++i) {                                      /* :735-755 chain TIMs */
        int root = clique_out[i], leaf = (i != M - 1) ? clique_out[i + 1] : clique_out[0];
        ps[2 * i] = src[2 * leaf] - src[2 * root];  ps[2 * i + 1] = src[2 * leaf + 1] - src[2 * root + 1];
        pd[2 * i] = tgt[2 * leaf] - tgt[2 * root];  pd[2 * i + 1] = tgt[2 * leaf + 1] - tgt[2 * root + 1];
        double r_sq_l = tgt[2 * leaf] * tgt[2 * leaf] + tgt[2 * leaf + 1] * tgt[2 * leaf + 1] + 1.0;   /* (x,y,1) */
        double r_sq_r = tgt[2 * root] * tgt[2 * root] + tgt[2 * root + 1] * tgt[2 * root + 1] + 1.0;
        w0[i] = 1.0 / sqrt(r_sq_l + r_sq_r);
        if (w0[i] > max_w) max_w = w0[i];
    }
    for (int i = 0; i < M; ++i) w0[i] /= max_w;                        /* :758-760 */
    for (int i = 0; i < 2 * M; ++i) pd[i] *= (1 / 1.0);                /* :762 scale == 1 */

    orora_gnc_rotation(ps, pd, w0, M, p, info->R, wt, rot_inl, &info->gnc_iters,
                       &info->gnc_early_return, &info->cost);
    if (info->gnc_early_return) {
        /* :351 returns before the inlier mask is written: rotation_inliers_mask_ keeps its
         * resize()d (uninitialised) content in the reference; count is meaningless there */
        for (int i = 0; i < M; ++i) rot_inl[i] = 0;
    }
    int nrot = 0;
    for (int i = 0; i < M; ++i) {                                      /* :776-789 */
        int prev = (i == 0) ? M - 1 : i - 1;
        if (rot_inl[prev] && rot_inl[i]) ++nrot;
    }
    info->n_rot_inliers = nrot;

    /* :798-804 rotate the clique's source points, then COTE */
    for (int i = 0; i < M; ++i) {
        int c = clique_out[i];
        double x = src[2 * c], y = src[2 * c + 1];
        ps[2 * i] = info->R[0] * x + info->R[2] * y;
        ps[2 * i + 1] = info->R[1] * x + info->R[3] * y;
        pd[2 * i] = tgt[2 * c];
        pd[2 * i + 1] = tgt[2 * c + 1];
    }
    orora_translation(ps, pd, M, p->nb_radial, p->nb_tangential, info->t, tr_inl);
    int nf = 0;
    for (int i = 0; i < M; ++i) if (tr_inl[i]) final_inliers_out[nf++] = clique_out[i];
    info->n_final_inliers = nf;
    info->valid = 1;
    T_out[0] = info->R[0]; T_out[1] = info->R[1]; T_out[4] = info->R[2]; T_out[5] = info->R[3];
    T_out[12] = info->t[0]; T_out[13] = info->t[1]; T_out[14] = 0.0;  /* :821 */
    free(ps); free(pd); free(w0); free(wt); free(rot_inl); free(tr_inl);
    return 1;
}

/* ---- ORORA.hpp:37-93 calcTransform (state: previous velocity) ---------------------------------
 * prev/curr = matched points 2xN (prev_matched.feat_in_cart_, curr_matched.feat_in_cart_).
 * vel_x_inout = curr_vel_se3_(0) carried between calls (SE3tose3(T)(0)/dt = T(0,3)/dt). */
int orora_calc_transform(const double *prev, const double *curr, int N, const orora_params *p,
                         double delta_t, double beta, int use_doppler, int check_stop_motion,
                         int stop_thr, double *vel_x_inout, double *T_out, int *clique_out,
                         int *final_inliers_out, orora_info *info)
{
    double prev_vel = *vel_x_inout;
    double *pc = (double *)malloc(2 * (size_t)(N > 0 ? N : 1) * sizeof(double));
    double *cc = (double *)malloc(2 * (size_t)(N > 0 ? N : 1) * sizeof(double));
    if (use_doppler) {
        orora_doppler(prev, N, prev_vel, beta, pc);
        orora_doppler(curr, N, prev_vel, beta, cc);
    } else {
        memcpy(pc, prev, 2 * (size_t)N * sizeof(double));
        memcpy(cc, curr, 2 * (size_t)N * sizeof(double));
    }
    int ok = 0;
    if (check_stop_motion && N > stop_thr) {                           /* :74-79 */
        for (int i = 0; i < 16; ++i) T_out[i] = (i % 5 == 0) ? 1.0 : 0.0;
        memset(info, 0, sizeof(*info));
        info->R[0] = 1; info->R[3] = 1;
    } else {
        ok = orora_solve(cc, pc, N, p, T_out, clique_out, final_inliers_out, info);   /* src = curr, tgt = prev */
    }
    *vel_x_inout = T_out[12] / delta_t;                                /* :89 + association.cpp:215 */
    free(pc); free(cc);
    return ok;
}
