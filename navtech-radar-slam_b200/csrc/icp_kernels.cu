// icp_kernels.cu -- the loop-verification step behind the ScanContext detector on sm_100a:
// point-to-point ICP of the current keyframe cloud against the +-25-keyframe submap around the loop
// candidate, and the fitness gate (reference: pgo/SC-A-LOAM/src/laserPosegraphOptimization.cpp:355-406,
// doICPVirtualRelative -> pcl::IterativeClosestPoint with setMaxCorrespondenceDistance(150),
// setMaximumIterations(100), setTransformationEpsilon(1e-6), setEuclideanFitnessEpsilon(1e-6),
// setRANSACIterations(0); accept iff hasConverged() && getFitnessScore() <= 0.3).
//
// PCL is a third-party dependency that is not vendored in the reference (SC/CMakeLists.txt find_package(PCL));
// what is restated here is the published algorithm of pcl::IterativeClosestPoint (registration/impl/icp.hpp,
// default_convergence_criteria.hpp, transformation_estimation_svd.hpp of PCL 1.8-1.10):
//   repeat  correspondences: nearest target point of every (transformed) source point, kept when the squared
//                            distance <= max_correspondence_distance^2            (CorrespondenceEstimation)
//           fewer than 3 correspondences -> not converged, stop
//           rigid transform of the pairs: centroids, 3x3 cross covariance, SVD, R = V diag(1,1,det) U^T,
//                            t = c_tgt - R c_src                                     (TransformationEstimationSVD)
//           source <- T source; final <- T final; ++iterations
//           converged when iterations >= max, or T is within (translation^2 <= transformation_epsilon and
//                            cos(angle) >= 1 - transformation_epsilon), or |mse - mse_prev| / mse_prev <
//                            euclidean_fitness_epsilon, or |mse - mse_prev| < 1e-12    (DefaultConvergenceCriteria;
//                            mse = mean squared correspondence distance BEFORE this iteration's transform)
//   fitness = mean squared nearest-neighbour distance of the aligned source (getFitnessScore, no range limit)
// Arithmetic: points and distances in float like PCL's point types and kd-tree metric, sums in double in a fixed
// order (PCL's float Umeyama is less accurate; parity with PCL itself cannot be pinned here -- see DESIGN.md).
// Nearest neighbours by exhaustive search (tiles of the target in shared memory, strict '<' while scanning up:
// the lowest index among equidistant targets); an exact kd-tree search returns the same pairs except such ties.
//
// One kernel launch per ICP iteration: every CTA applies the pending transform to its source points, finds
// their nearest targets and reduces its 17 partial sums; the last CTA to finish (ticket) adds the partials in
// block order, solves the 3x3 problem, updates the transforms and decides convergence.  Launches after
// convergence return at once (a device flag), so the host enqueues max_iterations launches without reading back.
// Bound: the target tile loop, N_src x N_tgt distance evaluations per iteration out of shared memory (compute /
// LDS-bound; the clouds are a few hundred KB and stay in L2).
#include <float.h>
#include <math.h>
#include <string.h>
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int ICP_THREADS = 128;
constexpr int ICP_TILE = 1024;              // target points staged per tile (12 KB)
constexpr int NSUM = 17;                    // n, sum s (3), sum t (3), sum s t^T (9), sum d2

struct IcpState {
    double T_inc[16];                       // the transform the NEXT launch applies to the source (row-major 4x4)
    double T_final[16];                     // final_transformation_ (row-major)
    double prev_mse;
    double fitness;
    unsigned int ticket;
    int done, converged, state, iterations, n_corr, similar;
    int pad_;
};

__device__ __forceinline__ float3 xform(const double *T, float3 p)   // pcl::transformPointCloud: float arithmetic
{
    const float r00 = (float)T[0], r01 = (float)T[1], r02 = (float)T[2], tx = (float)T[3];
    const float r10 = (float)T[4], r11 = (float)T[5], r12 = (float)T[6], ty = (float)T[7];
    const float r20 = (float)T[8], r21 = (float)T[9], r22 = (float)T[10], tz = (float)T[11];
    float3 o;
    o.x = r00 * p.x + r01 * p.y + r02 * p.z + tx;
    o.y = r10 * p.x + r11 * p.y + r12 * p.z + ty;
    o.z = r20 * p.x + r21 * p.y + r22 * p.z + tz;
    return o;
}

// one-sided Jacobi SVD of a 3x3 matrix (double): A = U diag(s) V^T, U and V orthonormal (columns of U for
// vanishing singular values are completed by cross products)
__device__ void svd3(const double A[9], double U[9], double V[9], int *jmin)
{
    double B[9];
    for (int i = 0; i < 9; ++i) { B[i] = A[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double al = 0, be = 0, ga = 0;
                for (int k = 0; k < 3; ++k) { al += B[3 * k + p] * B[3 * k + p]; be += B[3 * k + q] * B[3 * k + q]; ga += B[3 * k + p] * B[3 * k + q]; }
                off = fmax(off, fabs(ga) / (sqrt(al * be) + 1e-300));
                if (fabs(ga) <= 1e-300) continue;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < 3; ++k) {
                    const double bp = B[3 * k + p], bq = B[3 * k + q];
                    B[3 * k + p] = c * bp - s * bq; B[3 * k + q] = s * bp + c * bq;
                    const double vp = V[3 * k + p], vq = V[3 * k + q];
                    V[3 * k + p] = c * vp - s * vq; V[3 * k + q] = s * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    double nrm[3];
    for (int j = 0; j < 3; ++j) nrm[j] = sqrt(B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j]);
    const double big = fmax(nrm[0], fmax(nrm[1], nrm[2]));
    *jmin = (nrm[0] <= nrm[1] && nrm[0] <= nrm[2]) ? 0 : (nrm[1] <= nrm[2] ? 1 : 2);      // the smallest singular value's column
    bool ok[3];
    for (int j = 0; j < 3; ++j) {
        ok[j] = nrm[j] > 1e-12 * big && nrm[j] > 0.0;
        for (int k = 0; k < 3; ++k) U[3 * k + j] = ok[j] ? B[3 * k + j] / nrm[j] : 0.0;
    }
    // complete U: a missing column is the cross product of the other two (or any unit vector orthogonal to one)
    for (int j = 0; j < 3; ++j) {
        if (ok[j]) continue;
        const int a = (j + 1) % 3, b = (j + 2) % 3;
        double ua[3] = {U[a], U[3 + a], U[6 + a]}, ub[3] = {U[b], U[3 + b], U[6 + b]};
        if (!ok[a] && !ok[b]) { ua[0] = 1; ua[1] = 0; ua[2] = 0; ub[0] = 0; ub[1] = 1; ub[2] = 0; U[a] = 1; U[3 + a] = 0; U[6 + a] = 0; U[b] = 0; U[3 + b] = 1; U[6 + b] = 0; ok[a] = ok[b] = true; }
        else if (!ok[a] || !ok[b]) {
            const double *u = ok[a] ? ua : ub;
            double e[3] = {0, 0, 0};
            const int m = (fabs(u[0]) <= fabs(u[1]) && fabs(u[0]) <= fabs(u[2])) ? 0 : (fabs(u[1]) <= fabs(u[2]) ? 1 : 2);
            e[m] = 1.0;
            double w[3] = {u[1] * e[2] - u[2] * e[1], u[2] * e[0] - u[0] * e[2], u[0] * e[1] - u[1] * e[0]};
            const double wn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
            const int miss = ok[a] ? b : a;
            for (int k = 0; k < 3; ++k) U[3 * k + miss] = w[k] / wn;
            ok[miss] = true;
            for (int k = 0; k < 3; ++k) { ua[k] = U[3 * k + a]; ub[k] = U[3 * k + b]; }
        }
        U[j] = ua[1] * ub[2] - ua[2] * ub[1];
        U[3 + j] = ua[2] * ub[0] - ua[0] * ub[2];
        U[6 + j] = ua[0] * ub[1] - ua[1] * ub[0];
        ok[j] = true;
    }
}

__device__ double det3(const double M[9])
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// mode 0: one ICP iteration; mode 2: apply the pending transform, then the fitness of the aligned source (no update)
__global__ void __launch_bounds__(ICP_THREADS)
icp_iter_kernel(float *__restrict__ src, int ns, const float *__restrict__ tgt, int nt, double max_d2, int max_iter,
                double trans_eps, double fit_eps, int min_corr, int mode, double *__restrict__ partial,
                IcpState *__restrict__ st)
{
    __shared__ float tile[ICP_TILE * 3];
    __shared__ double red[ICP_THREADS / 32][NSUM];
    __shared__ int is_last;
    if (mode == 0 && st->done) return;
    const int i = blockIdx.x * ICP_THREADS + threadIdx.x;
    float3 p = make_float3(0.f, 0.f, 0.f);
    if (i < ns) {
        p = make_float3(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
        p = xform(st->T_inc, p);                          // the previous iteration's transform, applied lazily
        src[3 * i] = p.x; src[3 * i + 1] = p.y; src[3 * i + 2] = p.z;
    }
    float best = FLT_MAX;
    int bi = -1;
    for (int base = 0; base < nt; base += ICP_TILE) {
        const int n = min(ICP_TILE, nt - base);
        __syncthreads();
        for (int e = threadIdx.x; e < 3 * n; e += ICP_THREADS) tile[e] = tgt[3 * (size_t)base + e];
        __syncthreads();
        if (i < ns) {
            for (int j = 0; j < n; ++j) {
                const float dx = p.x - tile[3 * j], dy = p.y - tile[3 * j + 1], dz = p.z - tile[3 * j + 2];
                const float d2 = dx * dx + dy * dy + dz * dz;
                if (d2 < best) { best = d2; bi = base + j; }
            }
        }
    }
    double v[NSUM];
#pragma unroll
    for (int k = 0; k < NSUM; ++k) v[k] = 0.0;
    const bool valid = i < ns && bi >= 0 && (mode == 2 || (double)best <= max_d2);
    if (valid) {
        const double sx = p.x, sy = p.y, sz = p.z;
        const double tx = tgt[3 * (size_t)bi], ty = tgt[3 * (size_t)bi + 1], tz = tgt[3 * (size_t)bi + 2];
        v[0] = 1.0; v[1] = sx; v[2] = sy; v[3] = sz; v[4] = tx; v[5] = ty; v[6] = tz;
        v[7] = sx * tx; v[8] = sx * ty; v[9] = sx * tz; v[10] = sy * tx; v[11] = sy * ty; v[12] = sy * tz;
        v[13] = sz * tx; v[14] = sz * ty; v[15] = sz * tz; v[16] = (double)best;
    }
    // fixed-order reduction: shuffle tree inside the warp, warps in order, blocks in order (last block)
#pragma unroll
    for (int k = 0; k < NSUM; ++k)
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        for (int k = 0; k < NSUM; ++k) red[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x < NSUM) {
        double s = 0.0;
        for (int w = 0; w < ICP_THREADS / 32; ++w) s += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * NSUM + threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(&st->ticket, 1u);
        is_last = (tk == gridDim.x - 1);
        if (is_last) st->ticket = 0u;
    }
    __syncthreads();
    if (!is_last || threadIdx.x != 0) return;
    __threadfence();
    double S[NSUM];
    for (int k = 0; k < NSUM; ++k) S[k] = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b)
        for (int k = 0; k < NSUM; ++k) S[k] += __ldcg(partial + (size_t)b * NSUM + k);
    if (mode == 2) { st->fitness = S[0] > 0 ? S[16] / S[0] : DBL_MAX; return; }
    const int cnt = (int)S[0];
    st->n_corr = cnt;
    if (cnt < min_corr) {                                                               // CONVERGENCE_CRITERIA_NO_CORRESPONDENCES
        st->done = 1; st->converged = 0; st->state = 5;
        for (int k = 0; k < 16; ++k) st->T_inc[k] = (k % 5 == 0) ? 1.0 : 0.0;            // nothing pending
        return;
    }
    const double n = S[0];
    const double cs[3] = {S[1] / n, S[2] / n, S[3] / n}, ct[3] = {S[4] / n, S[5] / n, S[6] / n};
    double H[9];                                       // sum (s - cs)(t - ct)^T
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) H[3 * a + b] = S[7 + 3 * a + b] - n * cs[a] * ct[b];
    double U[9], V[9];
    int jmin;
    svd3(H, U, V, &jmin);
    // R = V diag(1, 1, det(V) det(U)) U^T, the sign on the smallest singular value's pair of columns
    const double sgn = (det3(U) * det3(V) < 0) ? -1.0 : 1.0;
    double R[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            double r = 0.0;
            for (int j = 0; j < 3; ++j) r += (j == jmin ? sgn : 1.0) * V[3 * a + j] * U[3 * b + j];
            R[3 * a + b] = r;
        }
    double T[16] = {R[0], R[1], R[2], 0, R[3], R[4], R[5], 0, R[6], R[7], R[8], 0, 0, 0, 0, 1};
    for (int a = 0; a < 3; ++a) T[4 * a + 3] = ct[a] - (R[3 * a] * cs[0] + R[3 * a + 1] * cs[1] + R[3 * a + 2] * cs[2]);
    double F[16];
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += T[4 * a + k] * st->T_final[4 * k + b];
            F[4 * a + b] = s;
        }
    for (int k = 0; k < 16; ++k) { st->T_final[k] = F[k]; st->T_inc[k] = T[k]; }
    const int it = ++st->iterations;
    // DefaultConvergenceCriteria::hasConverged
    bool conv = false;
    int state = 0;
    if (it >= max_iter) { conv = true; state = 1; }                                    // ITERATIONS
    else {
        bool similar = false;
        const double cos_angle = 0.5 * (T[0] + T[5] + T[10] - 1.0);
        const double tr2 = T[3] * T[3] + T[7] * T[7] + T[11] * T[11];
        if (cos_angle >= 1.0 - trans_eps && tr2 <= trans_eps) { conv = true; state = 2; similar = true; }     // TRANSFORM (max_iterations_similar_transforms_ = 0)
        const double mse = S[16] / n;
        if (!conv) {
            if (fabs(mse - st->prev_mse) < 1e-12) { conv = true; state = 3; }            // ABS_MSE
            else if (fabs(mse - st->prev_mse) / st->prev_mse < fit_eps) { conv = true; state = 4; }   // REL_MSE
        }
        st->prev_mse = mse;
        st->similar = similar ? st->similar + 1 : 0;
    }
    if (conv) { st->done = 1; st->converged = 1; st->state = state; }
}

}  // namespace

extern "C" {

void nrs_icp_default_params(nrs_icp_params *p)
{
    if (!p) return;
    p->max_correspondence_distance = 150.0;      // laserPosegraphOptimization.cpp:377
    p->max_iterations = 100;                     // :378
    p->transformation_epsilon = 1e-6;            // :379
    p->euclidean_fitness_epsilon = 1e-6;         // :380
    p->fitness_threshold = 0.3;                  // :390
    p->min_correspondences = 3;                  // pcl::Registration::min_number_correspondences_
}

int nrs_icp_align(nrs_ctx *ctx, const float *src_xyz, int ns, const float *tgt_xyz, int nt, const nrs_icp_params *p,
                  double T_out[16], nrs_icp_info *info)
{
    NRS_CHECK_ARG(ctx && p && T_out && info && ns >= 0 && nt >= 0 && (ns == 0 || src_xyz) && (nt == 0 || tgt_xyz), "bad argument");
    NRS_CHECK_ARG(p->max_iterations >= 1 && p->max_iterations <= 10000, "max_iterations out of range");
    std::lock_guard<std::mutex> lk(ctx->mu);
    NRS_CUDA(cudaSetDevice(ctx->device));
    memset(info, 0, sizeof(*info));
    for (int i = 0; i < 16; ++i) T_out[i] = (i % 5 == 0) ? 1.0 : 0.0;
    if (ns == 0 || nt == 0) { info->state = 5; info->fitness = DBL_MAX; return NRS_OK; }     // empty cloud: PCL refuses to align
    const int blocks = (ns + ICP_THREADS - 1) / ICP_THREADS;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_src = take((size_t)ns * 12), o_tgt = take((size_t)nt * 12), o_part = take((size_t)blocks * NSUM * 8), o_st = take(sizeof(IcpState));
    int rc = ctx->scratch[2].reserve(off);
    if (rc) return rc;
    unsigned char *b = ctx->scratch[2].as<unsigned char>();
    float *d_src = (float *)(b + o_src), *d_tgt = (float *)(b + o_tgt);
    double *d_part = (double *)(b + o_part);
    IcpState *d_st = (IcpState *)(b + o_st);
    cudaStream_t st = ctx->stream;
    IcpState h = {};
    for (int i = 0; i < 16; ++i) h.T_inc[i] = h.T_final[i] = (i % 5 == 0) ? 1.0 : 0.0;
    h.prev_mse = DBL_MAX;                        // correspondences_prev_mse_ starts at numeric_limits<double>::max()
    NRS_CUDA(cudaMemcpyAsync(d_src, src_xyz, (size_t)ns * 12, cudaMemcpyHostToDevice, st));
    NRS_CUDA(cudaMemcpyAsync(d_tgt, tgt_xyz, (size_t)nt * 12, cudaMemcpyHostToDevice, st));
    NRS_CUDA(cudaMemcpyAsync(d_st, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    const double max_d2 = p->max_correspondence_distance * p->max_correspondence_distance;
    IcpState hs;
    // iterations in chunks of 8 launches; a finished alignment turns the remaining launches of a chunk into no-ops
    for (int it = 0; it < p->max_iterations;) {
        const int chunk = std::min(8, p->max_iterations - it);
        for (int k = 0; k < chunk; ++k) {
            icp_iter_kernel<<<blocks, ICP_THREADS, 0, st>>>(d_src, ns, d_tgt, nt, max_d2, p->max_iterations, p->transformation_epsilon,
                                                            p->euclidean_fitness_epsilon, p->min_correspondences, 0, d_part, d_st);
            NRS_LAUNCHED();
        }
        it += chunk;
        NRS_CUDA(cudaMemcpyAsync(&hs, d_st, sizeof(hs), cudaMemcpyDeviceToHost, st));
        NRS_CUDA(cudaStreamSynchronize(st));
        if (hs.done) break;
    }
    // the last computed transform is still pending on the source (it is applied lazily by the next launch): the
    // fitness pass applies it, then measures the aligned source against the target without a range limit
    icp_iter_kernel<<<blocks, ICP_THREADS, 0, st>>>(d_src, ns, d_tgt, nt, max_d2, p->max_iterations, p->transformation_epsilon,
                                                    p->euclidean_fitness_epsilon, p->min_correspondences, 2, d_part, d_st);
    NRS_LAUNCHED();
    NRS_CUDA(cudaMemcpyAsync(&hs, d_st, sizeof(hs), cudaMemcpyDeviceToHost, st));
    NRS_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) T_out[4 * c + r] = hs.T_final[4 * r + c];       // column-major like Eigen
    info->converged = hs.converged;
    info->iterations = hs.iterations;
    info->state = hs.state;
    info->n_correspondences = hs.n_corr;
    info->fitness = hs.fitness;
    info->accepted = (hs.converged && hs.fitness <= p->fitness_threshold) ? 1 : 0;      // laserPosegraphOptimization.cpp:391
    return NRS_OK;
}

}  // extern "C"
