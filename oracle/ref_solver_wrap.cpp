// ref_solver_wrap.cpp -- extern "C" doors into the REFERENCE's own estimator, compiled from the sources where they lie
// (outlier-robust-radar-odometry/include/ORORA.hpp, include/solver/nonminimal_solver.hpp, include/solver/graph.h,
// src/graph.cc, unmodified) against the stand-in headers of oracle/refshim/.  What that pins and what it does not:
//   * pinned: the reference's LOGIC -- Doppler compensation and the velocity carried from pair to pair, the stop-motion
//     gate, TIM construction and its index map, the scale-consistency test, graph construction, the chain TIMs over the
//     sorted clique, GNC-TLS (initial weights, mu schedule, EPSILON only before the loop, the thresholds, the 0.4 inlier
//     rule and the wrap-around rotation-inlier count), COTE per axis (the sorted 2N interval ends, running sums,
//     minCoeff's first minimum, the inlier test), the composition of the 4 x 4 result;
//   * NOT pinned: Eigen's floating-point evaluation order (Eigen is not in this image; the stand-in evaluates products
//     and sums in plain index order and the 2 x 2 SVD in closed form) -- hence the comparison is to north_star's pose
//     tolerance with identical index sets, not bit for bit; and PMC (un-vendored): refshim/pmc/pmc.h hands the exact
//     search to the oracle's own max-clique routine, so A13 stays unpinned.
// association.cpp's SE3tose3 (needs Eigen's ComplexEigenSolver) is replaced below by what it returns on this path.
// Recipe: oracle/Makefile `ref`; output oracle/_ref/libref_solver.so.  TEST INFRASTRUCTURE ONLY (tests/test_ref_solver.py).
#include <cstring>

#include "radar_loader.hpp"      // RadarFeature (the reference's)

// association.cpp:179-222 ends with `rho = T.block(0, 3, 3, 1)` UNCONDITIONALLY (:211-212 overwrite the J^-1 branch), so
// xi(0..2) is the translation; xi(3..5) = phi * axis is read only by motionDeskewing (off in the shipped configuration).
Eigen::VectorXd SE3tose3(Eigen::MatrixXd T)
{
    Eigen::VectorXd xi = Eigen::VectorXd::Zero(6);
    for (int i = 0; i < 3; ++i) xi(i) = std::isnan(T(i, 3)) ? 0 : T(i, 3);
    double trace = T(0, 0) + T(1, 1) + T(2, 2);
    double phi = std::acos((trace - 1) / 2);
    if (std::isnan(phi)) phi = 0.0;
    xi(5) = (T(1, 0) < 0 ? -phi : phi);       // planar motion: the axis is +-z
    return xi;
}

#include "ORORA.hpp"
#include "pmc/pmc.h"             // the stand-in's callback type

namespace {
struct OpenSolver : orora::NonMinimalSolver {
    using orora::NonMinimalSolver::max_clique_;
    using orora::NonMinimalSolver::final_inliers_;
    using orora::NonMinimalSolver::scale_inliers_mask_;
};
Eigen::MatrixXd mat2(const double *p, int n)
{
    Eigen::MatrixXd m(2, n);
    if (n) std::memcpy(m.data(), p, sizeof(double) * 2 * (size_t)n);
    return m;
}
orora::NonMinimalSolver::Params make_params(double noise_bound, double cbar2, int max_iter, double gnc_factor, double cost_thr)
{
    orora::NonMinimalSolver::Params p;
    orora::setParams(noise_bound, cbar2, false, max_iter, gnc_factor, cost_thr, p);          // nonminimal_solver.hpp:959-973
    return p;
}
}  // namespace

extern "C" {

RefCvCallbacks g_ref_cv;
RefMaxCliqueFn g_ref_max_clique;
void ref_solver_set_max_clique(RefMaxCliqueFn f) { g_ref_max_clique = f; }

// NonMinimalSolver on its own (nonminimal_solver.hpp:665-825): src = current frame's matches, tgt = previous frame's.
// counts6: [0] valid, [1] scale-consistent TIM pairs (graph edges), [2] clique size, [3] rotation inliers, [4] final inliers
int ref_solver_run(const double *src_2xN, const double *tgt_2xN, int N, double noise_bound, double cbar2, int max_iter,
                   double gnc_factor, double cost_thr, double noise_bound_radial, double noise_bound_tangential, double *T16,
                   int *clique, int *final_inliers, int *counts6)
{
    OpenSolver s;
    s.reset(make_params(noise_bound, cbar2, max_iter, gnc_factor, cost_thr), noise_bound_radial, noise_bound_tangential);
    s.setInputSource(mat2(src_2xN, N));
    s.setInputTarget(mat2(tgt_2xN, N));
    Eigen::Matrix4d guess = Eigen::Matrix4d::Identity(), out;
    s.computeTransformation(guess, out);
    std::memcpy(T16, out.data(), sizeof(double) * 16);
    int edges = 0;
    for (long i = 0; i < s.scale_inliers_mask_.cols(); ++i) edges += s.scale_inliers_mask_(0, i) ? 1 : 0;
    counts6[0] = s.solution_.valid ? 1 : 0; counts6[1] = edges; counts6[2] = (int)s.max_clique_.size();
    counts6[3] = s.getNumRotaionInliers(); counts6[4] = (int)s.final_inliers_.size();
    for (size_t i = 0; i < s.max_clique_.size(); ++i) clique[i] = s.max_clique_[i];
    for (size_t i = 0; i < s.final_inliers_.size(); ++i) final_inliers[i] = s.final_inliers_[i];
    return 0;
}

// orora::ORORA (ORORA.hpp:13-181), the object the node keeps across scan pairs
void *ref_orora_create(const char *dataset, const char *deskewing_target, int use_doppler, int use_deskewing, int check_stop_motion,
                       int num_feat_thr, double beta, double noise_bound_radial, double noise_bound_tangential)
{
    return new orora::ORORA(dataset, deskewing_target, use_doppler != 0, use_deskewing != 0, check_stop_motion != 0, num_feat_thr, beta,
                            noise_bound_radial, noise_bound_tangential);
}
void ref_orora_destroy(void *h) { delete (orora::ORORA *)h; }

// calcTransform (ORORA.hpp:37-93): prev / curr = matched points 2 x N column-major; returns the 4 x 4 (column-major) and
// the public feat_compensated_ of both features
int ref_orora_calc(void *h, const double *prev_2xN, const double *curr_2xN, int N, double noise_bound, double cbar2, int max_iter,
                   double gnc_factor, double cost_thr, double delta_t, double *T16, double *prev_comp_2xN, double *curr_comp_2xN)
{
    RadarFeature prev, curr;
    prev.feat_in_cart_ = mat2(prev_2xN, N);
    curr.feat_in_cart_ = mat2(curr_2xN, N);
    const Eigen::MatrixXd T = ((orora::ORORA *)h)->calcTransform(make_params(noise_bound, cbar2, max_iter, gnc_factor, cost_thr), delta_t, prev, curr);
    std::memcpy(T16, T.data(), sizeof(double) * 16);
    if (N) {
        std::memcpy(prev_comp_2xN, prev.feat_compensated_.data(), sizeof(double) * 2 * (size_t)N);
        std::memcpy(curr_comp_2xN, curr.feat_compensated_.data(), sizeof(double) * 2 * (size_t)N);
    }
    return 0;
}

}  // extern "C"
