// ref_sc_wrap.cpp -- extern "C" doors into the REFERENCE's own ScanContext manager, compiled from the source where it
// lies (pgo/SC-A-LOAM/include/scancontext/Scancontext.cpp + Scancontext.h, unmodified, with the reference's own vendored
// nanoflann and KDTreeVectorOfVectorsAdaptor) against the stand-in headers of oracle/refshim/.  Eigen is not in this
// image: the stand-in MatrixXd evaluates norm / dot / mean in the order Eigen 3.3's SSE2 redux uses (recalled, see
// refshim/Eigen/Dense) -- so this pins SCManager's LOGIC (bin assignment, key construction, circshift, the 60-shift
// alignment, the +-3 search window, the arg-mins and their strictness, the kd-tree candidate flow with its rebuild
// period and exclusion window, the between-session one-shot tree) exactly, and its floating-point reductions up to
// that one stated assumption.  Recipe: oracle/Makefile `ref`; output oracle/_ref/libref_sc.so.  TEST INFRASTRUCTURE
// ONLY (tests/test_ref_sc.py); nothing in the product links or loads it.
#include <cstring>

#include "scancontext/Scancontext.h"

extern "C" {

RefCvCallbacks g_ref_cv;                // the OpenCV stand-in's callback table (Scancontext.h includes OpenCV, calls none of it)

void *ref_sc_create() { return new SCManager(); }
void ref_sc_destroy(void *m) { delete (SCManager *)m; }
void ref_sc_set_thres(void *m, double t) { ((SCManager *)m)->setSCdistThres(t); }
int ref_sc_size(void *m) { return (int)((SCManager *)m)->polarcontexts_.size(); }

static pcl::PointCloud<SCPointType> cloud_from(const float *xyz, int n)
{
    pcl::PointCloud<SCPointType> c;
    c.points.resize((size_t)n);
    for (int i = 0; i < n; ++i) { c.points[i].x = xyz[3 * i]; c.points[i].y = xyz[3 * i + 1]; c.points[i].z = xyz[3 * i + 2]; }
    return c;
}
static MatrixXd mat_from(const double *p, int rows, int cols)
{
    MatrixXd m(rows, cols);
    std::memcpy(m.data(), p, sizeof(double) * (size_t)rows * cols);
    return m;
}

// Scancontext.cpp:151-195 -> 20 x 60 doubles, column-major
void ref_sc_make(void *m, const float *xyz, int n, double *desc)
{
    auto c = cloud_from(xyz, n);
    const MatrixXd d = ((SCManager *)m)->makeScancontext(c);
    std::memcpy(desc, d.data(), sizeof(double) * (size_t)d.size());
}
// :198-227
void ref_sc_keys(void *m, const double *desc, double *ringkey20, double *sectorkey60)
{
    MatrixXd d = mat_from(desc, 20, 60);
    const MatrixXd r = ((SCManager *)m)->makeRingkeyFromScancontext(d), s = ((SCManager *)m)->makeSectorkeyFromScancontext(d);
    std::memcpy(ringkey20, r.data(), sizeof(double) * 20);
    std::memcpy(sectorkey60, s.data(), sizeof(double) * 60);
}
// :93-113
int ref_sc_fast_align(void *m, const double *vkey1, const double *vkey2)
{
    MatrixXd a = mat_from(vkey1, 1, 60), b = mat_from(vkey2, 1, 60);
    return ((SCManager *)m)->fastAlignUsingVkey(a, b);
}
// :69-90
double ref_sc_dist_direct(void *m, const double *sc1, const double *sc2)
{
    MatrixXd a = mat_from(sc1, 20, 60), b = mat_from(sc2, 20, 60);
    return ((SCManager *)m)->distDirectSC(a, b);
}
// :116-148
double ref_sc_distance(void *m, const double *sc1, const double *sc2, int *shift)
{
    MatrixXd a = mat_from(sc1, 20, 60), b = mat_from(sc2, 20, 60);
    const std::pair<double, int> r = ((SCManager *)m)->distanceBtnScanContext(a, b);
    *shift = r.second;
    return r.first;
}
// :39-59
void ref_sc_circshift(const double *mat, int rows, int cols, int shift, double *out)
{
    MatrixXd a = mat_from(mat, rows, cols);
    const MatrixXd s = circshift(a, shift);
    std::memcpy(out, s.data(), sizeof(double) * (size_t)rows * cols);
}
// :249-260 / :236-247
void ref_sc_make_and_save(void *m, const float *xyz, int n) { auto c = cloud_from(xyz, n); ((SCManager *)m)->makeAndSaveScancontextAndKeys(c); }
void ref_sc_save(void *m, const double *desc) { ((SCManager *)m)->saveScancontextAndKeys(mat_from(desc, 20, 60)); }
void ref_sc_get(void *m, int idx, double *desc, float *ringkey20_f32)
{
    SCManager *s = (SCManager *)m;
    std::memcpy(desc, s->polarcontexts_[(size_t)idx].data(), sizeof(double) * 1200);
    std::memcpy(ringkey20_f32, s->polarcontext_invkeys_mat_[(size_t)idx].data(), sizeof(float) * 20);
}
// :331-422
int ref_sc_detect(void *m, float *yaw) { const std::pair<int, float> r = ((SCManager *)m)->detectLoopClosureID(); *yaw = r.second; return r.first; }
// :267-328
int ref_sc_detect_between(void *m, const float *key20, const double *desc, float *yaw)
{
    std::vector<float> k(key20, key20 + 20);
    MatrixXd d = mat_from(desc, 20, 60);
    const std::pair<int, float> r = ((SCManager *)m)->detectLoopClosureIDBetweenSession(k, d);
    *yaw = r.second;
    return r.first;
}

}  // extern "C"
