// oracle/stdsort.cpp -- TEST INFRASTRUCTURE (see oracle/__init__.py).
// cen2019features orders its candidates with an UNSTABLE std::sort on the float response alone
// (outlier-robust-radar-odometry/src/features.cpp:133-149,224), so among equal responses the order -- and with it which
// members of the tie group at the max_points cut-off get marked -- is whatever this toolchain's libstdc++ introsort
// produces from the row-major candidate list.  The oracle's default is the stable order (what the GPU path implements);
// this helper gives the oracle the libstdc++ order too, so that it can be compared EXACTLY with the reference's own
// compiled code (oracle/_ref/libref_frontend.so) and the tie sensitivity can be measured instead of argued.
#include <algorithm>
#include <vector>

namespace {
struct Cand {
    float i;
    int a, r;
};
}  // namespace

extern "C" void oracle_std_sort_desc(const float *h, int *a, int *r, int n)
{
    std::vector<Cand> v((size_t)n);
    for (int k = 0; k < n; ++k) v[(size_t)k] = Cand{h[k], a[k], r[k]};
    std::sort(v.begin(), v.end(), [](const Cand &p1, const Cand &p2) { return p1.i > p2.i; });
    for (int k = 0; k < n; ++k) { a[k] = v[(size_t)k].a; r[k] = v[(size_t)k].r; }
}
