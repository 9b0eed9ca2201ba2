// Stand-in for <pcl/filters/voxel_grid.h> (TEST INFRASTRUCTURE, see oracle/refshim/README.md).
// NOT A PIN: PCL is not in this image, so this VoxelGrid is the same RECALLED restatement of pcl::VoxelGrid::applyFilter
// that oracle/frontend_ref.py::voxelize states (float arithmetic, voxel index ijk0 + ijk1*dx + ijk2*dx*dy with
// ijk = floor(p * inverse_leaf) - min_b, points ordered by index (ties in input order), centroid = float running sum /
// count, output in ascending index order).  It exists so that RadarLoader::extractFeature compiles and runs with
// use_voxelization on; row A5 stays "parity unpinned" whatever this file says.
#ifndef REFSHIM_PCL_VOXEL_GRID
#define REFSHIM_PCL_VOXEL_GRID
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>
#include "pcl/point_cloud.h"
#include "pcl/point_types.h"
namespace pcl {
template <class T> class VoxelGrid {
public:
    void setInputCloud(const typename PointCloud<T>::Ptr &c) { in_ = c; }
    void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
    void filter(PointCloud<T> &out)
    {
        out.points.clear();
        const std::vector<T> &p = in_->points;
        if (p.empty()) return;
        float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
        float mx[3] = {-mn[0], -mn[1], -mn[2]};
        for (const T &q : p) {
            const float v[3] = {q.x, q.y, q.z};
            for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], v[k]); mx[k] = std::max(mx[k], v[k]); }
        }
        float inv[3];
        int min_b[3], max_b[3], div_b[3];
        for (int k = 0; k < 3; ++k) {
            inv[k] = 1.0f / leaf_[k];
            min_b[k] = (int)std::floor(mn[k] * inv[k]);
            max_b[k] = (int)std::floor(mx[k] * inv[k]);
            div_b[k] = max_b[k] - min_b[k] + 1;
        }
        const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
        struct IdxPt { unsigned idx; unsigned pt; };
        std::vector<IdxPt> ip(p.size());
        for (size_t i = 0; i < p.size(); ++i) {
            const int i0 = (int)std::floor(p[i].x * inv[0]) - min_b[0], i1 = (int)std::floor(p[i].y * inv[1]) - min_b[1],
                      i2 = (int)std::floor(p[i].z * inv[2]) - min_b[2];
            ip[i] = IdxPt{(unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), (unsigned)i};
        }
        std::stable_sort(ip.begin(), ip.end(), [](const IdxPt &a, const IdxPt &b) { return a.idx < b.idx; });
        for (size_t a = 0; a < ip.size();) {
            size_t b = a;
            float sx = 0, sy = 0, sz = 0;
            while (b < ip.size() && ip[b].idx == ip[a].idx) { sx += p[ip[b].pt].x; sy += p[ip[b].pt].y; sz += p[ip[b].pt].z; ++b; }
            const float n = (float)(b - a);
            T c;
            c.x = sx / n; c.y = sy / n; c.z = sz / n;
            out.points.push_back(c);
            a = b;
        }
    }
private:
    typename PointCloud<T>::Ptr in_;
    float leaf_[3] = {1, 1, 1};
};
}  // namespace pcl
#endif
