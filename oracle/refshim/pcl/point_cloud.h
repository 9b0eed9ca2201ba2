// Stand-in for <pcl/point_cloud.h> (TEST INFRASTRUCTURE, see oracle/refshim/README.md)
#ifndef REFSHIM_PCL_POINT_CLOUD
#define REFSHIM_PCL_POINT_CLOUD
#include <memory>
#include <vector>
namespace pcl {
template <class T> struct PointCloud {
    using Ptr = std::shared_ptr<PointCloud<T>>;
    using ConstPtr = std::shared_ptr<const PointCloud<T>>;
    std::vector<T> points;
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); }
    size_t size() const { return points.size(); }
};
}  // namespace pcl
#endif
