// Stand-in for <pcl/point_cloud.h> (TEST INFRASTRUCTURE, see oracle/refshim/README.md)
#ifndef REFSHIM_PCL_POINT_CLOUD
#define REFSHIM_PCL_POINT_CLOUD
#include <fstream>            // the real PCL headers pull these in; the reference's solver header relies on that
#include <memory>
#include <unordered_map>
#include <vector>
namespace pcl {
template <class T> struct PointCloud {
    using Ptr = std::shared_ptr<PointCloud<T>>;
    using ConstPtr = std::shared_ptr<const PointCloud<T>>;
    std::vector<T> points;
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); }
    size_t size() const { return points.size(); }
    void reserve(size_t n) { points.reserve(n); }
};
}  // namespace pcl
namespace boost {             // PCL's headers used to bring boost::shared_ptr along; the solver header relies on that
template <class T> using shared_ptr = std::shared_ptr<T>;
}
#endif
