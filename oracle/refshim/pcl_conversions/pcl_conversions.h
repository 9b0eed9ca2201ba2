// Stand-in for <pcl_conversions/pcl_conversions.h> (TEST INFRASTRUCTURE, see oracle/refshim/README.md): the ROS message
// type conversion.hpp's cloud2msg names; never called by the pinned functions.
#ifndef REFSHIM_PCL_CONVERSIONS
#define REFSHIM_PCL_CONVERSIONS
#include <cstdlib>
#include <string>
#include "pcl/point_cloud.h"
namespace sensor_msgs {
struct PointCloud2 { struct { std::string frame_id; } header; };
}
namespace pcl {
template <class T> inline void toROSMsg(const PointCloud<T> &, sensor_msgs::PointCloud2 &) { std::abort(); }
}
#endif
