// Stand-in for <pcl/point_types.h> (TEST INFRASTRUCTURE, see oracle/refshim/README.md)
#ifndef REFSHIM_PCL_POINT_TYPES
#define REFSHIM_PCL_POINT_TYPES
namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; };
struct PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; };
}  // namespace pcl
#endif
