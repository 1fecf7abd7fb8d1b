// TEST INFRASTRUCTURE ONLY: the reference's map_util.h (thirdparty/jps3d/include/jps_collision/map_util.h:11,:30) names
// pcl::PointCloud<pcl::PointXYZ>::Ptr in the signature of readMap and reads ->points[i].x/.y/.z; nothing else of PCL.
#pragma once
#include <memory>
#include <vector>
namespace pcl
{
struct PointXYZ
{
  float x, y, z;
};
template <class T>
struct PointCloud
{
  typedef std::shared_ptr<PointCloud<T>> Ptr;
  std::vector<T> points;
};
}  // namespace pcl
