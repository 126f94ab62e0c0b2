// pcl::transformPointCloud(in, out, Eigen::Affine3d): double arithmetic, float32 store; in place allowed.
#ifndef PPCR_COMPAT_PCL_TRANSFORMS_H
#define PPCR_COMPAT_PCL_TRANSFORMS_H
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
namespace pcl {
template <typename PointT>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Affine3d& T)
{
    if (&in != &out) {
        out.points.resize(in.points.size());
        out.width = in.width;
        out.height = in.height;
        out.is_dense = in.is_dense;
    }
    for (std::size_t i = 0; i < in.points.size(); ++i) {
        const double x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
        PointT p = in.points[i];
        p.x = static_cast<float>(T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3));
        p.y = static_cast<float>(T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3));
        p.z = static_cast<float>(T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3));
        out.points[i] = p;
    }
}
inline double rad2deg(double a) { return a * 180.0 / M_PI; }
}  // namespace pcl
#endif
