// The metric helpers of the reference's utilities.hpp.
//
// calculateMSE (.../utilities.hpp:16-26; call sites src/prob_point_cloud_registration.cc:59,115,121,133 and the CLI :186)
// is the only one the reference calls.  Despite its name it is the MEAN EUCLIDEAN DISTANCE between corresponding points,
// evaluated in float32 like pcl::euclideanDistance; it runs on the host here as there (the per-iteration uses of it inside
// align() are replayed on the device, ppcr_replay_metrics).
//
// The closest-point helpers (:28-234) each build a pcl::KdTreeFLANN on cloud2 and reduce the nearestKSearch(k = 1) SQUARED
// distances of cloud1's points.  Here they share one device call, ppcr_closest_point_metrics: octree build + 1-NN search +
// sort + reduction on the GPU.  Same signatures, same index rules (the reference's "median" is element (n+1)/2 resp. the
// mean of elements n/2 and n/2+1 of the sorted vector), same DBL_MAX for fewer than 10 inliers.
#ifndef PROB_POINT_CLOUD_REGISTRATION_UTILITIES_HPP
#define PROB_POINT_CLOUD_REGISTRATION_UTILITIES_HPP
#include <cassert>
#include <cmath>
#include <stdexcept>
#include <string>

#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "ppcr.h"

namespace prob_point_cloud_registration {
inline double calculateMSE(pcl::PointCloud<pcl::PointXYZ>::Ptr first_cloud, pcl::PointCloud<pcl::PointXYZ>::Ptr second_cloud)
{
    assert(first_cloud->size() == second_cloud->size());
    double mse = 0;
    for (std::size_t i = 0; i < first_cloud->size(); i++) {
        const float dx = first_cloud->at(i).x - second_cloud->at(i).x;
        const float dy = first_cloud->at(i).y - second_cloud->at(i).y;
        const float dz = first_cloud->at(i).z - second_cloud->at(i).z;
        mse += std::sqrt(dx * dx + dy * dy + dz * dz);
    }
    return mse / first_cloud->size();
}

// one device pass serves all seven helpers
inline ppcr_closest_metrics closestPointMetrics(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1,
                                                pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2, double factor = 3.0)
{
    ppcr_closest_metrics m;
    const ppcr_status s = ppcr_closest_point_metrics(reinterpret_cast<const float*>(cloud1->points.data()),
                                                     static_cast<int64_t>(cloud1->size()),
                                                     reinterpret_cast<const float*>(cloud2->points.data()),
                                                     static_cast<int64_t>(cloud2->size()), factor, nullptr, &m, nullptr);
    if (s != PPCR_OK) throw std::runtime_error(std::string("ppcr_closest_point_metrics: ") + ppcr_last_error());
    return m;
}

inline double averageClosestDistance(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2)
{
    return closestPointMetrics(cloud1, cloud2).average_closest_distance;  // :28-46
}
inline double sumSquaredError(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2)
{
    return closestPointMetrics(cloud1, cloud2).sum_squared_error;  // :48-65
}
inline double robustSumSquaredError(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2)
{
    return closestPointMetrics(cloud1, cloud2).robust_sum_squared_error;  // :67-101
}
inline double robustSumSquaredError(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2,
                                    double factor)
{
    return closestPointMetrics(cloud1, cloud2, factor).robust_sum_squared_error_factor;  // :103-138
}
inline double robustAveragedSumSquaredError(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1,
                                            pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2)
{
    return closestPointMetrics(cloud1, cloud2).robust_averaged_sum_squared_error;  // :140-175
}
inline double medianClosestDistance(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2)
{
    return closestPointMetrics(cloud1, cloud2).median_closest_distance;  // :177-199
}
inline double robustMedianClosestDistance(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1,
                                          pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2)
{
    return closestPointMetrics(cloud1, cloud2).robust_median_closest_distance;  // :201-234
}

// :250-262
inline Eigen::Quaterniond euler2Quaternion(const double roll, const double pitch, const double yaw)
{
    const Eigen::AngleAxisd rollAngle(roll, Eigen::Vector3d::UnitX());
    const Eigen::AngleAxisd pitchAngle(pitch, Eigen::Vector3d::UnitY());
    const Eigen::AngleAxisd yawAngle(yaw, Eigen::Vector3d::UnitZ());
    return yawAngle * pitchAngle * rollAngle;
}
}  // namespace prob_point_cloud_registration
#endif
