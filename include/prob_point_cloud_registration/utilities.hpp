// calculateMSE -- the one helper of the reference's utilities.hpp that is actually called
// (.../utilities.hpp:16-26; call sites src/prob_point_cloud_registration.cc:59,115,121,133 and the CLI :186).
// Despite its name it is the MEAN EUCLIDEAN DISTANCE between corresponding points, evaluated in float32 like
// pcl::euclideanDistance.  The other helpers of that header are dead code in the reference and are not provided.
#ifndef PROB_POINT_CLOUD_REGISTRATION_UTILITIES_HPP
#define PROB_POINT_CLOUD_REGISTRATION_UTILITIES_HPP
#include <cassert>
#include <cmath>

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

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
}  // namespace prob_point_cloud_registration
#endif
