// prob_point_cloud_registration::ProbPointCloudRegistration -- the reference's public class
// (reference: include/prob_point_cloud_registration/prob_point_cloud_registration.h:18-64), same constructors and
// member functions, implemented over the C ABI of libppcr_cuda.so (include/ppcr.h): the whole outer loop of align()
// runs on one B200.  Drop-in for programs written against the reference:
//
//   * the source cloud is deep-copied, the caller's is never modified            (registration.cc:22)
//   * the target cloud is aliased and voxel-filtered IN PLACE when target_filter_size > 0   (:19,34-41)
//   * hasConverged() mutates the stall counter exactly like the reference        (:138-158)
//   * transformation() is history.back(); calling it before any iteration ran is undefined in the reference and
//     throws std::out_of_range here
//   * report() returns the 12-column CSV of :44-46,120-129 when params.summary is set
//
// There is no CPU fallback: construction throws std::runtime_error when no sm_100 GPU is usable.
#ifndef PROB_POINT_CLOUD_REGISTRATION_POINT_CLOUD_REGISTRATION_HPP
#define PROB_POINT_CLOUD_REGISTRATION_POINT_CLOUD_REGISTRATION_HPP

#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "prob_point_cloud_registration/output_stream.hpp"
#include "prob_point_cloud_registration/prob_point_cloud_registration_params.hpp"

struct ppcr_handle;

namespace prob_point_cloud_registration {

class ProbPointCloudRegistration {
public:
    ProbPointCloudRegistration(pcl::PointCloud<pcl::PointXYZ>::Ptr source_cloud,
                               pcl::PointCloud<pcl::PointXYZ>::Ptr target_cloud,
                               ProbPointCloudRegistrationParams parameters);
    ProbPointCloudRegistration(pcl::PointCloud<pcl::PointXYZ>::Ptr source_cloud,
                               pcl::PointCloud<pcl::PointXYZ>::Ptr target_cloud,
                               ProbPointCloudRegistrationParams parameters,
                               pcl::PointCloud<pcl::PointXYZ>::Ptr ground_truth_cloud);
    ~ProbPointCloudRegistration();
    ProbPointCloudRegistration(const ProbPointCloudRegistration&) = delete;
    ProbPointCloudRegistration& operator=(const ProbPointCloudRegistration&) = delete;

    void align();
    bool hasConverged();
    inline Eigen::Affine3d transformation() { return transformation_history_.at(transformation_history_.size() - 1); }
    inline std::vector<Eigen::Affine3d> transformation_history() { return transformation_history_; }
    inline std::string report() { return report_.str(); }

private:
    void init();

    ProbPointCloudRegistrationParams parameters_;
    pcl::PointCloud<pcl::PointXYZ>::Ptr target_cloud_;
    pcl::PointCloud<pcl::PointXYZ>::Ptr source_cloud_;        // full-resolution copy, moved alongside for the MSE metrics
    pcl::PointCloud<pcl::PointXYZ>::Ptr prev_source_cloud_;
    pcl::PointCloud<pcl::PointXYZ>::Ptr ground_truth_cloud_;
    bool ground_truth_;
    double mse_ground_truth_;
    double mse_prev_it_;
    int reported_iterations_;  // outer iterations already folded into history / report
    OutputStream output_stream_;
    std::vector<Eigen::Affine3d> transformation_history_;
    std::stringstream report_;
    ppcr_handle* handle_;
};

}  // namespace prob_point_cloud_registration

#endif
