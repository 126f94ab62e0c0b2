// ProbPointCloudRegistration over the C ABI of libppcr_cuda.so.
//
// The reference class (src/prob_point_cloud_registration.cc) owns the outer loop on the host: kd-tree, Ceres
// problem, cloud move, convergence test, one outer iteration at a time.  Here the constructor hands both clouds to
// the device once (ppcr_create: copy, voxel filters, target octree), align() is a single call that runs every outer
// iteration on the GPU (ppcr_align), and the host only reads back what the reference exposes afterwards: the pose
// history, the per-iteration statistics for report(), and -- because the reference filters the caller's target in
// place -- the filtered target cloud.  The verbose / ground-truth / summary diagnostics of the reference are replayed
// on the host from the per-iteration increments; they are outside the timed hot path.
#include "prob_point_cloud_registration/prob_point_cloud_registration.h"

#include <cmath>
#include <limits>
#include <stdexcept>

#include <pcl/common/transforms.h>

#include "ppcr.h"
#include "prob_point_cloud_registration/utilities.hpp"

namespace prob_point_cloud_registration {

namespace {

void check(ppcr_status s, const char* what)
{
    if (s != PPCR_OK) throw std::runtime_error(std::string(what) + ": " + ppcr_last_error());
}

ppcr_params to_c(const ProbPointCloudRegistrationParams& p)
{
    ppcr_params c;
    ppcr_default_params(&c);
    c.max_neighbours = p.max_neighbours;
    c.dof = p.dof;
    c.radius = p.radius;
    c.n_iter = p.n_iter;
    c.cost_drop_thresh = p.cost_drop_thresh;
    c.n_cost_drop_it = p.n_cost_drop_it;
    c.verbose = p.verbose ? 1 : 0;
    c.summary = p.summary ? 1 : 0;
    for (int k = 0; k < 4; ++k) c.initial_rotation[k] = p.initial_rotation[k];
    for (int k = 0; k < 3; ++k) c.initial_translation[k] = p.initial_translation[k];
    c.source_filter_size = p.source_filter_size;
    c.target_filter_size = p.target_filter_size;
    return c;
}

// The finite points of a cloud that is flagged !is_dense, in order.  PCL's own stages drop the others: VoxelGrid skips
// non-finite points when the input is not dense, and KdTreeFLANN leaves them out of the tree it builds
// (registration.cc:27-30,37-40,66-67).  The C ABI takes finite clouds only.
bool finite_points(const pcl::PointCloud<pcl::PointXYZ>& cloud, pcl::PointCloud<pcl::PointXYZ>::VectorType* out)
{
    if (cloud.is_dense) return false;
    out->clear();
    out->reserve(cloud.size());
    for (const auto& p : cloud.points)
        if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z)) out->push_back(p);
    return out->size() != cloud.size();
}

Eigen::Affine3d to_affine(const double* T)
{
    Eigen::Affine3d A;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) A(r, c) = T[4 * r + c];
    return A;
}

}  // namespace

ProbPointCloudRegistration::ProbPointCloudRegistration(pcl::PointCloud<pcl::PointXYZ>::Ptr source_cloud,
                                                       pcl::PointCloud<pcl::PointXYZ>::Ptr target_cloud,
                                                       ProbPointCloudRegistrationParams parameters)
    : parameters_(parameters),
      target_cloud_(target_cloud),
      ground_truth_(false),
      mse_ground_truth_(0),
      mse_prev_it_(0),
      reported_iterations_(0),
      output_stream_(parameters.verbose),
      handle_(nullptr)
{
    source_cloud_ = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>(*source_cloud);  // the caller's source is never touched
    init();
}

ProbPointCloudRegistration::ProbPointCloudRegistration(pcl::PointCloud<pcl::PointXYZ>::Ptr source_cloud,
                                                       pcl::PointCloud<pcl::PointXYZ>::Ptr target_cloud,
                                                       ProbPointCloudRegistrationParams parameters,
                                                       pcl::PointCloud<pcl::PointXYZ>::Ptr ground_truth_cloud)
    : ProbPointCloudRegistration(source_cloud, target_cloud, parameters)
{
    ground_truth_cloud_ = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>(*ground_truth_cloud);
    ground_truth_ = true;
    mse_ground_truth_ = calculateMSE(source_cloud_, ground_truth_cloud_);
    output_stream_ << "Initial MSE w.r.t. ground truth: " << mse_ground_truth_ << "\n";
}

ProbPointCloudRegistration::~ProbPointCloudRegistration()
{
    if (handle_) ppcr_destroy(handle_);
}

void ProbPointCloudRegistration::init()
{
    static_assert(sizeof(pcl::PointXYZ) == 16, "clouds are handed to the C ABI as 16-byte records");
    if (parameters_.source_filter_size > 0)
        output_stream_ << "Filtering source point cloud with leaf of size " << parameters_.source_filter_size << "\n";
    if (parameters_.target_filter_size > 0)
        output_stream_ << "Filtering target point cloud with leaf of size " << parameters_.target_filter_size << "\n";
    const ppcr_params cp = to_c(parameters_);
    const float* src = source_cloud_->empty() ? nullptr : reinterpret_cast<const float*>(source_cloud_->points.data());
    const float* tgt = target_cloud_->empty() ? nullptr : reinterpret_cast<const float*>(target_cloud_->points.data());
    int64_t n_src = static_cast<int64_t>(source_cloud_->size()), n_tgt = static_cast<int64_t>(target_cloud_->size());
    // Clouds flagged !is_dense (e.g. organised depth-sensor PCDs): the target's non-finite points never reach the search
    // structure, and a voxel-filtered source loses them in the filter.  An unfiltered source keeps them, as in the reference;
    // such a point finds no neighbour (every comparison with NaN fails) and contributes nothing.
    pcl::PointCloud<pcl::PointXYZ>::VectorType finite_src, finite_tgt;
    if (finite_points(*target_cloud_, &finite_tgt)) {
        tgt = finite_tgt.empty() ? nullptr : reinterpret_cast<const float*>(finite_tgt.data());
        n_tgt = static_cast<int64_t>(finite_tgt.size());
    }
    if (parameters_.source_filter_size > 0 && finite_points(*source_cloud_, &finite_src)) {
        src = finite_src.empty() ? nullptr : reinterpret_cast<const float*>(finite_src.data());
        n_src = static_cast<int64_t>(finite_src.size());
    }
    check(ppcr_create(src, n_src, tgt, n_tgt, &cp, &handle_), "ppcr_create");
    if (parameters_.target_filter_size > 0) {
        // the reference runs pcl::VoxelGrid on the caller's target cloud in place (registration.cc:34-41)
        int64_t n = 0;
        check(ppcr_filtered_target(handle_, nullptr, &n), "ppcr_filtered_target");
        pcl::PointCloud<pcl::PointXYZ>::VectorType filtered(static_cast<std::size_t>(n));
        if (n > 0) check(ppcr_filtered_target(handle_, reinterpret_cast<float*>(filtered.data()), &n), "ppcr_filtered_target");
        for (auto& p : filtered) p.data[3] = 1.0f;
        target_cloud_->points.swap(filtered);
        target_cloud_->width = static_cast<std::uint32_t>(target_cloud_->points.size());
        target_cloud_->height = 1;
        target_cloud_->is_dense = true;
    }
    if (parameters_.summary) {
        prev_source_cloud_ = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>(*source_cloud_);
        report_ << "iter, n_success_steps, initial_cost, final_cost, tx, ty, tz, roll, pitch, yaw, mse_prev_iter, mse_gtruth"
                << std::endl;
    }
}

void ProbPointCloudRegistration::align()
{
    check(ppcr_align(handle_), "ppcr_align");
    int32_t n = 0;
    check(ppcr_history(handle_, nullptr, &n), "ppcr_history");
    if (n > reported_iterations_) {
        std::vector<double> poses(static_cast<std::size_t>(n) * 16);
        std::vector<ppcr_iter_stats> stats(static_cast<std::size_t>(n));
        int32_t cap = n;
        check(ppcr_history(handle_, poses.data(), &cap), "ppcr_history");
        cap = n;
        check(ppcr_iteration_stats(handle_, stats.data(), &cap), "ppcr_iteration_stats");
        const bool replay = ground_truth_ || parameters_.summary;
        // The per-iteration diagnostics (:110-122): the full-resolution source copy follows the same increments as the
        // registered cloud, and its mean distance to the ground truth / to its previous position is taken after each.  One
        // device pass per iteration over the cloud (ppcr_replay_metrics) instead of a host loop per iteration.
        std::vector<double> mse_gt(static_cast<std::size_t>(n)), mse_prev(static_cast<std::size_t>(n));
        if (replay && !source_cloud_->empty()) {
            const int first = reported_iterations_;
            // (a ground truth of another size: let calculateMSE fail the way the reference's does, on .at())
            if (ground_truth_ && ground_truth_cloud_->size() != source_cloud_->size()) (void)calculateMSE(source_cloud_, ground_truth_cloud_);
            check(ppcr_replay_metrics(handle_, reinterpret_cast<float*>(source_cloud_->points.data()),
                                      ground_truth_ ? reinterpret_cast<const float*>(ground_truth_cloud_->points.data()) : nullptr,
                                      static_cast<int64_t>(source_cloud_->size()), first, n - first, mse_gt.data() + first,
                                      mse_prev.data() + first),
                  "ppcr_replay_metrics");
            if (parameters_.summary) *prev_source_cloud_ = *source_cloud_;
        }
        for (int it = reported_iterations_; it < n; ++it) {
            const Eigen::Affine3d current_trans = to_affine(&poses[16 * static_cast<std::size_t>(it)]);
            transformation_history_.push_back(current_trans);
            output_stream_ << "Outer iteration " << it << ": " << stats[it].n_correspondences << " correspondences, "
                           << stats[it].lm_iterations << " LM iterations (" << stats[it].num_successful_steps
                           << " successful), cost " << stats[it].initial_cost << " -> " << stats[it].final_cost << "\n";
            if (ground_truth_) {
                mse_ground_truth_ = mse_gt[static_cast<std::size_t>(it)];
                output_stream_ << "MSE w.r.t. ground truth: " << mse_ground_truth_ << "\n";
            }
            if (parameters_.summary) {
                mse_prev_it_ = mse_prev[static_cast<std::size_t>(it)];
                const Eigen::Vector3d rpy = current_trans.rotation().eulerAngles(0, 1, 2);
                report_ << it << ", " << stats[it].num_successful_steps << ", " << stats[it].initial_cost << ", "
                        << stats[it].final_cost << ", " << current_trans.translation().x() << ", "
                        << current_trans.translation().y() << ", " << current_trans.translation().z() << ", "
                        << pcl::rad2deg(rpy(0, 0)) << ", " << pcl::rad2deg(rpy(1, 0)) << ", " << pcl::rad2deg(rpy(2, 0))
                        << ", " << mse_prev_it_ << ", " << mse_ground_truth_ << std::endl;
            }
        }
        reported_iterations_ = n;
    }
    if (ground_truth_) {
        mse_ground_truth_ = calculateMSE(source_cloud_, ground_truth_cloud_);
        std::cout << "MSE w.r.t. ground truth: " << mse_ground_truth_ << std::endl;
    }
}

bool ProbPointCloudRegistration::hasConverged()
{
    int32_t done = 0;
    check(ppcr_has_converged(handle_, &done), "ppcr_has_converged");
    return done != 0;
}

}  // namespace prob_point_cloud_registration
