// ProbPointCloudRegistrationParams -- the reference's parameter aggregate, field for field
// (reference: include/prob_point_cloud_registration/prob_point_cloud_registration_params.hpp:5-18).
#ifndef PROB_POINT_CLOUD_REGISTRATION_POINT_CLOUD_REGISTRATION_PARAMS_HPP
#define PROB_POINT_CLOUD_REGISTRATION_POINT_CLOUD_REGISTRATION_PARAMS_HPP

namespace prob_point_cloud_registration {
struct ProbPointCloudRegistrationParams {
    int max_neighbours = 20;        // cap on the neighbour set of a source point
    double dof = 5;                 // t-distribution degrees of freedom; +inf = Gaussian model
    double radius = 1;              // neighbourhood radius (the CLI's default is 3)
    int n_iter = 1000;              // cap on outer iterations
    double cost_drop_thresh = 0.01; // relative cost drop below which an outer iteration counts as "unuseful"
    double n_cost_drop_it = 5;      // unuseful iterations tolerated (a double in the reference as well)
    bool verbose = false;
    bool summary = false;           // collect the per-iteration report (CLI --dump)
    double initial_rotation[4] = {1, 0, 0, 0};  // w, x, y, z: start of every incremental solve
    double initial_translation[3] = {0, 0, 0};
    double source_filter_size = 0;  // voxel leaf, 0 = off
    double target_filter_size = 0;
};
}  // namespace prob_point_cloud_registration

#endif
