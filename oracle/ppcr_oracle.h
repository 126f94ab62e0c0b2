/*
 * ppcr_oracle.h -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libppcr_cuda.so, the C++ class, the CLI)
 * never links, loads or calls anything in oracle/.
 *
 * Parity status ("pinned" = checked against the reference's own golden vectors):
 *   weights            PINNED   by test/ProbabilisticWeightsTest.cc:35-66 (G1, G2)
 *   inner solve        PINNED at the fixed point by test/PointCloudRegistrationTest.cc:30-116 (G3, G4)
 *   radius search      parity UNPINNED (PCL/FLANN absent; semantics restated from their published behaviour)
 *   voxel grid         parity UNPINNED (PCL absent)
 *   outer loop/align   parity UNPINNED (follows src/prob_point_cloud_registration.cc:63-158 line by line,
 *                      but Ceres' trust-region minimiser is restated, not linked)
 *
 * All citations are into /root/reference (read-only; nothing is read from it at run time).
 */
#ifndef PPCR_ORACLE_H
#define PPCR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Plain-C mirror of ProbPointCloudRegistrationParams
 * (include/prob_point_cloud_registration/prob_point_cloud_registration_params.hpp:5-18). */
typedef struct oracle_params {
    int32_t max_neighbours;
    int32_t n_iter;
    double dof;              /* +inf selects the Gaussian model (CLI -u, src/..._ex.cc:93-98) */
    double radius;
    double cost_drop_thresh;
    double n_cost_drop_it;   /* a double in the reference */
    int32_t verbose;
    int32_t summary;
    double initial_rotation[4];    /* w, x, y, z */
    double initial_translation[3];
    double source_filter_size;
    double target_filter_size;
} oracle_params;

/* Options of the restated ceres::Solve (src/prob_point_cloud_registration.cc:88-98). */
typedef struct oracle_solver_options {
    double function_tolerance;  /* reference: 10e-6 in align(), 10e-5 in the unit tests */
    int32_t max_num_iterations; /* reference: INT_MAX */
    int32_t inner_kind;         /* 0 = faithful (dual-number autodiff + dense Householder QR),
                                   1 = fast (analytic Jacobian rows + 7x7 normal equations, OpenMP) */
    int32_t num_threads;        /* OpenMP threads for kind 1 and for the grid search; <=0 = all */
} oracle_solver_options;

typedef struct oracle_solve_summary {
    double initial_cost;
    double final_cost;
    int32_t num_iterations;       /* LM iterations run, excluding iteration zero */
    int32_t num_successful_steps; /* includes iteration zero, like Ceres */
    int32_t termination;          /* 0 function tol, 1 parameter tol, 2 gradient tol, 3 min radius,
                                     4 max iterations, 5 no residuals, 6 too many invalid steps */
    int32_t num_nonmonotonic_steps; /* accepted steps that did not lower the minimum cost (use_nonmonotonic_steps) */
} oracle_solve_summary;

typedef struct oracle_iter_stats {
    double initial_cost;
    double final_cost;
    double cost_drop;
    int64_t n_correspondences;
    int32_t lm_iterations;
    int32_t num_successful_steps;
} oracle_iter_stats;

/* FLANN-semantics radius search (src/prob_point_cloud_registration.cc:72-81).
 * Points are 16-byte xyzw records like pcl::PointXYZ.  out_idx/out_d2 are [n_src][cap] (cap >= the
 * effective limit), rows sorted ascending by (d2, index).  use_grid!=0 runs the CPU uniform-grid
 * version (identical results, used for large clouds). Returns total correspondences. */
int64_t oracle_radius_search(const float* src_xyzw, int64_t n_src, const float* tgt_xyzw, int64_t n_tgt,
                             double radius, int32_t max_nn, int32_t cap, int32_t use_grid, int32_t num_threads,
                             int32_t* out_idx, float* out_d2, int32_t* out_count);

/* ProbabilisticWeights::updateWeights (probabilistic_weights.hpp:48-105) on a CSR pattern. */
void oracle_update_weights(int64_t n_rows, const int64_t* row_ptr, const double* squared_errors,
                           double dof, int32_t dimension, double* out_weights);

/* WeightUpdaterCallback::operator() (weight_updater_callback.hpp:36-63): residuals at (q,t) then weights. */
void oracle_callback_weights(const float* src_xyzw, const float* tgt_xyzw, int64_t n_rows,
                             const int64_t* row_ptr, const int32_t* col_idx, const double* rotation_wxyz,
                             const double* translation, double dof, double* out_sq_err, double* out_weights);

/* ProbPointCloudRegistrationIteration ctor + solve + transformation()
 * (prob_point_cloud_registration_iteration.hpp:24-67) with the restated Ceres minimiser.
 * out_rotation (w,x,y,z, NOT normalised, as Ceres leaves it), out_translation, out_T = 4x4 row-major. */
int32_t oracle_iteration_solve(const float* src_xyzw, int64_t n_src, const float* tgt_xyzw, int64_t n_tgt,
                               const int64_t* row_ptr, const int32_t* col_idx, const oracle_params* params,
                               const oracle_solver_options* opts, double* out_rotation, double* out_translation,
                               double* out_T, oracle_solve_summary* summary);

/* pcl::transformPointCloud(cloud, cloud, Affine3d) restated: double math, float store, in place. */
void oracle_transform(float* xyzw, int64_t n, const double* T_rowmajor4x4);

/* pcl::VoxelGrid<PointXYZ> (default settings) restated.  out must hold n points. Returns count,
 * or -1 when PCL would warn about index overflow and return the input unchanged (out = in). */
int64_t oracle_voxel_grid(const float* xyzw, int64_t n, double leaf, float* out_xyzw);

/* ProbPointCloudRegistration ctor + align() (src/prob_point_cloud_registration.cc:15-158).
 * src/tgt are copied.  history: up to max_hist 4x4 row-major matrices; stats: one per outer iteration.
 * out_filtered_source (n_src*4 floats, may be NULL) receives the moved, filtered source cloud;
 * *n_filtered_src / *n_filtered_tgt the sizes after voxel filtering.
 * Returns the number of outer iterations performed. */
int32_t oracle_align(const float* src_xyzw, int64_t n_src, const float* tgt_xyzw, int64_t n_tgt,
                     const oracle_params* params, const oracle_solver_options* opts, int32_t use_grid,
                     double* history, oracle_iter_stats* stats, int32_t max_hist,
                     float* out_filtered_source, int64_t* n_filtered_src, int64_t* n_filtered_tgt);

/* calculateMSE (utilities.hpp:16-26): mean Euclidean distance between same-sized clouds. */
double oracle_calculate_mse(const float* a_xyzw, const float* b_xyzw, int64_t n);

/* The closest-point metric helpers of utilities.hpp:28-234 in one call (1-NN squared distances of cloud1 in cloud2, brute
 * force).  out[9]: 0 averageClosestDistance, 1 sumSquaredError, 2 robustSumSquaredError, 3 robustSumSquaredError(factor),
 * 4 robustAveragedSumSquaredError, 5 medianClosestDistance, 6 robustMedianClosestDistance, 7 / 8 the number of distances
 * inside the [median / 3, median * 3] and [median / factor, median * factor] windows.  out_d2 (may be NULL): the n1
 * squared distances.  Returns -1 when a cloud is empty (the reference reads out of bounds there). */
int32_t oracle_closest_metrics(const float* c1_xyzw, int64_t n1, const float* c2_xyzw, int64_t n2, double factor,
                               double* out9, float* out_d2);

/* Upper triangle of J^T W J (28, row-major), J^T W r (7) and the cost (1) of the problem of iteration.hpp:24-50 evaluated at
 * x_e = (w,x,y,z,tx,ty,tz) with the weights refreshed at x_w (WeightUpdaterCallback), analytic Jacobian rows in float64. */
void oracle_normal_eq(const float* src_xyzw, const float* tgt_xyzw, int64_t n_rows, const int64_t* row_ptr,
                      const int32_t* col_idx, double dof, const double* x_w, const double* x_e, double* out36);

int32_t oracle_max_threads(void);

/* Accepted non-monotonic LM steps over every solve since the last reset (tests: proves a scenario exercises the
 * "callback sees the lowest-cost iterate" rule of update_state_every_iteration). */
int64_t oracle_nonmonotonic_steps(int32_t reset);

#ifdef __cplusplus
}
#endif
#endif
