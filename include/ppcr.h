/*
 * ppcr.h -- C ABI of libppcr_cuda.so, the B200 (sm_100a) implementation of the per-iteration hot path of
 * iralabdisco/probabilistic_point_clouds_registration.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The reference is a single C++
 * process with no FFI of its own, so each entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  The reference-facing C++ class
 * (include/prob_point_cloud_registration/prob_point_cloud_registration.h) is a thin wrapper over these calls;
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions: every call returns ppcr_status (0 = ok); no exceptions cross the ABI; ppcr_last_error() gives
 * the message of the calling thread's last failure.  Clouds are arrays of 16-byte records (x, y, z, pad) --
 * the memory layout of pcl::PointXYZ -- so a pcl::PointCloud's points.data() can be passed as is.
 * A handle owns its device memory and runs on one CUDA stream; use one handle per host thread.
 * There is no CPU fallback: every entry point fails with PPCR_ERR_NO_DEVICE when no sm_100 GPU is usable.
 */
#ifndef PPCR_H
#define PPCR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t ppcr_status;
enum {
    PPCR_OK = 0,
    PPCR_ERR_INVALID = 1,     /* bad argument (null pointer, negative size, non-finite parameter) */
    PPCR_ERR_NO_DEVICE = 2,   /* no usable CUDA device / wrong architecture */
    PPCR_ERR_CUDA = 3,        /* a CUDA runtime call failed */
    PPCR_ERR_UNSUPPORTED = 4, /* parameter outside the implemented range (see ppcr_create) */
    PPCR_ERR_TIMEOUT = 5,     /* a peer rank did not arrive (sharded mode) */
    PPCR_ERR_SMALL_BUFFER = 6 /* caller's output buffer too small; required size written back */
};

/* Plain-C mirror of ProbPointCloudRegistrationParams
 * (include/prob_point_cloud_registration/prob_point_cloud_registration_params.hpp:5-18), field for field. */
typedef struct ppcr_params {
    int32_t max_neighbours;       /* :6  default 20 */
    int32_t n_iter;               /* :9  default 1000 */
    double dof;                   /* :7  default 5; +inf = Gaussian model (CLI -u) */
    double radius;                /* :8  default 1 (the CLI's default is 3) */
    double cost_drop_thresh;      /* :10 default 0.01 */
    double n_cost_drop_it;        /* :11 default 5 (a double in the reference) */
    int32_t verbose;              /* :12 */
    int32_t summary;              /* :13 */
    double initial_rotation[4];   /* :14 w,x,y,z; start of EVERY incremental solve (iteration.hpp:31-34) */
    double initial_translation[3];/* :15 */
    double source_filter_size;    /* :16 voxel leaf, 0 = off */
    double target_filter_size;    /* :17 */
} ppcr_params;

/* Engine knobs with no reference counterpart.  Zero-initialise for defaults. */
typedef struct ppcr_options {
    int32_t device;            /* CUDA device ordinal */
    int32_t input_on_device;   /* 1: src/tgt pointers passed to ppcr_create_ex are device pointers */
    int32_t driver;            /* 0 auto, 1 host-stepped ticks, 2 CUDA-graph WHILE loop (no host round trip) */
    int32_t ticks_per_sync;    /* host-stepped driver: ticks enqueued between flag read-backs (default 4) */
    double function_tolerance; /* inner LM tolerance; 0 = the reference's 10e-6 (src/..registration.cc:97) */
    int32_t leaf_capacity;     /* octree nodes holding more target points than this are split; 0 = default (32) */
    int32_t exact_weights;     /* 0 (default): float32 row arithmetic for the weights (|dw/w| ~ 1e-7, the bar is 1e-5),
                                  float64 sums across rows; 1: float64 log1p/exp per correspondence */
    void* stream;              /* cudaStream_t to run on; NULL = the handle creates its own */
    int32_t record_stage_times;/* 1: bracket kernels with CUDA events (host-stepped driver only) */
    int32_t reserved[7];
} ppcr_options;

/* One row per outer iteration: what ceres::Solver::Summary feeds into
 * src/prob_point_cloud_registration.cc:119-129 plus the association size. */
typedef struct ppcr_iter_stats {
    double initial_cost;
    double final_cost;
    double cost_drop;
    int64_t n_correspondences;
    int32_t lm_iterations;
    int32_t num_successful_steps;
} ppcr_iter_stats;

typedef struct ppcr_stage_times {   /* accumulated milliseconds + launch counts since ppcr_create */
    float grid_build_ms, search_ms, eval_ms, controller_ms, transform_ms, voxel_ms;
    int32_t search_launches, eval_launches, controller_launches, transform_launches, total_launches;
    int32_t ticks;
    int32_t exchanges;        /* sharded pairs: moment exchanges (= evaluations) so far */
    float exchange_wait_ms;   /* sharded pairs: time the controller block spent sending its moments and waiting for the peers' */
} ppcr_stage_times;

typedef struct ppcr_handle ppcr_handle;

const char* ppcr_last_error(void);
const char* ppcr_version(void);
void ppcr_default_params(ppcr_params* p);   /* struct defaults, params.hpp:6-17 */
void ppcr_default_options(ppcr_options* o);

/* Replaces the ProbPointCloudRegistration constructor (src/prob_point_cloud_registration.cc:15-49): copies the
 * source, voxel-filters source and target when the leaf sizes are > 0, builds the target octree.
 * max_neighbours <= 0 (pcl: every target within the radius) and > 128 run with rows of 128; ppcr_align then returns
 * PPCR_ERR_UNSUPPORTED if a source point has 128 or more targets within the radius (never a truncated row). */
ppcr_status ppcr_create(const float* src_xyzw, int64_t n_src, const float* tgt_xyzw, int64_t n_tgt,
                        const ppcr_params* params, ppcr_handle** out);
ppcr_status ppcr_create_ex(const float* src_xyzw, int64_t n_src, const float* tgt_xyzw, int64_t n_tgt,
                           const ppcr_params* params, const ppcr_options* options, ppcr_handle** out);
void ppcr_destroy(ppcr_handle* h);

/* ProbPointCloudRegistration::align(), src/prob_point_cloud_registration.cc:63-136: the whole outer loop
 * (search, weights, LM solve, compose, move the cloud, convergence test) runs on the device. */
ppcr_status ppcr_align(ppcr_handle* h);

/* ProbPointCloudRegistration::hasConverged(), :138-158, same mutating semantics.  *out = 0/1. */
ppcr_status ppcr_has_converged(ppcr_handle* h, int32_t* out);

/* transformation_history() / transformation(), prob_point_cloud_registration.h:32-40.
 * T receives up to *n_inout row-major 4x4 doubles; *n_inout returns the number of outer iterations run. */
ppcr_status ppcr_history(ppcr_handle* h, double* T4x4_rowmajor, int32_t* n_inout);
/* The increment of every outer iteration (registration.transformation() at :101-112, the transform both source
 * clouds are moved by), same layout as ppcr_history.  Lets a host replay the per-iteration diagnostics of the
 * reference (MSE w.r.t. ground truth / previous iteration, :114-129) without a device round trip per iteration. */
ppcr_status ppcr_increment_history(ppcr_handle* h, double* T4x4_rowmajor, int32_t* n_inout);
ppcr_status ppcr_iteration_stats(ppcr_handle* h, ppcr_iter_stats* out, int32_t* n_inout);

/* Clouds as the reference holds them after the call: the (voxel-filtered, moved) source and the filtered target
 * (the reference filters the caller's target IN PLACE, :34-41; the C++ wrapper writes this back). */
ppcr_status ppcr_filtered_source(ppcr_handle* h, float* out_xyzw, int64_t* n_inout);
ppcr_status ppcr_filtered_target(ppcr_handle* h, float* out_xyzw, int64_t* n_inout);

/* Current data association of the handle (the CSR pattern of :69-83), for parity tests:
 * idx is [n_src][max_neighbours] in source order, each row sorted by target index (the column order of the
 * reference's CSR after setFromTriplets, :82-83), padded with -1; count[n_src]. */
ppcr_status ppcr_association(ppcr_handle* h, int32_t* idx, int32_t* count, int64_t n_src, int32_t max_neighbours);

ppcr_status ppcr_get_stage_times(ppcr_handle* h, ppcr_stage_times* out);

/* Re-runs one kernel of the handle's current state `reps` times between two CUDA events on the handle's stream
 * and returns the average milliseconds and the algorithmic bytes of one launch.  which: 0 search from scratch
 * (pruning bound = the radius), 1 weights + normal equations + controller, 2 cloud move, 3 target tree build,
 * 4 search fused with an (identity) cloud move and warm-started from the previous search's m-th distances. */
ppcr_status ppcr_time_kernel(ppcr_handle* h, int32_t which, int32_t reps, int32_t flush_l2, float* avg_ms,
                             double* algorithmic_bytes);

/* ---- stage-level entry points (parity tests, ncu) ------------------------------------------------------ */

/* pcl::VoxelGrid<PointXYZ> as used at :24-41.  *n_out receives the filtered size; returns PPCR_OK with
 * *n_out = n and out = in when PCL would refuse the leaf size (index overflow). out must hold n points. */
ppcr_status ppcr_voxel_filter(const float* xyzw, int64_t n, double leaf, float* out_xyzw, int64_t* n_out);

/* Timing of the same filter (bench.py's roofline entry): `reps` runs between two CUDA events after one warm-up run; the
 * cloud is a host pointer unless options->input_on_device.  *n_out = filtered size (-1: PCL would refuse the leaf). */
ppcr_status ppcr_time_voxel_filter(const float* xyzw, int64_t n, double leaf, const ppcr_options* options, int32_t reps,
                                   float* avg_ms, double* algorithmic_bytes, int64_t* n_out);

/* The radius search loop at :72-81 (pcl::KdTreeFLANN::radiusSearch semantics).
 * out_idx/out_d2: [n_src][max_nn], rows sorted ascending by (d2, index); out_count[n_src].
 * leaf_capacity 0 = default; the result does not depend on it. */
ppcr_status ppcr_radius_search(const float* src_xyzw, int64_t n_src, const float* tgt_xyzw, int64_t n_tgt,
                               double radius, int32_t max_nn, int32_t leaf_capacity, int32_t* out_idx, float* out_d2,
                               int32_t* out_count);

/* WeightUpdaterCallback + one Jacobian/cost evaluation (weight_updater_callback.hpp:36-63,
 * probabilistic_weights.hpp:48-105, error_term.hpp:21-37) for an explicit association given as
 * idx[n_src][max_nn] / count[n_src].  pose_w = pose at which the weights are refreshed, pose_e = pose at which
 * residuals are evaluated; both (w,x,y,z,tx,ty,tz).  Outputs: weights[n_src][max_nn] (may be NULL),
 * normal_eq[36] = upper triangle of J^T W J (28, row-major), J^T W r (7), cost (1).
 * dimension = the `dimension` argument of ProbabilisticWeights (3 in production, iteration.hpp:17,29; the
 * reference's unit test uses 1, test/ProbabilisticWeightsTest.cc:39,55). */
ppcr_status ppcr_weights_normal_eq(const float* src_xyzw, int64_t n_src, const float* tgt_xyzw, int64_t n_tgt,
                                   const int32_t* idx, const int32_t* count, int32_t max_nn, double dof,
                                   int32_t dimension, const double* pose_w, const double* pose_e,
                                   int32_t fast_weights, double* weights, double* normal_eq);

/* ProbPointCloudRegistrationIteration ctor + solve + transformation() (iteration.hpp:24-67) for an explicit
 * association: the unit-test entry of the reference (test/PointCloudRegistrationTest.cc:49-60).
 * out_pose = (w,x,y,z,tx,ty,tz) as Ceres leaves it; out_T = row-major 4x4. */
ppcr_status ppcr_iteration_solve(const float* src_xyzw, int64_t n_src, const float* tgt_xyzw, int64_t n_tgt,
                                 const int32_t* idx, const int32_t* count, int32_t max_nn,
                                 const ppcr_params* params, const ppcr_options* options /* may be NULL */,
                                 double function_tolerance, double* out_pose, double* out_T, ppcr_iter_stats* stats);

/* The reference's per-iteration diagnostics on the DEVICE (src/prob_point_cloud_registration.cc:110-122,132-135 with
 * calculateMSE, include/.../utilities.hpp:16-26): replays the increments of outer iterations [first, first + count) of a
 * finished ppcr_align on a full-resolution cloud -- x <- float(dT x) in place, like :110 -- and returns per iteration the mean
 * Euclidean distance of the moved cloud to the ground truth (mse_gt, "MSE w.r.t. ground truth"; gt_xyzw may be NULL) and to
 * its previous position (mse_prev, the report's mse_prev_iter).  cloud_xyzw (n points, host) comes back moved by all of them.
 * One pass over the cloud per iteration, sums in a fixed order. */
ppcr_status ppcr_replay_metrics(ppcr_handle* h, float* cloud_xyzw, const float* gt_xyzw, int64_t n, int32_t first, int32_t count,
                                double* mse_gt, double* mse_prev);

/* pcl::transformPointCloud(cloud, cloud, Affine3d) at :110-112: double math, float store, in place. */
ppcr_status ppcr_transform(float* xyzw, int64_t n, const double* T4x4_rowmajor);
/* the same on options->device / options->stream, in place on a device-resident cloud when options->input_on_device */
ppcr_status ppcr_transform_ex(float* xyzw, int64_t n, const double* T4x4_rowmajor, const ppcr_options* options);

/* The closest-point metric helpers of include/prob_point_cloud_registration/utilities.hpp:28-234 in one call.  Each of them
 * builds a kd-tree on cloud2, takes nearestKSearch(k = 1) of every cloud1 point -- a SQUARED distance, float -- and reduces
 * that vector; here: one octree build, one 1-NN search, one radix sort, one reduction pass.  The reference's "median" is its
 * own index rule on the sorted vector (element (n+1)/2 for odd n, mean of n/2 and n/2+1 for even n; NaN where that reads out
 * of bounds, which is undefined behaviour in the reference); the robust variants keep the entries within
 * [median / f, median * f] (f = 3, or `factor`) and return DBL_MAX when fewer than 10 remain. */
typedef struct ppcr_closest_metrics {
    double average_closest_distance;           /* averageClosestDistance          :28-46  */
    double sum_squared_error;                  /* sumSquaredError                 :48-65  */
    double robust_sum_squared_error;           /* robustSumSquaredError           :67-101 */
    double robust_sum_squared_error_factor;    /* robustSumSquaredError(.., factor) :103-138 */
    double robust_averaged_sum_squared_error;  /* robustAveragedSumSquaredError   :140-175 */
    double median_closest_distance;            /* medianClosestDistance           :177-199 */
    double robust_median_closest_distance;     /* robustMedianClosestDistance     :201-234 */
    int64_t n_filtered, n_filtered_factor;     /* entries inside the two windows */
} ppcr_closest_metrics;
/* cloud1 / cloud2: host pointers unless options->input_on_device; options may be NULL.  out_d2 (may be NULL, host): the n1
 * squared distances in cloud1's order. */
ppcr_status ppcr_closest_point_metrics(const float* cloud1_xyzw, int64_t n1, const float* cloud2_xyzw, int64_t n2,
                                       double factor, const ppcr_options* options, ppcr_closest_metrics* out, float* out_d2);

/* ---- batch of independent pairs (new surface; no reference counterpart) -------------------------------- */

typedef struct ppcr_pair {
    const float* src_xyzw; int64_t n_src;
    const float* tgt_xyzw; int64_t n_tgt;
} ppcr_pair;

/* Registers n_pairs independent pairs with the same parameters on one device.  `slots` lanes (host threads, each with
 * its own stream and device-side iteration loop) pull pairs from a shared counter, so the set-up of one pair overlaps the
 * iterations of the others and small pairs fill the device together; slots <= 0 picks the default (6).
 * out_T: [n_pairs][16] final transforms; out_n_outer[n_pairs]; out_corr[n_pairs] = sum over outer iterations of
 * the association size.  Pair buffers are host pointers unless options->input_on_device. */
ppcr_status ppcr_align_batch(const ppcr_pair* pairs, int32_t n_pairs, const ppcr_params* params,
                             const ppcr_options* options, int32_t slots, double* out_T, int32_t* out_n_outer,
                             int64_t* out_corr);

/* The same over several GPUs of one process (device_ids[], n_dev): `slots` lanes on EVERY listed device, all
 * drawing pairs from one counter -- no exchange between devices, results land in the caller's arrays by pair index.  Pair buffers
 * must be host pointers when n_dev > 1 (options->device and options->stream are ignored then). */
ppcr_status ppcr_align_batch_devices(const ppcr_pair* pairs, int32_t n_pairs, const ppcr_params* params,
                                     const ppcr_options* options, const int32_t* device_ids, int32_t n_dev, int32_t slots,
                                     double* out_T, int32_t* out_n_outer, int64_t* out_corr);

/* ---- page-locked host memory for clouds (new surface) ------------------------------------------------------ */

/* A cloud that is read from a file straight into page-locked memory goes to the device at the full speed of the link and
 * asynchronously (ppcr_create's copy of a pageable buffer is staged by the driver).  ppcr_host_alloc returns NULL when no
 * CUDA device is usable or the allocation is refused -- fall back to malloc; ppcr_host_free returns 1 if `p` was one of
 * its blocks (and frees it), 0 otherwise (the caller frees it its own way).  The PCL stand-in's PointCloud uses them
 * (include/ppcr_compat/pcl/point_cloud.h); with the real PCL, cudaHostRegister the cloud's vector instead. */
void* ppcr_host_alloc(size_t bytes);
int32_t ppcr_host_free(void* p);

/* ---- one pair sharded over several GPUs (new surface) --------------------------------------------------- */

/* Every rank passes ITS slice of the source and the whole target.  The per-iteration exchange is a 32-double
 * all-reduce written straight into the peers' mailboxes over NVLink from inside the reduction kernel.
 * Call order on every rank: ppcr_create_ex -> ppcr_shard_export -> (exchange the tokens out of band,
 * e.g. torch.distributed.all_gather) -> ppcr_shard_connect -> ppcr_align.  The ranks may be processes (one per GPU: the
 * mailboxes are mapped over CUDA IPC) or host threads of one process (one per GPU: plain peer access).  The mailbox of a device
 * and the mappings of its peers are made once per process and re-used by every later sharded handle. */
#define PPCR_SHARD_TOKEN_BYTES 128
ppcr_status ppcr_shard_export(ppcr_handle* h, int32_t rank, int32_t world, uint8_t* token_out);
ppcr_status ppcr_shard_connect(ppcr_handle* h, const uint8_t* tokens /* [world][PPCR_SHARD_TOKEN_BYTES], rank order */);

/* The same for the GPUs of ONE process: the host source is dealt to the listed devices in runs of 8192 consecutive points (round
 * robin), every device gets the whole target, one host thread per device builds a handle, the mailboxes are connected by peer
 * access and the ranks run align() together.  out_T: [*n_inout][16] accumulated poses per outer iteration (ppcr_history),
 * *n_inout <- outer iterations run; out_corr (may be NULL) <- association size summed over them.  params->source_filter_size
 * must be 0 (a voxel filter would act on every share separately): PPCR_ERR_UNSUPPORTED. */
ppcr_status ppcr_align_sharded(const float* src_xyzw, int64_t n_src, const float* tgt_xyzw, int64_t n_tgt, const ppcr_params* params,
                               const ppcr_options* options, const int32_t* device_ids, int32_t n_dev, double* out_T,
                               int32_t* n_inout, int64_t* out_corr);

#ifdef __cplusplus
}
#endif
#endif
