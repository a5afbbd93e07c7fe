/*
 * voxgraph_b200.h — C-ABI of the B200-native voxgraph hot paths.
 *
 * Drop-in boundary for (paths relative to /root/reference/voxgraph/):
 *   b1  RegistrationCostFunction::Evaluate
 *         include/voxgraph/backend/constraint/cost_functions/registration_cost_function.h:47-48
 *         src/backend/constraint/cost_functions/registration_cost_function.cpp:58-298
 *   b2  PoseGraph  (include/voxgraph/backend/pose_graph.h:21-55, src/backend/pose_graph.cpp)
 *   b3  voxblox::TsdfIntegratorBase::integratePointCloud as driven by
 *         PointcloudIntegrator::integratePointcloud
 *         (src/frontend/measurement_processors/pointcloud_integrator.cpp:66-84)
 *
 * Conventions: plain pointers and sizes, caller owns every host buffer, the context
 * owns all device memory, every function returns an int status (no exceptions or
 * aborts cross the boundary; vgx_last_error() gives the message).  One context per
 * GPU; a context is not thread-safe.  There is NO CPU fallback: every entry point
 * needs a CUDA device and fails with VGX_ERR_CUDA without one.
 */
#ifndef VOXGRAPH_B200_H_
#define VOXGRAPH_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGX_OK 0
#define VGX_ZERO_WEIGHT 1      /* Evaluate() would return false (cpp:273) */
#define VGX_ERR_INVALID (-1)
#define VGX_ERR_CUDA (-2)
#define VGX_ERR_NOT_FOUND (-3)
#define VGX_ERR_NOMEM (-4)
#define VGX_ERR_NCCL (-5)
#define VGX_ERR_CAPACITY (-6)

typedef struct vgx_ctx vgx_ctx;

/* ------------------------------------------------------------------ context */
int vgx_device_count(void);
int vgx_ctx_create(int device, vgx_ctx** out);
void vgx_ctx_destroy(vgx_ctx* ctx);
const char* vgx_last_error(const vgx_ctx* ctx);
/* cudaStream_t all work of this context is enqueued on (for event timing). */
void* vgx_ctx_stream(vgx_ctx* ctx);
int vgx_ctx_synchronize(vgx_ctx* ctx);

/* Kernel accounting: when enabled the library brackets each of its kernels with
 * CUDA events on the context stream. which: 0 = registration reduce kernel,
 * 1 = registration emit kernel, 2 = TSDF integrate kernel, 3 = TSDF allocate kernel,
 * 4 = Cholesky/LM kernels, 5 = everything else, 6 = registration pose set-up, 7 = per-constraint sums,
 * 8 = assembly of the normal equations (+ the peer exchange when it is fused into it). */
int vgx_profile_enable(vgx_ctx* ctx, int on);
int vgx_profile_reset(vgx_ctx* ctx);
int vgx_profile_get(vgx_ctx* ctx, int which, double* total_ms, int64_t* launches);
/* total kernels launched by the library since reset (counted even when disabled) */
int64_t vgx_launch_count(vgx_ctx* ctx);

/* ------------------------------------------------------------------ submaps */
/* Replaces voxblox::Layer<TsdfVoxel|EsdfVoxel> + Block + block hash: dense 16^3 bricks
 * in HBM behind a GPU open-addressing hash keyed on the voxblox block index.
 * A voxel is (distance, weight); observed <=> weight > 1e-6 (upload an ESDF layer with
 * weight = observed ? 1 : 0).  Voxel order inside a block: x + vps*(y + vps*z). */

/* Upload a finished layer (replaces any previous content of submap_id). */
int vgx_submap_upload(vgx_ctx* ctx, uint32_t submap_id, float voxel_size, int voxels_per_side,
                      int n_blocks, const int32_t* block_idx /* n x 3 */,
                      const float* distance /* n x vps^3 */,
                      const float* weight /* n x vps^3 */);
/* Create an empty (active) submap that vgx_tsdf_integrate fills. */
int vgx_submap_create(vgx_ctx* ctx, uint32_t submap_id, float voxel_size, int voxels_per_side,
                      int capacity_blocks);
/* VoxgraphSubmap::finishSubmap (voxgraph_submap.cpp:84-107), device part: freezes the
 * layer and builds the registration view of the bricks. */
int vgx_submap_finish(vgx_ctx* ctx, uint32_t submap_id);
/* The rest of finishSubmap (voxgraph_submap.cpp:84-107) on the device, so that a submap integrated
 * on the GPU never leaves it between HP1 and HP2:
 *   findRelevantVoxelIndices (cpp:144-201)  -> registration points of type VGX_POINTS_VOXELS
 *   findIsosurfaceVertices   (cpp:203-243)  -> registration points of type VGX_POINTS_ISOSURFACE
 *                                              (+ the isosurface block list overlapsWith uses)
 *   getSubmapFrameSurfaceObb (cpp:280-324)
 * Order of the points: blocks in allocation order, voxels in linear order, cube edges 0..11 (the
 * reference iterates a hash map; see oracle/vg_oracle.h vgo_find_isosurface_vertices). */
typedef struct vgx_registration_filter {   /* VoxgraphSubmap::Config::RegistrationFilter (h:26-30) */
  double min_voxel_weight;                 /* 1 */
  double max_voxel_distance;               /* 0.3 */
  int use_esdf_distance;                   /* reference default true: relevant voxels carry the ESDF
                                              distance (cpp:185-189; needs vgx_submap_generate_esdf);
                                              vgx_registration_filter_default sets 0 = TSDF distance */
} vgx_registration_filter;
void vgx_registration_filter_default(vgx_registration_filter* f);
/* On a finished submap (uploaded or integrated). filter may be NULL (defaults). */
int vgx_submap_extract_points(vgx_ctx* ctx, uint32_t submap_id, const vgx_registration_filter* filter);
/* vgx_submap_finish + vgx_submap_extract_points. */
int vgx_submap_finish_ex(vgx_ctx* ctx, uint32_t submap_id, const vgx_registration_filter* filter);
int vgx_submap_num_points(vgx_ctx* ctx, uint32_t submap_id, int point_type, int* n);
/* Copies the registration points back (inspection / parity); any output may be NULL. */
int vgx_submap_download_points(vgx_ctx* ctx, uint32_t submap_id, int point_type, int max_n,
                               float* xyz /* n x 3 */, float* distance, float* weight, int* n);
/* Submap-frame surface OBB; VGX_ZERO_WEIGHT when no voxel qualified (box stays +-inf). */
int vgx_submap_surface_obb(vgx_ctx* ctx, uint32_t submap_id, float obb_min[3], float obb_max[3]);
/* TsdfEsdfSubmap::generateEsdf (finishSubmap, voxgraph_submap.cpp:86) on the device: ESDF bricks
 * + their registration view from the resident TSDF bricks of a finished submap, so that the
 * reference's default registration branch (use_esdf_distance = true,
 * registration_cost_function.h:35, cpp:133-140) has a device producer.  Restates
 * voxblox::EsdfIntegrator's batch update as its fixed point: observed <=> TSDF weight >=
 * min_weight; |tsdf| < min_distance_m is copied (fixed); elsewhere the quasi-Euclidean distance
 * propagated through the 26-neighbourhood of same-sign observed voxels, clipped at max_distance_m
 * (default_distance_m where nothing reaches).  sweeps (may be NULL): relaxation sweeps run. */
typedef struct vgx_esdf_config {      /* voxblox::EsdfIntegrator::Config subset */
  float max_distance_m;               /* 2.0 */
  float default_distance_m;           /* 2.0 */
  float min_distance_m;               /* 0.2 */
  float min_weight;                   /* 1e-6 */
} vgx_esdf_config;
void vgx_esdf_config_default(vgx_esdf_config* cfg);
int vgx_submap_generate_esdf(vgx_ctx* ctx, uint32_t submap_id, const vgx_esdf_config* cfg, int* sweeps);
/* ESDF bricks in block (allocation) order: distance and observed (1 / 0) per voxel. */
int vgx_submap_download_esdf(vgx_ctx* ctx, uint32_t submap_id, int max_blocks, float* distance,
                             float* observed, int* n_blocks);

/* PoseGraphInterface::updateOverlappingSubmapList (pose_graph_interface.cpp:109-147) over
 * VoxgraphSubmap::overlapsWith (voxgraph_submap.cpp:245-278): mission-frame surface-AABB rejection,
 * then "any isosurface block centre of the first submap lands in an allocated block of the second",
 * for every i < j of the list.  poses: n x 7 = T_mission_submap [qw qx qy qz tx ty tz] (float, as
 * cblox stores them).  Needs vgx_submap_extract_points on every submap.  pairs: 2 ids per pair. */
int vgx_find_overlapping_pairs(vgx_ctx* ctx, int n, const uint32_t* submap_ids, const float* poses,
                               int max_pairs, uint32_t* pairs, int* n_pairs);
int vgx_submap_free(vgx_ctx* ctx, uint32_t submap_id);
int vgx_submap_block_count(vgx_ctx* ctx, uint32_t submap_id, int* n_blocks);
/* Layer geometry and state (any output may be NULL). */
int vgx_submap_info(vgx_ctx* ctx, uint32_t submap_id, float* voxel_size, int* voxels_per_side,
                    int* n_blocks, int* finished);
/* Copies blocks back in allocation order (parity checks / saving). */
int vgx_submap_download(vgx_ctx* ctx, uint32_t submap_id, int max_blocks, int32_t* block_idx,
                        float* distance, float* weight, int* n_blocks);

/* Device-to-device hand-over of a finished submap between two contexts (BASELINE configs[4]: one
 * GPU integrates, the other registers; the finished submap crosses NVLink once).  peek: read-only
 * device pointers of the resident bricks (block indices n x 3 int32, voxels n x vps^3 x
 * {distance, weight} float) - valid until the submap is modified or freed; the caller moves them
 * with its own transport (ncclSend / cudaMemcpyPeer).  upload_device: same as vgx_submap_upload but
 * from device memory of this context's GPU, interleaved (distance, weight) voxels. */
int vgx_submap_peek_device(vgx_ctx* ctx, uint32_t submap_id, const int32_t** d_block_idx,
                           const float** d_distance_weight, int* n_blocks);
int vgx_submap_upload_device(vgx_ctx* ctx, uint32_t submap_id, float voxel_size, int voxels_per_side,
                             int n_blocks, const int32_t* d_block_idx, const float* d_distance_weight);

/* Registration points of a submap = WeightedSampler<RegistrationPoint> items
 * (include/voxgraph/frontend/submap_collection/registration_point.h:6-12).
 * point_type: 0 = kVoxels, 1 = kIsosurfacePoints (voxgraph_submap.h RegistrationPointType). */
#define VGX_POINTS_VOXELS 0
#define VGX_POINTS_ISOSURFACE 1
int vgx_submap_upload_points(vgx_ctx* ctx, uint32_t submap_id, int point_type, int n,
                             const float* xyz /* n x 3 */, const float* distance,
                             const float* weight);

/* ------------------------------------------------------------------ b3: TSDF integration */
typedef struct vgx_tsdf_config {          /* voxblox::TsdfIntegratorBase::Config */
  float default_truncation_distance;      /* voxgraph_mapper.yaml:23  0.6 */
  float max_weight;                       /* 10000 */
  int voxel_carving_enabled;              /* 1 */
  float min_ray_length_m;                 /* 0.1 */
  float max_ray_length_m;                 /* yaml:24  16 */
  int use_const_weight;                   /* yaml:25  1 */
  int allow_clear;                        /* 1 */
  int use_weight_dropoff;                 /* yaml:26  1 */
  int use_sparsity_compensation_factor;   /* yaml:27  1 */
  float sparsity_compensation_factor;     /* yaml:28  20 */
  float start_voxel_subsampling_factor;   /* FastTsdfIntegrator, 2 */
  int max_consecutive_ray_collisions;     /* FastTsdfIntegrator, 2 */
  int mode;                               /* 0 simple (every ray, every voxel; the updates of each
                                             voxel are applied in ray order = bit-identical to the
                                             single-threaded reference), 1 fast (FastTsdfIntegrator
                                             scheduling, what voxgraph runs) */
  int deterministic;                      /* ignored (kept for layout compatibility): mode 0 is always
                                             ray ordered */
} vgx_tsdf_config;

typedef struct vgx_tsdf_stats {
  int64_t rays_valid;
  int64_t rays_cast;
  int64_t voxel_updates;
  int64_t blocks_allocated;
  int64_t saturated_batches;   /* mode 0: 32-update batches of sensor-adjacent voxels collapsed in
                                  closed form (distance pinned at +trunc, exact integer weight sum) */
} vgx_tsdf_stats;

void vgx_tsdf_config_default(vgx_tsdf_config* cfg);
/* integratePointCloud(T_G_C, points_C, colors): T_G_C = [qw qx qy qz tx ty tz];
 * rgba may be NULL. Host buffers; copies are part of the call. */
int vgx_tsdf_integrate(vgx_ctx* ctx, uint32_t submap_id, const float T_G_C[7], int n,
                       const float* points_C /* n x 3 */, const uint8_t* rgba /* n x 4 or NULL */,
                       const vgx_tsdf_config* cfg, vgx_tsdf_stats* stats /* may be NULL */);

/* ------------------------------------------------------------------ b1: cost function */
typedef struct vgx_reg_config {           /* RegistrationCostFunction::Config (h:17-41) */
  int registration_point_type;            /* VGX_POINTS_* ; default isosurface */
  double no_correspondence_cost;          /* 0 */
  float sampling_ratio;                   /* -1: every registration point, in order (cpp:115-117).
                                             Otherwise (voxgraph_mapper.yaml:34 uses 0.05):
                                             num_residuals = int(ratio * K) points are drawn with
                                             probability proportional to their weight from the
                                             reference submap's WeightedSampler (a default-seeded
                                             std::mt19937 per submap and point type,
                                             weighted_sampler_inl.h:19-28) and their weight is
                                             forced to 1 (cpp:118-122). */
  int use_esdf_distance;                  /* h:35 (reference default true): interpolate the reading
                                             submap's ESDF (vgx_submap_generate_esdf) instead of its
                                             TSDF (cpp:133-153).  vgx_reg_config_default sets 0. */
} vgx_reg_config;
void vgx_reg_config_default(vgx_reg_config* cfg);

/* Number of residuals of the (reference -> reading) cost function (cpp:45-55). */
int vgx_reg_num_residuals(vgx_ctx* ctx, uint32_t reference_submap_id, const vgx_reg_config* cfg,
                          int* num_residuals);
/* Evaluate(parameters, residuals, jacobians): Ceres layout, residuals[K] and two K x 4
 * row-major Jacobian blocks (either may be NULL). Returns VGX_OK, or VGX_ZERO_WEIGHT
 * where Evaluate returns false.  In sampling mode every call draws a fresh sample from the
 * reference submap's generator, exactly as every Evaluate does in the reference. */
int vgx_reg_eval_emit(vgx_ctx* ctx, uint32_t reference_submap_id, uint32_t reading_submap_id,
                      const vgx_reg_config* cfg, const double reference_pose[4],
                      const double reading_pose[4], double* residuals, double* jac_reference,
                      double* jac_reading);

/* WeightedSampler::getRandomItem (weighted_sampler_inl.h:19-28) n times on the submap's
 * generator: writes the n drawn point indices and advances the generator (inspection / parity
 * tests; the same routine feeds the sampling mode of b1 and b2). */
int vgx_submap_draw_samples(vgx_ctx* ctx, uint32_t submap_id, int point_type, int n,
                            int32_t* indices);

/* ------------------------------------------------------------------ b2: pose graph */
/* addSubmapNode for every node (replaces the node set). xyzyaw: n x 4 [x,y,z,yaw]. */
int vgx_graph_set_nodes(vgx_ctx* ctx, int n, const uint32_t* submap_ids, const double* xyzyaw,
                        const uint8_t* constant);
int vgx_graph_set_poses(vgx_ctx* ctx, const double* xyzyaw);
int vgx_graph_get_poses(vgx_ctx* ctx, double* xyzyaw);
/* add{Relative,Absolute}PoseConstraint: m edges, observed [tx,ty,tz,yaw], sqrt-information
 * (lower LLT factor, row-major 4x4, constraint.cpp:8-14). */
int vgx_graph_set_relative_edges(vgx_ctx* ctx, int m, const uint32_t* ids_a, const uint32_t* ids_b,
                                 const double* t_obs_xyzyaw, const double* sqrt_info);
/* One entry per registration residual block (reference -> reading). The caller adds the
 * mirrored block for isosurface points as PoseGraph::addRegistrationConstraint does
 * (pose_graph.cpp:63-71). Replaces the previous list (resetRegistrationConstraints). */
int vgx_graph_set_registration_constraints(vgx_ctx* ctx, int p, const uint32_t* reference_ids,
                                           const uint32_t* reading_ids, const vgx_reg_config* cfg);
/* Same with one config per residual block (the reference keeps a Config per constraint,
 * registration_constraint.h:15-21). */
int vgx_graph_set_registration_constraints_v(vgx_ctx* ctx, int p, const uint32_t* reference_ids,
                                             const uint32_t* reading_ids,
                                             const vgx_reg_config* cfgs /* p entries */);
/* Sampling mode inside the fused evaluation: the reference re-draws on every Evaluate, i.e. the
 * cost Ceres compares between two LM iterations is stochastic.  The fused path draws ONCE per
 * constraint when the constraint list is (re)built - PoseGraphInterface resets and re-adds the
 * registration constraints before every optimize() (pose_graph_interface.cpp:149-175), so every
 * solve gets a fresh draw from the submap's generator - and keeps it for all evaluations of
 * that solve.  Every rank draws for every constraint, so the streams stay identical across
 * ranks.  get: the indices in use (n = num_residuals of that block); set: replace them
 * (parity tests against an oracle fed with the same list). constraint = index in the list. */
int vgx_graph_get_sample_indices(vgx_ctx* ctx, int constraint, int max_n, int32_t* indices, int* n);
int vgx_graph_set_sample_indices(vgx_ctx* ctx, int constraint, int n, const int32_t* indices);
int vgx_graph_num_registration_residuals(vgx_ctx* ctx, int64_t* local, int64_t* global);

/* Whole-problem evaluation at the current poses, fused on the device: cost = 1/2 sum r^2,
 * gradient = J^T r (4 per node), H = J^T J (dense 4N x 4N row-major). Outputs may be NULL.
 * With a communicator the result is the all-reduced global one. */
int vgx_graph_eval(vgx_ctx* ctx, int exclude_registration, double* cost, double* gradient,
                   double* H);
/* Enqueue only (no device->host copy, no sync): for device-side timing. */
int vgx_graph_eval_async(vgx_ctx* ctx, int exclude_registration);
/* Sum of squared residuals per registration constraint (pose_graph.cpp:194-207). */
int vgx_graph_registration_costs(vgx_ctx* ctx, double* per_constraint);

typedef struct vgx_solver_options {       /* ceres::Solver::Options subset, pose_graph.cpp:91-97 */
  int max_num_iterations;                 /* 50 */
  double parameter_tolerance;             /* 3e-3 */
  double function_tolerance;              /* 1e-6 */
  double gradient_tolerance;              /* 1e-10 */
  double initial_trust_region_radius;     /* 1e4 */
  double max_trust_region_radius;         /* 1e16 */
  double min_trust_region_radius;         /* 1e-32 */
  double min_relative_decrease;           /* 1e-3 */
  double min_lm_diagonal;                 /* 1e-6 */
  double max_lm_diagonal;                 /* 1e32 */
  double max_solver_time_s;               /* 4 */
  int jacobi_scaling;                     /* 1 */
  int exclude_registration;               /* PoseGraph::optimize(bool) */
} vgx_solver_options;

typedef struct vgx_solver_summary {       /* what voxgraph reads off ceres::Solver::Summary */
  int iterations;
  int num_successful_steps;
  int num_residual_evals;
  int termination;  /* 0 parameter tol, 1 function tol, 2 gradient tol, 3 max iterations,
                       4 max time, 5 min radius, 6 failure */
  double initial_cost, final_cost;
  double total_time_s;
} vgx_solver_summary;

void vgx_solver_options_default(vgx_solver_options* o);
/* PoseGraph::optimize: Levenberg-Marquardt with Ceres' default trust-region schedule,
 * device resident; optimised poses are written to xyzyaw_out (n x 4, may be NULL) and
 * stay the graph's current poses. */
int vgx_graph_solve(vgx_ctx* ctx, const vgx_solver_options* opts, double* xyzyaw_out,
                    vgx_solver_summary* summary);

/* PoseGraph::getEdgeCovarianceMap (pose_graph.cpp:117-163): the 4 x 4 covariance blocks
 * Cov(x_a, x_b) (row-major, rows = parameters of a) that ceres::Covariance reports for the requested
 * submap pairs: blocks of (J^T J)^-1 over the free nodes at the current poses; zero for a constant
 * node.  Fails (VGX_ERR_INVALID) where Covariance::Compute returns false (rank-deficient J). */
int vgx_graph_edge_covariances(vgx_ctx* ctx, int m, const uint32_t* ids_a, const uint32_t* ids_b,
                               double* covariances /* m x 16 */);

/* ------------------------------------------------------------------ multi-GPU */
/* One process per GPU. Registration constraints are sharded over the ranks of the
 * communicator (every rank passes the same full list and holds the submaps it needs);
 * the per-node / per-edge normal-equation blocks are summed with one ncclAllReduce per
 * evaluation. */
/* Host-only: the constraint -> rank partition the library uses. Constraints are ordered by
 * locality key (the reading submap id; NULL = input order; ties keep input order) and that
 * sequence is cut into nranks contiguous pieces of equal residual count, so a rank gathers from
 * few reading submaps (L2 locality) and loads differ by at most one constraint.
 * owner[i] in [0, nranks). */
int vgx_shard_constraints(int nranks, int n, const int32_t* num_residuals,
                          const uint32_t* locality_keys, int32_t* owner);
int vgx_comm_unique_id(uint8_t id[128]);
int vgx_comm_init(vgx_ctx* ctx, int nranks, int rank, const uint8_t id[128]);
int vgx_comm_destroy(vgx_ctx* ctx);
/* NVLink peer-memory exchange (single node, <= 8 ranks): replaces the NCCL all-reduce of the
 * packed normal equations by a one-shot all-gather-reduce over CUDA-IPC mapped peer buffers:
 * the assembly kernel PUSHES every element of the rank's partial, tagged with the evaluation's
 * epoch (one 16-byte store {lo32, tag, hi32, tag}), into its slot of every peer's exported
 * region while it is produced; each rank then spins on the tagged elements of its own region and
 * sums the partials in rank order (no load crosses NVLink, no fence or flag; bit-identical on all
 * ranks).  export: allocate 2 parities x 8 ranks x capacity_doubles x 16 bytes + get the 64-byte
 * IPC handle; import: map the peers (handles = nranks x 64 bytes, own entry ignored). Works with
 * or without vgx_comm_init; when both are set the peer path is used. */
/* Self-check of the sharded evaluation: while suspended (on = 1) the context behaves as a
 * single-rank context - it evaluates EVERY constraint locally and skips the exchange - so a rank
 * can compare the all-reduced result with the full problem computed on one GPU (all submaps are
 * replicated anyway).  Local to the calling rank; on = 0 restores the rank layout. */
int vgx_comm_suspend(vgx_ctx* ctx, int on);
int vgx_comm_p2p_export(vgx_ctx* ctx, uint64_t capacity_doubles, uint8_t handle[64]);
int vgx_comm_p2p_import(vgx_ctx* ctx, int nranks, int rank, const uint8_t* handles);

#ifdef __cplusplus
}
#endif
#endif /* VOXGRAPH_B200_H_ */
