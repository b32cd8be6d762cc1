/* panovlm_b200 — C ABI of the B200-native correspondence-and-residual hot path of PanoVLM.
 *
 * The reference (3dv-casia/PanoVLM) has no plugin/FFI interface; the surface this library drops in behind is
 * Ceres' (SURVEY.md §8b):
 *   - ceres::EvaluationCallback::PrepareForEvaluation(evaluate_jacobians, new_point)  ->  pvb_blocks_evaluate()
 *   - ceres::CostFunction::Evaluate(parameters, residuals, jacobians) for the functors of
 *     base/CostFunction.h:567-1022 (one AutoDiffCostFunction<F,1,3,3,3,3> per correspondence, built by
 *     util/Optimization.cpp:329-441, 506-607)                                         ->  pvb_blocks_residuals()/
 *                                                                                          pvb_blocks_jacobians() rows
 *   - the association loops those builders call (lidar_mapping/LidarFeatureAssociate.cpp:442-476, 550-630,
 *     joint_optimization/CameraLidarLineAssociate.cpp:340-475)                        ->  pvb_frames_* / pvb_*_votes
 *   - the bulk SE(3) + equirectangular projection util/Visualization.h:408-441       ->  pvb_project_*
 *   - the solve the reduced system feeds (lidar_mapping/LidarOdometry.cpp:78-80)       ->  pvb_blocks_solve_lm(),
 *                                                                                          pvb_dense_*
 * Conventions: every function returns 0 on success and a negative code on error (no exceptions cross the ABI);
 * pvb_last_error() gives the message.  The caller owns every host array it passes; the context owns device and
 * pinned buffers; pointers returned by getters stay valid until the next evaluate call on the same context.
 * Pose blocks are 6 doubles (angle-axis aa_lw[3], t_lw[3]) = the parameter blocks of
 * lidar_mapping/LidarOdometry.cpp:23-33 (world -> sensor).  Rotation matrices are row-major.
 * There is no CPU fallback: every entry point needs a CUDA device (sm_100a build).
 */
#ifndef PANOVLM_B200_H_
#define PANOVLM_B200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pvb_ctx pvb_ctx;

enum { PVB_OK = 0, PVB_ERR_CUDA = -1, PVB_ERR_ARG = -2, PVB_ERR_STATE = -3, PVB_ERR_NOMEM = -4 };

/* residual types = the reference functors (base/CostFunction.h) */
enum {
  PVB_P2PLANE_METER = 0,      /* Point2Plane_Meter   :567-619   consts: p[0..2] plane[3..6] weight[7]                      */
  PVB_P2PLANE_ANGLE = 1,      /* Point2Plane_Angle   :630-729   consts: p[0..2] plane[3..6] (weight ignored, as upstream)   */
  PVB_P2LINE_METER = 2,       /* Point2Line_Meter    :769-829   consts: p[0..2] a[3..5] dir[6..8] weight[9]                */
  PVB_P2LINE_ANGLE = 3,       /* Point2Line_Angle    :836-934   consts: p[0..2] a[3..5] dir[6..8] (dir = (a-b)/|a-b|)       */
  PVB_PLANE2PLANE_GLOBAL = 4, /* Plane2Plane_Global  :350-425   consts: n[0..2] a[3..5] b[6..8] weight[9]                  */
  PVB_PLANE_IOU = 5,          /* PlaneIOUResidual    :433-507   consts: plane[0..3] mid_nei[4..6] mid_ref[7..9] angle[10] weight[11] */
  /* calibration-mode functors (CameraLidarOptimizer.cpp:58-63, LidarOdometry_test.cpp:127): one relative pose = the `ref` block,
     the `nei` block is ignored (zero Jacobian columns)                                                                          */
  PVB_PLANE2PLANE_RELATIVE = 6, /* Plane2Plane_Relative :294-346  consts as PVB_PLANE2PLANE_GLOBAL; residual in DEGREES (:334)    */
  PVB_PLANE_RELATIVE_IOU = 7,   /* PlaneRelativeIOUResidual :509-563  consts as PVB_PLANE_IOU                                      */
  PVB_LINE2LINE_ANGLE = 8       /* Line2Line_Angle     :984-1022  consts: dir_ref[0..2] dir_nei[3..5] (unit); rotation blocks only */
};

/* ---- lifecycle --------------------------------------------------------------------------------------------- */
int pvb_create(int device, pvb_ctx** out);
void pvb_destroy(pvb_ctx* ctx);
const char* pvb_last_error(const pvb_ctx* ctx);
int pvb_set_stream(pvb_ctx* ctx, void* cuda_stream); /* run on a caller-owned cudaStream_t (NULL: own stream)  */
int pvb_synchronize(pvb_ctx* ctx);
long pvb_kernel_launches(const pvb_ctx* ctx);        /* kernels this context has launched so far                */
void* pvb_stream(const pvb_ctx* ctx);                 /* the cudaStream_t work is enqueued on                    */

/* ---- A. correspondence-list mode: the ceres::CostFunction / EvaluationCallback surface -------------------- */
/* Registers n residual blocks (== n problem.AddResidualBlock calls, util/Optimization.cpp:417,549,555,592,602).
 * ref/nei index pose blocks; huber <= 0 means loss == nullptr; consts is n x 12 (layout per type above).        */
int pvb_blocks_set(pvb_ctx* ctx, long n, const int* type, const int* ref, const int* nei, const int* normalize,
                   const double* huber, const double* consts, int n_pose_blocks);
/* == PrepareForEvaluation: evaluates every block at `poses` (n_pose_blocks x 6, host).  want_rows: per-block
 * residual + 1x12 Jacobian rows are brought back to pinned host mirrors (original block order); 1 = rows with the
 * robust loss applied (Ceres' Corrector: r and J scaled by sqrt(rho')), 2 = RAW rows - the caller registers the
 * loss with its solver (ceres::HuberLoss(huber[i])), which then also computes the robust cost itself;
 * want_system: the per-edge normal equations (12x12 upper + gradient, 92 doubles) are reduced on the device.    */
int pvb_blocks_evaluate(pvb_ctx* ctx, const double* poses, int want_rows, int want_system);
const double* pvb_blocks_residuals(const pvb_ctx* ctx); /* n doubles (loss-corrected for want_rows = 1, raw for 2)   */
const double* pvb_blocks_jacobians(const pvb_ctx* ctx); /* n x 12 row-major [d aa_ref | d t_ref | d aa_nei | d t_nei] */
int pvb_blocks_cost(const pvb_ctx* ctx, double* cost, long* n_residuals);
/* device time (CUDA events) of the residual+Jacobian kernel of the last pvb_blocks_evaluate                          */
int pvb_blocks_kernel_time_ms(pvb_ctx* ctx, float* ms);
int pvb_blocks_num_edges(const pvb_ctx* ctx);
int pvb_blocks_edges(const pvb_ctx* ctx, int* ref, int* nei);
/* per edge: H upper-triangular row-major (78) | g (12) | cost | n_residuals  (parameter order aa_r,t_r,aa_n,t_n) */
int pvb_blocks_edge_systems(const pvb_ctx* ctx, double* out92);
/* the same systems where the last evaluate left them (pinned host memory, n_edges x 92 doubles; valid until the next
 * evaluate; NULL without want_system) - what include/panovlm_b200_reduced.hpp turns into one 13-residual block per edge */
const double* pvb_blocks_edge_systems_ptr(const pvb_ctx* ctx);
/* dense 6nb x 6nb J^T J, J^T r assembled on the host from the edge systems of the last evaluate                 */
int pvb_blocks_dense_system(const pvb_ctx* ctx, double* H, double* g, double* cost);
/* the ceres::Solve(SetOptionsLidar(...)) step (LidarOdometry.cpp:78-80): trust-region LM over the registered
 * blocks, evaluation on the device.  summary6 = initial_cost, final_cost, iterations,
 * successful, unsuccessful, termination (0 max-iter, 1 function tol, 2 gradient tol, 3 parameter tol, 4 failure) */
int pvb_blocks_solve_lm(pvb_ctx* ctx, double* poses, const unsigned char* is_const, int max_iterations, double* summary6);
/* where the linear algebra of the LM step runs — the analogue of SetOptionsLidar's linear_solver_type choice
 * (util/Optimization.cpp:647-662): HOST = dense Cholesky on the host cores; DEVICE = the edge systems stay in HBM, dense assembly,
 * Jacobi scaling, damping, blocked FP64 Cholesky and the triangular solves are kernels (pvb_solver.cuh), only 6N-vectors cross PCIe;
 * PCG = the same matrix as 6x6 blocks (no dense matrix, no factorisation) + block-Jacobi preconditioned conjugate gradients on the
 * device, converged to rounding (relative residual 1e-12) - the counterpart of the reference's sparse / iterative solvers for large pose
 * graphs; AUTO (default) = DEVICE from 256 free unknowns, PCG above 500 pose blocks (SetOptionsLidar: sparse from 50, iterative from 2000).  */
enum { PVB_SOLVER_AUTO = 0, PVB_SOLVER_HOST = 1, PVB_SOLVER_DEVICE = 2, PVB_SOLVER_PCG = 3 };
int pvb_blocks_set_linear_solver(pvb_ctx* ctx, int kind);
/* statistics of the PCG solver since the context was created: number of solves and total CG iterations */
int pvb_blocks_pcg_stats(const pvb_ctx* ctx, long* solves, long* iterations);
/* the device Cholesky solve alone: x = A^-1 b, A symmetric positive definite n x n row-major (parity / timing entry);
 * factor_ms (may be NULL) = device time of factorisation + substitution                                                             */
int pvb_cholesky_solve(pvb_ctx* ctx, const double* A, int n, const double* b, double* x, float* factor_ms);

/* ---- multi-GPU pose graphs: one process per GPU, edges sharded by reference frame -------------------------------------
 * Every rank registers only ITS residual blocks but the GLOBAL edge list, so all ranks reduce into the same layout (n_edges x 92 doubles,
 * edges of other ranks stay zero); the hook runs once per evaluation on the device buffer of edge systems, on the context's stream, before
 * anything reads them: there the caller does ONE sum-allreduce (NCCL) and every rank continues with the complete normal equations -
 * pvb_blocks_solve_lm then takes identical steps on all ranks.
 * pvb_blocks_set_edge_list: edges sorted by (ref, nei), unique; must hold the edge of every block registered afterwards; n_edges = 0 returns
 * to the single-GPU behaviour (edges derived from the blocks).  Call it before pvb_blocks_set.                                          */
typedef void (*pvb_reduce_hook)(void* user, double* device_edge_systems, long n_doubles, void* cuda_stream);
int pvb_blocks_set_edge_list(pvb_ctx* ctx, int n_edges, const int* ref, const int* nei);
int pvb_blocks_set_reduce_hook(pvb_ctx* ctx, pvb_reduce_hook hook, void* user);

/* ---- B. frames: transform to world + point-to-plane association per pose-graph edge ------------------------ */
typedef struct {
  const float* surf_target; int n_target; /* ref side: surfLessFlat, n x 4 (x,y,z,intensity=class), sensor frame */
  const float* surf_query; int n_query;   /* nei side: surfFlat                                                   */
} pvb_frame;

typedef struct {
  double plane_tolerance; /* lidar_plane_tolerance (base/Config.h:117)                                           */
  float dist_threshold;   /* point_to_plane_dis_threshold (Config.h:116), compared squared in float32             */
  int k;                  /* 10 in the reference (LidarFeatureAssociate.cpp:574); 5 and 10 are built              */
  double cell_size;       /* uniform-grid cell (m); <= 0: chosen per cloud from its extent and point count        */
} pvb_assoc_params;

int pvb_frames_set(pvb_ctx* ctx, int n_frames, const pvb_frame* frames);
/* AssociatePoint2Plane for every edge (ref[e], nei[e]) at the given poses (n_frames x 6): clouds -> world
 * (float32), cell-sorted targets, exact k-NN + plane fit.  *n_assoc = number of accepted correspondences.      */
int pvb_frames_associate_point2plane(pvb_ctx* ctx, const double* poses, int n_edges, const int* ref, const int* nei,
                                     const pvb_assoc_params* prm, long* n_assoc);
/* correspondences of the last association, edge-major, query order within an edge (the reference's push_back
 * order): edge index, query index in nei.surfFlat, point in the nei sensor frame, plane in the ref sensor frame */
int pvb_frames_get_point2plane(const pvb_ctx* ctx, long cap, int* edge, int* query, double* point3, double* plane4);
/* AddLidarPointToPlaneResidual (util/Optimization.cpp:506-562) without a host round trip: the association above, then the accepted
 * correspondences are compacted on the device (prefix sum) and written as residual blocks straight into the blocks-mode buffers -
 * equivalent to pvb_frames_get_point2plane + pvb_build_point2plane_blocks_edges + pvb_blocks_set, edge-major in the reference's order.
 * Frame f uses pose block block_offset + f of n_pose_blocks (joint problems keep the cameras in front).  n_extra host-built blocks
 * (line-to-line, camera-LiDAR, ...; arrays as pvb_blocks_set, may be 0 / NULL) are appended behind them.  *n_blocks = total.          */
int pvb_frames_point2plane_blocks(pvb_ctx* ctx, const double* poses, int n_edges, const int* ref, const int* nei, const pvb_assoc_params* prm,
                                  int angle_residual, int normalize_distance, double weight, int block_offset, int n_pose_blocks, long n_extra,
                                  const int* x_type, const int* x_ref, const int* x_nei, const int* x_normalize, const double* x_huber,
                                  const double* x_consts, long* n_blocks);
/* debug/parity view of the k-NN itself for one edge: indices into ref.surfLessFlat and float32 squared distances */
int pvb_frames_knn(pvb_ctx* ctx, const double* poses, int ref, int nei, const pvb_assoc_params* prm, int* idx, float* d2);

/* AssociatePoint2Line (LidarFeatureAssociate.cpp:478-548) on the frames' cornerLessSharp clouds (n x 4 float32, sensor frame):
 * 5 nearest reference corner points, PCA line test (FormLine(pts, 10, 0.05)) in the world frame; per correspondence the query in the
 * neighbour's sensor frame and the two synthetic line points c +- 0.1 d in the reference sensor frame.                           */
int pvb_frames_set_corners(pvb_ctx* ctx, int n_frames, const float* const* corner, const int* n_corner);
int pvb_frames_associate_point2line(pvb_ctx* ctx, const double* poses, int n_edges, const int* ref, const int* nei, float dist_threshold,
                                    double cell_size, long* n_assoc);
int pvb_frames_get_point2line(const pvb_ctx* ctx, long cap, int* edge, int* query, double* point3, double* a3, double* b3);

/* ---- C. dense ICP sweep (BASELINE.json configs[4]): fused transform + k-NN + plane fit + residual + reduce --- */
typedef struct {
  double plane_tolerance;
  float dist_threshold;
  int k;
  int residual_type; /* PVB_P2PLANE_METER or PVB_P2PLANE_ANGLE                                                    */
  int normalize;     /* normalize_distance (Config.h:113)                                                         */
  double huber;      /* 0.2 (metre) / 2 deg (angle), util/Optimization.cpp:513-517; <= 0: none                    */
  double weight;     /* lidar_weight                                                                              */
} pvb_dense_params;

/* target cloud in the world frame (n x 4 float32: x,y,z,class); builds the cell-sorted layout in HBM             */
int pvb_dense_set_target(pvb_ctx* ctx, const float* xyzc, long n, double cell_size);
/* source frames in their sensor frames, concatenated; frame f = [offsets[f], offsets[f+1]).  Points of a frame
 * are re-ordered along a Morton curve on upload so that a thread block works on a compact patch.              */
int pvb_dense_set_sources(pvb_ctx* ctx, const float* xyzc, const int* offsets, int n_frames);
/* one Gauss-Newton evaluation at poses_lw (n_frames x 6): per frame 29 doubles = H upper 6x6 (21) | g (6) | cost |
 * n_residuals w.r.t. the frame's own pose (the target/reference pose is constant = identity).                  */
int pvb_dense_evaluate(pvb_ctx* ctx, const double* poses_lw, const pvb_dense_params* prm, double* out_sys29);
/* same, but leaves the reduced systems on the device (for an on-stream allreduce).  If *dev_sys29 is non-NULL on entry
 * it is a caller-owned device buffer (n_frames x 29 doubles) the result is written to; otherwise it receives the
 * context's own buffer.                                                                                            */
int pvb_dense_evaluate_device(pvb_ctx* ctx, const double* poses_lw, const pvb_dense_params* prm, double** dev_sys29);
/* Search-radius hints of the dense path.  Every evaluation stores, per source point, its world position and the squared
 * distance of its K-th neighbour; the next evaluation bounds that point's search by sqrt(tau_old) + |q - q_old| (the K old
 * neighbours still lie within it).  A hint only shortens the walk - a search that finds fewer than K points below it starts
 * again without it - so results never depend on it.  enable = 0 ignores the stored hints (they are still rewritten);
 * pvb_dense_reset_hints forgets them (the next evaluation is a cold search).  Hints are forgotten whenever the target or the
 * layout of the source frames changes.                                                                              */
int pvb_dense_set_hints(pvb_ctx* ctx, int enable);
int pvb_dense_reset_hints(pvb_ctx* ctx);
/* debug: out2[0] = tiles of the fused kernel staged through TMA so far, out2[1] = tiles that used the global-memory path */
int pvb_debug_counters(pvb_ctx* ctx, unsigned long long* out2);
/* dense mode: how often the queries were sorted by target cell, and how often a fresh upload re-used the previous permutation (same layout, poses near those of the
 * last sort, order still local when last measured) */
int pvb_dense_order_stats(const pvb_ctx* ctx, long* sorts, long* reuses);
/* device time (CUDA events on the context's stream) of the fused associate+residual kernel of the last dense evaluate */
int pvb_dense_kernel_time_ms(pvb_ctx* ctx, float* ms);
/* Gauss-Newton/LM step per frame from the reduced 6x6 systems (host, 64 tiny solves): poses updated in place.   */
int pvb_dense_gauss_newton_step(const double* sys29, int n_frames, double lambda, double* poses_lw);
/* per-query view of the last evaluate for parity: valid flag, point (nei frame), plane, residual, 6 Jacobian cols */
int pvb_dense_get_rows(pvb_ctx* ctx, const double* poses_lw, const pvb_dense_params* prm, unsigned char* valid,
                       double* point3, double* plane4, double* residual, double* jac6);

/* ---- D. bulk SE(3) + equirectangular projection (util/Visualization.h:408-441) -------------------------- */
/* uvd: n x 3 float32 = pixel.x, pixel.y (FastAtan2 path, base/Math.h:15-29), depth = |p|                          */
int pvb_project_equirect(pvb_ctx* ctx, const float* xyzi, long n, const double* T_cl16, int rows, int cols, float* uvd);
/* the sparse depth image itself: uint16(depth*256) splatted over (size+1)^2 windows, last point wins             */
int pvb_project_depth_image(pvb_ctx* ctx, const float* xyzi, long n, const double* T_cl16, int rows, int cols, int size, uint16_t* image);

/* ---- E. line-to-line vote matrix (LidarFeatureAssociate.cpp:459-473) ------------------------------------- */
/* M[nei_seg][ref_seg] += 1 for every nei corner point (world, float32) within dist_threshold of the ref line     */
int pvb_line_votes(pvb_ctx* ctx, const double* ref_lines_world6, int n_ref_lines, const float* nei_corner_world, int n_points,
                   const int* p2s_off, const int* p2s_ids, int n_nei_lines, double dist_threshold, int* M);

/* ---- F. camera-LiDAR line association votes (CameraLidarLineAssociate.cpp:389-414) ------------------------ */
/* counts[image_line][lidar_segment] of AssociateByAngle's inner loop; cloud is cornerLessSharp in the LiDAR frame */
int pvb_angle_votes(pvb_ctx* ctx, int rows, int cols, const float* lines4, int n_lines, const float* cloud_local, int n_points,
                    const int* p2s_off, const int* p2s_ids, int n_segments, const double* T_cl16, int* counts);

/* ---- G. host-side builders around the kernels (C++ in the library, no device math of their own) ----------- */
/* FindNeighbors (lidar_mapping/LidarFeatureAssociate.cpp:19-111): k nearest frame centres (float32, self excluded),
 * previous / next frame with a valid pose, loop candidates within 20 m more than 200 frames away from all but one of
 * the current neighbours; frames without a pose get their temporal neighbours.  t_wl: n x 3 frame positions.
 * out_offsets: n+1, out_neighbors: up to cap entries.  Returns the number of neighbour entries or < 0.          */
int pvb_find_neighbors(int n_frames, const double* t_wl, const unsigned char* pose_valid, const unsigned char* frame_valid,
                       int neighbor_size, int* out_offsets, int* out_neighbors, int cap);

/* CameraLidarOptimizer::NeighborEachFrame (joint_optimization/CameraLidarOptimizer.cpp:551-610): per image the LiDAR frames its lines are
 * associated with - a temporal window of neighbor_size indices, or the neighbor_size nearest LiDAR centres (+ previous / next index).
 * t_wc / t_wl: camera / LiDAR positions (n x 3); CSR output as pvb_find_neighbors; returns the number of entries or < 0.  Host only.       */
int pvb_neighbor_each_frame(int n_frames, int n_lidars, int neighbor_size, int temporal, const double* t_wc, const unsigned char* frame_pose_valid,
                            const double* t_wl, const unsigned char* lidar_pose_valid, const unsigned char* lidar_valid, int* out_offsets,
                            int* out_neighbors, int cap);
/* CameraLidarOptimizer::LidarMaskByTrack (:612-642) after GenerateTracks (pvb_generate_line_tracks): mask[seg_off[f] + line] = 1 for every
 * LiDAR line that belongs to a track.  Host only.                                                                                      */
int pvb_lidar_mask_by_track(int n_tracks, const int* track_off, const int* feat_frame, const int* feat_line, int n_lidars, const int* seg_off,
                            unsigned char* mask);

typedef struct {
  const float* corner_local; int n_corner;      /* cornerLessSharp, n x 4, sensor frame                                 */
  const int* p2s_off; const int* p2s_ids;       /* point_to_segment as CSR (sensors/Velodyne.h:89)                       */
  int n_segments;
  const double* segment_coeffs;                 /* n_segments x 6: point + unit direction, sensor frame (Velodyne.h:87)   */
  const double* end_points;                     /* n_segments x 2 x 3, sensor frame (Velodyne.h:88); may be NULL for A2   */
  const double* R_wl; const double* t_wl;       /* pose (row-major R)                                                      */
} pvb_line_frame;

/* AssociateLine2Line (LidarFeatureAssociate.cpp:442-476) + FindAssociations (:120-197): vote matrix on the device, the
 * per-segment arg-max / 7 degree / many-to-one tail on the host.  Outputs sized for nei->n_segments entries:
 * neighbour line, reference line, and the two synthetic end points c +- 0.1 d of the reference line (its sensor frame). */
int pvb_line2line_associate(pvb_ctx* ctx, const pvb_line_frame* ref, const pvb_line_frame* nei, double dist_threshold,
                            int* n_out, int* nei_line, int* ref_line, double* point_a3, double* point_b3);

/* ---- segment-based variants of the corner association (LidarFeatureAssociate.cpp:238-440; `point_to_line_residual` paths) ---- */
/* pcl::KdTreeFLANN::nearestKSearch(k = 5) of every neighbour corner point in the reference's cornerLessSharp cloud, both moved to the
 * world frame (float32) with the given T_wl: idx5 = n_nei x 5 indices into ref_local (unordered), a row of -1 when the 5th neighbour
 * is beyond dist_threshold (:261, :416) or the reference cloud has fewer than 5 points.  cell_size <= 0: chosen from the cloud.     */
int pvb_pair_knn5(pvb_ctx* ctx, const float* ref_local, int n_ref, const double* R_ref, const double* t_ref, const float* nei_local, int n_nei,
                  const double* R_nei, const double* t_nei, float dist_threshold, double cell_size, int* idx5);
/* arg-min / min over the infinite lines (point + direction, world) of PointToLineDistance3D for every point (world float32, n x 4);
 * the first minimum wins (:335-341); line = -1 when there are no lines                                                             */
int pvb_nearest_line(pvb_ctx* ctx, const double* lines_world6, int n_lines, const float* points_world, int n_points, int* line, double* dist);
/* AssociatePoint2LineSegmentKNN (:238-317): a query whose 5 neighbours all belong to one reference segment is associated with that
 * segment's line; outputs per association: query index in nei.cornerLessSharp, reference segment, the query in the neighbour's sensor
 * frame and the synthetic line points c +- 0.1 d in the reference sensor frame.  The *_tail forms take the k-NN result as input.     */
int pvb_point2line_segment_knn_associate(pvb_ctx* ctx, const pvb_line_frame* ref, const pvb_line_frame* nei, float dist_threshold, long cap, long* n_out,
                                         int* query, int* ref_line, double* point3, double* a3, double* b3);
int pvb_point2line_segment_knn_tail(const pvb_line_frame* ref, const pvb_line_frame* nei, const int* idx5, long cap, long* n_out, int* query, int* ref_line,
                                    double* point3, double* a3, double* b3);
/* AssociatePoint2LineSegment (:319-383): nearest reference line, accepted within dist_threshold                                     */
int pvb_point2line_segment_associate(pvb_ctx* ctx, const pvb_line_frame* ref, const pvb_line_frame* nei, float dist_threshold, long cap, long* n_out,
                                     int* query, int* ref_line, double* point3, double* a3, double* b3);
/* AssociateLine2LineKNN (:385-440): votes from segments holding >= 3 of a point's 5 neighbours, then FindAssociations (:120-197);
 * outputs as pvb_line2line_associate                                                                                                */
int pvb_line2line_knn_associate(pvb_ctx* ctx, const pvb_line_frame* ref, const pvb_line_frame* nei, float dist_threshold, int* n_out, int* nei_line,
                                int* ref_line, double* point_a3, double* point_b3);
int pvb_line2line_knn_tail(const pvb_line_frame* ref, const pvb_line_frame* nei, const int* idx5, int* n_out, int* nei_line, int* ref_line, double* point_a3,
                           double* point_b3);

/* ---- LiDAR line tracks (lidar_mapping/LidarLineMatch.cpp:36-86, util/Tracks.cpp:58-186) --------------------------------------- */
/* TrackBuilder::Build + Filter(min_track_length) + ExportTracks on line matches given as CSR: pair p = (frame pair_a[p], frame pair_b[p]),
 * its matches match_off[p]..match_off[p+1] = (line of a, line of b).  Output CSR: track t = features track_off[t]..track_off[t+1] as
 * (feat_frame, feat_line), ascending; tracks numbered as LidarLineMatch assigns ids (:80-81).  Host only.                            */
int pvb_line_tracks_build(int n_pairs, const int* pair_a, const int* pair_b, const int* match_off, const int* match_a, const int* match_b, int min_track_length,
                          int allow_multiple_map, int cap_features, int* n_tracks, int* track_off, int* feat_frame, int* feat_line);
/* the gate of AddLidarLineToLineResidual2 (util/Optimization.cpp:343-400): keep[i] = the reference line and the neighbour line of
 * association i share a track.  Host only.                                                                                          */
int pvb_line_tracks_gate(int n_tracks, const int* track_off, const int* feat_frame, const int* feat_line, int ref_frame, int nei_frame, int n, const int* ref_line,
                         const int* nei_line, unsigned char* keep);
/* LidarLineMatch::GenerateTracks: AssociateLine2Line(frames[nei], frames[i], dist_threshold = 0.3) for every frame i with a valid pose and
 * every neighbour (CSR nbr_off / nbr_ids from pvb_find_neighbors), then pvb_line_tracks_build(allow_multiple_map = 1).              */
int pvb_generate_line_tracks(pvb_ctx* ctx, int n_frames, const pvb_line_frame* frames, const unsigned char* pose_valid, const int* nbr_off, const int* nbr_ids,
                             double dist_threshold, int min_track_length, int cap_features, int* n_tracks, int* track_off, int* feat_frame, int* feat_line);

/* The vote matrices of AssociateLine2Line for many frame pairs in two launches (all corner clouds -> world, all pairs' votes) and one
 * download.  lines_world: every frame's segment lines in the world frame (TransformLines), concatenated in frame order (6 doubles each);
 * M: pair p's n_segments(nei) x n_segments(ref) matrix at m_off[p] (m_total ints in all); world_out (may be NULL): the world-frame corner
 * clouds concatenated in frame order (x, y, z valid).                                                                                  */
int pvb_line_votes_batch(pvb_ctx* ctx, int n_frames, const pvb_line_frame* frames, const double* lines_world, int n_pairs, const int* pair_ref, const int* pair_nei,
                         const long long* m_off, long long m_total, double dist_threshold, int* M, float* world_out);
/* AddLidarLineToLineResidual2 (util/Optimization.cpp:329-441) for a whole pose graph in ONE call: AssociateLine2Line of every edge
 * (ref[e], nei[e]) from the batched device pass above, the FindAssociations tails on the host cores, the line-track gate (:383-400;
 * tracks as pvb_generate_line_tracks returns them, n_tracks < 0: no gate) and one Point2Line block per point of every kept neighbour
 * segment, appended edge by edge in the reference's order.  Returns the new block count (>= at) or a negative error code.              */
int pvb_frames_line2line_blocks(pvb_ctx* ctx, int n_frames, const pvb_line_frame* frames, int n_edges, const int* ref, const int* nei, double dist_threshold,
                                int n_tracks, const int* track_off, const int* feat_frame, const int* feat_line, int angle_residual, int normalize_distance,
                                double weight, long at, long cap, int* type, int* ref_out, int* nei_out, int* normalize, double* huber, double* consts);
/* The same family with the tails on the device as well (csrc/pvb_lines.cuh): the vote matrices stay in HBM, FindAssociations
 * (lidar_mapping/LidarFeatureAssociate.cpp:120-197) and the track gate run as one thread per edge, and the Point2Line blocks are written by the
 * device straight into the block arrays when the NEXT pvb_frames_point2plane_blocks call sizes them (placed after that call's extra blocks,
 * one reduction edge per pose-graph edge; with pvb_blocks_set_edge_list: filed under the global edges).  Only the per-edge block counts come
 * back to the host.  The block constants are bit-identical to pvb_frames_line2line_blocks'.  *n_blocks = the line blocks that wait; they are
 * dropped by the next pvb_line_votes_batch / pvb_generate_line_tracks call or consumed by the next pvb_frames_point2plane_blocks.           */
int pvb_frames_line2line_blocks_device(pvb_ctx* ctx, int n_frames, const pvb_line_frame* frames, int n_edges, const int* ref, const int* nei, double dist_threshold,
                                       int n_tracks, const int* track_off, const int* feat_frame, const int* feat_line, int angle_residual, int normalize_distance,
                                       double weight, long* n_blocks);

/* CameraLidarLineAssociate::AssociateByAngle (joint_optimization/CameraLidarLineAssociate.cpp:340-475) followed by
 * Filter(false, filter_by_length) (:628-715) and, unless multiple_association, UniqueLinePair (:754-876): per (image line, LiDAR
 * segment) vote counts on the device; acceptance tests, projected-length filter, one-to-one reduction and the transform back to the
 * LiDAR frame on the host.  image_line_mask / lidar_line_mask (may be NULL = all lines take part, :351-366): 0 excludes a line.
 * The reference's default is multiple_association = false (CameraLidarLineAssociate.h:109); the joint optimisation passes true with
 * masks (CameraLidarOptimizer.cpp:362-364).  Outputs sized for cap pairs.                                                           */
int pvb_camera_lidar_associate(pvb_ctx* ctx, int rows, int cols, const float* lines4, int n_lines, const pvb_line_frame* lidar,
                               const double* T_cl16, int filter_by_length, int multiple_association, const unsigned char* image_line_mask,
                               const unsigned char* lidar_line_mask, int cap, int* n_out, int* image_line, int* lidar_line,
                               double* start3, double* end3, float* angle);
/* ---- pixel-space CameraLidarLineAssociate::Associate, first stage (joint_optimization/CameraLidarLineAssociate.cpp:22-102): the fallback
 * for frames without LiDAR line segments (CameraLidarOptimizer.cpp:360-367).  The RANSAC line fit that follows in the reference (:105-110) is
 * PCL's randomised SACSegmentation and has no deterministic counterpart here; these entry points deliver its input.                        */
/* image lines -> sub-line mid points (BreakToSegments(line, 70), seam pieces skipped, :38-54); returns their number.  Host only.            */
int pvb_pixel_sub_lines(int rows, int cols, const float* lines4, int n_lines, int cap, float* mid2, int* sub_to_line);
/* cv::flann knnSearch(k = 3) of every projected LiDAR point among the mid points (:74-79): cloud -> camera frame (float32) -> pixel (CamToImage
 * with FastAtan2) -> idx3 = the 3 nearest mid points by ascending float squared L2 (-1 when there are fewer), d2_3 / pixel2 optional views   */
int pvb_pixel_knn3(pvb_ctx* ctx, int rows, int cols, const float* mid2, int n_mid, const float* cloud_local, int n_points, const double* T_cl16,
                   int* idx3, float* d2_3, float* pixel2);
/* the two above + the 60 px gate (:81): line3 = n_points x 3 image-line index of the k-th nearest mid point, or -1                              */
int pvb_pixel_line_neighbors(pvb_ctx* ctx, int rows, int cols, const float* lines4, int n_lines, const float* cloud_local, int n_points,
                             const double* T_cl16, int* line3, float* d2_3, float* pixel2);
/* `line_lidar` (:83-97) as CSR: per image line the LiDAR points that chose it (ascending, with multiplicity), emptied below min_points (6).
 * Returns the number of entries.  Host only.                                                                                                */
int pvb_pixel_line_candidates(int n_lines, int n_points, const int* line3, int min_points, int cap, int* line_off, int* lidar_idx);
/* FitLineRANSAC (:717-752) + the end-point tail of the pixel-space Associate (:117-137) for ONE candidate list (camera-frame points, `stride` floats per
 * point, x y z first).  PARITY UNPINNED for the sample-consensus part: the reference calls pcl::SACSegmentation (a system dependency that is neither in the
 * reference tree nor installed here); this restates PCL 1.10's RANSAC over SACMODEL_LINE (boost mt19937 seeded 12345, uniform_int(0, INT_MAX), index shuffle
 * sampling, adaptive iteration count) with the reference's settings: dist_threshold 0.1, PCL defaults max_iterations 50, probability 0.99.  The reference's
 * own code after it is restated as written: float32 centroid / covariance of the inliers, direction = eigenvector of the largest eigenvalue (pcl::eigen33),
 * farthest inlier pair (whose POSITIONS in the inlier list index the candidate list, :136-137), ProjectPoint2Line3D in double.
 * Returns the number of inliers; fewer than 3 = no line (the reference's `false`; outputs untouched).  inliers: ascending, at most `cap`.  Host only.        */
int pvb_pixel_fit_line(const float* xyz, int n, int stride, double dist_threshold, int max_iterations, double probability, float* coeff6, int cap, int* inliers,
                       double* start3, double* end3);
/* The fit of every candidate list of one image (the loop of Associate :88-146) on all host threads: list l = cloud_cam[lidar_idx[line_off[l] .. line_off[l+1])]
 * (camera-frame cloud, 4 floats per point; CSR from pvb_pixel_line_candidates).  n_inliers[l] < 3 = no line for image line l; coeff6 / start3 / end3 are per line. */
int pvb_pixel_fit_lines(const float* cloud_cam, int n_points, int n_lines, const int* line_off, const int* lidar_idx, double dist_threshold, int max_iterations,
                        double probability, int* n_inliers, float* coeff6, double* start3, double* end3);
/* CameraLidarLineAssociate::Filter (:628-715) alone, on pairs whose LiDAR end points are in the CAMERA frame: the angle branch (great-circle
 * planes within 5 deg, LiDAR ends inside the image line's arc, both ends within 0.4 of the image plane at radius 5; angle[i] = plane angle in
 * degrees, :652) and the projected-length branch (100 .. 2000 px).  keep[i] = 1 when the pair survives.  Host only.                          */
int pvb_filter_line_pairs(int rows, int cols, int n, const float* image_line4, const double* start3, const double* end3, int filter_by_angle,
                          int filter_by_length, unsigned char* keep, float* angle);
/* UniqueLinePair alone (host): candidates in input order -> one-to-one pairs, ascending image line                                  */
int pvb_unique_line_pairs(int n, const int* image_line, const int* lidar_line, const float* score, int* n_out, int* out_image, int* out_lidar,
                          float* out_score);

/* Residual-block builders == the AddResidualBlock loops of util/Optimization.cpp.  Each appends to caller arrays
 * (type, ref, nei, normalize: int; huber: double; consts: 12 doubles per block) starting at index `at` and returns the new
 * count (or < 0).  Angle residuals use HuberLoss(2 deg) for planes and no loss for line-to-line (Optimization.cpp:417).  */
int pvb_build_point2plane_blocks(long n, const double* point3, const double* plane4, int ref_block, int nei_block, int angle_residual,
                                 int normalize_distance, double weight, long at, long cap, int* type, int* ref, int* nei, int* normalize,
                                 double* huber, double* consts);           /* Optimization.cpp:506-562 */
/* the same for the correspondences of many edges at once: correspondence i belongs to edge[i] (as pvb_frames_get_point2plane returns them),
 * whose pose blocks are edge_ref_block[edge[i]] / edge_nei_block[edge[i]]; blocks are appended in the order of the correspondences             */
int pvb_build_point2plane_blocks_edges(long n, const int* edge, const double* point3, const double* plane4, int n_edges, const int* edge_ref_block,
                                       const int* edge_nei_block, int angle_residual, int normalize_distance, double weight, long at, long cap,
                                       int* type, int* ref, int* nei, int* normalize, double* huber, double* consts);
/* AddLidarPointToLineResidual (Optimization.cpp:443-504): Point2Line_Angle / _Meter with HuberLoss(2 deg / 0.2) */
int pvb_build_point2line_blocks(long n, const double* point3, const double* a3, const double* b3, int ref_block, int nei_block, int angle_residual,
                                int normalize_distance, double weight, long at, long cap, int* type, int* ref, int* nei, int* normalize,
                                double* huber, double* consts);
/* one block per point of the neighbour segment (world float32 -> neighbour sensor frame, Optimization.cpp:403-431) */
int pvb_build_line2line_blocks(const pvb_line_frame* nei, const float* nei_corner_world, int nei_line, const double* point_a3,
                               const double* point_b3, int ref_block, int nei_block, int angle_residual, int normalize_distance, double weight,
                               long at, long cap, int* type, int* ref, int* nei_out, int* normalize, double* huber, double* consts);
/* two blocks per pair: Plane2Plane_Global + PlaneIOUResidual, HuberLoss(3 deg) (Optimization.cpp:564-607) */
int pvb_build_camera_lidar_blocks(int rows, int cols, int n_pairs, const float* image_line4, const double* start3, const double* end3,
                                  const float* pair_weight, int cam_block, int lidar_block, double weight, long at, long cap, int* type, int* ref,
                                  int* nei, int* normalize, double* huber, double* consts);
/* calibration mode, CameraLidarOptimizer::Optimize(line_pairs, T_cl) (joint_optimization/CameraLidarOptimizer.cpp:32-64): per pair
 * Plane2Plane_Relative + HuberLoss(2 deg) and PlaneRelativeIOUResidual (weight 2, no loss) on the ONE relative pose block (aa_cl, t_cl)    */
int pvb_build_calibration_blocks(int rows, int cols, int n_pairs, const float* image_line4, const double* start3, const double* end3, int pose_block,
                                 long at, long cap, int* type, int* ref, int* nei, int* normalize, double* huber, double* consts);
/* pcl::transformPointCloud of one cloud on the device (sensors/Velodyne.cpp:1790-1806): n x 4 float32 in, n x 4 out     */
int pvb_transform_cloud(pvb_ctx* ctx, const float* xyzi, long n, const double* R9, const double* t3, float* out);

/* ---- H. per-frame preprocessing either side of the path -------------------------------------------------- */
/* SlerpPose (base/Geometry.hpp:572-583): poses are 4x4 row-major T_world<-local; ratio 0 -> pose_w1, 1 -> pose_w2. Host only.           */
int pvb_slerp_pose(const double* pose_w1_16, const double* pose_w2_16, double ratio, double* out16);
/* The pose of the END of every sweep as LidarOdometry::UndistortLidars chooses it (lidar_mapping/LidarOdometry.cpp:203-243, sweep 0.1 s
 * + gap_time between sweeps): interpolated towards the next usable frame, extrapolated for the last frame; has_end[i] = 0 where the
 * reference saves the raw cloud (`goto save_undistort`).  poses16: n x 16; pose_valid / frame_valid: IsPoseValid() / valid. Host only.
 * A frame without a pose is expected as the reference stores it (R = 0, t = inf).  Like the reference, the search only steps over frames that
 * have NEITHER a pose NOR valid data (`!IsPoseValid() && !valid`, :220, :231): next to a valid frame without a pose the end pose is NaN
 * (has_end = 1) and the undistorted sweep comes out as NaN - identical to LidarOdometry::UndistortLidars (tests/golden/ref_velodyne.npz).   */
int pvb_undistort_end_poses(int n, const double* poses16, const unsigned char* pose_valid, const unsigned char* frame_valid, float gap_time,
                            double* out_pose16, unsigned char* has_end);
/* Pose text files (util/FileIO.cpp:11-73 ReadPoseT, :168-191 ExportPoseT): one line per frame, [name ]r00 r01 r02 tx r10 r11 r12 ty r20 r21 r22 tz
 * with 6 significant digits (the reference's default ostream precision); inf / nan marks a frame without pose.  R9 row-major.  Host only.
 * read: returns the number of poses (<= cap) or < 0; invalid lines are kept (valid = 0, R = 0, t = +inf) only when with_invalid; names may be NULL,
 * else cap x name_len characters.                                                                                                          */
int pvb_write_poses_text(const char* path, int n, const double* R9, const double* t3, const char* const* names);
int pvb_read_poses_text(const char* path, int with_invalid, int cap, double* R9, double* t3, unsigned char* valid, char* names, int name_len);
/* Velodyne::UndistortCloud (sensors/Velodyne.cpp:1642-1674) for a batch of frames in one launch: point i of a frame's n points (scan
 * order) is moved by slerp(identity, q_se, float(i)/float(n)) and the same fraction of t_se, where (q_se, t_se) = T_wl^-1 T_we.
 * xyzi / out: concatenated n x 4 float32 clouds, frame f = [offsets[f], offsets[f+1]); T_wl16 / T_we16: n_frames x 16 (start / end
 * pose of the sweep); frames with has_end[f] == 0 (NULL: none) pass through unchanged.                                                 */
int pvb_undistort_clouds(pvb_ctx* ctx, const float* xyzi, const int* offsets, int n_frames, const double* T_wl16, const double* T_we16,
                         const unsigned char* has_end, float* out);

/* ---- I. camera-camera reprojection residuals and their bundle adjustment -------------------------------------- */
/* Registers the observation list of AddCameraResidual (util/Optimization.cpp:172-222, ANGLE_RESIDUAL_1): observation i = one
 * problem.AddResidualBlock(PanoramaReprojResidual_1Angle(bearing_i, weight), HuberLoss(huber), aa_cw[cam_i], t_cw[cam_i], point_3d[point_i])
 * (base/CostFunction.h:218-247).  bearing3: the key point on the unit sphere (eq.ImageToCam), normalised here like the functor's
 * constructor does; huber <= 0: no loss (the reference passes 4 deg).  Camera blocks are 6 doubles (aa_cw, t_cw), points 3 doubles.   */
int pvb_reproj_set(pvb_ctx* ctx, long n_obs, const int* cam, const int* point, const double* bearing3, double weight, double huber, int n_cams,
                   long n_points);
/* == PrepareForEvaluation for these blocks.  want_rows: residual + 1x9 Jacobian row [d aa_cw | d t_cw | d point] per observation in the
 * caller's order (pinned host mirrors; 1 = loss-corrected, 2 = raw rows for a solver that applies the loss itself - not together with
 * want_system); want_system: the blocks of the normal equations are reduced on the device — per camera
 * J_c^T J_c (upper 6x6, 21) and gradient (6), per point J_p^T J_p (upper 3x3, 6) and gradient (3), per observation the 6x3 coupling
 * block J_c^T J_p.                                                                                                                  */
int pvb_reproj_evaluate(pvb_ctx* ctx, const double* cams6, const double* points3, int want_rows, int want_system);
const double* pvb_reproj_residuals(const pvb_ctx* ctx);   /* n_obs doubles (loss-corrected / raw)       */
const double* pvb_reproj_jacobians(const pvb_ctx* ctx);   /* n_obs x 9 row-major                        */
int pvb_reproj_cost(const pvb_ctx* ctx, double* cost);
/* the reduced blocks of the last evaluate (any pointer may be NULL): cam_H21 n_cams x 21, cam_g6 n_cams x 6, pt_H6 n_points x 6,
 * pt_g3 n_points x 3, obs_E18 n_obs x 18 (row-major 6x3, caller's observation order)                                                 */
int pvb_reproj_blocks(pvb_ctx* ctx, double* cam_H21, double* cam_g6, double* pt_H6, double* pt_g3, double* obs_E18);
int pvb_reproj_kernel_time_ms(pvb_ctx* ctx, float* ms);   /* device time of the residual + Jacobian kernel  */
/* SfMGlobalBA (util/Optimization.cpp:10-82): ceres::Solve over cameras and structure with Ceres' trust-region defaults.  Each step
 * eliminates the points on the device (Schur complement, the role of SetOptionsSfM's DENSE_SCHUR / SPARSE_SCHUR, :611-636), factors the
 * reduced camera system with the blocked FP64 Cholesky and back-substitutes the points.  cam_param_const: n_cams x 6 flags (NULL: all
 * free) — SetParameterBlockConstant on aa_cw / t_cw sets three of them (:40-43, :54-55); point_const: n_points flags (:45-47).
 * cams6 / points3 are updated in place; summary6 as pvb_blocks_solve_lm.                                                            */
int pvb_reproj_solve_lm(pvb_ctx* ctx, double* cams6, double* points3, const unsigned char* cam_param_const, const unsigned char* point_const,
                        int max_iterations, double* summary6);
/* CameraLidarOptimizer::Optimize's solve (joint_optimization/CameraLidarOptimizer.cpp:387-548): ONE trust-region problem over the pose blocks
 * [cameras | LiDARs] and the structure points with the reprojection observations (pvb_reproj_set; cam = index of the camera's pose block) and
 * the LiDAR-LiDAR / camera-LiDAR residual blocks (pvb_blocks_set) together.  pose_param_const: n_pose_blocks x 6 flags (NULL: all free), e.g.
 * camera 0 constant (:490-491) or the refine_* switches (:466-488); point_const as pvb_reproj_solve_lm (refine_structure, :462-465).        */
int pvb_joint_solve_lm(pvb_ctx* ctx, double* poses6, double* points3, const unsigned char* pose_param_const, const unsigned char* point_const,
                       int max_iterations, double* summary6);
/* The observation loop of AddCameraResidual (util/Optimization.cpp:187-219): tracks as CSR (track t = features track_off[t] ..
 * track_off[t+1]: frame feat_frame[f], key point feat_xy[2f..2f+1] in pixels); features of frames without a valid pose are skipped;
 * the key point is rounded to the pixel grid and mapped to the unit sphere in float32 exactly like Equirectangular::ImageToCam(cv::Point2i)
 * (sensors/Equirectangular.h:155-164).  Returns the number of observations written (cam = frame, point = track) or < 0.  Host only.  */
int pvb_build_reproj_observations(int rows, int cols, long n_tracks, const int* track_off, const int* feat_frame, const float* feat_xy,
                                  const unsigned char* pose_valid, long cap, int* cam, int* point, double* bearing3);

#ifdef __cplusplus
}
#endif
#endif /* PANOVLM_B200_H_ */
