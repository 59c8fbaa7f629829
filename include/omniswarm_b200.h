/*
 * omniswarm_b200.h -- C ABI of libomniswarm_b200.so
 *
 * B200-native (sm_100a) replacement for the ONE compute-heavy path of HKUST-Aerial-Robotics/Omni-swarm:
 * the swarm_loop keyframe front-end and the swarm_localization pose-graph solve.  The reference has no
 * plugin ABI for this path (the boundary is C++ member calls inside one process);
 * every entry point below names the reference call site it replaces (paths relative to /root/reference).
 * INTEGRATION.md shows the thin C++ adapter classes a maintainer adds so that loop_cam.cpp /
 * loop_detector.cpp / swarm_localization_solver.cpp keep their signatures.
 *
 * Conventions
 *   - plain C types only; every function returns an osb_status (0 = OK) and never throws;
 *   - "host" entry points take HOST pointers and do their own H2D/D2H (the drop-in calls);
 *     "_dev" entry points take DEVICE pointers plus a cudaStream_t (passed as void*) and never synchronise:
 *     they exist so that a caller that already owns device memory (bench.py, the multi-GPU keyframe exchange)
 *     can keep everything resident;
 *   - all device memory is owned by the opaque handles for their lifetime (as the reference's
 *     TensorRTInferenceGeneric does, swarm_loop/src/tensorrt_generic.cpp:99-120);
 *   - handles are internally serialised (one stream + one mutex per handle): the reference calls
 *     LoopDetector from the ROS thread and the LCM thread without a lock;
 *   - there is NO CPU fallback: without a CUDA device every create() returns OSB_ERR_NO_DEVICE.
 */
#ifndef OMNISWARM_B200_H
#define OMNISWARM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int osb_status;
#define OSB_OK 0
#define OSB_ERR_INVALID 1    /* bad argument (null pointer, size mismatch, batch > max_batch ...) */
#define OSB_ERR_CUDA 2       /* a CUDA call failed; osb_last_error() has the text */
#define OSB_ERR_CAPACITY 3   /* database / solver capacity exceeded */
#define OSB_ERR_NO_DEVICE 4  /* no CUDA device: this library has no CPU path */

#define OSB_SP_DESC_RAW_LEN 256   /* SP_DESC_RAW_LEN, swarm_loop/include/swarm_loop/loop_defines.h */
#define OSB_FEATURE_DESC_SIZE 64  /* FEATURE_DESC_SIZE, loop_defines.h:67 */
#define OSB_DEEP_DESC_SIZE 4096   /* DEEP_DESC_SIZE,    loop_defines.h:30 */
#define OSB_MAX_DIRS 4            /* MAX_DIRS for STEREO_FISHEYE, swarm_loop/src/swarm_loop.cpp:277-278 */
#define OSB_MAX_KPTS 200          /* superpoint_max_num, swarm_loop/launch/nodelet-sfisheye.launch:30 */
#define OSB_REMOTE_MAGIN_NUMBER 1000000  /* REMOTE_MAGIN_NUMBER, swarm_loop/include/swarm_loop/loop_detector.h:22 */

const char* osb_last_error(void);          /* thread-local text of the last failure */
const char* osb_version(void);
int osb_device_count(void);                /* 0 when no CUDA device is visible */

/* ------------------------------------------------------------------------------------------------------------
 * SuperPoint  -- replaces class SuperPointTensorRT (swarm_loop/include/swarm_loop/superpoint_tensorrt.h:20-28,
 *               constructed at swarm_loop/src/loop_cam.cpp:26, called at loop_cam.cpp:542).
 * `weights`: float32 blob, for each layer of swarm_loop/superpoint.ipynb:143-158 in definition order
 *            (conv1a,conv1b,conv2a,conv2b,conv3a,conv3b,conv4a,conv4b,convPa,convPb,convDa,convDb):
 *            weight in PyTorch OIHW order, then bias.  1 300 865 floats.  (Replaces the .trt engine path:
 *            TensorRT engine binaries are device-specific and unusable on B200.)
 * `pca_comp` [64][256] row-major = components_.csv, `pca_mean` [256] = mean_.csv (superpoint_tensorrt.cpp:110-111).
 * infer():   images  [batch][height][width] uint8 (the reference asserts the size, superpoint_tensorrt.cpp:122)
 *            n_kpts  [batch]                          number of keypoints per image (<= max_num)
 *            kpts    [batch][max_num][2] float (x,y)  ordered by descending confidence (NMS2, :304-308)
 *            desc    [batch][max_num][64] float       PCA-projected descriptors (:221)
 *            rows >= n_kpts[b] are left untouched.
 * -----------------------------------------------------------------------------------------------------------*/
typedef struct osb_superpoint osb_superpoint;
osb_status osb_superpoint_create(osb_superpoint** out, const float* weights, size_t n_weights, int width,
                                 int height, float thres, int max_num, const float* pca_comp,
                                 const float* pca_mean, int max_batch);
osb_status osb_superpoint_destroy(osb_superpoint* h);
osb_status osb_superpoint_infer(osb_superpoint* h, const uint8_t* images, int batch, int32_t* n_kpts,
                                float* kpts, float* desc);
osb_status osb_superpoint_infer_dev(osb_superpoint* h, const uint8_t* images_dev, int batch, int32_t* n_kpts_dev,
                                    float* kpts_dev, float* desc_dev, void* stream);
/* Stage-wise parity hooks (tests only; not used by the reference call sites):
 *  set_heatmap: upload a caller-supplied `semi` [batch][H][W] and `desc` [batch][256][H/8][W/8] (the two engine
 *               outputs, superpoint_tensorrt.cpp:139-140) and run ONLY getKeyPoints+NMS2+computeDescriptors.
 *  read:        copy an intermediate of the last infer() back: what = 0 semi [H][W], 1 desc [256][H/8][W/8],
 *               2 confidences of the returned keypoints [max_num], 3 NMS survivor plane as float [H][W],
 *               4 keypoint kernel counters [8]: candidates, survivors, NMS rounds, 0, SM cycles of its 4 phases,
 *               5 cycle counters [16] of the fused conv1a+conv1b kernel's CTA 0 from the last infer() with profiling on
 *                 (producer wait / compute / wait, loop total, MMA wait TMEM / wait tile / issue, epilogue wait / work, tiles). */
osb_status osb_superpoint_postprocess(osb_superpoint* h, const float* semi, const float* desc_nchw, int batch,
                                      int32_t* n_kpts, float* kpts, float* desc);
osb_status osb_superpoint_read(osb_superpoint* h, int what, int image, float* out, size_t n_floats);
/* per-layer device time of the last infer() (bench / profiling aid; tensor-core path only): enable, run infer(), then
 * layer_ms fills ms[0..11] = conv1a, conv1b(+pool), conv2a, conv2b(+pool), conv3a, conv3b(+pool), conv4a, conv4b, convPa,
 * convPb, convDa, convDb in milliseconds (CUDA events on the handle's stream). */
osb_status osb_superpoint_set_profiling(osb_superpoint* h, int enable);
osb_status osb_superpoint_layer_ms(osb_superpoint* h, float* ms, int n);

/* ------------------------------------------------------------------------------------------------------------
 * NetVLAD global descriptor -- replaces class MobileNetVLADTensorRT
 *   (swarm_loop/include/swarm_loop/mobilenetvlad_tensorrt.h:10-21, loop_cam.cpp:27 ctor, loop_cam.cpp:554 call).
 * images [batch][H][W] uint8 (converted to float UNSCALED, mobilenetvlad_tensorrt.cpp:8-10) -> out [batch][4096].
 * The hfnet MobileNetVLAD architecture is not part of the reference; DESIGN.md pins the stand-in whose weight
 * blob layout is omniswarm_b200/synth.py::netvlad_layer_table().
 * -----------------------------------------------------------------------------------------------------------*/
typedef struct osb_netvlad osb_netvlad;
osb_status osb_netvlad_create(osb_netvlad** out, const float* weights, size_t n_weights, int width, int height,
                              int max_batch);
osb_status osb_netvlad_destroy(osb_netvlad* h);
osb_status osb_netvlad_infer(osb_netvlad* h, const uint8_t* images, int batch, float* out);
osb_status osb_netvlad_infer_dev(osb_netvlad* h, const uint8_t* images_dev, int batch, float* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Keyframe database -- replaces faiss::IndexFlatIP(4096) (swarm_loop/include/swarm_loop/loop_detector.h:27-29;
 *   add: loop_detector.cpp:166,169; search: loop_detector.cpp:213; ntotal: loop_detector.cpp:291).
 * Exact float32 inner product, top-k by descending score, ties by ascending row id, ids = -1 and
 * scores = -inf where fewer than k rows exist.  Rows live in HBM, row-major [capacity][dim].
 * -----------------------------------------------------------------------------------------------------------*/
typedef struct osb_db osb_db;
osb_status osb_db_create(osb_db** out, int dim, int64_t capacity);
osb_status osb_db_destroy(osb_db* h);
osb_status osb_db_add(osb_db* h, int64_t n, const float* x, int64_t* first_id);
osb_status osb_db_add_dev(osb_db* h, int64_t n, const float* x_dev, int64_t* first_id, void* stream);
osb_status osb_db_search(osb_db* h, int64_t nq, const float* q, int k, float* scores, int64_t* ids);
osb_status osb_db_search_dev(osb_db* h, int64_t nq, const float* q_dev, int k, float* scores_dev,
                             int64_t* ids_dev, void* stream);
int64_t osb_db_size(osb_db* h);
osb_status osb_db_reset(osb_db* h);        /* ntotal = 0 (rows stay allocated) */
/* Row-sharded search (alternative for the 50 k-row sweep): each GPU scans its shard with
 * osb_db_search_dev, the per-shard top-k lists are all-gathered, and this call merges n_lists lists of k candidates per
 * query (cand_* are [nq][n_lists][k], ids < 0 = padding) into the global top-k with faiss::IndexFlatIP's order
 * (loop_detector.cpp:213): score descending, ties by ascending id.  id_offset_dev[l] (may be null) is added IN PLACE to the
 * ids of list l (shard-local row -> global row). */
osb_status osb_topk_merge_dev(int nq, int n_lists, int k, const float* cand_scores_dev, int64_t* cand_ids_dev,
                              const int64_t* id_offset_dev, float* scores_dev, int64_t* ids_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Local descriptor matcher -- replaces cv::BFMatcher(cv::NORM_L2, crossCheck = true).match(query, train)
 *   (swarm_loop/src/loop_cam.cpp:147-150 stereo up/down; swarm_loop/src/loop_detector.cpp:564-567 loop pairs).
 * For every query row the nearest train row by L2 (first minimum wins), kept only if mutual; output sorted by
 * query index.  Batched over `n_pairs` independent (query, train) pairs:
 *   q [n_pairs][max_n][dim], t [n_pairs][max_n][dim], nq/nt [n_pairs] valid row counts,
 *   out: qi/ti/dist [n_pairs][max_n], n_out [n_pairs].   max_n <= 256, dim == 64.
 * -----------------------------------------------------------------------------------------------------------*/
typedef struct osb_matcher osb_matcher;
osb_status osb_matcher_create(osb_matcher** out, int max_pairs, int max_n, int dim);
osb_status osb_matcher_destroy(osb_matcher* h);
osb_status osb_matcher_match(osb_matcher* h, int n_pairs, const float* q, const int32_t* nq, const float* t,
                             const int32_t* nt, int32_t* qi, int32_t* ti, float* dist, int32_t* n_out);
osb_status osb_matcher_match_dev(osb_matcher* h, int n_pairs, const float* q_dev, const int32_t* nq_dev,
                                 const float* t_dev, const int32_t* nt_dev, int32_t* qi_dev, int32_t* ti_dev,
                                 float* dist_dev, int32_t* n_out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Pose-graph solve -- replaces the body of SwarmLocalizationSolver::solve_once
 *   (swarm_localization/src/swarm_localization_solver.cpp:1668-1725): the three setup_problem_with_* walks
 *   (:1064-1214) become a flat factor list built by the adapter, ceres::Solve (:1712) becomes a GPU
 *   Levenberg-Marquardt with a block-Jacobi preconditioned CG on the 4x4-block normal equations.
 * poses   [n_nodes][4] double (x,y,z,yaw) in/out  -- the 4-vectors of EstimatePoses (swarm_localization_solver.hpp:46-50)
 * fixed   [n_nodes] 1 = SetParameterBlockConstant (solver.cpp:1196-1206)
 * type/ia/ib [n_factors]: factor kind and the two pose blocks
 * huber   [n_factors] 1 = ceres::HuberLoss(1.0) (solver.cpp:1077-1082,1138-1141), 0 = no loss (ego motion :1178)
 * payload [n_factors][OSB_PAYLOAD_LEN] double:
 *   OSB_FACTOR_DISTANCE  (DistanceMeasurementFactor, swarm_localization_factors.hpp:203-224): [d, sqrt_inf]
 *   OSB_FACTOR_RELPOSE   (RelativePoseFactor4d, :226-271): [meas x,y,z,yaw, sqrt_inf 4x4 row-major]
 *   OSB_FACTOR_DETECTION (DroneDetection4dFactor, :273-367): [dir 3, tan_base 2x3, inv_dep, flags, extrinsic_z,
 *                         dposea 4, dposeb 4, DETECTION_SPHERE_STD, DETECTION_INV_DEP_STD]; flags bit0 enable_depth,
 *                         bit1 enable_dpose
 * -----------------------------------------------------------------------------------------------------------*/
#define OSB_FACTOR_DISTANCE 0
#define OSB_FACTOR_RELPOSE 1
#define OSB_FACTOR_DETECTION 2
#define OSB_PAYLOAD_LEN 24

typedef struct {
  int32_t max_iterations;        /* ceres max_num_iterations = 1000 (solver.cpp:1697) */
  int32_t max_pcg_iterations;    /* per LM step */
  double max_time_s;             /* max_solver_time_in_seconds (solver.cpp:1702); <= 0 disables */
  double function_tolerance;     /* Ceres default 1e-6 */
  double gradient_tolerance;     /* Ceres default 1e-10 */
  double parameter_tolerance;    /* Ceres default 1e-8 */
  double pcg_tolerance;          /* relative residual of the inner solve */
  double initial_trust_radius;   /* Ceres default 1e4 */
  int32_t preconditioner;        /* OSB_PRECOND_AUTO (chain block-tridiagonal when the graph fits the one-cluster fast
                                    path, else block-Jacobi) or OSB_PRECOND_BLOCK_JACOBI */
  int32_t inner_precision;       /* arithmetic INSIDE the PCG (Jacobian blocks, direction/residual vectors, preconditioner):
                                    OSB_INNER_AUTO = fp32 when pcg_tolerance >= 1e-4 (LM only needs an inexact step), else
                                    fp64.  Residuals, costs, gradient, poses and all LM decisions are always fp64. */
} osb_solve_options;
#define OSB_PRECOND_AUTO 0
#define OSB_PRECOND_BLOCK_JACOBI 1
#define OSB_INNER_AUTO 0
#define OSB_INNER_FP64 1
#define OSB_INNER_FP32 2

typedef struct {
  double initial_cost;           /* 1/2 sum rho(|r|^2), as ceres Summary::initial_cost */
  double final_cost;             /* Summary::final_cost (solver.cpp:1721) */
  double solve_ms;               /* device time of the solve kernel(s), CUDA events */
  int32_t iterations;            /* LM iterations taken (accepted + rejected) */
  int32_t pcg_iterations;        /* total inner iterations */
  int32_t n_residuals;           /* problem.NumResiduals() (solver.cpp:1724) */
  int32_t termination;           /* 0 converged (function tol), 1 gradient tol, 2 parameter tol, 3 max iterations,
                                    4 time limit, 5 failure (non-finite cost) */
} osb_solve_summary;

typedef struct osb_solver osb_solver;
void osb_solve_default_options(osb_solve_options* o);
osb_status osb_solver_create(osb_solver** out, int max_nodes, int max_factors);
osb_status osb_solver_destroy(osb_solver* h);
osb_status osb_solver_solve(osb_solver* h, int n_nodes, double* poses, const uint8_t* fixed, int n_factors,
                            const int32_t* type, const int32_t* ia, const int32_t* ib, const double* payload,
                            const uint8_t* huber, const osb_solve_options* opt, osb_solve_summary* summary);
/* Resident graph: instead of re-flattening the whole window for every solve (the reference rebuilds its
 * ceres::Problem each time: setup_problem_with_sferror / _loops_and_detections / _ego_motion,
 * swarm_localization_solver.cpp:1064-1214), the adapter appends what add_new_swarm_frame / add_new_loop_connection
 * (swarm_localization_solver.hpp:197-214) bring.  The factor list stays in device memory, only new factors cross PCIe,
 * and the poses persist between solves like the reference's est_poses.  Node ids are append order; drop_oldest renumbers
 * (sliding window, solver.cpp:186-202).  osb_solver_solve and the resident graph can be mixed on one handle. */
osb_status osb_solver_graph_clear(osb_solver* h);
osb_status osb_solver_graph_add_nodes(osb_solver* h, int n, const double* poses /*[n][4]*/, const uint8_t* fixed /*[n] or NULL*/,
                                      int32_t* first_id);
osb_status osb_solver_graph_add_factors(osb_solver* h, int m, const int32_t* type, const int32_t* ia, const int32_t* ib,
                                        const double* payload, const uint8_t* huber);
osb_status osb_solver_graph_set_fixed(osb_solver* h, int node, int fixed);
osb_status osb_solver_graph_set_poses(osb_solver* h, int first, int n, const double* poses);
osb_status osb_solver_graph_get_poses(osb_solver* h, int first, int n, double* poses);
osb_status osb_solver_graph_size(osb_solver* h, int32_t* n_nodes, int32_t* n_factors);
osb_status osb_solver_graph_drop_oldest(osb_solver* h, int n_nodes);
osb_status osb_solver_solve_resident(osb_solver* h, const osb_solve_options* opt, osb_solve_summary* summary);
/* profiling aid: SM-clock cycles block 0 spent in the phases of the LAST solve, summed over its CG iterations:
 * out[0] factor phase, [1] barrier after it, [2] node phase 1, [3] reduction 1, [4] node phase 2, [5] reduction 2,
 * [6] number of CG iterations, [7] whole kernel; [8] CTAs, [9] 1 = one thread-block cluster (hardware barrier) /
 * 0 = cooperative grid, [10] bit 0 = Jacobians in shared memory, bit 1 = chain preconditioner, bit 2 = fp32 inner
 * arithmetic, [11] threads per CTA. */
osb_status osb_solver_phase_cycles(osb_solver* h, double* out12);
/* profiling aid: SM-clock cycles each warp spent in the chain-preconditioner sweeps of the LAST solve, [16 CTAs][8 warps] */
osb_status osb_solver_chain_cycles(osb_solver* h, double* out128);
/* host-only (no GPU needed): the node numbering the solver uses for its chain preconditioner -- a greedy maximum-weight
 * path cover of the factor graph (on a swarm graph: every drone's odometry chain).  order_out[i] = caller's node id of
 * internal node i; link_out[i] = 1 iff internal node i-1 precedes i on its path and i % 16 != 0. */
osb_status osb_solver_chain_plan(int n_nodes, const uint8_t* fixed, int n_factors, const int32_t* type,
                                 const int32_t* ia, const int32_t* ib, const double* payload, int32_t* order_out,
                                 uint8_t* link_out);
/* residual + analytic Jacobian of every factor at `poses` (parity hook for the factor kernels):
 * r [n_factors][4], Ja/Jb [n_factors][4][4] (rows >= the factor's residual count are zero), un-robustified. */
osb_status osb_solver_linearize(osb_solver* h, int n_nodes, const double* poses, int n_factors,
                                const int32_t* type, const int32_t* ia, const int32_t* ib, const double* payload,
                                double* r, double* Ja, double* Jb);

/* ------------------------------------------------------------------------------------------------------------
 * PCM outlier rejection of loop edges -- replaces SwarmLocalOutlierRejection::OutlierRejectionLoopEdgesPCM
 *   (swarm_localization/src/swarm_outlier_rejection/swarm_outlier_rejection.cpp:173-297), the stage that feeds
 *   get_good_loops() to the solve: pairwise consistency of the n loop edges of one drone pair
 *   (err = odom_a * p_edge2 * odom_b^-1 * p_edge1^-1, 6-D log map, squared Mahalanobis distance < pcm_thres, :190-235) and
 *   FMC::maxCliqueHeu on the consistency graph (third_party/fast_max-clique_finder/src/findCliqueHeu.cpp:120-244, restated
 *   literally incl. its prunings and candidate order).
 * osb_loop_edge: what the adapter reads off a Swarm::LoopEdge and the two DroneTrajectory objects (swarm_msgs is not in the
 *   reference tree, so its arithmetic is defined in oracle/pcm_ref.py): relative_pose as (x y z, qw qx qy qz),
 *   get_covariance() 6x6 row-major (translation block first), the ego-motion pose of drone id_a at ts_a and of drone id_b at
 *   ts_b, and the accumulated trajectory length at those stamps.  The odometry between two stamps is pose(ts1)^-1 * pose(ts2)
 *   with covariance |len2 - len1| * diag(odom_pos_cov_per_m x3, odom_ang_cov_per_m x3).
 * Edges are in insertion order (all_loops[id_a][id_b]); edges of another drone pair are never consistent.
 * Outputs: clique [n] = indices of the kept loops in maxCliqueHeu's order (good_loops_set, :291-296), clique_size; optional
 *   adj [n][n] (1 = consistent) and smd [n][n] (squared Mahalanobis distances, +inf where undefined).  n <= 4096. */
typedef struct {
  int32_t id_a, id_b;
  double rel_pose[7];
  double cov[36];
  double odom_a[7];
  double odom_b[7];
  double len_a, len_b;
} osb_loop_edge;
osb_status osb_pcm(const osb_loop_edge* edges, int n, double pcm_thres, double odom_pos_cov_per_m, double odom_ang_cov_per_m,
                   int32_t* clique, int32_t* clique_size, uint8_t* adj, double* smd);
osb_status osb_pcm_dev(const osb_loop_edge* edges_dev, int n, double pcm_thres, double odom_pos_cov_per_m,
                       double odom_ang_cov_per_m, int32_t* clique_dev, int32_t* clique_size_dev, uint8_t* adj_dev,
                       double* smd_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Geometric filter of the loop matcher (first half) -- the inlier mask of
 *   cv::findHomography(old_2d, new_2d, CV_RANSAC, 3, mask)            swarm_loop/src/loop_detector.cpp:589-598
 * for n_pairs correspondence sets at once.  src = old_2d, dst = new_2d, [n_pairs][max_n][2] floats, n[pair] points each
 * (max_n <= 256).  mask [n_pairs][max_n] (1 = inlier), n_inliers [n_pairs], winner [n_pairs] (hypothesis index, -1 if
 * none; may be NULL in the host variant).  Fewer than 4 points: empty mask (the reference rejects the pair, :598-600).
 * OpenCV's RANSAC is randomised; this one is deterministic in (points, seed): 512 hypotheses drawn by a counter-based
 * hash, same 4-point model / error / threshold rule, first best hypothesis wins (oracle/geometry_ref.py, pinned against
 * cv2 on well-separated data).  No process-wide state: the _dev entry point takes its scratch from the stream-ordered
 * allocator of `stream`, so concurrent callers on different streams are safe. */
osb_status osb_homography_ransac(const float* src, const float* dst, const int32_t* n, int n_pairs, int max_n,
                                 float thresh, uint32_t seed, uint8_t* mask, int32_t* n_inliers, int32_t* winner);
osb_status osb_homography_ransac_dev(const float* src_dev, const float* dst_dev, const int32_t* n_dev, int n_pairs,
                                     int max_n, float thresh, uint32_t seed, uint8_t* mask_dev, int32_t* n_inliers_dev,
                                     int32_t* winner_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Keypoints -> 3-D landmarks + landmarks_flag -- replaces the per-keypoint loops of
 *   LoopCam::generate_stereo_image_descriptor (swarm_loop/src/loop_cam.cpp:393-432: liftProjective, triangulatePoint :73-106,
 *   err <= TRIANGLE_THRES, in front of the up camera) and generate_gray_depth_image_descriptor (:276-302: depth look-up in
 *   mm, DEPTH_NEAR_THRES < dep < DEPTH_FAR_THRES, lift through pose_cam).  Arrays are [n_dirs][max_n][...]; intrinsics =
 *   fx fy cx cy of the distortion-free flattened pinhole (HOST pointer in both variants); poses are 7 doubles
 *   (x y z, qw qx qy qz), already pose_drone * extrinsic; nothing is lifted when a direction has <= accept_min_3d_pts
 *   keypoints (:385-391, :267).  stereo_match[d][i] = index of the down keypoint matched to up keypoint i, or -1. */
osb_status osb_stereo_lift(const float* kp_up, const float* kp_down, const int32_t* stereo_match, const int32_t* n_up,
                           const int32_t* n_down, int n_dirs, int max_n, const double* intrinsics, const double* pose_up,
                           const double* pose_down, double triangle_thres, int accept_min_3d_pts, float* pts3d,
                           uint8_t* flag_up, uint8_t* flag_down);
osb_status osb_stereo_lift_dev(const float* kp_up_dev, const float* kp_down_dev, const int32_t* stereo_match_dev,
                               const int32_t* n_up_dev, const int32_t* n_down_dev, int n_dirs, int max_n,
                               const double* intrinsics, const double* pose_up_dev, const double* pose_down_dev,
                               double triangle_thres, int accept_min_3d_pts, float* pts3d_dev, uint8_t* flag_up_dev,
                               uint8_t* flag_down_dev, void* stream);
osb_status osb_depth_lift(const float* kp, const int32_t* n, int n_dirs, int max_n, const uint16_t* depth_mm, int height,
                          int width, const double* intrinsics, const double* pose_cam, double near_thres, double far_thres,
                          int accept_min_3d_pts, float* pts3d, uint8_t* flag);
osb_status osb_depth_lift_dev(const float* kp_dev, const int32_t* n_dev, int n_dirs, int max_n, const uint16_t* depth_mm_dev,
                              int height, int width, const double* intrinsics, const double* pose_cam_dev, double near_thres,
                              double far_thres, int accept_min_3d_pts, float* pts3d_dev, uint8_t* flag_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Relative pose of a loop candidate (second half) -- replaces LoopDetector::compute_relative_pose
 *   (swarm_loop/src/loop_detector.cpp:355-413: cv::solvePnPRansac(matched_3d_now, matched_2d_norm_old, K = I, iterations
 *   100 / 1000 in init mode, reprojection error 3), PnPRestoCamPose, DeltaPose, RPerror :338-351, pnp_result_verify :317-336)
 *   and LoopDetector::check_loop_odometry_consistency (:294-315): everything between the matched landmarks and the LoopEdge.
 * pts3d [n_cand][max_n][3] = matched_3d_now (landmarks in the NEW drone's body frame), pts2d [n_cand][max_n][2] =
 *   matched_2d_norm_old (normalised coordinates in the OLD camera), n [n_cand] valid counts, max_n <= 1024.
 * OpenCV's RANSAC is randomised; this one is deterministic in (points, seed, prior): hypothesis h fits 4 hashed
 *   correspondences by a fixed-schedule Levenberg-Marquardt started from `prior` (the odometry prediction the reference
 *   computes and leaves unused, initial_old_cam_pose :377-382), the first hypothesis with the most inliers
 *   (reprojection error^2 <= thresh^2, no cheirality test, as cv::projectPoints) wins, and the pose is the LM minimiser over
 *   its inliers -- the definition of solvePnPRansac's result (oracle/pnp_ref.py, pinned against cv2.solvePnPRansac).
 * Poses are 7 doubles (x y z, qw qx qy qz); Swarm::Pose / DeltaPose / quat2eulers are defined in oracle/pnp_ref.py
 *   (swarm_msgs is not in the reference tree). */
typedef struct {
  int32_t iterations;            /* 100, or 1000 in init mode (:385-391) */
  float reproj_thresh;           /* 3 (:393) */
  uint32_t seed;
  int32_t is_4dof;               /* DeltaPose(..., is_4dof) (:403) */
  int32_t min_loop_num;          /* MIN_LOOP_NUM, or INIT_MODE_MIN_LOOP_NUM in init mode (:328-332) */
  int32_t same_drone;            /* 1: drone_id_a == drone_id_b, run the odometry-consistency check (:294-298) */
  double rperr_thres;            /* RPERR_THRES */
  double accept_loop_yaw_rad;    /* ACCEPT_LOOP_YAW_RAD */
  double max_loop_dis;           /* MAX_LOOP_DIS */
  double odometry_consistency_threshold;
  double prior[7];               /* initial (R, t) guess, x_cam_old = R X_now + t: (drone_pose_now^-1 drone_pose_old extrinsic)^-1 */
  double extrinsic[7];           /* old camera in the old drone's body frame */
  double drone_pose_now[7];
  double drone_pose_old[7];
  double odom_rel[7];            /* same_drone: ego-motion between the two stamps (get_relative_pose_by_ts) */
  double odom_edge_cov[36];      /* same_drone: odometry covariance + edge covariance, 6x6 row-major (:303) */
} osb_pnp_params;
typedef struct {
  int32_t pnp_success;           /* a model with >= 4 inliers was found */
  int32_t n_inliers;
  int32_t winner;                /* winning hypothesis, -1 if none */
  int32_t verified;              /* pnp_result_verify (:317-336) */
  int32_t odometry_consistent;   /* check_loop_odometry_consistency (:294-315); 1 for inter-drone loops */
  int32_t reserved;
  double rperr;                  /* RPerror (:338-351) */
  double md;                     /* squared Mahalanobis distance of the odometry check */
  double pose_cam[7];            /* (t, q) of the PnP solution: x_cam_old = R X_now + t */
  double dp_old_to_new[4];       /* DP_old_to_new: x y z yaw -- the LoopEdge's relative pose */
} osb_pnp_result;
osb_status osb_pnp_ransac(const float* pts3d, const float* pts2d, const int32_t* n, int n_cand, int max_n,
                          const osb_pnp_params* params, uint8_t* mask /*[n_cand][max_n]*/, osb_pnp_result* results);
osb_status osb_pnp_ransac_dev(const float* pts3d_dev, const float* pts2d_dev, const int32_t* n_dev, int n_cand, int max_n,
                              const osb_pnp_params* params_dev, uint8_t* mask_dev, osb_pnp_result* results_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Keyframe front-end -- the per-keyframe pipeline of LoopCam::on_flattened_images (loop_cam.cpp:178-229 ->
 *   generate_stereo_image_descriptor :341-523) followed by LoopDetector::on_image_recv's database work
 *   (loop_detector.cpp:89-104,150-287) and compute_correspond_features' matcher (loop_detector.cpp:539-587),
 *   kept resident on the GPU.  Triangulation / PnP / homography RANSAC stay host code in the reference and
 *   are out of scope.
 *
 * One keyframe = n_dirs directions x (up image, down image) (STEREO_FISHEYE: n_dirs=4, 8 images).
 * osb_keyframe_record is the fixed-size record every drone contributes to the swarm-wide exchange
 * (replaces LoopNet::broadcast_fisheye_desc, swarm_loop/src/loop_net.cpp:20-120): on the 8-GPU box it is
 * the unit of ONE ncclAllGather.
 * -----------------------------------------------------------------------------------------------------------*/
typedef struct {
  int32_t drone_id;
  int32_t msg_id;
  int32_t n_dirs;
  int32_t reserved;
  int32_t n_kpts[OSB_MAX_DIRS];                                        /* landmark_num per direction (up image) */
  int32_t n_kpts_down[OSB_MAX_DIRS];
  float global_desc[OSB_MAX_DIRS][OSB_DEEP_DESC_SIZE];                 /* image_desc */
  float local_desc[OSB_MAX_DIRS][OSB_MAX_KPTS][OSB_FEATURE_DESC_SIZE]; /* feature_descriptor (up image) */
  float kpts[OSB_MAX_DIRS][OSB_MAX_KPTS][2];                           /* landmarks_2d (up image) */
  int32_t stereo_match[OSB_MAX_DIRS][OSB_MAX_KPTS];                    /* down-image keypoint index matched to each
                                                                          up keypoint by the cross-check matcher, or -1 */
  float landmarks_3d[OSB_MAX_DIRS][OSB_MAX_KPTS][3];                   /* landmarks_3d (loop_cam.cpp:405-432): triangulated
                                                                          in the drone's odometry frame; zero without flag */
  int32_t landmarks_flag[OSB_MAX_DIRS][OSB_MAX_KPTS];                  /* landmarks_flag.  With osb_frontend_set_cameras: 1
                                                                          iff the stereo pair triangulates (err <=
                                                                          TRIANGLE_THRES, in front of the camera, :423);
                                                                          without cameras: stereo_match >= 0, a superset */
} osb_keyframe_record;

typedef struct {
  int32_t hit_id;                /* database image id (remote ids + OSB_REMOTE_MAGIN_NUMBER) or -1 */
  int32_t hit_dir;               /* imgid2dir of the hit (direction_old) or -1 */
  float hit_score;               /* `distance` of query_from_database (-1 when untouched) */
  int32_t accepted;              /* id != -1 && distance > -1 (loop_detector.cpp:265) */
  int32_t swapped;               /* 1 when the hit is a remote keyframe and the query keyframe is our own: the reference
                                    then calls compute_loop(old, new) (loop_detector.cpp:113-118), i.e. the DATABASE
                                    frame is the matcher's query side ("new") and the current keyframe the train side */
  int32_t hit_msg_id;            /* msg_id of the matched keyframe = imgid2fisheye[best_image_id], the key of
                                    fisheyeframe_database (loop_detector.cpp:272-275,105-111); -1 without a hit, and for rows
                                    put in with osb_frontend_db_load */
  int32_t hit_drone_id;          /* drone_id of the matched keyframe (fisheyeframe_database[msg_id].drone_id); -1 without a hit */
  int32_t dir_new[OSB_MAX_DIRS]; /* direction pairing of compute_correspond_features (loop_detector.cpp:455-465), */
  int32_t dir_old[OSB_MAX_DIRS]; /* one slot per pair with landmarks on both sides, -1 = unused slot */
  int32_t n_matches[OSB_MAX_DIRS];                 /* cross-check matches new-vs-old per direction pair */
  int32_t match_new[OSB_MAX_DIRS][OSB_MAX_KPTS];   /* queryIdx */
  int32_t match_old[OSB_MAX_DIRS][OSB_MAX_KPTS];   /* trainIdx */
  /* geometric filter (osb_frontend_config::geometric_filter; loop_detector.cpp:569-598): per direction pair, the
   * matches whose NEW landmark has a 3-D flag (stereo_match >= 0) and that pass the homography-RANSAC mask of
   * findHomography(old_2d, new_2d, RANSAC, 3).  geo_valid = 0 when fewer than 4 flagged matches remained (the reference
   * returns false for the pair).  Untouched (zero) when the filter is off. */
  int32_t geo_valid[OSB_MAX_DIRS];
  int32_t n_geo[OSB_MAX_DIRS];
  int32_t geo_new[OSB_MAX_DIRS][OSB_MAX_KPTS];
  int32_t geo_old[OSB_MAX_DIRS][OSB_MAX_KPTS];
} osb_loop_result;

typedef struct osb_frontend osb_frontend;
typedef struct {
  int32_t width, height, n_dirs, max_num;
  float sp_thres;                /* superpoint_thres */
  int32_t self_id;
  int32_t db_capacity;           /* rows per database (local and remote) */
  double inner_product_thres;    /* INNER_PRODUCT_THRES (query_thres) */
  double init_mode_product_thres;/* INIT_MODE_PRODUCT_THRES */
  int32_t match_index_dist;      /* MATCH_INDEX_DIST */
  int32_t query_dir;             /* direction queried: 1 for STEREO_FISHEYE, 0 for pinhole (loop_detector.cpp:249-257) */
  int32_t zero_bottom_quarter;   /* 1 for STEREO_FISHEYE: blank rows [3H/4, H) of every image (loop_cam.cpp:535-538) */
  int32_t accept_min_3d_pts;     /* ACCEPT_MIN_3D_PTS: stereo match skipped when n_kpts <= this (loop_cam.cpp:385-391) */
  int32_t geometric_filter;      /* 1: run the 3-D-flag + homography-RANSAC filter of loop_detector.cpp:569-598 in query */
  int32_t ransac_seed;           /* seed of the deterministic RANSAC (osb_homography_ransac) */
} osb_frontend_config;

osb_status osb_frontend_create(osb_frontend** out, const osb_frontend_config* cfg, const float* sp_weights,
                               size_t n_sp_weights, const float* pca_comp, const float* pca_mean,
                               const float* nv_weights, size_t n_nv_weights);
osb_status osb_frontend_destroy(osb_frontend* h);
/* extract: images_up/down [n_dirs][H][W] uint8 HOST (pinned or pageable) -> record written to `record_dev`
 * (DEVICE pointer, sizeof(osb_keyframe_record)); no synchronisation.  msg_id is stored in the record. */
osb_status osb_frontend_extract(osb_frontend* h, const uint8_t* images_up, const uint8_t* images_down,
                                int32_t msg_id, osb_keyframe_record* record_dev, void* stream);
/* same with images already on the device */
osb_status osb_frontend_extract_dev(osb_frontend* h, const uint8_t* images_up_dev, const uint8_t* images_down_dev,
                                    int32_t msg_id, osb_keyframe_record* record_dev, void* stream);
/* ingest: add `n_records` records (DEVICE array, e.g. the all-gather output; records whose drone_id == self go to the
 * local database, others to the remote one; index `skip` is ignored, pass -1 to ingest all) -- add_to_database. */
osb_status osb_frontend_ingest(osb_frontend* h, const osb_keyframe_record* records_dev, int n_records, int skip,
                               void* stream);
/* the same for ONE record that the caller knows to be this drone's own (the record extract has just written: on_image_recv's
 * add_to_database, loop_detector.cpp:89).  The routing is still by drone_id on the device; what the promise buys is that the
 * host does not have to assume the record might have gone to the remote database, so a drone that has never received a
 * foreign keyframe does not scan an empty remote store on every query. */
osb_status osb_frontend_ingest_own(osb_frontend* h, const osb_keyframe_record* record_dev, void* stream);
/* query: run query_from_database for the keyframe in `record_dev` (must already be ingested if it is an own keyframe,
 * as on_image_recv does) and, on a hit, the per-direction cross-check match against the stored keyframe; the result
 * is written to `result_dev` (DEVICE).  init_mode / nonkeyframe as loop_detector.cpp:176. */
osb_status osb_frontend_query(osb_frontend* h, const osb_keyframe_record* record_dev, int init_mode,
                              int nonkeyframe, osb_loop_result* result_dev, void* stream);
/* the whole single-drone step with HOST buffers: extract + ingest(own) + query, one synchronisation at the end. */
osb_status osb_frontend_process(osb_frontend* h, const uint8_t* images_up, const uint8_t* images_down,
                                int32_t msg_id, osb_keyframe_record* record_host, osb_loop_result* result_host);
/* synchronise `stream` and refresh the host-side row counts (call once per keyframe round when driving the
 * extract / ingest / query stages separately, e.g. around the multi-GPU all-gather) */
osb_status osb_frontend_finish(osb_frontend* h, void* stream);
int64_t osb_frontend_db_size(osb_frontend* h, int remote);
osb_status osb_frontend_db_reset(osb_frontend* h);
/* bulk-load rows into a database without running the networks (benchmark / replay set-up): global descriptors
 * [n][4096] and optional local descriptors [n][max_num][64] + counts [n] (HOST). */
osb_status osb_frontend_db_load(osb_frontend* h, int remote, int64_t n, const float* global_desc,
                                const float* local_desc, const int32_t* n_kpts);
/* Stereo triangulation inside extract: with the cameras set, every keyframe's record carries the
 * reference's landmarks_3d / landmarks_flag (loop_cam.cpp:393-432) -- up/down keypoints lifted through the distortion-free
 * pinhole of the flattened images (intrinsics fx fy cx cy), triangulated between pose_drone * left_extrinsic[d] and
 * pose_drone * right_extrinsic[d] (7 doubles each: x y z, qw qx qy qz), kept iff err <= triangle_thres and the point is in
 * front of the up camera.  set_drone_pose gives the pose_drone of the NEXT extract (msg.pose_drone, :394). */
osb_status osb_frontend_set_cameras(osb_frontend* h, const double* intrinsics /*[4]*/, const double* left_extrinsics /*[n_dirs][7]*/,
                                    const double* right_extrinsics /*[n_dirs][7]*/, double triangle_thres);
osb_status osb_frontend_set_drone_pose(osb_frontend* h, const double* pose_drone /*[7]*/);
/* landmarks_2d [n][max_num][2] and stereo_match [n][max_num] (>= 0 <=> landmarks_flag) of rows loaded with
 * osb_frontend_db_load -- what the geometric filter reads when such a row is the loop hit */
osb_status osb_frontend_db_set_geometry(osb_frontend* h, int remote, int64_t first_row, int64_t n, const float* kpts,
                                        const int32_t* stereo_match);
/* stage timing (CUDA events on the caller's stream, recorded only while enabled).  After a synchronising call
 * (process / finish) stage_ms returns the device time of the LAST extract+ingest+query sequence:
 * [0] SuperPoint network  [1] keypoints + descriptors (NetVLAD runs concurrently on a second stream)
 * [2] NetVLAD remainder not hidden behind [1]  [3] stereo match + record pack
 * [4] add_to_database  [5] database scans (remote + local)  [6] acceptance rule + per-direction match  [7] unused */
osb_status osb_frontend_set_profiling(osb_frontend* h, int enable);
osb_status osb_frontend_stage_ms(osb_frontend* h, float* ms8);
/* ------------------------------------------------------------------------------------------------------------
 * Swarm-wide keyframe exchange -- replaces LoopNet::broadcast_fisheye_desc / image_desc_callback
 *   (swarm_loop/src/loop_net.cpp:20-120,142-172; called from swarm_loop/src/swarm_loop.cpp:167): the LCM UDP multicast of
 *   one header + one message per landmark becomes ONE ncclAllGather of the fixed-size osb_keyframe_record per keyframe
 *   round (drone = rank; NVLink on the 8-GPU box).  NCCL is opened with dlopen on first use (no link-time dependency).
 * unique_id: rank 0 creates the 128-byte communicator id and hands it to the other drones over any channel they already
 *            share (a ROS parameter, a file, the LCM channel); every rank then calls init with the same id.
 * exchange:  record_dev (this drone's record) -> gathered_dev[world] in rank order, enqueued on `stream`, no
 *            synchronisation.  osb_frontend_ingest(gathered_dev, world, ...) routes own / foreign records by drone_id.
 * exchange_async + wait: the reference's exchange is asynchronous (loop_net.cpp:142-172), so the collective may run on
 *            the handle's own stream behind an event of `stream` while the next keyframe is extracted; osb_swarm_wait
 *            makes `stream` wait for the LAST exchange_async.  One exchange in flight per handle; the caller
 *            double-buffers record_dev / gathered_dev.   world == 1: a device copy, no NCCL needed.
 *            Transport of exchange_async: when the ranks can map each other's memory (CUDA IPC, one node), every record is
 *            PUSHED into the peers' inboxes by the copy engines over NVLink and completion travels as 32-bit round stamps
 *            that the streams wait on (cuStreamWaitValue32) -- no kernel, no SM, nothing spinning while a peer is late;
 *            otherwise (or with OSB_SWARM_P2P=0) it is the ncclAllGather on the handle's stream.  Every rank must call
 *            exchange_async and wait the same number of times. */
#define OSB_SWARM_ID_BYTES 128
typedef struct osb_swarm osb_swarm;
osb_status osb_swarm_unique_id(uint8_t* id_out /*[OSB_SWARM_ID_BYTES]*/);
osb_status osb_swarm_init(osb_swarm** out, const uint8_t* id /*[OSB_SWARM_ID_BYTES], may be NULL when world == 1*/, int rank,
                          int world);
osb_status osb_swarm_destroy(osb_swarm* h);
osb_status osb_swarm_exchange(osb_swarm* h, const osb_keyframe_record* record_dev, osb_keyframe_record* gathered_dev,
                              void* stream);
osb_status osb_swarm_exchange_async(osb_swarm* h, const osb_keyframe_record* record_dev, osb_keyframe_record* gathered_dev,
                                    void* stream);
osb_status osb_swarm_wait(osb_swarm* h, void* stream);
int osb_swarm_transport(osb_swarm* h);     /* what exchange_async uses: 1 = peer-to-peer copy engines (CUDA IPC inboxes, stream
                                              stamps; no SM, no NCCL kernel), 0 = ncclAllGather on the side stream */
int osb_swarm_rank(osb_swarm* h);
int osb_swarm_world(osb_swarm* h);

/* The convolution kernels are persistent (one CTA per SM, statically strided tiles): they run at full speed only when all of
 * their CTAs are resident.  A host that keeps another kernel on the GPU beside the front-end -- the pose-graph solve holds a
 * 16-CTA cluster for milliseconds -- caps them at n_sms (process-wide; 0 = every SM). */
void osb_set_sm_budget(int n_sms);
/* number of kernels launched by this library since it was loaded (all handles), for bench.py's gpu_launches */
int64_t osb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* OMNISWARM_B200_H */
