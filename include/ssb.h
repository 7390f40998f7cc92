/* ssb.h — C-ABI of the B200-native backend for the two numeric hot paths of
 * hridaybavle/semantic_slam.
 *
 * Plain C: opaque handles, plain pointers and sizes, int status codes (0/positive = ok, negative =
 * error, text via ssb_last_error()).  No torch / Eigen / g2o / PCL types cross this boundary.
 * Every entry point cites the reference interface (file:line under /root/reference) it replaces.
 * Matrices are row-major.  SE3 values are 3x4 [R|t] (12 doubles), i.e. the top three rows of the
 * Eigen::Isometry3d the reference passes.  A handle is not thread-safe (the reference drives these
 * calls from a single thread, src/semantic_graph_SLAM_node.cpp:14-20); distinct handles are
 * independent.
 */
#ifndef SSB_H
#define SSB_H

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_OK 0
#define SSB_ERR_INVALID (-1)  /* bad handle / id / argument            */
#define SSB_ERR_CUDA (-2)     /* CUDA runtime error (see ssb_last_error) */
#define SSB_ERR_NUMERIC (-3)  /* non-finite data reached the solver    */
#define SSB_ERR_COMM (-4)     /* NCCL / multi-GPU set-up error         */

/* ------------------------------------------------------------------------------------------- */
/* Path (1): ps_graph_slam::GraphSLAM  (include/ps_graph_slam/graph_slam.hpp:27-150,           */
/*           src/ps_graph_slam/graph_slam.cpp:40-239) — replaces g2o::SparseOptimizer "lm_var". */
/* ------------------------------------------------------------------------------------------- */
typedef struct ssb_graph ssb_graph;

typedef struct ssb_graph_opts {
  int device;           /* CUDA device ordinal, -1 = current device                               */
  int verbose;          /* graph_slam.cpp:41 verbose_                                             */
  int max_pcg_iters;    /* cap on PCG iterations per damped solve (default 20000)                 */
  double pcg_tol;       /* stop when sqrt(r'M^-1 r) <= pcg_tol * sqrt(r0'M^-1 r0) (default 1e-8)   */
  int preconditioner;   /* 0 = block-Jacobi on the Schur complement; 1 = + rigid-body coarse level (one
                           aggregate per CTA); 2 = + a middle level of 5-pose aggregates (3-level additive);
                           3 = as 2 with the 5-pose aggregates of a CTA coupled exactly inside two groups (on-chip kernel) */
  int coarse_group;     /* poses per coarse aggregate when preconditioner == 1 (default 32)        */
  int reserved[4];      /* [0] = 1: force the streaming PCG kernel; [1] = n: re-invert the coarse matrix every n-th solve;
                           [2]: internal (CTAs per rank of a sharded graph, set by ssb_graph_attach_local);
                           [3]: initial launch counter of the cell tags (test hook for the tag wrap-around) */
} ssb_graph_opts;

typedef struct ssb_lm_stats {
  int iterations;        /* outer LM iterations performed (SparseOptimizer::optimize return)      */
  int terminated;        /* 1 if the algorithm returned Terminate (10 failed trials or rho == 0)   */
  int total_trials;      /* damped solves performed                                               */
  int total_pcg_iters;   /* PCG iterations summed over all solves                                 */
  double chi2_initial;   /* graph->chi2() before (graph_slam.cpp:202)                             */
  double chi2_final;     /* graph->chi2() after  (graph_slam.cpp:212)                             */
  double lambda_final;
  double ms_prepare;     /* host CSR build + H2D upload                                           */
  double ms_device;      /* CUDA-event time of the whole LM loop on the device                    */
  double ms_total;       /* wall clock of the call                                                */
  long long kernel_launches; /* kernels launched by this call                                     */
  double ms_pcg;         /* CUDA-event time spent inside the PCG kernel (sum over trials)         */
} ssb_lm_stats;

/* fills `o` with the defaults */
void ssb_graph_default_opts(ssb_graph_opts* o);

/* GraphSLAM::GraphSLAM(bool verbose)  graph_slam.cpp:40-97  (solver "lm_var", ParameterSE3Offset id 0
 * = identity).  opts may be NULL.  Returns NULL on failure. */
ssb_graph* ssb_graph_create(const ssb_graph_opts* opts);
/* GraphSLAM::~GraphSLAM  graph_slam.cpp:102 */
void ssb_graph_destroy(ssb_graph* g);

/* GraphSLAM::add_se3_node  graph_slam.cpp:104-115.  Vertex id = number of vertices so far; the first
 * vertex ever added is fixed (:109-111).  Returns the vertex id (>= 0) or an error. */
int ssb_graph_add_se3_node(ssb_graph* g, const double T34[12]);
/* GraphSLAM::add_point_xyz_node  graph_slam.cpp:127-134 */
int ssb_graph_add_point_xyz_node(ssb_graph* g, const double xyz[3]);
/* GraphSLAM::add_se3_edge  graph_slam.cpp:136-148 (g2o::EdgeSE3, measurement = relative pose,
 * information 6x6).  Returns the edge id. */
int ssb_graph_add_se3_edge(ssb_graph* g, int v1, int v2, const double Z34[12], const double info[36]);
/* GraphSLAM::add_se3_point_xyz_edge  graph_slam.cpp:150-166 (g2o::EdgeSE3PointXYZ, parameter id 0,
 * no robust kernel — the reference passes an uninitialised pointer). */
int ssb_graph_add_se3_point_xyz_edge(ssb_graph* g, int v_se3, int v_xyz, const double xyz[3], const double info[9]);
/* GraphSLAM::add_point_xyz_point_xyz_edge  graph_slam.cpp:168-180 (g2o::EdgePointXYZ; never called by
 * the reference).  Accepted and stored; ssb_graph_optimize reports SSB_ERR_INVALID while such an edge
 * exists (not supported by the Schur back-end in this round). */
int ssb_graph_add_point_xyz_point_xyz_edge(ssb_graph* g, int v1, int v2, const double xyz[3], const double info[9]);

/* Plane landmarks: the API the reference keeps commented out — GraphSLAM::add_plane_node  graph_slam.hpp:44 /
 * graph_slam.cpp:117-125 (g2o::VertexPlane, 3 DoF, estimate = 4 plane coefficients, normalised on entry like
 * g2o::Plane3D) and GraphSLAM::add_se3_plane_edge  graph_slam.hpp:74-75 with the reference's own edge type
 * g2o::EdgeSE3Plane  include/g2o/edge_se3_plane.hpp:8-48 (error = (X^-1 * plane).ominus(measurement) =
 * (d azimuth, d elevation, d distance), information 3x3).  g2o differentiates this edge numerically; this
 * back-end uses exact Jacobians.  Plane vertices are eliminated by the Schur step like point landmarks. */
int ssb_graph_add_plane_node(ssb_graph* g, const double coeffs[4]);
int ssb_graph_add_se3_plane_edge(ssb_graph* g, int v_se3, int v_plane, const double plane[4], const double info[9]);
int ssb_graph_get_plane(ssb_graph* g, int vid, double coeffs[4]);
int ssb_graph_set_plane(ssb_graph* g, int vid, const double coeffs[4]);

int ssb_graph_num_vertices(const ssb_graph* g); /* graph->vertices().size() */
int ssb_graph_num_edges(const ssb_graph* g);    /* graph->edges().size()    */

/* VertexSE3::estimate() / VertexPointXYZ::estimate()  (semantic_graph_slam.cpp:94-95, data_association.h:378) */
int ssb_graph_get_se3(ssb_graph* g, int vid, double T34[12]);
int ssb_graph_get_point_xyz(ssb_graph* g, int vid, double xyz[3]);
int ssb_graph_set_se3(ssb_graph* g, int vid, const double T34[12]);
int ssb_graph_set_point_xyz(ssb_graph* g, int vid, const double xyz[3]);
/* OptimizableGraph::Vertex::setFixed */
int ssb_graph_set_fixed(ssb_graph* g, int vid, int fixed);
/* OptimizableGraph::Vertex::hessianIndex(): position among the non-fixed vertices in id order, -1 if
 * fixed (semantic_graph_slam.cpp:188-190).  Valid after an optimize call. */
int ssb_graph_hessian_index(ssb_graph* g, int vid);
/* bulk read-back: all SE3 estimates (12 doubles each, id order) and all landmark estimates (3 each, id order;
 * for a plane vertex its unit normal — use ssb_graph_get_plane for all 4 coefficients) */
int ssb_graph_get_all(ssb_graph* g, double* se3_out, double* xyz_out);

/* SparseOptimizer::chi2() at the current estimates (graph_slam.cpp:202,212) */
int ssb_graph_chi2(ssb_graph* g, double* chi2_out);

/* GraphSLAM::optimize  graph_slam.cpp:182-219:  returns 0 and does nothing when the graph has fewer
 * than 10 edges (:184-186); otherwise initializeOptimization + optimize(max_iterations) (the reference
 * passes 1024, :205) and returns 1.  Negative = error.  stats may be NULL. */
int ssb_graph_optimize(ssb_graph* g, int max_iterations, ssb_lm_stats* stats);
/* Splits of ssb_graph_optimize used by the benchmark and by incremental callers:
 *   ssb_graph_prepare           : initializeOptimization analogue (graph_slam.cpp:199) — hessian indices, CSR
 *                                 edge tables, H2D upload of whatever changed on the host.
 *   ssb_graph_optimize_resident : the LM loop on the state already resident on the device (no upload, no
 *                                 read-back of the estimates); requires ssb_graph_prepare.
 *   ssb_graph_invalidate        : force the next prepare/optimize to rebuild and re-upload the edge tables.
 *   ssb_graph_set_all           : bulk counterpart of ssb_graph_get_all. */
int ssb_graph_prepare(ssb_graph* g);
int ssb_graph_optimize_resident(ssb_graph* g, int max_iterations, ssb_lm_stats* stats);
int ssb_graph_invalidate(ssb_graph* g);
int ssb_graph_set_all(ssb_graph* g, const double* se3_in, const double* xyz_in);
/* per-iteration records of the last optimize: 6 doubles each
 * (chi2 before, chi2 after, lambda after, rho of last trial, trials, pcg iterations). Returns count. */
int ssb_graph_get_history(ssb_graph* g, double* out6n, int cap);

/* GraphSLAM::computeLandmarkMarginals  graph_slam.cpp:221-234 <- getAndSetLandmarkCov
 * semantic_graph_slam.cpp:181-205: for each listed XYZ vertex the 3x3 block (H^-1)[v,v] of the last
 * built (undamped) system.  out9n: n row-major 3x3 blocks.  Returns 1 on success, 0 if unavailable.
 * One Schur-complement PCG solve per column; graphs that fill a fraction of the chip are laid out up to 16 times side by
 * side and solved 16 columns per launch (one CG recurrence per copy).  On a sharded graph every rank computes the blocks
 * on an unsharded copy of the graph on its own GPU from the gathered final estimates (identical bits on every rank). */
int ssb_graph_landmark_marginals(ssb_graph* g, const int* vids, int n, double* out9n);

/* GraphSLAM::save  graph_slam.cpp:236-239 (g2o text format: VERTEX_SE3:QUAT, VERTEX_TRACKXYZ,
 * VERTEX_PLANE, EDGE_SE3:QUAT, EDGE_SE3_TRACKXYZ, EDGE_SE3_PLANE, PARAMS_SE3OFFSET, FIX) and its inverse. */
int ssb_graph_save_g2o(ssb_graph* g, const char* path);
int ssb_graph_load_g2o(ssb_graph* g, const char* path);

/* test hook: device linearisation of edge `eid` at the current estimates.  err[D], Ji[D*di], Jj[D*dj]
 * row-major with (D,di,dj) = (6,6,6) for SE3 edges and (3,6,3) for SE3-XYZ and SE3-plane edges. */
int ssb_graph_edge_linearize(ssb_graph* g, int eid, double* err, double* Ji, double* Jj);
/* test hook: solve (H + lambda I) x = b once at the current linearisation with the Schur/PCG device
 * path; x is returned in hessian-index order (6 per SE3, 3 per XYZ vertex).  Returns PCG iterations. */
int ssb_graph_solve_once(ssb_graph* g, double lambda, double* x, int x_len);

/* One graph sharded over several GPUs by contiguous keyframe range (the reference re-optimises the
 * whole graph every tick, semantic_graph_slam.cpp:81, graph_slam.cpp:199-205, so graph length is the scaling axis).
 * Every rank replays the same add_* calls on its own handle; ssb_graph_prepare / optimize / chi2 / snapshot /
 * restore are then collective (every rank must call them in the same order).  Inside optimize each rank linearises,
 * solves and updates its own keyframe range; the kernels write boundary data straight into the neighbours' memory
 * over NVLink (csrc/ssb_peer.cuh) — the communicator is only used to hand over the memory handles.
 *   ssb_graph_attach_comm  : one process per GPU; `unique_id` = 128 bytes from ssb_comm_unique_id on rank 0,
 *                            broadcast by the host (NCCL, resolved at run time).  world == 1 detaches.
 *   ssb_graph_attach_local : the ranks are host threads of this process (one handle each, same or different
 *                            devices), paired through `group_key`.  cta_per_rank = 74 or 37 lets 2 or 4 shards share
 *                            ONE GPU ("virtual shards": the whole protocol on a single-GPU box); 0 = a GPU per rank.
 *   ssb_shard_ranges       : own keyframe range [out[0], out[1]) of `rank` (host-only).
 *   ssb_graph_shard_info   : the plan for the current graph (host-only): out[0..1] own keyframe range, out[2] local
 *                            keyframes (own + ghosts), out[3] owned / out[4] touched landmarks, out[5] local edges. */
int ssb_comm_unique_id(unsigned char id_out[128]);
int ssb_graph_attach_comm(ssb_graph* g, int rank, int world, const unsigned char unique_id[128]);
int ssb_graph_attach_local(ssb_graph* g, int rank, int world, const char* group_key, int cta_per_rank);
int ssb_shard_ranges(int n_poses, int n_landmarks, int world, int rank, int out4[4]);
int ssb_graph_shard_info(ssb_graph* g, int world, int rank, int out6[6]);
/* The same plan from bare index lists (host-only, no GPU, no handle).  pl_pose / pl_lm: keyframe and landmark index
 * (within their kind) of every pose-landmark edge; pp_i / pp_j: keyframes of every pose-pose edge.  For `rank`:
 * out[0..1] own keyframe range, out[2] local keyframes, out[3] owned / out[4] touched landmarks, out[5] local
 * pose-landmark edges, out[6] u pushes (keyframe, rank) and out[7] v pushes (landmark part, rank) per PCG iteration.
 * ghost_out (NULL or n_poses ints): the ghost keyframes; push_to (NULL or world ints): own keyframes pushed per rank. */
int ssb_shard_plan(int n_poses, int n_landmarks, const int* pl_pose, const int* pl_lm, int n_pl, const int* pp_i, const int* pp_j,
                   int n_pp, int world, int rank, int out8[8], int* ghost_out, int* push_to);

/* ------------------------------------------------------------------------------------------- */
/* Path (2): planar_segmentation RANSAC plane fit on bbox-cropped depth clouds                  */
/*   plane_segmentation::segmentPointCloudData  src/planar_segmentation/plane_segmentation.cpp:24-82 */
/*   plane_segmentation::compute2DConvexHull -> pcl::SACSegmentation::segment  :631-647          */
/* ------------------------------------------------------------------------------------------- */
typedef struct ssb_cloud_layout {     /* sensor_msgs::PointCloud2 fields used at :44-61 */
  int width, height;                  /* 640 x 480 (:35)                               */
  int point_step, row_step;           /* bytes                                         */
  int off_x, off_y, off_z, off_rgb;   /* fields[0..3].offset                           */
} ssb_cloud_layout;

typedef struct ssb_bbox {             /* msg/ObjectInfo.msg:3-6 */
  int tl_x, tl_y, width, height;
} ssb_bbox;

typedef struct ssb_ransac_opts {
  double threshold;      /* seg.setDistanceThreshold(0.01)  :645                                */
  int refine;            /* seg.setOptimizeCoefficients(true) :641                              */
  int mode;              /* 0 = fixed-K (score every hypothesis), 1 = PCL adaptive-k replay      */
  int max_iterations;    /* PCL default 50 (mode 1)                                             */
  double probability;    /* PCL default 0.99 (mode 1)                                           */
  int device;            /* CUDA device ordinal, -1 = current                                   */
  int reserved[3];
} ssb_ransac_opts;

typedef struct ssb_plane_result {
  int status;            /* 0 ok, 1 "spurious" bbox (:34-38), 2 no valid model                  */
  int n_points;          /* width*height of the crop                                            */
  int best_hyp;          /* winning hypothesis (first best, PCL keeps strictly-better only)     */
  int best_count;        /* its inlier count                                                    */
  int iterations;        /* hypotheses consumed                                                 */
  int refined_count;     /* inliers of the refined model (selectWithinDistance)                 */
  float coef[4];         /* winning 3-point model (a,b,c,d)                                     */
  float refined[4];      /* after optimizeModelCoefficients                                     */
  float centroid[3];     /* centroid of the winning model's inliers (what optimizeModelCoefficients
                            computes; zeros when refine is off or fewer than 4 inliers)             */
  int reserved;
} ssb_plane_result;

typedef struct ssb_ransac ssb_ransac;   /* persistent device buffers + stream */

void ssb_ransac_default_opts(ssb_ransac_opts* o);
ssb_ransac* ssb_ransac_create(int device);
void ssb_ransac_destroy(ssb_ransac* r);

/* Crops every bbox out of the PointCloud2 byte buffer (K6), scores `n_hyp` 3-point plane hypotheses
 * per crop given as index triples into the row-major crop (K7), refines the winner by PCA over its
 * inliers and re-selects inliers (K8).
 *   msg      : host pointer to PointCloud2.data (height*row_step bytes)
 *   triples  : host int32 [n_boxes][n_hyp][3]
 *   results  : host [n_boxes]
 *   counts   : host int32 [n_boxes][n_hyp] inlier count per hypothesis (-1 = not evaluated), may be NULL
 *   mask     : host uint8, concatenation over non-spurious boxes of the refined inlier mask, may be NULL
 * Returns SSB_OK or an error. */
int ssb_ransac_plane_batch(ssb_ransac* r, const void* msg, const ssb_cloud_layout* layout, const ssb_bbox* boxes,
                           int n_boxes, const int* triples, int n_hyp, const ssb_ransac_opts* opts,
                           ssb_plane_result* results, int* counts, unsigned char* mask);

/* Device-resident variant used for the `value` leg of the benchmark: cloud and triples are uploaded
 * once with ssb_ransac_upload, then ssb_ransac_run_resident re-runs crop + score + refine on the device
 * without host traffic (results stay on the device until ssb_ransac_fetch). */
int ssb_ransac_upload(ssb_ransac* r, const void* msg, const ssb_cloud_layout* layout, const ssb_bbox* boxes,
                      int n_boxes, const int* triples, int n_hyp, const ssb_ransac_opts* opts);
int ssb_ransac_run_resident(ssb_ransac* r);
int ssb_ransac_fetch(ssb_ransac* r, ssb_plane_result* results, int* counts, unsigned char* mask);
/* CUDA stream the handle launches on (for event timing by the caller), as an opaque pointer */
void* ssb_ransac_stream(ssb_ransac* r);
/* CUDA-event timing of the last run: out[0] = whole device pipeline (crop..finish) in ms, out[1] = the
 * point x hypothesis sweep kernel alone */
int ssb_ransac_timing(ssb_ransac* r, double out[2]);
/* number of kernels launched so far on this handle */
long long ssb_ransac_launch_count(ssb_ransac* r);

/* The 3-point sample stream pcl::RandomSampleConsensus draws for a model over `n_indices` points (host only, no GPU):
 * SampleConsensusModel's boost::mt19937 seeded with `seed` (PCL: 12345u unless `random`), read through
 * boost::uniform_int<>(0, INT_MAX) (= the 32-bit output >> 1), and drawIndexSample's partial Fisher-Yates shuffle
 * of `shuffled_indices_`, whose state carries over from one draw to the next
 * (pcl/sample_consensus/sac_model.h: drawIndexSample; called per RANSAC iteration through getSamples by
 * pcl::SACSegmentation::segment, plane_segmentation.cpp:647).  triples: int32 [n_draws][3], indices into the model's
 * point set (the row-major crop).  A draw PCL would reject (isSampleGood: collinear) stays in the stream; the adaptive
 * replay (ssb_ransac_opts.mode = 1) skips it exactly where PCL would draw again, so the evaluated hypotheses coincide.
 * Returns SSB_OK; n_indices < 3 gives a stream of zeros (PCL selects no sample). */
int ssb_ransac_pcl_samples(int n_indices, int n_draws, unsigned seed, int* triples);

/* plane_segmentation::segmentPointCloudData alone (K6): out = n x 4 floats (x,y,z,rgb), row-major
 * organised crop.  Returns n = width*height, or -1 for a spurious box. */
int ssb_crop_bbox(ssb_ransac* r, const void* msg, const ssb_cloud_layout* layout, const ssb_bbox* box, float* out);

/* ------------------------------------------------------------------------------------------- */
/* Per-frame landmark association (host step between the two hot paths)                          */
/*   data_association::find_matches / associate_lanmarks   include/ps_graph_slam/data_association.h:75-235 */
/*   semantic_tools::transformNormalsToWorld / dist        include/tools.h:18-135,293-297        */
/* Host code in the reference too (a few detections x a few hundred landmarks per frame); restated in   */
/* single precision with the reference's evaluation order and quirks so that indices are bit-exact.     */
/* ------------------------------------------------------------------------------------------- */
typedef struct ssb_assoc ssb_assoc;

typedef struct ssb_assoc_opts {      /* ros params of data_association::init  :43-67 */
  double maha_dist_thres;            /* 0.5  */
  double eq_dist_thres;              /* 1.21 */
  double land_noise_low;             /* 0.5  -> Q_ = land_noise_low * I3 */
  double land_noise_high;            /* 0.9 (unused by the live code) */
  int use_maha_dist;                 /* 1 */
  int use_eq_dist;                   /* 0 (the shipped yamls set use_maha_dist 0 / use_eq_dist 1) */
  int use_rtab_map_odom;             /* 0 */
  int strict;                        /* 0 = reproduce the stale distance_min / nearest id of :100-107;
                                        1 = reset them for every detection */
} ssb_assoc_opts;

typedef struct ssb_detection {       /* detected_object  include/planar_segmentation/detected_object.h:14-24 */
  int type;                          /* std::string type, as an integer class id */
  int plane_type;                    /* std::string plane_type ("horizontal"/"vertical"), as an integer id */
  float pose[3];                     /* centroid in the camera frame */
  float normal[4];                   /* normal_orientation */
} ssb_detection;

typedef struct ssb_landmark_obs {    /* landmark  include/ps_graph_slam/landmark.h:16-35 as returned per detection */
  int is_new_landmark;               /* 1: caller adds a VertexPointXYZ at `pose` (semantic_graph_slam.cpp:160-166) */
  int id;                            /* index into the mapped landmark list (-1: detection produced no landmark) */
  int type, plane_type;
  float pose[3];                     /* world frame */
  float local_pose[3];               /* robot frame: the measurement of the SE3-XYZ edge (:171-174) */
  float normal[4];                   /* world frame */
  float covariance[9];               /* = Q_ */
  double information[9];             /* covariance.inverse().cast<double>() (:170), row-major */
} ssb_landmark_obs;

typedef struct ssb_planar_region {   /* one pcl::PlanarRegion as consumed by multiPlaneSegmentation :160-255 */
  float centroid[3];                 /* regions[i].getCentroid(), camera frame */
  float model[4];                    /* regions[i].getCoefficients() */
  int contour_points;                /* regions[i].getContour().size() (gate > 100, :169) */
  float area;                        /* pcl::calculatePolygonArea(contour) (gate >= planar_area, :195) */
} ssb_planar_region;

typedef struct ssb_detected_object { /* detected_object  include/planar_segmentation/detected_object.h:14-24 */
  int type;
  int plane_type;                    /* 0 "horizontal", 1 "vertical" */
  float prob;
  float num_points;
  float pose[3];                     /* camera frame */
  float world_pose[3];
  float normal_orientation[4];
} ssb_detected_object;

/* plane_segmentation::multiPlaneSegmentation's region post-processing (src/planar_segmentation/plane_segmentation.cpp:
 * 117-132,160-255) + point_cloud_segmentation::segmentPlanarSurfaces (include/planar_segmentation/
 * point_cloud_segmentation.h:26-103): gates, horizontal/vertical classification against gravity, normal sign
 * conventions, camera -> world.  Host code (a handful of regions per frame).  Returns the number of objects written
 * to out[] (<= n) or an error. */
int ssb_segment_planar_surfaces(const ssb_planar_region* regions, int n, const float robot_pose[6], float cam_angle,
                                int object_type, float prob, float planar_area, ssb_detected_object* out);

/* The LIVE segmentation path: plane_segmentation::computeNormalsFromPointCloud (src/planar_segmentation/
 * plane_segmentation.cpp:84-106 -> pcl::IntegralImageNormalEstimation, COVARIANCE_MATRIX, depth-change factor 0.03,
 * smoothing 20) + plane_segmentation::multiPlaneSegmentation's PCL call (:136-156 -> pcl::OrganizedMultiPlaneSegmentation::
 * segmentAndRefine, 2 degrees / 0.02 m) + pcl::calculatePolygonArea (:189), for every bbox crop of a frame on the device
 * (csrc/ssb_organized.cuh).  Together with ssb_segment_planar_surfaces this is point_cloud_segmentation::
 * segmentallPointCloudData (include/planar_segmentation/point_cloud_segmentation.h:105-181) without the class filter. */
typedef struct ssb_organized_opts {
  float max_depth_change_factor;   /* 0.03  plane_segmentation.cpp:98 */
  float normal_smoothing_size;     /* 20    :99 */
  int min_inliers;                 /* num_point_seg (500; the shipped yamls use 100), :7,139 */
  float angular_threshold;         /* 0.017453 * 2 rad, :140 */
  float distance_threshold;        /* 0.02 m (depth dependent inside PCL), :141 */
  float maximum_curvature;         /* 0.001, PCL default */
  int norm_point_thres;            /* crops with fewer points are skipped (5000), :8,93 */
  int reserved[3];
} ssb_organized_opts;
void ssb_organized_default_opts(ssb_organized_opts* o);
/* regions: [n_boxes][max_regions] in PCL's order (by label); n_regions[b] = regions found (may exceed max_regions),
 * -1 = spurious bbox, -2 = crop skipped (below norm_point_thres).  Optional outputs (NULL to skip): n_inliers
 * [n_boxes][max_regions] (after the refinement); per point, concatenated over the non-spurious boxes in row-major crop
 * order: normals_out [][4] (nx, ny, nz, curvature; NaN where PCL leaves NaN), labels_out (after the refinement, -1 = none),
 * dist_out (the smoothing distance map).  At most 64 regions per crop are kept. */
int ssb_organized_planes(ssb_ransac* r, const void* msg, const ssb_cloud_layout* layout, const ssb_bbox* boxes, int n_boxes,
                         const ssb_organized_opts* opts, int max_regions, ssb_planar_region* regions, int* n_regions, int* n_inliers,
                         float* normals_out, int* labels_out, float* dist_out);
double ssb_organized_last_ms(ssb_ransac* r);

/* ------------------------------------------------------------------------------------------- */
/* The DORMANT plane-clustering chain (never called by the reference's live path)       */
/*   plane_segmentation::clusterAndSegmentAllPlanes   src/planar_segmentation/plane_segmentation.cpp:261-294   */
/*   computeKmeans -> cv::kmeans                      :525-535                                                  */
/*   compute2DConvexHull: SACSegmentation + pcl::ProjectInliers + pcl::ConvexHull   :631-664                    */
/* Device code: csrc/ssb_cluster.cuh.  Uses the buffers and the stream of an ssb_ransac handle.                   */
/* ------------------------------------------------------------------------------------------- */
/* plane_segmentation::computeKmeans (:525-535): cv::kmeans(points, K, labels, TermCriteria(EPS + ITER, max_count, epsilon),
 * attempts, cv::KMEANS_RANDOM_CENTERS, centroids) with OpenCV's arithmetic (labels and centres bit-identical to cv::kmeans of
 * OpenCV 4.x for the same RNG state).  data: host float [n][dims] (dims <= 4, K <= 8); rng_state: cv::theRNG().state before the
 * call, updated to its value after it (OpenCV's initial state is 0xffffffff); labels: host int [n]; centers: host float
 * [K][dims]; compactness: the return value of cv::kmeans.  Returns SSB_OK or an error (n < K is an error: cv::kmeans throws). */
int ssb_kmeans(ssb_ransac* r, const float* data, int n, int dims, int K, int max_count, double epsilon, int attempts,
               unsigned long long* rng_state, int* labels, float* centers, double* compactness);

/* pcl::ProjectInliers(SACMODEL_PLANE) + pcl::ConvexHull::reconstruct of compute2DConvexHull (:649-662): the points of pts4
 * (host float [n][4]) selected by mask (host uint8 [n]) are projected onto the plane coef (a, b, c, d) and the vertices of their
 * planar convex hull are returned in PCL's output order (decreasing angle about the hull's centroid): rows3 [max_rows][3] = the
 * projected hull points, src (optional) = their indices in pts4.  Returns the number of hull vertices (may exceed max_rows;
 * only max_rows are written) or a negative error; *n_inliers = points projected. */
int ssb_project_hull(ssb_ransac* r, const float* pts4, const unsigned char* mask, int n, const float coef[4], float* rows3, int* src,
                     int max_rows, int* n_inliers);

typedef struct ssb_cluster_opts {
  int num_centroids_normals;     /* 4     include/planar_segmentation/plane_segmentation.h:41 */
  int num_centroids_distance;    /* 2     :42 */
  int kmeans_attempts;           /* 10    plane_segmentation.cpp:531 */
  int kmeans_max_count;          /* 10    :529 */
  double kmeans_epsilon;         /* 0.01  :530 */
  int min_cluster_points;        /* 500: a distance cluster is kept when it has MORE points (:419) */
  float centroid_tolerance;      /* 0.3: filterCentroids keeps normals within +-0.3 per component of the horizontal normal (:511-516) */
  int ransac_hypotheses;         /* 0 = PCL's adaptive stopping rule on a 512-sample stream; > 0 = score that many */
  unsigned ransac_seed;          /* 12345: seed of PCL's sample stream (ssb_ransac_pcl_samples), restarted for every cluster
                                    like the pcl::SACSegmentation object compute2DConvexHull builds per call (:637) */
  int reserved[4];
} ssb_cluster_opts;

typedef struct ssb_plane_cluster {   /* one entry of final_normals_with_distances + its hull (:399-425, :431-477) */
  float normal[3];               /* the (filtered) k-means centroid of the normals */
  float distance;                /* the k-means centroid of the signed distances */
  int normal_label, distance_label;
  int n_points;                  /* points of the cluster (> min_cluster_points) */
  int n_inliers;                 /* RANSAC inliers that were projected onto the plane */
  float coef[4];                 /* the refined plane of compute2DConvexHull's SACSegmentation */
  int row0, n_rows;              /* its rows in rows8 */
} ssb_plane_cluster;

void ssb_cluster_default_opts(ssb_cluster_opts* o);
/* plane_segmentation::clusterAndSegmentAllPlanes (:261-294).  cloud4 / normals4: host float [n][4] (x y z rgb as produced by
 * ssb_crop_bbox; nx ny nz curvature as produced by ssb_organized_planes' normals_out), transformation_mat: row-major 4x4.
 * rows8: the reference's final_pose_vec, one row per hull vertex (x, y, z, nx, ny, nz, d, 0).  Optional outputs (NULL to skip):
 * labels_out [n] = first k-means label per point (-1 where the normal is NaN), centers_out [num_centroids_normals][3].
 * Returns 1, 0 when there were not enough normals (<= 10, :316-320), or a negative error. */
int ssb_cluster_planes(ssb_ransac* r, const float* cloud4, const float* normals4, int n, const float transformation_mat[16],
                       const ssb_cluster_opts* opts, unsigned long long* rng_state, float* rows8, int max_rows, int* n_rows,
                       ssb_plane_cluster* clusters, int max_clusters, int* n_clusters, int* labels_out, float* centers_out);

void ssb_assoc_default_opts(ssb_assoc_opts* o);
ssb_assoc* ssb_assoc_create(const ssb_assoc_opts* opts);   /* data_association::data_association + init */
void ssb_assoc_destroy(ssb_assoc* a);
/* data_association::find_matches: one ssb_landmark_obs per detection (out[n]); robot_pose = x y z roll pitch yaw
 * (ros_utils.hpp:90-106 matrix2vector).  Returns n or an error. */
int ssb_assoc_find_matches(ssb_assoc* a, const ssb_detection* dets, int n, const float robot_pose[6], float cam_angle,
                           ssb_landmark_obs* out);
/* landmarks_[id].node->estimate() as used by landmarkMeasurementModel (:375-389): the caller refreshes it
 * from the graph after every optimize (assignLandmarkNode :391-393 binds the node in the reference) */
int ssb_assoc_set_landmark_estimate(ssb_assoc* a, int id, const double xyz[3]);
/* data_association::setLandmarkCovs  :395-397 */
int ssb_assoc_set_landmark_cov(ssb_assoc* a, int id, const float cov[9]);
/* data_association::getMappedLandmarks  :399 */
int ssb_assoc_num_landmarks(const ssb_assoc* a);
int ssb_assoc_get_landmark(const ssb_assoc* a, int id, ssb_landmark_obs* out);
/* Test hook: the two float 3 x 3 inverses of the association step (row-major in and out).  kind 0 = the fixed-size
 * Eigen::Matrix3f::inverse() of `information = covariance.inverse()` (semantic_graph_slam.cpp:170, cofactors), kind 1 = the
 * dynamic-size Eigen::MatrixXf::inverse() of the Mahalanobis gate's Q (data_association.h:175-184, PartialPivLU + solve). */
int ssb_assoc_inverse3(int kind, const float a[9], float r[9]);

/* ------------------------------------------------------------------------------------------- */
const char* ssb_last_error(void);
/* "sm_100a" build tag, CUDA runtime version */
const char* ssb_build_info(void);
/* CUDA stream of a graph handle (opaque cudaStream_t) for caller-side event timing */
void* ssb_graph_stream(ssb_graph* g);
/* device-resident benchmark helpers: snapshot the current estimates on the device / restore them,
 * so repeated optimize() calls start from the same state without host traffic */
int ssb_graph_snapshot(ssb_graph* g);
int ssb_graph_restore(ssb_graph* g);

#ifdef __cplusplus
}
#endif
#endif /* SSB_H */
