/*
 * cuboid_cuda.h — C ABI of libcuboid_cuda.so, the B200 (sm_100a) drop-in for the per-frame
 * point-cloud hot path of dash-robotics/perception (cuboid_detection + object_detection).
 *
 * The reference has no function boundary on this path: PCL objects are built on the stack inside
 * three ROS callbacks. Each entry point below names the PCL call sequence it
 * replaces. Abbreviations (paths relative to the reference root):
 *   gps.cpp = cuboid_detection/src/ground_plane_segmentation.cpp
 *   icp.cpp = cuboid_detection/src/iterative_closest_point.cpp
 *   opd.cpp = object_detection/src/object_pose_detection.cpp
 *
 * Conventions: plain pointers and sizes, host buffers in / caller-allocated host buffers out, the
 * library copies in and out and keeps no caller pointer after return. Every call returns an int
 * status (0 = CUBOID_OK, negative = error); "no plane" / "not converged" are data, not errors.
 * A handle is thread-compatible (one call in flight), matching the single ros::spin() thread of
 * each node (gps.cpp:153, icp.cpp:236, opd.cpp:488). Nothing throws across this boundary.
 * There is no CPU fallback: without a CUDA device cuboid_create fails with CUBOID_E_NO_DEVICE.
 */
#ifndef CUBOID_CUDA_H
#define CUBOID_CUDA_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUBOID_ABI_VERSION 1
#define CUBOID_MAX_CLUSTERS 16
#define CUBOID_MAX_TEMPLATES 8

enum {
    CUBOID_OK = 0,
    CUBOID_E_INVALID = -1,     /* bad argument */
    CUBOID_E_NO_DEVICE = -2,   /* no usable CUDA device */
    CUBOID_E_CUDA = -3,        /* CUDA runtime error (see cuboid_last_error) */
    CUBOID_E_CAPACITY = -4,    /* caller buffer or internal capacity too small */
    CUBOID_E_NO_TEMPLATE = -5, /* template slot empty (icp.cpp:159-163: "Couldn't read the template") */
    CUBOID_E_UNSUPPORTED = -6
};

/* frame status bits (cuboid_frame_result.status) */
enum {
    CUBOID_W_VOXEL_OVERFLOW = 1, /* dx*dy*dz > INT32_MAX: PCL warns and carries on; so do we */
    CUBOID_W_RNG_EXHAUSTED = 2,  /* RANSAC sampler ran past the precomputed mt19937 table */
    CUBOID_W_CLUSTERS_TRUNCATED = 4,
    CUBOID_W_CLUSTER_RANGE = 8   /* a non-plane point lies more than 5e5 clustering cells (2.6e5 * cluster_tol) from the origin:
                                    k_cluster's cell argument is not guaranteed for this frame */
};

/* ICP convergence states = pcl::registration::DefaultConvergenceCriteria */
enum {
    CUBOID_ICP_NOT_CONVERGED = 0,
    CUBOID_ICP_ITERATIONS = 1,
    CUBOID_ICP_TRANSFORM = 2,
    CUBOID_ICP_ABS_MSE = 3,
    CUBOID_ICP_REL_MSE = 4,
    CUBOID_ICP_NO_CORRESPONDENCES = 5
};

/* Mirrors the launch parameters the nodes read plus the constants they hard-code. */
typedef struct {
    float fx, fy, cx, cy, depth_scale;   /* README.md:78 ; 0.001 m per depth unit */
    int32_t _pad0;
    double pass_z_min, pass_z_max;       /* 0, 0.9        gps.cpp:56 */
    double pass_x_min, pass_x_max;       /* -0.2, 0.2     gps.cpp:64 */
    double pass_z2_min, pass_z2_max;     /* 0, 0.75       opd.cpp:335 */
    int32_t use_pass_z2;                 /* opd.cpp:331-336 only */
    float leaf;                          /* voxel_size    cuboid_detection/launch/ground_plane_segmentation.launch:16 */
    double sac_threshold;                /* distance_threshold  gps.cpp:89 */
    int32_t sac_max_iter;                /* 1000          gps.cpp:88 */
    uint32_t sac_seed;                   /* 12345: SACSegmentation's fixed boost::mt19937 seed */
    double sac_prob;                     /* 0.99 (PCL default) */
    int32_t sac_refine;                  /* setOptimizeCoefficients(true)  gps.cpp:85 */
    int32_t extract_negative;            /* invert        gps.cpp:100 */
    double cluster_tol;                  /* 0.02          opd.cpp:356; accepted range [1e-4, 1e3] m */
    int32_t cluster_min, cluster_max;    /* 200, 25000    opd.cpp:357-358 */
    int32_t use_cluster;                 /* 1: opd.cpp:346-413 per-cluster ICP; 0: icp.cpp:156-178 whole cloud */
    int32_t icp_max_iter;                /* 5000          icp.cpp:173 */
    double icp_tf_eps;                   /* 1e-9          icp.cpp:174 */
    double icp_rel_mse;                  /* icp_fitness_score as setEuclideanFitnessEpsilon  icp.cpp:176 */
    double icp_max_corr_dist;            /* sqrt(DBL_MAX): icp.cpp:175 is commented out; a finite value drops the pairs beyond it */
    double icp_fitness_gate;             /* icp_fitness_score as the acceptance gate         icp.cpp:182 */
    int32_t n_guess;                     /* initial-pose hypotheses per cluster (>=1) in cuboid_process_* */
    int32_t guess_mode;                  /* 0: absolute 4x4 guesses; 1: 3x3 rotations about the cluster centroid */
} cuboid_params;

typedef struct {
    int32_t size;        /* points in the cluster (ICP source) */
    int32_t converged;   /* icp.hasConverged() */
    int32_t iterations;
    int32_t best_guess;  /* lowest fitness, ties -> lowest guess id */
    int32_t state;       /* CUBOID_ICP_* */
    int32_t accepted;    /* converged && fitness < icp_fitness_gate   icp.cpp:182 / opd.cpp:235 */
    double fitness;      /* icp.getFitnessScore() */
    float T[16];         /* icp.getFinalTransformation(), row-major, source(camera) -> template */
    uint64_t corr_hash;  /* position-keyed hash of every iteration's correspondences (parity tap) */
} cuboid_cluster_result;

typedef struct {
    int32_t status;            /* CUBOID_W_* bits */
    int32_t n_points;          /* after PassThrough z, x          gps.cpp:53-65 */
    int32_t n_voxels;          /* after VoxelGrid                 gps.cpp:69-73 */
    int32_t min_b[3];          /* VoxelGrid min_b_ */
    int32_t div_b[3];          /* VoxelGrid div_b_ */
    int32_t plane_found;
    float plane_coeff[4];      /* pcl::ModelCoefficients published at gps.cpp:105-107 */
    int32_t n_inliers_pre;     /* best-model inliers before optimizeModelCoefficients */
    int32_t n_inliers;         /* inliers->indices.size() after the refine + reselect */
    int32_t sac_iterations;
    int32_t sac_draws;
    int32_t n_remain;          /* after ExtractIndices (+ PassThrough z2) */
    int32_t n_clusters;
    uint64_t points_hash, voxel_key_hash, voxel_hash, inlier_hash, remain_hash, cluster_hash; /* parity taps */
    cuboid_cluster_result cluster[CUBOID_MAX_CLUSTERS];
} cuboid_frame_result;

typedef struct cuboid_handle cuboid_handle; /* opaque: device buffers, streams, templates */

/* ---- lifetime -------------------------------------------------------------------------- */
void cuboid_default_params(cuboid_params* p);   /* cuboid_detection launch defaults */
/* max_points bounds one frame/cloud (e.g. 640*480); max_batch bounds frames resident per chunk. */
int cuboid_create(cuboid_handle** out, const cuboid_params* p, int device, int max_points, int max_batch);
int cuboid_destroy(cuboid_handle* h);
int cuboid_set_params(cuboid_handle* h, const cuboid_params* p);
/* replaces pcl::io::loadPCDFile per callback (icp.cpp:159, opd.cpp:398): upload once, reuse (quirk Q7) */
int cuboid_set_template(cuboid_handle* h, int slot, const float* xyz, int stride_bytes, int n);
/* n_guess entries of 16 floats (guess_mode 0, row-major 4x4) or 9 floats (guess_mode 1); NULL = identity only */
int cuboid_set_guesses(cuboid_handle* h, const float* guesses, int n_guess, int guess_mode);

/* ---- stage entry points (host in, host out) --------------------------------------------- */
/* stage 1a — realsense2_camera's deprojection (absent from the reference; README.md:78 intrinsics).
 * Writes all w*h points row-major (x,y,z,1); *n_out = w*h. */
int cuboid_unproject(cuboid_handle* h, const uint16_t* depth, int w, int hgt, float* xyzw_out, int cap, int* n_out);
/* PassThrough z + PassThrough x + VoxelGrid on a PointCloud2-style blob (gps.cpp:49-73, opd.cpp:273-298).
 * key_per_point_out (n int32, voxel idx of each SURVIVING point in input order, compacted) may be NULL. */
int cuboid_preprocess(cuboid_handle* h, const void* pts, int point_step, int xoff, int yoff, int zoff, int n,
                      float* vox_xyzw_out, int cap, int* n_vox, int32_t* key_per_point_out, int* n_pass);
/* PointCloud2 inputs that carry a packed "rgb" / "rgba" field (realsense2_camera's cloud: x, y, z, rgb): tell the library where it
 * sits inside a record (sensor_msgs/PointField offset, 4-byte aligned; -1 = none, the default). PassThrough and ExtractIndices
 * copy whole PCLPointCloud2 records and VoxelGrid<PCLPointCloud2> (gps.cpp:69-73, downsample_all_data_ = true) averages r, g, b
 * per voxel, so the cloud the node republishes (gps.cpp:110-112) keeps its colour. With a field declared, the w component of every
 * xyzw point the library returns for a PointCloud2 input (cuboid_preprocess, cuboid_batch_fetch what = 0 / 2 / 4) holds that packed
 * word - per voxel (int(mean r) << 16) | (int(mean g) << 8) | int(mean b) - instead of pcl::PointXYZ's 1.0f padding. Other extra
 * fields are not carried. */
int cuboid_set_cloud_fields(cuboid_handle* h, int rgb_offset);
/* SACSegmentation::segment + ExtractIndices (gps.cpp:76-101). triplets NULL -> internal seeded sampler
 * (boost::mt19937(12345) stream shared with the oracle). Any output pointer may be NULL. */
int cuboid_segment_plane(cuboid_handle* h, const float* xyzw, int n, const int32_t* triplets, int n_triplets,
                         float coeff_out[4], int32_t* inlier_idx_out, int* n_inl, int32_t* inlier_pre_out,
                         int* n_inl_pre, float* remain_xyzw_out, int* n_remain, int* iters_run, int* plane_found);
/* surface_normal_estimation node (cuboid_detection/src/surface_normal_estimation.cpp): the coarse cuboid pose the
 * reference meant as ICP's initial guess (icp.cpp:165-167, 227). Three SACSegmentation runs on the shrinking cloud
 * (:196-205): plane 0 with SACMODEL_PERPENDICULAR_PLANE, planes 1 and 2 with SACMODEL_PARALLEL_PLANE, setAxis(axis) =
 * the table normal of ground_plane_segmentation, setEpsAngle(eps_angle = 0.1), optimize = true, 1000 iterations,
 * distance_threshold from launch/surface_normal_estimation.launch (0.015); pcl::compute3DCentroid of each plane;
 * then the ordering, handedness fix, projection and pose of :207-237 / :64-96. */
typedef struct {
    float coeff[3][4];      /* plane i in segmentation order (published as normal_z/y/x after the ordering) */
    float midpoint[3][3];   /* pcl::compute3DCentroid of plane i's inliers */
    int32_t n_plane[3];     /* inliers of plane i */
    int32_t found[3];
    int32_t n_in[3];        /* points the i-th segmentation ran on */
    int32_t n_left;         /* points left after the third plane */
    int32_t order[3];       /* segmentation index of normals[0], [1], [2] after the point-count ordering (:207-221) */
    float Rt[16];           /* :228-237 row-major, columns (normals[2], normals[1], normals[0], centroid) */
    double pose7[7];        /* geometry_msgs/Pose published on /surface_segmentation/pose: x y z qx qy qz qw */
} cuboid_surface_result;
int cuboid_surface_normals(cuboid_handle* h, const float* xyzw, int n, const float axis[3], double eps_angle,
                           double distance_threshold, cuboid_surface_result* out);
/* the host arithmetic of the same callback on already-segmented planes (coeff 3x4, midpoint 3x3) */
void cuboid_surface_pose(const float coeff[12], const float midpoint[9], const int32_t n_plane[3], float Rt[16],
                         int32_t order[3], double pose7[7]);
/* bbox_filter node (cuboid_detection/src/bbox_filter.cpp:30-51 within_bbox, :89-103 ExtractIndices): keeps, in order, the
 * points whose projection through the 3x4 CameraInfo matrix P (row-major doubles, bbox_filter.cpp:54-66) lies strictly
 * inside the rectangle bbox = (x1, y1, x2, y2) of color_object_detection/Rectangle (bbox_filter.cpp:68-74).
 * idx_out / xyzw_out (cap entries) may be NULL. */
int cuboid_bbox_filter(cuboid_handle* h, const void* pts, int point_step, int xoff, int yoff, int zoff, int n,
                       const double P[12], const int32_t bbox[4], int32_t* idx_out, float* xyzw_out, int cap, int* n_out);
/* The same predicate fused into the extraction of cuboid_process_* / cuboid_segment_plane, i.e. the node chain
 * ground_plane_segmentation -> bbox_filter -> iterative_closest_point (launch/iterative_closest_point.launch:15).
 * enable = 0 (default) restores the reference's default wiring (ICP on /ground_plane_segmentation/points). */
int cuboid_set_bbox_filter(cuboid_handle* h, const double P[12], const int32_t bbox[4], int enable);
/* search::KdTree + EuclideanClusterExtraction (opd.cpp:346-362). idx_sorted_out: n entries;
 * offsets_out: cap_clusters+1 entries. Clusters ordered size-descending, ties by smallest member. */
int cuboid_cluster(cuboid_handle* h, const float* xyzw, int n, int32_t* idx_sorted_out, int32_t* offsets_out,
                   int cap_clusters, int* n_clusters);
/* IterativeClosestPoint::align + getFitnessScore (icp.cpp:170-182, opd.cpp:220-235).
 * guesses_4x4 NULL -> identity. corr_trace (cap_trace_iters*n_src int32) and T_trace (cap_trace_iters*16)
 * are the per-iteration debug taps of the FIRST guess; NULL to skip. */
int cuboid_icp(cuboid_handle* h, const float* src_xyzw, int n_src, int tmpl_slot, const float* guesses_4x4,
               int n_guess, float best_T_out[16], double* fitness_out, int* converged, int* iters, int* state,
               int* best_guess, float* aligned_xyzw_out, int32_t* corr_trace, float* T_trace,
               int cap_trace_iters, uint64_t* corr_hash);

/* ---- whole-callback entry points --------------------------------------------------------- */
/* One PointCloud2 message through the whole chain (gps.cpp:43-112 + icp.cpp:136-203, or opd.cpp:270-442). */
int cuboid_process_cloud(cuboid_handle* h, const void* pts, int point_step, int xoff, int yoff, int zoff, int n,
                         int tmpl_slot, cuboid_frame_result* result);
/* Throughput entry: n_frames independent depth frames (host memory), stage 1a included. */
int cuboid_process_batch(cuboid_handle* h, const uint16_t* depth, int w, int hgt, int n_frames, int tmpl_slot,
                         cuboid_frame_result* results);
/* Same work on frames already resident in device memory (depth_dev = device pointer); results stay on the
 * device until cuboid_batch_results. Used for kernel-only timing. stages: bit0 preprocess+voxel,
 * bit1 plane segmentation, bit2 clustering, bit3 ICP (lower bits required by higher ones). */
int cuboid_process_batch_device(cuboid_handle* h, const void* depth_dev, int w, int hgt, int n_frames,
                                int tmpl_slot, int stages);
int cuboid_batch_results(cuboid_handle* h, cuboid_frame_result* results, int n_frames);
/* Parity taps on the most recent batch: copy one frame's intermediate array to the host.
 * what: 0 points(xyzw) 1 voxel key per point(int32) 2 voxels(xyzw) 3 inlier indices(int32)
 *       4 remaining points(xyzw) 5 cluster index list(int32) 6 cluster offsets(int32) 7 points per voxel(int32)
 * Only frames of the last resident chunk (the last max_batch frames of the batch) can be fetched. */
int cuboid_batch_fetch(cuboid_handle* h, int frame, int what, void* out, int cap_bytes, int* n_items);

/* ---- publish side (icp.cpp:179, 55-128) — host arithmetic, kept here so a node does one call -- */
void cuboid_pose_from_transform(const float T[16], double H_out[16], double pose7_out[7]);
void cuboid_bbox_corners(const double H[16], double l, double w, double hgt, float corners_xyzw_out[32]);

/* ---- object_pose_detection service bookkeeping (opd.cpp:212-247, 365-441) — host logic over one frame result ------------
 * The service runs ICP per cluster against the requested template, retries the (deterministic) ICP until it converged
 * with fitness < icp_fitness_score or 11 attempts were made (:215-246, quirk Q4), scores every cluster by
 * |cluster points - template points| (:411), takes the first minimum below 1000 (:415-422) and succeeds iff that
 * difference is < 250 (:429, quirk Q5). It then reads icp_transforms[argmin], a vector with one entry per ATTEMPT
 * (:230 vs :426, quirk Q3): `reference_cluster` is the cluster whose transform that actually is; `argmin` is the
 * cluster the node meant. H_* = getFinalTransformation().cast<double>().inverse() of those clusters (:229). */
typedef struct {
    int32_t n_clusters;
    int32_t argmin;                          /* -1: no cluster within 1000 points of the template (the node then reads out of bounds) */
    int32_t success;                         /* res.success */
    int32_t reference_cluster;               /* cluster owning icp_transforms[argmin] in the reference (== argmin unless an earlier cluster retried) */
    int32_t attempts[CUBOID_MAX_CLUSTERS];   /* ICP runs the reference would have made for cluster i: 1 or 11 */
    double diff_score[CUBOID_MAX_CLUSTERS];
    double icp_score[CUBOID_MAX_CLUSTERS];
    double H_argmin[16], H_reference[16];    /* row-major 4x4, identity when undefined */
} cuboid_object_selection;
int cuboid_select_object(const cuboid_frame_result* frame, int template_points, double icp_fitness_score,
                         cuboid_object_selection* out);

/* ---- multi-GPU -------------------------------------------------------------------
 * Frames are independent: the caller shards them over GPUs (one handle per GPU), no exchange step exists.
 * Second axis, for single-frame latency: the n_guess initial-pose hypotheses of every cluster (icp.cpp:165-178 runs ONE; the
 * hypotheses are north_star's extension) are split over the GPUs. Rank r uploads its slice [g0, g1) with cuboid_set_guesses,
 * declares g0 with cuboid_set_guess_offset (best_guess in every result then carries GLOBAL ids), runs the same frame, turns
 * every (frame, cluster) result into an 80-byte record, all-gathers the records (ncclAllGather / MPI / torch.distributed —
 * the library does not link a communication library) and calls cuboid_reduce_guess_records on the gathered array.
 * The winner is the lexicographic minimum of (fitness, guess id) on the exact double: identical for 1, 2, 4, 8 GPUs. */
typedef struct {
    double fitness;          /* icp.getFitnessScore() of this rank's best hypothesis for the cluster */
    int32_t guess_id;        /* global hypothesis id */
    int32_t iter_state;      /* iterations | state << 24 | converged << 28 */
    float T[16];
} cuboid_guess_record;       /* 80 bytes */
int cuboid_set_guess_offset(cuboid_handle* h, int id_offset);
void cuboid_guess_record_from_result(const cuboid_cluster_result* c, cuboid_guess_record* out);
/* recs: n_ranks records of ONE (frame, cluster); out keeps its size field, gets the winner's pose, fitness, iterations, state,
 * converged, best_guess and accepted = converged && fitness < gate; corr_hash (a per-rank parity tap) is cleared. */
int cuboid_reduce_guess_records(const cuboid_guess_record* recs, int n_ranks, double gate, cuboid_cluster_result* out);
/* Legacy 64-bit key (fitness bits with the low 16 mantissa bits replaced by the guess id) for a single MIN all-reduce. It
 * orders fitness values that agree to 2^-36 relative by guess id, i.e. it is NOT the exact rule above: use the records. */
uint64_t cuboid_pack_fitness_key(double fitness, int32_t guess_id);
void cuboid_unpack_fitness_key(uint64_t key, double* fitness, int32_t* guess_id);

/* ---- introspection ------------------------------------------------------------------------- */
const char* cuboid_strerror(int status);
const char* cuboid_last_error(cuboid_handle* h);
int cuboid_abi_version(void);
int cuboid_params_size(void);
int cuboid_frame_result_size(void);
/* kernel launches issued by this handle since creation (bench.py's gpu_launches) */
int64_t cuboid_launch_count(cuboid_handle* h);
/* last batch: device time in ms of stage s (0 preprocess,1 voxel,2 plane,3 cluster,4 icp), CUDA events */
int cuboid_stage_ms(cuboid_handle* h, float ms_out[5]);
/* options: CUBOID_OPT_ICP_CULL (default 1; 0 = plain brute force over every template chunk, same results),
 *          CUBOID_OPT_TAPS (default 1; 0 = do not keep the per-point voxel key / per-voxel count arrays and leave the
 *          parity hashes points_hash / voxel_key_hash / voxel_hash / corr_hash at 0; every other output is unchanged),
 *          CUBOID_OPT_STAGES (default 15; stage bits cuboid_process_batch / cuboid_process_cloud run, as in
 *          cuboid_process_batch_device: e.g. 3 = ground-plane segmentation only),
 *          CUBOID_OPT_FRONTEND (default 1: stages 1a+1b run as ONE kernel, one thread-block cluster per frame;
 *          0 = the unfused kernels, kept as the byte-for-byte cross-check of the fused one),
 *          CUBOID_OPT_PIPELINE (default 1: inside one cuboid_process_batch call every sub-chunk of 256 frames runs all
 *          its stages on its own stream as soon as its depth copy lands -- best for ONE handle called in a loop;
 *          0 = only the front end follows the copies sub-chunk by sub-chunk and plane / clusters / ICP are launched
 *          once over the whole chunk -- best when several handles, one host thread each, share the GPU and overlap
 *          each other's copies; same results either way) */
enum { CUBOID_OPT_ICP_CULL = 1, CUBOID_OPT_TAPS = 2, CUBOID_OPT_STAGES = 3, CUBOID_OPT_FRONTEND = 4, CUBOID_OPT_PIPELINE = 5 };
int cuboid_set_option(cuboid_handle* h, int option, int value);
/* last batch: out[0] = source-template pairs the ICP kernel actually evaluated, out[1] = pairs of the
 * brute-force equivalent (S*T per nearest-neighbour pass). Roofline accounting for the culled kernel. */
int cuboid_icp_work(cuboid_handle* h, uint64_t out[2]);
/* developer counters of the ICP kernel (all zero unless the library was built with -DCUBOID_ICP_STATS); reset != 0 clears them */
int cuboid_debug_counters(cuboid_handle* h, uint64_t out[32], int reset);
/* un-fused FP32 (FMUL+FADD) and FFMA throughput micro-benchmark, lane-ops/s -> TFLOP/s */
int cuboid_measure_fp32_peak(cuboid_handle* h, double* unfused_tflops, double* ffma_tflops);

#ifdef __cplusplus
}
#endif
#endif
