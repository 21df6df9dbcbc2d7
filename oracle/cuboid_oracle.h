/*
 * cuboid_oracle.h — CPU restatement of the reference's per-frame point-cloud hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under perception_b200/ may include, link or call this.
 * Allowed callers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in PCL (+Eigen, FLANN, Boost.Random), which
 * is neither in /root/reference nor installable offline. The reference holds
 * no golden vectors for the path. What IS pinned: the make_cuboid.py templates
 * (byte-identical regeneration), the image_geometry unprojection KAT, and mt19937's standard
 * known answer. Everything else restates PCL 1.7.2/1.8.1 + Eigen 3.2/3.3 from their published
 * algorithms; each function cites the reference call site it serves.
 *
 * mode: ORC_CANONICAL (0) = the order the CUDA path must reproduce bit-for-bit
 *       ORC_LITERAL   (1) = closest to what a PCL binary does where that is knowable
 *                           (std::sort inside voxels, sequential float/double sums, libm float trig)
 */
#ifndef CUBOID_ORACLE_H
#define CUBOID_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_CANONICAL 0
#define ORC_LITERAL 1
#define ORC_MAX_CLUSTERS 16

/* mirrors include/cuboid_cuda.h cuboid_params field for field (kept separate on purpose) */
typedef struct {
    float fx, fy, cx, cy, depth_scale;
    int32_t _pad0;
    double pass_z_min, pass_z_max, pass_x_min, pass_x_max;
    double pass_z2_min, pass_z2_max;
    int32_t use_pass_z2;
    float leaf;
    double sac_threshold;
    int32_t sac_max_iter;
    uint32_t sac_seed;
    double sac_prob;
    int32_t sac_refine;
    int32_t extract_negative;
    double cluster_tol;
    int32_t cluster_min, cluster_max;
    int32_t use_cluster;
    int32_t icp_max_iter;
    double icp_tf_eps, icp_rel_mse, icp_max_corr_dist;
    double icp_fitness_gate;
    int32_t n_guess;
    int32_t guess_mode; /* 0: guesses are absolute 4x4; 1: rotations applied about the cluster centroid */
} orc_params;

typedef struct {
    int32_t size;
    int32_t converged;
    int32_t iterations;
    int32_t best_guess;
    int32_t state;
    int32_t accepted; /* converged && fitness < gate (icp.cpp:182) */
    double fitness;
    float T[16]; /* row-major final transformation (source -> template) */
    uint64_t corr_hash;
} orc_cluster_result;

typedef struct {
    int32_t status;
    int32_t n_points;
    int32_t n_voxels;
    int32_t min_b[3];
    int32_t div_b[3];
    int32_t plane_found;
    float plane_coeff[4];
    int32_t n_inliers_pre;
    int32_t n_inliers;
    int32_t sac_iterations;
    int32_t sac_draws;
    int32_t n_remain;
    int32_t n_clusters;
    uint64_t points_hash, voxel_key_hash, voxel_hash, inlier_hash, remain_hash, cluster_hash;
    orc_cluster_result cluster[ORC_MAX_CLUSTERS];
} orc_frame_result;

/* ICP convergence states (PCL DefaultConvergenceCriteria) */
enum { ORC_ICP_NOT_CONVERGED = 0, ORC_ICP_ITERATIONS = 1, ORC_ICP_TRANSFORM = 2, ORC_ICP_ABS_MSE = 3,
       ORC_ICP_REL_MSE = 4, ORC_ICP_NO_CORRESPONDENCES = 5 };

/* a0: stage 1a (intrinsics README.md:78) — all w*h points, row-major, xyzw stride 4 floats */
int orc_unproject(const uint16_t* depth, int w, int h, float fx, float fy, float cx, float cy,
                  float depth_scale, float* xyzw_out);

/* a1: pcl::PassThrough<PCLPointCloud2> (gps.cpp:53-57,61-65; opd.cpp:332-336). field 0/1/2 = x/y/z.
 * src_index_out may be NULL. Returns kept count. */
int orc_passthrough(const float* xyzw, int n, int field, double lo, double hi, float* xyzw_out,
                    int32_t* src_index_out);

/* a2: pcl::VoxelGrid<PCLPointCloud2> (gps.cpp:69-73). key_per_point (n) and voxel_key/voxel_count (V)
 * may be NULL. Returns V; *overflow set when dx*dy*dz > INT32_MAX (PCL warns and carries on). */
int orc_voxel_grid(const float* xyzw, int n, float leaf, int mode, float* vox_xyzw_out,
                   int32_t* key_per_point, int32_t* voxel_key, int32_t* voxel_count, int32_t min_b[3],
                   int32_t div_b[3], int* overflow);

/* a3: pcl::SACSegmentation PLANE/RANSAC/optimize (gps.cpp:76-93). triplets_in NULL -> seeded sampler.
 * Returns 1 when a plane was found, 0 when not. Outputs may be NULL where noted in the .cpp. */
int orc_sac_plane(const float* xyzw, int n, double thr, int max_iter, double prob, uint32_t seed,
                  int refine, int mode, const int32_t* triplets_in, int n_triplets_in, float coeff_out[4],
                  int32_t* inliers_out, int* n_inliers, float coeff_pre[4], int32_t* inliers_pre,
                  int* n_inliers_pre, int* iters_run, int32_t* triplets_used, int cap_triplets,
                  int* n_draws, double* k_margin);

/* a4: pcl::ExtractIndices (gps.cpp:96-101; opd.cpp:378-383). Returns count. */
int orc_extract(const float* xyzw, int n, const int32_t* idx, int n_idx, int negative, float* out,
                int32_t* src_index_out);

/* f1: surface_normal_estimation (cuboid_detection/src/surface_normal_estimation.cpp:105-165 getNormal, :182-233 callback).
 * SACSegmentation with SACMODEL_PERPENDICULAR_PLANE (model_type 1) / SACMODEL_PARALLEL_PLANE (model_type 2), setAxis,
 * setEpsAngle(0.1), optimize = true; semantics [PCL-recall], see cuboid_oracle.cpp: model_valid. */
int orc_sac_plane_model(const float* xyzw, int n, int model_type, const float axis[3], double eps_angle, double thr, int max_iter,
                        double prob, uint32_t seed, int refine, int mode, float coeff_out[4], int32_t* inliers_out, int* n_inliers,
                        int32_t* inliers_pre, int* n_inliers_pre, int* iters_run);
typedef struct {
    float coeff[3][4];      /* plane i of the loop at :196-205 (0: perpendicular model, 1 and 2: parallel model), segmentation order */
    float midpoint[3][3];   /* pcl::compute3DCentroid of its inliers */
    int32_t n_plane[3], found[3], n_in[3];
    int32_t n_left;
    int32_t order[3];       /* segmentation index of the plane that ends up as normals[0], [1], [2] after the sort at :207-221 */
    float Rt[16];           /* :228-237, row-major: columns (n2, n1, n0, centroid) */
} orc_surface_result;
int orc_surface_normals(const float* xyzw, int n, const float axis[3], double eps_angle, double thr, int max_iter, double prob,
                        uint32_t seed, int mode, orc_surface_result* out);
void orc_surface_pose(const float coeff[12], const float midpoint[9], const int32_t n_plane[3], float Rt[16], int32_t order[3]);

/* f4: object_pose_detection service bookkeeping, simulated literally with the node's own containers (opd.cpp:212-247 retry
 * loop pushing one transform per attempt, :365-441 diff scores, argmin, success). sizes/converged/fitness per cluster in
 * cluster order. Returns success; *argmin, *reference_cluster (owner of icp_transforms[argmin], -1 if out of range), attempts[]. */
int orc_select_object(const int32_t* sizes, const int32_t* converged, const double* fitness, int n_clusters, int template_points,
                      double icp_fitness_score, int32_t* argmin, int32_t* reference_cluster, int32_t* attempts);

/* f3: bbox_filter's within_bbox + ExtractIndices (cuboid_detection/src/bbox_filter.cpp:30-51, 89-103): keep point i iff its
 * projection through the 3x4 CameraInfo P lies strictly inside the rectangle (x1,y1,x2,y2). Pinned by the reference
 * source itself (no third-party arithmetic). idx_out may be NULL. Returns count. */
int orc_bbox_filter(const float* xyzw, int n, const double P[12], const int32_t bbox[4], float* out_xyzw, int32_t* idx_out);

/* a5: KdTree + EuclideanClusterExtraction (opd.cpp:346-362). idx_sorted_out has room for n, offsets_out
 * for n_clusters+1 (cap n/min+2). Returns number of kept clusters. */
int orc_cluster(const float* xyzw, int n, double tol, int min_size, int max_size, int32_t* idx_sorted_out,
                int32_t* offsets_out);

/* a7+a8: IterativeClosestPoint::align + getFitnessScore (icp.cpp:170-182). guess NULL = identity.
 * corr_trace (cap_iters*n_src int32) and T_trace (cap_iters*16 float) may be NULL. */
int orc_icp(const float* src_xyzw, int n_src, const float* tgt_xyzw, int n_tgt, const float* guess16,
            int max_iter, double tf_eps, double rel_mse, double max_corr_dist, int mode, float T_out[16],
            double* fitness, int* converged, int* iters, int* state, float* aligned_xyzw,
            int32_t* corr_trace, float* T_trace, int cap_iters, uint64_t* corr_hash);

/* a9: final.cast<double>().inverse() + pose + bbox corners (icp.cpp:179, 55-128) */
void orc_pose_from_transform(const float T[16], double H_out[16], double pose7_out[7]);
void orc_bbox_corners(const double H[16], double l, double w, double h, float corners_out[8 * 4]);

/* guess builder for guess_mode 1: G = T(c) R T(-c), c = canonical centroid of the cluster */
void orc_guess_about_centroid(const float* src_xyzw, int n, const float R9[9], float G16[16]);

/* whole frame, as cuboid_process_batch does it. guesses: n_guess*16 (mode 0) or n_guess*9 (mode 1), NULL = identity. */
int orc_process_frame(const orc_params* p, const uint16_t* depth, int w, int h, const float* tmpl_xyzw,
                      int n_tmpl, const float* guesses, int mode, orc_frame_result* out);
/* same, from an already-unprojected PointCloud2-style blob (the node's real input, gps.cpp:43-49) */
int orc_process_cloud(const orc_params* p, const void* blob, int point_step, int xoff, int yoff, int zoff,
                      int n, const float* tmpl_xyzw, int n_tmpl, const float* guesses, int mode,
                      orc_frame_result* out);

/* A.8: ASCII PCD reader (strtof). Returns point count or <0. xyzw_out NULL -> count only. */
int orc_load_pcd(const char* path, float* xyzw_out, int cap);

/* known answers */
uint32_t orc_mt19937_nth(uint32_t seed, int n); /* n-th raw output (1-based) */
uint64_t orc_hash_i32(const int32_t* v, int n);
uint64_t orc_hash_f32x3(const float* xyzw, int n);
int orc_params_size(void);
int orc_frame_result_size(void);

#ifdef __cplusplus
}
#endif
#endif
