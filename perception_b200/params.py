"""ctypes mirrors of include/cuboid_cuda.h structs + the reference's launch defaults."""
import ctypes as C
import math

MAX_CLUSTERS = 16


class CuboidParams(C.Structure):
    _fields_ = [
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("depth_scale", C.c_float),
        ("_pad0", C.c_int32),
        ("pass_z_min", C.c_double), ("pass_z_max", C.c_double), ("pass_x_min", C.c_double), ("pass_x_max", C.c_double),
        ("pass_z2_min", C.c_double), ("pass_z2_max", C.c_double),
        ("use_pass_z2", C.c_int32), ("leaf", C.c_float),
        ("sac_threshold", C.c_double), ("sac_max_iter", C.c_int32), ("sac_seed", C.c_uint32), ("sac_prob", C.c_double),
        ("sac_refine", C.c_int32), ("extract_negative", C.c_int32),
        ("cluster_tol", C.c_double), ("cluster_min", C.c_int32), ("cluster_max", C.c_int32),
        ("use_cluster", C.c_int32), ("icp_max_iter", C.c_int32),
        ("icp_tf_eps", C.c_double), ("icp_rel_mse", C.c_double), ("icp_max_corr_dist", C.c_double),
        ("icp_fitness_gate", C.c_double),
        ("n_guess", C.c_int32), ("guess_mode", C.c_int32),
    ]


class ClusterResult(C.Structure):
    _fields_ = [
        ("size", C.c_int32), ("converged", C.c_int32), ("iterations", C.c_int32), ("best_guess", C.c_int32),
        ("state", C.c_int32), ("accepted", C.c_int32), ("fitness", C.c_double), ("T", C.c_float * 16),
        ("corr_hash", C.c_uint64),
    ]


class FrameResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("n_points", C.c_int32), ("n_voxels", C.c_int32),
        ("min_b", C.c_int32 * 3), ("div_b", C.c_int32 * 3), ("plane_found", C.c_int32),
        ("plane_coeff", C.c_float * 4), ("n_inliers_pre", C.c_int32), ("n_inliers", C.c_int32),
        ("sac_iterations", C.c_int32), ("sac_draws", C.c_int32), ("n_remain", C.c_int32), ("n_clusters", C.c_int32),
        ("points_hash", C.c_uint64), ("voxel_key_hash", C.c_uint64), ("voxel_hash", C.c_uint64),
        ("inlier_hash", C.c_uint64), ("remain_hash", C.c_uint64), ("cluster_hash", C.c_uint64),
        ("cluster", ClusterResult * MAX_CLUSTERS),
    ]


# D435 depth intrinsics recorded at /root/reference/README.md:78
D435_FX = 384.0898742675781
D435_FY = 384.0898742675781
D435_CX = 322.4656677246094
D435_CY = 240.64073181152344


def default_params(variant="cuboid"):
    """Launch defaults. 'cuboid' = cuboid_detection launch files (leaf .005, thr .015, whole-cloud ICP after
    clustering per north_star); 'object' = object_detection.launch (leaf .001, thr .01, pass z2, clusters)."""
    p = CuboidParams()
    p.fx, p.fy, p.cx, p.cy, p.depth_scale = D435_FX, D435_FY, D435_CX, D435_CY, 0.001
    p.pass_z_min, p.pass_z_max, p.pass_x_min, p.pass_x_max = 0.0, 0.9, -0.2, 0.2
    p.pass_z2_min, p.pass_z2_max, p.use_pass_z2 = 0.0, 0.75, 0
    p.leaf = 0.005
    p.sac_threshold, p.sac_max_iter, p.sac_seed, p.sac_prob, p.sac_refine = 0.015, 1000, 12345, 0.99, 1
    p.extract_negative = 1
    p.cluster_tol, p.cluster_min, p.cluster_max, p.use_cluster = 0.02, 200, 25000, 1
    p.icp_max_iter, p.icp_tf_eps = 5000, 1e-9
    p.icp_rel_mse = 0.0004
    p.icp_fitness_gate = 0.0004
    p.icp_max_corr_dist = math.sqrt(1.7976931348623157e308)
    p.n_guess, p.guess_mode = 1, 0
    if variant == "object":
        p.leaf, p.sac_threshold, p.use_pass_z2 = 0.001, 0.01, 1
    elif variant == "cuboid_nocluster":
        p.use_cluster = 0
    elif variant == "multi8":
        p.pass_x_min, p.pass_x_max = -0.4, 0.4
    elif variant == "hd720":
        p.fx = p.fy = 640.0
        p.cx, p.cy = 642.4656677, 360.6407318
        p.pass_z_max, p.pass_x_min, p.pass_x_max = 2.0, -1.0, 1.0
        p.leaf = 0.002
    elif variant != "cuboid":
        raise ValueError(variant)
    return p
