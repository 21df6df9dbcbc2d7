"""ctypes wrapper around oracle/libcuboid_oracle.so — TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg, and by
nothing under perception_b200/. PARITY UNPINNED: see oracle/cuboid_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcuboid_oracle.so")
CANONICAL, LITERAL = 0, 1
MAX_CLUSTERS = 16


class Params(C.Structure):
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


def build(force=False):
    src = [os.path.join(HERE, "cuboid_oracle.cpp"), os.path.join(HERE, "cuboid_oracle.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in src):
        return LIB
    r = subprocess.run(["make", "-C", HERE, "-B"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_hash_i32.restype = C.c_uint64
        L.orc_hash_f32x3.restype = C.c_uint64
        L.orc_mt19937_nth.restype = C.c_uint32
        assert L.orc_params_size() == C.sizeof(Params), (L.orc_params_size(), C.sizeof(Params))
        assert L.orc_frame_result_size() == C.sizeof(FrameResult)
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def params_from(src):
    """Copy any struct with the same field names (e.g. perception_b200.params.CuboidParams) into oracle Params."""
    p = Params()
    for name, _ in Params._fields_:
        setattr(p, name, getattr(src, name))
    return p


def unproject(depth, fx, fy, cx, cy, scale=0.001):
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    h, w = d.shape
    out = np.empty((h * w, 4), np.float32)
    lib().orc_unproject(_p(d), w, h, C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_float(scale), _p(out))
    return out


def passthrough(pts, field, lo, hi):
    pts = _f32(pts)
    out = np.empty_like(pts)
    idx = np.empty(len(pts), np.int32)
    k = lib().orc_passthrough(_p(pts), len(pts), field, C.c_double(lo), C.c_double(hi), _p(out), _p(idx))
    return out[:k].copy(), idx[:k].copy()


def voxel_grid(pts, leaf, mode=CANONICAL):
    pts = _f32(pts)
    n = len(pts)
    out = np.empty((max(n, 1), 4), np.float32)
    kpp = np.empty(max(n, 1), np.int32)
    vk = np.empty(max(n, 1), np.int32)
    vc = np.empty(max(n, 1), np.int32)
    mb = (C.c_int32 * 3)()
    db = (C.c_int32 * 3)()
    ovf = C.c_int(0)
    V = lib().orc_voxel_grid(_p(pts), n, C.c_float(leaf), mode, _p(out), _p(kpp), _p(vk), _p(vc), mb, db, C.byref(ovf))
    return dict(vox=out[:V].copy(), key_per_point=kpp[:n].copy(), voxel_key=vk[:V].copy(), voxel_count=vc[:V].copy(),
                min_b=list(mb), div_b=list(db), overflow=ovf.value)


def sac_plane(pts, thr=0.015, max_iter=1000, prob=0.99, seed=12345, refine=1, mode=CANONICAL, triplets=None):
    pts = _f32(pts)
    n = len(pts)
    coeff = (C.c_float * 4)()
    coeff_pre = (C.c_float * 4)()
    inl = np.empty(max(n, 1), np.int32)
    inl_pre = np.empty(max(n, 1), np.int32)
    n_inl, n_pre, iters, draws = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    cap = 12000
    trips = np.empty((cap, 3), np.int32)
    km = C.c_double(0)
    tin = None if triplets is None else np.ascontiguousarray(triplets, dtype=np.int32)
    found = lib().orc_sac_plane(_p(pts), n, C.c_double(thr), max_iter, C.c_double(prob), C.c_uint32(seed), refine, mode,
                                _p(tin), 0 if tin is None else len(tin), coeff, _p(inl), C.byref(n_inl), coeff_pre,
                                _p(inl_pre), C.byref(n_pre), C.byref(iters), _p(trips), cap, C.byref(draws), C.byref(km))
    return dict(found=bool(found), coeff=np.array(list(coeff), np.float32), coeff_pre=np.array(list(coeff_pre), np.float32),
                inliers=inl[:n_inl.value].copy(), inliers_pre=inl_pre[:n_pre.value].copy(), iters=iters.value,
                draws=draws.value, triplets=trips[:min(draws.value, cap)].copy(), k_margin=km.value)


def extract(pts, idx, negative=True):
    pts = _f32(pts)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    out = np.empty((max(len(pts), len(idx), 1), 4), np.float32)
    src = np.empty(max(len(pts), len(idx), 1), np.int32)
    k = lib().orc_extract(_p(pts), len(pts), _p(idx), len(idx), int(negative), _p(out), _p(src))
    return out[:k].copy(), src[:k].copy()


class SurfaceResult(C.Structure):
    _fields_ = [("coeff", (C.c_float * 4) * 3), ("midpoint", (C.c_float * 3) * 3), ("n_plane", C.c_int32 * 3),
                ("found", C.c_int32 * 3), ("n_in", C.c_int32 * 3), ("n_left", C.c_int32), ("order", C.c_int32 * 3),
                ("Rt", C.c_float * 16)]


def sac_plane_model(pts, model_type, axis, eps_angle=0.1, thr=0.015, max_iter=1000, prob=0.99, seed=12345, refine=1, mode=CANONICAL):
    """SACSegmentation with SACMODEL_PERPENDICULAR_PLANE (1) / SACMODEL_PARALLEL_PLANE (2), setAxis + setEpsAngle."""
    pts = _f32(pts)
    n = len(pts)
    ax = np.ascontiguousarray(axis, dtype=np.float32).reshape(3)
    coeff = np.zeros(4, np.float32)
    inl = np.empty(max(n, 1), np.int32)
    pre = np.empty(max(n, 1), np.int32)
    ni, npre, it = C.c_int(0), C.c_int(0), C.c_int(0)
    found = lib().orc_sac_plane_model(_p(pts), n, int(model_type), _p(ax), C.c_double(eps_angle), C.c_double(thr), int(max_iter),
                                      C.c_double(prob), C.c_uint32(seed), int(refine), int(mode), _p(coeff), _p(inl), C.byref(ni),
                                      _p(pre), C.byref(npre), C.byref(it))
    return dict(found=bool(found), coeff=coeff, inliers=inl[:ni.value].copy(), inliers_pre=pre[:npre.value].copy(), iters=it.value)


def surface_normals(pts, axis, eps_angle=0.1, thr=0.015, max_iter=1000, prob=0.99, seed=12345, mode=CANONICAL):
    """surface_normal_estimation callback: three constrained planes + the coarse cuboid pose."""
    pts = _f32(pts)
    ax = np.ascontiguousarray(axis, dtype=np.float32).reshape(3)
    out = SurfaceResult()
    ok = lib().orc_surface_normals(_p(pts), len(pts), _p(ax), C.c_double(eps_angle), C.c_double(thr), int(max_iter),
                                   C.c_double(prob), C.c_uint32(seed), int(mode), C.byref(out))
    return bool(ok), out


def bbox_filter(pts, P, bbox):
    """bbox_filter.cpp: points whose projection through the 3x4 matrix P lies strictly inside (x1, y1, x2, y2)."""
    pts = _f32(pts)
    Pd = np.ascontiguousarray(P, dtype=np.float64).reshape(12)
    bb = np.ascontiguousarray(bbox, dtype=np.int32).reshape(4)
    out = np.empty((max(len(pts), 1), 4), np.float32)
    idx = np.empty(max(len(pts), 1), np.int32)
    k = lib().orc_bbox_filter(_p(pts), len(pts), _p(Pd), _p(bb), _p(out), _p(idx))
    return out[:k].copy(), idx[:k].copy()


def cluster(pts, tol=0.02, min_size=200, max_size=25000):
    pts = _f32(pts)
    n = len(pts)
    idx = np.empty(max(n, 1), np.int32)
    off = np.empty(n // max(min_size, 1) + 3, np.int32)
    k = lib().orc_cluster(_p(pts), n, C.c_double(tol), min_size, max_size, _p(idx), _p(off))
    off = off[:k + 1].copy()
    return idx[:off[-1]].copy(), off


def icp(src, tgt, guess=None, max_iter=5000, tf_eps=1e-9, rel_mse=0.0004, max_corr_dist=None, mode=CANONICAL,
        trace_iters=0):
    src, tgt = _f32(src), _f32(tgt)
    if max_corr_dist is None:
        max_corr_dist = np.sqrt(np.finfo(np.float64).max)
    T = (C.c_float * 16)()
    fit, conv, iters, state, h = C.c_double(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_uint64(0)
    aligned = np.empty((max(len(src), 1), 4), np.float32)
    ct = np.full((trace_iters, len(src)), -9, np.int32) if trace_iters else None
    tt = np.zeros((trace_iters, 16), np.float32) if trace_iters else None
    g = None if guess is None else _f32(guess).reshape(16)
    lib().orc_icp(_p(src), len(src), _p(tgt), len(tgt), _p(g), max_iter, C.c_double(tf_eps), C.c_double(rel_mse),
                  C.c_double(max_corr_dist), mode, T, C.byref(fit), C.byref(conv), C.byref(iters), C.byref(state),
                  _p(aligned), _p(ct), _p(tt), trace_iters, C.byref(h))
    return dict(T=np.array(list(T), np.float32).reshape(4, 4), fitness=fit.value, converged=conv.value,
                iters=iters.value, state=state.value, aligned=aligned[:len(src)].copy(), corr_trace=ct, T_trace=tt,
                corr_hash=h.value)


def pose_from_transform(T):
    T = _f32(T).reshape(16)
    H = (C.c_double * 16)()
    pose = (C.c_double * 7)()
    lib().orc_pose_from_transform(_p(T), H, pose)
    return np.array(list(H)).reshape(4, 4), np.array(list(pose))


def bbox_corners(H, l, w, h):
    Hd = np.ascontiguousarray(H, dtype=np.float64).reshape(16)
    out = np.empty((8, 4), np.float32)
    lib().orc_bbox_corners(_p(Hd), C.c_double(l), C.c_double(w), C.c_double(h), _p(out))
    return out


def guess_about_centroid(src, R):
    src = _f32(src)
    R = _f32(R).reshape(9)
    G = np.empty(16, np.float32)
    lib().orc_guess_about_centroid(_p(src), len(src), _p(R), _p(G))
    return G.reshape(4, 4)


def process_frame(params, depth, tmpl, guesses=None, mode=CANONICAL):
    p = params if isinstance(params, Params) else params_from(params)
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    h, w = d.shape
    t = _f32(tmpl) if tmpl is not None else None
    g = _f32(guesses) if guesses is not None else None
    out = FrameResult()
    lib().orc_process_frame(C.byref(p), _p(d), w, h, _p(t), 0 if t is None else len(t), _p(g), mode, C.byref(out))
    return out


def process_cloud(params, blob, point_step, xoff, yoff, zoff, n, tmpl, guesses=None, mode=CANONICAL):
    p = params if isinstance(params, Params) else params_from(params)
    b = np.ascontiguousarray(blob)
    t = _f32(tmpl) if tmpl is not None else None
    g = _f32(guesses) if guesses is not None else None
    out = FrameResult()
    lib().orc_process_cloud(C.byref(p), _p(b), point_step, xoff, yoff, zoff, n, _p(t), 0 if t is None else len(t), _p(g),
                            mode, C.byref(out))
    return out


def load_pcd(path):
    n = lib().orc_load_pcd(path.encode(), None, 0)
    if n < 0:
        raise IOError("orc_load_pcd(%s) = %d" % (path, n))
    out = np.empty((n, 4), np.float32)
    lib().orc_load_pcd(path.encode(), _p(out), n)
    return out


def hash_i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return int(lib().orc_hash_i32(_p(a), len(a)))


def hash_pts(a):
    a = _f32(a)
    return int(lib().orc_hash_f32x3(_p(a), len(a)))


def mt19937_nth(seed, n):
    return int(lib().orc_mt19937_nth(C.c_uint32(seed), n))
