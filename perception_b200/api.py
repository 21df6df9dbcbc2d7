"""Host-side mirror of the libcuboid_cuda C ABI (include/cuboid_cuda.h) over ctypes.

This is the Python face of the same boundary the patched ROS nodes bind in C++: every
method is one C-ABI call with host buffers in and host buffers out. Method names follow the PCL call
sequence of the reference callbacks they replace:

    preprocess      PassThrough z, PassThrough x, VoxelGrid           gps.cpp:53-73
    segment_plane   SACSegmentation::segment + ExtractIndices         gps.cpp:76-101
    cluster         KdTree + EuclideanClusterExtraction               opd.cpp:346-362
    icp             IterativeClosestPoint::align + getFitnessScore    icp.cpp:170-182
    process_cloud   one PointCloud2 message through the whole chain   gps.cpp:43-112 + icp.cpp:136-203
    process_batch   n depth frames, stage 1a included (throughput entry)

There is no CPU fallback: load() raises when the CUDA library is missing and CuboidCuda() raises when
cuboid_create fails (e.g. no CUDA device).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build
from .params import MAX_CLUSTERS, ClusterResult, CuboidParams, FrameResult, default_params  # noqa: F401

OK = 0
E_INVALID, E_NO_DEVICE, E_CUDA, E_CAPACITY, E_NO_TEMPLATE, E_UNSUPPORTED = -1, -2, -3, -4, -5, -6
STAGE_PREPROCESS, STAGE_PLANE, STAGE_CLUSTER, STAGE_ICP = 1, 2, 4, 8

# every symbol include/cuboid_cuda.h declares (tests check the built library exports all of them)
ABI_SYMBOLS = [
    "cuboid_default_params", "cuboid_create", "cuboid_destroy", "cuboid_set_params", "cuboid_set_template",
    "cuboid_set_guesses", "cuboid_unproject", "cuboid_preprocess", "cuboid_segment_plane", "cuboid_cluster",
    "cuboid_icp", "cuboid_process_cloud", "cuboid_process_batch", "cuboid_process_batch_device",
    "cuboid_batch_results", "cuboid_batch_fetch", "cuboid_pose_from_transform", "cuboid_bbox_corners",
    "cuboid_pack_fitness_key", "cuboid_unpack_fitness_key", "cuboid_strerror", "cuboid_last_error",
    "cuboid_abi_version", "cuboid_params_size", "cuboid_frame_result_size", "cuboid_launch_count",
    "cuboid_stage_ms", "cuboid_measure_fp32_peak", "cuboid_set_option", "cuboid_icp_work",
    "cuboid_debug_counters", "cuboid_set_cloud_fields", "cuboid_set_guess_offset", "cuboid_guess_record_from_result", "cuboid_reduce_guess_records",
    "cuboid_bbox_filter", "cuboid_set_bbox_filter", "cuboid_surface_normals", "cuboid_surface_pose", "cuboid_select_object",
]
OPT_ICP_CULL, OPT_TAPS, OPT_STAGES, OPT_FRONTEND, OPT_PIPELINE = 1, 2, 3, 4, 5


class SurfaceResult(C.Structure):
    """cuboid_surface_result (include/cuboid_cuda.h)."""
    _fields_ = [("coeff", (C.c_float * 4) * 3), ("midpoint", (C.c_float * 3) * 3), ("n_plane", C.c_int32 * 3),
                ("found", C.c_int32 * 3), ("n_in", C.c_int32 * 3), ("n_left", C.c_int32), ("order", C.c_int32 * 3),
                ("Rt", C.c_float * 16), ("pose7", C.c_double * 7)]


class ObjectSelection(C.Structure):
    """cuboid_object_selection (include/cuboid_cuda.h)."""
    _fields_ = [("n_clusters", C.c_int32), ("argmin", C.c_int32), ("success", C.c_int32), ("reference_cluster", C.c_int32),
                ("attempts", C.c_int32 * MAX_CLUSTERS), ("diff_score", C.c_double * MAX_CLUSTERS),
                ("icp_score", C.c_double * MAX_CLUSTERS), ("H_argmin", C.c_double * 16), ("H_reference", C.c_double * 16)]


def select_object(frame_result, template_points, icp_fitness_score):
    """object_pose_detection service bookkeeping over one FrameResult (host logic, no GPU)."""
    out = ObjectSelection()
    st = load().cuboid_select_object(C.byref(frame_result), int(template_points), float(icp_fitness_score), C.byref(out))
    if st != OK:
        raise CuboidError(st, "cuboid_select_object")
    return out


class GuessRecord(C.Structure):
    """cuboid_guess_record (include/cuboid_cuda.h): what one rank contributes per (frame, cluster) to the hypothesis-sharded mode."""
    _fields_ = [("fitness", C.c_double), ("guess_id", C.c_int32), ("iter_state", C.c_int32), ("T", C.c_float * 16)]


def guess_record(cluster_result):
    out = GuessRecord()
    load().cuboid_guess_record_from_result(C.byref(cluster_result), C.byref(out))
    return out


def reduce_guess_records(records, gate, into):
    """records: sequence / ctypes array of GuessRecord of ONE (frame, cluster), one per rank; `into` (a ClusterResult) receives the winner."""
    arr = (GuessRecord * len(records))(*records)
    st = load().cuboid_reduce_guess_records(arr, len(records), float(gate), C.byref(into))
    if st != OK:
        raise CuboidError(st, "cuboid_reduce_guess_records")
    return into


class CuboidError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("libcuboid_cuda: %s (status %d)" % (msg, status))
        self.status = status


_lib = None


def load():
    """dlopen perception_b200/libcuboid_cuda.so. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.CUDA_LIB
    if not os.path.exists(path):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). libcuboid_cuda has no CPU fallback." % path)
    L = C.CDLL(path)
    vp, ip, i32 = C.c_void_p, C.POINTER(C.c_int), C.c_int
    L.cuboid_create.argtypes = [C.POINTER(vp), C.POINTER(CuboidParams), i32, i32, i32]
    L.cuboid_destroy.argtypes = [vp]
    L.cuboid_set_params.argtypes = [vp, C.POINTER(CuboidParams)]
    L.cuboid_set_template.argtypes = [vp, i32, vp, i32, i32]
    L.cuboid_set_guesses.argtypes = [vp, vp, i32, i32]
    L.cuboid_unproject.argtypes = [vp, vp, i32, i32, vp, i32, ip]
    L.cuboid_preprocess.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, i32, ip, vp, ip]
    L.cuboid_segment_plane.argtypes = [vp, vp, i32, vp, i32, vp, vp, ip, vp, ip, vp, ip, ip, ip]
    L.cuboid_cluster.argtypes = [vp, vp, i32, vp, vp, i32, ip]
    L.cuboid_icp.argtypes = [vp, vp, i32, i32, vp, i32, vp, C.POINTER(C.c_double), ip, ip, ip, ip, vp, vp, vp, i32,
                             C.POINTER(C.c_uint64)]
    L.cuboid_process_cloud.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, C.POINTER(FrameResult)]
    L.cuboid_process_batch.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    L.cuboid_process_batch_device.argtypes = [vp, vp, i32, i32, i32, i32, i32]
    L.cuboid_batch_results.argtypes = [vp, vp, i32]
    L.cuboid_batch_fetch.argtypes = [vp, i32, i32, vp, i32, ip]
    L.cuboid_pose_from_transform.argtypes = [vp, vp, vp]
    L.cuboid_pose_from_transform.restype = None
    L.cuboid_bbox_corners.argtypes = [vp, C.c_double, C.c_double, C.c_double, vp]
    L.cuboid_bbox_corners.restype = None
    L.cuboid_pack_fitness_key.argtypes = [C.c_double, C.c_int32]
    L.cuboid_pack_fitness_key.restype = C.c_uint64
    L.cuboid_unpack_fitness_key.argtypes = [C.c_uint64, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    L.cuboid_unpack_fitness_key.restype = None
    L.cuboid_set_guess_offset.argtypes = [vp, i32]
    L.cuboid_set_cloud_fields.argtypes = [vp, i32]
    L.cuboid_guess_record_from_result.argtypes = [vp, vp]
    L.cuboid_guess_record_from_result.restype = None
    L.cuboid_reduce_guess_records.argtypes = [vp, i32, C.c_double, vp]
    L.cuboid_strerror.argtypes = [i32]
    L.cuboid_strerror.restype = C.c_char_p
    L.cuboid_last_error.argtypes = [vp]
    L.cuboid_last_error.restype = C.c_char_p
    L.cuboid_launch_count.argtypes = [vp]
    L.cuboid_launch_count.restype = C.c_int64
    L.cuboid_stage_ms.argtypes = [vp, vp]
    L.cuboid_measure_fp32_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.cuboid_set_option.argtypes = [vp, i32, i32]
    L.cuboid_icp_work.argtypes = [vp, vp]
    L.cuboid_debug_counters.argtypes = [vp, vp, i32]
    L.cuboid_bbox_filter.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, ip]
    L.cuboid_set_bbox_filter.argtypes = [vp, vp, vp, i32]
    L.cuboid_surface_normals.argtypes = [vp, vp, i32, vp, C.c_double, C.c_double, C.POINTER(SurfaceResult)]
    L.cuboid_surface_pose.argtypes = [vp, vp, vp, vp, vp, vp]
    L.cuboid_surface_pose.restype = None
    L.cuboid_select_object.argtypes = [C.POINTER(FrameResult), i32, C.c_double, C.POINTER(ObjectSelection)]
    if L.cuboid_params_size() != C.sizeof(CuboidParams) or L.cuboid_frame_result_size() != C.sizeof(FrameResult):
        raise ImportError("struct layout mismatch between params.py and include/cuboid_cuda.h")
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _xyzw(a):
    a = _f32(a)
    if a.ndim != 2 or a.shape[1] not in (3, 4):
        raise ValueError("expected [n,3] or [n,4] points")
    if a.shape[1] == 3:
        a = np.concatenate([a, np.ones((len(a), 1), np.float32)], axis=1)
    return np.ascontiguousarray(a)


def pack_fitness_key(fitness, guess_id):
    return int(load().cuboid_pack_fitness_key(float(fitness), int(guess_id)))


def unpack_fitness_key(key):
    f, g = C.c_double(0), C.c_int32(0)
    load().cuboid_unpack_fitness_key(C.c_uint64(key), C.byref(f), C.byref(g))
    return f.value, g.value


def pose_from_transform(T):
    """icp.cpp:179 + publish_pose: (H = T^-1 in double, [x,y,z,qx,qy,qz,qw])."""
    T = _f32(T).reshape(16)
    H = np.empty(16, np.float64)
    pose = np.empty(7, np.float64)
    load().cuboid_pose_from_transform(_ptr(T), _ptr(H), _ptr(pose))
    return H.reshape(4, 4), pose


def bbox_corners(H, l, w, h):
    Hd = np.ascontiguousarray(H, dtype=np.float64).reshape(16)
    out = np.empty((8, 4), np.float32)
    load().cuboid_bbox_corners(_ptr(Hd), float(l), float(w), float(h), _ptr(out))
    return out


class CuboidCuda:
    """One handle = one node's worth of state (thread-compatible, one call in flight)."""

    def __init__(self, params=None, device=0, max_points=640 * 480, max_batch=1):
        self.lib = load()
        self.params = params if params is not None else default_params()
        self.max_points, self.max_batch = int(max_points), int(max_batch)
        self._h = C.c_void_p()
        st = self.lib.cuboid_create(C.byref(self._h), C.byref(self.params), int(device), self.max_points, self.max_batch)
        if st != OK:
            self._h = C.c_void_p()
            raise CuboidError(st, "cuboid_create: " + self.lib.cuboid_strerror(st).decode())

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.cuboid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, st, what):
        if st != OK:
            detail = self.lib.cuboid_last_error(self._h).decode()
            raise CuboidError(st, "%s: %s %s" % (what, self.lib.cuboid_strerror(st).decode(), detail))

    def set_params(self, params):
        self._ck(self.lib.cuboid_set_params(self._h, C.byref(params)), "cuboid_set_params")
        self.params = params

    def set_template(self, slot, pts):
        pts = _f32(pts)
        self._ck(self.lib.cuboid_set_template(self._h, int(slot), _ptr(pts), pts.strides[0], len(pts)), "cuboid_set_template")

    def set_guesses(self, guesses, mode=0):
        if guesses is None:
            self._ck(self.lib.cuboid_set_guesses(self._h, None, 1, 0), "cuboid_set_guesses")
            return
        g = _f32(guesses).reshape(-1, 16 if mode == 0 else 9)
        self._ck(self.lib.cuboid_set_guesses(self._h, _ptr(g), len(g), int(mode)), "cuboid_set_guesses")

    def set_cloud_fields(self, rgb_offset=-1):
        """Offset of the packed rgb / rgba field of PointCloud2 inputs (-1: none): carried through PassThrough, VoxelGrid, ExtractIndices."""
        self._ck(self.lib.cuboid_set_cloud_fields(self._h, int(rgb_offset)), "cuboid_set_cloud_fields")

    def set_guess_offset(self, id_offset):
        """Hypothesis-sharded mode: this handle holds guesses [id_offset, id_offset + n) of the global list."""
        self._ck(self.lib.cuboid_set_guess_offset(self._h, int(id_offset)), "cuboid_set_guess_offset")

    # -- stages -------------------------------------------------------------------------------
    def unproject(self, depth):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        hgt, w = d.shape
        out = np.empty((hgt * w, 4), np.float32)
        n = C.c_int(0)
        self._ck(self.lib.cuboid_unproject(self._h, _ptr(d), w, hgt, _ptr(out), len(out), C.byref(n)), "cuboid_unproject")
        return out[:n.value]

    def preprocess(self, pts, point_step=None, xoff=0, yoff=4, zoff=8, n=None):
        """pts: float32 [n,>=3] (point_step = row stride) or a raw uint8 blob with explicit point_step / offsets."""
        a = np.ascontiguousarray(pts)
        if point_step is None:
            a = _f32(a)
            point_step, n = a.strides[0] if a.ndim == 2 and len(a) else 16, len(a)
        n = int(n)
        vox = np.empty((max(n, 1), 4), np.float32)
        kpp = np.empty(max(n, 1), np.int32)
        nv, npass = C.c_int(0), C.c_int(0)
        self._ck(self.lib.cuboid_preprocess(self._h, _ptr(a), int(point_step), xoff, yoff, zoff, n, _ptr(vox), len(vox), C.byref(nv),
                                            _ptr(kpp), C.byref(npass)), "cuboid_preprocess")
        return dict(vox=vox[:nv.value].copy(), key_per_point=kpp[:npass.value].copy(), n_pass=npass.value)

    def segment_plane(self, pts, triplets=None):
        p = _xyzw(pts)
        n = len(p)
        coeff = np.zeros(4, np.float32)
        inl = np.empty(max(n, 1), np.int32)
        pre = np.empty(max(n, 1), np.int32)
        rem = np.empty((max(n, 1), 4), np.float32)
        ni, npre, nr, it, found = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        t = None if triplets is None else np.ascontiguousarray(triplets, dtype=np.int32)
        self._ck(self.lib.cuboid_segment_plane(self._h, _ptr(p), n, _ptr(t), 0 if t is None else len(t), _ptr(coeff), _ptr(inl),
                                               C.byref(ni), _ptr(pre), C.byref(npre), _ptr(rem), C.byref(nr), C.byref(it),
                                               C.byref(found)), "cuboid_segment_plane")
        return dict(found=bool(found.value), coeff=coeff, inliers=inl[:ni.value].copy(), inliers_pre=pre[:npre.value].copy(),
                    remain=rem[:nr.value].copy(), iters=it.value)

    def surface_normals(self, pts, axis, eps_angle=0.1, distance_threshold=0.015):
        """surface_normal_estimation callback on the non-plane cloud; axis = the table normal."""
        p = _xyzw(pts)
        ax = np.ascontiguousarray(axis, dtype=np.float32).reshape(3)
        out = SurfaceResult()
        self._ck(self.lib.cuboid_surface_normals(self._h, _ptr(p), len(p), _ptr(ax), float(eps_angle), float(distance_threshold),
                                                 C.byref(out)), "cuboid_surface_normals")
        return out

    def bbox_filter(self, pts, P, bbox):
        """bbox_filter node: (kept points xyzw, their indices), order preserved."""
        p = _xyzw(pts)
        n = len(p)
        Pd = np.ascontiguousarray(P, dtype=np.float64).reshape(12)
        bb = np.ascontiguousarray(bbox, dtype=np.int32).reshape(4)
        idx = np.empty(max(n, 1), np.int32)
        out = np.empty((max(n, 1), 4), np.float32)
        k = C.c_int(0)
        self._ck(self.lib.cuboid_bbox_filter(self._h, _ptr(p), 16, 0, 4, 8, n, _ptr(Pd), _ptr(bb), _ptr(idx), _ptr(out), max(n, 1),
                                             C.byref(k)), "cuboid_bbox_filter")
        return out[:k.value].copy(), idx[:k.value].copy()

    def set_bbox_filter(self, P=None, bbox=None):
        """Fuse the bbox_filter predicate into the extraction of the pipeline entries (None, None switches it off)."""
        if P is None:
            self._ck(self.lib.cuboid_set_bbox_filter(self._h, None, None, 0), "cuboid_set_bbox_filter")
            return
        Pd = np.ascontiguousarray(P, dtype=np.float64).reshape(12)
        bb = np.ascontiguousarray(bbox, dtype=np.int32).reshape(4)
        self._ck(self.lib.cuboid_set_bbox_filter(self._h, _ptr(Pd), _ptr(bb), 1), "cuboid_set_bbox_filter")

    def cluster(self, pts, cap_clusters=1024):
        p = _xyzw(pts)
        n = len(p)
        idx = np.empty(max(n, 1), np.int32)
        off = np.zeros(cap_clusters + 1, np.int32)
        k = C.c_int(0)
        self._ck(self.lib.cuboid_cluster(self._h, _ptr(p), n, _ptr(idx), _ptr(off), cap_clusters, C.byref(k)), "cuboid_cluster")
        off = off[:k.value + 1].copy()
        return idx[:off[-1]].copy() if k.value else idx[:0].copy(), off

    def icp(self, src, slot=0, guesses=None, trace_iters=0):
        s = _xyzw(src)
        n = len(s)
        g = None if guesses is None else _f32(guesses).reshape(-1, 16)
        T = np.zeros(16, np.float32)
        fit = C.c_double(0)
        conv, iters, state, bg = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        aligned = np.empty((max(n, 1), 4), np.float32)
        ct = np.full((trace_iters, n), -9, np.int32) if trace_iters else None
        tt = np.zeros((trace_iters, 16), np.float32) if trace_iters else None
        h = C.c_uint64(0)
        self._ck(self.lib.cuboid_icp(self._h, _ptr(s), n, int(slot), _ptr(g), 1 if g is None else len(g), _ptr(T), C.byref(fit),
                                     C.byref(conv), C.byref(iters), C.byref(state), C.byref(bg), _ptr(aligned), _ptr(ct), _ptr(tt),
                                     trace_iters, C.byref(h)), "cuboid_icp")
        return dict(T=T.reshape(4, 4), fitness=fit.value, converged=conv.value, iters=iters.value, state=state.value,
                    best_guess=bg.value, aligned=aligned[:n].copy(), corr_trace=ct, T_trace=tt, corr_hash=h.value)

    # -- whole callbacks -----------------------------------------------------------------------
    def process_cloud(self, pts, slot=0, point_step=None, xoff=0, yoff=4, zoff=8, n=None):
        a = np.ascontiguousarray(pts)
        if point_step is None:
            a = _f32(a)
            point_step, n = a.strides[0] if len(a) else 16, len(a)
        out = FrameResult()
        self._ck(self.lib.cuboid_process_cloud(self._h, _ptr(a), int(point_step), xoff, yoff, zoff, int(n), int(slot), C.byref(out)),
                 "cuboid_process_cloud")
        return out

    def process_batch(self, depth, slot=0):
        """depth: uint16 [n,h,w] host array (numpy, or anything exposing a CPU pointer via .ctypes/.data_ptr)."""
        n, hgt, w, ptr, _keep = _depth_desc(depth)
        res = (FrameResult * n)()
        self._ck(self.lib.cuboid_process_batch(self._h, ptr, w, hgt, n, int(slot), C.cast(res, C.c_void_p)), "cuboid_process_batch")
        return res

    def process_batch_device(self, depth_dev_ptr, w, hgt, n, slot=0, stages=15):
        self._ck(self.lib.cuboid_process_batch_device(self._h, C.c_void_p(int(depth_dev_ptr)), w, hgt, n, int(slot), int(stages)),
                 "cuboid_process_batch_device")

    def batch_results(self, n):
        res = (FrameResult * n)()
        self._ck(self.lib.cuboid_batch_results(self._h, C.cast(res, C.c_void_p), n), "cuboid_batch_results")
        return res

    _FETCH = {"points": (0, np.float32, 4), "voxel_keys": (1, np.int32, 1), "voxels": (2, np.float32, 4),
              "inliers": (3, np.int32, 1), "remain": (4, np.float32, 4), "cluster_idx": (5, np.int32, 1),
              "cluster_offsets": (6, np.int32, 1), "voxel_counts": (7, np.int32, 1)}

    def fetch(self, frame, what):
        code, dt, width = self._FETCH[what]
        cap = self.max_points + 8
        buf = np.empty((cap, width), dt)
        n = C.c_int(0)
        self._ck(self.lib.cuboid_batch_fetch(self._h, int(frame), code, _ptr(buf), buf.nbytes, C.byref(n)), "cuboid_batch_fetch")
        out = buf[:n.value].copy()
        return out if width > 1 else out.reshape(-1)

    # -- introspection ---------------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.cuboid_launch_count(self._h))

    def stage_ms(self):
        ms = np.zeros(5, np.float32)
        self._ck(self.lib.cuboid_stage_ms(self._h, _ptr(ms)), "cuboid_stage_ms")
        return dict(zip(["preprocess", "voxel", "plane", "cluster", "icp"], [float(x) for x in ms]))

    def set_option(self, option, value):
        self._ck(self.lib.cuboid_set_option(self._h, int(option), int(value)), "cuboid_set_option")

    def icp_work(self):
        """(pairs evaluated, brute-force-equivalent pairs) of the last batch."""
        w = np.zeros(2, np.uint64)
        self._ck(self.lib.cuboid_icp_work(self._h, _ptr(w)), "cuboid_icp_work")
        return int(w[0]), int(w[1])

    def debug_counters(self, reset=True):
        """Developer counters of k_icp (zeros unless the library was built with -DCUBOID_ICP_STATS)."""
        w = np.zeros(32, np.uint64)
        self._ck(self.lib.cuboid_debug_counters(self._h, _ptr(w), 1 if reset else 0), "cuboid_debug_counters")
        return [int(x) for x in w]

    def measure_fp32_peak(self):
        a, b = C.c_double(0), C.c_double(0)
        self._ck(self.lib.cuboid_measure_fp32_peak(self._h, C.byref(a), C.byref(b)), "cuboid_measure_fp32_peak")
        return a.value, b.value


def _depth_desc(depth):
    if hasattr(depth, "data_ptr"):  # torch CPU tensor (e.g. pinned)
        if depth.is_cuda:
            raise ValueError("process_batch takes HOST buffers; use process_batch_device for device memory")
        n, hgt, w = depth.shape
        return int(n), int(hgt), int(w), C.c_void_p(depth.data_ptr()), depth
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    if d.ndim == 2:
        d = d[None]
    n, hgt, w = d.shape
    return n, hgt, w, d.ctypes.data_as(C.c_void_p), d
