"""The oracle against everything the reference pins for this path + self-consistency.

PARITY UNPINNED for PCL arithmetic: the reference holds no golden outputs of the hot path. What is pinned here:
make_cuboid.py templates (byte-identical), the image_geometry unprojection KAT, mt19937's known answer.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from perception_b200 import pcd, synth
from perception_b200.params import default_params

from conftest import GOLD, bits


def test_make_cuboid_byte_identical_to_reference_script(golden):
    for t in golden["templates"]:
        a = [float(x) for x in t["args"][1::2]]
        kw = dict(zip(["L", "W", "H", "density"], a))
        txt = pcd.pcd_text(pcd.make_cuboid(**kw))
        assert hashlib.sha256(txt.encode()).hexdigest() == t["sha256"]
        assert t["byte_identical_to_shipped"]
        assert open(os.path.join(GOLD, t["file"])).read() == txt


def test_pcd_readers_agree(tmpl30):
    o = O.load_pcd(os.path.join(GOLD, "template_cuboid_L200_W100_H30_3faces.pcd"))
    assert o.shape == (7250, 4) and np.array_equal(bits(o), bits(tmpl30))
    assert np.array_equal(bits(pcd.template_points(0.2, 0.1, 0.03, 0.002)), bits(tmpl30))
    assert np.all(tmpl30[:, 3] == 1.0)


def test_image_geometry_kat(golden):
    k = golden["image_geometry_kat"]
    d = np.zeros((480, 640), np.uint16)
    d[k["pixel"][1], k["pixel"][0]] = 1000
    p = O.unproject(d, k["fx"], k["fy"], k["cx"], k["cy"], 0.001)[k["pixel"][1] * 640 + k["pixel"][0]]
    assert np.allclose(p[:3], k["ray"], rtol=0, atol=2e-7)


def test_mt19937_known_answer():
    assert O.mt19937_nth(5489, 10000) == 4123659995


def test_passthrough_float_vs_double_edges():
    # 0.2f = 0.20000000298 > 0.2 is dropped, -0.2f dropped, 0.9f = 0.89999998 kept, zeros pass
    pts = np.array([[0.2, 0, 0.5, 1], [-0.2, 0, 0.5, 1], [0.1, 0, 0.9, 1], [0, 0, 0, 1], [np.nan, 0, 0.5, 1],
                    [0.1, np.inf, 0.5, 1], [0.19999999, 0, 0.5, 1]], np.float32)
    pz, iz = O.passthrough(pts, 2, 0.0, 0.9)
    px, ix = O.passthrough(pz, 0, -0.2, 0.2)
    assert list(iz[ix]) == [2, 3, 6]


def test_voxel_inverse_leaf_and_keys():
    assert np.float32(1.0) / np.float32(0.001) == np.float32(999.99994)
    assert np.float32(1.0) / np.float32(0.002) == np.float32(499.99997)
    rng = np.random.default_rng(1)
    pts = np.ones((5000, 4), np.float32)
    pts[:, :3] = rng.uniform(-0.2, 0.2, (5000, 3)).astype(np.float32)
    vg = O.voxel_grid(pts, 0.005)
    inv = np.float32(1.0) / np.float32(0.005)
    ijk = np.floor(pts[:, :3] * inv).astype(np.int64) - np.array(vg["min_b"])
    key = ijk[:, 0] + ijk[:, 1] * vg["div_b"][0] + ijk[:, 2] * vg["div_b"][0] * vg["div_b"][1]
    assert np.array_equal(key, vg["key_per_point"])
    assert np.array_equal(np.unique(key), vg["voxel_key"]) and vg["voxel_count"].sum() == 5000
    # canonical centroid = sequential float sum in ascending point order / count
    k0 = vg["voxel_key"][17]
    sel = pts[key == k0]
    s = sel[0, :3].copy()
    for q in sel[1:]:
        s = (s + q[:3]).astype(np.float32)
    assert np.array_equal(bits(s / np.float32(len(sel))), bits(vg["vox"][17, :3]))


def test_voxel_literal_vs_canonical_gap(stage_data):
    """SURVEY.md A.2 hazard: std::sort order inside a voxel moves centroids by last bits only."""
    lit = O.voxel_grid(stage_data["passed"], 0.005, O.LITERAL)
    can = stage_data["vg"]
    assert np.array_equal(lit["voxel_key"], can["voxel_key"]) and np.array_equal(lit["key_per_point"], can["key_per_point"])
    d = np.abs(lit["vox"] - can["vox"]).max()
    assert 0 < d < 1e-6


def test_plane_fit_recovers_planted_plane(stage_data):
    sac = stage_data["sac"]
    assert sac["found"] and 1 <= sac["iters"] <= 1000
    n = sac["coeff"][:3]
    t = np.radians(55.0)   # planted: cos(t)*y + sin(t)*z = 0.45 in the camera frame
    assert abs(np.linalg.norm(n) - 1) < 1e-5
    assert abs(abs(n @ np.array([0, np.cos(t), np.sin(t)])) - 1) < 2e-4
    assert abs(abs(sac["coeff"][3]) - 0.45) < 2e-3
    assert sac["k_margin"] > 1e-6      # the adaptive-k loop test is nowhere near an ulp of pow/log
    assert np.all(np.diff(sac["inliers"]) > 0)
    # replaying the recorded triplets reproduces the same answer
    again = O.sac_plane(stage_data["vg"]["vox"], triplets=sac["triplets"])
    assert np.array_equal(again["inliers"], sac["inliers"]) and np.array_equal(bits(again["coeff"]), bits(sac["coeff"]))


def test_sampler_matches_pcl_draw_rule(stage_data):
    """first draw = three swaps of shuffled[i] with shuffled[i + rnd() % (V - i)], rnd() = mt19937(12345)() >> 1"""
    V = len(stage_data["vg"]["vox"])
    sh = list(range(V))
    for i in range(3):
        r = O.mt19937_nth(12345, i + 1) >> 1
        j = i + r % (V - i)
        sh[i], sh[j] = sh[j], sh[i]
    assert list(stage_data["sac"]["triplets"][0]) == sh[:3]


def test_extract_and_cluster(stage_data):
    V = len(stage_data["vg"]["vox"])
    assert len(stage_data["remain"]) == V - len(stage_data["sac"]["inliers"])
    idx, off = stage_data["cidx"], stage_data["coff"]
    assert len(off) == 2 and off[1] >= 1000 and np.all(np.diff(idx) > 0)
    # two blobs 5 cm apart -> two clusters, larger first, indices ascending inside each
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 0.03, (300, 3)); b = rng.uniform(0, 0.03, (400, 3)) + [0.08, 0, 0]
    pts = np.ones((700, 4), np.float32); pts[:, :3] = np.vstack([a, b])
    perm = rng.permutation(700); pts = pts[perm]
    idx, off = O.cluster(pts, 0.02, 200, 25000)
    assert list(off) == [0, 400, 700]
    assert set(perm[idx[:400]]) == set(range(300, 700)) and np.all(np.diff(idx[:400]) > 0) and np.all(np.diff(idx[400:]) > 0)
    idx, off = O.cluster(pts, 0.02, 350, 25000)
    assert list(off) == [0, 400]
    idx, off = O.cluster(pts[:0], 0.02, 1, 25000)
    assert list(off) == [0]


def _rot(ax, ang):
    ax = np.asarray(ax, float) / np.linalg.norm(ax)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def test_icp_recovers_planted_pose(tmpl30):
    """north_star tolerances: 1e-4 rad, 1e-5 m, fitness 1e-6 — noise-free subset of the template, small motion."""
    rng = np.random.default_rng(5)
    sub = tmpl30[rng.choice(len(tmpl30), 1200, replace=False)]
    # the template is a 2 mm lattice: any in-plane lattice shift is also a perfect match, so stay well inside half a cell
    R, t = _rot([0.2, -0.5, 1.0], 0.004), np.array([0.0006, -0.0004, 0.0005])
    src = np.ones((1200, 4), np.float32)
    src[:, :3] = (sub[:, :3].astype(np.float64) - t) @ R   # src = R^T (tgt - t)  =>  T maps src -> tgt
    r = O.icp(src, tmpl30, rel_mse=1e-12, max_iter=200)
    assert r["converged"]
    Rg = r["T"][:3, :3].astype(np.float64)
    D = Rg @ R.T   # arccos(trace) is ill-conditioned near 0: take the angle from the skew part instead
    ang = np.linalg.norm([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]]) / 2
    assert ang < 1e-4 and np.abs(r["T"][:3, 3] - t).max() < 1e-5 and r["fitness"] < 1e-6


def test_icp_nn_equals_bruteforce_lowest_index(tmpl30):
    rng = np.random.default_rng(9)
    src = np.ones((257, 4), np.float32)
    src[:, :3] = rng.uniform(-0.1, 0.1, (257, 3)).astype(np.float32)
    src[:8, :3] = tmpl30[100:108, :3] + np.float32(0.001) * np.array([1, 0, 0], np.float32)  # exact lattice ties are possible
    r = O.icp(src, tmpl30, max_iter=1, trace_iters=1)
    d = src[:, None, :3] - tmpl30[None, :, :3]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32) + d[..., 2] * d[..., 2]
    assert np.array_equal(r["corr_trace"][0], d2.argmin(axis=1))   # argmin returns the first (lowest) index on ties


def test_icp_max_correspondence_distance_drops_pairs_like_correspondence_estimation(tmpl30):
    """setMaxCorrespondenceDistance (icp.cpp:175, commented out upstream). PCL's CorrespondenceEstimation::determineCorrespondences
    skips a source point whose nearest squared distance is greater than max_dist^2 ([PCL-recall]); the transformation is estimated
    from the remaining pairs in their order, and fewer than 3 pairs (min_number_correspondences_) end ICP unconverged. Checked
    against a brute-force float32 scan and a float64 Kabsch fit of the kept pairs."""
    rng = np.random.default_rng(21)
    src = np.ones((400, 4), np.float32)
    src[:, :3] = tmpl30[rng.integers(0, len(tmpl30), 400), :3] + rng.normal(0, 0.004, (400, 3)).astype(np.float32)
    maxd = 0.005
    r = O.icp(src, tmpl30, max_iter=1, max_corr_dist=maxd, trace_iters=1)
    d = src[:, None, :3] - tmpl30[None, :, :3]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32) + d[..., 2] * d[..., 2]
    nn, dmin = d2.argmin(axis=1), d2.min(axis=1)
    keep = ~(dmin.astype(np.float64) > maxd * maxd)
    assert 20 < keep.sum() < 380                                           # the distance rejects, and not everything
    assert np.array_equal(r["corr_trace"][0], np.where(keep, nn, -1))
    a, b = src[keep, :3].astype(np.float64), tmpl30[nn[keep], :3].astype(np.float64)   # Kabsch on the kept pairs only
    ca, cb = a.mean(0), b.mean(0)
    U, _, Vt = np.linalg.svd((b - cb).T @ (a - ca))
    S = np.diag([1.0, 1.0, np.sign(np.linalg.det(U) * np.linalg.det(Vt))])
    R = U @ S @ Vt
    assert np.allclose(r["T"][:3, :3], R, atol=2e-5) and np.allclose(r["T"][:3, 3], cb - R @ ca, atol=2e-6)
    far = O.icp(src, tmpl30, max_corr_dist=1e-7)
    assert (far["iters"], far["state"], far["converged"]) == (0, 5, 0) and np.array_equal(far["T"], np.eye(4, dtype=np.float32))
    same = O.icp(src, tmpl30, max_corr_dist=10.0)
    dflt = O.icp(src, tmpl30)
    assert same["corr_hash"] == dflt["corr_hash"] and np.array_equal(bits(same["T"]), bits(dflt["T"]))


def test_icp_edge_cases(tmpl30):
    r = O.icp(tmpl30[:2], tmpl30)
    assert not r["converged"] and r["iters"] == 0 and r["state"] == 5 and np.array_equal(r["T"], np.eye(4, dtype=np.float32))
    r = O.icp(tmpl30[:500], tmpl30, max_iter=3, rel_mse=0.0)
    assert r["converged"] and r["state"] in (1, 2, 3) and r["fitness"] < 1e-12


def test_pipeline_matches_golden_regression(golden, tmpl30):
    for g in golden["oracle_frames"]:
        d = synth.depth_frame(g["kind"], g["seed"])
        assert hashlib.sha256(d.tobytes()).hexdigest() == g["depth_sha256"]
        r = O.process_frame(default_params("cuboid"), d, tmpl30)
        for k in ("n_points", "n_voxels", "n_inliers_pre", "n_inliers", "sac_iterations", "n_remain", "n_clusters",
                  "points_hash", "voxel_key_hash", "voxel_hash", "inlier_hash", "remain_hash", "cluster_hash"):
            assert getattr(r, k) == g[k], k
        c = r.cluster[0]
        assert c.iterations == g["icp_iterations"] and c.corr_hash == g["icp_corr_hash"] and c.state == g["icp_state"]
        assert list(c.T) == g["icp_T"] and c.fitness == g["icp_fitness"]


def test_pipeline_literal_vs_canonical_within_tolerance(frame0, tmpl30, params):
    """The canonical choices (stable voxel order, tree sums, correctly-rounded trig) stay inside north_star's
    pose tolerances of the literal restatement on the reference frame (measured gap recorded in DESIGN.md)."""
    a = O.process_frame(params, frame0, tmpl30, mode=O.CANONICAL)
    b = O.process_frame(params, frame0, tmpl30, mode=O.LITERAL)
    assert a.voxel_key_hash == b.voxel_key_hash and a.n_voxels == b.n_voxels
    assert np.abs(np.array(list(a.plane_coeff)) - np.array(list(b.plane_coeff))).max() < 1e-6
    assert abs(a.n_inliers - b.n_inliers) <= 2
    Ta, Tb = np.array(list(a.cluster[0].T)).reshape(4, 4), np.array(list(b.cluster[0].T)).reshape(4, 4)
    assert abs(a.cluster[0].fitness - b.cluster[0].fitness) < 1e-6
    assert np.abs(Ta - Tb).max() < 5e-4


def test_pose_and_bbox():
    R, t = _rot([1, 2, 3], 0.7), np.array([0.1, -0.2, 0.5])
    T = np.eye(4, dtype=np.float32); T[:3, :3] = R; T[:3, 3] = t
    H, pose = O.pose_from_transform(T)
    assert np.allclose(H, np.linalg.inv(T.astype(np.float64)), atol=1e-12)
    q = pose[3:]
    assert abs(np.linalg.norm(q) - 1) < 1e-6
    x, y, z, w = q
    Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                   [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                   [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    assert np.allclose(Rq, H[:3, :3], atol=1e-6)
    c = O.bbox_corners(H, 0.2, 0.1, 0.03)
    assert c.shape == (8, 4) and np.allclose(c[:, :3].mean(0), H[:3, 3], atol=1e-6)


def test_bbox_filter_restates_the_reference_arithmetic(stage_data):
    """bbox_filter.cpp:30-51 is plain arithmetic in the reference's own source (no PCL inside): double accumulation stored to
    float, float divide, strict comparisons against the int rectangle. Checked against an independent numpy evaluation."""
    rem = stage_data["remain"]
    P = np.array([615.0, 0.0, 322.5, 0.0, 0.0, 615.5, 240.6, 0.0, 0.0, 0.0, 1.0, 0.0])   # a D435-like colour CameraInfo P
    bbox = (250, 120, 420, 330)
    out, idx = O.bbox_filter(rem, P, bbox)
    x, y, z = (rem[:, k].astype(np.float64) for k in range(3))
    with np.errstate(all="ignore"):
        u = (P[0] * x + P[1] * y + P[2] * z + P[3]).astype(np.float32)
        v = (P[4] * x + P[5] * y + P[6] * z + P[7]).astype(np.float32)
        w = (P[8] * x + P[9] * y + P[10] * z + P[11]).astype(np.float32)
        un, vn = u / w, v / w
    keep = (np.float32(bbox[0]) < un) & (un < np.float32(bbox[2])) & (np.float32(bbox[1]) < vn) & (vn < np.float32(bbox[3]))
    assert 0 < keep.sum() < len(rem)
    assert np.array_equal(idx, np.nonzero(keep)[0]) and np.array_equal(bits(out), bits(rem[keep]))
    # strict inequalities, w == 0 (NaN / inf never pass), READ_INFO-style empty rectangle
    edge = np.array([[0.0, 0.0, 1.0, 1.0], [1.0, 0.0, 1.0, 1.0], [0.5, 0.5, 0.0, 1.0], [0.0, 0.0, 0.0, 1.0]], np.float32)
    I = np.array([1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0])
    _, idx = O.bbox_filter(edge, I, (0, -1, 1, 1))
    assert list(idx) == []                      # u == x1 and u == x2 are outside, w == 0 gives inf / NaN
    _, idx = O.bbox_filter(edge, I, (-1, -1, 2, 1))
    assert list(idx) == [0, 1]
    assert len(O.bbox_filter(edge[:0], I, (0, 0, 1, 1))[1]) == 0


def _remain_of(kind, seed, params):
    from perception_b200 import synth
    d = synth.depth_frame(kind, seed)
    pts = O.unproject(d, params.fx, params.fy, params.cx, params.cy, params.depth_scale)
    pz, _ = O.passthrough(pts, 2, params.pass_z_min, params.pass_z_max)
    px, _ = O.passthrough(pz, 0, params.pass_x_min, params.pass_x_max)
    vg = O.voxel_grid(px, params.leaf)
    sac = O.sac_plane(vg["vox"], params.sac_threshold, params.sac_max_iter, params.sac_prob, params.sac_seed, 1)
    rem, _ = O.extract(vg["vox"], sac["inliers"], True)
    return rem, sac["coeff"]


def test_constrained_plane_models_respect_the_axis(params):
    """SACMODEL_PERPENDICULAR_PLANE / SACMODEL_PARALLEL_PLANE (surface_normal_estimation.cpp:118-123): the accepted plane's
    normal is within eps of the axis (resp. of perpendicular to it), and eps <= 0 degenerates to the plain plane model."""
    rem, table = _remain_of("tallbox", 1, params)
    axis = table[:3]
    top = O.sac_plane_model(rem, 1, axis, 0.1)
    rest, _ = O.extract(rem, top["inliers"], True)           # the node runs the parallel model on what the first plane left
    side = O.sac_plane_model(rest, 2, axis, 0.1)
    assert top["found"] and side["found"] and len(top["inliers"]) > 500 and len(side["inliers"]) > 300
    cosang = abs(float(np.dot(top["coeff"][:3], axis)) / np.linalg.norm(top["coeff"][:3]) / np.linalg.norm(axis))
    assert cosang > np.cos(0.1)
    sinang = abs(float(np.dot(side["coeff"][:3], axis)) / np.linalg.norm(side["coeff"][:3]) / np.linalg.norm(axis))
    assert sinang < np.sin(0.1) + 1e-6
    plain = O.sac_plane(rem, 0.015)
    for mt in (1, 2):
        same = O.sac_plane_model(rem, mt, axis, 0.0)
        assert np.array_equal(same["inliers"], plain["inliers"]) and np.array_equal(bits(same["coeff"]), bits(plain["coeff"]))
    # an axis no plane of the scene is perpendicular / parallel to within a tiny eps: every model fails isModelValid
    none = O.sac_plane_model(rem, 1, np.array([1.0, 0.0, 0.0], np.float32), 1e-4)
    assert len(none["inliers"]) == 0


def test_surface_normal_estimation_recovers_the_box_frame(params):
    """surface_normal_estimation.cpp:182-237 on a corner-on 200 x 75 x 100 mm box: three faces, a right-handed near-orthonormal
    frame whose z column is the table normal, and the point-count ordering of the node."""
    for seed in (0, 1, 2):
        rem, table = _remain_of("tallbox", seed, params)
        ok, r = O.surface_normals(rem, table[:3])
        assert ok and list(r.n_in)[0] == len(rem)
        n = list(r.n_plane)
        assert n[0] > 0 and n[1] > 0 and n[2] > 0 and sum(n) + r.n_left == len(rem)
        Rt = np.array(list(r.Rt), np.float64).reshape(4, 4)
        R = Rt[:3, :3]
        assert abs(np.linalg.det(R) - 1.0) < 0.02 and np.abs(R.T @ R - np.eye(3)).max() < 0.12
        cnt = [n[k] for k in r.order]
        assert cnt[0] >= cnt[1] >= cnt[2]
        zcol = R[:, 2]                                   # normals[0] = the plane with most points = the top face here
        assert abs(abs(float(zcol @ table[:3])) - 1.0) < 0.02
        assert 0.3 < Rt[2, 3] < 0.7                       # centroid in front of the camera


def test_surface_pose_host_function_matches_oracle():
    """cuboid_surface_pose (host arithmetic behind the C ABI, no GPU needed) against the oracle's restatement, including the
    node's exchange-sort quirk and the handedness flip."""
    import ctypes as C
    from perception_b200 import api
    L = api.load()
    rng = np.random.default_rng(5)
    for trial in range(40):
        coeff = rng.normal(size=(3, 4)).astype(np.float32)
        coeff[:, :3] /= np.linalg.norm(coeff[:, :3], axis=1, keepdims=True)
        mid = rng.uniform(-0.5, 0.5, size=(3, 3)).astype(np.float32)
        cnt = rng.integers(0, 6, size=3).astype(np.int32) if trial % 2 else rng.integers(0, 2000, size=3).astype(np.int32)
        Rt_o, ord_o = np.zeros(16, np.float32), np.zeros(3, np.int32)
        O.lib().orc_surface_pose(O._p(coeff), O._p(mid), O._p(cnt), O._p(Rt_o), O._p(ord_o))
        Rt_g, ord_g, pose = np.zeros(16, np.float32), np.zeros(3, np.int32), np.zeros(7, np.float64)
        L.cuboid_surface_pose(api._ptr(coeff), api._ptr(mid), api._ptr(cnt), api._ptr(Rt_g), api._ptr(ord_g), api._ptr(pose))
        assert np.array_equal(bits(Rt_g), bits(Rt_o)) and list(ord_g) == list(ord_o)
        assert np.allclose(pose[:3], Rt_g.reshape(4, 4)[:3, 3].astype(np.float64))
        assert abs(np.linalg.norm(pose[3:]) - 1.0) < 0.6    # tf's quaternion of a not-quite-orthonormal matrix


def _quat_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def scan_case(golden):
    """The reference's real marker scan, its template-frame copy and the capture pose published for them."""
    import os
    from perception_b200 import pcd
    from conftest import GOLD
    g = golden["object_scan"]
    cap = pcd.load_pcd(os.path.join(GOLD, g["capture"]))
    tpl = pcd.load_pcd(os.path.join(GOLD, g["template_frame"]))
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = _quat_R(g["rotation_xyzw"]), g["translation"]
    return cap, tpl, T


def test_icp_reproduces_the_published_capture_pose_of_a_real_scan(golden):
    """Known answer from the reference's own data (SURVEY 8c fixture 4): marker_ascii_tf.pcd IS marker_ascii.pcd moved by the
    pose in transforms.txt:74-83, so ICP of the capture onto the template-frame copy, started next to that pose, must return it (north_star tolerances: 1e-4 rad, 1e-5 m) with a vanishing fitness."""
    cap, tpl, T = scan_case(golden)
    assert len(cap) == len(tpl) == 597
    moved = cap[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    assert np.abs(moved - tpl[:, :3]).max() < 1e-7                  # the fixture's own relation
    # point-to-point ICP on a thin object slides into nearby local minima, so the start has to be inside the basin of the
    # one-to-one correspondence (point spacing of the scan is ~1 mm)
    for k, (ang, off) in enumerate(((0.002, 0.0005), (-0.001, -0.0003), (0.0015, 0.0006))):
        ax = np.eye(3)[k]
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        dR = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        G = np.eye(4)
        G[:3, :3], G[:3, 3] = dR @ T[:3, :3], T[:3, 3] + off
        r = O.icp(cap, tpl, guess=G.astype(np.float32), rel_mse=0.0004)
        assert r["converged"]
        Tr = r["T"].astype(np.float64)
        D = Tr[:3, :3] @ T[:3, :3].T
        rot_err = np.linalg.norm([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]]) / 2
        assert rot_err < 1e-4 and np.abs(Tr[:3, 3] - T[:3, 3]).max() < 1e-5, (rot_err, Tr[:3, 3] - T[:3, 3])
        assert r["fitness"] < 1e-12


def test_literal_vs_canonical_oracle_divergence_is_what_design_md_states(tmpl30):
    """north_star's tolerance (1e-4 rad, 1e-5 m, 1e-6 fitness) is established against the CANONICAL oracle only (the GPU is bit-exact
    against it). The oracle's own LITERAL mode (std::sort voxel order, sequential Eigen-style sums, host libm) is the closest thing to
    a second reading of PCL available offline; this test pins how far the two readings are apart on 256 bench frames, so that the
    claim in DESIGN.md section 5 stays a measured one. Integer outputs almost always agree; poses agree within tolerance when both
    modes stop ICP at the same iteration and differ by what the extra iterations move when last-bit differences shift the
    relative-MSE stop."""
    from concurrent.futures import ThreadPoolExecutor
    from perception_b200 import synth
    p = O.params_from(default_params("cuboid"))
    n = 256
    frames = synth.depth_batch("bench", range(n))

    def run(i):
        return O.process_frame(p, frames[i], tmpl30, mode=O.CANONICAL), O.process_frame(p, frames[i], tmpl30, mode=O.LITERAL)

    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        res = list(ex.map(run, range(n)))
    same_int = same_it = within = within_same = 0
    dt_same = []
    for a, b in res:
        si = all(getattr(a, k) == getattr(b, k) for k in ("n_points", "n_voxels", "voxel_key_hash", "n_inliers", "inlier_hash", "n_remain", "cluster_hash"))
        ca, cb = a.cluster[0], b.cluster[0]
        Ta, Tb = np.array(list(ca.T), np.float64).reshape(4, 4), np.array(list(cb.T), np.float64).reshape(4, 4)
        ang = np.linalg.norm(Ta[:3, :3] @ Tb[:3, :3].T - np.eye(3)) / np.sqrt(2.0)      # = the rotation angle for small angles
        tr = np.abs(Ta[:3, 3] - Tb[:3, 3]).max()
        ok = ang <= 1e-4 and tr <= 1e-5 and abs(ca.fitness - cb.fitness) <= 1e-6
        same_int += si
        same_it += ca.iterations == cb.iterations
        within += ok
        if ca.iterations == cb.iterations:
            within_same += ok
            dt_same.append(tr)
    # measured 2026-10 (this container, gcc 13, glibc 2.39): 255/256 integer-identical, 195/256 same stop iteration, 162/256 within
    # tolerance overall, 161/195 within tolerance when the stop iteration agrees; the bounds leave room for another libm
    assert same_int >= 0.98 * n
    assert same_it >= 0.65 * n
    assert within >= 0.5 * n
    assert within_same >= 0.75 * same_it
    assert np.median(dt_same) < 1e-5


def test_pcl_probe_and_harness_source():
    """no PCL here, so the probe must say so (and bench.py falls back to the restatement); the harness that would pin
    the oracle is committed source that follows the reference's call sequence."""
    from oracle import pcl_probe
    info = pcl_probe.find_pcl()
    assert set(info) == {"found", "how", "detail"}
    if not info["found"]:
        assert pcl_probe.build_harness(info) is None
    src = open(pcl_probe.HARNESS_SRC).read()
    for call in ("PassThrough<pcl::PCLPointCloud2>", "VoxelGrid<pcl::PCLPointCloud2>", "SACSegmentation<pcl::PointXYZ>", "ExtractIndices<pcl::PCLPointCloud2>",
                 "IterativeClosestPoint<pcl::PointXYZ, pcl::PointXYZ>", "setEuclideanFitnessEpsilon", "setTransformationEpsilon(1e-9)", "setMaximumIterations(5000)",
                 "ground_plane_segmentation.cpp:53-101", "iterative_closest_point.cpp:170-182"):
        assert call in src, call
