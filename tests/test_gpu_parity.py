"""GPU parity: libcuboid_cuda (through the C ABI) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): voxel keys, RANSAC inlier index sets and ICP correspondences bit-exact;
final ICP pose within 1e-4 rad / 1e-5 m, fitness within 1e-6 (they are in fact bit-equal here because the
CUDA path reproduces the oracle's canonical operation order)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as O
from perception_b200 import api, synth
from perception_b200.params import default_params

from conftest import bits

pytestmark = pytest.mark.gpu

ROT_TOL, TRANS_TOL, FIT_TOL = 1e-4, 1e-5, 1e-6


@pytest.fixture(scope="module")
def cc(params, tmpl30, tmpl100):
    h = api.CuboidCuda(params, device=0, max_points=640 * 480, max_batch=8)
    h.set_template(0, tmpl30)
    h.set_template(1, tmpl100)
    yield h
    h.close()


def _pose_close(Tg, To):
    Tg, To = np.asarray(Tg, np.float64).reshape(4, 4), np.asarray(To, np.float64).reshape(4, 4)
    D = Tg[:3, :3] @ To[:3, :3].T
    ang = np.linalg.norm([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]]) / 2
    return ang < ROT_TOL and np.abs(Tg[:3, 3] - To[:3, 3]).max() < TRANS_TOL


def _same_frame(r, ref, check_icp=True):
    for k in ("status", "n_points", "n_voxels", "plane_found", "n_inliers_pre", "n_inliers", "sac_iterations", "sac_draws",
              "n_remain", "n_clusters", "points_hash", "voxel_key_hash", "voxel_hash", "inlier_hash", "remain_hash",
              "cluster_hash"):
        assert getattr(r, k) == getattr(ref, k), (k, getattr(r, k), getattr(ref, k))
    assert list(r.min_b) == list(ref.min_b) and list(r.div_b) == list(ref.div_b)
    assert np.array_equal(bits(list(r.plane_coeff)), bits(list(ref.plane_coeff)))
    if not check_icp:
        return
    for c in range(min(r.n_clusters, 16)):
        a, b = r.cluster[c], ref.cluster[c]
        assert (a.size, a.converged, a.iterations, a.state, a.best_guess, a.accepted) == \
               (b.size, b.converged, b.iterations, b.state, b.best_guess, b.accepted)
        assert a.corr_hash == b.corr_hash                       # every iteration's correspondences, bit-exact
        assert _pose_close(list(a.T), list(b.T)) and abs(a.fitness - b.fitness) < FIT_TOL
        assert np.array_equal(bits(list(a.T)), bits(list(b.T))) and a.fitness == b.fitness


def test_unproject_bit_exact(cc, frame0, stage_data):
    g = cc.unproject(frame0)
    assert np.array_equal(bits(g), bits(stage_data["all"]))


def test_preprocess_voxel_keys_and_centroids_bit_exact(cc, stage_data, params):
    g = cc.preprocess(stage_data["all"])
    vg = stage_data["vg"]
    assert g["n_pass"] == len(stage_data["passed"])
    assert np.array_equal(g["key_per_point"], vg["key_per_point"])     # the bit-exact voxel key
    assert np.array_equal(bits(g["vox"]), bits(vg["vox"]))
    # PointCloud2-style blob: 32-byte records, z/x/y shuffled, NaN + inf rows (PassThrough must drop them)
    n = 20000
    blob = np.zeros((n + 3, 8), np.float32)
    src = stage_data["all"][40000:40000 + n]
    blob[:n, 5], blob[:n, 1], blob[:n, 2] = src[:, 0], src[:, 1], src[:, 2]
    blob[n] = np.nan
    blob[n + 1, 5], blob[n + 1, 1], blob[n + 1, 2] = 0.1, np.inf, 0.5
    blob[n + 2, 5], blob[n + 2, 1], blob[n + 2, 2] = 0.2, 0.0, 0.5   # 0.2f > 0.2 -> dropped
    g = cc.preprocess(blob.view(np.uint8).reshape(-1), point_step=32, xoff=20, yoff=4, zoff=8, n=n + 3)
    pz, _ = O.passthrough(src, 2, params.pass_z_min, params.pass_z_max)
    px, _ = O.passthrough(pz, 0, params.pass_x_min, params.pass_x_max)
    ov = O.voxel_grid(px, params.leaf)
    assert g["n_pass"] == len(px) and np.array_equal(g["key_per_point"], ov["key_per_point"])
    assert np.array_equal(bits(g["vox"]), bits(ov["vox"]))


def test_segment_plane_inlier_sets_bit_exact(cc, stage_data):
    vox, sac = stage_data["vg"]["vox"], stage_data["sac"]
    g = cc.segment_plane(vox)
    assert g["found"] and g["iters"] == sac["iters"]
    assert np.array_equal(g["inliers_pre"], sac["inliers_pre"])
    assert np.array_equal(bits(g["coeff"]), bits(sac["coeff"]))
    assert np.array_equal(g["inliers"], sac["inliers"])
    assert np.array_equal(bits(g["remain"]), bits(stage_data["remain"]))
    # same seeded sample triplets handed in explicitly (north_star: "given the same seeded sample triplets")
    g2 = cc.segment_plane(vox, triplets=sac["triplets"])
    assert np.array_equal(g2["inliers"], sac["inliers"]) and np.array_equal(bits(g2["coeff"]), bits(sac["coeff"]))
    # a different explicit list: long enough that the adaptive stop decides, not the list end
    rng = np.random.default_rng(11)
    trips = np.stack([rng.choice(len(vox), 3, replace=False) for _ in range(64)]).astype(np.int32)
    o = O.sac_plane(vox, triplets=trips)
    g3 = cc.segment_plane(vox, triplets=trips)
    assert g3["iters"] == o["iters"] and np.array_equal(g3["inliers"], o["inliers"]) and np.array_equal(bits(g3["coeff"]), bits(o["coeff"]))


def test_segment_plane_edge_cases(cc):
    g = cc.segment_plane(np.zeros((0, 4), np.float32))
    assert not g["found"] and len(g["inliers"]) == 0 and len(g["remain"]) == 0
    two = np.array([[0, 0, 0.5, 1], [0.1, 0, 0.5, 1]], np.float32)
    g = cc.segment_plane(two)
    assert not g["found"] and np.array_equal(bits(g["remain"]), bits(two))   # ExtractIndices(negative) of nothing keeps all
    # collinear cloud: every sample fails isSampleGood -> no model, like the oracle
    line = np.ones((50, 4), np.float32)
    line[:, 0] = np.arange(50) * 0.01
    line[:, 1] = 0.0
    line[:, 2] = 0.5
    o = O.sac_plane(line)
    g = cc.segment_plane(line)
    assert g["found"] == o["found"] and np.array_equal(g["inliers"], o["inliers"])
    # rough cloud: many RANSAC iterations, several scoring rounds
    rng = np.random.default_rng(2)
    pts = np.ones((4000, 4), np.float32)
    pts[:, :3] = rng.uniform(-0.2, 0.2, (4000, 3)).astype(np.float32)
    pts[:1200, 2] = (0.3 + 0.002 * rng.standard_normal(1200)).astype(np.float32)
    o = O.sac_plane(pts)
    g = cc.segment_plane(pts)
    assert o["iters"] > 8 and g["iters"] == o["iters"]
    assert np.array_equal(g["inliers"], o["inliers"]) and np.array_equal(bits(g["coeff"]), bits(o["coeff"]))


def test_cluster_sets_equal(cc, stage_data):
    idx, off = cc.cluster(stage_data["remain"])
    assert np.array_equal(off, stage_data["coff"]) and np.array_equal(idx, stage_data["cidx"])
    rng = np.random.default_rng(3)
    blobs = [rng.uniform(0, 0.03, (n, 3)) + [0.07 * k, 0, 0] for k, n in enumerate([300, 400, 400, 150, 650])]
    pts = np.ones((sum(len(b) for b in blobs), 4), np.float32)
    pts[:, :3] = np.vstack(blobs)
    pts = pts[rng.permutation(len(pts))]
    oi, oo = O.cluster(pts, 0.02, 200, 25000)
    gi, go = cc.cluster(pts)
    assert list(go) == list(oo) and len(oo) == 5 and np.array_equal(gi, oi)   # 150-point blob filtered; 400/400 tie by smallest member
    gi, go = cc.cluster(pts[:0])
    assert list(go) == [0]


@pytest.mark.parametrize("n,box,cmin", [(1800, 0.28, 4), (10000, 0.5, 4), (30000, 0.72, 8)])   # the three memory tiers of k_cluster
def test_cluster_components_at_the_percolation_threshold(n, box, cmin):
    """Uniform random clouds with ~2.7 neighbours per point: components hinge on single borderline edges, cells
    are sparsely filled and hash buckets mix cells -- the fine-cell shortcut must still give the oracle's components."""
    from perception_b200.params import default_params
    p = default_params("cuboid")
    p.cluster_min, p.cluster_max = cmin, 25000
    h = api.CuboidCuda(p, device=0, max_points=32768, max_batch=1)
    try:
        for seed in range(3):
            rng = np.random.default_rng(100 + seed)
            pts = np.ones((n, 4), np.float32)
            pts[:, :3] = rng.uniform(-box / 2, box / 2, (n, 3))
            oi, oo = O.cluster(pts, p.cluster_tol, cmin, 25000)
            gi, go = h.cluster(pts)
            assert 20 < len(oo) <= 1024 and np.array_equal(go, oo) and np.array_equal(gi, oi)
        # lattices whose spacing is exactly / just under / just over the tolerance (strict '<' on the float distance)
        p.cluster_min = 1
        h.set_params(p)
        for spacing, n_expected in ((np.float32(0.02), None), (np.float32(0.019999), 1), (np.float32(0.020001), 1000)):
            g = np.arange(10, dtype=np.float32) * spacing
            pts = np.ones((1000, 4), np.float32)
            pts[:, :3] = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) - np.float32(0.11)
            oi, oo = O.cluster(pts, p.cluster_tol, 1, 25000)
            gi, go = h.cluster(pts)
            assert np.array_equal(go, oo) and np.array_equal(gi, oi)
            assert n_expected is None or len(oo) - 1 == n_expected
    finally:
        h.close()
    bad = default_params("cuboid")
    bad.cluster_tol = 1e-6
    with pytest.raises(api.CuboidError):
        api.CuboidCuda(bad, device=0, max_points=1024, max_batch=1)


def test_icp_per_iteration_correspondences_bit_exact(cc, stage_data, tmpl30):
    src = stage_data["remain"][stage_data["cidx"]]
    o = O.icp(src, tmpl30, trace_iters=128)
    g = cc.icp(src, 0, trace_iters=128)
    assert g["iters"] == o["iters"] and g["state"] == o["state"] and g["converged"] == o["converged"]
    n = o["iters"]
    assert np.array_equal(g["corr_trace"][:n], o["corr_trace"][:n])          # every iteration, every point
    assert np.array_equal(bits(g["T_trace"][:n]), bits(o["T_trace"][:n]))
    assert g["corr_hash"] == o["corr_hash"]
    assert _pose_close(g["T"], o["T"]) and abs(g["fitness"] - o["fitness"]) < FIT_TOL
    assert np.array_equal(bits(g["T"]), bits(o["T"])) and np.array_equal(bits(g["aligned"]), bits(o["aligned"]))


def test_icp_guesses_and_streamed_template(cc, stage_data, tmpl30):
    src = stage_data["remain"][stage_data["cidx"]]
    c = src[:, :3].mean(0)
    gs = []
    for k in range(4):
        a = k * np.pi / 2
        G = np.eye(4, dtype=np.float32)
        G[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
        G[:3, 3] = c - G[:3, :3] @ c
        gs.append(G)
    gs[0] = np.eye(4, dtype=np.float32)
    outs = [O.icp(src, tmpl30, guess=G) for G in gs]
    best = int(np.argmin([o["fitness"] for o in outs]))
    g = cc.icp(src, 0, guesses=np.stack(gs))
    assert g["best_guess"] == best and g["corr_hash"] == outs[best]["corr_hash"]
    assert np.array_equal(bits(g["T"]), bits(outs[best]["T"])) and g["fitness"] == outs[best]["fitness"]
    # a template larger than shared memory (21 400-point class) streams through the window: same answers
    rng = np.random.default_rng(4)
    big = np.ones((21400, 4), np.float32)
    big[:, :3] = rng.uniform(-0.1, 0.1, (21400, 3)).astype(np.float32)
    big[:7250] = tmpl30
    cc.set_template(2, big)
    o = O.icp(src, big, max_iter=6)
    p = default_params("cuboid")
    p.icp_max_iter = 6
    cc.set_params(p)
    try:
        g = cc.icp(src, 2)
    finally:
        cc.set_params(default_params("cuboid"))
    assert g["corr_hash"] == o["corr_hash"] and np.array_equal(bits(g["T"]), bits(o["T"])) and g["fitness"] == o["fitness"]


def test_icp_exact_ties_resolve_to_lowest_template_index(cc):
    """Dyadic lattice template + sources at exact cell/face/edge centres: 8-, 4- and 2-way exact float ties that
    straddle the kd-ordered chunks. The culled kernel must return the lowest ORIGINAL template index, like a
    brute-force scan in template order with strict '<' (and like the oracle's exact KD-tree)."""
    g = np.arange(16, dtype=np.float32) / np.float32(256.0)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    tm = np.ones((4096, 4), np.float32)
    tm[:, 0], tm[:, 1], tm[:, 2] = X.ravel(), Y.ravel(), Z.ravel()
    rng = np.random.default_rng(8)
    tm = tm[rng.permutation(4096)]                       # original order unrelated to space
    cells = rng.integers(0, 15, (900, 3)).astype(np.float32) / np.float32(256.0)
    h = np.float32(1.0 / 512.0)
    src = np.ones((900, 4), np.float32)
    src[:, :3] = cells
    src[:300, :3] += h                                   # cell centres: 8-way ties
    src[300:600, 0] += h; src[300:600, 1] += h           # face centres: 4-way ties
    src[600:, 2] += h                                    # edge midpoints: 2-way ties
    d = src[:, None, :3] - tm[None, :, :3]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32) + d[..., 2] * d[..., 2]
    assert ((d2 == d2.min(axis=1, keepdims=True)).sum(axis=1)[:300] == 8).all()
    cc.set_template(3, tm)
    p = default_params("cuboid")
    p.icp_max_iter = 2
    cc.set_params(p)
    try:
        o = O.icp(src, tm, max_iter=2, trace_iters=2)
        for cull in (1, 0):
            cc.set_option(api.OPT_ICP_CULL, cull)
            gq = cc.icp(src, 3, trace_iters=2)
            assert np.array_equal(gq["corr_trace"][0], d2.argmin(axis=1))      # numpy argmin = first = lowest index
            assert np.array_equal(gq["corr_trace"], o["corr_trace"]) and gq["corr_hash"] == o["corr_hash"]
            assert np.array_equal(bits(gq["T"]), bits(o["T"])) and gq["fitness"] == o["fitness"]
    finally:
        cc.set_option(api.OPT_ICP_CULL, 1)
        cc.set_params(default_params("cuboid"))


def test_icp_culling_is_exact_and_saves_work(cc, tmpl30, params):
    depth = synth.depth_batch("bench", [41, 42, 43])
    cc.set_option(api.OPT_ICP_CULL, 0)
    brute = cc.process_batch(depth)
    w0 = cc.icp_work()
    cc.set_option(api.OPT_ICP_CULL, 1)
    culled = cc.process_batch(depth)
    w1 = cc.icp_work()
    for a, b in zip(brute, culled):
        assert bytes(a) == bytes(b)
    assert w0[1] == w1[1] and w0[0] >= w0[1] and w1[0] < 0.5 * w0[0]


def test_icp_edge_cases(cc, tmpl30):
    o = O.icp(tmpl30[:2], tmpl30)
    g = cc.icp(tmpl30[:2], 0)
    assert (g["converged"], g["iters"], g["state"]) == (o["converged"], o["iters"], o["state"]) == (0, 0, 5)
    assert np.array_equal(g["T"], np.eye(4, dtype=np.float32)) and g["fitness"] == o["fitness"]
    for n in (3, 255, 256, 257, 1024, 1025):
        src = tmpl30[5:5 + n].copy()
        src[:, :3] += np.float32(0.0007)
        o = O.icp(src, tmpl30)
        g = cc.icp(src, 0)
        assert g["corr_hash"] == o["corr_hash"] and g["iters"] == o["iters"], n
        assert np.array_equal(bits(g["T"]), bits(o["T"])) and g["fitness"] == o["fitness"], n
    with pytest.raises(api.CuboidError) as e:
        cc.icp(tmpl30[:10], 5)
    assert e.value.status == api.E_NO_TEMPLATE


@pytest.mark.parametrize("maxd,near", [(10.0, False), (0.49, False), (0.45, False), (0.43, False), (0.008, True), (0.005, True), (0.003, True), (1e-6, True)])
def test_icp_finite_max_correspondence_distance(cc, stage_data, tmpl30, maxd, near):
    """icp.setMaxCorrespondenceDistance (iterative_closest_point.cpp:175, commented out upstream: the default is sqrt(DBL_MAX)).
    With a finite distance CorrespondenceEstimation drops the pairs beyond it, the transformation estimate and the MSE run over the
    rest in their original order, and fewer than 3 pairs end ICP unconverged. Every iteration's correspondences (-1 = dropped) and
    transformation, the final transformation, the aligned cloud and the fitness must equal the oracle's: from the identity (the cloud
    is half a metre from the template: distances that drop nothing, 227, 1356 and all of the 1410 pairs of the first iteration) and
    from a guess 6 mm off the aligned pose (distances that keep dropping tens to hundreds of pairs in every iteration)."""
    src = stage_data["remain"][stage_data["cidx"]]
    G = None
    if near:
        G = O.icp(src, tmpl30)["T"].copy()
        G[:3, 3] += np.float32(0.006)
    o = O.icp(src, tmpl30, guess=G, max_corr_dist=maxd, trace_iters=160)
    p = default_params("cuboid")
    p.icp_max_corr_dist = maxd
    cc.set_params(p)
    try:
        g = cc.icp(src, 0, guesses=None if G is None else [G], trace_iters=160)
    finally:
        cc.set_params(default_params("cuboid"))
    assert (g["iters"], g["state"], g["converged"]) == (o["iters"], o["state"], o["converged"])
    n = min(o["iters"], 160)
    assert np.array_equal(g["corr_trace"][:n], o["corr_trace"][:n])
    assert np.array_equal(bits(g["T_trace"][:n]), bits(o["T_trace"][:n]))
    assert g["corr_hash"] == o["corr_hash"]
    assert np.array_equal(bits(g["T"]), bits(o["T"])) and np.array_equal(bits(g["aligned"]), bits(o["aligned"]))
    assert g["fitness"] == o["fitness"]
    if maxd in (0.43, 1e-6):
        assert (g["iters"], g["state"], g["converged"]) == (0, 5, 0)
    elif maxd < 10.0:
        assert (o["corr_trace"][0] < 0).any()                 # the distance really rejects on this cloud


def test_process_batch_with_a_finite_max_correspondence_distance(tmpl30):
    """The whole-frame entry with the same parameter: results equal the oracle's frame by frame (one distance under which ICP still
    converges with part of the first iteration's pairs dropped, one under which it finds no correspondences at all)."""
    depth = synth.depth_batch("bench", [20, 21, 22])
    for maxd in (0.47, 0.2):
        p = default_params("cuboid")
        p.icp_max_corr_dist = maxd
        with api.CuboidCuda(p, max_points=640 * 480, max_batch=3) as h:
            h.set_template(0, tmpl30)
            res = h.process_batch(depth)
        for i in range(3):
            _same_frame(res[i], O.process_frame(p, depth[i], tmpl30))
        if maxd == 0.2:
            assert all(r.cluster[0].state == 5 and not r.cluster[0].converged for r in res)


def test_process_batch_matches_oracle_frame_by_frame(cc, tmpl30, params):
    seeds = [0, 1, 2, 3, 4]
    depth = synth.depth_batch("bench", seeds)
    res = cc.process_batch(depth)
    for i, s in enumerate(seeds):
        _same_frame(res[i], O.process_frame(params, depth[i], tmpl30))
    # intermediate arrays of one frame, fetched through the parity taps
    f = 3
    ref_pts = O.unproject(depth[f], params.fx, params.fy, params.cx, params.cy, params.depth_scale)
    pz, _ = O.passthrough(ref_pts, 2, params.pass_z_min, params.pass_z_max)
    px, _ = O.passthrough(pz, 0, params.pass_x_min, params.pass_x_max)
    vg = O.voxel_grid(px, params.leaf)
    assert np.array_equal(bits(cc.fetch(f, "points")), bits(px))
    assert np.array_equal(cc.fetch(f, "voxel_keys"), vg["key_per_point"])
    assert np.array_equal(bits(cc.fetch(f, "voxels")), bits(vg["vox"]))
    assert np.array_equal(cc.fetch(f, "voxel_counts"), vg["voxel_count"])
    sac = O.sac_plane(vg["vox"])
    assert np.array_equal(cc.fetch(f, "inliers"), sac["inliers"])


def test_process_cloud_entry_matches_depth_entry(cc, frame0, stage_data, tmpl30, params):
    r = cc.process_cloud(stage_data["all"])
    ref = O.process_cloud(params, stage_data["all"], 16, 0, 4, 8, len(stage_data["all"]), tmpl30)
    _same_frame(r, ref)
    _same_frame(r, O.process_frame(params, frame0, tmpl30))


def test_batch_edge_frames(cc, tmpl30, params):
    """empty / degenerate frames inside a batch: all-invalid depth, everything beyond pass limits, plane only."""
    depth = np.zeros((4, 480, 640), np.uint16)
    depth[1] = 2000                                   # z = 2 m: PassThrough z drops everything
    depth[2] = synth.depth_frame("plane_only", 5)     # no object: nothing to cluster
    depth[3] = synth.depth_frame("cuboid1", 9)
    res = cc.process_batch(depth)
    for i in range(4):
        _same_frame(res[i], O.process_frame(params, depth[i], tmpl30))
    assert res[1].n_points == 0 and res[1].n_voxels == 0 and res[1].plane_found == 0
    assert res[0].n_voxels == 1                       # all-zero depth collapses into the origin voxel
    assert res[2].n_clusters == 0 and res[3].n_clusters == 1


def test_multi_guess_rotations_about_centroid(tmpl30, params):
    depth = synth.depth_batch("bench", [21, 22])
    rots = []
    for k in range(4):
        a = k * np.pi / 2
        rots.append(np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32))
    rots[0] = np.eye(3, dtype=np.float32)
    rots = np.stack(rots)
    p = default_params("cuboid")
    p.n_guess, p.guess_mode = 4, 1
    with api.CuboidCuda(p, max_points=640 * 480, max_batch=2) as h:
        h.set_template(0, tmpl30)
        h.set_guesses(rots, mode=1)
        res = h.process_batch(depth)
    for i in range(2):
        _same_frame(res[i], O.process_frame(p, depth[i], tmpl30, guesses=rots))


def test_bench_64_rotation_guess_set_matches_oracle(tmpl30):
    """BASELINE config 3 as bench.py runs it: identity + 63 rotations about the cluster centroid, full iteration budget, two frames."""
    import bench
    rots = bench.guess_rotations()
    assert rots.shape == (64, 3, 3)
    depth = synth.depth_batch("bench", [11, 12])
    p = default_params("cuboid")
    p.n_guess, p.guess_mode = 64, 1
    with api.CuboidCuda(p, max_points=640 * 480, max_batch=2) as h:
        h.set_template(0, tmpl30)
        h.set_guesses(rots, mode=1)
        res = h.process_batch(depth)
    for i in range(2):
        _same_frame(res[i], O.process_frame(p, depth[i], tmpl30, guesses=rots))


def test_multi_object_frame_and_second_template(tmpl100):
    p = default_params("multi8")
    depth = synth.depth_batch("multi8", [0])
    with api.CuboidCuda(p, max_points=640 * 480, max_batch=1) as h:
        h.set_template(0, tmpl100)
        res = h.process_batch(depth)
    ref = O.process_frame(p, depth[0], tmpl100)
    assert ref.n_clusters >= 4
    _same_frame(res[0], ref)


def test_multi8_bench_config_four_frames_full_iterations(tmpl30):
    """BASELINE configs[3] as bench.py runs it (8 cuboids per frame, the 200x100x30 template, default iteration budget):
    four frames, every cluster's ICP trajectory against the oracle."""
    from concurrent.futures import ThreadPoolExecutor
    p = default_params("multi8")
    depth = synth.depth_batch("multi8", [1, 2, 3, 4])
    with api.CuboidCuda(p, max_points=640 * 480, max_batch=4) as h:
        h.set_template(0, tmpl30)
        res = h.process_batch(depth)
    with ThreadPoolExecutor(4) as ex:
        refs = list(ex.map(lambda i: O.process_frame(p, depth[i], tmpl30), range(4)))
    for i in range(4):
        assert refs[i].n_clusters == 8
        _same_frame(res[i], refs[i])


def test_object_detection_variant_small_leaf(tmpl30):
    """object_detection.launch: leaf 0.001 (inv_leaf = 999.99994f), threshold 0.01, extra PassThrough z [0, 0.75]."""
    p = default_params("object")
    p.icp_max_iter = 8
    depth = synth.depth_batch("cuboid1", [3])
    with api.CuboidCuda(p, max_points=640 * 480, max_batch=1) as h:
        h.set_template(0, tmpl30)
        res = h.process_batch(depth)
    _same_frame(res[0], O.process_frame(p, depth[0], tmpl30))


def test_hd720_small_leaf_stress_frame(tmpl30):
    """BASELINE configs[4]: 1280x720, 2 mm leaf: ~870k points, ~620k voxels, 29-bit keys (4 radix passes)."""
    p = default_params("hd720")
    p.icp_max_iter = 12
    depth = synth.depth_batch("hd720", [0])
    with api.CuboidCuda(p, max_points=1280 * 720, max_batch=1) as h:
        h.set_template(0, tmpl30)
        res = h.process_batch(depth)
    ref = O.process_frame(p, depth[0], tmpl30)
    assert ref.n_points > 800000 and ref.n_voxels > 500000
    _same_frame(res[0], ref)


def test_hd720_bench_config_four_frames_full_iterations(tmpl30):
    """BASELINE configs[4] as bench.py runs it: four 1280x720 frames, 2 mm leaf, default (5000) ICP iteration budget."""
    from concurrent.futures import ThreadPoolExecutor
    p = default_params("hd720")
    depth = synth.depth_batch("hd720", [1, 2, 3, 4])
    with api.CuboidCuda(p, max_points=1280 * 720, max_batch=4) as h:
        h.set_template(0, tmpl30)
        res = h.process_batch(depth)
    with ThreadPoolExecutor(4) as ex:
        refs = list(ex.map(lambda i: O.process_frame(p, depth[i], tmpl30), range(4)))
    for i in range(4):
        _same_frame(res[i], refs[i])


def test_stage_mask_segmentation_only(cc, params, tmpl30):
    depth = synth.depth_batch("plane_var", [5, 6])
    cc.set_option(api.OPT_STAGES, 3)
    try:
        res = cc.process_batch(depth)
    finally:
        cc.set_option(api.OPT_STAGES, 15)
    for i in range(2):
        ref = O.process_frame(params, depth[i], None)
        for k in ("n_points", "n_voxels", "n_inliers", "n_remain", "voxel_key_hash", "voxel_hash", "inlier_hash", "remain_hash"):
            assert getattr(res[i], k) == getattr(ref, k), k
        assert res[i].n_clusters == 0


def test_full_size_batch_properties(cc, tmpl30, params):
    """BASELINE-sized work: a 64-frame batch in 8-frame chunks must equal the same frames run one by one
    (determinism, chunk independence) and a sample must equal the oracle."""
    seeds = list(range(100, 164))
    depth = synth.depth_batch("bench", seeds)
    res = cc.process_batch(depth)
    again = cc.process_batch(depth[::-1].copy())
    for i in range(64):
        a, b = res[i], again[63 - i]
        assert bytes(a) == bytes(b)
    for i in (0, 31, 63):
        _same_frame(res[i], O.process_frame(params, depth[i], tmpl30))
    assert all(r.n_clusters == 1 and r.cluster[0].converged for r in res)


@pytest.mark.parametrize("cluster,threads", [(1, 512), (2, 512), (4, 512), (8, 512), (1, 1024), (4, 1024), (8, 1024), (1, 256), (8, 256)])
def test_fused_frontend_equals_unfused_kernels(tmpl30, params, cluster, threads, monkeypatch):
    """Stages 1a+1b as one cluster-per-frame kernel (frontend.cuh) against the unfused kernels: every result byte and
    every fetched intermediate array equal, for each cluster size, on normal, empty and degenerate frames."""
    depth = np.concatenate([synth.depth_batch("bench", [40, 41, 42, 43, 44]), np.zeros((3, 480, 640), np.uint16)])
    depth[6] = 2000
    depth[7] = synth.depth_frame("plane_only", 3)
    monkeypatch.setenv("CUBOID_FE_CLUSTER", str(cluster))
    monkeypatch.setenv("CUBOID_FE_THREADS", str(threads))
    out = {}
    for fused in (0, 1):
        with api.CuboidCuda(params, max_points=640 * 480, max_batch=8) as h:
            h.set_template(0, tmpl30)
            h.set_option(api.OPT_FRONTEND, fused)
            res = h.process_batch(depth)
            out[fused] = ([bytes(r) for r in res],
                          {w: [h.fetch(f, w) for f in (0, 4, 5, 7)] for w in ("points", "voxel_keys", "voxels", "voxel_counts")})
    assert out[0][0] == out[1][0]
    for w, arrs in out[0][1].items():
        for x, y in zip(arrs, out[1][1][w]):
            assert x.tobytes() == y.tobytes(), w
    ref = O.process_frame(params, depth[2], tmpl30)
    with api.CuboidCuda(params, max_points=640 * 480, max_batch=8) as h:
        h.set_template(0, tmpl30)
        _same_frame(h.process_batch(depth)[2], ref)


def test_bbox_filter_node_and_fused_predicate(cc, stage_data, tmpl30, params):
    """bbox_filter.cpp as a stand-alone stage and fused into the extraction of the pipeline."""
    rem = stage_data["remain"]
    P = np.array([615.0, 0.0, 322.5, 0.0, 0.0, 615.5, 240.6, 0.0, 0.0, 0.0, 1.0, 0.0])
    bbox = (200, 100, 460, 380)
    ref_pts, ref_idx = O.bbox_filter(rem, P, bbox)
    got_pts, got_idx = cc.bbox_filter(rem, P, bbox)
    assert 0 < len(ref_idx) < len(rem)
    assert np.array_equal(got_idx, ref_idx) and np.array_equal(bits(got_pts), bits(ref_pts))
    edge = np.array([[0.0, 0.0, 1.0, 1.0], [1.0, 0.0, 1.0, 1.0], [0.5, 0.5, 0.0, 1.0], [0.0, 0.0, 0.0, 1.0]], np.float32)
    I = np.array([1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0])
    for bb in ((0, -1, 1, 1), (-1, -1, 2, 1)):
        assert np.array_equal(cc.bbox_filter(edge, I, bb)[1], O.bbox_filter(edge, I, bb)[1])
    assert len(cc.bbox_filter(edge[:0], I, (0, 0, 1, 1))[1]) == 0
    # fused: ground_plane_segmentation -> bbox_filter -> clustering + ICP, against the oracle's stages composed the same way
    depth = synth.depth_batch("bench", [11, 12])
    cc.set_bbox_filter(P, bbox)
    try:
        res = cc.process_batch(depth)
        rem_gpu = [cc.fetch(i, "remain") for i in range(2)]
    finally:
        cc.set_bbox_filter(None)
    for i in range(2):
        pts = O.unproject(depth[i], params.fx, params.fy, params.cx, params.cy, params.depth_scale)
        pz, _ = O.passthrough(pts, 2, params.pass_z_min, params.pass_z_max)
        px, _ = O.passthrough(pz, 0, params.pass_x_min, params.pass_x_max)
        vg = O.voxel_grid(px, params.leaf)
        sac = O.sac_plane(vg["vox"], params.sac_threshold, params.sac_max_iter, params.sac_prob, params.sac_seed, 1)
        rem_o, _ = O.extract(vg["vox"], sac["inliers"], True)
        kept, _ = O.bbox_filter(rem_o, P, bbox)
        assert 0 < len(kept) < len(rem_o)
        assert np.array_equal(bits(rem_gpu[i]), bits(kept)) and res[i].n_remain == len(kept)
        cidx, coff = O.cluster(kept, params.cluster_tol, params.cluster_min, params.cluster_max)
        assert res[i].n_clusters == len(coff) - 1
        if len(coff) > 1:
            ref = O.icp(kept[cidx[coff[0]:coff[1]]], tmpl30, rel_mse=params.icp_rel_mse)
            c = res[i].cluster[0]
            assert (c.iterations, c.converged) == (ref["iters"], ref["converged"]) and c.corr_hash == ref["corr_hash"]
            assert np.array_equal(bits(list(c.T)), bits(ref["T"].reshape(-1))) and c.fitness == ref["fitness"]


def test_surface_normal_estimation_matches_oracle(cc, params):
    """the three constrained-plane RANSACs, centroids and pose of surface_normal_estimation.cpp, every
    float bit-equal to the oracle (same sampler stream, same isModelValid, same sequential sums)."""
    for kind, seed in (("tallbox", 0), ("tallbox", 1), ("tallbox", 2), ("bench", 7), ("plane_only", 2)):
        d = synth.depth_frame(kind, seed)
        pts = O.unproject(d, params.fx, params.fy, params.cx, params.cy, params.depth_scale)
        pz, _ = O.passthrough(pts, 2, params.pass_z_min, params.pass_z_max)
        px, _ = O.passthrough(pz, 0, params.pass_x_min, params.pass_x_max)
        vg = O.voxel_grid(px, params.leaf)
        sac = O.sac_plane(vg["vox"], params.sac_threshold, params.sac_max_iter, params.sac_prob, params.sac_seed, 1)
        rem, _ = O.extract(vg["vox"], sac["inliers"], True)
        ok, ref = O.surface_normals(rem, sac["coeff"][:3], 0.1, 0.015)
        got = cc.surface_normals(rem, sac["coeff"][:3], 0.1, 0.015)
        assert list(got.n_plane) == list(ref.n_plane) and list(got.found) == list(ref.found) and list(got.n_in) == list(ref.n_in)
        assert got.n_left == ref.n_left and list(got.order) == list(ref.order)
        for i in range(3):
            assert np.array_equal(bits(list(got.coeff[i])), bits(list(ref.coeff[i]))), (kind, seed, i)
            assert np.array_equal(bits(list(got.midpoint[i])), bits(list(ref.midpoint[i]))), (kind, seed, i)
        assert np.array_equal(bits(list(got.Rt)), bits(list(ref.Rt)))
        if kind == "tallbox":
            assert ok and min(got.n_plane) > 50
    # constrained single segmentation through the same kernel: inlier sets equal
    empty = cc.surface_normals(np.zeros((0, 4), np.float32), [0.0, 0.0, 1.0])
    assert list(empty.n_plane) == [0, 0, 0] and empty.n_left == 0


def test_icp_on_the_reference_real_scan_returns_the_published_pose(golden, params):
    """The reference's real marker scan pair (tests/golden, SURVEY 8c fixture 4) through the GPU ICP: bit-equal to the oracle
    and equal to the capture pose of transforms.txt:74-83 within the north_star tolerances."""
    from test_oracle import scan_case
    cap, tpl, T = scan_case(golden)
    K = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], float)
    dR = np.eye(3) + np.sin(0.002) * K + (1 - np.cos(0.002)) * K @ K
    G = np.eye(4)
    G[:3, :3], G[:3, 3] = dR @ T[:3, :3], T[:3, 3] + 0.0005
    with api.CuboidCuda(params, max_points=640 * 480, max_batch=1) as h:
        h.set_template(0, tpl)
        g = h.icp(cap, 0, guesses=[G.astype(np.float32)])
        far = h.icp(cap, 0)                                   # from identity: whatever minimum ICP slides into, same as the oracle
    ref = O.icp(cap, tpl, guess=G.astype(np.float32), rel_mse=params.icp_rel_mse)
    assert g["iters"] == ref["iters"] and g["corr_hash"] == ref["corr_hash"] and g["fitness"] == ref["fitness"]
    assert np.array_equal(bits(g["T"]), bits(ref["T"]))
    assert _pose_close(g["T"], T) and g["fitness"] < 1e-12
    ref_far = O.icp(cap, tpl, rel_mse=params.icp_rel_mse)
    assert far["iters"] == ref_far["iters"] and far["corr_hash"] == ref_far["corr_hash"] and np.array_equal(bits(far["T"]), bits(ref_far["T"]))


@pytest.mark.parametrize("knob,values", [("CUBOID_ICP_NSUB", ("1", "2", "3", "4")), ("CUBOID_ICP_SLICE", ("1", "8", "5000")),
                                         ("CUBOID_ICP_OUTWARD", ("0", "1")), ("CUBOID_ICP_QUEUED", ("0", "1")), ("CUBOID_ICP_TABLE", ("0", "1")), ("CUBOID_FE_HASH", ("0", "1")), ("CUBOID_FE_ONEPASS", ("0", "1")), ("CUBOID_FE_RUNS", ("0", "1")), ("CUBOID_FE_SOLO", ("0", "1")), ("CUBOID_ICP_LOCAL", ("0", "1")), ("CUBOID_ICP_SEEDGRID", ("0", "1")), ("CUBOID_FE_CLUSTER_SMALL", ("1", "8")), ("CUBOID_SAC_WIDE", ("0", "1", "2")), ("CUBOID_NNT_H_MM", ("0.7", "2.5")),
                                         ("CUBOID_PIPELINE", ("0", "1"))])
def test_execution_knobs_do_not_change_results(tmpl30, params, knob, values, monkeypatch):
    """How the work is scheduled must never show in the results: sub-workers per CTA, iterations per time slice, outward search
    vs descent from the root, queued (work-list) vs per-lane walk of the surviving subtrees, pipelined sub-chunks vs one chunk - every result byte of a 40-frame batch is the same."""
    depth = np.concatenate([synth.depth_batch("bench", list(range(60, 96))), synth.depth_batch("tallbox", [0, 1]),
                            np.zeros((2, 480, 640), np.uint16)])
    out = []
    for v in values:
        monkeypatch.setenv(knob, v)
        monkeypatch.setenv("CUBOID_SUB_BATCH", "16")
        with api.CuboidCuda(params, max_points=640 * 480, max_batch=40) as h:
            h.set_template(0, tmpl30)
            out.append([bytes(r) for r in h.process_batch(depth)])
    for o in out[1:]:
        assert o == out[0]


def test_taps_off_only_blanks_the_parity_hashes(tmpl30, params):
    """CUBOID_OPT_TAPS = 0 (what bench.py times) drops the parity instrumentation - the point / voxel / correspondence
    hashes read 0 - and changes no other result byte."""
    depth = np.concatenate([synth.depth_batch("bench", list(range(300, 312))), np.zeros((1, 480, 640), np.uint16)])
    with api.CuboidCuda(params, max_points=640 * 480, max_batch=13) as h:
        h.set_template(0, tmpl30)
        on = h.process_batch(depth)
        h.set_option(api.OPT_TAPS, 0)
        off = h.process_batch(depth)
    for a, b in zip(on, off):
        assert b.points_hash == 0 and b.voxel_key_hash == 0 and b.voxel_hash == 0 and b.cluster[0].corr_hash == 0
        assert b.inlier_hash == a.inlier_hash and b.remain_hash == a.remain_hash and b.cluster_hash == a.cluster_hash
        c = type(a).from_buffer_copy(bytes(a))
        c.points_hash = c.voxel_key_hash = c.voxel_hash = 0
        for k in range(len(c.cluster)):
            c.cluster[k].corr_hash = 0
        assert bytes(c) == bytes(b)
    assert on[0].points_hash != 0 and on[0].cluster[0].corr_hash != 0


def test_pipeline_option_and_concurrent_handles_return_the_same_bytes(tmpl30, params, monkeypatch):
    """CUBOID_OPT_PIPELINE only moves launches between streams: a handle gives the same bytes with the sub-chunk pipeline
    (one handle called in a loop) and with chunk-wide launches (several handles sharing the GPU), also when two handles
    run at the same time from two host threads, as bench.py's end-to-end arm drives them."""
    import threading
    monkeypatch.setenv("CUBOID_SUB_BATCH", "8")
    depth = np.concatenate([synth.depth_batch("bench", list(range(400, 420))), np.zeros((1, 480, 640), np.uint16)])
    hs = [api.CuboidCuda(params, max_points=640 * 480, max_batch=21) for _ in range(2)]
    try:
        for h in hs:
            h.set_template(0, tmpl30)
        ref = [bytes(r) for r in hs[0].process_batch(depth)]                 # pipelined sub-chunks of 8 frames
        out = [None, None]
        for h in hs:
            h.set_option(api.OPT_PIPELINE, 0)

        def drive(k):
            out[k] = [[bytes(r) for r in hs[k].process_batch(depth)] for _ in range(3)]

        thr = [threading.Thread(target=drive, args=(k,)) for k in range(2)]
        for t in thr:
            t.start()
        for t in thr:
            t.join()
        assert all(o == ref for k in range(2) for o in out[k])
    finally:
        for h in hs:
            h.close()


def test_rgb_field_is_carried_through_voxelgrid_and_extraction(frame0, tmpl30, params):
    """gps.cpp:49-112 on a PointCloud2 with x, y, z, rgb (point_step 20, rgb at 16): PassThrough / ExtractIndices copy whole
    records and VoxelGrid<PCLPointCloud2> (downsample_all_data_) averages r, g, b per voxel: float sums of the channel values
    (exact integers), float division by the count, truncation, repacked as r << 16 | g << 8 | b  [PCL-recall: voxel_grid.cpp
    'RGB special case']. xyz results must not change by a bit; the published remainder keeps the colour of its voxels."""
    with api.CuboidCuda(params, max_points=640 * 480, max_batch=1) as h:
        h.set_template(0, tmpl30)
        cloud = h.unproject(frame0)
        n = len(cloud)
        rng = np.random.default_rng(5)
        rgb = rng.integers(0, 1 << 24, n, dtype=np.uint32) | (rng.integers(0, 256, n, dtype=np.uint32) << 24)   # alpha byte must be dropped
        blob = np.zeros((n, 5), np.float32)
        blob[:, :3] = cloud[:, :3]
        blob[:, 3] = 7.0                                   # a foreign field between z and rgb
        blob[:, 4] = rgb.view(np.float32)
        plain = h.process_cloud(blob, point_step=20, n=n)
        vox_plain, rem_plain = h.fetch(0, "voxels"), h.fetch(0, "remain")
        assert (vox_plain[:, 3] == 1.0).all() and (rem_plain[:, 3] == 1.0).all()
        h.set_cloud_fields(16)
        col = h.process_cloud(blob, point_step=20, n=n)
        pts, keys = h.fetch(0, "points"), h.fetch(0, "voxel_keys")
        vox, rem, inl = h.fetch(0, "voxels"), h.fetch(0, "remain"), h.fetch(0, "inliers")
        h.set_cloud_fields(-1)
    a, b = type(plain).from_buffer_copy(bytes(plain)), type(col).from_buffer_copy(bytes(col))
    assert bytes(a) == bytes(b)                            # counts, hashes, plane, clusters, poses: unchanged
    assert np.array_equal(bits(vox[:, :3]), bits(vox_plain[:, :3])) and np.array_equal(bits(rem[:, :3]), bits(rem_plain[:, :3]))
    # survivors keep their own colour (PassThrough copies records), in order
    keep = np.isfinite(cloud[:, :3]).all(axis=1) & ~((cloud[:, 2] > 0.9) | (cloud[:, 2] < 0.0)) & ~((cloud[:, 0] > 0.2) | (cloud[:, 0] < -0.2))
    assert np.array_equal(pts[:, 3].view(np.uint32), rgb[keep])
    # per-voxel colour
    w = pts[:, 3].view(np.uint32)
    uk, inv, cnt = np.unique(keys, return_inverse=True, return_counts=True)
    assert len(uk) == len(vox)
    want = np.zeros(len(uk), np.uint32)
    for shift in (16, 8, 0):
        s = np.zeros(len(uk), np.float32)
        np.add.at(s, inv, ((w >> shift) & 255).astype(np.float32))
        want |= (s / cnt.astype(np.float32)).astype(np.int32).astype(np.uint32) << shift
    assert np.array_equal(vox[:, 3].view(np.uint32), want)
    # ExtractIndices(negative) republishes the voxel records that are not plane inliers, colour included
    mask = np.ones(len(vox), bool)
    mask[inl] = False
    assert np.array_equal(bits(rem), bits(vox[mask]))


def test_one_pass_front_end_equals_two_pass(tmpl30, params, monkeypatch):
    """Depth input without parity taps (what bench.py times): the front end writes its sort keys relative to STATIC bounds (pass-through
    limits, intrinsics) in ONE pass over the pixels instead of reducing min / max first, and sorts one record per RUN of consecutive
    survivors of a voxel instead of one per point (CUBOID_FE_RUNS). Every result byte and every fetched array must equal the two-pass
    path's, on normal, empty, constant and plane-only frames, on frames whose origin voxel (all zero-depth pixels) is made of tens of
    thousands of runs of every length, and on a near wall whose voxels hold more than a hundred points each; the tapped run must agree
    on all it shares."""
    depth = np.concatenate([synth.depth_batch("bench", [50, 51, 52, 53]), synth.depth_batch("plane_var", [7, 8]), np.zeros((3, 480, 640), np.uint16),
                            synth.depth_batch("bench", [54, 55, 56, 57])])
    depth[7] = 2000
    depth[8, 100:110, 300:320] = 450            # a handful of survivors in an otherwise empty frame
    rng = np.random.default_rng(5)
    depth[9][rng.random((480, 640)) < 0.5] = 0  # origin voxel: ~150k points in runs of 1, 2, 3 ... between surviving pixels
    depth[10][:, ::2] = 0                       # ... in runs of exactly one
    depth[11] = 300                             # a wall at 0.3 m: ~10 x 10 pixels per 5 mm voxel
    depth[11, ::7, ::5] = 0
    depth[12][200:, :] = 0                      # whole zero rows: runs of 32
    n = len(depth)
    out = {}
    for onepass, runs in (("0", "0"), ("1", "0"), ("1", "1"), ("1", "s")):
        monkeypatch.setenv("CUBOID_FE_ONEPASS", onepass)
        monkeypatch.setenv("CUBOID_FE_RUNS", "0" if runs == "s" else runs)
        monkeypatch.setenv("CUBOID_FE_SOLO", "0" if runs == "s" else "1")     # "1s": one-pass mode inside the general kernel instance
        with api.CuboidCuda(params, max_points=640 * 480, max_batch=n) as h:
            h.set_template(0, tmpl30)
            h.set_option(api.OPT_TAPS, 0)
            res = h.process_batch(depth)
            out[onepass + runs] = ([bytes(r) for r in res], {w: [h.fetch(f, w) for f in range(n)] for w in ("points", "voxels", "remain", "inliers")})
    for other in ("10", "11", "1s"):
        assert out["00"][0] == out[other][0]
        for w, arrs in out["00"][1].items():
            for f, arr in enumerate(arrs):
                assert np.array_equal(bits(arr) if arr.dtype == np.float32 else arr, bits(out[other][1][w][f]) if arr.dtype == np.float32 else out[other][1][w][f]), (other, w, f)
    out["1"] = out["11"]
    with api.CuboidCuda(params, max_points=640 * 480, max_batch=n) as h:    # taps on: two-pass by construction, checked against the oracle elsewhere
        h.set_template(0, tmpl30)
        tapped = h.process_batch(depth)
    for i in range(n):
        a = type(tapped[i]).from_buffer_copy(out["1"][0][i])
        for k in ("status", "n_points", "n_voxels", "n_inliers", "n_remain", "n_clusters", "inlier_hash", "remain_hash", "cluster_hash"):
            assert getattr(a, k) == getattr(tapped[i], k), (i, k)
        assert list(a.min_b) == list(tapped[i].min_b) and list(a.div_b) == list(tapped[i].div_b)


def test_candidate_table_at_cell_boundaries_and_beyond_the_band(cc, tmpl30):
    """The nearest-neighbour candidate table (nn_table.cuh) maps a query to a 1 mm cell with float arithmetic; its records are built
    for the cell widened by a margin. Adversarial queries: exactly on cell faces / edges / corners of the grid and one ulp either
    side of them, at every distance from the template surface up to and beyond the 25 mm band, plus points far outside the grid.
    Iteration-0 correspondences must equal a float32 brute-force scan in template order (first minimum = lowest index)."""
    tm = np.ascontiguousarray(tmpl30[:, :3], dtype=np.float32)
    rng = np.random.default_rng(11)
    h = np.float32(1e-3)
    org = (tm.min(axis=0).astype(np.float64) - 25.0 * 1e-3).astype(np.float32)        # the grid origin cuboid_set_template derives
    n = 6000
    base = tm[rng.integers(0, len(tm), n)] + rng.normal(0, 1, (n, 3)).astype(np.float32) * rng.choice([0.001, 0.004, 0.012, 0.03, 0.08], (n, 1)).astype(np.float32)
    cell = np.floor((base - org) / h).astype(np.float32)
    snapped = (org + cell * h).astype(np.float32)                                     # on a grid face in every axis (a cell corner)
    q = base.copy()
    kind = rng.integers(0, 4, n)
    for ax in range(3):
        on = (kind == ax) | (kind == 3)                                               # faces (one axis) and corners (all three)
        q[on, ax] = snapped[on, ax]
    nudge = rng.integers(-1, 2, (n, 3))
    q = np.where(nudge < 0, np.nextafter(q, np.float32(-np.inf)), np.where(nudge > 0, np.nextafter(q, np.float32(np.inf)), q)).astype(np.float32)
    q[:50] += np.float32(0.7)                                                         # far outside the grid: BVH path
    src = np.ones((n, 4), np.float32)
    src[:, :3] = q
    d = q[:, None, :] - tm[None, :, :]
    d2 = ((d[..., 0] * d[..., 0]) + d[..., 1] * d[..., 1]).astype(np.float32) + d[..., 2] * d[..., 2]
    want = d2.argmin(axis=1)
    p = default_params("cuboid")
    p.icp_max_iter = 1
    cc.set_params(p)
    try:
        cc.set_template(5, tmpl30)
        g = cc.icp(src, 5, trace_iters=1)
    finally:
        cc.set_params(default_params("cuboid"))
    assert np.array_equal(g["corr_trace"][0], want)


@pytest.mark.parametrize("seed", range(10))
def test_random_parameter_sets_match_oracle(seed, tmpl30, tmpl100):
    """Seeded parameter fuzz: leaf size, pass-through limits, RANSAC threshold, cluster tolerance / sizes, ICP epsilon, template, scene
    kind - one frame each through cuboid_process_batch (taps on: the two-pass front end) and once more without taps (the one-pass
    front end with its static key bounds), against the oracle."""
    rng = np.random.default_rng(1000 + seed)
    p = default_params("cuboid")
    p.leaf = float(rng.choice([0.003, 0.004, 0.005, 0.0065, 0.008, 0.011]))
    p.pass_z_max = float(rng.uniform(0.6, 1.2))
    p.pass_z_min = float(rng.choice([0.0, 0.05, 0.2]))
    xw = float(rng.uniform(0.12, 0.45))
    p.pass_x_min, p.pass_x_max = -xw, float(xw * rng.uniform(0.7, 1.0))
    p.sac_threshold = float(rng.choice([0.006, 0.01, 0.015, 0.02]))
    p.cluster_tol = float(rng.choice([0.012, 0.02, 0.03]))
    p.cluster_min = int(rng.choice([50, 200, 400]))
    p.icp_rel_mse = p.icp_fitness_gate = float(rng.choice([1e-4, 4e-4, 2e-3]))
    p.icp_max_iter = int(rng.choice([40, 5000]))
    p.use_cluster = int(rng.integers(0, 2))
    tm = tmpl30 if rng.integers(0, 2) else tmpl100
    kind = str(rng.choice(["bench", "plane_var", "tallbox", "multi8"]))
    depth = synth.depth_batch(kind, [int(rng.integers(0, 500))])
    ref = O.process_frame(p, depth[0], tm)
    with api.CuboidCuda(p, max_points=640 * 480, max_batch=1) as h:
        h.set_template(0, tm)
        res = h.process_batch(depth)
        _same_frame(res[0], ref)
        h.set_option(api.OPT_TAPS, 0)
        fast = h.process_batch(depth)
    a, b = type(ref).from_buffer_copy(bytes(fast[0])), type(ref).from_buffer_copy(bytes(res[0]))
    b.points_hash = b.voxel_key_hash = b.voxel_hash = 0
    for k in range(len(b.cluster)):
        b.cluster[k].corr_hash = 0
    assert bytes(a) == bytes(b)
