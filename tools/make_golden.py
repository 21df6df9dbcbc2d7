"""Regenerates tests/golden/ — run in the build container, where /root/reference exists (it does not on the GPU box).

1. Runs the reference's own cuboid_detection/templates/make_cuboid.py (unmodified, from /root/reference) for the two
   shipped centred templates and stores what it writes; also records whether each output is byte-identical to the
   .pcd the reference ships (it is: SURVEY.md probe B3).
2. Records the image_geometry unprojection known answer (vision_opencv/image_geometry/test/utest.cpp:25-27,56-65).
3. Stores oracle outputs for two seeded frames (a regression pin of the oracle against itself, NOT a reference pin).
4. Copies the reference's real object scans with a published answer (SURVEY.md 8c fixture 4): object_detection/templates/
   marker_ascii.pcd (the capture, camera frame) and marker_ascii_tf.pcd (the same 597 points in the template frame), plus the
   capture pose listed for them in object_detection/templates/transforms.txt:74-83. These are data, not code.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    os.makedirs(GOLD, exist_ok=True)
    script = os.path.join(REF, "cuboid_detection", "templates", "make_cuboid.py")
    meta = {"generator": script, "templates": []}
    cases = [
        (["-L", "0.2", "-W", "0.1", "-H", "0.03"], "template_cuboid_L200_W100_H30_3faces.pcd"),
        (["-L", "0.2", "-W", "0.075", "-H", "0.1", "-d", "0.005"], "template_cuboid_L200_W75_H100_3faces.pcd"),
    ]
    for args, name in cases:
        with tempfile.TemporaryDirectory() as td:
            subprocess.run([sys.executable, script] + args, cwd=td, check=True, capture_output=True)
            out = os.path.join(td, name)
            shipped = os.path.join(REF, "cuboid_detection", "templates", name)
            same = open(out, "rb").read() == open(shipped, "rb").read()
            shutil.copy(out, os.path.join(GOLD, name))
            meta["templates"].append({"file": name, "args": args, "sha256": sha(out), "byte_identical_to_shipped": same,
                                      "shipped_sha256": sha(shipped)})
    meta["image_geometry_kat"] = {
        "source": "vision_opencv/image_geometry/test/utest.cpp:25-27,56-65",
        "fx": 295.53402059708782, "fy": 295.53402059708782, "cx": 285.55760765075684, "cy": 223.29617881774902,
        "pixel": [100, 100], "ray": [-0.62787224048135637, -0.41719792045817677, 1.0],
    }
    from oracle import pyoracle as O
    from perception_b200 import pcd, synth
    from perception_b200.params import default_params
    tm = pcd.load_pcd(os.path.join(GOLD, cases[0][1]))
    frames = []
    for kind, seed in (("cuboid1", 0), ("bench", 7)):
        d = synth.depth_frame(kind, seed)
        r = O.process_frame(default_params("cuboid"), d, tm)
        c = r.cluster[0]
        frames.append({"kind": kind, "seed": seed, "depth_sha256": hashlib.sha256(d.tobytes()).hexdigest(),
                       "n_points": r.n_points, "n_voxels": r.n_voxels, "min_b": list(r.min_b), "div_b": list(r.div_b),
                       "n_inliers_pre": r.n_inliers_pre, "n_inliers": r.n_inliers, "sac_iterations": r.sac_iterations,
                       "plane_coeff_bits": [int.from_bytes(__import__("struct").pack("<f", v), "little") for v in r.plane_coeff],
                       "n_remain": r.n_remain, "n_clusters": r.n_clusters, "points_hash": r.points_hash,
                       "voxel_key_hash": r.voxel_key_hash, "voxel_hash": r.voxel_hash, "inlier_hash": r.inlier_hash,
                       "remain_hash": r.remain_hash, "cluster_hash": r.cluster_hash, "icp_iterations": c.iterations,
                       "icp_state": c.state, "icp_fitness": c.fitness, "icp_T": list(c.T), "icp_corr_hash": c.corr_hash})
    meta["oracle_frames"] = frames
    scans = os.path.join(REF, "object_detection", "templates")
    for name in ("marker_ascii.pcd", "marker_ascii_tf.pcd"):
        shutil.copy(os.path.join(scans, name), os.path.join(GOLD, name))
        os.chmod(os.path.join(GOLD, name), 0o644)
    meta["object_scan"] = {
        "source": "object_detection/templates/marker_ascii.pcd, marker_ascii_tf.pcd, transforms.txt:74-83",
        "capture": "marker_ascii.pcd", "template_frame": "marker_ascii_tf.pcd",
        "sha256": {n: sha(os.path.join(GOLD, n)) for n in ("marker_ascii.pcd", "marker_ascii_tf.pcd")},
        "translation": [-0.0108, -0.4808, -0.3096],
        "rotation_xyzw": [-0.342422592276, 0.0607978149491, 0.0275301284879, 0.937172602044],
        "relation": "template_frame = R(rotation) * capture + translation (verified to 3e-8 when the fixture was made)",
    }
    json.dump(meta, open(os.path.join(GOLD, "golden.json"), "w"), indent=1)
    print(json.dumps(meta["templates"], indent=1))


if __name__ == "__main__":
    main()
