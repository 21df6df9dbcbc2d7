"""Run-time probe for a real PCL install — TEST INFRASTRUCTURE, used only by
`bench.py --impl reference`, the bench's cpu_baseline leg and tests/.

PCL / Eigen / FLANN / Boost / ROS are absent from the build container and cannot be installed offline, so the oracle is
"parity unpinned". Should a PCL ever be present where the bench runs (a prebuilt oracle/_ref/pcl_harness, `pkg-config
pcl_registration*`, or a PCLConfig.cmake), `find_pcl()` says so, `build_harness()` compiles oracle/pcl_harness/pcl_harness.cpp
(the reference's literal call sequence) into oracle/_ref/ and `run_harness()` executes it on one frame: the bench then reports
the real-PCL time (cpu_baseline.kind = "reference") and the restatement-vs-PCL difference with no further work."""
import glob
import json
import os
import shutil
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
HARNESS_SRC = os.path.join(HERE, "pcl_harness", "pcl_harness.cpp")
HARNESS_BIN = os.path.join(REF_DIR, "pcl_harness")
MODULES = ["pcl_registration", "pcl_segmentation", "pcl_filters", "pcl_io", "pcl_common"]


def find_pcl():
    """-> dict(found, how, detail). Looks for, in order: a prebuilt harness under oracle/_ref/ or baseline/_ref/, pkg-config modules
    pcl_registration[-1.x], a PCLConfig.cmake in the usual prefixes."""
    for cand in (HARNESS_BIN, os.path.join(os.path.dirname(HERE), "baseline", "_ref", "pcl_harness")):
        if os.path.isfile(cand) and os.access(cand, os.X_OK):
            return {"found": True, "how": "prebuilt", "detail": cand}
    pc = shutil.which("pkg-config")
    if pc:
        try:
            mods = subprocess.run([pc, "--list-all"], capture_output=True, text=True, timeout=20).stdout.split("\n")
            reg = sorted(m.split()[0] for m in mods if m.startswith("pcl_registration"))
            if reg:
                suffix = reg[-1][len("pcl_registration"):]
                return {"found": True, "how": "pkg-config", "detail": suffix}
        except Exception:
            pass
    pats = ["/usr/lib/*/cmake/pcl/PCLConfig.cmake", "/usr/lib/cmake/pcl*/PCLConfig.cmake", "/usr/share/pcl*/PCLConfig.cmake",
            "/usr/local/share/pcl*/PCLConfig.cmake", "/usr/local/lib/cmake/pcl*/PCLConfig.cmake", "/opt/ros/*/share/pcl*/PCLConfig.cmake",
            os.path.join(os.path.dirname(HERE), "baseline", "_ref", "**", "PCLConfig.cmake")]
    for pat in pats:
        hits = glob.glob(pat, recursive=True)
        if hits:
            return {"found": True, "how": "cmake", "detail": hits[0]}
    return {"found": False, "how": None, "detail": "no prebuilt harness, no pcl_registration pkg-config module, no PCLConfig.cmake"}


def build_harness(info=None):
    """Compile the harness with g++ and pkg-config flags (never the reference's own build system). Returns the binary path or None."""
    info = info or find_pcl()
    if not info["found"]:
        return None
    if info["how"] == "prebuilt":
        return info["detail"]
    if info["how"] != "pkg-config":
        return None            # a cmake-only install needs a hand-written flag set: report it, do not guess
    mods = [m + info["detail"] for m in MODULES]
    flags = subprocess.run(["pkg-config", "--cflags", "--libs"] + mods, capture_output=True, text=True, timeout=20)
    if flags.returncode != 0:
        return None
    os.makedirs(REF_DIR, exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++14", HARNESS_SRC, "-o", HARNESS_BIN] + flags.stdout.split()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    return HARNESS_BIN if r.returncode == 0 else None


def run_harness(binary, cloud_xyz, template_pcd, workdir, leaf=0.005, thr=0.015, fitness=0.0004, runs=10):
    """Run the literal PCL sequence on one cloud; returns the parsed dump (see pcl_harness.cpp) or None."""
    os.makedirs(workdir, exist_ok=True)
    cin, cout = os.path.join(workdir, "cloud.bin"), os.path.join(workdir, "pcl_out.bin")
    pts = np.ascontiguousarray(cloud_xyz[:, :3], dtype=np.float32)
    with open(cin, "wb") as f:
        f.write(struct.pack("<i", len(pts)))
        f.write(pts.tobytes())
    r = subprocess.run([binary, cin, template_pcd, cout, repr(leaf), repr(thr), repr(fitness), str(runs)], capture_output=True, text=True, timeout=1800)
    if r.returncode != 0:
        return None
    raw = open(cout, "rb").read()
    o = 0

    def take(fmt):
        nonlocal o
        v = struct.unpack_from(fmt, raw, o)
        o += struct.calcsize(fmt)
        return v

    n_pass, n_vox = take("<2i")
    vox = np.frombuffer(raw, np.float32, 3 * n_vox, o).reshape(n_vox, 3); o += 12 * n_vox
    found, = take("<i")
    coeff = np.array(take("<4f"), np.float32)
    n_inl, = take("<i")
    inl = np.frombuffer(raw, np.int32, n_inl, o); o += 4 * n_inl
    n_rem, = take("<i")
    rem = np.frombuffer(raw, np.float32, 3 * n_rem, o).reshape(n_rem, 3); o += 12 * n_rem
    T = np.frombuffer(raw, np.float32, 16, o).reshape(4, 4); o += 64
    fit, = take("<d")
    conv, = take("<i")
    ms_seg, ms_icp = take("<2d")
    return dict(n_pass=n_pass, vox=vox, found=found, coeff=coeff, inliers=inl, remain=rem, T=T, fitness=fit, converged=conv,
                ms_segmentation=ms_seg, ms_icp=ms_icp, stdout=json.loads(r.stdout.strip().splitlines()[-1]) if r.stdout.strip() else None)


def compare_with_oracle(dump, oracle_frame):
    """Restatement-vs-PCL difference for the bench line: counts, plane coefficients, pose and fitness."""
    c = oracle_frame.cluster[0]
    To = np.array(list(c.T), np.float64).reshape(4, 4)
    Tp = dump["T"].astype(np.float64)
    return {"n_points_equal": bool(dump["n_pass"] == oracle_frame.n_points), "n_voxels_equal": bool(len(dump["vox"]) == oracle_frame.n_voxels),
            "n_inliers_equal": bool(len(dump["inliers"]) == oracle_frame.n_inliers), "n_remain_equal": bool(len(dump["remain"]) == oracle_frame.n_remain),
            "plane_coeff_max_abs_diff": float(np.abs(dump["coeff"].astype(np.float64) - np.array(list(oracle_frame.plane_coeff))).max()),
            "rotation_diff_rad": float(np.linalg.norm(To[:3, :3] @ Tp[:3, :3].T - np.eye(3)) / np.sqrt(2.0)),
            "translation_diff_m": float(np.abs(To[:3, 3] - Tp[:3, 3]).max()), "fitness_diff": float(abs(c.fitness - dump["fitness"]))}
