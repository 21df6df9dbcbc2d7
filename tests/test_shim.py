"""The drop-in boundary from the reference's side: a ROS-free C++ translation unit with the patched callback bodies of
INTEGRATION.md (tests/shim/node_shim.cpp) and a plain-C one (tests/shim/abi_c.c) are compiled against include/cuboid_cuda.h,
linked to perception_b200/libcuboid_cuda.so and run. Without a GPU they exercise the host-side entry points and the loud
failure of cuboid_create; on the GPU box the callbacks run a PointCloud2-shaped blob through the library and everything they
would publish is compared with the ctypes path and the oracle."""
import ctypes as C
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "shim")
LIBDIR = os.path.join(ROOT, "perception_b200")


def _build(tmp, src, exe, cc):
    out = os.path.join(tmp, exe)
    cmd = [cc, "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(SHIM, src), "-o", out,
           "-L", LIBDIR, "-lcuboid_cuda", "-Wl,-rpath," + LIBDIR]
    if cc == "g++":
        cmd.insert(1, "-std=c++14")
    else:
        cmd.insert(1, "-std=c99")
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_c_and_cpp_translation_units_link_and_run(tmp_path):
    exe = _build(str(tmp_path), "abi_c.c", "abi_c", "gcc")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stdout
    exe = _build(str(tmp_path), "node_shim.cpp", "node_shim", "g++")
    r = subprocess.run([exe, "selftest"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "selftest ok" in r.stdout


@pytest.mark.gpu
def test_patched_callbacks_publish_what_the_ctypes_path_and_the_oracle_compute(tmp_path, frame0, tmpl30, params):
    from oracle import pyoracle as O
    from perception_b200 import api
    from perception_b200.params import FrameResult
    exe = _build(str(tmp_path), "node_shim.cpp", "node_shim", "g++")
    with api.CuboidCuda(params, max_points=640 * 480, max_batch=1) as h:
        h.set_template(0, tmpl30)
        cloud = h.unproject(frame0)
        n = len(cloud)
        rgb = (np.arange(n, dtype=np.uint32) * 2654435761 & 0xFFFFFF).astype(np.uint32)
        blob = np.zeros((n, 8), np.float32)               # point_step 32: x y z pad rgb pad pad pad (realsense-like padding)
        blob[:, :3] = cloud[:, :3]
        blob[:, 4] = rgb.view(np.float32)
        with open(tmp_path / "cloud.bin", "wb") as f:
            f.write(struct.pack("<6i", n, 32, 0, 4, 8, 16))
            f.write(blob.tobytes())
        t4 = np.ascontiguousarray(np.concatenate([tmpl30[:, :3], np.ones((len(tmpl30), 1), np.float32)], axis=1), dtype=np.float32)
        with open(tmp_path / "tmpl.bin", "wb") as f:
            f.write(struct.pack("<i", len(t4)))
            f.write(t4.tobytes())
        r = subprocess.run([exe, "run", str(tmp_path / "cloud.bin"), str(tmp_path / "tmpl.bin"), str(tmp_path / "out.bin")],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
        # the same calls through ctypes
        h.set_cloud_fields(16)
        pre = h.preprocess(blob, point_step=32, n=n)
        seg = h.segment_plane(pre["vox"])
        icp = h.icp(seg["remain"], 0)
        whole = h.process_cloud(blob, point_step=32, n=n)
        h.set_cloud_fields(-1)
    raw = open(tmp_path / "out.bin", "rb").read()
    o = 0

    def take(fmt):
        nonlocal o
        v = struct.unpack_from(fmt, raw, o)
        o += struct.calcsize(fmt)
        return v

    found, = take("<i")
    coeff = np.array(take("<4f"), np.float32)
    n_vox, n_pass, n_rem = take("<3i")
    rem = np.frombuffer(raw, np.float32, 4 * n_rem, o).reshape(n_rem, 4); o += 16 * n_rem
    T = np.frombuffer(raw, np.float32, 16, o).copy(); o += 64
    fitness, = take("<d")
    converged, iters, state = take("<3i")
    H = np.frombuffer(raw, np.float64, 16, o).reshape(4, 4); o += 128
    pose = np.frombuffer(raw, np.float64, 7, o); o += 56
    corners = np.frombuffer(raw, np.float32, 32, o).reshape(8, 4); o += 128
    success, = take("<i")
    if success:
        aligned = np.frombuffer(raw, np.float32, 4 * n_rem, o).reshape(n_rem, 4); o += 16 * n_rem
    res = FrameResult.from_buffer_copy(raw[o:o + C.sizeof(FrameResult)]); o += C.sizeof(FrameResult)
    sel = api.ObjectSelection.from_buffer_copy(raw[o:o + C.sizeof(api.ObjectSelection)]); o += C.sizeof(api.ObjectSelection)
    assert o == len(raw)
    # ground_plane_segmentation: coefficients + republished cloud, byte for byte, colour included
    assert found == 1 and np.array_equal(coeff.view(np.uint32), seg["coeff"].view(np.uint32))
    assert n_vox == len(pre["vox"]) and n_pass == pre["n_pass"] and n_rem == len(seg["remain"])
    assert rem.tobytes() == seg["remain"].tobytes()
    assert (rem[:, 3].view(np.uint32) >> 24 == 0).all() and len(np.unique(rem[:, 3].view(np.uint32))) > 10     # colours, not 1.0f
    # the oracle agrees on the geometry
    ref = O.process_frame(params, frame0, tmpl30)
    assert n_pass == ref.n_points and n_vox == ref.n_voxels and n_rem == ref.n_remain
    assert np.array_equal(coeff.view(np.uint32), np.array(list(ref.plane_coeff), np.float32).view(np.uint32))
    # iterative_closest_point: transform, fitness, pose, bounding box
    assert np.array_equal(T.view(np.uint32), icp["T"].reshape(16).view(np.uint32)) and fitness == icp["fitness"]
    assert (converged, iters, state) == (icp["converged"], icp["iters"], icp["state"]) and success == 1
    Hh, ph = api.pose_from_transform(T)
    assert np.array_equal(H, Hh) and np.array_equal(pose, ph)
    assert np.array_equal(corners, api.bbox_corners(H, 0.2, 0.1, 0.03))
    assert aligned.tobytes() == icp["aligned"].tobytes()
    # object_pose_detection service
    assert bytes(res) == bytes(whole)
    want = api.select_object(whole, len(tmpl30), 0.0004)
    assert bytes(sel) == bytes(want)
