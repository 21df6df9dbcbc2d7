"""Seeded synthetic D435-shaped depth frames (the reference's bag files are git-ignored: SURVEY.md §0.5).

Scene kinds follow BASELINE.json's configs:
  cuboid1   table + one 200x100x30 mm cuboid, camera 0.45 m up, 55 deg down           (configs 0, 2)
  plane_var camera height U(0.35,0.6), tilt U(40,65) deg, one cuboid                  (config 1)
  bench     cuboid1 with mild camera jitter so every frame holds the cuboid           (bench workload)
  multi8    8 cuboids (200x100x30 and 200x75x100) >= 4 cm apart                       (config 3)
  hd720     1280x720, fx=fy=640 (free choice: no 720p intrinsics in the reference)    (config 4)
The ray caster itself is csrc/synth.c (host C, counter-based RNG); this module only draws scene parameters.
"""
import ctypes as C
import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import build as _build
from .params import D435_CX, D435_CY, D435_FX, D435_FY


class _Box(C.Structure):
    _fields_ = [("L", C.c_double), ("W", C.c_double), ("H", C.c_double), ("px", C.c_double), ("py", C.c_double),
                ("yaw", C.c_double)]


class Scene(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double),
                ("cy", C.c_double), ("cam_h", C.c_double), ("tilt", C.c_double), ("noise_sigma", C.c_double),
                ("invalid_frac", C.c_double), ("seed", C.c_uint64), ("n_boxes", C.c_int32), ("_pad", C.c_int32),
                ("box", _Box * 16)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        path = _build.SYNTH_LIB
        if not os.path.exists(path):
            _build.build_synth()
        _lib = C.CDLL(path)
        _lib.synth_depth_frame.argtypes = [C.POINTER(Scene), C.c_void_p]
        _lib.synth_depth_frame.restype = None
        assert _lib.synth_scene_size() == C.sizeof(Scene)
    return _lib


def make_scene(kind="cuboid1", seed=0):
    rng = np.random.default_rng([0xC0B01D, int(seed)])
    sc = Scene()
    sc.w, sc.h, sc.fx, sc.fy, sc.cx, sc.cy = 640, 480, D435_FX, D435_FY, D435_CX, D435_CY
    sc.noise_sigma, sc.invalid_frac, sc.seed = 0.0008, 0.005, int(seed)
    cam_h, tilt = 0.45, math.radians(55.0)
    if kind == "plane_var":
        cam_h, tilt = rng.uniform(0.35, 0.6), math.radians(rng.uniform(40.0, 65.0))
    elif kind == "bench":
        cam_h, tilt = rng.uniform(0.42, 0.48), math.radians(rng.uniform(52.0, 58.0))
    elif kind == "hd720":
        sc.w, sc.h, sc.fx, sc.fy, sc.cx, sc.cy = 1280, 720, 640.0, 640.0, 642.4656677, 360.6407318
        cam_h, tilt = 0.8, math.radians(60.0)
    sc.cam_h, sc.tilt = cam_h, tilt
    ax = cam_h / math.tan(tilt)  # where the optical axis meets the table (world X)
    if kind in ("cuboid1", "plane_var", "bench", "hd720"):
        sc.n_boxes = 1
        b = sc.box[0]
        b.L, b.W, b.H = 0.2, 0.1, 0.03
        b.px, b.py = ax + rng.uniform(-0.03, 0.03), rng.uniform(-0.04, 0.04)
        b.yaw = rng.uniform(0.0, 2.0 * math.pi)
    elif kind == "multi8":
        sc.n_boxes = 8
        # 2 rows (world X) x 4 columns (world Y), long side along X: >= 9 cm between neighbours
        k = 0
        for ix in range(2):
            for iy in range(4):
                b = sc.box[k]
                if (ix + iy) % 2 == 0:
                    b.L, b.W, b.H = 0.2, 0.1, 0.03
                else:
                    b.L, b.W, b.H = 0.2, 0.075, 0.1
                b.px = 0.27 + 0.30 * ix + rng.uniform(-0.005, 0.005)
                b.py = (iy - 1.5) * 0.2 + rng.uniform(-0.005, 0.005)
                b.yaw = rng.uniform(-0.08, 0.08)
                k += 1
    elif kind == "tallbox":   # one 200 x 75 x 100 mm cuboid seen corner-on: top face and two side faces (surface_normal_estimation)
        sc.n_boxes = 1
        b = sc.box[0]
        b.L, b.W, b.H = 0.2, 0.075, 0.1
        b.px, b.py = ax + rng.uniform(-0.02, 0.02), rng.uniform(-0.03, 0.03)
        b.yaw = math.radians(rng.uniform(25.0, 65.0))
    elif kind == "plane_only":
        sc.n_boxes = 0
    else:
        raise ValueError(kind)
    return sc


def depth_frame(kind="cuboid1", seed=0):
    lib = _load()
    sc = make_scene(kind, seed)
    out = np.empty((sc.h, sc.w), dtype=np.uint16)
    lib.synth_depth_frame(C.byref(sc), out.ctypes.data)
    return out


def depth_batch(kind, seeds, threads=None):
    """Frames for the given seeds -> uint16 [len(seeds), h, w]. ctypes releases the GIL, so threads scale."""
    lib = _load()
    seeds = list(seeds)
    sc0 = make_scene(kind, seeds[0] if seeds else 0)
    out = np.empty((len(seeds), sc0.h, sc0.w), dtype=np.uint16)

    def one(i):
        sc = make_scene(kind, seeds[i])
        lib.synth_depth_frame(C.byref(sc), out[i].ctypes.data)

    threads = threads or min(32, os.cpu_count() or 1)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(len(seeds))))
    return out
