import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(GOLD, "golden.json")))


@pytest.fixture(scope="session")
def tmpl30():
    from perception_b200 import pcd
    return pcd.load_pcd(os.path.join(GOLD, "template_cuboid_L200_W100_H30_3faces.pcd"))


@pytest.fixture(scope="session")
def tmpl100():
    from perception_b200 import pcd
    return pcd.load_pcd(os.path.join(GOLD, "template_cuboid_L200_W75_H100_3faces.pcd"))


@pytest.fixture(scope="session")
def frame0():
    from perception_b200 import synth
    return synth.depth_frame("cuboid1", 0)


@pytest.fixture(scope="session")
def params():
    from perception_b200.params import default_params
    return default_params("cuboid")


@pytest.fixture(scope="session")
def stage_data(frame0, params):
    """Oracle outputs of every stage for frame 0 (inputs for per-stage GPU parity)."""
    from oracle import pyoracle as O
    pts = O.unproject(frame0, params.fx, params.fy, params.cx, params.cy, params.depth_scale)
    pz, _ = O.passthrough(pts, 2, params.pass_z_min, params.pass_z_max)
    px, _ = O.passthrough(pz, 0, params.pass_x_min, params.pass_x_max)
    vg = O.voxel_grid(px, params.leaf)
    sac = O.sac_plane(vg["vox"], params.sac_threshold, params.sac_max_iter, params.sac_prob, params.sac_seed, 1)
    rem, _ = O.extract(vg["vox"], sac["inliers"], True)
    cidx, coff = O.cluster(rem, params.cluster_tol, params.cluster_min, params.cluster_max)
    return dict(all=pts, passed=px, vg=vg, sac=sac, remain=rem, cidx=cidx, coff=coff)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
