"""ASCII PCD v0.7 reader/writer and the cuboid template generator.

Replaces pcl::io::loadPCDFile<PointXYZ> (icp.cpp:159, opd.cpp:398) on the host side: the node (or this
module) parses the template ONCE and hands it to cuboid_set_template, instead of re-reading it in every
callback. ``make_cuboid`` restates cuboid_detection/templates/make_cuboid.py:38-66;
tests/golden holds the files that script writes and the test checks byte equality.
"""
import numpy as np

_HEADER = """# .PCD v0.7 - Point Cloud Data file format
VERSION 0.7
FIELDS x y z
SIZE 4 4 4
TYPE F F F
COUNT 1 1 1
WIDTH %d
HEIGHT 1
VIEWPOINT 0 0 0 1 0 0 0
POINTS %d
DATA ascii
"""


def make_cuboid(L=0.2, W=0.1, H=0.075, density=0.002):
    """Three faces (z=-H/2, y=-W/2, x=-L/2) of an LxWxH cuboid on a `density` lattice, float64 [n,3]."""
    X = np.arange(-L / 2.0, L / 2.0, density)
    Y = np.arange(-W / 2.0, W / 2.0, density)
    Z = np.arange(-H / 2.0, H / 2.0, density)

    def face(a, b):
        aa, bb = np.meshgrid(a, b)
        return aa.ravel(), bb.ravel()

    xy = face(X, Y)
    xz = face(X, Z)
    yz = face(Y, Z)
    f0 = np.stack([xy[0], xy[1], np.full(xy[0].shape, -H / 2.0)], axis=1)
    f1 = np.stack([xz[0], np.full(xz[0].shape, -W / 2.0), xz[1]], axis=1)
    f2 = np.stack([np.full(yz[0].shape, -L / 2.0), yz[0], yz[1]], axis=1)
    return np.vstack([f0, f1, f2])


def cuboid_filename(L, W, H):
    return "template_cuboid_L%d_W%d_H%d_3faces.pcd" % (L * 1000, W * 1000, H * 1000)


def pcd_text(points):
    pts = np.asarray(points)
    body = "".join("%f %f %f\n" % (p[0], p[1], p[2]) for p in pts)
    return _HEADER % (len(pts), len(pts)) + body


def write_pcd(path, points):
    with open(path, "w") as f:
        f.write(pcd_text(points))


def load_pcd(path):
    """-> float32 [n,4] (x,y,z,1.0), the pcl::PointXYZ layout (16-byte stride, data[3] = 1)."""
    with open(path, "r") as f:
        n = None
        fields = None
        for line in f:
            if line.startswith("FIELDS"):
                fields = line.split()[1:]
            elif line.startswith("POINTS"):
                n = int(line.split()[1])
            elif line.startswith("DATA"):
                if "ascii" not in line:
                    raise ValueError("only ASCII PCD is supported (the reference templates are ASCII)")
                break
        if n is None or fields is None or fields[:3] != ["x", "y", "z"]:
            raise ValueError("not an x y z PCD file: %s" % path)
        # parse each decimal literal straight to float32 (strtof semantics; not float64 -> float32 double rounding)
        rows = [ln.split() for ln in f if ln.strip()]
    if len(rows) < n:
        raise ValueError("PCD truncated")
    out = np.ones((n, 4), dtype=np.float32)
    flat = np.array([r[:3] for r in rows[:n]], dtype="U32")
    out[:, :3] = flat.astype(np.float32)
    return out


def template_points(L=0.2, W=0.1, H=0.03, density=0.002):
    """What loadPCDFile returns for the file make_cuboid.py writes: the %f-rounded text parsed to float32."""
    pts = make_cuboid(L, W, H, density)
    txt = np.array([["%f" % v for v in p] for p in pts], dtype="U32")
    out = np.ones((len(pts), 4), dtype=np.float32)
    out[:, :3] = txt.astype(np.float32)
    return out
