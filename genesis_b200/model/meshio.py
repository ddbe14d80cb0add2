"""Minimal collision-mesh loading for the model builder (build time only, host).

The reference loads meshes through MuJoCo/trimesh and convexifies them with ``trimesh.convex.convex_hull``
(qhull) followed by an optional decimation to 500 faces (genesis/utils/mesh.py:357,469;
genesis/engine/mesh.py:116).  trimesh/fast_simplification are not available here, so this module reads
STL/OBJ directly and takes the qhull convex hull via scipy (same qhull library).  Decimation is NOT
reproduced: hulls that exceed ``decimate_face_num`` faces are kept as-is (documented in DESIGN.md).
"""

import os
import struct

import numpy as np
from scipy.spatial import ConvexHull


def load_stl(path):
    with open(path, "rb") as f:
        data = f.read()
    # Binary STL: 80-byte header, uint32 n_tri, then 50 bytes per triangle.
    if len(data) >= 84:
        (n_tri,) = struct.unpack_from("<I", data, 80)
        if 84 + 50 * n_tri == len(data):
            tri = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n_tri, offset=84)
            return tri["v"].reshape(-1, 3).astype(np.float64)
    verts = []
    for line in data.decode("ascii", errors="ignore").splitlines():
        parts = line.split()
        if len(parts) == 4 and parts[0] == "vertex":
            verts.append([float(parts[1]), float(parts[2]), float(parts[3])])
    return np.asarray(verts, dtype=np.float64)


def load_obj(path):
    verts = []
    with open(path, "r", errors="ignore") as f:
        for line in f:
            if line.startswith("v "):
                p = line.split()
                verts.append([float(p[1]), float(p[2]), float(p[3])])
    return np.asarray(verts, dtype=np.float64)


def load_vertices(path):
    ext = os.path.splitext(path)[1].lower()
    if ext == ".stl":
        return load_stl(path)
    if ext == ".obj":
        return load_obj(path)
    raise ValueError(f"Unsupported collision mesh format: {path}")


def convex_hull(points):
    """Returns (verts[N,3], faces[M,3]) of the convex hull with outward-oriented faces."""
    points = np.asarray(points, dtype=np.float64)
    hull = ConvexHull(points)
    used = np.unique(hull.simplices)
    remap = -np.ones(len(points), dtype=np.int64)
    remap[used] = np.arange(len(used))
    verts = points[used]
    faces = remap[hull.simplices]
    # orient outward using qhull's facet equations
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    n = np.cross(b - a, c - a)
    flip = np.einsum("ij,ij->i", n, hull.equations[:, :3]) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    return verts, faces


def mesh_mass_properties(verts, faces):
    """Volume, center of mass and unit-density inertia (about the COM) of a closed triangle mesh."""
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    vol6 = np.einsum("ij,ij->i", a, np.cross(b, c))
    volume = vol6.sum() / 6.0
    if abs(volume) < 1e-18:
        return 0.0, verts.mean(axis=0), np.zeros((3, 3))
    com = ((a + b + c) * vol6[:, None]).sum(axis=0) / (24.0 * volume)
    # covariance integral over tetrahedra (origin, a, b, c)
    S = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            S[i, j] = (
                vol6
                * (
                    2 * (a[:, i] * a[:, j] + b[:, i] * b[:, j] + c[:, i] * c[:, j])
                    + a[:, i] * b[:, j] + a[:, j] * b[:, i]
                    + a[:, i] * c[:, j] + a[:, j] * c[:, i]
                    + b[:, i] * c[:, j] + b[:, j] * c[:, i]
                )
            ).sum() / 120.0
    S -= volume * np.outer(com, com)
    inertia = np.trace(S) * np.eye(3) - S
    return volume, com, inertia


def box_mesh(extents):
    """Box vertices in trimesh.creation.box order (x slowest, z fastest) + triangulated faces."""
    hx, hy, hz = 0.5 * np.asarray(extents, dtype=np.float64)
    verts = np.array([[sx * hx, sy * hy, sz * hz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
    _, faces = convex_hull(verts)
    return verts, faces


def icosphere(radius, subdivisions=2):
    t = (1.0 + 5.0**0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    verts = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    faces = list(f)
    for _ in range(subdivisions):
        cache, new_faces = {}, []

        def mid(i, j):
            key = (min(i, j), max(i, j))
            if key not in cache:
                m = verts[i] + verts[j]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new_faces += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = new_faces
    return np.asarray(verts) * radius, np.asarray(faces, dtype=np.int64)


def cylinder_mesh(radius, height, sections=32):
    """Cylinder along z, ``sections`` ring vertices per cap (trimesh.creation.cylinder default tessellation:
    the reference turns URDF/MJCF cylinders into such a mesh, genesis/utils/mjcf.py:529-536)."""
    theta = np.linspace(0.0, 2.0 * np.pi, sections, endpoint=False)
    ring = np.stack([radius * np.cos(theta), radius * np.sin(theta)], axis=1)
    verts = np.concatenate(
        [np.concatenate([ring, np.full((sections, 1), -0.5 * height)], axis=1),
         np.concatenate([ring, np.full((sections, 1), 0.5 * height)], axis=1)]
    )
    return convex_hull(verts)


def capsule_mesh(radius, height, count=(8, 12)):
    """Capsule along z (cylinder of ``height`` plus hemispherical caps): ``count`` = (latitude, longitude) sections like
    trimesh.creation.capsule, which the reference uses for collision capsules (genesis/utils/mjcf.py:517-526).  The
    longitude angles contain 0 / 90 / 180 / 270 degrees and the poles are vertices, so the rest-pose AABB is exact."""
    n_lat, n_lon = count
    phi = np.linspace(0.0, 2.0 * np.pi, n_lon, endpoint=False)
    verts = [[0.0, 0.0, 0.5 * height + radius], [0.0, 0.0, -0.5 * height - radius]]
    for k in range(1, n_lat // 2 + 1):
        th = 0.5 * np.pi * k / (n_lat // 2)  # pole -> equator
        rr, zz = radius * np.sin(th), radius * np.cos(th)
        for p in phi:
            verts.append([rr * np.cos(p), rr * np.sin(p), 0.5 * height + zz])
            verts.append([rr * np.cos(p), rr * np.sin(p), -0.5 * height - zz])
    return convex_hull(np.asarray(verts))
