"""Host-side (numpy, fp64) spatial helpers used only at model-build time.

Quaternions are (w, x, y, z).  These mirror the numpy twins in the reference's
``genesis/utils/geom.py`` (quat_to_R, R_to_quat, transform_by_quat, ...) in behaviour; they are
never used on the step path.
"""

import numpy as np


def quat_mul(u, v):
    w1, x1, y1, z1 = u
    w2, x2, y2, z2 = v
    return np.array(
        [
            w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
            w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
            w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
            w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
        ]
    )


def quat_normalize(q):
    q = np.asarray(q, dtype=np.float64)
    n = np.linalg.norm(q)
    return q / n if n > 0 else np.array([1.0, 0.0, 0.0, 0.0])


def quat_inv(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat_to_R(q):
    w, x, y, z = quat_normalize(q)
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ]
    )


def R_to_quat(R):
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    if q[0] < 0:
        q = -q
    return quat_normalize(q)


def rotate(v, q):
    return quat_to_R(q) @ np.asarray(v, dtype=np.float64)


def transform_pos_quat(pos, quat, t_pos, t_quat):
    """(pos, quat) expressed in frame (t_pos, t_quat) -> parent frame."""
    return np.asarray(t_pos) + rotate(pos, t_quat), quat_normalize(quat_mul(t_quat, quat))


def inv_transform_pos_quat(pos, quat, t_pos, t_quat):
    qi = quat_inv(quat_normalize(t_quat))
    return rotate(np.asarray(pos) - np.asarray(t_pos), qi), quat_normalize(quat_mul(qi, quat))


def rpy_to_R(rpy):
    """URDF fixed-axis roll/pitch/yaw -> rotation matrix (R = Rz(y) Ry(p) Rx(r))."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def axis_angle_to_quat(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(axis)
    if n < 1e-14:
        return np.array([1.0, 0.0, 0.0, 0.0])
    axis = axis / n
    return np.concatenate([[np.cos(angle / 2)], np.sin(angle / 2) * axis])


def principal_inertia(I):
    """Diagonalise a symmetric inertia tensor: returns (diag[3] sorted descending, quat of the principal frame).

    MuJoCo's compiler (used by the reference for both MJCF and URDF, genesis/utils/mjcf.py:177,240-243) stores
    ``body_inertia`` (principal moments) and ``body_iquat``.  Dynamics only depend on R diag R^T.
    """
    I = 0.5 * (np.asarray(I, dtype=np.float64) + np.asarray(I, dtype=np.float64).T)
    off = abs(I[0, 1]) + abs(I[0, 2]) + abs(I[1, 2])
    if off < 1e-14 * max(1.0, np.trace(I)):
        return np.diag(I).copy(), np.array([1.0, 0.0, 0.0, 0.0])
    w, V = np.linalg.eigh(I)
    order = np.argsort(-w)
    w, V = w[order], V[:, order]
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]
    return w, R_to_quat(V)
