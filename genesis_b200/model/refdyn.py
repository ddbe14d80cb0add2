"""Build-time fp64 numpy kinematics/CRBA used to derive invweight and meaninertia at qpos0.

Restates RigidSolver._init_invweight_and_meaninertia (genesis/engine/solvers/rigid/rigid_solver.py:492-655)
and kernel_init_meaninertia (abd/misc.py:103-123).  The kinematics follow
abd/forward_kinematics.py:337-736 and the CRBA abd/forward_dynamics.py:292-541, written with plain numpy
matrices (single configuration, no batching).  Host only; never on the step path.
"""

import numpy as np

from . import geom_np as gn
from ..constants import JOINT_TYPE


def forward_kinematics(m, qpos):
    nL = m.n_links
    pos = np.zeros((nL, 3))
    quat = np.zeros((nL, 4))
    xanchor = np.zeros((m.n_joints, 3))
    xaxis = np.zeros((m.n_joints, 3))
    for l in range(nL):
        p = m.links_parent[l]
        if p != -1:
            lp = pos[p] + gn.rotate(m.links_pos[l], quat[p])
            lq = gn.quat_normalize(gn.quat_mul(quat[p], m.links_quat[l]))
        else:
            lp, lq = m.links_pos[l].copy(), m.links_quat[l].copy()
        for j in range(m.links_joint_start[l], m.links_joint_end[l]):
            t = m.joints_type[j]
            qs, ds = m.joints_q_start[j], m.joints_dof_start[j]
            if t == JOINT_TYPE.FREE:
                xanchor[j] = qpos[qs: qs + 3]
                xaxis[j] = [0, 0, 1]
                lp = qpos[qs: qs + 3].copy()
                lq = gn.quat_normalize(qpos[qs + 3: qs + 7])
            elif t == JOINT_TYPE.REVOLUTE:
                axis = m.dofs_motion_ang[ds]
                xanchor[j] = gn.rotate(m.joints_pos[j], lq) + lp
                xaxis[j] = gn.rotate(axis, lq)
                lq = gn.quat_normalize(gn.quat_mul(lq, gn.axis_angle_to_quat(axis, qpos[qs] - m.qpos0[qs])))
                lp = xanchor[j] - gn.rotate(m.joints_pos[j], lq)
            elif t == JOINT_TYPE.PRISMATIC:
                axis = m.dofs_motion_vel[ds]
                xanchor[j] = gn.rotate(m.joints_pos[j], lq) + lp
                xaxis[j] = gn.rotate(axis, lq)
                lp = lp + xaxis[j] * (qpos[qs] - m.qpos0[qs])
            elif t == JOINT_TYPE.SPHERICAL:  # forward_kinematics.py:700-716: rotate about the anchor by the local quaternion
                xanchor[j] = gn.rotate(m.joints_pos[j], lq) + lp
                xaxis[j] = [0, 0, 1]
                lq = gn.quat_normalize(gn.quat_mul(lq, gn.quat_normalize(qpos[qs: qs + 4])))
                lp = xanchor[j] - gn.rotate(m.joints_pos[j], lq)
        pos[l], quat[l] = lp, lq
    return pos, quat, xanchor, xaxis


def com_and_cdof(m, pos, quat, xanchor, xaxis):
    nL, nD = m.n_links, m.n_dofs
    i_pos_w = np.zeros((nL, 3))
    i_quat = np.zeros((nL, 4))
    mass_sum = np.zeros(nL)
    com_sum = np.zeros((nL, 3))
    for l in range(nL):
        i_pos_w[l] = pos[l] + gn.rotate(m.links_inertial_pos[l], quat[l])
        i_quat[l] = gn.quat_normalize(gn.quat_mul(quat[l], m.links_inertial_quat[l]))
        r = m.links_root[l]
        mass_sum[r] += m.links_inertial_mass[l]
        com_sum[r] += m.links_inertial_mass[l] * i_pos_w[l]
    root_com = np.zeros((nL, 3))
    for l in range(nL):
        r = m.links_root[l]
        root_com[l] = com_sum[r] / mass_sum[r] if mass_sum[r] > 0 else i_pos_w[r]
    i_pos = i_pos_w - root_com
    cdof_ang = np.zeros((nD, 3))
    cdof_vel = np.zeros((nD, 3))
    for l in range(nL):
        for j in range(m.links_joint_start[l], m.links_joint_end[l]):
            t, ds = m.joints_type[j], m.joints_dof_start[j]
            off = root_com[l] - xanchor[j]
            if t == JOINT_TYPE.REVOLUTE:
                cdof_ang[ds] = xaxis[j]
                cdof_vel[ds] = np.cross(xaxis[j], off)
            elif t == JOINT_TYPE.PRISMATIC:
                cdof_vel[ds] = xaxis[j]
            elif t == JOINT_TYPE.FREE:
                cdof_vel[ds: ds + 3] = np.eye(3)
                RT = gn.quat_to_R(quat[l]).T
                for i in range(3):
                    cdof_ang[ds + 3 + i] = RT[i]
                    cdof_vel[ds + 3 + i] = np.cross(RT[i], off)
            elif t == JOINT_TYPE.SPHERICAL:  # body-frame angular dofs (forward_kinematics.py:559-562)
                RT = gn.quat_to_R(quat[l]).T
                for i in range(3):
                    cdof_ang[ds + i] = RT[i]
                    cdof_vel[ds + i] = np.cross(RT[i], off)
    return i_pos, i_quat, root_com, cdof_ang, cdof_vel


def mass_matrix(m, i_pos, i_quat, cdof_ang, cdof_vel):
    nL, nD = m.n_links, m.n_dofs
    crb_I = np.zeros((nL, 3, 3))
    crb_pos = np.zeros((nL, 3))
    crb_mass = np.zeros(nL)
    for l in range(nL):
        R = gn.quat_to_R(i_quat[l])
        x, y, z = i_pos[l]
        hhT = np.array([[y * y + z * z, -x * y, -x * z], [-x * y, x * x + z * z, -y * z], [-x * z, -y * z, x * x + y * y]])
        mass = m.links_inertial_mass[l]
        crb_I[l] = R @ m.links_inertial_i[l].reshape(3, 3) @ R.T + hhT * mass
        crb_pos[l] = i_pos[l] * mass
        crb_mass[l] = mass
    for l in range(nL - 1, -1, -1):
        p = m.links_parent[l]
        if p != -1:
            crb_I[p] += crb_I[l]
            crb_pos[p] += crb_pos[l]
            crb_mass[p] += crb_mass[l]
    f_ang = np.zeros((nD, 3))
    f_vel = np.zeros((nD, 3))
    for d in range(nD):
        l = m.dofs_link[d]
        f_ang[d] = crb_I[l] @ cdof_ang[d] + np.cross(crb_pos[l], cdof_vel[d])
        f_vel[d] = crb_mass[l] * cdof_vel[d] - np.cross(crb_pos[l], cdof_ang[d])
    M = np.zeros((nD, nD))
    for i in range(nD):
        for j in range(i + 1):
            if m.dofs_entity[i] == m.dofs_entity[j]:
                M[i, j] = (f_ang[i] @ cdof_ang[j] + f_vel[i] @ cdof_vel[j]) * m.mass_parent_mask[i, j]
                M[j, i] = M[i, j]
    M[np.arange(nD), np.arange(nD)] += m.dofs_armature
    return M


def invweight_and_meaninertia(m):
    nL, nD = m.n_links, m.n_dofs
    links_invweight = np.zeros((nL, 2))
    dofs_invweight = np.zeros(nD)
    if nD == 0:
        return links_invweight, dofs_invweight, 1.0
    pos, quat, xanchor, xaxis = forward_kinematics(m, m.qpos0)
    i_pos, i_quat, root_com, cdof_ang, cdof_vel = com_and_cdof(m, pos, quat, xanchor, xaxis)
    M = mass_matrix(m, i_pos, i_quat, cdof_ang, cdof_vel)
    Minv = np.linalg.inv(M)
    for l in range(nL):
        jacp = np.zeros((3, nD))
        jacr = np.zeros((3, nD))
        a = l
        while a != -1:
            for d in range(m.links_dof_start[a], m.links_dof_end[a]):
                jacp[:, d] = cdof_vel[d] + np.cross(cdof_ang[d], i_pos[l])
                jacr[:, d] = cdof_ang[d]
            a = m.links_parent[a]
        jac = np.concatenate([jacp, jacr], axis=0)
        A = np.diag(jac @ Minv @ jac.T)
        tran, rot = A[:3].mean(), A[3:].mean()
        eps = float(np.finfo(np.float32).eps)
        if tran < eps and rot > eps:
            tran = rot
        elif rot < eps and tran > eps:
            rot = tran
        links_invweight[l] = (tran, rot)
    dMinv = np.diag(Minv)
    for j in range(m.n_joints):
        t, ds = m.joints_type[j], m.joints_dof_start[j]
        if t == JOINT_TYPE.FREE:
            dofs_invweight[ds: ds + 3] = dMinv[ds: ds + 3].mean()
            dofs_invweight[ds + 3: ds + 6] = dMinv[ds + 3: ds + 6].mean()
        elif t == JOINT_TYPE.SPHERICAL:
            dofs_invweight[ds: ds + 3] = dMinv[ds: ds + 3].mean()
        elif t != JOINT_TYPE.FIXED:
            dofs_invweight[ds] = dMinv[ds]
    # kernel_init_meaninertia: running sum divided by n_dofs after every entity (abd/misc.py:113-121)
    meaninertia = 0.0
    for e in range(m.n_entities):
        for d in range(m.entities_dof_start[e], m.entities_dof_end[e]):
            meaninertia += M[d, d]
        meaninertia /= nD
    return links_invweight, dofs_invweight, meaninertia


def neutral_geom_vertices(m, g):
    pos, quat, _, _ = forward_kinematics(m, m.qpos0)
    l = m.geoms_link[g]
    gp = pos[l] + gn.rotate(m.geoms_pos[g], quat[l])
    gq = gn.quat_normalize(gn.quat_mul(quat[l], m.geoms_quat[g]))
    v = m.verts[m.geoms_vert_start[g]: m.geoms_vert_end[g]]
    return v @ gn.quat_to_R(gq).T + gp
