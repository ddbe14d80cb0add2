// gsr_kernels.cuh -- hand-written sm_100a kernels of the batched rigid-body step.
//
// Mapping: one CUDA thread owns one environment for a whole phase ("lane = env").  All per-env arrays are rows of
// B consecutive environments (gsr_types.h), so every load/store of a warp is one fully coalesced 128-byte
// transaction; the model is read from __constant__ memory with warp-uniform addresses (single broadcast).  The
// ~65 Quadrants launches of one reference substep are fused into four kernels:
//   k_dynamics   : CRBA mass matrix, L^T D L factor, actuation / passive forces, RNE bias, acc_smooth
//   k_collide    : geom AABBs, pair-ordered broad phase, narrow phase (plane / box / MPR) with multi-contact, sort
//   k_constraint : row assembly (equality | frictionloss | contacts | limits) + Newton/CG solver + contact forces
//   k_integrate  : link accelerations, semi-implicit Euler, FK + COM + motion subspaces + geom poses + velocities
// Reference citations use the abbreviations of oracle/rigid_oracle.hpp (FK, FD, CS, BP, NP, MPR, SF, CT, BOX, RS).
#pragma once
#include "gsr_types.h"
#include "gsr_math.cuh"

__constant__ DevModel c_m;
__constant__ Layout c_lay;

struct Env {
  float* F;
  int* I;
  size_t B;
};
#define FA(field, i) S.F[((size_t)(c_lay.field + (i))) * S.B]
#define IA(field, i) S.I[((size_t)(c_lay.field + (i))) * S.B]

__device__ __forceinline__ f3 ld3(const Env& S, int off, int i) {
  const float* p = S.F + ((size_t)(off + 3 * i)) * S.B;
  return f3(p[0], p[S.B], p[2 * S.B]);
}
__device__ __forceinline__ void st3(const Env& S, int off, int i, f3 v) {
  float* p = S.F + ((size_t)(off + 3 * i)) * S.B;
  p[0] = v.x; p[S.B] = v.y; p[2 * S.B] = v.z;
}
__device__ __forceinline__ f4 ld4(const Env& S, int off, int i) {
  const float* p = S.F + ((size_t)(off + 4 * i)) * S.B;
  return f4(p[0], p[S.B], p[2 * S.B], p[3 * S.B]);
}
__device__ __forceinline__ void st4(const Env& S, int off, int i, f4 v) {
  float* p = S.F + ((size_t)(off + 4 * i)) * S.B;
  p[0] = v.w; p[S.B] = v.x; p[2 * S.B] = v.y; p[3 * S.B] = v.z;
}
__device__ __forceinline__ sym3 ld6(const Env& S, int off, int i) {
  const float* p = S.F + ((size_t)(off + 6 * i)) * S.B;
  sym3 r; r.xx = p[0]; r.xy = p[S.B]; r.xz = p[2 * S.B]; r.yy = p[3 * S.B]; r.yz = p[4 * S.B]; r.zz = p[5 * S.B];
  return r;
}
__device__ __forceinline__ void st6(const Env& S, int off, int i, const sym3& v) {
  float* p = S.F + ((size_t)(off + 6 * i)) * S.B;
  p[0] = v.xx; p[S.B] = v.xy; p[2 * S.B] = v.xz; p[3 * S.B] = v.yy; p[4 * S.B] = v.yz; p[5 * S.B] = v.zz;
}
__device__ __forceinline__ f3 c3(const float* p) { return f3(p[0], p[1], p[2]); }
__device__ __forceinline__ f4 c4(const float* p) { return f4(p[0], p[1], p[2], p[3]); }

// =====================================================================================================
// Kinematics: FK:580-736 (link poses, joint anchors/axes, dofs_pos), FK:337-576 (COM, c-frame inertia, cdof),
// FK:828-864 (geom poses), FK:993-1116 (c-frame velocities, cdof_dot).
// =====================================================================================================
__device__ void dev_kinematics(const Env& S) {
  const float EPS = c_m.eps;
  for (int i_e = 0; i_e < c_m.nE; i_e++) {
    const int ls = c_m.e_link_start[i_e], le = c_m.e_link_end[i_e];
    // ---- pass 1: link poses; accumulate subtree mass * COM at the root link (kept in root_COM rows meanwhile)
    for (int l = ls; l < le; l++) {
      if (c_m.l_root[l] == l) st3(S, c_lay.root_COM, l, f3());
    }
    for (int l = ls; l < le; l++) {
      f3 pos = c3(c_m.l_pos[l]);
      f4 quat = c4(c_m.l_quat[l]);
      const int p = c_m.l_parent[l];
      if (p != -1) {
        f3 pp = ld3(S, c_lay.links_pos, p);
        f4 pq = ld4(S, c_lay.links_quat, p);
        pos = pp + rotate(c3(c_m.l_pos[l]), pq);
        quat = quat_compose(c4(c_m.l_quat[l]), pq);
      }
      for (int j = c_m.l_joint_start[l]; j < c_m.l_joint_end[l]; j++) {
        const int type = c_m.j_type[j], qs = c_m.j_q_start[j], ds = c_m.j_dof_start[j];
        const f3 jpos = c3(c_m.j_pos[j]);
        f3 xanchor, xaxis(0.f, 0.f, 1.f);
        if (type == JOINT_FREE) {
          pos = f3(FA(qpos, qs), FA(qpos, qs + 1), FA(qpos, qs + 2));
          f4 q(FA(qpos, qs + 3), FA(qpos, qs + 4), FA(qpos, qs + 5), FA(qpos, qs + 6));
          float n = 1.0f / sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
          quat = f4(q.w * n, q.x * n, q.y * n, q.z * n);
          xanchor = pos;
          f3 xyz = quat_to_xyz(quat, EPS);
          FA(dofs_pos, ds) = pos.x; FA(dofs_pos, ds + 1) = pos.y; FA(dofs_pos, ds + 2) = pos.z;
          FA(dofs_pos, ds + 3) = xyz.x; FA(dofs_pos, ds + 4) = xyz.y; FA(dofs_pos, ds + 5) = xyz.z;
        } else if (type == JOINT_REVOLUTE) {
          const f3 axis = c3(c_m.d_motion_ang[ds]);
          xanchor = rotate(jpos, quat) + pos;
          xaxis = rotate(axis, quat);
          const float dq = FA(qpos, qs) - c_m.qpos0[qs];
          FA(dofs_pos, ds) = dq;
          quat = quat_compose(rotvec_to_quat(axis * dq, EPS), quat);
          pos = xanchor - rotate(jpos, quat);
        } else if (type == JOINT_PRISMATIC) {
          const f3 axis = c3(c_m.d_motion_vel[ds]);
          xanchor = rotate(jpos, quat) + pos;
          xaxis = rotate(axis, quat);
          const float dq = FA(qpos, qs) - c_m.qpos0[qs];
          FA(dofs_pos, ds) = dq;
          pos = pos + xaxis * dq;
        } else if (type == JOINT_SPHERICAL) {
          xanchor = rotate(jpos, quat) + pos;
          f4 qloc(FA(qpos, qs), FA(qpos, qs + 1), FA(qpos, qs + 2), FA(qpos, qs + 3));
          f3 xyz = quat_to_xyz(qloc, EPS);
          FA(dofs_pos, ds) = xyz.x; FA(dofs_pos, ds + 1) = xyz.y; FA(dofs_pos, ds + 2) = xyz.z;
          quat = quat_compose(qloc, quat);
          pos = xanchor - rotate(jpos, quat);
        }
        st3(S, c_lay.joints_xanchor, j, xanchor);
        st3(S, c_lay.joints_xaxis, j, xaxis);
      }
      if (p == -1 && c_m.l_is_fixed[l]) {  // fixed root links keep the pose the user gave them
        pos = ld3(S, c_lay.links_pos, l);
        quat = ld4(S, c_lay.links_quat, l);
      } else {
        st3(S, c_lay.links_pos, l, pos);
        st4(S, c_lay.links_quat, l, quat);
      }
      const int r = c_m.l_root[l];
      const f3 ipw = pos + rotate(c3(c_m.l_ipos[l]), quat);
      f3 acc = ld3(S, c_lay.root_COM, r);
      st3(S, c_lay.root_COM, r, acc + ipw * c_m.l_mass[l]);
    }
    // ---- subtree COM per root (mass sums are model constants: no per-env mass shift in this build)
    for (int l = ls; l < le; l++) {
      if (c_m.l_root[l] != l) continue;
      float mass_sum = 0.f;
      for (int k = ls; k < le; k++) if (c_m.l_root[k] == l) mass_sum += c_m.l_mass[k];
      f3 com;
      if (mass_sum > EPS) com = ld3(S, c_lay.root_COM, l) / mass_sum;
      else com = ld3(S, c_lay.links_pos, l) + rotate(c3(c_m.l_ipos[l]), ld4(S, c_lay.links_quat, l));
      st3(S, c_lay.root_COM, l, com);
    }
    // ---- pass 2: c-frame inertia, motion subspaces, c-frame velocities
    for (int l = ls; l < le; l++) {
      const int r = c_m.l_root[l];
      const f3 com = ld3(S, c_lay.root_COM, r);
      if (r != l) st3(S, c_lay.root_COM, l, com);
      const f3 pos = ld3(S, c_lay.links_pos, l);
      const f4 quat = ld4(S, c_lay.links_quat, l);
      const f3 ipos = pos + rotate(c3(c_m.l_ipos[l]), quat) - com;
      const f4 iquat = quat_compose(c4(c_m.l_iquat[l]), quat);
      {  // geom.py:320-336 transform_inertia_by_trans_quat
        const float mass = c_m.l_mass[l];
        const mat3 R = quat_to_R(iquat, EPS);
        const float* Iin = c_m.l_inertia[l];
        float RI[3][3];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) RI[a][b] = R.m[a][0] * Iin[b] + R.m[a][1] * Iin[3 + b] + R.m[a][2] * Iin[6 + b];
        sym3 I;
        const float x = ipos.x, y = ipos.y, z = ipos.z;
        I.xx = RI[0][0] * R.m[0][0] + RI[0][1] * R.m[0][1] + RI[0][2] * R.m[0][2] + (y * y + z * z) * mass;
        I.xy = RI[0][0] * R.m[1][0] + RI[0][1] * R.m[1][1] + RI[0][2] * R.m[1][2] - x * y * mass;
        I.xz = RI[0][0] * R.m[2][0] + RI[0][1] * R.m[2][1] + RI[0][2] * R.m[2][2] - x * z * mass;
        I.yy = RI[1][0] * R.m[1][0] + RI[1][1] * R.m[1][1] + RI[1][2] * R.m[1][2] + (x * x + z * z) * mass;
        I.yz = RI[1][0] * R.m[2][0] + RI[1][1] * R.m[2][1] + RI[1][2] * R.m[2][2] - y * z * mass;
        I.zz = RI[2][0] * R.m[2][0] + RI[2][1] * R.m[2][1] + RI[2][2] * R.m[2][2] + (x * x + y * y) * mass;
        st6(S, c_lay.cinr_I, l, I);
        st3(S, c_lay.cinr_pos, l, ipos * mass);
      }
      // motion subspaces (FK:517-576) and c-frame velocity recursion (FK:993-1116)
      f3 cvel_vel, cvel_ang;
      const int p = c_m.l_parent[l];
      if (p != -1) { cvel_vel = ld3(S, c_lay.cd_vel, p); cvel_ang = ld3(S, c_lay.cd_ang, p); }
      for (int j = c_m.l_joint_start[l]; j < c_m.l_joint_end[l]; j++) {
        const int type = c_m.j_type[j], ds = c_m.j_dof_start[j];
        const f3 xaxis = ld3(S, c_lay.joints_xaxis, j);
        const f3 off = com - ld3(S, c_lay.joints_xanchor, j);
        if (type == JOINT_FREE) {
          const mat3 R = quat_to_R(quat, EPS);
          for (int k = 0; k < 3; k++) {
            f3 v; set_comp(v, k, 1.f);
            st3(S, c_lay.cdof_ang, ds + k, f3());
            st3(S, c_lay.cdof_vel, ds + k, v);
            st3(S, c_lay.cdofd_ang, ds + k, f3());
            st3(S, c_lay.cdofd_vel, ds + k, f3());
            cvel_vel += v * FA(vel, ds + k);
          }
          f3 a3[3], v3[3];
          for (int k = 0; k < 3; k++) {
            a3[k] = mat_col(R, k);
            v3[k] = cross(a3[k], off);
            st3(S, c_lay.cdof_ang, ds + 3 + k, a3[k]);
            st3(S, c_lay.cdof_vel, ds + 3 + k, v3[k]);
            // geom.py:379-383 motion_cross_motion(cvel, cdof)
            st3(S, c_lay.cdofd_ang, ds + 3 + k, cross(cvel_ang, a3[k]));
            st3(S, c_lay.cdofd_vel, ds + 3 + k, cross(cvel_ang, v3[k]) + cross(cvel_vel, a3[k]));
          }
          for (int k = 0; k < 3; k++) {
            const float w = FA(vel, ds + 3 + k);
            cvel_vel += v3[k] * w;
            cvel_ang += a3[k] * w;
          }
        } else if (type == JOINT_SPHERICAL) {
          const mat3 R = quat_to_R(quat, EPS);
          f3 a3[3], v3[3];
          for (int k = 0; k < 3; k++) {
            a3[k] = mat_col(R, k);
            v3[k] = cross(a3[k], off);
            st3(S, c_lay.cdof_ang, ds + k, a3[k]);
            st3(S, c_lay.cdof_vel, ds + k, v3[k]);
            st3(S, c_lay.cdofd_ang, ds + k, cross(cvel_ang, a3[k]));
            st3(S, c_lay.cdofd_vel, ds + k, cross(cvel_ang, v3[k]) + cross(cvel_vel, a3[k]));
          }
          for (int k = 0; k < 3; k++) {
            const float w = FA(vel, ds + k);
            cvel_vel += v3[k] * w;
            cvel_ang += a3[k] * w;
          }
        } else if (type == JOINT_REVOLUTE || type == JOINT_PRISMATIC) {
          f3 ca, cv;
          if (type == JOINT_REVOLUTE) { ca = xaxis; cv = cross(xaxis, off); }
          else { cv = xaxis; }
          st3(S, c_lay.cdof_ang, ds, ca);
          st3(S, c_lay.cdof_vel, ds, cv);
          st3(S, c_lay.cdofd_ang, ds, cross(cvel_ang, ca));
          st3(S, c_lay.cdofd_vel, ds, cross(cvel_ang, cv) + cross(cvel_vel, ca));
          const float w = FA(vel, ds);
          cvel_vel += cv * w;
          cvel_ang += ca * w;
        }
      }
      st3(S, c_lay.cd_vel, l, cvel_vel);
      st3(S, c_lay.cd_ang, l, cvel_ang);
    }
    // ---- geoms (FK:828-864): fixed geoms are only initialised once (gsr_init_state)
    for (int g = c_m.e_geom_start[i_e]; g < c_m.e_geom_end[i_e]; g++) {
      if (c_m.g_is_fixed[g]) continue;
      const int l = c_m.g_link[g];
      const f3 lp = ld3(S, c_lay.links_pos, l);
      const f4 lq = ld4(S, c_lay.links_quat, l);
      st3(S, c_lay.geoms_pos, g, lp + rotate(c3(c_m.g_pos[g]), lq));
      st4(S, c_lay.geoms_quat, g, quat_compose(c4(c_m.g_quat[g]), lq));
    }
  }
}

// =====================================================================================================
// Smooth dynamics (FD:146-227): CRBA, L^T D L, actuation, RNE bias, acc_smooth = M^-1 qf_smooth
// =====================================================================================================
__device__ __forceinline__ float& Mref(const Env& S, int off, int i, int j) { return S.F[((size_t)(off + i * c_m.nD + j)) * S.B]; }

// FD:545-602 (factor) operating on rows [ds, de) of the matrix stored at `off`
__device__ void dev_factor_block(const Env& S, int off, int ds, int de) {
  for (int i = de - 1; i >= ds; i--) {
    const float Dinv = 1.0f / Mref(S, off, i, i);
    FA(mass_Dinv, i) = Dinv;
    for (int j = i - 1; j >= ds; j--) {
      const float a = Mref(S, off, i, j) * Dinv;
      for (int k = ds; k <= j; k++) Mref(S, off, j, k) -= a * Mref(S, off, i, k);
      Mref(S, off, i, j) = a;
    }
    Mref(S, off, i, i) = 1.0f;
  }
}
// FD:813-894 (solve) on rows [ds, de): out = L^-1 D^-1 L^-T vec, vec/out are field offsets (may alias)
__device__ void dev_solve_block(const Env& S, int vec_off, int out_off, int ds, int de) {
  const int L = c_lay.mass_L;
  for (int i = de - 1; i >= ds; i--) {
    float c = S.F[((size_t)(vec_off + i)) * S.B];
    for (int j = i + 1; j < de; j++) c -= Mref(S, L, j, i) * S.F[((size_t)(out_off + j)) * S.B];
    S.F[((size_t)(out_off + i)) * S.B] = c;
  }
  for (int i = ds; i < de; i++) S.F[((size_t)(out_off + i)) * S.B] *= FA(mass_Dinv, i);
  for (int i = ds; i < de; i++) {
    float c = S.F[((size_t)(out_off + i)) * S.B];
    for (int j = ds; j < i; j++) c -= Mref(S, L, i, j) * S.F[((size_t)(out_off + j)) * S.B];
    S.F[((size_t)(out_off + i)) * S.B] = c;
  }
}

__device__ void dev_dynamics(const Env& S) {
  const int nD = c_m.nD, nL = c_m.nL;
  const float dt = c_m.dt;
  // ---- composite rigid body inertia, leaf -> root (FD:306-388), kept in the per-env scratch rows:
  // crb_I (6 nL) | crb_pos (3 nL) | crb_mass (nL) | cfrc_ang (3 nL) | cfrc_vel (3 nL)
  const int CI = c_lay.scratch, CP = c_lay.scratch + 6 * nL, CM = c_lay.scratch + 9 * nL;
  for (int l = 0; l < nL; l++) {
    st6(S, CI, l, ld6(S, c_lay.cinr_I, l));
    st3(S, CP, l, ld3(S, c_lay.cinr_pos, l));
    S.F[((size_t)(CM + l)) * S.B] = c_m.l_mass[l];
  }
  for (int i_e = 0; i_e < c_m.nE; i_e++)
    for (int l = c_m.e_link_end[i_e] - 1; l >= c_m.e_link_start[i_e]; l--) {
      const int p = c_m.l_parent[l];
      if (p == -1) continue;
      sym3 a = ld6(S, CI, p), b = ld6(S, CI, l);
      a.xx += b.xx; a.xy += b.xy; a.xz += b.xz; a.yy += b.yy; a.yz += b.yz; a.zz += b.zz;
      st6(S, CI, p, a);
      st3(S, CP, p, ld3(S, CP, p) + ld3(S, CP, l));
      S.F[((size_t)(CM + p)) * S.B] += S.F[((size_t)(CM + l)) * S.B];
    }
  // ---- mass matrix (FD:390-541): M[i][j] = f_i . cdof_j for j ancestor-or-self dof of i (lower triangle), mirrored
  for (int i = 0; i < nD; i++) {
    const int l = c_m.d_link[i];
    f3 f_ang, f_vel;
    inertial_mul(ld3(S, CP, l), ld6(S, CI, l), S.F[((size_t)(CM + l)) * S.B], ld3(S, c_lay.cdof_vel, i), ld3(S, c_lay.cdof_ang, i), f_ang, f_vel);
    const int es = c_m.e_dof_start[c_m.d_entity[i]];
    for (int j = es; j <= i; j++) {
      float v = 0.f;
      if (c_m.mass_mask[i][j]) v = dot(f_ang, ld3(S, c_lay.cdof_ang, j)) + dot(f_vel, ld3(S, c_lay.cdof_vel, j));
      if (i == j) {
        v += c_m.d_armature[i];
        if (c_m.integrator == INTEGRATOR_APPROX_IMPLICITFAST) {
          v += c_m.d_damping[i] * dt;
          if (IA(ctrl_mode, i) <= CTRL_VELOCITY) v -= c_m.d_act_bias[i][2] * dt;
        }
      }
      Mref(S, c_lay.mass_mat, i, j) = v;
      Mref(S, c_lay.mass_mat, j, i) = v;
      Mref(S, c_lay.mass_L, i, j) = v;
    }
  }
  // ---- factor (FD:545-602)
  for (int i_e = 0; i_e < c_m.nE; i_e++) dev_factor_block(S, c_lay.mass_L, c_m.e_dof_start[i_e], c_m.e_dof_end[i_e]);

  // ---- actuation + passive forces (FD:956-1182)
  for (int l = 0; l < nL; l++) {
    const int ds = c_m.l_dof_start[l], de = c_m.l_dof_end[l];
    if (de - ds <= 0) continue;
    const int jt = c_m.j_type[c_m.l_joint_start[l]];
    for (int d = ds; d < de; d++) {
      const int mode = IA(ctrl_mode, d);
      const float v = FA(vel, d), dp = FA(dofs_pos, d);
      float force = 0.f;
      if (mode == CTRL_FORCE) force = FA(ctrl_force, d);
      else if (mode == CTRL_VELOCITY) force = -c_m.d_act_bias[d][2] * (FA(ctrl_vel, d) - v);
      else if (mode == CTRL_POSITION && !(jt == JOINT_FREE && d >= ds + 3))
        force = c_m.d_kp[d] * (FA(ctrl_pos, d) - dp) + c_m.d_act_bias[d][0] + (c_m.d_kp[d] + c_m.d_act_bias[d][1]) * dp +
                c_m.d_act_bias[d][2] * (v - FA(ctrl_vel, d));
      FA(qf_applied, d) = fminf(fmaxf(force, c_m.d_force_range[d][0]), c_m.d_force_range[d][1]);
      float passive = -c_m.d_damping[d] * v;
      if (jt != JOINT_FREE && jt != JOINT_FIXED) passive += -dp * c_m.d_stiffness[d];
      FA(qf_passive, d) = passive;
    }
    if (jt == JOINT_FREE && (IA(ctrl_mode, ds + 3) == CTRL_POSITION || IA(ctrl_mode, ds + 4) == CTRL_POSITION || IA(ctrl_mode, ds + 5) == CTRL_POSITION)) {
      f3 xyz(FA(dofs_pos, ds + 3), FA(dofs_pos, ds + 4), FA(dofs_pos, ds + 5));
      f3 cxyz(FA(ctrl_pos, ds + 3), FA(ctrl_pos, ds + 4), FA(ctrl_pos, ds + 5));
      f3 rv = quat_to_rotvec(quat_compose(xyz_to_quat(cxyz), quat_inv(xyz_to_quat(xyz))), c_m.eps);
      for (int k = 0; k < 3; k++) {
        const int d = ds + 3 + k;
        float force = c_m.d_kp[d] * comp(rv, k) + c_m.d_act_bias[d][0] + (c_m.d_kp[d] + c_m.d_act_bias[d][1]) * FA(dofs_pos, d) +
                      c_m.d_act_bias[d][2] * (FA(vel, d) - FA(ctrl_vel, d));
        FA(qf_applied, d) = fminf(fmaxf(force, c_m.d_force_range[d][0]), c_m.d_force_range[d][1]);
      }
    }
  }

  // ---- RNE bias forces (FD:1186-1486).  cdd (link c-frame bias acceleration) lives in the cacc rows until
  // k_integrate overwrites them with the true link accelerations; cfrc is accumulated in the scratch rows.
  const int FAo = c_lay.scratch + 10 * nL, FVo = c_lay.scratch + 13 * nL;
  const f3 grav = c3(c_m.gravity);
  for (int l = 0; l < nL; l++) {
    const int p = c_m.l_parent[l];
    f3 cdd_vel, cdd_ang;
    if (p == -1) cdd_vel = -grav * (1.0f - c_m.e_gravcomp[c_m.l_entity[l]]);
    else { cdd_vel = ld3(S, c_lay.cacc_lin, p); cdd_ang = ld3(S, c_lay.cacc_ang, p); }
    for (int d = c_m.l_dof_start[l]; d < c_m.l_dof_end[l]; d++) {
      const float v = FA(vel, d);
      cdd_vel += ld3(S, c_lay.cdofd_vel, d) * v;
      cdd_ang += ld3(S, c_lay.cdofd_ang, d) * v;
    }
    st3(S, c_lay.cacc_lin, l, cdd_vel);
    st3(S, c_lay.cacc_ang, l, cdd_ang);
    const sym3 I = ld6(S, c_lay.cinr_I, l);
    const f3 ip = ld3(S, c_lay.cinr_pos, l);
    const float mass = c_m.l_mass[l];
    const f3 cv = ld3(S, c_lay.cd_vel, l), ca = ld3(S, c_lay.cd_ang, l);
    f3 f1a, f1v, f2a, f2v;
    inertial_mul(ip, I, mass, cdd_vel, cdd_ang, f1a, f1v);
    inertial_mul(ip, I, mass, cv, ca, f2a, f2v);
    // geom.py:372-376 motion_cross_force
    const f3 f3v = cross(ca, f2v);
    const f3 f3a = cross(ca, f2a) + cross(cv, f2v);
    st3(S, FVo, l, f1v + f3v + ld3(S, c_lay.cfrc_applied_vel, l));
    st3(S, FAo, l, f1a + f3a + ld3(S, c_lay.cfrc_applied_ang, l));
  }
  for (int i_e = 0; i_e < c_m.nE; i_e++)
    for (int l = c_m.e_link_end[i_e] - 1; l >= c_m.e_link_start[i_e]; l--) {
      const int p = c_m.l_parent[l];
      if (p == -1) continue;
      st3(S, FVo, p, ld3(S, FVo, p) + ld3(S, FVo, l));
      st3(S, FAo, p, ld3(S, FAo, p) + ld3(S, FAo, l));
    }
  for (int d = 0; d < nD; d++) {
    const int l = c_m.d_link[d];
    const float bias = dot(ld3(S, c_lay.cdof_ang, d), ld3(S, FAo, l)) + dot(ld3(S, c_lay.cdof_vel, d), ld3(S, FVo, l));
    FA(qf_bias, d) = bias;
    const float f = FA(qf_passive, d) - bias + FA(qf_applied, d);
    FA(force, d) = f;
    FA(qf_smooth, d) = f;
  }
  // ---- acc_smooth (FD:1506-1563)
  for (int i_e = 0; i_e < c_m.nE; i_e++) dev_solve_block(S, c_lay.force, c_lay.acc_smooth, c_m.e_dof_start[i_e], c_m.e_dof_end[i_e]);
  for (int d = 0; d < nD; d++) FA(acc, d) = FA(acc_smooth, d);
}

// =====================================================================================================
// Collision detection (collider.py:684-808): AABBs, broad phase, narrow phase, canonical contact order
// =====================================================================================================
// FK:1297-1319: world AABB of the rest-pose box.  The reference transforms the 8 corners and takes min/max; the
// identical box is obtained from centre +- |R| * half-extent with a third of the arithmetic.
__device__ __forceinline__ void dev_geom_aabb(const Env& S, int g) {
  const f3 gp = ld3(S, c_lay.geoms_pos, g);
  const mat3 R = quat_to_R(ld4(S, c_lay.geoms_quat, g), 0.f);
  const f3 lo = c3(c_m.g_aabb_lo[g]), hi = c3(c_m.g_aabb_hi[g]);
  const f3 c = (lo + hi) * 0.5f, h = (hi - lo) * 0.5f;
  const f3 wc(R.m[0][0] * c.x + R.m[0][1] * c.y + R.m[0][2] * c.z + gp.x, R.m[1][0] * c.x + R.m[1][1] * c.y + R.m[1][2] * c.z + gp.y,
              R.m[2][0] * c.x + R.m[2][1] * c.y + R.m[2][2] * c.z + gp.z);
  const f3 wh(fabsf(R.m[0][0]) * h.x + fabsf(R.m[0][1]) * h.y + fabsf(R.m[0][2]) * h.z,
              fabsf(R.m[1][0]) * h.x + fabsf(R.m[1][1]) * h.y + fabsf(R.m[1][2]) * h.z,
              fabsf(R.m[2][0]) * h.x + fabsf(R.m[2][1]) * h.y + fabsf(R.m[2][2]) * h.z);
  st3(S, c_lay.aabb_min, g, wc - wh);
  st3(S, c_lay.aabb_max, g, wc + wh);
}

// SF:138-180 support point of a mesh-backed geom from its 180x180 direction table (L2-resident, read-only path)
__device__ f3 dev_support_mesh(f3 d, int g) {
  const float PI = 3.14159265358979323846f;
  const float theta = atan2f(d.y, d.x), phi = acosf(d.z);
  const int res = 180;
  const float ii = (theta + PI) / PI / 2.f * res, jj = phi / PI * res;
  float dot_max = -1e20f;
  f3 best;
  const float* tab = c_m.support_v + 3 * (size_t)c_m.g_support_start[g];
#pragma unroll
  for (int i4 = 0; i4 < 4; i4++) {
    int i = (i4 & 1) ? ((int)ceilf(ii)) % res : ((int)floorf(ii)) % res;
    int j;
    if (i4 >> 1) { j = (int)fminf(fmaxf(ceilf(jj), 0.f), (float)(res - 1)); if (j == res - 1) j = res - 2; }
    else { j = (int)fminf(fmaxf(floorf(jj), 0.f), (float)(res - 1)); if (j == 0) j = 1; }
    const float* p = tab + 3 * (i * res + j);
    const f3 v(__ldg(p), __ldg(p + 1), __ldg(p + 2));
    const float dt = dot(v, d);
    if (dt > dot_max) { best = v; dot_max = dt; }
  }
  return best;
}

// MPR:146-176 support_driver
__device__ f3 dev_support(f3 d, int g, f3 pos, f4 quat) {
  const int type = c_m.g_type[g];
  if (type == GEOM_SPHERE) return pos + d * c_m.g_data[g][0];
  if (type == GEOM_BOX) {
    const f3 db = rotate_inv(d, quat);
    const f3 v((db.x < 0.f ? -1.f : 1.f) * c_m.g_data[g][0] * 0.5f, (db.y < 0.f ? -1.f : 1.f) * c_m.g_data[g][1] * 0.5f,
               (db.z < 0.f ? -1.f : 1.f) * c_m.g_data[g][2] * 0.5f);
    return transform(v, pos, quat);
  }
  if (type == GEOM_CAPSULE) {
    const float r = c_m.g_data[g][0], hl = 0.5f * c_m.g_data[g][1];
    const f3 axis = rotate(f3(0.f, 0.f, 1.f), quat);
    const float side = dot(d, axis) < 0.f ? -1.f : 1.f;
    return pos + axis * (hl * side) + d * r;
  }
  if (type == GEOM_ELLIPSOID) {
    const float a = c_m.g_data[g][0], b = c_m.g_data[g][1], c = c_m.g_data[g][2];
    const f3 dl = rotate_inv(d, quat);
    const float n = sqrtf(a * a * dl.x * dl.x + b * b * dl.y * dl.y + c * c * dl.z * dl.z);
    return pos + rotate(f3(a * a * dl.x / n, b * b * dl.y / n, c * c * dl.z / n), quat);
  }
  return transform(dev_support_mesh(rotate_inv(d, quat), g), pos, quat);
}

// ---- MPR (mpr.py) ---------------------------------------------------------------------------------
struct Simplex {
  f3 v[4], v1[4], v2[4];
};
__device__ __forceinline__ void simplex_set(Simplex& s, int i, f3 v, f3 v1, f3 v2) {
  // explicit selects keep the simplex in registers (no dynamically indexed local memory)
  if (i == 1) { s.v[1] = v; s.v1[1] = v1; s.v2[1] = v2; }
  else if (i == 2) { s.v[2] = v; s.v1[2] = v1; s.v2[2] = v2; }
  else { s.v[3] = v; s.v1[3] = v1; s.v2[3] = v2; }
}
struct PairPose { int ga, gb; f3 pa, pb; f4 qa, qb; };
__device__ __forceinline__ void mpr_support(const PairPose& P, f3 d, f3& v, f3& v1, f3& v2) {
  v1 = dev_support(d, P.ga, P.pa, P.qa);
  v2 = dev_support(-d, P.gb, P.pb, P.qb);
  v = v1 - v2;
}
__device__ __forceinline__ f3 portal_dir(const Simplex& s) { return normalized(cross(s.v[2] - s.v[1], s.v[3] - s.v[1])); }
__device__ __forceinline__ bool reach_tolerance(const Simplex& s, f3 v, f3 dir) {  // MPR:134-143
  const float dv4 = dot(v, dir);
  const float dot1 = fminf(fminf(dv4 - dot(s.v[1], dir), dv4 - dot(s.v[2], dir)), dv4 - dot(s.v[3], dir));
  return dot1 < c_m.ccd_tolerance + c_m.ccd_eps * fmaxf(1.0f, dot1);
}
__device__ __forceinline__ void expand_portal(Simplex& s, f3 v, f3 v1, f3 v2) {  // MPR:426-442
  const f3 v4v0 = cross(v, s.v[0]);
  int is;
  if (dot(s.v[1], v4v0) > 0.f) is = dot(s.v[2], v4v0) > 0.f ? 1 : 3;
  else is = dot(s.v[3], v4v0) > 0.f ? 2 : 1;
  simplex_set(s, is, v, v1, v2);
}
// MPR:445-598
__device__ int discover_portal(Simplex& s, const PairPose& P, f3 ca, f3 cb) {
  const float CE = c_m.ccd_eps;
  s.v1[0] = ca; s.v2[0] = cb; s.v[0] = ca - cb;
  if (fabsf(s.v[0].x) < CE && fabsf(s.v[0].y) < CE && fabsf(s.v[0].z) < CE) s.v[0].x += 10.0f * CE;
  f3 dir = -normalized(s.v[0]);
  f3 v, v1, v2;
  mpr_support(P, dir, v, v1, v2);
  s.v[1] = v; s.v1[1] = v1; s.v2[1] = v2;
  if (dot(v, dir) < CE) return -1;
  dir = cross(s.v[0], s.v[1]);
  if (dot(dir, dir) < CE) {
    if (fabsf(s.v[1].x) < CE && fabsf(s.v[1].y) < CE && fabsf(s.v[1].z) < CE) return 1;
    return 2;
  }
  dir = normalized(dir);
  mpr_support(P, dir, v, v1, v2);
  if (dot(v, dir) < CE) return -1;
  s.v[2] = v; s.v1[2] = v1; s.v2[2] = v2;
  dir = normalized(cross(s.v[1] - s.v[0], s.v[2] - s.v[0]));
  if (dot(dir, s.v[0]) > 0.f) {
    f3 t;
    t = s.v[1]; s.v[1] = s.v[2]; s.v[2] = t;
    t = s.v1[1]; s.v1[1] = s.v1[2]; s.v1[2] = t;
    t = s.v2[1]; s.v2[1] = s.v2[2]; s.v2[2] = t;
    dir = -dir;
  }
  int num_trials = 0;
  while (true) {
    mpr_support(P, dir, v, v1, v2);
    if (dot(v, dir) < CE) return -1;
    bool cont = false;
    if (dot(cross(s.v[1], v), s.v[0]) < -CE) { s.v[2] = v; s.v1[2] = v1; s.v2[2] = v2; cont = true; }
    if (!cont && dot(cross(v, s.v[2]), s.v[0]) < -CE) { s.v[1] = v; s.v1[1] = v1; s.v2[1] = v2; cont = true; }
    if (cont) {
      dir = normalized(cross(s.v[1] - s.v[0], s.v[2] - s.v[0]));
      if (++num_trials == 15) return -1;
    } else {
      s.v[3] = v; s.v1[3] = v1; s.v2[3] = v2;
      return 0;
    }
  }
}
// MPR:601-819 func_mpr_contact (centre guess + portal discovery/refinement + penetration)
__device__ void dev_mpr_contact(const PairPose& P, f3 normal_ws, bool& is_col, f3& normal, float& pen, f3& pos) {
  const float CE = c_m.ccd_eps;
  f3 ca = transform(c3(c_m.g_center[P.ga]), P.pa, P.qa), cb = transform(c3(c_m.g_center[P.gb]), P.pb, P.qb);
  if (fabsf(normal_ws.x) > CE || fabsf(normal_ws.y) > CE || fabsf(normal_ws.z) > CE) {
    const f3 loa = c3(c_m.g_aabb_lo[P.ga]), hia = c3(c_m.g_aabb_hi[P.ga]), lob = c3(c_m.g_aabb_lo[P.gb]), hib = c3(c_m.g_aabb_hi[P.gb]);
    ca = transform((hia + loa) * 0.5f, P.pa, P.qa);
    cb = transform((hib + lob) * 0.5f, P.pb, P.qb);
    const f3 delta = ca - cb;
    if (norm(cross(normal_ws, normalized(delta))) > 0.01f) {
      const f3 offset = normal_ws * dot(delta, normal_ws) - delta;
      const float on = norm(offset);
      if (on > c_m.eps) {
        const f3 dir = offset / on;
        const f3 dla = rotate_inv(dir, P.qa), dlb = rotate_inv(dir, P.qb);
        const f3 bsa = hia - loa, bsb = hib - lob;
        const float la = bsa.x * fabsf(dla.x) + bsa.y * fabsf(dla.y) + bsa.z * fabsf(dla.z);
        const float lb = bsb.x * fabsf(dlb.x) + bsb.y * fabsf(dlb.y) + bsb.z * fabsf(dlb.z);
        const float ratio = fminf(on / (la + lb), 0.5f);
        ca = ca + dir * (la * ratio);
        cb = cb - dir * (lb * ratio);
      }
    }
  }
  Simplex s;
  int res = discover_portal(s, P, ca, cb);
  is_col = false; pos = f3(); normal = f3(); pen = 0.f;
  if (res == 1) { is_col = true; normal = -normalized(s.v[0]); pos = (s.v1[1] + s.v2[1]) * 0.5f; }
  else if (res == 2) { is_col = true; pen = norm(s.v[1]); normal = -normalized(s.v[1]); pos = (s.v1[1] + s.v2[1]) * 0.5f; }
  else if (res == 0) {
    // MPR:232-278 refine portal
    while (true) {
      const f3 dir = portal_dir(s);
      if (dot(s.v[1], dir) > -CE) { res = 0; break; }
      f3 v, v1, v2;
      mpr_support(P, dir, v, v1, v2);
      if (!(dot(v, dir) > -CE) || reach_tolerance(s, v, dir)) { res = -1; break; }
      expand_portal(s, v, v1, v2);
    }
    if (res >= 0) {
      // MPR:338-423 find penetration
      int it = 0;
      while (true) {
        const f3 dir = portal_dir(s);
        f3 v, v1, v2;
        mpr_support(P, dir, v, v1, v2);
        if (reach_tolerance(s, v, dir) || it > c_m.ccd_iterations) {
          pen = dot(dir, s.v[1]);
          normal = -dir;
          is_col = true;
          // MPR:281-316 find_pos: barycentric weights of the origin's projection on the portal
          const float b1 = dot(cross(s.v[2], s.v[3]), dir), b2 = dot(cross(s.v[3], s.v[1]), dir), b3 = dot(cross(s.v[1], s.v[2]), dir);
          const float sum = b1 + b2 + b3;
          const f3 p1 = s.v1[1] * b1 + s.v1[2] * b2 + s.v1[3] * b3, p2 = s.v2[1] * b1 + s.v2[2] * b2 + s.v2[3] * b3;
          pos = (p1 + p2) * (0.5f / sum);
          break;
        }
        expand_portal(s, v, v1, v2);
        it++;
      }
    }
  }
}

// CT:181-222 func_add_contact; `nc` is the running per-env contact count (register)
__device__ __forceinline__ void dev_add_contact(const Env& S, int& nc, int ga, int gb, f3 normal, f3 pos, float pen) {
  if (nc < c_m.max_contact_pairs) {
    IA(contact_geom_a, nc) = ga; IA(contact_geom_b, nc) = gb;
    IA(contact_link_a, nc) = c_m.g_link[ga]; IA(contact_link_b, nc) = c_m.g_link[gb];
    st3(S, c_lay.contact_normal, nc, normal);
    st3(S, c_lay.contact_pos, nc, pos);
    FA(contact_penetration, nc) = pen;
    const float fa = c_m.g_friction[ga] * FA(friction_ratio, ga), fb = c_m.g_friction[gb] * FA(friction_ratio, gb);
    FA(contact_friction, nc) = fmaxf(fmaxf(fa, fb), 1e-2f);
    nc++;
  } else {
    IA(errno_, 0) |= ERR_CONTACT_OVERFLOW;
  }
}
// CT:317-336
__device__ __forceinline__ float dev_tolerance(int ga, int gb) {
  float size = norm(c3(c_m.g_aabb_hi[gb]) - c3(c_m.g_aabb_lo[gb]));
  if (c_m.g_type[ga] != GEOM_PLANE) size = fminf(norm(c3(c_m.g_aabb_hi[ga]) - c3(c_m.g_aabb_lo[ga])), size);
  return 0.5f * c_m.mc_tolerance * size;
}
// CT:415-436
__device__ __forceinline__ void rotate_frame(f3 pos, f4 quat, f3 cpos, f4 qrot, f3& npos, f4& nquat) {
  nquat = quat_compose(quat, qrot);
  const f3 rel = cpos - pos;
  npos = pos - (rotate(rel, qrot) - rel);
}

// NP:519-970 func_convex_convex_contact for one broad-phase pair (type[ga] <= type[gb], not plane-box)
__device__ void dev_convex_convex(const Env& S, int& nc, int ga, int gb, int i_pair) {
  const float EPS = c_m.eps;
  const int ta = c_m.g_type[ga], tb = c_m.g_type[gb];
  const bool multi = c_m.enable_multi_contact && ta != GEOM_SPHERE && ta != GEOM_ELLIPSOID && tb != GEOM_SPHERE && tb != GEOM_ELLIPSOID;
  const float tolerance = dev_tolerance(ga, gb);
  const f3 ga_pos0 = ld3(S, c_lay.geoms_pos, ga), gb_pos0 = ld3(S, c_lay.geoms_pos, gb);
  const f4 ga_quat0 = ld4(S, c_lay.geoms_quat, ga), gb_quat0 = ld4(S, c_lay.geoms_quat, gb);
  PairPose P;
  P.ga = ga; P.gb = gb; P.pa = ga_pos0; P.pb = gb_pos0; P.qa = ga_quat0; P.qb = gb_quat0;
  bool is_col_0 = false, is_col = false;
  float pen = 0.f;
  f3 normal_0, cpos_0, normal, cpos, axis_0, axis_1;
  f4 qrot(0.f, 0.f, 0.f, 0.f);
  int n_con = 0;
  const bool use_cache = ta != GEOM_PLANE;
  for (int i_det = 0; i_det < 5; i_det++) {
    if (i_det > 0 && !(multi && is_col_0)) break;
    if (i_det > 0) {
      const f3 axis = axis_0 * (float)(2 * (i_det % 2) - 1) + axis_1 * (float)(1 - 2 * ((i_det / 2) % 2));
      qrot = rotvec_to_quat(axis * c_m.mc_perturbation, EPS);
      rotate_frame(ga_pos0, ga_quat0, cpos_0, qrot, P.pa, P.qa);
      rotate_frame(gb_pos0, gb_quat0, cpos_0, quat_inv(qrot), P.pb, P.qb);
    }
    if (ta == GEOM_PLANE) {
      normal = -normalized(rotate(c3(c_m.g_data[ga]), P.qa));
      const f3 v1 = dev_support(normal, gb, P.pb, P.qb);
      pen = dot(normal, v1 - P.pa);
      cpos = v1 - normal * (0.5f * pen);
      is_col = pen > 0.f;
    } else {
      f3 normal_ws;
      const bool valid = (IA(cache_valid, i_pair >> 5) >> (i_pair & 31)) & 1;
      if (valid) normal_ws = ld3(S, c_lay.cache_normal, i_pair);
      bool guess = fabsf(normal_ws.x) > EPS || fabsf(normal_ws.y) > EPS || fabsf(normal_ws.z) > EPS;
      dev_mpr_contact(P, normal_ws, is_col, normal, pen, cpos);
      if (i_det == 0 && !is_col && guess) dev_mpr_contact(P, f3(), is_col, normal, pen, cpos);
    }
    if (i_det == 0) {
      is_col_0 = is_col; normal_0 = normal; cpos_0 = cpos;
      if (is_col_0) {
        dev_add_contact(S, nc, ga, gb, normal, cpos, pen);
        if (multi) {
          // CT:353-412 func_contact_orthogonals: perturbation axes from the principal inertia frame of the smaller geom
          int g = gb;
          if (ta != GEOM_PLANE) {
            const f3 sa = c3(c_m.g_aabb_hi[ga]), sb = c3(c_m.g_aabb_hi[gb]);
            g = (sa.x * sa.y * sa.z < sb.x * sb.y * sb.z) ? ga : gb;
          }
          const int l = c_m.g_link[g];
          const mat3 rot = quat_to_R(quat_compose(c4(c_m.l_iquat[l]), ld4(S, c_lay.links_quat, l)), EPS);
          int axis_idx = 0;
          float amax = 0.f;
          for (int i = 0; i < 3; i++) { const float a = fabsf(dot(mat_col(rot, i), normal)); if (a > amax) { amax = a; axis_idx = i; } }
          axis_idx = (axis_idx + 1) % 3;
          axis_0 = mat_col(rot, axis_idx);
          axis_0 = normalized(axis_0 - normal * dot(normal, axis_0));
          axis_1 = cross(normal, axis_0);
          n_con = 1;
        }
      }
      if (use_cache) {
        int w = IA(cache_valid, i_pair >> 5);
        if (is_col_0) { st3(S, c_lay.cache_normal, i_pair, normal); w |= (1 << (i_pair & 31)); }
        else w &= ~(1 << (i_pair & 31));
        IA(cache_valid, i_pair >> 5) = w;
      }
    } else if (is_col) {
      const f3 cpa = rotate((cpos - normal * (0.5f * pen)) - cpos_0, quat_inv(qrot)) + cpos_0;
      const f3 cpb = rotate((cpos + normal * (0.5f * pen)) - cpos_0, qrot) + cpos_0;
      cpos = (cpa + cpb) * 0.5f;
      f3 twist = cross(normal, normal_0);
      const float mcp = c_m.mc_perturbation;
      twist = f3(fminf(fmaxf(twist.x, -mcp), mcp), fminf(fmaxf(twist.y, -mcp), mcp), fminf(fmaxf(twist.z, -mcp), mcp));
      normal = normal + cross(twist, normal);
      pen = dot(normal, cpb - cpa);
      bool repeated = false;
      for (int c = 0; c < n_con; c++) {
        const f3 prev = ld3(S, c_lay.contact_pos, nc - 1 - c);
        if (norm(cpos - prev) < tolerance) { repeated = true; break; }
      }
      if (!repeated && pen > -tolerance) {
        dev_add_contact(S, nc, ga, gb, normal, cpos, fmaxf(pen, 0.f));
        n_con++;
      }
    }
  }
}

// BOX:24-96 func_plane_box_contact
__device__ void dev_plane_box(const Env& S, int& nc, int ga, int gb) {
  const f3 pa = ld3(S, c_lay.geoms_pos, ga), pb = ld3(S, c_lay.geoms_pos, gb);
  const f4 qa = ld4(S, c_lay.geoms_quat, ga), qb = ld4(S, c_lay.geoms_quat, gb);
  const f3 normal = -normalized(rotate(c3(c_m.g_data[ga]), qa));
  const f3 v1 = dev_support(normal, gb, pb, qb);
  float pen = dot(normal, v1 - pa);
  if (pen > 0.f) {
    f3 cpos = v1 - normal * (0.5f * pen);
    dev_add_contact(S, nc, ga, gb, normal, cpos, pen);
    if (c_m.enable_multi_contact) {
      int n_con = 1;
      const f3 cpos_0 = cpos;
      const float tolerance = dev_tolerance(ga, gb);
      for (int iv = c_m.g_vert_start[gb]; iv < c_m.g_vert_end[gb]; iv++) {
        if (n_con < c_m.n_contacts_per_pair) {
          const float* vp = c_m.verts + 3 * (size_t)iv;
          const f3 pc = transform(f3(__ldg(vp), __ldg(vp + 1), __ldg(vp + 2)), pb, qb);
          pen = dot(normal, pc - pa);
          if (pen > 0.f) {
            cpos = pc - normal * (0.5f * pen);
            if (norm(cpos - cpos_0) > tolerance) { dev_add_contact(S, nc, ga, gb, normal, cpos, pen); n_con++; }
          }
        }
      }
    }
  }
}

__device__ void dev_collide(const Env& S) {
  // ---- AABBs of moving geoms (fixed ones were set at init)
  for (int g = 0; g < c_m.nG; g++)
    if (!c_m.g_is_fixed[g]) dev_geom_aabb(S, g);
  if (c_m.nP == 0) return;
  // ---- BP:74-139 clear the previous contacts (only the slots that were in use)
  const int n_prev = IA(n_contacts, 0);
  for (int c = 0; c < n_prev; c++) {
    IA(contact_geom_a, c) = -1; IA(contact_geom_b, c) = -1; IA(contact_link_a, c) = -1; IA(contact_link_b, c) = -1;
    FA(contact_penetration, c) = 0.f;
    st3(S, c_lay.contact_pos, c, f3()); st3(S, c_lay.contact_normal, c, f3()); st3(S, c_lay.contact_force, c, f3());
  }
  int nc = 0, n_broad = 0;
  // ---- BP:400-460 + NP:1989-2060: pair-ordered traversal; pass 0 = convex-convex, pass 1 = plane-box (NP:2135-2173)
  for (int pass = 0; pass < 2; pass++) {
    int cur_a = -1;
    f3 amin, amax;
    for (int p = 0; p < c_m.nP; p++) {
      const int ga0 = __ldg(c_m.pairs + 2 * p), gb0 = __ldg(c_m.pairs + 2 * p + 1);
      const bool plane_box = (c_m.g_type[ga0] == GEOM_PLANE && c_m.g_type[gb0] == GEOM_BOX) || (c_m.g_type[gb0] == GEOM_PLANE && c_m.g_type[ga0] == GEOM_BOX);
      if (pass == 1 && !plane_box) continue;
      if (ga0 != cur_a) { cur_a = ga0; amin = ld3(S, c_lay.aabb_min, ga0); amax = ld3(S, c_lay.aabb_max, ga0); }
      const f3 bmin = ld3(S, c_lay.aabb_min, gb0), bmax = ld3(S, c_lay.aabb_max, gb0);
      const bool sep = amax.x <= bmin.x || amin.x >= bmax.x || amax.y <= bmin.y || amin.y >= bmax.y || amax.z <= bmin.z || amin.z >= bmax.z;
      if (sep) {
        if (pass == 0 && c_m.g_type[ga0] != GEOM_PLANE && c_m.g_type[gb0] != GEOM_PLANE) {
          const int w = IA(cache_valid, p >> 5);
          if ((w >> (p & 31)) & 1) IA(cache_valid, p >> 5) = w & ~(1 << (p & 31));
        }
        continue;
      }
      if (pass == 0) {
        if (n_broad >= c_m.max_broad_pairs) { IA(errno_, 0) |= ERR_BROAD_OVERFLOW; continue; }
        n_broad++;
      }
      int ga = ga0, gb = gb0;
      if (c_m.g_type[ga] > c_m.g_type[gb]) { ga = gb0; gb = ga0; }
      if (pass == 0) { if (!plane_box) dev_convex_convex(S, nc, ga, gb, p); }
      else dev_plane_box(S, nc, ga, gb);
    }
  }
  IA(n_contacts, 0) = nc;
  // ---- CT:439-538 canonical order: stable insertion sort of pair groups by x of the group's first contact
  if (c_m.has_cc && c_m.sort_contacts && nc > 1) {
    float gk = 0.f;
    for (int i = 0; i < nc; i++) {
      if (i == 0 || IA(contact_geom_a, i) != IA(contact_geom_a, i - 1) || IA(contact_geom_b, i) != IA(contact_geom_b, i - 1)) gk = FA(contact_pos, 3 * i);
      FA(sort_key, i) = gk;
      IA(sort_idx, i) = i;
    }
    bool moved = false;
    for (int i = 1; i < nc; i++) {
      const float ck = FA(sort_key, i);
      if (FA(sort_key, i - 1) <= ck) continue;
      const int ci = IA(sort_idx, i);
      int j = i - 1;
      while (j >= 0) {
        const float kj = FA(sort_key, j);
        if (kj <= ck) break;
        FA(sort_key, j + 1) = kj; IA(sort_idx, j + 1) = IA(sort_idx, j);
        j--;
      }
      FA(sort_key, j + 1) = ck; IA(sort_idx, j + 1) = ci;
      moved = true;
    }
    if (moved) {
      for (int i = 0; i < nc; i++) {
        if (IA(sort_idx, i) == i) continue;
        const int t_ga = IA(contact_geom_a, i), t_gb = IA(contact_geom_b, i), t_la = IA(contact_link_a, i), t_lb = IA(contact_link_b, i);
        const f3 t_n = ld3(S, c_lay.contact_normal, i), t_p = ld3(S, c_lay.contact_pos, i);
        const float t_pen = FA(contact_penetration, i), t_fr = FA(contact_friction, i);
        int j = i;
        while (IA(sort_idx, j) != i) {
          const int src = IA(sort_idx, j);
          IA(contact_geom_a, j) = IA(contact_geom_a, src); IA(contact_geom_b, j) = IA(contact_geom_b, src);
          IA(contact_link_a, j) = IA(contact_link_a, src); IA(contact_link_b, j) = IA(contact_link_b, src);
          st3(S, c_lay.contact_normal, j, ld3(S, c_lay.contact_normal, src));
          st3(S, c_lay.contact_pos, j, ld3(S, c_lay.contact_pos, src));
          FA(contact_penetration, j) = FA(contact_penetration, src); FA(contact_friction, j) = FA(contact_friction, src);
          IA(sort_idx, j) = j;
          j = src;
        }
        IA(contact_geom_a, j) = t_ga; IA(contact_geom_b, j) = t_gb; IA(contact_link_a, j) = t_la; IA(contact_link_b, j) = t_lb;
        st3(S, c_lay.contact_normal, j, t_n); st3(S, c_lay.contact_pos, j, t_p);
        FA(contact_penetration, j) = t_pen; FA(contact_friction, j) = t_fr;
        IA(sort_idx, j) = j;
      }
    }
  }
}

// =====================================================================================================
// Constraint rows (CS:579-1286), row order [equality | frictionloss | 4 x contacts | joint limits]
// =====================================================================================================
#define JAC(c, d) S.F[((size_t)(c_lay.jac + (c) * nD + (d))) * S.B]

__device__ __forceinline__ void dev_zero_row(const Env& S, int c) {
  const int nD = c_m.nD;
  for (int d = 0; d < nD; d++) JAC(c, d) = 0.f;
}
__device__ __forceinline__ void dev_store_row_params(const Env& S, int c, float diag, float aref) {
  FA(efc_diag, c) = diag; FA(efc_aref, c) = aref; FA(efc_D, c) = 1.0f / diag;
}

__device__ int dev_build_rows(const Env& S, int& n_eq, int& n_floss) {
  const int nD = c_m.nD;
  const float EPS = c_m.eps;
  int n_con = 0;
  n_eq = 0; n_floss = 0;
  // ---- equalities (CS:880-937)
  for (int i = 0; i < c_m.nEq; i++) {
    if (c_m.eq_type[i] == EQ_JOINT) {  // CS:800-877
      const int j1 = c_m.eq_obj1[i], j2 = c_m.eq_obj2[i];
      const int q1 = c_m.j_q_start[j1], q2 = c_m.j_q_start[j2], d1 = c_m.j_dof_start[j1], d2 = c_m.j_dof_start[j2];
      const int c = n_con++;
      n_eq++;
      dev_zero_row(S, c);
      const float diff = FA(qpos, q2) - c_m.qpos0[q2];
      float pos = FA(qpos, q1) - c_m.qpos0[q1], deriv = 0.f, dp = 1.f;
      for (int k = 0; k < 5; k++) {
        pos -= dp * c_m.eq_data[i][k];
        if (k < 4) deriv += c_m.eq_data[i][k + 1] * dp * (float)(k + 1);
        dp *= diff;
      }
      JAC(c, d1) = 1.0f;
      JAC(c, d2) = -deriv;
      const float jac_qvel = FA(vel, d1) + (-deriv) * FA(vel, d2);
      const float invweight = c_m.d_invweight[d1] + c_m.d_invweight[d2];
      float imp, aref;
      imp_aref(c_m.eq_sol[i], -fabsf(pos), jac_qvel, pos, imp, aref);
      dev_store_row_params(S, c, fmaxf(invweight * (1.0f - imp) / imp, EPS), aref);
    } else if (c_m.eq_type[i] == EQ_CONNECT) {  // CS:688-797
      const int l1 = c_m.eq_obj1[i], l2 = c_m.eq_obj2[i];
      const f3 g1 = transform(f3(c_m.eq_data[i][0], c_m.eq_data[i][1], c_m.eq_data[i][2]), ld3(S, c_lay.links_pos, l1), ld4(S, c_lay.links_quat, l1));
      const f3 g2 = transform(f3(c_m.eq_data[i][3], c_m.eq_data[i][4], c_m.eq_data[i][5]), ld3(S, c_lay.links_pos, l2), ld4(S, c_lay.links_quat, l2));
      const float invweight = c_m.l_invweight[l1] + c_m.l_invweight[l2];
      const f3 pd = g1 - g2;
      for (int i3 = 0; i3 < 3; i3++) {
        const int c = n_con++;
        n_eq++;
        dev_zero_row(S, c);
        float jac_qvel = 0.f;
        for (int ab = 0; ab < 2; ab++) {
          const float sign = ab == 0 ? 1.f : -1.f;
          int link = ab == 0 ? l1 : l2;
          const f3 pos = ab == 0 ? g1 : g2;
          while (link > -1) {
            const f3 tpos = pos - ld3(S, c_lay.root_COM, link);
            for (int d = c_m.l_dof_end[link] - 1; d >= c_m.l_dof_start[link]; d--) {
              const f3 vel = ld3(S, c_lay.cdof_vel, d) - cross(tpos, ld3(S, c_lay.cdof_ang, d));
              const float jac = sign * comp(vel, i3);
              jac_qvel += jac * FA(vel, d);
              JAC(c, d) += jac;
            }
            link = c_m.l_parent[link];
          }
        }
        float imp, aref;
        imp_aref(c_m.eq_sol[i], -norm(pd), jac_qvel, comp(pd, i3), imp, aref);
        dev_store_row_params(S, c, fmaxf(invweight * (1.0f - imp) / imp, EPS), aref);
      }
    }
  }
  // ---- frictionloss (CS:1236-1286)
  for (int j = 0; j < c_m.nJ; j++)
    for (int d = c_m.j_dof_start[j]; d < c_m.j_dof_start[j] + c_m.j_n_dofs[j]; d++)
      if (c_m.d_frictionloss[d] > EPS) {
        float imp, aref;
        imp_aref(c_m.j_sol[j], 0.f, FA(vel, d), 0.f, imp, aref);
        const int c = n_con++;
        n_floss++;
        dev_store_row_params(S, c, fmaxf(c_m.d_invweight[d] * (1.0f - imp) / imp, EPS), aref);
        FA(efc_frictionloss, c) = c_m.d_frictionloss[d];
        dev_zero_row(S, c);
        JAC(c, d) = 1.0f;
      }
  // ---- contacts (CS:579-685): pyramidal friction, 4 rows per contact
  if (c_m.enable_collision) {
    const int nc = IA(n_contacts, 0);
    for (int ic = 0; ic < nc; ic++) {
      const int la = IA(contact_link_a, ic), lb = IA(contact_link_b, ic), ga = IA(contact_geom_a, ic), gb = IA(contact_geom_b, ic);
      const f3 cpos = ld3(S, c_lay.contact_pos, ic), cn = ld3(S, c_lay.contact_normal, ic);
      const float mu = FA(contact_friction, ic), pen = FA(contact_penetration, ic);
      float sol[7];
#pragma unroll
      for (int k = 0; k < 7; k++) sol[k] = 0.5f * (c_m.g_sol[ga][k] + c_m.g_sol[gb][k]);
      f3 d1, d2;
      orthogonals(cn, d1, d2);
      float invweight = c_m.l_invweight[la];
      if (lb > -1) invweight += c_m.l_invweight[lb];
      const int c0 = n_con + ic * 4;
      for (int i = 0; i < 4; i++) dev_zero_row(S, c0 + i);
      float jq[4] = {0.f, 0.f, 0.f, 0.f};
      const f3 n0 = d1 * (-mu) - cn, n1 = d1 * mu - cn, n2 = d2 * (-mu) - cn, n3 = d2 * mu - cn;
      for (int ab = 0; ab < 2; ab++) {
        const float sign = ab == 0 ? -1.f : 1.f;
        int link = ab == 0 ? la : lb;
        while (link > -1) {
          const f3 tpos = cpos - ld3(S, c_lay.root_COM, link);
          for (int d = c_m.l_dof_end[link] - 1; d >= c_m.l_dof_start[link]; d--) {
            const f3 diff = (ld3(S, c_lay.cdof_vel, d) - cross(tpos, ld3(S, c_lay.cdof_ang, d))) * sign;
            const float v = FA(vel, d);
            const float j0 = dot(diff, n0), j1 = dot(diff, n1), j2 = dot(diff, n2), j3 = dot(diff, n3);
            jq[0] += j0 * v; jq[1] += j1 * v; jq[2] += j2 * v; jq[3] += j3 * v;
            JAC(c0, d) += j0; JAC(c0 + 1, d) += j1; JAC(c0 + 2, d) += j2; JAC(c0 + 3, d) += j3;
          }
          link = c_m.l_parent[link];
        }
      }
      for (int i = 0; i < 4; i++) {
        float imp, aref;
        imp_aref(sol, -pen, jq[i], -pen, imp, aref);
        float diag = invweight + mu * mu * invweight;
        diag *= 2.f * mu * mu * (1.0f - imp) / imp;
        dev_store_row_params(S, c0 + i, fmaxf(diag, EPS), aref);
      }
    }
    n_con += nc * 4;
  }
  // ---- joint limits (CS:1176-1233)
  if (c_m.enable_joint_limit) {
    for (int j = 0; j < c_m.nJ; j++) {
      const int type = c_m.j_type[j];
      if (type != JOINT_REVOLUTE && type != JOINT_PRISMATIC) continue;
      const int q = c_m.j_q_start[j], d = c_m.j_dof_start[j];
      const float qv = FA(qpos, q);
      const float pmin = qv - c_m.d_limit[d][0], pmax = c_m.d_limit[d][1] - qv;
      const float pd = fminf(pmin, pmax);
      if (pd < 0.f) {
        const float jac = pmin < pmax ? 1.f : -1.f;
        float imp, aref;
        imp_aref(c_m.j_sol[j], pd, jac * FA(vel, d), pd, imp, aref);
        const int c = n_con++;
        dev_store_row_params(S, c, fmaxf(c_m.d_invweight[d] * (1.0f - imp) / imp, EPS), aref);
        dev_zero_row(S, c);
        JAC(c, d) = jac;
      }
    }
  }
  return n_con;
}

// =====================================================================================================
// Constraint solver: Newton (or PR-CG) with exact line search, monolith semantics (CS:2033-3261)
// =====================================================================================================
struct SolverCtx {
  int n_con, ne, nef;
  float cost, prev_cost, gauss;
  float qg0, qg1, qg2, es0, es1, es2;  // quad_gauss, eq_sum
  int ls_it;
};
#define HM(i, j) S.F[((size_t)(c_lay.nt_H + (i) * nD + (j))) * S.B]
#define MM(i, j) S.F[((size_t)(c_lay.mass_mat + (i) * nD + (j))) * S.B]

// CS:2619-2694 func_update_constraint_batch; returns whether any active flag changed (bookkeeping for the caller)
__device__ void dev_update_constraint(const Env& S, SolverCtx& X, bool track_prev) {
  const int nD = c_m.nD;
  X.prev_cost = X.cost;
  float cost = 0.f, gauss = 0.f;
  for (int c = 0; c < X.n_con; c++) {
    const float Ja = FA(Jaref, c), D = FA(efc_D, c);
    const int prev = IA(efc_active, c) & 1;
    int act = 1;
    float floss_force = 0.f;
    if (X.ne <= c && c < X.nef) {
      const float f = FA(efc_frictionloss, c), rf = FA(efc_diag, c) * f;
      const bool ln = Ja <= -rf, lp = Ja >= rf;
      act = !(ln || lp);
      floss_force = (ln ? f : 0.f) + (lp ? -f : 0.f);
      cost += (ln ? f * (-0.5f * rf - Ja) : 0.f) + (lp ? f * (-0.5f * rf + Ja) : 0.f);
    } else if (X.nef <= c) {
      act = Ja < 0.f;
    }
    IA(efc_active, c) = act | (track_prev ? (prev << 1) : (act << 1));
    FA(efc_force, c) = floss_force + (-Ja * D * (float)act);
    cost += 0.5f * (Ja * Ja * D * (float)act);
  }
  for (int d = 0; d < nD; d++) {
    float q = 0.f;
    for (int c = 0; c < X.n_con; c++) q += JAC(c, d) * FA(efc_force, c);
    FA(qf_constraint, d) = q;
    const float v = 0.5f * (FA(Ma, d) - FA(force, d)) * (FA(qacc, d) - FA(acc_smooth, d));
    gauss += v;
  }
  X.gauss = gauss;
  X.cost = cost + gauss;
}

// CS:1394-1452 + CS:1594-1621: H = M + J^T D J (lower), then in-place Cholesky
__device__ void dev_hessian_cholesky(const Env& S, const SolverCtx& X) {
  const int nD = c_m.nD;
  const float EPS = c_m.eps;
  for (int i = 0; i < nD; i++)
    for (int j = 0; j <= i; j++) HM(i, j) = (c_m.d_entity[i] == c_m.d_entity[j]) ? MM(i, j) : 0.f;
  for (int c = 0; c < X.n_con; c++) {
    if (!(IA(efc_active, c) & 1)) continue;
    const float D = FA(efc_D, c);
    for (int i = 0; i < nD; i++) {
      const float Jci = JAC(c, i);
      if (fabsf(Jci) > EPS) {
        const float s = Jci * D;
        for (int j = 0; j <= i; j++) HM(i, j) += JAC(c, j) * s;
      }
    }
  }
  for (int i = 0; i < nD; i++) {
    float tmp = HM(i, i);
    for (int j = 0; j < i; j++) { const float h = HM(i, j); tmp -= h * h; }
    const float Lii = sqrtf(fmaxf(tmp, EPS));
    HM(i, i) = Lii;
    const float inv = 1.0f / Lii;
    for (int j = i + 1; j < nD; j++) {
      float dt = 0.f;
      for (int k = 0; k < i; k++) dt += HM(j, k) * HM(i, k);
      HM(j, i) = (HM(j, i) - dt) * inv;
    }
  }
}

// CS:1763-1822 rank-1 Cholesky up/down-dates for rows whose active flag flipped; true when degenerate
__device__ bool dev_hessian_incremental(const Env& S, const SolverCtx& X) {
  const int nD = c_m.nD;
  const float EPS = c_m.eps;
  for (int c = 0; c < X.n_con; c++) {
    const int a = IA(efc_active, c);
    if (!((a ^ (a >> 1)) & 1)) continue;
    const float sign = (a & 1) ? 1.f : -1.f;
    const float sq = sqrtf(FA(efc_D, c));
    for (int d = 0; d < nD; d++) FA(nt_vec, d) = JAC(c, d) * sq;
    for (int k = 0; k < nD; k++) {
      const float vk = FA(nt_vec, k);
      if (fabsf(vk) > EPS) {
        const float Lkk = HM(k, k);
        const float tmp = Lkk * Lkk + sign * vk * vk;
        if (tmp < EPS) return true;
        const float r = sqrtf(tmp), cc = r / Lkk, cinv = 1.0f / cc, s = vk / Lkk;
        HM(k, k) = r;
        for (int i = k + 1; i < nD; i++) {
          const float Lik = (HM(i, k) + s * FA(nt_vec, i) * sign) * cinv;
          HM(i, k) = Lik;
          FA(nt_vec, i) = FA(nt_vec, i) * cc - s * Lik;
        }
      }
    }
  }
  return false;
}

// CS:2721-2750 gradient and preconditioned gradient (Newton: Cholesky solve CS:1892-1910; CG: M^-1 via L^T D L)
__device__ void dev_update_gradient(const Env& S) {
  const int nD = c_m.nD;
  for (int d = 0; d < nD; d++) FA(grad, d) = FA(Ma, d) - FA(force, d) - FA(qf_constraint, d);
  if (c_m.solver_type == SOLVER_CG) {
    for (int i_e = 0; i_e < c_m.nE; i_e++) dev_solve_block(S, c_lay.grad, c_lay.Mgrad, c_m.e_dof_start[i_e], c_m.e_dof_end[i_e]);
  } else {
    for (int i = 0; i < nD; i++) {
      float c = FA(grad, i);
      for (int j = 0; j < i; j++) c -= HM(i, j) * FA(Mgrad, j);
      FA(Mgrad, i) = c / HM(i, i);
    }
    for (int i = nD - 1; i >= 0; i--) {
      float c = FA(Mgrad, i);
      for (int j = i + 1; j < nD; j++) c -= HM(j, i) * FA(Mgrad, j);
      FA(Mgrad, i) = c / HM(i, i);
    }
  }
}

struct LsPoint { float alpha, cost, d0, d1; };

// CS:2154-2354 accumulation of the piecewise-quadratic cost at step `alpha` over friction and contact rows
__device__ __forceinline__ LsPoint dev_ls_eval(const Env& S, const SolverCtx& X, float alpha) {
  float t0 = X.qg0 + X.es0, t1 = X.qg1 + X.es1, t2 = X.qg2 + X.es2;
  for (int c = X.ne; c < X.nef; c++) {
    const float Ja = FA(Jaref, c), jv = FA(jv, c), D = FA(efc_D, c), f = FA(efc_frictionloss, c), rf = FA(efc_diag, c) * f;
    float q0 = D * (0.5f * Ja * Ja), q1 = D * (jv * Ja), q2 = D * (0.5f * jv * jv);
    const float x = Ja + alpha * jv;
    const bool ln = x <= -rf, lp = x >= rf;
    if (ln || lp) { q0 = (ln ? f * (-0.5f * rf - Ja) : 0.f) + (lp ? f * (-0.5f * rf + Ja) : 0.f); q1 = (ln ? -f * jv : 0.f) + (lp ? f * jv : 0.f); q2 = 0.f; }
    t0 += q0; t1 += q1; t2 += q2;
  }
  for (int c = X.nef; c < X.n_con; c++) {
    const float Ja = FA(Jaref, c), jv = FA(jv, c), D = FA(efc_D, c);
    if (Ja + alpha * jv < 0.f) { t0 += D * (0.5f * Ja * Ja); t1 += D * (jv * Ja); t2 += D * (0.5f * jv * jv); }
  }
  LsPoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * t2 + alpha * t1 + t0;
  p.d0 = 2.f * alpha * t2 + t1;
  p.d1 = 2.f * t2;
  if (p.d1 <= 0.f) p.d1 = c_m.eps;
  return p;
}
// CS:2357-2388
__device__ __forceinline__ int dev_update_bracket(LsPoint& p, const LsPoint* cand, float& next_alpha) {
  int flag = 0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    if (p.d0 < 0.f && cand[i].d0 < 0.f && p.d0 < cand[i].d0) { p = cand[i]; flag = 1; }
    else if (p.d0 > 0.f && cand[i].d0 > 0.f && p.d0 > cand[i].d0) { p = cand[i]; flag = 2; }
  }
  next_alpha = p.alpha;
  if (flag > 0) next_alpha = p.alpha - p.d0 / p.d1;
  return flag;
}

// CS:2426-2597 func_linesearch_batch (includes CS:2033-2151 init + alpha=0 evaluation)
__device__ float dev_linesearch(const Env& S, SolverCtx& X) {
  const int nD = c_m.nD;
  const float EPS = c_m.eps;
  float snorm = 0.f;
  for (int d = 0; d < nD; d++) { const float s = FA(search, d); snorm += s * s; }
  snorm = sqrtf(snorm);
  const float gtol = c_m.tolerance * c_m.ls_tolerance * snorm * (c_m.meaninertia * (float)max(1, nD));
  X.ls_it = 0;
  if (snorm < EPS) return 0.f;
  // mv = M search, jv = J search, quadratic coefficients of the Gauss term
  float qg1 = 0.f, qg2 = 0.f;
  for (int i = 0; i < nD; i++) {
    const int e = c_m.d_entity[i];
    float mv = 0.f;
    for (int j = c_m.e_dof_start[e]; j < c_m.e_dof_end[e]; j++) mv += MM(i, j) * FA(search, j);
    FA(mv, i) = mv;
    const float s = FA(search, i);
    qg1 += s * FA(Ma, i) - s * FA(force, i);
    qg2 += 0.5f * s * mv;
  }
  X.qg0 = X.gauss; X.qg1 = qg1; X.qg2 = qg2;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int c = 0; c < X.n_con; c++) {
    float jv = 0.f;
    for (int d = 0; d < nD; d++) jv += JAC(c, d) * FA(search, d);
    FA(jv, c) = jv;
    if (c < X.ne) {
      const float Ja = FA(Jaref, c), D = FA(efc_D, c);
      s0 += D * (0.5f * Ja * Ja); s1 += D * (jv * Ja); s2 += D * (0.5f * jv * jv);
    }
  }
  X.es0 = s0; X.es1 = s1; X.es2 = s2;
  const LsPoint p0 = dev_ls_eval(S, X, 0.f);
  X.ls_it = 1;
  LsPoint p1 = dev_ls_eval(S, X, p0.alpha - p0.d0 / p0.d1);
  X.ls_it++;
  if (p0.cost < p1.cost) p1 = p0;
  if (fabsf(p1.d0) < gtol) return p1.alpha;
  const float direction = p1.d0 < 0.f ? 1.f : -1.f;
  bool p2update = false;
  LsPoint p2 = p1;
  while (p1.d0 * direction <= -gtol && X.ls_it < c_m.ls_iterations) {
    p2 = p1; p2update = true;
    p1 = dev_ls_eval(S, X, p1.alpha - p1.d0 / p1.d1);
    X.ls_it++;
    if (fabsf(p1.d0) < gtol) return p1.alpha;
  }
  if (X.ls_it >= c_m.ls_iterations) return p1.alpha;
  if (!p2update) return p1.alpha;
  float a0 = p1.alpha - p1.d0 / p1.d1, a1 = p1.alpha, a2 = (p1.alpha + p2.alpha) * 0.5f;
  while (X.ls_it < c_m.ls_iterations) {
    LsPoint cand[3];
    cand[0] = dev_ls_eval(S, X, a0); cand[1] = dev_ls_eval(S, X, a1); cand[2] = dev_ls_eval(S, X, a2);
    X.ls_it += 3;
    float best_alpha = 0.f, best_cost = 0.f;
    bool found = false;
#pragma unroll
    for (int i = 0; i < 3; i++)
      if (fabsf(cand[i].d0) < gtol && (!found || cand[i].cost < best_cost)) { best_alpha = cand[i].alpha; best_cost = cand[i].cost; found = true; }
    if (found) return best_alpha;
    float p1n, p2n;
    const int b1 = dev_update_bracket(p1, cand, p1n), b2 = dev_update_bracket(p2, cand, p2n);
    if (b1 == 0 && b2 == 0) return a2;
    a0 = p1n; a1 = p2n; a2 = (p1.alpha + p2.alpha) * 0.5f;
  }
  if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
  if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
  return 0.f;
}

// CS:210-245 resolve(): init (CS:2935-3059), monolith body (CS:3163-3190), update_qacc (CS:3240-3261),
// contact forces (CS:3198-3237)
__device__ void dev_solve(const Env& S, int n_con, int n_eq, int n_floss) {
  const int nD = c_m.nD;
  const float EPS = c_m.eps;
  SolverCtx X;
  X.n_con = n_con; X.ne = n_eq; X.nef = n_eq + n_floss; X.cost = 0.f; X.prev_cost = 0.f; X.gauss = 0.f; X.ls_it = 0;
  const bool warm = n_con > 0 && IA(is_warmstart, 0);
  for (int d = 0; d < nD; d++) FA(qacc, d) = warm ? FA(qacc_ws, d) : FA(acc_smooth, d);
  int iters = 0;
  if (n_con > 0) {
    for (int i = 0; i < nD; i++) {  // CS:2910-2929
      const int e = c_m.d_entity[i];
      float s = 0.f;
      for (int j = c_m.e_dof_start[e]; j < c_m.e_dof_end[e]; j++) s += MM(i, j) * FA(qacc, j);
      FA(Ma, i) = s;
    }
    for (int c = 0; c < n_con; c++) {  // CS:2887-2907
      float s = -FA(efc_aref, c);
      for (int d = 0; d < nD; d++) s += JAC(c, d) * FA(qacc, d);
      FA(Jaref, c) = s;
    }
    dev_update_constraint(S, X, false);
    if (c_m.solver_type == SOLVER_NEWTON) dev_hessian_cholesky(S, X);
    dev_update_gradient(S);
    for (int d = 0; d < nD; d++) FA(search, d) = -FA(Mgrad, d);
    const float tol_scaled = (c_m.meaninertia * (float)max(1, nD)) * c_m.tolerance;
    for (int it = 0; it < c_m.iterations; it++) {  // CS:3062-3139 func_solve_iter
      iters++;
      const float alpha = dev_linesearch(S, X);
      if (fabsf(alpha) < EPS) break;
      for (int d = 0; d < nD; d++) {
        FA(qacc, d) += FA(search, d) * alpha;
        FA(Ma, d) += FA(mv, d) * alpha;
        if (c_m.solver_type == SOLVER_CG) { FA(cg_prev_grad, d) = FA(grad, d); FA(cg_prev_Mgrad, d) = FA(Mgrad, d); }
      }
      for (int c = 0; c < n_con; c++) FA(Jaref, c) += FA(jv, c) * alpha;
      dev_update_constraint(S, X, true);
      if (c_m.solver_type == SOLVER_NEWTON) {
        if (dev_hessian_incremental(S, X)) dev_hessian_cholesky(S, X);
      }
      dev_update_gradient(S);
      // CS:2841-2884 termination / new search direction
      float gn = 0.f;
      for (int d = 0; d < nD; d++) { const float g = FA(grad, d); gn += g * g; }
      gn = sqrtf(gn);
      const bool improved = gn > tol_scaled && (X.prev_cost - X.cost) > tol_scaled;
      if (!improved) break;
      if (c_m.solver_type == SOLVER_NEWTON) {
        for (int d = 0; d < nD; d++) FA(search, d) = -FA(Mgrad, d);
      } else {
        float beta = 0.f, pg = 0.f;
        for (int d = 0; d < nD; d++) { beta += FA(grad, d) * (FA(Mgrad, d) - FA(cg_prev_Mgrad, d)); pg += FA(cg_prev_Mgrad, d) * FA(cg_prev_grad, d); }
        beta = fmaxf(beta / fmaxf(EPS, pg), 0.f);
        for (int d = 0; d < nD; d++) FA(search, d) = -FA(Mgrad, d) + beta * FA(search, d);
      }
    }
  } else {
    for (int d = 0; d < nD; d++) FA(qf_constraint, d) = 0.f;
  }
  IA(solver_iters, 0) = iters;
  // ---- CS:3240-3261 func_update_qacc
  bool has_nan = false;
  for (int d = 0; d < nD; d++) {
    const float a = FA(qacc, d);
    FA(acc, d) = a;
    FA(force, d) = FA(qf_smooth, d) + FA(qf_constraint, d);
    FA(qacc_ws, d) = a;
    has_nan |= isnan(a);
  }
  if (has_nan) IA(errno_, 0) |= ERR_NAN_FORCE;
  IA(is_warmstart, 0) = 1;
  // ---- CS:3198-3237 func_update_contact_force
  if (c_m.enable_collision && c_m.nP > 0) {
    for (int l = 0; l < c_m.nL; l++) st3(S, c_lay.links_contact_force, l, f3());
    const int nc = IA(n_contacts, 0), cs = X.nef;
    for (int ic = 0; ic < nc; ic++) {
      const f3 cn = ld3(S, c_lay.contact_normal, ic);
      const float mu = FA(contact_friction, ic);
      f3 d1, d2;
      orthogonals(cn, d1, d2);
      const float f0 = FA(efc_force, cs + ic * 4), f1 = FA(efc_force, cs + ic * 4 + 1), f2 = FA(efc_force, cs + ic * 4 + 2), f3_ = FA(efc_force, cs + ic * 4 + 3);
      const f3 force = (d1 * (-mu) - cn) * f0 + (d1 * mu - cn) * f1 + (d2 * (-mu) - cn) * f2 + (d2 * mu - cn) * f3_;
      st3(S, c_lay.contact_force, ic, force);
      const int la = IA(contact_link_a, ic), lb = IA(contact_link_b, ic);
      st3(S, c_lay.links_contact_force, la, ld3(S, c_lay.links_contact_force, la) - force);
      st3(S, c_lay.links_contact_force, lb, ld3(S, c_lay.links_contact_force, lb) + force);
    }
  }
}

// =====================================================================================================
// Integration (RS:2783-2891 kernel_step_2): link accelerations, semi-implicit Euler, NaN guard
// =====================================================================================================
__device__ bool dev_integrate(const Env& S) {
  const int nD = c_m.nD, nL = c_m.nL;
  const float dt = c_m.dt, EPS = c_m.eps;
  // FD:1186-1285 update_acc(update_cacc=True): consistent pos/vel/acc snapshot before integration
  for (int l = 0; l < nL; l++) {
    const int p = c_m.l_parent[l];
    f3 lin, ang;
    if (p != -1) { lin = ld3(S, c_lay.cacc_lin, p); ang = ld3(S, c_lay.cacc_ang, p); }
    for (int d = c_m.l_dof_start[l]; d < c_m.l_dof_end[l]; d++) {
      const float v = FA(vel, d), a = FA(acc, d);
      lin += ld3(S, c_lay.cdofd_vel, d) * v + ld3(S, c_lay.cdof_vel, d) * a;
      ang += ld3(S, c_lay.cdofd_ang, d) * v + ld3(S, c_lay.cdof_ang, d) * a;
    }
    st3(S, c_lay.cacc_lin, l, lin);
    st3(S, c_lay.cacc_ang, l, ang);
  }
  if (c_m.integrator != INTEGRATOR_APPROX_IMPLICITFAST) {
    // FD:1787-1859 func_implicit_damping: refactor M + dt*(damping - d act/d vel) and re-solve acc = M'^-1 force
    for (int i_e = 0; i_e < c_m.nE; i_e++) {
      const int ds = c_m.e_dof_start[i_e], de = c_m.e_dof_end[i_e];
      bool mask = false;
      for (int d = ds; d < de; d++) {
        if (c_m.d_damping[d] > EPS) mask = true;
        if (c_m.integrator != INTEGRATOR_EULER && IA(ctrl_mode, d) <= CTRL_VELOCITY && fabsf(c_m.d_act_bias[d][2]) > EPS) mask = true;
      }
      if (!mask) continue;
      for (int i = ds; i < de; i++) {
        for (int j = ds; j <= i; j++) Mref(S, c_lay.mass_L, i, j) = Mref(S, c_lay.mass_mat, i, j);
        float dd = c_m.d_damping[i] * dt;
        if (c_m.integrator == INTEGRATOR_IMPLICITFAST && IA(ctrl_mode, i) <= CTRL_VELOCITY) dd -= c_m.d_act_bias[i][2] * dt;
        Mref(S, c_lay.mass_L, i, i) += dd;
      }
      dev_factor_block(S, c_lay.mass_L, ds, de);
      dev_solve_block(S, c_lay.force, c_lay.acc, ds, de);
    }
  }
  // FD:1566-1707 func_integrate + diff.py:25-54 (commit only if finite)
  bool valid = true;
  for (int d = 0; d < nD; d++) {
    const float vn = FA(vel, d) + FA(acc, d) * dt;
    valid &= !isnan(vn);
    FA(mv, d) = vn;  // vel_next parked in the solver's mv rows (dead after the solve)
  }
  for (int l = 0; l < nL; l++) {
    const int ds = c_m.l_dof_start[l], qs = c_m.l_q_start[l], qe = c_m.l_q_end[l];
    if (c_m.l_dof_end[l] - ds <= 0) continue;
    const int jt = c_m.j_type[c_m.l_joint_start[l]];
    if (jt == JOINT_FREE || jt == JOINT_SPHERICAL) {
      const int ro = jt == JOINT_FREE ? 3 : 0;
      const f4 rot0(FA(qpos, qs + ro), FA(qpos, qs + ro + 1), FA(qpos, qs + ro + 2), FA(qpos, qs + ro + 3));
      const f3 ang = f3(FA(mv, ds + ro), FA(mv, ds + ro + 1), FA(mv, ds + ro + 2)) * dt;
      const f4 rot = quat_compose(rotvec_to_quat(ang, EPS), rot0);
      valid &= !(isnan(rot.w) || isnan(rot.x) || isnan(rot.y) || isnan(rot.z));
    }
  }
  if (!valid) { IA(errno_, 0) |= ERR_NAN_ACC; return false; }
  for (int l = 0; l < nL; l++) {
    const int ds = c_m.l_dof_start[l], qs = c_m.l_q_start[l], qe = c_m.l_q_end[l];
    if (c_m.l_dof_end[l] - ds <= 0) continue;
    const int jt = c_m.j_type[c_m.l_joint_start[l]];
    if (jt == JOINT_FREE)
      for (int k = 0; k < 3; k++) FA(qpos, qs + k) += FA(mv, ds + k) * dt;
    if (jt == JOINT_FREE || jt == JOINT_SPHERICAL) {
      const int ro = jt == JOINT_FREE ? 3 : 0;
      const f4 rot0(FA(qpos, qs + ro), FA(qpos, qs + ro + 1), FA(qpos, qs + ro + 2), FA(qpos, qs + ro + 3));
      const f3 ang = f3(FA(mv, ds + ro), FA(mv, ds + ro + 1), FA(mv, ds + ro + 2)) * dt;
      const f4 rot = quat_compose(rotvec_to_quat(ang, EPS), rot0);
      FA(qpos, qs + ro) = rot.w; FA(qpos, qs + ro + 1) = rot.x; FA(qpos, qs + ro + 2) = rot.y; FA(qpos, qs + ro + 3) = rot.z;
    } else {
      for (int k = 0; k < qe - qs; k++) FA(qpos, qs + k) += FA(mv, ds + k) * dt;
    }
  }
  for (int d = 0; d < nD; d++) FA(vel, d) = FA(mv, d);
  return true;
}

// =====================================================================================================
// Kernels
// =====================================================================================================
#define GSR_BLOCK 64

__device__ __forceinline__ bool make_env(Env& S, float* F, int* I, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return false;
  S.F = F + b; S.I = I + b; S.B = (size_t)B;
  return true;
}

__global__ void __launch_bounds__(GSR_BLOCK) k_dynamics(float* F, int* I, int B) {
  Env S;
  if (!make_env(S, F, I, B)) return;
  dev_dynamics(S);
}
__global__ void __launch_bounds__(GSR_BLOCK) k_collide(float* F, int* I, int B) {
  Env S;
  if (!make_env(S, F, I, B)) return;
  dev_collide(S);
}
__global__ void __launch_bounds__(GSR_BLOCK) k_constraint(float* F, int* I, int B) {
  Env S;
  if (!make_env(S, F, I, B)) return;
  int n_eq, n_floss;
  const int n_con = dev_build_rows(S, n_eq, n_floss);
  IA(n_constraints, 0) = n_con; IA(n_constraints_equality, 0) = n_eq; IA(n_constraints_frictionloss, 0) = n_floss;
  dev_solve(S, n_con, n_eq, n_floss);
}
__global__ void __launch_bounds__(GSR_BLOCK) k_integrate(float* F, int* I, int B, int clear_ext) {
  Env S;
  if (!make_env(S, F, I, B)) return;
  dev_integrate(S);
  dev_kinematics(S);
  if (clear_ext)  // RS:2603-2612 clear_external_force
    for (int l = 0; l < c_m.nL; l++) { st3(S, c_lay.cfrc_applied_vel, l, f3()); st3(S, c_lay.cfrc_applied_ang, l, f3()); }
}
// FK for a masked subset of envs (set_qpos / set_dofs_velocity / reset paths, RS:1545-1572)
__global__ void __launch_bounds__(GSR_BLOCK) k_forward(float* F, int* I, int B, const uint8_t* mask, int init_fixed) {
  Env S;
  if (!make_env(S, F, I, B)) return;
  if (mask && !mask[blockIdx.x * blockDim.x + threadIdx.x]) return;
  dev_kinematics(S);
  for (int g = 0; g < c_m.nG; g++) {
    if (c_m.g_is_fixed[g]) {
      if (!init_fixed) continue;
      const int l = c_m.g_link[g];
      const f3 lp = ld3(S, c_lay.links_pos, l);
      const f4 lq = ld4(S, c_lay.links_quat, l);
      st3(S, c_lay.geoms_pos, g, lp + rotate(c3(c_m.g_pos[g]), lq));
      st4(S, c_lay.geoms_quat, g, quat_compose(c4(c_m.g_quat[g]), lq));
    }
    dev_geom_aabb(S, g);
  }
}
// collider.clear + constraint_solver.reset for masked envs (collider.py:563-607, solver.py:114-179)
__global__ void __launch_bounds__(GSR_BLOCK) k_reset(float* F, int* I, int B, const uint8_t* mask) {
  Env S;
  if (!make_env(S, F, I, B)) return;
  if (mask && !mask[blockIdx.x * blockDim.x + threadIdx.x]) return;
  for (int d = 0; d < c_m.nD; d++) FA(qacc_ws, d) = 0.f;
  IA(is_warmstart, 0) = 0;
  for (int w = 0; w < c_m.n_cache_words; w++) IA(cache_valid, w) = 0;
  const int nc = IA(n_contacts, 0);
  for (int c = 0; c < nc; c++) {
    IA(contact_geom_a, c) = -1; IA(contact_geom_b, c) = -1; IA(contact_link_a, c) = -1; IA(contact_link_b, c) = -1;
    FA(contact_penetration, c) = 0.f;
    st3(S, c_lay.contact_pos, c, f3()); st3(S, c_lay.contact_normal, c, f3()); st3(S, c_lay.contact_force, c, f3());
  }
  IA(n_contacts, 0) = 0; IA(n_constraints, 0) = 0; IA(n_constraints_equality, 0) = 0; IA(n_constraints_frictionloss, 0) = 0;
  IA(errno_, 0) = 0;
}
