// gsr_dynamics.cuh -- forward kinematics and smooth dynamics, one lane group per environment, all operands in shared memory.
// Reference citations use the abbreviations of oracle/rigid_oracle.hpp (FK, FD, CS, BP, NP, MPR, SF, CT, BOX, RS).
//
// The reference walks the kinematic tree serially (one thread per entity, FK:580-736, FD:306-388, FD:1186-1400).  Here the
// tree recursions are restated as data-parallel sums: every c-frame quantity is expressed about the same point (the
// subtree COM) in world axes, so "propagate from the parent" is a plain sum over the ancestor set and "accumulate into the
// parent" a plain sum over the descendant set -- each lane evaluates its link's (or dof's) sum directly from a 64-bit
// ancestor mask, in the reference's summation order.  Only the pose composition of FK is a true scan; it is done by pointer
// jumping in ceil(log2(depth)) rounds.
#pragma once
#include "gsr_group.cuh"
#undef IN
#undef WK
#if GSR_LDS_MASK & 1
#define IN(field) IN_SHARED(field)
#define WK(field) WK_SHARED(field)
#else
#define IN(field) IN_GENERIC(field)
#define WK(field) WK_GENERIC(field)
#endif

// ---------------------------------------------------------------------------------------------------------------------
// FK:580-736 one link's pose from its parent's world pose (pp, pq); also joint anchors / axes and dofs_pos when `final`.
// ---------------------------------------------------------------------------------------------------------------------
template <int G>
__device__ GSR_FK_NOINLINE void fk_link(const Ctx<G>& C, int l, bool has_parent, f3 pp, f4 pq, const float* qpos, f3& pos, f4& quat, bool final) {
  const float EPS = CM(eps);
  pos = f3(GM(l_pos[l][0]), GM(l_pos[l][1]), GM(l_pos[l][2]));
  quat = f4(GM(l_quat[l][0]), GM(l_quat[l][1]), GM(l_quat[l][2]), GM(l_quat[l][3]));
  if (has_parent) {
    pos = pp + rotate(pos, pq);
    quat = quat_compose(quat, pq);
  }
  const int js = GM(l_joint_start[l]), je = GM(l_joint_end[l]);
  for (int j = js; j < je; j++) {
    const int type = GM(j_type[j]), qs = GM(j_q_start[j]), ds = GM(j_dof_start[j]);
    const f3 jpos(GM(j_pos[j][0]), GM(j_pos[j][1]), GM(j_pos[j][2]));
    f3 xanchor, xaxis(0.f, 0.f, 1.f);
    if (type == JOINT_FREE) {
      pos = f3(qpos[qs], qpos[qs + 1], qpos[qs + 2]);
      const f4 q(qpos[qs + 3], qpos[qs + 4], qpos[qs + 5], qpos[qs + 6]);
      const float n = 1.0f / sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
      quat = f4(q.w * n, q.x * n, q.y * n, q.z * n);
      xanchor = pos;
      if (final) {
        const f3 xyz = quat_to_xyz(quat, EPS);
        float* dp = IN(dofs_pos);
        dp[ds] = pos.x; dp[ds + 1] = pos.y; dp[ds + 2] = pos.z; dp[ds + 3] = xyz.x; dp[ds + 4] = xyz.y; dp[ds + 5] = xyz.z;
      }
    } else if (type == JOINT_REVOLUTE) {
      const f3 axis(GM(d_motion_ang[ds][0]), GM(d_motion_ang[ds][1]), GM(d_motion_ang[ds][2]));
      xanchor = rotate(jpos, quat) + pos;
      xaxis = rotate(axis, quat);
      const float dq = qpos[qs] - GM(qpos0[qs]);
      if (final) IN(dofs_pos)[ds] = dq;
      quat = quat_compose(rotvec_to_quat(axis * dq, EPS), quat);
      pos = xanchor - rotate(jpos, quat);
    } else if (type == JOINT_PRISMATIC) {
      const f3 axis(GM(d_motion_vel[ds][0]), GM(d_motion_vel[ds][1]), GM(d_motion_vel[ds][2]));
      xanchor = rotate(jpos, quat) + pos;
      xaxis = rotate(axis, quat);
      const float dq = qpos[qs] - GM(qpos0[qs]);
      if (final) IN(dofs_pos)[ds] = dq;
      pos = pos + xaxis * dq;
    } else if (type == JOINT_SPHERICAL) {
      xanchor = rotate(jpos, quat) + pos;
      const f4 qloc(qpos[qs], qpos[qs + 1], qpos[qs + 2], qpos[qs + 3]);
      if (final) {
        const f3 xyz = quat_to_xyz(qloc, EPS);
        float* dp = IN(dofs_pos);
        dp[ds] = xyz.x; dp[ds + 1] = xyz.y; dp[ds + 2] = xyz.z;
      }
      quat = quat_compose(qloc, quat);
      pos = xanchor - rotate(jpos, quat);
    }
    if (final) {
      st3(WK(xanchor), j, xanchor);
      st3(WK(xaxis), j, xaxis);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
#ifndef GSR_KIN_INLINE
#define GSR_KIN_INLINE __noinline__
#endif
// Kinematics of the whole environment from qpos / vel (FK:580-736, FK:337-576, FK:993-1116): rebuilds the kinematic cache.
// ---------------------------------------------------------------------------------------------------------------------
template <int G>
__device__ GSR_KIN_INLINE void phase_kinematics(const Ctx<G>& C) {  // one copy for its two call sites in k_step (start / epilogue)
  const int nL = CM(nL), nD = CM(nD), nJ = CM(nJ);
  const float EPS = CM(eps);
  const int lane = C.lane;
  const float* qpos = IN(qpos);
  const float* vel = IN(vel);
  float* lpos = IN(links_pos);
  float* lquat = IN(links_quat);
  // ---- 1. local transforms X_l (pose relative to the parent frame; absolute for roots), pointer-jumping scan to world poses
  float* pa = WK(pose_a);
  float* pb = WK(pose_b);
  int* anc_a = reinterpret_cast<int*>(pa) + 7 * nL;
  int* anc_b = reinterpret_cast<int*>(pb) + 7 * nL;
  for (int l = lane; l < nL; l += G) {
    const int p = GM(l_parent[l]);
    f3 pos; f4 quat;
    if (p == -1 && GM(l_is_fixed[l])) { pos = ld3(lpos, l); quat = ld4(lquat, l); }  // fixed roots keep the user-set pose (FK:732-736)
    else fk_link(C, l, false, f3(), f4(), qpos, pos, quat, false);
    st3(pa, l, pos); st4(pa + 3 * nL, l, quat);
    anc_a[l] = p;
  }
  C.sync();
  int rounds = 0;
  for (int d = 1; d < CM(max_depth); d <<= 1) rounds++;
  for (int r = 0; r < rounds; r++) {
    for (int l = lane; l < nL; l += G) {
      const int a = anc_a[l];
      f3 pos = ld3(pa, l); f4 quat = ld4(pa + 3 * nL, l);
      int na = a;
      if (a >= 0) {
        const f3 ap = ld3(pa, a); const f4 aq = ld4(pa + 3 * nL, a);
        pos = ap + rotate(pos, aq);
        quat = quat_compose(quat, aq);
        na = anc_a[a];
      }
      st3(pb, l, pos); st4(pb + 3 * nL, l, quat);
      anc_b[l] = na;
    }
    C.sync();
    float* t = pa; pa = pb; pb = t;
    int* ti = anc_a; anc_a = anc_b; anc_b = ti;
  }
  // ---- 2. final pose of every link the reference way, from the parent's scanned world pose; anchors, axes, dofs_pos
  for (int l = lane; l < nL; l += G) {
    const int p = GM(l_parent[l]);
    if (p == -1 && GM(l_is_fixed[l])) continue;
    f3 pp; f4 pq;
    if (p >= 0) { pp = ld3(pa, p); pq = ld4(pa + 3 * nL, p); }
    f3 pos; f4 quat;
    fk_link(C, l, p >= 0, pp, pq, qpos, pos, quat, true);
    st3(lpos, l, pos); st4(lquat, l, quat);
  }
  C.sync();
  // ---- 3. subtree COM (FK:337-420): mass-weighted mean of the inertial frame origins of every link below the same root
  float* mi = pb;  // scratch: m_l * ipos_w[l] (3 nL words) in the scan buffer that does not hold the result
  const bool shifted = C.SI()[c_h.has_shift] != 0;
  for (int l = lane; l < nL; l += G) {
    const f3 ipw = ld3(lpos, l) + rotate(link_ipos(C, l, shifted), ld4(lquat, l));
    st3(mi, l, ipw * link_mass(C, l, shifted));
  }
  C.sync();
  float* rcom = IN(root_COM);
  for (int l = lane; l < nL; l += G) {
    const int r = GM(l_root[l]);
    f3 s;
    float ms = shifted ? 0.f : GM(l_root_mass[r]);
    for (int k = 0; k < nL; k++)
      if (CM(l_root[k]) == r) { s += ld3(mi, k); if (shifted) ms += CM(l_mass[k]) + C.rec[c_h.mass_shift + k]; }
    f3 com;
    if (ms > EPS) com = s / ms;
    else com = ld3(lpos, r) + rotate(link_ipos(C, r, shifted), ld4(lquat, r));
    st3(rcom, l, com);
  }
  C.sync();
  // ---- 4. c-frame inertia (geom.py:320-336) per link, motion subspaces per joint (FK:517-576)
  for (int l = lane; l < nL; l += G) {
    const f3 com = ld3(rcom, l);
    const f4 quat = ld4(lquat, l);
    const f3 ipos = ld3(lpos, l) + rotate(link_ipos(C, l, shifted), quat) - com;
    const f4 iquat = quat_compose(f4(GM(l_iquat[l][0]), GM(l_iquat[l][1]), GM(l_iquat[l][2]), GM(l_iquat[l][3])), quat);
    const float mass = link_mass(C, l, shifted);
    const mat3 R = quat_to_R(iquat, EPS);
    float Iin[9];
#pragma unroll
    for (int k = 0; k < 9; k++) Iin[k] = GM(l_inertia[l][k]);
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
    st6(IN(cinr_I), l, I);
    st3(IN(cinr_pos), l, ipos * mass);
  }
  float* cda = IN(cdof_ang);
  float* cdv = IN(cdof_vel);
  for (int j = lane; j < nJ; j += G) {
    const int type = GM(j_type[j]), ds = GM(j_dof_start[j]);
    if (type == JOINT_FIXED) continue;
    const int l = GM(d_link[ds]);
    const f3 off = ld3(rcom, l) - ld3(WK(xanchor), j);
    const f3 xaxis = ld3(WK(xaxis), j);
    if (type == JOINT_REVOLUTE) {
      st3(cda, ds, xaxis); st3(cdv, ds, cross(xaxis, off));
    } else if (type == JOINT_PRISMATIC) {
      st3(cda, ds, f3()); st3(cdv, ds, xaxis);
    } else {
      const mat3 R = quat_to_R(ld4(lquat, l), EPS);
      int o = ds;
      if (type == JOINT_FREE) {
        for (int k = 0; k < 3; k++) { f3 v; set_comp(v, k, 1.f); st3(cda, ds + k, f3()); st3(cdv, ds + k, v); }
        o = ds + 3;
      }
      for (int k = 0; k < 3; k++) { const f3 a = mat_col(R, k); st3(cda, o + k, a); st3(cdv, o + k, cross(a, off)); }
    }
  }
  C.sync();
  // ---- 5. c-frame velocities (FK:993-1116): cd[l] = sum over the dofs on the path to the root; cdof_dot[d] = cvel x cdof[d]
  // with cvel the velocity accumulated before joint(d)'s own dofs (after the translational dofs for a free joint).
  for (int l = lane; l < nL; l += G) {
    const unsigned long long mask = GM(l_dofmask[l]);
    f3 vv, va;
    FOR_BITS_UP(mask, d) { const float w = vel[d]; vv += ld3(cdv, d) * w; va += ld3(cda, d) * w; } END_BITS
    st3(IN(cd_vel), l, vv); st3(IN(cd_ang), l, va);
  }
  for (int d = lane; d < nD; d += G) {
    const int l = GM(d_link[d]), j = GM(d_joint[d]);
    const int type = GM(j_type[j]), ds = GM(j_dof_start[j]);
    const unsigned long long mask = GM(l_dofmask[l]);
    int lim = ds;
    if (type == JOINT_FREE && d >= ds + 3) lim = ds + 3;
    f3 vv, va;
    FOR_BITS_UP(mask & ((1ull << lim) - 1ull), k) { const float w = vel[k]; vv += ld3(cdv, k) * w; va += ld3(cda, k) * w; } END_BITS
    f3 da, dv;
    if (!(type == JOINT_FREE && d < ds + 3)) {  // geom.py:379-383 motion_cross_motion(cvel, cdof)
      const f3 a = ld3(cda, d), v = ld3(cdv, d);
      da = cross(va, a);
      dv = cross(va, v) + cross(vv, a);
    }
    st3(IN(cdofd_ang), d, da); st3(IN(cdofd_vel), d, dv);
  }
  C.sync();
}

// geom pose from its link's pose (FK:828-864)
template <int G>
__device__ __forceinline__ void geom_pose_from_link(const Ctx<G>& C, int g, f3& gp, f4& gq) {
  const int l = GM(g_link[g]);
  const f3 lp = ld3(IN(links_pos), l);
  const f4 lq = ld4(IN(links_quat), l);
  gp = lp + rotate(f3(GM(g_pos[g][0]), GM(g_pos[g][1]), GM(g_pos[g][2])), lq);
  gq = quat_compose(f4(GM(g_quat[g][0]), GM(g_quat[g][1]), GM(g_quat[g][2]), GM(g_quat[g][3])), lq);
}
// ... as the narrow phase reads it: from the collision scratch (models whose layout keeps the poses: c_s.gpos >= 0) or recomputed
// with the same arithmetic (small-group models, where 7 words per geom of workspace cost resident environments)
template <int G>
__device__ __forceinline__ void geom_pose(const Ctx<G>& C, int g, f3& gp, f4& gq) {
  if (c_s.gpos >= 0) { gp = ld3(WK(gpos), g); gq = ld4(WK(gquat), g); }
  else geom_pose_from_link(C, g, gp, gq);
}

// geom poses (collision scratch, when kept) and world AABBs (FK:1297-1319)
template <int G>
__device__ void phase_geoms(const Ctx<G>& C) {
  const int nG = CM(nG);
  float* aabb = WK(aabb);
  const bool keep = c_s.gpos >= 0;
  for (int g = C.lane; g < nG; g += G) {
    f3 gp;
    f4 gq;
    geom_pose_from_link(C, g, gp, gq);
    if (keep) { st3(WK(gpos), g, gp); st4(WK(gquat), g, gq); }
    // the reference transforms the 8 corners of the rest-pose box and takes min / max; the identical box is
    // centre +- |R| * half-extent
    const mat3 R = quat_to_R(gq, 0.f);
    const f3 lo(GM(g_aabb_lo[g][0]), GM(g_aabb_lo[g][1]), GM(g_aabb_lo[g][2])), hi(GM(g_aabb_hi[g][0]), GM(g_aabb_hi[g][1]), GM(g_aabb_hi[g][2]));
    const f3 c = (lo + hi) * 0.5f, h = (hi - lo) * 0.5f;
    const f3 wc(R.m[0][0] * c.x + R.m[0][1] * c.y + R.m[0][2] * c.z + gp.x, R.m[1][0] * c.x + R.m[1][1] * c.y + R.m[1][2] * c.z + gp.y,
                R.m[2][0] * c.x + R.m[2][1] * c.y + R.m[2][2] * c.z + gp.z);
    const f3 wh(fabsf(R.m[0][0]) * h.x + fabsf(R.m[0][1]) * h.y + fabsf(R.m[0][2]) * h.z,
                fabsf(R.m[1][0]) * h.x + fabsf(R.m[1][1]) * h.y + fabsf(R.m[1][2]) * h.z,
                fabsf(R.m[2][0]) * h.x + fabsf(R.m[2][1]) * h.y + fabsf(R.m[2][2]) * h.z);
    st3(aabb, g, wc - wh); st3(aabb + 3 * nG, g, wc + wh);
  }
  C.sync();
  // accessor-visible copy (get_AABB): the scratch is recycled by the solver rows
  float* o_min = C.rec + c_h.aabb_min;
  float* o_max = C.rec + c_h.aabb_max;
  for (int i = C.lane; i < 3 * nG; i += G) { o_min[i] = aabb[i]; o_max[i] = aabb[3 * nG + i]; }
}

// ---------------------------------------------------------------------------------------------------------------------
// L^T D L solve (FD:813-894) of the vector x (shared memory, in place): x <- L^-1 D^-1 L^-T x
// ---------------------------------------------------------------------------------------------------------------------
template <int G>
__device__ GSR_MASS_NOINLINE void mass_solve(const Ctx<G>& C, float* x) {
  const int nD = CM(nD), MS = c_s.MS;
  const float* L = WKS(L);
  const float* Dinv = WKS(Dinv);
  if (nD <= G) {
    // one lane per unknown, the running right-hand side in a register; the pivot value travels by shuffle, so the
    // 2 (nD - 1) elimination steps need no shared-memory round trip and no barrier (same products, same order)
    const int lane = C.lane;
    const bool own = lane < nD;
    float s = own ? x[lane] : 0.f;
    for (int j = nD - 1; j > 0; j--) {  // L^T w = y
      const float xj = __shfl_sync(C.gmask, s, j, G);
      if (lane < j) s -= L[j * MS + lane] * xj;
    }
    if (own) s *= Dinv[lane];
    for (int j = 0; j < nD - 1; j++) {  // L x = z
      const float xj = __shfl_sync(C.gmask, s, j, G);
      if (own && lane > j) s -= L[lane * MS + j] * xj;
    }
    if (own) x[lane] = s;
    C.sync();
    return;
  }
  if (nD <= 2 * G) {
    // same scheme with two unknowns per lane (rows lane and lane + G): Franka's 9 dofs on 8 lanes
    const int lane = C.lane, r1 = lane + G;
    const bool own1 = r1 < nD;
    float s0 = x[lane], s1 = own1 ? x[r1] : 0.f;
    for (int j = nD - 1; j > 0; j--) {  // L^T w = y
      const float xj = __shfl_sync(C.gmask, j < G ? s0 : s1, j < G ? j : j - G, G);
      if (lane < j) s0 -= L[j * MS + lane] * xj;
      if (r1 < j) s1 -= L[j * MS + r1] * xj;
    }
    s0 *= Dinv[lane];
    if (own1) s1 *= Dinv[r1];
    for (int j = 0; j < nD - 1; j++) {  // L x = z
      const float xj = __shfl_sync(C.gmask, j < G ? s0 : s1, j < G ? j : j - G, G);
      if (lane > j) s0 -= L[lane * MS + j] * xj;
      if (own1 && r1 > j) s1 -= L[r1 * MS + j] * xj;
    }
    x[lane] = s0;
    if (own1) x[r1] = s1;
    C.sync();
    return;
  }
  for (int j = nD - 1; j > 0; j--) {  // L^T w = y: once w_j is final, eliminate it from every i < j
    const float xj = x[j];
    for (int i = C.lane; i < j; i += G) x[i] -= L[j * MS + i] * xj;
    C.sync();
  }
  for (int i = C.lane; i < nD; i += G) x[i] *= Dinv[i];
  C.sync();
  for (int j = 0; j < nD - 1; j++) {  // L x = z
    const float xj = x[j];
    for (int i = j + 1 + C.lane; i < nD; i += G) x[i] -= L[i * MS + j] * xj;
    C.sync();
  }
}

// FD:545-602 factor of the matrix stored in WK(L) (lower part), in place; D^-1 to WK(Dinv)
template <int G>
__device__ GSR_MASS_NOINLINE void mass_factor(const Ctx<G>& C) {
  const int nD = CM(nD), MS = c_s.MS;
  float* L = WKS(L);
  float* Dinv = WKS(Dinv);
  for (int i = nD - 1; i >= 0; i--) {
    const float dinv = 1.0f / L[i * MS + i];
    // (j, k), k <= j < i, over the ancestor dofs of i only: entries of other branches are structurally zero and stay zero
    const int t1 = CM(fac_start[i + 1]);
    for (int t = CM(fac_start[i]) + C.lane; t < t1; t += G) {
      const unsigned tr = __ldg(CM(fac_tasks) + t);
      const int j = (int)(tr >> 8), k = (int)(tr & 255u);
      L[j * MS + k] -= (L[i * MS + j] * dinv) * L[i * MS + k];
    }
    C.sync();
    for (int j = C.lane; j < i; j += G) L[i * MS + j] *= dinv;
    if (C.lane == 0) { Dinv[i] = dinv; L[i * MS + i] = 1.0f; }
    C.sync();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Smooth dynamics (FD:146-227): CRBA, L^T D L, actuation, RNE bias, acc_smooth = M^-1 qf_smooth
// ---------------------------------------------------------------------------------------------------------------------
template <int G>
__device__ void phase_dynamics(const Ctx<G>& C) {
  const int nL = CM(nL), nD = CM(nD), MS = c_s.MS;
  const float dt = CM(dt);
  const int lane = C.lane;
  const float* vel = IN(vel);
  const float* cda = IN(cdof_ang);
  const float* cdv = IN(cdof_vel);
  // ---- composite rigid body inertia (FD:306-388): crb[l] = sum of cinr over the subtree of l (leaf-to-root order)
  const bool shifted = C.SI()[c_h.has_shift] != 0;
  float* crb = WK(crb);
  for (int l = lane; l < nL; l += G) {
    float s[10];
#pragma unroll
    for (int k = 0; k < 10; k++) s[k] = 0.f;
    FOR_BITS_DOWN(GM(l_descmask[l]), d) {
      const float* I = IN(cinr_I) + 6 * d;
      const float* P = IN(cinr_pos) + 3 * d;
#pragma unroll
      for (int k = 0; k < 6; k++) s[k] += I[k];
#pragma unroll
      for (int k = 0; k < 3; k++) s[6 + k] += P[k];
      s[9] += link_mass(C, d, shifted);
    } END_BITS
#pragma unroll
    for (int k = 0; k < 10; k++) crb[10 * l + k] = s[k];
  }
  C.sync();
  // ---- f_d = crb[link(d)] * cdof_d (FD:390-430)
  float* fdof = WK(fdof);
  for (int d = lane; d < nD; d += G) {
    const int l = GM(d_link[d]);
    const float* c = crb + 10 * l;
    sym3 I; I.xx = c[0]; I.xy = c[1]; I.xz = c[2]; I.yy = c[3]; I.yz = c[4]; I.zz = c[5];
    f3 fa, fv;
    inertial_mul(f3(c[6], c[7], c[8]), I, c[9], ld3(cdv, d), ld3(cda, d), fa, fv);
    st3(fdof, 2 * d, fa); st3(fdof, 2 * d + 1, fv);
  }
  C.sync();
  // ---- mass matrix (FD:432-541): M[i][j] = f_i . cdof_j for dof j on the path of dof i (lower triangle), mirrored
  float* M = WK(M);
  float* L = WK(L);
  const int ntri = nD * (nD + 1) / 2;
  const int* mode = reinterpret_cast<const int*>(C.rec) + c_h.ctrl_mode;  // PRE block: read from the record
  for (int t = lane; t < ntri; t += G) {
    int i, j;
    tri_decode(t, i, j);
    float v = 0.f;
    if (bit64(GM(l_dofmask[GM(d_link[i])]), j)) v = dot(ld3(fdof, 2 * i), ld3(cda, j)) + dot(ld3(fdof, 2 * i + 1), ld3(cdv, j));
    if (i == j) {
      v += dof_armature(C, i);
      if (CM(integrator) == INTEGRATOR_APPROX_IMPLICITFAST) {
        v += dof_damping(C, i) * dt;
        if (mode[i] <= CTRL_VELOCITY) v -= dof_bias(C, i, 2) * dt;
      }
    }
    M[t] = v;  // M: packed lower triangle in task order (gsr_group.cuh sym_row_dot)
  }
  C.sync();
  // the factorisation works on a copy, in the square layout of the dense-algebra routines (f_d may have lived in that slot until now)
  for (int t = lane; t < ntri; t += G) {
    int i, j;
    tri_decode(t, i, j);
    L[i * MS + j] = M[t];
  }
  C.sync();
  mass_factor(C);

  // ---- actuation + passive forces (FD:956-1182)
  float* qf_applied = WK(qf_applied);
  float* qf_passive = WK(qf_passive);
  const float* dofs_pos = IN(dofs_pos);
  for (int d = lane; d < nD; d += G) {
    const int j = GM(d_joint[d]);
    const int jt = GM(j_type[j]), ds = GM(j_dof_start[j]);
    const int md = mode[d];
    const float v = vel[d], dp = dofs_pos[d];
    const float kp = dof_kp(C, d), b0 = dof_bias(C, d, 0), b1 = dof_bias(C, d, 1), b2 = dof_bias(C, d, 2);
    float force = 0.f;
    if (md == CTRL_FORCE) force = C.rec[c_h.ctrl_force + d];
    else if (md == CTRL_VELOCITY) force = -b2 * (C.rec[c_h.ctrl_vel + d] - v);
    else if (md == CTRL_POSITION) {
      if (!(jt == JOINT_FREE && d >= ds + 3)) {
        force = kp * (IN(ctrl_pos)[d] - dp) + b0 + (kp + b1) * dp + b2 * (v - C.rec[c_h.ctrl_vel + d]);
      }
    }
    if (jt == JOINT_FREE && d >= ds + 3 && (mode[ds + 3] == CTRL_POSITION || mode[ds + 4] == CTRL_POSITION || mode[ds + 5] == CTRL_POSITION)) {
      // rotational PD of a free joint acts on the rotation-vector error (FD:1040-1090)
      const f3 xyz(dofs_pos[ds + 3], dofs_pos[ds + 4], dofs_pos[ds + 5]);
      const f3 cxyz(IN(ctrl_pos)[ds + 3], IN(ctrl_pos)[ds + 4], IN(ctrl_pos)[ds + 5]);
      const f3 rv = quat_to_rotvec(quat_compose(xyz_to_quat(cxyz), quat_inv(xyz_to_quat(xyz))), CM(eps));
      force = kp * comp(rv, d - ds - 3) + b0 + (kp + b1) * dp + b2 * (v - C.rec[c_h.ctrl_vel + d]);
    }
    const float applied = fminf(fmaxf(force, dof_force_range(C, d, 0)), dof_force_range(C, d, 1));
    qf_applied[d] = applied;                // workspace slot shared with `acc` (free until the solver ends), read back below
    C.rec[c_h.qf_applied + d] = applied;    // accessor-visible copy
    float passive = -dof_damping(C, d) * v;
    if (jt != JOINT_FREE && jt != JOINT_FIXED) passive += -dp * dof_stiffness(C, d);
    qf_passive[d] = passive;
  }
  // ---- RNE bias forces (FD:1186-1486)
  float* cdd = WK(cdd);    // per link: vel3, ang3
  float* cfrc = WK(cfrc);  // per link: vel3, ang3 (own contribution)
  const f3 grav(CM(gravity[0]), CM(gravity[1]), CM(gravity[2]));
  const bool has_ext = C.SI()[c_h.has_ext_force] != 0;
  for (int l = lane; l < nL; l += G) {
    const unsigned long long mask = GM(l_dofmask[l]);
    f3 cdd_vel = -grav * (1.0f - GM(e_gravcomp[GM(l_entity[l])])), cdd_ang;
    FOR_BITS_UP(mask, d) { const float w = vel[d]; cdd_vel += ld3(IN(cdofd_vel), d) * w; cdd_ang += ld3(IN(cdofd_ang), d) * w; } END_BITS
    const sym3 I = ld6(IN(cinr_I), l);
    const f3 ip = ld3(IN(cinr_pos), l);
    const float mass = link_mass(C, l, shifted);
    const f3 cv = ld3(IN(cd_vel), l), ca = ld3(IN(cd_ang), l);
    f3 f1a, f1v, f2a, f2v;
    inertial_mul(ip, I, mass, cdd_vel, cdd_ang, f1a, f1v);
    inertial_mul(ip, I, mass, cv, ca, f2a, f2v);
    const f3 f3v = cross(ca, f2v);  // geom.py:372-376 motion_cross_force
    const f3 f3a = cross(ca, f2a) + cross(cv, f2v);
    f3 ev, ea;
    if (has_ext) { ev = ld3(C.rec + c_h.cfrc_applied_vel, l); ea = ld3(C.rec + c_h.cfrc_applied_ang, l); }
    st3(cfrc, 2 * l, f1v + f3v + ev);
    st3(cfrc, 2 * l + 1, f1a + f3a + ea);
  }
  C.sync();
  for (int l = lane; l < nL; l += G) {  // subtree totals, leaf-to-root order (FD:1375-1400)
    f3 sv, sa;
    FOR_BITS_DOWN(GM(l_descmask[l]), d) { sv += ld3(cfrc, 2 * d); sa += ld3(cfrc, 2 * d + 1); } END_BITS
    st3(cdd, 2 * l, sv); st3(cdd, 2 * l + 1, sa);
  }
  C.sync();
  float* force = WK(force);
  float* qf_bias = WK(qf_bias);
  float* acc_smooth = WK(acc_smooth);
  for (int d = lane; d < nD; d += G) {
    const int l = GM(d_link[d]);
    const float bias = dot(ld3(cda, d), ld3(cdd, 2 * l + 1)) + dot(ld3(cdv, d), ld3(cdd, 2 * l));
    qf_bias[d] = bias;
    const float f = qf_passive[d] - bias + qf_applied[d];
    force[d] = f;
    acc_smooth[d] = f;
  }
  C.sync();
  mass_solve(C, acc_smooth);  // FD:1506-1563
}

// ---------------------------------------------------------------------------------------------------------------------
// Integration (RS:2783-2891 kernel_step_2): link accelerations, implicit damping, semi-implicit Euler with NaN guard
// ---------------------------------------------------------------------------------------------------------------------
template <int G>
__device__ void phase_integrate(const Ctx<G>& C) {
  const int nL = CM(nL), nD = CM(nD), nQ = CM(nQ), MS = c_s.MS;
  const float dt = CM(dt), EPS = CM(eps);
  const int lane = C.lane;
  float* vel = IN(vel);
  float* qpos = IN(qpos);
  float* acc = WK(acc);
  // FD:1186-1285 update_acc(update_cacc=True): link accelerations from the constrained acceleration
  float* cacc = WK(cacc);
  for (int l = lane; l < nL; l += G) {
    const unsigned long long mask = GM(l_dofmask[l]);
    f3 lin, ang;
    FOR_BITS_UP(mask, d) {
      const float v = vel[d], a = acc[d];
      lin += ld3(IN(cdofd_vel), d) * v + ld3(IN(cdof_vel), d) * a;
      ang += ld3(IN(cdofd_ang), d) * v + ld3(IN(cdof_ang), d) * a;
    } END_BITS
    st3(cacc, 2 * l, lin); st3(cacc, 2 * l + 1, ang);
  }
  if (CM(integrator) != INTEGRATOR_APPROX_IMPLICITFAST) {
    // FD:1787-1859 func_implicit_damping: refactor M + dt (damping - d act / d vel) and re-solve acc = M'^-1 force
    const int* mode = reinterpret_cast<const int*>(C.rec) + c_h.ctrl_mode;  // PRE block: read from the record
    auto dof_damped = [&](int k) {
      return dof_damping(C, k) > EPS || (CM(integrator) != INTEGRATOR_EULER && mode[k] <= CTRL_VELOCITY && fabsf(dof_bias(C, k, 2)) > EPS);
    };
    // the reference refactors entity by entity and skips entities without damping / velocity-dependent actuation (FD:1806-1832)
    auto ent_active = [&](int e) {
      bool a = false;
      for (int k = 0; k < nD; k++)
        if (CM(d_entity[k]) == e && dof_damped(k)) a = true;
      return a;
    };
    bool need = false;
    for (int d = 0; d < nD; d++) need |= dof_damped(d);
    if (need) {
      float* L = WK(L);
      const float* M = WK(M);
      const int ntri = nD * (nD + 1) / 2;
      C.sync();
      for (int t = lane; t < ntri; t += G) {
        int i, j;
        tri_decode(t, i, j);
        float v = M[t];
        if (i == j && ent_active(GM(d_entity[i]))) {
          v += dof_damping(C, i) * dt;
          if (CM(integrator) == INTEGRATOR_IMPLICITFAST && mode[i] <= CTRL_VELOCITY) v -= dof_bias(C, i, 2) * dt;
        }
        L[i * MS + j] = v;
      }
      C.sync();
      mass_factor(C);
      float* x = WK(vel_next);  // scratch (the next velocity is formed below)
      for (int d = lane; d < nD; d += G) x[d] = WK(force)[d] + WK(qfrc_c)[d];  // dofs_state.force = qf_smooth + qf_constraint
      C.sync();
      mass_solve(C, x);
      for (int d = lane; d < nD; d += G)
        if (ent_active(GM(d_entity[d]))) acc[d] = x[d];
    }
  }
  C.sync();
  // FD:1566-1707 func_integrate + diff.py:25-54 (commit only if finite)
  float* vn = WK(vel_next);
  float* qn = WK(qpos_next);
  bool bad = false;
  for (int d = lane; d < nD; d += G) { const float v = vel[d] + acc[d] * dt; vn[d] = v; bad |= isnan(v); }
  C.sync();
  for (int l = lane; l < nL; l += G) {
    const int ds = GM(l_dof_start[l]), de = GM(l_dof_end[l]), qs = GM(l_q_start[l]), qe = GM(l_q_end[l]);
    if (de - ds <= 0) continue;
    const int jt = GM(j_type[GM(l_joint_start[l])]);
    if (jt == JOINT_FREE)
      for (int k = 0; k < 3; k++) qn[qs + k] = qpos[qs + k] + vn[ds + k] * dt;
    if (jt == JOINT_FREE || jt == JOINT_SPHERICAL) {
      const int ro = jt == JOINT_FREE ? 3 : 0;
      const f4 rot0(qpos[qs + ro], qpos[qs + ro + 1], qpos[qs + ro + 2], qpos[qs + ro + 3]);
      const f3 ang = f3(vn[ds + ro], vn[ds + ro + 1], vn[ds + ro + 2]) * dt;
      const f4 rot = quat_compose(rotvec_to_quat(ang, EPS), rot0);
      qn[qs + ro] = rot.w; qn[qs + ro + 1] = rot.x; qn[qs + ro + 2] = rot.y; qn[qs + ro + 3] = rot.z;
    } else {
      for (int k = 0; k < qe - qs; k++) qn[qs + k] = qpos[qs + k] + vn[ds + k] * dt;
    }
  }
  C.sync();
  for (int q = lane; q < nQ; q += G) bad |= isnan(qn[q]);
  const bool invalid = C.any(bad);
  if (invalid) {
    if (lane == 0) C.SI()[c_h.errno_] |= ERR_NAN_ACC;
  } else {
    for (int d = lane; d < nD; d += G) vel[d] = vn[d];
    for (int q = lane; q < nQ; q += G) qpos[q] = qn[q];
  }
  C.sync();
}
