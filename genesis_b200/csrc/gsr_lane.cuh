// gsr_lane.cuh -- lane-per-environment step kernel (k_step_lane): one thread owns one environment for a whole substep.
//
// Why a second formulation next to the lane-group kernel (gsr_kernels.cuh): for small trees at very large batch
// (Franka, 30000 envs) the group kernel spends ~10x the instructions of serial code on lane bookkeeping (partial passes of
// 9 dofs / 12 links over 8 lanes, ballots, group barriers, lane-varying model-table loads).  With one environment per
// lane every loop is warp-uniform: model tables become constant-bank operands, there are no shuffles or barriers inside a
// phase, and the tree recursions are the reference's own serial recursions (FK:580-736, FD:306-388, FD:1186-1400).
//
// Memory: the workspace of the 32 environments of a warp is interleaved in shared memory, word w of lane t at
// S[w * 33 + t] (stride 33: conflict-free both for the per-lane accesses of the math -- all lanes read the same w -- and for
// the transposed, coalesced record <-> workspace copies where the 32 lanes move 32 consecutive words of ONE environment).
// HBM keeps the env-major records of gsr_types.h, so the public (B, item) views and the lane-group kernels (k_forward for
// the set_qpos paths) are unchanged.  The kinematic cache of the record is not read back: the step recomputes the
// kinematics from qpos / vel (bitwise the values the previous step wrote) and only writes the accessor-visible outputs.
//
// Every function is __host__ __device__ and free of warp intrinsics outside lane_coop_*: tests/lane_emu compiles this
// header for the host (stride 1, one "lane") to check the arithmetic against the oracle without a GPU.  That harness is test
// infrastructure; libgsrigid.so only contains the device kernel.
#pragma once
#include "gsr_group.cuh"
#include "gsr_epa.cuh"

#ifdef __CUDA_ARCH__
#define GSR_LS 33
#else
#define GSR_LS 1
#endif
#define GSR_LANE_STRIDE_DEV 33
#ifdef GSR_LANE_EMU  // tests/lane_emu: the same functions compiled for the host as well
#define LHD __host__ __device__
#else
#define LHD __device__
#endif

__constant__ LaneLayout c_l;

// lane-varying model-table reads (narrow phase after divergence) go through the read-only path like the group kernel
#ifdef __CUDA_ARCH__
#define LG(x) (__ldg(&(C.gm->x)))
#define LDG(p) __ldg(p)
#else
#define LG(x) (c_m.x)
#define LDG(p) (*(p))
#endif

LHD __forceinline__ int lane_ctz64(unsigned long long m) {
#ifdef __CUDA_ARCH__
  return __ffsll((long long)m) - 1;
#else
  return __builtin_ctzll(m);
#endif
}
LHD __forceinline__ int lane_f2i(float f) { int i; memcpy(&i, &f, 4); return i; }
LHD __forceinline__ float lane_i2f(int i) { float f; memcpy(&f, &i, 4); return f; }
LHD __forceinline__ void lane_atomic_or(int* p, int v) {
#ifdef __CUDA_ARCH__
  atomicOr(p, v);
#else
  *p |= v;
#endif
}
LHD __forceinline__ void lane_atomic_and(int* p, int v) {
#ifdef __CUDA_ARCH__
  atomicAnd(p, v);
#else
  *p &= v;
#endif
}

struct LArr {  // interleaved workspace array of this lane
  float* p;
  LHD __forceinline__ float& operator[](int i) const { return p[i * GSR_LS]; }
};
struct LIArr {
  int* p;
  LHD __forceinline__ int& operator[](int i) const { return p[i * GSR_LS]; }
};
struct GArr {  // array that lives either in the workspace (es = lane stride) or in the env's global record (es = 1)
  float* p;
  int es;
  LHD __forceinline__ float& operator[](int i) const { return p[i * es]; }
};
LHD __forceinline__ f3 ld3(LArr a, int i) { return f3(a[3 * i], a[3 * i + 1], a[3 * i + 2]); }
LHD __forceinline__ void st3(LArr a, int i, f3 v) { a[3 * i] = v.x; a[3 * i + 1] = v.y; a[3 * i + 2] = v.z; }
LHD __forceinline__ f4 ld4(LArr a, int i) { return f4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]); }
LHD __forceinline__ void st4(LArr a, int i, f4 v) { a[4 * i] = v.w; a[4 * i + 1] = v.x; a[4 * i + 2] = v.y; a[4 * i + 3] = v.z; }
LHD __forceinline__ f3 gld3(const float* p, int i) { return f3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
LHD __forceinline__ void gst3(float* p, int i, f3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }
LHD __forceinline__ int tri_idx(int i, int j) { return ((i * (i + 1)) >> 1) + j; }  // j <= i
LHD __forceinline__ int sym_idx(int i, int j) { return i >= j ? tri_idx(i, j) : tri_idx(j, i); }

struct LaneCtx {
  float* S;         // this lane's workspace base (word w at S[w * GSR_LS])
  float* rec;       // this env's HOT record
  float* contacts;  // CONTACT records
  float* cache;     // CACHE record
  float* spill;     // SPILL record
  const DevModel* gm;
  // cooperative record <-> workspace copies
  float* Sw;        // workspace base of the warp (lane 0)
  float* rec0;      // record of the warp's first environment
  int n_live, lane;
  bool live;        // false for the padding lanes of the last warp (they shadow the last live environment)
  LHD __forceinline__ LArr A(int off) const { return LArr{S + off * GSR_LS}; }
  LHD __forceinline__ LIArr I(int off) const { return LIArr{reinterpret_cast<int*>(S) + off * GSR_LS}; }
};

// ---- cooperative copies: 32 lanes move 32 consecutive words of one environment per instruction (coalesced) ----------------
LHD __forceinline__ void lane_coop_load(const LaneCtx& C, int w0, int rec_off, int n) {
#ifdef __CUDA_ARCH__
  __syncwarp();
  for (int e = 0; e < 32; e++) {  // padding lanes of the last warp shadow its last live environment
    const float* src = C.rec0 + (size_t)(e < C.n_live ? e : C.n_live - 1) * c_h.hot_stride + rec_off;
    for (int k = C.lane; k < n; k += 32) C.Sw[(w0 + k) * GSR_LS + e] = src[k];
  }
  __syncwarp();
#else
  for (int k = 0; k < n; k++) C.S[w0 + k] = C.rec[rec_off + k];
#endif
}
LHD __forceinline__ void lane_coop_flush(const LaneCtx& C, int w0, int rec_off, int n) {
#ifdef __CUDA_ARCH__
  __syncwarp();
  for (int e = 0; e < C.n_live; e++) {
    float* dst = C.rec0 + (size_t)e * c_h.hot_stride + rec_off;
    for (int k = C.lane; k < n; k += 32) dst[k] = C.Sw[(w0 + k) * GSR_LS + e];
  }
  __syncwarp();
#else
  for (int k = 0; k < n; k++) C.rec[rec_off + k] = C.S[w0 + k];
#endif
}

// per-environment inertial parameters: model value plus the domain-randomisation shift of the record (FK:377-382, 437)
LHD __forceinline__ float lane_link_mass(const LaneCtx& C, int l, bool shifted) {
  float m = CM(l_mass[l]);
  if (shifted) m += C.rec[c_h.mass_shift + l];
  return m;
}
LHD __forceinline__ f3 lane_link_ipos(const LaneCtx& C, int l, bool shifted) {
  f3 p(CM(l_ipos[l][0]), CM(l_ipos[l][1]), CM(l_ipos[l][2]));
  if (shifted) p += gld3(C.rec + c_h.com_shift, l);
  return p;
}

// ---------------------------------------------------------------------------------------------------------------------
// Kinematics (FK:580-736, FK:337-576, FK:993-1116) from qpos / vel: link poses, dofs_pos, subtree COM, motion subspaces,
// c-frame velocities (to workspace offset `cvel_off`: vel[3 nL] then ang[3 nL]) and cdof_dot.
// ---------------------------------------------------------------------------------------------------------------------
LHD __noinline__ void lane_fk_link(const LaneCtx& C, int l, bool has_parent, f3 pp, f4 pq, f3& pos, f4& quat) {
  const float EPS = CM(eps);
  const LArr qpos = C.A(c_h.qpos), dofs_pos = C.A(c_h.dofs_pos);
  const LArr cda = C.A(c_l.cdof), cdv = C.A(c_l.cdof + 3 * CM(nD));
  pos = f3(CM(l_pos[l][0]), CM(l_pos[l][1]), CM(l_pos[l][2]));
  quat = f4(CM(l_quat[l][0]), CM(l_quat[l][1]), CM(l_quat[l][2]), CM(l_quat[l][3]));
  if (has_parent) {
    pos = pp + rotate(pos, pq);
    quat = quat_compose(quat, pq);
  }
  const int js = CM(l_joint_start[l]), je = CM(l_joint_end[l]);
  for (int j = js; j < je; j++) {
    const int type = CM(j_type[j]), qs = CM(j_q_start[j]), ds = CM(j_dof_start[j]);
    const f3 jpos(CM(j_pos[j][0]), CM(j_pos[j][1]), CM(j_pos[j][2]));
    f3 xanchor, xaxis(0.f, 0.f, 1.f);
    if (type == JOINT_FIXED) continue;
    if (type == JOINT_FREE) {
      pos = f3(qpos[qs], qpos[qs + 1], qpos[qs + 2]);
      const f4 q(qpos[qs + 3], qpos[qs + 4], qpos[qs + 5], qpos[qs + 6]);
      const float n = 1.0f / sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
      quat = f4(q.w * n, q.x * n, q.y * n, q.z * n);
      xanchor = pos;
      const f3 xyz = quat_to_xyz(quat, EPS);
      dofs_pos[ds] = pos.x; dofs_pos[ds + 1] = pos.y; dofs_pos[ds + 2] = pos.z;
      dofs_pos[ds + 3] = xyz.x; dofs_pos[ds + 4] = xyz.y; dofs_pos[ds + 5] = xyz.z;
    } else if (type == JOINT_REVOLUTE) {
      const f3 axis(CM(d_motion_ang[ds][0]), CM(d_motion_ang[ds][1]), CM(d_motion_ang[ds][2]));
      xanchor = rotate(jpos, quat) + pos;
      xaxis = rotate(axis, quat);
      const float dq = qpos[qs] - CM(qpos0[qs]);
      dofs_pos[ds] = dq;
      quat = quat_compose(rotvec_to_quat(axis * dq, EPS), quat);
      pos = xanchor - rotate(jpos, quat);
    } else if (type == JOINT_PRISMATIC) {
      const f3 axis(CM(d_motion_vel[ds][0]), CM(d_motion_vel[ds][1]), CM(d_motion_vel[ds][2]));
      xanchor = rotate(jpos, quat) + pos;
      xaxis = rotate(axis, quat);
      const float dq = qpos[qs] - CM(qpos0[qs]);
      dofs_pos[ds] = dq;
      pos = pos + xaxis * dq;
    } else {  // JOINT_SPHERICAL
      xanchor = rotate(jpos, quat) + pos;
      const f4 qloc(qpos[qs], qpos[qs + 1], qpos[qs + 2], qpos[qs + 3]);
      const f3 xyz = quat_to_xyz(qloc, EPS);
      dofs_pos[ds] = xyz.x; dofs_pos[ds + 1] = xyz.y; dofs_pos[ds + 2] = xyz.z;
      quat = quat_compose(qloc, quat);
      pos = xanchor - rotate(jpos, quat);
    }
    // parked in the joint's first dof slot until the subtree COM is known (converted in lane_kinematics)
    st3(cda, ds, xaxis);
    st3(cdv, ds, xanchor);
  }
}

LHD __noinline__ void lane_kinematics(const LaneCtx& C, int cvel_off) {
  const int nL = CM(nL), nD = CM(nD), nJ = CM(nJ);
  const float EPS = CM(eps);
  const LArr lpos = C.A(c_l.lpos), lquat = C.A(c_l.lquat), rcom = C.A(c_l.rcom);
  const LArr vel = C.A(c_h.vel);
  const LArr cda = C.A(c_l.cdof), cdv = C.A(c_l.cdof + 3 * nD), cdda = C.A(c_l.cdof + 6 * nD), cddv = C.A(c_l.cdof + 9 * nD);
  // ---- 1. poses root -> leaf, mass-weighted inertial origins per tree
  const bool shifted = C.I(c_h.has_shift)[0] != 0;
  for (int r = 0; r < 3 * CM(n_roots); r++) rcom[r] = 0.f;
  for (int l = 0; l < nL; l++) {
    const int p = CM(l_parent[l]);
    f3 pos; f4 quat;
    if (p == -1 && CM(l_is_fixed[l])) {  // fixed roots keep the user-set pose (FK:732-736)
      pos = gld3(C.rec + c_h.links_pos, l);
      const float* q = C.rec + c_h.links_quat + 4 * l;
      quat = f4(q[0], q[1], q[2], q[3]);
    } else {
      f3 pp; f4 pq;
      if (p >= 0) { pp = ld3(lpos, p); pq = ld4(lquat, p); }
      lane_fk_link(C, l, p >= 0, pp, pq, pos, quat);
    }
    st3(lpos, l, pos); st4(lquat, l, quat);
    const f3 ipw = pos + rotate(lane_link_ipos(C, l, shifted), quat);
    const int rs = CM(l_rootslot[l]);
    st3(rcom, rs, ld3(rcom, rs) + ipw * lane_link_mass(C, l, shifted));
  }
  // ---- 2. subtree COM (FK:337-420)
  for (int l = 0; l < nL; l++) {
    if (CM(l_parent[l]) != -1) continue;
    const int rs = CM(l_rootslot[l]);
    float ms = CM(l_root_mass[l]);
    if (shifted) {
      ms = 0.f;
      for (int k = 0; k < nL; k++)
        if (CM(l_root[k]) == l) ms += lane_link_mass(C, k, true);
    }
    f3 com;
    if (ms > EPS) com = ld3(rcom, rs) / ms;
    else com = ld3(lpos, l) + rotate(lane_link_ipos(C, l, shifted), ld4(lquat, l));
    st3(rcom, rs, com);
  }
  // ---- 3. motion subspaces (FK:517-576) from the parked axis / anchor
  for (int j = 0; j < nJ; j++) {
    const int type = CM(j_type[j]), ds = CM(j_dof_start[j]);
    if (type == JOINT_FIXED) continue;
    const int l = CM(d_link[ds]);
    const f3 xaxis = ld3(cda, ds);
    const f3 off = ld3(rcom, CM(l_rootslot[l])) - ld3(cdv, ds);
    if (type == JOINT_REVOLUTE) {
      st3(cdv, ds, cross(xaxis, off));
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
  // ---- 4. c-frame velocities root -> leaf and cdof_dot = cvel x cdof (FK:993-1116, geom.py:379-383)
  const LArr cv = C.A(cvel_off), ca = C.A(cvel_off + 3 * nL);
  for (int l = 0; l < nL; l++) {
    const int p = CM(l_parent[l]);
    f3 vv, va;
    if (p >= 0) { vv = ld3(cv, p); va = ld3(ca, p); }
    const int js = CM(l_joint_start[l]), je = CM(l_joint_end[l]);
    for (int j = js; j < je; j++) {
      const int type = CM(j_type[j]), ds = CM(j_dof_start[j]), nd = CM(j_n_dofs[j]);
      int d0 = ds;
      if (type == JOINT_FREE) {  // the translational dofs come first and have no cdof_dot
        for (int d = ds; d < ds + 3; d++) {
          st3(cdda, d, f3()); st3(cddv, d, f3());
          const float w = vel[d];
          vv += ld3(cdv, d) * w; va += ld3(cda, d) * w;
        }
        d0 = ds + 3;
      }
      for (int d = d0; d < ds + nd; d++) {
        const f3 a = ld3(cda, d), v = ld3(cdv, d);
        st3(cdda, d, cross(va, a));
        st3(cddv, d, cross(va, v) + cross(vv, a));
      }
      for (int d = d0; d < ds + nd; d++) { const float w = vel[d]; vv += ld3(cdv, d) * w; va += ld3(cda, d) * w; }
    }
    st3(cv, l, vv); st3(ca, l, va);
  }
}

// c-frame inertia of every link (geom.py:320-336): I[6 nL] then first moment pos[3 nL] at workspace offset `off`
LHD __noinline__ void lane_cinr(const LaneCtx& C, int off) {
  const int nL = CM(nL);
  const float EPS = CM(eps);
  const LArr lpos = C.A(c_l.lpos), lquat = C.A(c_l.lquat), rcom = C.A(c_l.rcom);
  const LArr cI = C.A(off), cP = C.A(off + 6 * nL);
  const bool shifted = C.I(c_h.has_shift)[0] != 0;
  for (int l = 0; l < nL; l++) {
    const f3 com = ld3(rcom, CM(l_rootslot[l]));
    const f4 quat = ld4(lquat, l);
    const f3 ipos = ld3(lpos, l) + rotate(lane_link_ipos(C, l, shifted), quat) - com;
    const f4 iquat = quat_compose(f4(CM(l_iquat[l][0]), CM(l_iquat[l][1]), CM(l_iquat[l][2]), CM(l_iquat[l][3])), quat);
    const float mass = lane_link_mass(C, l, shifted);
    const mat3 R = quat_to_R(iquat, EPS);
    float RI[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++)
        RI[a][b] = R.m[a][0] * CM(l_inertia[l][b]) + R.m[a][1] * CM(l_inertia[l][3 + b]) + R.m[a][2] * CM(l_inertia[l][6 + b]);
    const float x = ipos.x, y = ipos.y, z = ipos.z;
    cI[6 * l + 0] = RI[0][0] * R.m[0][0] + RI[0][1] * R.m[0][1] + RI[0][2] * R.m[0][2] + (y * y + z * z) * mass;
    cI[6 * l + 1] = RI[0][0] * R.m[1][0] + RI[0][1] * R.m[1][1] + RI[0][2] * R.m[1][2] - x * y * mass;
    cI[6 * l + 2] = RI[0][0] * R.m[2][0] + RI[0][1] * R.m[2][1] + RI[0][2] * R.m[2][2] - x * z * mass;
    cI[6 * l + 3] = RI[1][0] * R.m[1][0] + RI[1][1] * R.m[1][1] + RI[1][2] * R.m[1][2] + (x * x + z * z) * mass;
    cI[6 * l + 4] = RI[1][0] * R.m[2][0] + RI[1][1] * R.m[2][1] + RI[1][2] * R.m[2][2] - y * z * mass;
    cI[6 * l + 5] = RI[2][0] * R.m[2][0] + RI[2][1] * R.m[2][1] + RI[2][2] * R.m[2][2] + (x * x + y * y) * mass;
    st3(cP, l, ipos * mass);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// L^T D L factor / solve on packed lower triangles (FD:545-602, FD:813-894)
// ---------------------------------------------------------------------------------------------------------------------
LHD __noinline__ void lane_mass_factor(const LaneCtx& C) {
  const int nD = CM(nD);
  const LArr L = C.A(c_l.L), Dinv = C.A(c_l.Dinv);
  for (int i = nD - 1; i >= 0; i--) {
    const int ri = tri_idx(i, 0);
    const float dinv = 1.0f / L[ri + i];
    for (int j = 0; j < i; j++) {
      const float a = L[ri + j] * dinv;
      if (a == 0.f) continue;  // structural zeros (branches of the tree) are identical in every lane
      const int rj = tri_idx(j, 0);
      for (int k = 0; k <= j; k++) L[rj + k] -= a * L[ri + k];
    }
    for (int j = 0; j < i; j++) L[ri + j] *= dinv;
    Dinv[i] = dinv;
    L[ri + i] = 1.0f;
  }
}
LHD __noinline__ void lane_mass_solve(const LaneCtx& C, LArr x) {
  const int nD = CM(nD);
  const LArr L = C.A(c_l.L), Dinv = C.A(c_l.Dinv);
  for (int j = nD - 1; j > 0; j--) {
    const float xj = x[j];
    const int rj = tri_idx(j, 0);
    for (int i = 0; i < j; i++) x[i] -= L[rj + i] * xj;
  }
  for (int i = 0; i < nD; i++) x[i] *= Dinv[i];
  for (int i = 1; i < nD; i++) {  // L x = z (row oriented: x_i -= sum_{j<i} L_ij x_j, same products as the column sweep)
    const int ri = tri_idx(i, 0);
    float s = x[i];
    for (int j = 0; j < i; j++) s -= L[ri + j] * x[j];
    x[i] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Smooth dynamics (FD:146-227): CRBA, L^T D L, actuation, RNE bias, acc_smooth = M^-1 qf_smooth
// ---------------------------------------------------------------------------------------------------------------------
LHD __noinline__ void lane_dynamics(const LaneCtx& C) {
  const int nL = CM(nL), nD = CM(nD);
  const float dt = CM(dt);
  const LArr vel = C.A(c_h.vel), dofs_pos = C.A(c_h.dofs_pos);
  const LIArr mode = C.I(c_h.ctrl_mode);
  const LArr cda = C.A(c_l.cdof), cdv = C.A(c_l.cdof + 3 * nD), cdda = C.A(c_l.cdof + 6 * nD), cddv = C.A(c_l.cdof + 9 * nD);
  const LArr crbI = C.A(c_l.crb), crbP = C.A(c_l.crb + 6 * nL);
  const LArr cfv = C.A(c_l.cfrc), cfa = C.A(c_l.cfrc + 3 * nL);
  const LArr cv = C.A(c_l.cvel), ca = C.A(c_l.cvel + 3 * nL);
  const LArr cdd_v = C.A(c_l.cdd), cdd_a = C.A(c_l.cdd + 3 * nL);
  lane_cinr(C, c_l.crb);
  const bool shifted = C.I(c_h.has_shift)[0] != 0;
  // ---- RNE bias forces, root -> leaf (FD:1186-1400): cdd = parent's + cdof_dot * vel; own cfrc = I cdd + cvel x* (I cvel)
  const f3 grav(CM(gravity[0]), CM(gravity[1]), CM(gravity[2]));
  const bool has_ext = C.I(c_h.has_ext_force)[0] != 0;
  for (int l = 0; l < nL; l++) {
    const int p = CM(l_parent[l]);
    f3 dv, da;
    if (p >= 0) { dv = ld3(cdd_v, p); da = ld3(cdd_a, p); }
    else dv = -grav * (1.0f - CM(e_gravcomp[CMI(l_entity[l])]));
    for (int d = CM(l_dof_start[l]); d < CM(l_dof_end[l]); d++) { const float w = vel[d]; dv += ld3(cddv, d) * w; da += ld3(cdda, d) * w; }
    st3(cdd_v, l, dv); st3(cdd_a, l, da);
    sym3 I;
    I.xx = crbI[6 * l]; I.xy = crbI[6 * l + 1]; I.xz = crbI[6 * l + 2]; I.yy = crbI[6 * l + 3]; I.yz = crbI[6 * l + 4]; I.zz = crbI[6 * l + 5];
    const f3 ip = ld3(crbP, l);
    const float mass = lane_link_mass(C, l, shifted);
    const f3 lv = ld3(cv, l), la = ld3(ca, l);
    f3 f1a, f1v, f2a, f2v;
    inertial_mul(ip, I, mass, dv, da, f1a, f1v);
    inertial_mul(ip, I, mass, lv, la, f2a, f2v);
    const f3 f3v = cross(la, f2v);  // geom.py:372-376 motion_cross_force
    const f3 f3a = cross(la, f2a) + cross(lv, f2v);
    f3 ev, ea;
    if (has_ext) { ev = gld3(C.rec + c_h.cfrc_applied_vel, l); ea = gld3(C.rec + c_h.cfrc_applied_ang, l); }
    st3(cfv, l, f1v + f3v + ev);
    st3(cfa, l, f1a + f3a + ea);
  }
  // ---- leaf -> root: subtree wrench (FD:1375-1400) and composite inertia (FD:306-388)
  for (int l = nL - 1; l >= 0; l--) {
    const int p = CM(l_parent[l]);
    if (p < 0) continue;
    st3(cfv, p, ld3(cfv, p) + ld3(cfv, l));
    st3(cfa, p, ld3(cfa, p) + ld3(cfa, l));
#pragma unroll
    for (int k = 0; k < 6; k++) crbI[6 * p + k] += crbI[6 * l + k];
#pragma unroll
    for (int k = 0; k < 3; k++) crbP[3 * p + k] += crbP[3 * l + k];
  }
  // ---- mass matrix rows (FD:390-541), packed lower triangle; M also seeds the factor
  const LArr M = C.A(c_l.M), L = C.A(c_l.L);
  const LArr qf_applied = C.A(c_l.qf_applied), force = C.A(c_l.force), acc_smooth = C.A(c_l.acc_smooth);
  const bool diag_out = CM(diagnostics) && C.live;
  for (int i = 0; i < nD; i++) {
    const int l = CM(d_link[i]);
    sym3 I;
    I.xx = crbI[6 * l]; I.xy = crbI[6 * l + 1]; I.xz = crbI[6 * l + 2]; I.yy = crbI[6 * l + 3]; I.yz = crbI[6 * l + 4]; I.zz = crbI[6 * l + 5];
    f3 fa, fv;
    float submass = CM(l_submass[l]);
    if (shifted) {  // composite mass of the subtree with the per-environment shifts (leaf-to-root sum like FD:382-388)
      submass = 0.f;
      for (int k = nL - 1; k >= l; k--)
        if ((CM(l_descmask[l]) >> k) & 1ull) submass += lane_link_mass(C, k, true);
    }
    inertial_mul(ld3(crbP, l), I, submass, ld3(cdv, i), ld3(cda, i), fa, fv);
    const unsigned long long mask = CM(l_dofmask[l]);
    const int ri = tri_idx(i, 0);
    for (int j = 0; j <= i; j++) {
      float v = 0.f;
      if ((mask >> j) & 1ull) v = dot(fa, ld3(cda, j)) + dot(fv, ld3(cdv, j));
      if (i == j) {
        v += CM(d_armature[i]);
        if (CM(integrator) == INTEGRATOR_APPROX_IMPLICITFAST) {
          v += CM(d_damping[i]) * dt;
          if (mode[i] <= CTRL_VELOCITY) v -= CM(d_act_bias[i][2]) * dt;
        }
      }
      M[ri + j] = v; L[ri + j] = v;
    }
    // ---- actuation + passive forces (FD:956-1182)
    const int jn = CM(d_joint[i]);
    const int jt = CM(j_type[jn]), ds = CM(j_dof_start[jn]);
    const int md = mode[i];
    const float v = vel[i], dp = dofs_pos[i];
    const float kp = CM(d_kp[i]), b0 = CM(d_act_bias[i][0]), b1 = CM(d_act_bias[i][1]), b2 = CM(d_act_bias[i][2]);
    const LArr ctrl_pos = C.A(c_h.ctrl_pos), ctrl_vel = C.A(c_h.ctrl_vel), ctrl_force = C.A(c_h.ctrl_force);
    float fo = 0.f;
    if (md == CTRL_FORCE) fo = ctrl_force[i];
    else if (md == CTRL_VELOCITY) fo = -b2 * (ctrl_vel[i] - v);
    else if (md == CTRL_POSITION) {
      if (!(jt == JOINT_FREE && i >= ds + 3)) fo = kp * (ctrl_pos[i] - dp) + b0 + (kp + b1) * dp + b2 * (v - ctrl_vel[i]);
    }
    if (jt == JOINT_FREE && i >= ds + 3 && (mode[ds + 3] == CTRL_POSITION || mode[ds + 4] == CTRL_POSITION || mode[ds + 5] == CTRL_POSITION)) {
      // rotational PD of a free joint acts on the rotation-vector error (FD:1040-1090)
      const f3 xyz(dofs_pos[ds + 3], dofs_pos[ds + 4], dofs_pos[ds + 5]);
      const f3 cxyz(ctrl_pos[ds + 3], ctrl_pos[ds + 4], ctrl_pos[ds + 5]);
      const f3 rv = quat_to_rotvec(quat_compose(xyz_to_quat(cxyz), quat_inv(xyz_to_quat(xyz))), CM(eps));
      fo = kp * comp(rv, i - ds - 3) + b0 + (kp + b1) * dp + b2 * (v - ctrl_vel[i]);
    }
    const float applied = fminf(fmaxf(fo, CM(d_force_range[i][0])), CM(d_force_range[i][1]));
    float passive = -CM(d_damping[i]) * v;
    if (jt != JOINT_FREE && jt != JOINT_FIXED) passive += -dp * CM(d_stiffness[i]);
    const float bias = dot(ld3(cda, i), ld3(cfa, l)) + dot(ld3(cdv, i), ld3(cfv, l));
    const float f = passive - bias + applied;
    qf_applied[i] = applied; force[i] = f; acc_smooth[i] = f;
    if (diag_out) { C.rec[c_h.qf_passive + i] = passive; C.rec[c_h.qf_bias + i] = bias; C.rec[c_h.qf_smooth + i] = f; }
  }
  if (diag_out) {
    for (int i = 0; i < nD; i++)
      for (int j = 0; j < nD; j++) C.rec[c_h.mass_mat + i * nD + j] = M[sym_idx(i, j)];
  }
  lane_mass_factor(C);
  lane_mass_solve(C, acc_smooth);  // FD:1506-1563
  if (diag_out)
    for (int i = 0; i < nD; i++) { C.rec[c_h.acc_smooth + i] = acc_smooth[i]; C.rec[c_h.mass_Dinv + i] = C.A(c_l.Dinv)[i]; }
}

// ---------------------------------------------------------------------------------------------------------------------
// Collision detection (collider.py:684-808), serial per environment
// ---------------------------------------------------------------------------------------------------------------------
LHD __forceinline__ void lane_geom_pose(const LaneCtx& C, int g, f3& gp, f4& gq) {
  const int l = LG(g_link[g]);
  const f3 lp = ld3(C.A(c_l.lpos), l);
  const f4 lq = ld4(C.A(c_l.lquat), l);
  gp = lp + rotate(f3(LG(g_pos[g][0]), LG(g_pos[g][1]), LG(g_pos[g][2])), lq);
  gq = quat_compose(f4(LG(g_quat[g][0]), LG(g_quat[g][1]), LG(g_quat[g][2]), LG(g_quat[g][3])), lq);
}

// geom poses from link poses (FK:828-864) and world AABBs (FK:1297-1319): min[3 nG] then max[3 nG] in the collision scratch
LHD __noinline__ void lane_geoms(const LaneCtx& C) {
  const int nG = CM(nG);
  const LArr amin = C.A(c_l.aabb), amax = C.A(c_l.aabb + 3 * nG);
  const LArr lpos = C.A(c_l.lpos), lquat = C.A(c_l.lquat);
  for (int g = 0; g < nG; g++) {
    const int l = CM(g_link[g]);
    const f3 lp = ld3(lpos, l);
    const f4 lq = ld4(lquat, l);
    const f3 gp = lp + rotate(f3(CM(g_pos[g][0]), CM(g_pos[g][1]), CM(g_pos[g][2])), lq);
    const f4 gq = quat_compose(f4(CM(g_quat[g][0]), CM(g_quat[g][1]), CM(g_quat[g][2]), CM(g_quat[g][3])), lq);
    // identical to the min / max over the 8 transformed rest-pose corners: centre +- |R| half-extent
    const mat3 R = quat_to_R(gq, 0.f);
    const f3 lo(CM(g_aabb_lo[g][0]), CM(g_aabb_lo[g][1]), CM(g_aabb_lo[g][2])), hi(CM(g_aabb_hi[g][0]), CM(g_aabb_hi[g][1]), CM(g_aabb_hi[g][2]));
    const f3 c = (lo + hi) * 0.5f, h = (hi - lo) * 0.5f;
    const f3 wc(R.m[0][0] * c.x + R.m[0][1] * c.y + R.m[0][2] * c.z + gp.x, R.m[1][0] * c.x + R.m[1][1] * c.y + R.m[1][2] * c.z + gp.y,
                R.m[2][0] * c.x + R.m[2][1] * c.y + R.m[2][2] * c.z + gp.z);
    const f3 wh(fabsf(R.m[0][0]) * h.x + fabsf(R.m[0][1]) * h.y + fabsf(R.m[0][2]) * h.z,
                fabsf(R.m[1][0]) * h.x + fabsf(R.m[1][1]) * h.y + fabsf(R.m[1][2]) * h.z,
                fabsf(R.m[2][0]) * h.x + fabsf(R.m[2][1]) * h.y + fabsf(R.m[2][2]) * h.z);
    st3(amin, g, wc - wh); st3(amax, g, wc + wh);
  }
}

// SF:138-180 support point of a mesh-backed geom from its 180x180 direction table
LHD __noinline__ f3 lane_support_mesh(const LaneCtx& C, f3 d, int g) {
  const float PI = 3.14159265358979323846f;
  const float theta = atan2f(d.y, d.x), phi = acosf(d.z);
  const int res = 180;
  const float ii = (theta + PI) / PI / 2.f * res, jj = phi / PI * res;
  float dot_max = -1e20f;
  f3 best;
  const float* tab = CM(support_v) + 3 * (size_t)LG(g_support_start[g]);
#pragma unroll
  for (int i4 = 0; i4 < 4; i4++) {
    int i = (i4 & 1) ? ((int)ceilf(ii)) % res : ((int)floorf(ii)) % res;
    int j;
    if (i4 >> 1) { j = (int)fminf(fmaxf(ceilf(jj), 0.f), (float)(res - 1)); if (j == res - 1) j = res - 2; }
    else { j = (int)fminf(fmaxf(floorf(jj), 0.f), (float)(res - 1)); if (j == 0) j = 1; }
    const float* t = tab + 3 * (i * res + j);
    const f3 v(LDG(t), LDG(t + 1), LDG(t + 2));
    const float dt = dot(v, d);
    if (dt > dot_max) { best = v; dot_max = dt; }
  }
  return best;
}
// MPR:146-176 support_driver
LHD __noinline__ f3 lane_support(const LaneCtx& C, f3 d, int g, f3 pos, f4 quat) {
  const int type = LG(g_type[g]);
  if (type == GEOM_SPHERE) return pos + d * LG(g_data[g][0]);
  if (type == GEOM_BOX) {
    const f3 db = rotate_inv(d, quat);
    const f3 v((db.x < 0.f ? -1.f : 1.f) * LG(g_data[g][0]) * 0.5f, (db.y < 0.f ? -1.f : 1.f) * LG(g_data[g][1]) * 0.5f,
               (db.z < 0.f ? -1.f : 1.f) * LG(g_data[g][2]) * 0.5f);
    return transform(v, pos, quat);
  }
  if (type == GEOM_CAPSULE) {
    const float r = LG(g_data[g][0]), hl = 0.5f * LG(g_data[g][1]);
    const f3 axis = rotate(f3(0.f, 0.f, 1.f), quat);
    const float side = dot(d, axis) < 0.f ? -1.f : 1.f;
    return pos + axis * (hl * side) + d * r;
  }
  if (type == GEOM_ELLIPSOID) {
    const float a = LG(g_data[g][0]), b = LG(g_data[g][1]), c = LG(g_data[g][2]);
    const f3 dl = rotate_inv(d, quat);
    const float n = sqrtf(a * a * dl.x * dl.x + b * b * dl.y * dl.y + c * c * dl.z * dl.z);
    return pos + rotate(f3(a * a * dl.x / n, b * b * dl.y / n, c * c * dl.z / n), quat);
  }
  return transform(lane_support_mesh(C, rotate_inv(d, quat), g), pos, quat);
}

// ---- MPR (mpr.py), simplex in registers ------------------------------------------------------------------------------------
struct LSimplex { f3 v[4], v1[4], v2[4]; };
struct LPairPose { int ga, gb; f3 pa, pb; f4 qa, qb; };
LHD __forceinline__ void lsimplex_set(LSimplex& s, int i, f3 v, f3 v1, f3 v2) {
  if (i == 1) { s.v[1] = v; s.v1[1] = v1; s.v2[1] = v2; }
  else if (i == 2) { s.v[2] = v; s.v1[2] = v1; s.v2[2] = v2; }
  else { s.v[3] = v; s.v1[3] = v1; s.v2[3] = v2; }
}
LHD __forceinline__ void lane_mpr_support(const LaneCtx& C, const LPairPose& P, f3 d, f3& v, f3& v1, f3& v2) {
  v1 = lane_support(C, d, P.ga, P.pa, P.qa);
  v2 = lane_support(C, -d, P.gb, P.pb, P.qb);
  v = v1 - v2;
}
struct LaneSupport {  // support functor of the pair for gsr_epa.cuh
  const LaneCtx& C;
  const LPairPose& P;
  LHD __forceinline__ void operator()(f3 d, f3& v, f3& v1, f3& v2) const { lane_mpr_support(C, P, d, v, v1, v2); }
};
LHD __forceinline__ f3 lportal_dir(const LSimplex& s) { return normalized(cross(s.v[2] - s.v[1], s.v[3] - s.v[1])); }
LHD __forceinline__ bool lreach_tolerance(const LSimplex& s, f3 v, f3 dir) {  // MPR:134-143
  const float dv4 = dot(v, dir);
  const float dot1 = fminf(fminf(dv4 - dot(s.v[1], dir), dv4 - dot(s.v[2], dir)), dv4 - dot(s.v[3], dir));
  return dot1 < CM(ccd_tolerance) + CM(ccd_eps) * fmaxf(1.0f, dot1);
}
LHD __forceinline__ void lexpand_portal(LSimplex& s, f3 v, f3 v1, f3 v2) {  // MPR:426-442
  const f3 v4v0 = cross(v, s.v[0]);
  int is;
  if (dot(s.v[1], v4v0) > 0.f) is = dot(s.v[2], v4v0) > 0.f ? 1 : 3;
  else is = dot(s.v[3], v4v0) > 0.f ? 2 : 1;
  lsimplex_set(s, is, v, v1, v2);
}
// MPR:445-598
LHD __noinline__ int lane_discover_portal(const LaneCtx& C, LSimplex& s, const LPairPose& P, f3 ca, f3 cb) {
  const float CE = CM(ccd_eps);
  s.v1[0] = ca; s.v2[0] = cb; s.v[0] = ca - cb;
  if (fabsf(s.v[0].x) < CE && fabsf(s.v[0].y) < CE && fabsf(s.v[0].z) < CE) s.v[0].x += 10.0f * CE;
  f3 dir = -normalized(s.v[0]);
  f3 v, v1, v2;
  lane_mpr_support(C, P, dir, v, v1, v2);
  s.v[1] = v; s.v1[1] = v1; s.v2[1] = v2;
  if (dot(v, dir) < CE) return -1;
  dir = cross(s.v[0], s.v[1]);
  if (dot(dir, dir) < CE) {
    if (fabsf(s.v[1].x) < CE && fabsf(s.v[1].y) < CE && fabsf(s.v[1].z) < CE) return 1;
    return 2;
  }
  dir = normalized(dir);
  lane_mpr_support(C, P, dir, v, v1, v2);
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
    lane_mpr_support(C, P, dir, v, v1, v2);
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
// MPR:601-819 func_mpr_contact (centre guess + portal discovery / refinement + penetration)
LHD __noinline__ void lane_mpr_contact(const LaneCtx& C, const LPairPose& P, f3 normal_ws, bool& is_col, f3& normal, float& pen, f3& pos,
                                       LSimplex& s, bool& tetra) {
  const float CE = CM(ccd_eps);
  f3 ca = transform(f3(LG(g_center[P.ga][0]), LG(g_center[P.ga][1]), LG(g_center[P.ga][2])), P.pa, P.qa);
  f3 cb = transform(f3(LG(g_center[P.gb][0]), LG(g_center[P.gb][1]), LG(g_center[P.gb][2])), P.pb, P.qb);
  if (fabsf(normal_ws.x) > CE || fabsf(normal_ws.y) > CE || fabsf(normal_ws.z) > CE) {
    const f3 loa(LG(g_aabb_lo[P.ga][0]), LG(g_aabb_lo[P.ga][1]), LG(g_aabb_lo[P.ga][2])), hia(LG(g_aabb_hi[P.ga][0]), LG(g_aabb_hi[P.ga][1]), LG(g_aabb_hi[P.ga][2]));
    const f3 lob(LG(g_aabb_lo[P.gb][0]), LG(g_aabb_lo[P.gb][1]), LG(g_aabb_lo[P.gb][2])), hib(LG(g_aabb_hi[P.gb][0]), LG(g_aabb_hi[P.gb][1]), LG(g_aabb_hi[P.gb][2]));
    ca = transform((hia + loa) * 0.5f, P.pa, P.qa);
    cb = transform((hib + lob) * 0.5f, P.pb, P.qb);
    const f3 delta = ca - cb;
    if (norm(cross(normal_ws, normalized(delta))) > 0.01f) {
      const f3 offset = normal_ws * dot(delta, normal_ws) - delta;
      const float on = norm(offset);
      if (on > CM(eps)) {
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
  int res = lane_discover_portal(C, s, P, ca, cb);
  is_col = false; pos = f3(); normal = f3(); pen = 0.f;
  tetra = false;  // true when s ends as the interior point + final portal, a tetrahedron around the origin (EPA seed)
  if (res == 1) { is_col = true; normal = -normalized(s.v[0]); pos = (s.v1[1] + s.v2[1]) * 0.5f; }
  else if (res == 2) { is_col = true; pen = norm(s.v[1]); normal = -normalized(s.v[1]); pos = (s.v1[1] + s.v2[1]) * 0.5f; }
  else if (res == 0) {
    while (true) {  // MPR:232-278 refine portal
      const f3 dir = lportal_dir(s);
      if (dot(s.v[1], dir) > -CE) { res = 0; break; }
      f3 v, v1, v2;
      lane_mpr_support(C, P, dir, v, v1, v2);
      if (!(dot(v, dir) > -CE) || lreach_tolerance(s, v, dir)) { res = -1; break; }
      lexpand_portal(s, v, v1, v2);
    }
    if (res >= 0) {
      int it = 0;
      while (true) {  // MPR:338-423 find penetration
        const f3 dir = lportal_dir(s);
        f3 v, v1, v2;
        lane_mpr_support(C, P, dir, v, v1, v2);
        if (lreach_tolerance(s, v, dir) || it > CM(ccd_iterations)) {
          pen = dot(dir, s.v[1]);
          normal = -dir;
          is_col = true;
          // MPR:281-316 find_pos: barycentric weights of the origin's projection on the portal
          const float b1 = dot(cross(s.v[2], s.v[3]), dir), b2 = dot(cross(s.v[3], s.v[1]), dir), b3 = dot(cross(s.v[1], s.v[2]), dir);
          const float sum = b1 + b2 + b3;
          const f3 p1 = s.v1[1] * b1 + s.v1[2] * b2 + s.v1[3] * b3, p2 = s.v2[1] * b1 + s.v2[2] * b2 + s.v2[3] * b3;
          pos = (p1 + p2) * (0.5f / sum);
          tetra = true;
          break;
        }
        lexpand_portal(s, v, v1, v2);
        it++;
      }
    }
  }
}

struct LContacts {
  int n;
  float pen[5];
  f3 pos[5], nrm[5];
  LHD __forceinline__ void add(f3 normal, f3 p, float pe) {
    if (n < 5) { nrm[n] = normal; pos[n] = p; pen[n] = pe; n++; }
  }
};
// CT:317-336
LHD __forceinline__ float lane_tolerance(const LaneCtx& C, int ga, int gb) {
  float size = norm(f3(LG(g_aabb_hi[gb][0]) - LG(g_aabb_lo[gb][0]), LG(g_aabb_hi[gb][1]) - LG(g_aabb_lo[gb][1]), LG(g_aabb_hi[gb][2]) - LG(g_aabb_lo[gb][2])));
  if (LG(g_type[ga]) != GEOM_PLANE)
    size = fminf(norm(f3(LG(g_aabb_hi[ga][0]) - LG(g_aabb_lo[ga][0]), LG(g_aabb_hi[ga][1]) - LG(g_aabb_lo[ga][1]), LG(g_aabb_hi[ga][2]) - LG(g_aabb_lo[ga][2]))), size);
  return 0.5f * CM(mc_tolerance) * size;
}
// CT:415-436
LHD __forceinline__ void lane_rotate_frame(f3 pos, f4 quat, f3 cpos, f4 qrot, f3& npos, f4& nquat) {
  nquat = quat_compose(quat, qrot);
  const f3 rel = cpos - pos;
  npos = pos - (rotate(rel, qrot) - rel);
}
// NP:519-970 func_convex_convex_contact for one broad-phase pair (type[ga] <= type[gb], not plane-box)
LHD __noinline__ void lane_convex_convex(const LaneCtx& C, LContacts& out, int ga, int gb, int i_pair) {
  const float EPS = CM(eps);
  const int ta = LG(g_type[ga]), tb = LG(g_type[gb]);
  const bool multi = CM(enable_multi_contact) && ta != GEOM_SPHERE && ta != GEOM_ELLIPSOID && tb != GEOM_SPHERE && tb != GEOM_ELLIPSOID;
  const float tolerance = lane_tolerance(C, ga, gb);
  f3 ga_pos0, gb_pos0; f4 ga_quat0, gb_quat0;
  lane_geom_pose(C, ga, ga_pos0, ga_quat0);
  lane_geom_pose(C, gb, gb_pos0, gb_quat0);
  LPairPose P;
  P.ga = ga; P.gb = gb; P.pa = ga_pos0; P.pb = gb_pos0; P.qa = ga_quat0; P.qb = gb_quat0;
  bool is_col_0 = false, is_col = false;
  float pen = 0.f;
  f3 normal_0, cpos_0, normal, cpos, axis_0, axis_1;
  f4 qrot(0.f, 0.f, 0.f, 0.f);
  const bool use_cache = ta != GEOM_PLANE;
  int* cvalid = reinterpret_cast<int*>(C.cache) + c_h.cache_valid;
  for (int i_det = 0; i_det < 5; i_det++) {
    if (i_det > 0 && !(multi && is_col_0)) break;
    if (i_det > 0) {
      const f3 axis = axis_0 * (float)(2 * (i_det % 2) - 1) + axis_1 * (float)(1 - 2 * ((i_det / 2) % 2));
      qrot = rotvec_to_quat(axis * CM(mc_perturbation), EPS);
      lane_rotate_frame(ga_pos0, ga_quat0, cpos_0, qrot, P.pa, P.qa);
      lane_rotate_frame(gb_pos0, gb_quat0, cpos_0, quat_inv(qrot), P.pb, P.qb);
    }
    if (ta == GEOM_CAPSULE && tb == GEOM_CAPSULE) {  // NP:637-648 analytic capsule pairs
      capsule_capsule_contact(P.pa, P.qa, LG(g_data[ga][0]), 0.5f * LG(g_data[ga][1]), P.pb, P.qb, LG(g_data[gb][0]), 0.5f * LG(g_data[gb][1]), EPS,
                              is_col, normal, cpos, pen);
    } else if (ta == GEOM_SPHERE && tb == GEOM_CAPSULE) {  // NP:649-661 (types are sorted: A is the sphere)
      sphere_capsule_contact(P.pa, LG(g_data[ga][0]), P.pb, P.qb, LG(g_data[gb][0]), 0.5f * LG(g_data[gb][1]), EPS, is_col, normal, cpos, pen);
    } else if (ta == GEOM_PLANE) {
      normal = -normalized(rotate(f3(LG(g_data[ga][0]), LG(g_data[ga][1]), LG(g_data[ga][2])), P.qa));
      const f3 v1 = lane_support(C, normal, gb, P.pb, P.qb);
      pen = dot(normal, v1 - P.pa);
      cpos = v1 - normal * (0.5f * pen);
      is_col = pen > 0.f;
    } else {
      f3 normal_ws;
      const bool valid = (cvalid[i_pair >> 5] >> (i_pair & 31)) & 1;
      if (valid) normal_ws = gld3(C.cache, i_pair);
      const bool guess = fabsf(normal_ws.x) > EPS || fabsf(normal_ws.y) > EPS || fabsf(normal_ws.z) > EPS;
      LSimplex sx;
      bool tetra = false, guess_used = guess;
      lane_mpr_contact(C, P, normal_ws, is_col, normal, pen, cpos, sx, tetra);
      if (i_det == 0 && !is_col && guess) { lane_mpr_contact(C, P, f3(), is_col, normal, pen, cpos, sx, tetra); guess_used = false; }
      // NP:726-733: MPR's answer is replaced by the exact one when the overlap is large or no direction was cached
      if (CM(gjk_fallback) && is_col && tetra && pen > tolerance && (!guess_used || CM(mc_tolerance) * pen >= 0.25f * tolerance)) {
        const LaneSupport sup{C, P};
        f3 n2, c2;
        float p2 = 0.f;
        if (epa_penetration(sx.v, sx.v1, sx.v2, sup, n2, p2, c2)) { normal = n2; pen = p2; cpos = c2; is_col = p2 > 0.f; }
      }
    }
    if (i_det == 0) {
      is_col_0 = is_col; normal_0 = normal; cpos_0 = cpos;
      if (is_col_0) {
        out.add(normal, cpos, pen);
        if (multi) {
          // CT:353-412 func_contact_orthogonals: perturbation axes from the principal inertia frame of the smaller geom
          int g = gb;
          if (ta != GEOM_PLANE) {
            const float va = LG(g_aabb_hi[ga][0]) * LG(g_aabb_hi[ga][1]) * LG(g_aabb_hi[ga][2]);
            const float vb = LG(g_aabb_hi[gb][0]) * LG(g_aabb_hi[gb][1]) * LG(g_aabb_hi[gb][2]);
            g = (va < vb) ? ga : gb;
          }
          const int l = LG(g_link[g]);
          const mat3 rot = quat_to_R(quat_compose(f4(LG(l_iquat[l][0]), LG(l_iquat[l][1]), LG(l_iquat[l][2]), LG(l_iquat[l][3])), ld4(C.A(c_l.lquat), l)), EPS);
          int axis_idx = 0;
          float amax = 0.f;
          for (int i = 0; i < 3; i++) { const float a = fabsf(dot(mat_col(rot, i), normal)); if (a > amax) { amax = a; axis_idx = i; } }
          axis_idx = (axis_idx + 1) % 3;
          axis_0 = mat_col(rot, axis_idx);
          axis_0 = normalized(axis_0 - normal * dot(normal, axis_0));
          axis_1 = cross(normal, axis_0);
        }
      }
      if (use_cache) {
        if (is_col_0) { gst3(C.cache, i_pair, normal); lane_atomic_or(&cvalid[i_pair >> 5], 1 << (i_pair & 31)); }
        else lane_atomic_and(&cvalid[i_pair >> 5], ~(1 << (i_pair & 31)));
      }
    } else if (is_col) {
      const f3 cpa = rotate((cpos - normal * (0.5f * pen)) - cpos_0, quat_inv(qrot)) + cpos_0;
      const f3 cpb = rotate((cpos + normal * (0.5f * pen)) - cpos_0, qrot) + cpos_0;
      cpos = (cpa + cpb) * 0.5f;
      f3 twist = cross(normal, normal_0);
      const float mcp = CM(mc_perturbation);
      twist = f3(fminf(fmaxf(twist.x, -mcp), mcp), fminf(fmaxf(twist.y, -mcp), mcp), fminf(fmaxf(twist.z, -mcp), mcp));
      normal = normal + cross(twist, normal);
      pen = dot(normal, cpb - cpa);
      bool repeated = false;
      for (int c = 0; c < out.n; c++)
        if (norm(cpos - out.pos[c]) < tolerance) repeated = true;
      if (!repeated && pen > -tolerance) out.add(normal, cpos, fmaxf(pen, 0.f));
    }
  }
}
// BOX:24-96 func_plane_box_contact
LHD __noinline__ void lane_plane_box(const LaneCtx& C, LContacts& out, int ga, int gb) {
  f3 pa, pb; f4 qa, qb;
  lane_geom_pose(C, ga, pa, qa);
  lane_geom_pose(C, gb, pb, qb);
  const f3 normal = -normalized(rotate(f3(LG(g_data[ga][0]), LG(g_data[ga][1]), LG(g_data[ga][2])), qa));
  const f3 v1 = lane_support(C, normal, gb, pb, qb);
  float pen = dot(normal, v1 - pa);
  if (pen > 0.f) {
    f3 cpos = v1 - normal * (0.5f * pen);
    out.add(normal, cpos, pen);
    if (CM(enable_multi_contact)) {
      int n_con = 1;
      const f3 cpos_0 = cpos;
      const float tolerance = lane_tolerance(C, ga, gb);
      const int vs = LG(g_vert_start[gb]), ve = LG(g_vert_end[gb]);
      for (int iv = vs; iv < ve; iv++) {
        if (n_con < CM(n_contacts_per_pair)) {
          const float* vp = CM(verts) + 3 * (size_t)iv;
          const f3 pc = transform(f3(LDG(vp), LDG(vp + 1), LDG(vp + 2)), pb, qb);
          pen = dot(normal, pc - pa);
          if (pen > 0.f) {
            cpos = pc - normal * (0.5f * pen);
            if (norm(cpos - cpos_0) > tolerance) { out.add(normal, cpos, pen); n_con++; }
          }
        }
      }
    }
  }
}

// narrow phase of one candidate pair: appends its contacts (unordered) to the staging records in the SPILL record
LHD __noinline__ int lane_narrow_pair(const LaneCtx& C, int p, int n_staged) {
  LContacts out;
  out.n = 0;
  const int ga0 = LDG(CM(pairs) + 2 * p), gb0 = LDG(CM(pairs) + 2 * p + 1);
  int ga = ga0, gb = gb0;
  if (LG(g_type[ga0]) > LG(g_type[gb0])) { ga = gb0; gb = ga0; }
  const int plane_box = LG(g_type[ga]) == GEOM_PLANE && LG(g_type[gb]) == GEOM_BOX;
  if (plane_box) lane_plane_box(C, out, ga, gb);
  else lane_convex_convex(C, out, ga, gb, p);
  const int cap = CM(max_contact_pairs);
  if (out.n > 0) {
    const float mu = fmaxf(fmaxf(LG(g_friction[ga]) * C.rec[c_h.friction_ratio + ga], LG(g_friction[gb]) * C.rec[c_h.friction_ratio + gb]), 1e-2f);
    const int la = LG(g_link[ga]), lb = LG(g_link[gb]);
    for (int k = 0; k < 5; k++) {
      if (k < out.n && n_staged + k < cap) {
        float* r = C.spill + (size_t)(n_staged + k) * GSR_CONTACT_WORDS;
        int* ri = reinterpret_cast<int*>(r);
        ri[CT_GA] = ga; ri[CT_GB] = gb; ri[CT_LA] = la; ri[CT_LB] = lb;
        r[CT_PEN] = out.pen[k]; r[CT_MU] = mu;
        r[CT_POS] = out.pos[k].x; r[CT_POS + 1] = out.pos[k].y; r[CT_POS + 2] = out.pos[k].z;
        r[CT_NRM] = out.nrm[k].x; r[CT_NRM + 1] = out.nrm[k].y; r[CT_NRM + 2] = out.nrm[k].z;
        ri[CT_FRC] = (plane_box << 24) | (p << 3) | k;  // canonical detection order (convex pairs first, pair, index)
        r[CT_FRC + 1] = 0.f; r[CT_FRC + 2] = 0.f;
        r[CT_KEY] = out.pos[0].x;                        // CT:455-470: sort key = x of the pair's first contact
      }
    }
  }
  int n_new = n_staged + out.n;
  if (n_new > cap) {
    C.I(c_h.errno_)[0] |= ERR_CONTACT_OVERFLOW;
    n_new = cap;
  }
  return n_new;
}

LHD __noinline__ void lane_collide(const LaneCtx& C) {
  const int nP = CM(nP), nG = CM(nG);
  const LArr amin = C.A(c_l.aabb), amax = C.A(c_l.aabb + 3 * nG);
  const LIArr queue = C.I(c_l.queue);
  const int n_prev = C.I(c_h.n_contacts)[0];
  int* cvalid = reinterpret_cast<int*>(C.cache) + c_h.cache_valid;
  // ---- broad phase (BP:400-460): all valid pairs, in pair order
  int qn = 0, n_broad = 0;
  for (int p = 0; p < nP; p++) {
    const int ga = LDG(CM(pairs) + 2 * p), gb = LDG(CM(pairs) + 2 * p + 1);
    const f3 a0 = ld3(amin, ga), a1 = ld3(amax, ga), b0 = ld3(amin, gb), b1 = ld3(amax, gb);
    const bool sep = a1.x <= b0.x || a0.x >= b1.x || a1.y <= b0.y || a0.y >= b1.y || a1.z <= b0.z || a0.z >= b1.z;
    if (sep) {
      if (CM(g_type[ga]) != GEOM_PLANE && CM(g_type[gb]) != GEOM_PLANE) {  // BP:120-133: separated pairs lose their warm start
        if ((cvalid[p >> 5] >> (p & 31)) & 1) lane_atomic_and(&cvalid[p >> 5], ~(1 << (p & 31)));
      }
    } else {
      if (n_broad < CM(max_broad_pairs)) { queue[qn++] = p; n_broad++; }
      else C.I(c_h.errno_)[0] |= ERR_BROAD_OVERFLOW;
    }
  }
  C.I(c_h.n_broad)[0] = n_broad;
  // ---- narrow phase, pair by pair
  int n = 0;
  for (int q = 0; q < qn; q++) n = lane_narrow_pair(C, queue[q], n);
  // ---- canonical order (CT:439-538) by ranking, then the final records (global) and the on-chip copy
  const bool sorted = CM(has_cc) && CM(sort_contacts);
  const int NCC = c_l.NCC;
  const LArr ct_s = C.A(c_l.ct);
  for (int c = 0; c < n; c++) {
    const float* r = C.spill + (size_t)c * GSR_CONTACT_WORDS;
    const float key = sorted ? r[CT_KEY] : 0.f;
    const int ord = reinterpret_cast<const int*>(r)[CT_FRC];
    int rank = 0;
    for (int o = 0; o < n; o++) {
      const float* ro = C.spill + (size_t)o * GSR_CONTACT_WORDS;
      const float ko = sorted ? ro[CT_KEY] : 0.f;
      const int oo = reinterpret_cast<const int*>(ro)[CT_FRC];
      rank += (ko < key || (ko == key && oo < ord)) ? 1 : 0;
    }
    float* dst = C.contacts + (size_t)rank * GSR_CONTACT_WORDS;
    for (int k = 0; k < GSR_CONTACT_WORDS; k++) {
      const float v = (k >= CT_FRC && k < CT_FRC + 3) ? 0.f : r[k];
      dst[k] = v;
      if (rank < NCC) ct_s[rank * GSR_CONTACT_WORDS + k] = v;
    }
  }
  // BP:74-139: slots that held a contact last step and do not any more read as cleared (-1 / 0)
  for (int c = n; c < n_prev; c++) {
    float* dst = C.contacts + (size_t)c * GSR_CONTACT_WORDS;
    for (int k = 0; k < 4; k++) reinterpret_cast<int*>(dst)[k] = -1;
    for (int k = 4; k < GSR_CONTACT_WORDS; k++) dst[k] = 0.f;
  }
  C.I(c_h.n_contacts)[0] = n;
}

// ---------------------------------------------------------------------------------------------------------------------
// Constraint rows (CS:579-1286) and the Newton / CG solver (CS:2033-3261, monolith body), serial per environment
// ---------------------------------------------------------------------------------------------------------------------
struct LRows {
  GArr J; int JS;  // J[r * JS + d]
  GArr aref, D, R, floss, Jaref, jv, force, active;
  GArr ct;         // contact records (16 words each)
  int n_con, ne, nef;
};
struct LSolver {
  float cost, prev_cost, gauss;
  float qg0, qg1, qg2, es0, es1, es2;
  int ls_it;
};
LHD __forceinline__ int& gi(GArr a, int i) { return reinterpret_cast<int&>(a[i]); }

// CS:984-1173 func_equality_weld: 3 position rows (anchor1 = data[3:6] in body 1, anchor2 = data[0:3] in body 2) + 3 orientation
// rows (relpose = data[6:10], torquescale = data[10]); translational / rotational invweight components
LHD __noinline__ void lane_weld_rows(const LaneCtx& C, const LRows& R, int row, int l1, int l2, const float* dat, const float* sol) {
  const int nD = CM(nD), JS = R.JS;
  const float EPS = CM(eps);
  const GArr J = R.J;
  const LArr vel = C.A(c_h.vel);
  const LArr cda = C.A(c_l.cdof), cdv = C.A(c_l.cdof + 3 * nD);
  const LArr lpos = C.A(c_l.lpos), lquat = C.A(c_l.lquat), rcom = C.A(c_l.rcom);
  const f4 q1 = ld4(lquat, l1), q2 = ld4(lquat, l2);
  const f3 g1 = transform(f3(dat[3], dat[4], dat[5]), ld3(lpos, l1), q1);
  const f3 g2 = transform(f3(dat[0], dat[1], dat[2]), ld3(lpos, l2), q2);
  const f3 t1 = g1 - ld3(rcom, LG(l_rootslot[l1])), t2 = g2 - ld3(rcom, LG(l_rootslot[l2]));
  const float ts = dat[10];
  const f4 q = quat_mul(q1, f4(dat[6], dat[7], dat[8], dat[9]));
  const f4 iq2 = quat_inv(q2);
  const f4 eq = quat_mul(iq2, q);
  const f3 pos_error = g1 - g2, rot_error = f3(eq.x, eq.y, eq.z) * ts;
  const unsigned long long m1 = LG(l_dofmask[l1]), m2 = LG(l_dofmask[l2]);
  float jq[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int d = 0; d < nD; d++) {
    f3 v, jr;
    const f3 ca = ld3(cda, d);
    if ((m1 >> d) & 1ull) { v += ld3(cdv, d) - cross(t1, ca); jr += ca; }
    if ((m2 >> d) & 1ull) { v -= ld3(cdv, d) - cross(t2, ca); jr -= ca; }
    const f4 q3 = quat_mul(quat_mul_axis(iq2, jr), q);
    const float jv[6] = {v.x, v.y, v.z, 0.5f * q3.x * ts, 0.5f * q3.y * ts, 0.5f * q3.z * ts};
    const float w = vel[d];
#pragma unroll
    for (int k = 0; k < 6; k++) { J[(row + k) * JS + d] = jv[k]; jq[k] += jv[k] * w; }
  }
  const float pos_imp = sqrtf(dot(pos_error, pos_error) + dot(rot_error, rot_error));
  for (int k = 0; k < 6; k++) {
    const float invweight = k < 3 ? LG(l_invweight[l1]) + LG(l_invweight[l2]) : LG(l_invweight_rot[l1]) + LG(l_invweight_rot[l2]);
    float imp, aref;
    imp_aref(sol, -pos_imp, jq[k], k < 3 ? comp(pos_error, k) : comp(rot_error, k - 3), imp, aref);
    const float diag = fmaxf(invweight * (1.0f - imp) / imp, EPS);
    R.R[row + k] = diag; R.aref[row + k] = aref; R.D[row + k] = 1.0f / diag; R.floss[row + k] = 0.f;
  }
}

LHD __noinline__ void lane_build_rows(const LaneCtx& C, LRows& R) {
  const int nD = CM(nD), nJ = CM(nJ);
  const float EPS = CM(eps);
  const LArr qpos = C.A(c_h.qpos), vel = C.A(c_h.vel);
  const LArr cda = C.A(c_l.cdof), cdv = C.A(c_l.cdof + 3 * nD);
  const LArr lpos = C.A(c_l.lpos), lquat = C.A(c_l.lquat), rcom = C.A(c_l.rcom);
  const int n_contacts = CM(enable_collision) ? C.I(c_h.n_contacts)[0] : 0;
  int n_eq = 0;
  for (int i = 0; i < CM(nEq); i++) n_eq += CM(eq_type[i]) == EQ_JOINT ? 1 : (CM(eq_type[i]) == EQ_CONNECT ? 3 : (CM(eq_type[i]) == EQ_WELD ? 6 : 0));
  int n_dyn = C.I(c_h.n_dyn_eq)[0];
  n_dyn = n_dyn < 0 ? 0 : (n_dyn > GSR_MAX_DYN_EQ ? GSR_MAX_DYN_EQ : n_dyn);
  n_eq += 6 * n_dyn;
  int n_floss = 0;
  for (int d = 0; d < nD; d++) n_floss += CM(d_frictionloss[d]) > EPS ? 1 : 0;
  int n_lim = 0;
  if (CM(enable_joint_limit)) {
    for (int j = 0; j < nJ; j++) {
      const int type = CM(j_type[j]);
      if (type == JOINT_REVOLUTE || type == JOINT_PRISMATIC) {
        const int q = CM(j_q_start[j]), d = CM(j_dof_start[j]);
        const float qv = qpos[q];
        if (fminf(qv - CM(d_limit[d][0]), CM(d_limit[d][1]) - qv) < 0.f) n_lim++;
      }
    }
  }
  const int lim_base = n_eq + n_floss + 4 * n_contacts;
  const int n_con = lim_base + n_lim;
  R.n_con = n_con; R.ne = n_eq; R.nef = n_eq + n_floss;
  // ---- storage: workspace when the rows fit, the SPILL record otherwise
  const int NCS = c_l.NCS, lenC = CM(len_constraints);
  R.JS = nD;
  if (n_con <= NCS) {
    const int es = GSR_LS;
    R.J = GArr{C.S + c_l.J * es, es};
    R.aref = GArr{C.S + c_l.r_aref * es, es}; R.D = GArr{C.S + c_l.r_D * es, es}; R.R = GArr{C.S + c_l.r_R * es, es};
    R.floss = GArr{C.S + c_l.r_floss * es, es}; R.Jaref = GArr{C.S + c_l.r_Jaref * es, es}; R.jv = GArr{C.S + c_l.r_jv * es, es};
    R.force = GArr{C.S + c_l.r_force * es, es}; R.active = GArr{C.S + c_l.r_active * es, es};
  } else {
    float* sp = C.spill;
    R.J = GArr{sp, 1}; sp += (size_t)lenC * (nD | 1);
    R.aref = GArr{sp, 1}; sp += lenC; R.D = GArr{sp, 1}; sp += lenC; R.R = GArr{sp, 1}; sp += lenC; R.floss = GArr{sp, 1}; sp += lenC;
    R.Jaref = GArr{sp, 1}; sp += lenC; R.jv = GArr{sp, 1}; sp += lenC; R.force = GArr{sp, 1}; sp += lenC; R.active = GArr{sp, 1};
  }
  R.ct = (n_contacts <= c_l.NCC) ? GArr{C.S + c_l.ct * GSR_LS, GSR_LS} : GArr{C.contacts, 1};
  if (n_con == 0) return;
  const GArr J = R.J;
  const int JS = R.JS;
  // ---- zero the non-contact rows
  for (int r0 = 0; r0 < R.nef + n_lim; r0++) {
    const int r = r0 < R.nef ? r0 : lim_base + (r0 - R.nef);
    for (int d = 0; d < nD; d++) J[r * JS + d] = 0.f;
  }
  // ---- equalities (CS:880-937)
  int row = 0;
  for (int i = 0; i < CM(nEq); i++) {
    if (CM(eq_type[i]) == EQ_JOINT) {  // CS:800-877
      const int j1 = CM(eq_obj1[i]), j2 = CM(eq_obj2[i]);
      const int q1 = CM(j_q_start[j1]), q2 = CM(j_q_start[j2]), d1 = CM(j_dof_start[j1]), d2 = CM(j_dof_start[j2]);
      const float diff = qpos[q2] - CM(qpos0[q2]);
      float pos = qpos[q1] - CM(qpos0[q1]), deriv = 0.f, dp = 1.f;
      for (int k = 0; k < 5; k++) {
        pos -= dp * CM(eq_data[i][k]);
        if (k < 4) deriv += CM(eq_data[i][k + 1]) * dp * (float)(k + 1);
        dp *= diff;
      }
      J[row * JS + d1] = 1.0f;
      J[row * JS + d2] = -deriv;
      const float jac_qvel = vel[d1] + (-deriv) * vel[d2];
      const float invweight = CM(d_invweight[d1]) + CM(d_invweight[d2]);
      float imp, aref;
      imp_aref(CM(eq_sol[i]), -fabsf(pos), jac_qvel, pos, imp, aref);
      const float diag = fmaxf(invweight * (1.0f - imp) / imp, EPS);
      R.R[row] = diag; R.aref[row] = aref; R.D[row] = 1.0f / diag; R.floss[row] = 0.f;
      row += 1;
    } else if (CM(eq_type[i]) == EQ_CONNECT) {  // CS:688-797
      const int l1 = CM(eq_obj1[i]), l2 = CM(eq_obj2[i]);
      const f3 g1 = transform(f3(CM(eq_data[i][0]), CM(eq_data[i][1]), CM(eq_data[i][2])), ld3(lpos, l1), ld4(lquat, l1));
      const f3 g2 = transform(f3(CM(eq_data[i][3]), CM(eq_data[i][4]), CM(eq_data[i][5])), ld3(lpos, l2), ld4(lquat, l2));
      const f3 t1 = g1 - ld3(rcom, CM(l_rootslot[l1])), t2 = g2 - ld3(rcom, CM(l_rootslot[l2]));
      const unsigned long long m1 = CM(l_dofmask[l1]), m2 = CM(l_dofmask[l2]);
      float jq[3] = {0.f, 0.f, 0.f};
      for (int d = 0; d < nD; d++) {
        f3 v;
        if ((m1 >> d) & 1ull) v += ld3(cdv, d) - cross(t1, ld3(cda, d));
        if ((m2 >> d) & 1ull) v -= ld3(cdv, d) - cross(t2, ld3(cda, d));
        J[row * JS + d] = v.x; J[(row + 1) * JS + d] = v.y; J[(row + 2) * JS + d] = v.z;
        const float w = vel[d];
        jq[0] += v.x * w; jq[1] += v.y * w; jq[2] += v.z * w;
      }
      const f3 pd = g1 - g2;
      const float invweight = CM(l_invweight[l1]) + CM(l_invweight[l2]);
      for (int k = 0; k < 3; k++) {
        float imp, aref;
        imp_aref(CM(eq_sol[i]), -norm(pd), jq[k], comp(pd, k), imp, aref);
        const float diag = fmaxf(invweight * (1.0f - imp) / imp, EPS);
        R.R[row + k] = diag; R.aref[row + k] = aref; R.D[row + k] = 1.0f / diag; R.floss[row + k] = 0.f;
      }
      row += 3;
    } else if (CM(eq_type[i]) == EQ_WELD) {
      lane_weld_rows(C, R, row, CM(eq_obj1[i]), CM(eq_obj2[i]), CM(eq_data[i]), CM(eq_sol[i]));
      row += 6;
    }
  }
  // dynamic welds of this environment (add_weld_constraint), after the static equalities like in the reference's table
  for (int k = 0; k < n_dyn; k++) {
    const float* rec_eq = C.rec + c_h.dyn_eq + k * GSR_DYN_EQ_WORDS;
    lane_weld_rows(C, R, row, lane_f2i(rec_eq[0]), lane_f2i(rec_eq[1]), rec_eq + 2, rec_eq + 13);
    row += 6;
  }
  // ---- frictionloss (CS:1236-1286), static rows in dof order
  if (n_floss > 0) {
    int r = n_eq;
    for (int d = 0; d < nD; d++)
      if (CM(d_frictionloss[d]) > EPS) {
        float imp, aref;
        imp_aref(CM(j_sol[CMI(d_joint[d])]), 0.f, vel[d], 0.f, imp, aref);
        const float diag = fmaxf(CM(d_invweight[d]) * (1.0f - imp) / imp, EPS);
        R.R[r] = diag; R.aref[r] = aref; R.D[r] = 1.0f / diag; R.floss[r] = CM(d_frictionloss[d]);
        J[r * JS + d] = 1.0f;
        r++;
      }
  }
  // ---- contacts (CS:579-685): pyramidal friction, 4 rows per contact
  const int c_base = R.nef;
  for (int ic = 0; ic < n_contacts; ic++) {
    const int o = ic * GSR_CONTACT_WORDS;
    const int la = lane_f2i(R.ct[o + CT_LA]), lb = lane_f2i(R.ct[o + CT_LB]);
    const int ga = lane_f2i(R.ct[o + CT_GA]), gb = lane_f2i(R.ct[o + CT_GB]);
    const f3 cpos(R.ct[o + CT_POS], R.ct[o + CT_POS + 1], R.ct[o + CT_POS + 2]), cn(R.ct[o + CT_NRM], R.ct[o + CT_NRM + 1], R.ct[o + CT_NRM + 2]);
    const float mu = R.ct[o + CT_MU], pen = R.ct[o + CT_PEN];
    f3 d1, d2;
    orthogonals(cn, d1, d2);
    const f3 n0 = d1 * (-mu) - cn, n1 = d1 * mu - cn, n2 = d2 * (-mu) - cn, n3 = d2 * mu - cn;
    const unsigned long long ma = LG(l_dofmask[la]), mb = lb > -1 ? LG(l_dofmask[lb]) : 0ull;
    const f3 ra = cpos - ld3(rcom, LG(l_rootslot[la]));
    f3 rb;
    if (lb > -1) rb = cpos - ld3(rcom, LG(l_rootslot[lb]));
    const int c0 = c_base + 4 * ic;
    float jq0 = 0.f, jq1 = 0.f, jq2 = 0.f, jq3 = 0.f;
    for (int d = 0; d < nD; d++) {
      f3 diff;
      const f3 cv = ld3(cdv, d), ca = ld3(cda, d);
      if ((ma >> d) & 1ull) diff -= cv - cross(ra, ca);
      if ((mb >> d) & 1ull) diff += cv - cross(rb, ca);
      const float j0 = dot(diff, n0), j1 = dot(diff, n1), j2 = dot(diff, n2), j3 = dot(diff, n3);
      J[c0 * JS + d] = j0; J[(c0 + 1) * JS + d] = j1; J[(c0 + 2) * JS + d] = j2; J[(c0 + 3) * JS + d] = j3;
      const float w = vel[d];
      jq0 += j0 * w; jq1 += j1 * w; jq2 += j2 * w; jq3 += j3 * w;
    }
    float sol[7];
#pragma unroll
    for (int k = 0; k < 7; k++) sol[k] = 0.5f * (LG(g_sol[ga][k]) + LG(g_sol[gb][k]));
    float invweight = LG(l_invweight[la]);
    if (lb > -1) invweight += LG(l_invweight[lb]);
    for (int k = 0; k < 4; k++) {
      const float jac_qvel = k == 0 ? jq0 : (k == 1 ? jq1 : (k == 2 ? jq2 : jq3));
      float imp, aref;
      imp_aref(sol, -pen, jac_qvel, -pen, imp, aref);
      float diag = invweight + mu * mu * invweight;
      diag *= 2.f * mu * mu * (1.0f - imp) / imp;
      diag = fmaxf(diag, EPS);
      const int c = c0 + k;
      R.R[c] = diag; R.aref[c] = aref; R.D[c] = 1.0f / diag; R.floss[c] = 0.f;
    }
  }
  // ---- joint limits (CS:1176-1233), in joint order
  if (CM(enable_joint_limit) && n_lim > 0) {
    int c = lim_base;
    for (int j = 0; j < nJ; j++) {
      const int type = CM(j_type[j]);
      if (!(type == JOINT_REVOLUTE || type == JOINT_PRISMATIC)) continue;
      const int q = CM(j_q_start[j]), d = CM(j_dof_start[j]);
      const float qv = qpos[q];
      const float pmin = qv - CM(d_limit[d][0]), pmax = CM(d_limit[d][1]) - qv;
      if (!(fminf(pmin, pmax) < 0.f)) continue;
      const float pd = fminf(pmin, pmax);
      const float jac = pmin < pmax ? 1.f : -1.f;
      float imp, aref;
      imp_aref(CM(j_sol[j]), pd, jac * vel[d], pd, imp, aref);
      const float diag = fmaxf(CM(d_invweight[d]) * (1.0f - imp) / imp, EPS);
      R.R[c] = diag; R.aref[c] = aref; R.D[c] = 1.0f / diag; R.floss[c] = 0.f;
      J[c * JS + d] = jac;
      c++;
    }
  }
}

// CS:2619-2694 func_update_constraint_batch
LHD __noinline__ void lane_update_constraint(const LaneCtx& C, const LRows& R, LSolver& X, bool track_prev) {
  const int nD = CM(nD);
  X.prev_cost = X.cost;
  float cost = 0.f;
  for (int c = 0; c < R.n_con; c++) {
    const float Ja = R.Jaref[c], D = R.D[c];
    const int prev = gi(R.active, c) & 1;
    int act = 1;
    float floss_force = 0.f;
    if (R.ne <= c && c < R.nef) {
      const float f = R.floss[c], rf = R.R[c] * f;
      const bool ln = Ja <= -rf, lp = Ja >= rf;
      act = !(ln || lp);
      floss_force = (ln ? f : 0.f) + (lp ? -f : 0.f);
      cost += (ln ? f * (-0.5f * rf - Ja) : 0.f) + (lp ? f * (-0.5f * rf + Ja) : 0.f);
    } else if (R.nef <= c) {
      act = Ja < 0.f;
    }
    gi(R.active, c) = act | (track_prev ? (prev << 1) : (act << 1));
    R.force[c] = floss_force + (-Ja * D * (float)act);
    cost += 0.5f * (Ja * Ja * D * (float)act);
  }
  float gauss = 0.f;
  const LArr qfrc = C.A(c_l.qfrc_c), Ma = C.A(c_l.Ma), force = C.A(c_l.force), qacc = C.A(c_l.qacc), acc_smooth = C.A(c_l.acc_smooth);
  for (int d = 0; d < nD; d++) {
    float q = 0.f;
    for (int c = 0; c < R.n_con; c++) q += R.J[c * R.JS + d] * R.force[c];
    qfrc[d] = q;
    gauss += 0.5f * (Ma[d] - force[d]) * (qacc[d] - acc_smooth[d]);
  }
  X.gauss = gauss;
  X.cost = cost + gauss;
}

// CS:1394-1452 + CS:1594-1621: H = M + J^T (D active) J (packed lower), in-place Cholesky, 1 / L_ii kept in Hinv
LHD __noinline__ void lane_hessian_cholesky(const LaneCtx& C, const LRows& R) {
  const int nD = CM(nD);
  const float EPS = CM(eps);
  const LArr H = C.A(c_l.H), Hinv = C.A(c_l.Hinv), M = C.A(c_l.M);
  for (int i = 0; i < nD; i++) {
    const int ri = tri_idx(i, 0);
    for (int j = 0; j <= i; j++) {
      float s = 0.f;
      for (int c = 0; c < R.n_con; c++)
        if (gi(R.active, c) & 1) s += R.J[c * R.JS + j] * (R.J[c * R.JS + i] * R.D[c]);
      H[ri + j] = s + M[ri + j];
    }
  }
  for (int i = 0; i < nD; i++) {
    const int ri = tri_idx(i, 0);
    float dt = 0.f;
    for (int k = 0; k < i; k++) dt += H[ri + k] * H[ri + k];
    const float Lii = sqrtf(fmaxf(H[ri + i] - dt, EPS));
    const float inv = 1.0f / Lii;
    H[ri + i] = Lii; Hinv[i] = inv;
    for (int j = i + 1; j < nD; j++) {
      const int rj = tri_idx(j, 0);
      float s = 0.f;
      for (int k = 0; k < i; k++) s += H[rj + k] * H[ri + k];
      H[rj + i] = (H[rj + i] - s) * inv;
    }
  }
}

// CS:1763-1822 rank-1 Cholesky up / down-dates for the rows whose active flag flipped; true when degenerate
LHD __noinline__ bool lane_hessian_incremental(const LaneCtx& C, const LRows& R) {
  const int nD = CM(nD);
  const float EPS = CM(eps);
  const LArr H = C.A(c_l.H), Hinv = C.A(c_l.Hinv), vec = C.A(c_l.prev_grad);
  for (int c = 0; c < R.n_con; c++) {
    const int a = gi(R.active, c);
    if (!((a ^ (a >> 1)) & 1)) continue;
    const float sign = (a & 1) ? 1.f : -1.f;
    const float sq = sqrtf(R.D[c]);
    for (int d = 0; d < nD; d++) vec[d] = R.J[c * R.JS + d] * sq;
    for (int k = 0; k < nD; k++) {
      const float vk = vec[k];
      if (fabsf(vk) > EPS) {
        const int rk = tri_idx(k, 0);
        const float Lkk = H[rk + k];
        const float tmp = Lkk * Lkk + sign * vk * vk;
        if (tmp < EPS) return true;
        const float r = sqrtf(tmp), cc = r / Lkk, cinv = 1.0f / cc, s = vk / Lkk;
        for (int i = k + 1; i < nD; i++) {
          const int ik = tri_idx(i, k);
          const float Lik = (H[ik] + s * vec[i] * sign) * cinv;
          H[ik] = Lik;
          vec[i] = vec[i] * cc - s * Lik;
        }
        H[rk + k] = r; Hinv[k] = 1.0f / r;
      }
    }
  }
  return false;
}

// CS:2721-2750 gradient and preconditioned gradient (Newton: Cholesky solve CS:1892-1910; CG: M^-1 via L^T D L)
LHD __noinline__ void lane_update_gradient(const LaneCtx& C) {
  const int nD = CM(nD);
  const LArr grad = C.A(c_l.grad), Mgrad = C.A(c_l.Mgrad), Ma = C.A(c_l.Ma), force = C.A(c_l.force), qfrc = C.A(c_l.qfrc_c);
  for (int d = 0; d < nD; d++) { const float g = Ma[d] - force[d] - qfrc[d]; grad[d] = g; Mgrad[d] = g; }
  if (CM(solver_type) == SOLVER_CG) {
    lane_mass_solve(C, Mgrad);
  } else {
    const LArr H = C.A(c_l.H), Hinv = C.A(c_l.Hinv);
    for (int i = 0; i < nD; i++) {  // forward substitution
      const int ri = tri_idx(i, 0);
      float s = Mgrad[i];
      for (int j = 0; j < i; j++) s -= H[ri + j] * Mgrad[j];
      Mgrad[i] = s * Hinv[i];
    }
    for (int i = nD - 1; i >= 0; i--) {  // backward substitution, column sweep
      const float xi = Mgrad[i] * Hinv[i];
      Mgrad[i] = xi;
      const int ri = tri_idx(i, 0);
      for (int j = 0; j < i; j++) Mgrad[j] -= H[ri + j] * xi;
    }
  }
}

struct LLsPoint { float alpha, cost, d0, d1; };
// CS:2154-2354 the piecewise-quadratic cost and its derivatives at step `alpha`
LHD __noinline__ LLsPoint lane_ls_eval(const LRows& R, const LSolver& X, float alpha) {
  float t0 = 0.f, t1 = 0.f, t2 = 0.f;
  for (int c = R.ne; c < R.n_con; c++) {
    const float Ja = R.Jaref[c], jv = R.jv[c], D = R.D[c];
    const float x = Ja + alpha * jv;
    if (c < R.nef) {
      const float f = R.floss[c], rf = R.R[c] * f;
      const bool ln = x <= -rf, lp = x >= rf;
      if (ln || lp) {
        t0 += (ln ? f * (-0.5f * rf - Ja) : 0.f) + (lp ? f * (-0.5f * rf + Ja) : 0.f);
        t1 += (ln ? -f * jv : 0.f) + (lp ? f * jv : 0.f);
      } else {
        t0 += D * (0.5f * Ja * Ja); t1 += D * (jv * Ja); t2 += D * (0.5f * jv * jv);
      }
    } else if (x < 0.f) {
      t0 += D * (0.5f * Ja * Ja); t1 += D * (jv * Ja); t2 += D * (0.5f * jv * jv);
    }
  }
  t0 += X.qg0 + X.es0; t1 += X.qg1 + X.es1; t2 += X.qg2 + X.es2;
  LLsPoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * t2 + alpha * t1 + t0;
  p.d0 = 2.f * alpha * t2 + t1;
  p.d1 = 2.f * t2;
  if (p.d1 <= 0.f) p.d1 = CM(eps);
  return p;
}
// CS:2357-2388
LHD __forceinline__ int lane_update_bracket(LLsPoint& p, const LLsPoint* cand, float& next_alpha) {
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
// CS:2426-2597 func_linesearch_batch (includes CS:2033-2151 init + alpha = 0 evaluation)
LHD __noinline__ float lane_linesearch(const LaneCtx& C, const LRows& R, LSolver& X) {
  const int nD = CM(nD);
  const float EPS = CM(eps);
  const LArr search = C.A(c_l.search), M = C.A(c_l.M), mv = C.A(c_l.mv), Ma = C.A(c_l.Ma), force = C.A(c_l.force);
  float sn = 0.f, qg1 = 0.f, qg2 = 0.f;
  for (int i = 0; i < nD; i++) {
    const float s = search[i];
    sn += s * s;
    float m = 0.f;
    for (int j = 0; j < nD; j++) m += M[sym_idx(i, j)] * search[j];
    mv[i] = m;
    qg1 += s * Ma[i] - s * force[i];
    qg2 += 0.5f * s * m;
  }
  const float snorm = sqrtf(sn);
  const float gtol = CM(tolerance) * CM(ls_tolerance) * snorm * (CM(meaninertia) * (float)(nD > 1 ? nD : 1));
  X.ls_it = 0;
  if (snorm < EPS) return 0.f;
  X.qg0 = X.gauss; X.qg1 = qg1; X.qg2 = qg2;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int c = 0; c < R.n_con; c++) {
    float jv = 0.f;
    for (int d = 0; d < nD; d++) jv += R.J[c * R.JS + d] * search[d];
    R.jv[c] = jv;
    if (c < R.ne) {
      const float Ja = R.Jaref[c], D = R.D[c];
      s0 += D * (0.5f * Ja * Ja); s1 += D * (jv * Ja); s2 += D * (0.5f * jv * jv);
    }
  }
  X.es0 = s0; X.es1 = s1; X.es2 = s2;
  const int ls_iterations = CM(ls_iterations);
  const LLsPoint p0 = lane_ls_eval(R, X, 0.f);
  X.ls_it = 1;
  LLsPoint p1 = lane_ls_eval(R, X, p0.alpha - p0.d0 / p0.d1);
  X.ls_it++;
  if (p0.cost < p1.cost) p1 = p0;
  if (fabsf(p1.d0) < gtol) return p1.alpha;
  const float direction = p1.d0 < 0.f ? 1.f : -1.f;
  bool p2update = false;
  LLsPoint p2 = p1;
  while (p1.d0 * direction <= -gtol && X.ls_it < ls_iterations) {
    p2 = p1; p2update = true;
    p1 = lane_ls_eval(R, X, p1.alpha - p1.d0 / p1.d1);
    X.ls_it++;
    if (fabsf(p1.d0) < gtol) return p1.alpha;
  }
  if (X.ls_it >= ls_iterations) return p1.alpha;
  if (!p2update) return p1.alpha;
  float a0 = p1.alpha - p1.d0 / p1.d1, a1 = p1.alpha, a2 = (p1.alpha + p2.alpha) * 0.5f;
  while (X.ls_it < ls_iterations) {
    LLsPoint cand[3];
    cand[0] = lane_ls_eval(R, X, a0); cand[1] = lane_ls_eval(R, X, a1); cand[2] = lane_ls_eval(R, X, a2);
    X.ls_it += 3;
    float best_alpha = 0.f, best_cost = 0.f;
    bool found = false;
#pragma unroll
    for (int i = 0; i < 3; i++)
      if (fabsf(cand[i].d0) < gtol && (!found || cand[i].cost < best_cost)) { best_alpha = cand[i].alpha; best_cost = cand[i].cost; found = true; }
    if (found) return best_alpha;
    float p1n, p2n;
    const int b1 = lane_update_bracket(p1, cand, p1n), b2 = lane_update_bracket(p2, cand, p2n);
    if (b1 == 0 && b2 == 0) return a2;
    a0 = p1n; a1 = p2n; a2 = (p1.alpha + p2.alpha) * 0.5f;
  }
  if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
  if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
  return 0.f;
}

// constraint/noslip.py: projected Gauss-Seidel on the friction-loss rows and the opposing pyramid edges with A = J M^-1 J^T.
// Only the rows M^-1 J^T are stored (SPILL record, global) and q = J^T f is updated incrementally (see gsr_solver.cuh).
LHD __noinline__ void lane_noslip(const LaneCtx& C, const LRows& R, int n_contacts) {
  const int nD = CM(nD), JS = R.JS, n_con = R.n_con;
  const float EPS = CM(eps);
  float* MJ = C.spill + c_h.spill_minvjt;  // [n_con][nD]
  const GArr b = R.jv;                     // dead after the primal solve
  const LArr q = C.A(c_l.grad), acc_smooth = C.A(c_l.acc_smooth), qfrc = C.A(c_l.qfrc_c);
  const LArr L = C.A(c_l.L), Dinv = C.A(c_l.Dinv);
  for (int r = 0; r < n_con; r++) {  // noslip.py:9-58
    float* x = MJ + (size_t)r * nD;
    for (int d = 0; d < nD; d++) x[d] = R.J[r * JS + d];
    for (int j = nD - 1; j > 0; j--) {
      const float xj = x[j];
      const int rj = tri_idx(j, 0);
      for (int i = 0; i < j; i++) x[i] -= L[rj + i] * xj;
    }
    for (int i = 0; i < nD; i++) x[i] *= Dinv[i];
    for (int i = 1; i < nD; i++) {
      const int ri = tri_idx(i, 0);
      float s = x[i];
      for (int j = 0; j < i; j++) s -= L[ri + j] * x[j];
      x[i] = s;
    }
    float v = -R.aref[r];
    for (int d = 0; d < nD; d++) v += R.J[r * JS + d] * acc_smooth[d];
    b[r] = v;
  }
  for (int d = 0; d < nD; d++) q[d] = qfrc[d];
  const float scale = 1.0f / (CM(meaninertia) * (float)(nD > 1 ? nD : 1));
  const int const_start = R.nef;
  for (int it = 0; it < CM(noslip_iterations); it++) {  // noslip.py:166-283
    float improvement = 0.f;
    if (it == 0)
      for (int c = 0; c < n_con; c++) improvement += 0.5f * R.force[c] * R.force[c] * R.R[c];
    for (int c = R.ne; c < R.nef; c++) {
      float res = b[c], Acc = 0.f;
      for (int d = 0; d < nD; d++) { const float mj = MJ[(size_t)c * nD + d]; res += mj * q[d]; Acc += R.J[c * JS + d] * mj; }
      const float old = R.force[c], fl = R.floss[c];
      float f = old - res / Acc;
      f = f < -fl ? -fl : (f > fl ? fl : f);
      const float delta = f - old;
      improvement -= 0.5f * delta * delta / Acc + delta * res;
      R.force[c] = f;
      for (int d = 0; d < nD; d++) q[d] += R.J[c * JS + d] * delta;
    }
    for (int ic = 0; ic < n_contacts; ic++) {
      for (int j2 = 0; j2 < 2; j2++) {
        const int j = const_start + 4 * ic + 2 * j2;
        float res0 = b[j], res1 = b[j + 1], a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
        for (int d = 0; d < nD; d++) {
          const float m0 = MJ[(size_t)j * nD + d], m1 = MJ[(size_t)(j + 1) * nD + d], j0 = R.J[j * JS + d], j1 = R.J[(j + 1) * JS + d], qd = q[d];
          res0 += m0 * qd; res1 += m1 * qd; a00 += j0 * m0; a01 += j1 * m0; a10 += j0 * m1; a11 += j1 * m1;
        }
        const float o0 = R.force[j], o1 = R.force[j + 1];
        const float bc0 = res0 - (a00 * o0 + a01 * o1), bc1 = res1 - (a10 * o0 + a11 * o1);
        const float mid = 0.5f * (o0 + o1);
        const float K1 = a00 + a11 - a01 - a10, K0 = mid * (a00 - a11) + bc0 - bc1;
        float f0 = mid, f1 = mid;
        if (!(K1 < EPS)) {
          const float y = -K0 / K1;
          if (y < -mid) { f0 = 0.f; f1 = 2.f * mid; }
          else if (y > mid) { f0 = 2.f * mid; f1 = 0.f; }
          else { f0 = mid + y; f1 = mid - y; }
        }
        float d0 = f0 - o0, d1 = f1 - o1;  // noslip.py:358-385 func_cost_change
        float change = 0.5f * (a00 * d0 * d0 + a01 * d0 * d1 + a10 * d1 * d0 + a11 * d1 * d1) + d0 * res0 + d1 * res1;
        if (change > EPS) { f0 = o0; f1 = o1; d0 = 0.f; d1 = 0.f; change = 0.f; }
        improvement -= change;
        R.force[j] = f0; R.force[j + 1] = f1;
        for (int d = 0; d < nD; d++) q[d] += R.J[j * JS + d] * d0 + R.J[(j + 1) * JS + d] * d1;
      }
    }
    improvement *= scale;
    if (improvement < CM(noslip_tolerance)) break;
  }
  // noslip.py:286-323 kernel_dual_finish
  const LArr x = C.A(c_l.tmpv), qacc = C.A(c_l.qacc), acc = C.A(c_l.acc);
  for (int d = 0; d < nD; d++) { qfrc[d] = q[d]; x[d] = q[d]; }
  lane_mass_solve(C, x);
  for (int d = 0; d < nD; d++) { const float a = x[d] + acc_smooth[d]; qacc[d] = a; acc[d] = a; }
}

// CS:210-245 resolve(): init (CS:2935-3059), monolith body (CS:3163-3190), update_qacc (CS:3240-3261), contact forces
LHD __noinline__ void lane_constraints(const LaneCtx& C) {
  const int nD = CM(nD), nL = CM(nL);
  const float EPS = CM(eps);
  LRows R;
  lane_build_rows(C, R);
  const int n_con = R.n_con;
  LSolver X;
  X.cost = 0.f; X.prev_cost = 0.f; X.gauss = 0.f; X.ls_it = 0;
  const LArr qacc = C.A(c_l.qacc), Ma = C.A(c_l.Ma), search = C.A(c_l.search), Mgrad = C.A(c_l.Mgrad), grad = C.A(c_l.grad);
  const LArr mv = C.A(c_l.mv), prev_grad = C.A(c_l.prev_grad), prev_Mgrad = C.A(c_l.prev_Mgrad);
  const LArr acc_smooth = C.A(c_l.acc_smooth), qacc_ws = C.A(c_h.qacc_ws), qfrc = C.A(c_l.qfrc_c), M = C.A(c_l.M);
  const bool warm = n_con > 0 && C.I(c_h.is_warmstart)[0];
  for (int d = 0; d < nD; d++) qacc[d] = warm ? qacc_ws[d] : acc_smooth[d];
  int iters = 0;
  if (n_con > 0) {
    for (int i = 0; i < nD; i++) {  // CS:2910-2929
      float s = 0.f;
      for (int j = 0; j < nD; j++) s += M[sym_idx(i, j)] * qacc[j];
      Ma[i] = s;
    }
    for (int c = 0; c < n_con; c++) {  // CS:2887-2907
      float s = -R.aref[c];
      for (int d = 0; d < nD; d++) s += R.J[c * R.JS + d] * qacc[d];
      R.Jaref[c] = s;
      gi(R.active, c) = 0;
    }
    lane_update_constraint(C, R, X, false);
    if (CM(solver_type) == SOLVER_NEWTON) lane_hessian_cholesky(C, R);
    lane_update_gradient(C);
    for (int d = 0; d < nD; d++) search[d] = -Mgrad[d];
    const float tol_scaled = (CM(meaninertia) * (float)(nD > 1 ? nD : 1)) * CM(tolerance);
    for (int it = 0; it < CM(iterations); it++) {  // CS:3062-3139 func_solve_iter
      iters++;
      const float alpha = lane_linesearch(C, R, X);
      if (fabsf(alpha) < EPS) break;
      for (int d = 0; d < nD; d++) {
        qacc[d] += search[d] * alpha;
        Ma[d] += mv[d] * alpha;
        if (CM(solver_type) == SOLVER_CG) { prev_grad[d] = grad[d]; prev_Mgrad[d] = Mgrad[d]; }
      }
      for (int c = 0; c < n_con; c++) R.Jaref[c] += R.jv[c] * alpha;
      lane_update_constraint(C, R, X, true);
      if (CM(solver_type) == SOLVER_NEWTON) {
        if (lane_hessian_incremental(C, R)) lane_hessian_cholesky(C, R);
      }
      lane_update_gradient(C);
      // CS:2841-2884 termination / new search direction
      float gn = 0.f;
      for (int d = 0; d < nD; d++) { const float g = grad[d]; gn += g * g; }
      gn = sqrtf(gn);
      const bool improved = gn > tol_scaled && (X.prev_cost - X.cost) > tol_scaled;
      if (!improved) break;
      if (CM(solver_type) == SOLVER_NEWTON) {
        for (int d = 0; d < nD; d++) search[d] = -Mgrad[d];
      } else {
        float beta = 0.f, pg = 0.f;
        for (int d = 0; d < nD; d++) { beta += grad[d] * (Mgrad[d] - prev_Mgrad[d]); pg += prev_Mgrad[d] * prev_grad[d]; }
        beta = fmaxf(beta / fmaxf(EPS, pg), 0.f);
        for (int d = 0; d < nD; d++) search[d] = -Mgrad[d] + beta * search[d];
      }
    }
  } else {
    for (int d = 0; d < nD; d++) qfrc[d] = 0.f;
  }
  // ---- CS:3240-3261 func_update_qacc
  bool has_nan = false;
  const LArr acc = C.A(c_l.acc);
  for (int d = 0; d < nD; d++) {
    const float a = qacc[d];
    acc[d] = a; qacc_ws[d] = a;
    has_nan |= isnan(a);
  }
  if (has_nan) C.I(c_h.errno_)[0] |= ERR_NAN_FORCE;
  C.I(c_h.is_warmstart)[0] = 1;
  const int nc = CM(enable_collision) ? C.I(c_h.n_contacts)[0] : 0;
  if (CM(noslip_iterations) > 0 && n_con > 0) lane_noslip(C, R, nc);  // CS:237-238
  // ---- CS:3198-3237 func_update_contact_force: per-contact force (global record) and per-link totals
  const LArr lcf = C.A(c_l.lcf);
  for (int i = 0; i < 3 * nL; i++) lcf[i] = 0.f;
  for (int ic = 0; ic < nc; ic++) {
    const int o = ic * GSR_CONTACT_WORDS;
    const f3 cn(R.ct[o + CT_NRM], R.ct[o + CT_NRM + 1], R.ct[o + CT_NRM + 2]);
    const float mu = R.ct[o + CT_MU];
    f3 d1, d2;
    orthogonals(cn, d1, d2);
    const int r0 = R.nef + 4 * ic;
    const f3 fo = (d1 * (-mu) - cn) * R.force[r0] + (d1 * mu - cn) * R.force[r0 + 1] + (d2 * (-mu) - cn) * R.force[r0 + 2] + (d2 * mu - cn) * R.force[r0 + 3];
    if (C.live) { float* g = C.contacts + (size_t)o; g[CT_FRC] = fo.x; g[CT_FRC + 1] = fo.y; g[CT_FRC + 2] = fo.z; }
    const int la = lane_f2i(R.ct[o + CT_LA]), lb = lane_f2i(R.ct[o + CT_LB]);
    st3(lcf, la, ld3(lcf, la) - fo);
    if (lb > -1) st3(lcf, lb, ld3(lcf, lb) + fo);
  }
  // counters are staged after lcf's neighbours are dead (the caller flushes c_l.cnt)
  X.ls_it = iters;
  // the solver scratch is dead now: stage the counters where the output block expects them
  const LIArr cnt = C.I(c_l.cnt);
  const int ne = R.ne, nfl = R.nef - R.ne;
  cnt[0] = n_con; cnt[1] = ne; cnt[2] = nfl; cnt[3] = iters;
}

// ---------------------------------------------------------------------------------------------------------------------
// Integration (RS:2783-2891 kernel_step_2): link accelerations, implicit damping, semi-implicit Euler with NaN guard
// ---------------------------------------------------------------------------------------------------------------------
LHD __noinline__ void lane_integrate(const LaneCtx& C) {
  const int nL = CM(nL), nD = CM(nD), nQ = CM(nQ);
  const float dt = CM(dt), EPS = CM(eps);
  const LArr vel = C.A(c_h.vel), qpos = C.A(c_h.qpos), acc = C.A(c_l.acc);
  const LArr cda = C.A(c_l.cdof), cdv = C.A(c_l.cdof + 3 * nD), cdda = C.A(c_l.cdof + 6 * nD), cddv = C.A(c_l.cdof + 9 * nD);
  // FD:1186-1285 update_acc(update_cacc=True): link accelerations from the constrained acceleration, root -> leaf
  const LArr clin = C.A(c_l.cacc), cang = C.A(c_l.cacc + 3 * nL);
  for (int l = 0; l < nL; l++) {
    const int p = CM(l_parent[l]);
    f3 lin, ang;
    if (p >= 0) { lin = ld3(clin, p); ang = ld3(cang, p); }
    for (int d = CM(l_dof_start[l]); d < CM(l_dof_end[l]); d++) {
      const float v = vel[d], a = acc[d];
      lin += ld3(cddv, d) * v + ld3(cdv, d) * a;
      ang += ld3(cdda, d) * v + ld3(cda, d) * a;
    }
    st3(clin, l, lin); st3(cang, l, ang);
  }
  if (CM(integrator) != INTEGRATOR_APPROX_IMPLICITFAST) {
    // FD:1787-1859 func_implicit_damping: refactor M + dt (damping - d act / d vel) and re-solve acc = M'^-1 force
    const LIArr mode = C.I(c_h.ctrl_mode);
    bool need = false;
    for (int d = 0; d < nD; d++)
      need |= CM(d_damping[d]) > EPS || (CM(integrator) != INTEGRATOR_EULER && mode[d] <= CTRL_VELOCITY && fabsf(CM(d_act_bias[d][2])) > EPS);
    if (need) {
      const LArr L = C.A(c_l.L), M = C.A(c_l.M), x = C.A(c_l.vn);
      for (int i = 0; i < nD; i++) {
        // the reference refactors entity by entity and skips entities without damping / velocity-dependent actuation
        const int e = CM(d_entity[i]);
        bool ent_active = false;
        for (int k = CM(e_dof_start[e]); k < CM(e_dof_end[e]); k++)
          ent_active |= CM(d_damping[k]) > EPS || (CM(integrator) != INTEGRATOR_EULER && mode[k] <= CTRL_VELOCITY && fabsf(CM(d_act_bias[k][2])) > EPS);
        const int ri = tri_idx(i, 0);
        for (int j = 0; j <= i; j++) {
          float v = M[ri + j];
          if (i == j && ent_active) {
            v += CM(d_damping[i]) * dt;
            if (CM(integrator) == INTEGRATOR_IMPLICITFAST && mode[i] <= CTRL_VELOCITY) v -= CM(d_act_bias[i][2]) * dt;
          }
          L[ri + j] = v;
        }
        x[i] = C.A(c_l.force)[i] + C.A(c_l.qfrc_c)[i];  // dofs_state.force = qf_smooth + qf_constraint
        C.A(c_l.qn)[i] = ent_active ? 1.f : 0.f;        // (qn is rebuilt below; nQ >= nD)
      }
      lane_mass_factor(C);
      lane_mass_solve(C, x);
      for (int d = 0; d < nD; d++)
        if (C.A(c_l.qn)[d] != 0.f) acc[d] = x[d];
    }
  }
  // stage acc / force / qf_applied / qf_constraint in record order (flushed by the caller)
  const LArr out4 = C.A(c_l.out4);
  for (int d = 0; d < nD; d++) {
    const float qc = C.A(c_l.qfrc_c)[d];
    out4[d] = acc[d]; out4[nD + d] = C.A(c_l.force)[d] + qc; out4[2 * nD + d] = C.A(c_l.qf_applied)[d]; out4[3 * nD + d] = qc;
  }
  // FD:1566-1707 func_integrate + diff.py:25-54 (commit only if finite)
  const LArr vn = C.A(c_l.vn), qn = C.A(c_l.qn);
  bool bad = false;
  for (int d = 0; d < nD; d++) { const float v = vel[d] + acc[d] * dt; vn[d] = v; bad |= isnan(v); }
  for (int l = 0; l < nL; l++) {
    const int ds = CM(l_dof_start[l]), de = CM(l_dof_end[l]), qs = CM(l_q_start[l]), qe = CM(l_q_end[l]);
    if (de - ds <= 0) continue;
    const int jt = CM(j_type[CMI(l_joint_start[l])]);
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
  for (int q = 0; q < nQ; q++) bad |= isnan(qn[q]);
  if (bad) {
    C.I(c_h.errno_)[0] |= ERR_NAN_ACC;
  } else {
    for (int d = 0; d < nD; d++) vel[d] = vn[d];
    for (int q = 0; q < nQ; q++) qpos[q] = qn[q];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// One substep of one environment (RS:912-980): kernel_step_1 -> _func_constraint_force -> kernel_step_2
// ---------------------------------------------------------------------------------------------------------------------
LHD __forceinline__ void lane_step(const LaneCtx& C, int clear_ext) {
  const int nL = CM(nL), nD = CM(nD), nG = CM(nG), nQ = CM(nQ);
  lane_coop_load(C, 0, 0, c_l.n_head);
  lane_kinematics(C, c_l.cvel);
  lane_dynamics(C);
  if (CM(enable_collision) && nG > 0) {
    lane_geoms(C);
    lane_coop_flush(C, c_l.aabb, c_h.aabb_min, 6 * nG);  // get_AABB: aabb_min[3 nG] aabb_max[3 nG] are adjacent in the record
    lane_collide(C);
  }
  lane_constraints(C);
  lane_coop_flush(C, c_l.lcf, c_h.links_contact_force, 3 * nL);
  lane_integrate(C);
  lane_coop_flush(C, c_l.cacc, c_h.cacc_lin, 6 * nL);
  lane_coop_flush(C, c_l.out4, c_h.acc, 4 * nD);
  lane_coop_flush(C, c_l.cnt, c_h.n_constraints, 4);
  // ---- kinematics of the new configuration (RS:2866-2891): what the getters and the next step's k_forward users see
  lane_kinematics(C, c_l.fcvel);
  lane_coop_flush(C, c_l.fcvel, c_h.cd_vel, 6 * nL);
  {
    const LArr ro = C.A(c_l.rcom_out), rcom = C.A(c_l.rcom);
    for (int l = 0; l < nL; l++) st3(ro, l, ld3(rcom, CM(l_rootslot[l])));
  }
  lane_coop_flush(C, c_l.rcom_out, c_h.root_COM, 3 * nL);
  lane_coop_flush(C, c_l.lpos, c_h.links_pos, 7 * nL);
  if (CM(diagnostics)) {  // internals of the kinematic cache (this kernel recomputes them; k_forward and the tests read them)
    lane_coop_flush(C, c_l.cdof, c_h.cdof_ang, 12 * nD);
    lane_cinr(C, c_l.cinr);
    lane_coop_flush(C, c_l.cinr, c_h.cinr_I, 9 * nL);
  }
  if (CM(diagnostics) && C.live) {  // geoms_state.pos / quat of the post-integration configuration (FK:828-864)
    for (int g = 0; g < nG; g++) {
      f3 gp; f4 gq;
      lane_geom_pose(C, g, gp, gq);
      gst3(C.rec + c_h.geoms_pos, g, gp);
      float* q = C.rec + c_h.geoms_quat + 4 * g;
      q[0] = gq.w; q[1] = gq.x; q[2] = gq.y; q[3] = gq.z;
    }
  }
  if (clear_ext && C.I(c_h.has_ext_force)[0]) {  // RS:2603-2612 clear_external_force
    if (C.live)
      for (int i = 0; i < 3 * nL; i++) { C.rec[c_h.cfrc_applied_vel + i] = 0.f; C.rec[c_h.cfrc_applied_ang + i] = 0.f; }
    C.I(c_h.has_ext_force)[0] = 0;
  }
  // ---- state head: qpos, vel | qacc_ws, flags, dofs_pos (the control targets in between are read-only)
  lane_coop_flush(C, c_h.qpos, c_h.qpos, nQ + nD);
  lane_coop_flush(C, c_h.qacc_ws, c_h.qacc_ws, c_l.n_head - c_h.qacc_ws);
}

#ifdef __CUDACC__
// One warp = 32 environments; warps are independent (no CTA-level synchronisation), W warps per CTA share the dynamic
// shared memory: warp w owns words [w * env_words * 33, (w + 1) * env_words * 33).
__global__ void __launch_bounds__(128) k_step_lane(float* state, int B, const DevModel* gm, int clear_ext) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env0 = (blockIdx.x * (blockDim.x >> 5) + warp) * 32;
  if (env0 >= B) return;
  LaneCtx C;
  C.n_live = min(32, B - env0);
  C.lane = lane;
  C.live = lane < C.n_live;
  const int env = env0 + min(lane, C.n_live - 1);  // padding lanes shadow the last live environment (never flushed)
  C.Sw = smem + (size_t)warp * c_l.env_words * GSR_LANE_STRIDE_DEV;
  C.S = C.Sw + lane;
  C.rec0 = state + (size_t)env0 * c_h.hot_stride;
  C.rec = state + (size_t)env * c_h.hot_stride;
  C.contacts = state + c_h.contact_base + (size_t)env * c_h.contact_stride;
  C.cache = state + c_h.cache_base + (size_t)env * c_h.cache_stride;
  C.spill = state + c_h.spill_base + (size_t)env * c_h.spill_stride;
  C.gm = gm;
  lane_step(C, clear_ext);
}
#endif
