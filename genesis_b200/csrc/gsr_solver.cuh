// gsr_solver.cuh -- constraint rows and the Newton / CG solver of one environment by one lane group.
//
// Rows live in shared memory (J[NCS][JS] plus eight per-row scalars) and never touch HBM; an environment with more rows
// than the on-chip capacity runs the same code on its SPILL record in global memory (generic pointers).  Lane mapping:
//   row assembly : one lane per (contact, dof) Jacobian task / per row for the scalar parameters
//   J v, J^T f   : one lane per row / per dof
//   H = M + J^T D J : one lane per lower-triangle entry, Cholesky column by column with the rows of a column on the lanes
//   line search  : one lane per row for the piecewise-quadratic sums, butterfly reductions give every lane the same
//                  totals so the control flow of the exact line search stays uniform inside the group
// Semantics follow the reference's monolith solver body (CS:2033-3261), which is what `gs.cpu` runs.
#pragma once
#include "gsr_group.cuh"
#include "gsr_dynamics.cuh"
#undef IN
#undef WK
#if GSR_LDS_MASK & 4
#define IN(field) IN_SHARED(field)
#define WK(field) WK_SHARED(field)
#else
#define IN(field) IN_GENERIC(field)
#define WK(field) WK_GENERIC(field)
#endif

struct Rows {
  float* J; int JS;
  float *aref, *D, *R, *floss, *Jaref, *jv, *force;
  int* active;
  const float* ct;  // contact records (16 words each)
  int n_con, ne, nef;
};

struct SolverState {
  float cost, prev_cost, gauss;
  float qg0, qg1, qg2, es0, es1, es2;
  int ls_it;
};

// CS:984-1173 func_equality_weld: 3 position rows (anchor1 = data[3:6] in body 1, anchor2 = data[0:3] in body 2) + 3 orientation
// rows (relpose = data[6:10], torquescale = data[10]); translational / rotational invweight components
template <int G>
__device__ __noinline__ void weld_rows(const Ctx<G>& C, Rows& R, int row, int l1, int l2, const float* dat, const float* sol) {
  const int nD = CM(nD), lane = C.lane, JS = R.JS;
  const float EPS = CM(eps);
  float* J = R.J;
  const float* vel = IN(vel);
  const f4 q1 = ld4(IN(links_quat), l1), q2 = ld4(IN(links_quat), l2);
  const f3 g1 = transform(f3(dat[3], dat[4], dat[5]), ld3(IN(links_pos), l1), q1);
  const f3 g2 = transform(f3(dat[0], dat[1], dat[2]), ld3(IN(links_pos), l2), q2);
  const f3 t1 = g1 - ld3(IN(root_COM), l1), t2 = g2 - ld3(IN(root_COM), l2);
  const float ts = dat[10];
  const f4 q = quat_mul(q1, f4(dat[6], dat[7], dat[8], dat[9]));
  const f4 iq2 = quat_inv(q2);
  const f4 eq = quat_mul(iq2, q);
  const f3 pos_error = g1 - g2, rot_error = f3(eq.x, eq.y, eq.z) * ts;
  const unsigned long long m1 = GM(l_dofmask[l1]), m2 = GM(l_dofmask[l2]);
  for (int d = lane; d < nD; d += G) {
    f3 v, jr;
    const f3 ca = ld3(IN(cdof_ang), d);
    if (bit64(m1, d)) { v += ld3(IN(cdof_vel), d) - cross(t1, ca); jr += ca; }
    if (bit64(m2, d)) { v -= ld3(IN(cdof_vel), d) - cross(t2, ca); jr -= ca; }
    J[row * JS + d] = v.x; J[(row + 1) * JS + d] = v.y; J[(row + 2) * JS + d] = v.z;
    const f4 q3 = quat_mul(quat_mul_axis(iq2, jr), q);
    J[(row + 3) * JS + d] = 0.5f * q3.x * ts; J[(row + 4) * JS + d] = 0.5f * q3.y * ts; J[(row + 5) * JS + d] = 0.5f * q3.z * ts;
  }
  C.sync();
  if (lane < 6) {
    float jac_qvel = 0.f;
    for (int d = 0; d < nD; d++) jac_qvel += J[(row + lane) * JS + d] * vel[d];
    const float pos_imp = sqrtf(dot(pos_error, pos_error) + dot(rot_error, rot_error));
    const float invweight = lane < 3 ? GM(l_invweight[l1]) + GM(l_invweight[l2]) : GM(l_invweight_rot[l1]) + GM(l_invweight_rot[l2]);
    float imp, aref;
    imp_aref(sol, -pos_imp, jac_qvel, lane < 3 ? comp(pos_error, lane) : comp(rot_error, lane - 3), imp, aref);
    const float diag = fmaxf(invweight * (1.0f - imp) / imp, EPS);
    R.R[row + lane] = diag; R.aref[row + lane] = aref; R.D[row + lane] = 1.0f / diag; R.floss[row + lane] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Row assembly (CS:579-1286), row order [equality | frictionloss | 4 x contacts | joint limits]
// ---------------------------------------------------------------------------------------------------------------------
template <int G>
__device__ GSR_ROWS_NOINLINE void build_rows(const Ctx<G>& C, Rows& R) {
  const int nD = CM(nD), nJ = CM(nJ), lane = C.lane;
  const float EPS = CM(eps);
  const float* qpos = IN(qpos);
  const float* vel = IN(vel);
  const bool off = CM(disable_constraint) != 0;  // RigidOptions.disable_constraint: collision detection still runs, no rows are built
  const int n_contacts = (CM(enable_collision) && !off) ? C.SI()[c_h.n_contacts] : 0;
  // ---- static counts: equality rows and friction-loss rows
  int n_eq = 0;
  for (int i = 0; i < (off ? 0 : CM(nEq)); i++) n_eq += CM(eq_type[i]) == EQ_JOINT ? 1 : (CM(eq_type[i]) == EQ_CONNECT ? 3 : (CM(eq_type[i]) == EQ_WELD ? 6 : 0));
  const int n_dyn = off ? 0 : min(max(C.SI()[c_h.n_dyn_eq], 0), GSR_MAX_DYN_EQ);
  n_eq += 6 * n_dyn;
  int n_floss = 0;
  for (int d = 0; d < (off ? 0 : nD); d++) n_floss += dof_frictionloss(C, d) > EPS ? 1 : 0;
  // ---- joint limits that are violated (CS:1176-1233), in joint order
  int n_lim = 0;
  const int lim_base = n_eq + n_floss + 4 * n_contacts;
  // first pass only counts, second pass (below) fills once the storage is chosen
  if (CM(enable_joint_limit) && !off) {
    for (int base = 0; base < nJ; base += G) {
      const int j = base + lane;
      bool viol = false;
      if (j < nJ) {
        const int type = GM(j_type[j]);
        if (type == JOINT_REVOLUTE || type == JOINT_PRISMATIC) {
          const int q = GM(j_q_start[j]), d = GM(j_dof_start[j]);
          const float qv = qpos[q];
          viol = fminf(qv - dof_limit(C, d, 0), dof_limit(C, d, 1) - qv) < 0.f;
        }
      }
      n_lim += __popc(C.ballot(viol));
    }
  }
  const int n_con = lim_base + n_lim;
  R.n_con = n_con; R.ne = n_eq; R.nef = n_eq + n_floss;
  // ---- storage: shared memory when the rows fit, the SPILL record otherwise
  const int NCS = c_s.NCS, lenC = CM(len_constraints);
  R.JS = c_s.JS;
  if (n_con <= NCS) {
    R.J = WK(J);
    R.aref = WK(r_aref); R.D = WK(r_D); R.R = WK(r_R); R.floss = WK(r_floss); R.Jaref = WK(r_Jaref); R.jv = WK(r_jv);
    R.force = WK(r_force); R.active = reinterpret_cast<int*>(WK(r_active));
  } else {
    float* sp = C.spill;
    R.J = sp; sp += (size_t)lenC * R.JS;
    R.aref = sp; sp += lenC; R.D = sp; sp += lenC; R.R = sp; sp += lenC; R.floss = sp; sp += lenC; R.Jaref = sp; sp += lenC;
    R.jv = sp; sp += lenC; R.force = sp; sp += lenC; R.active = reinterpret_cast<int*>(sp);
  }
  R.ct = (n_contacts <= c_s.NCC) ? WK(contacts) : C.contacts;
  if (n_con == 0) return;
  float* J = R.J;
  const int JS = R.JS;
  // ---- zero the non-contact rows
  for (int r0 = 0; r0 < R.nef + n_lim; r0++) {
    const int r = r0 < R.nef ? r0 : lim_base + (r0 - R.nef);
    for (int d = lane; d < nD; d += G) J[r * JS + d] = 0.f;
  }
  C.sync();
  // ---- equalities (CS:880-937)
  int row = 0;
  for (int i = 0; i < CM(nEq); i++) {
    if (CM(eq_type[i]) == EQ_JOINT) {  // CS:800-877
      if (lane == 0) {
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
        const float invweight = dof_invweight(C, d1) + dof_invweight(C, d2);
        float imp, aref;
        imp_aref(CM(eq_sol[i]), -fabsf(pos), jac_qvel, pos, imp, aref);
        const float diag = fmaxf(invweight * (1.0f - imp) / imp, EPS);
        R.R[row] = diag; R.aref[row] = aref; R.D[row] = 1.0f / diag; R.floss[row] = 0.f;
      }
      row += 1;
    } else if (CM(eq_type[i]) == EQ_CONNECT) {  // CS:688-797
      const int l1 = CM(eq_obj1[i]), l2 = CM(eq_obj2[i]);
      const f3 g1 = transform(f3(CM(eq_data[i][0]), CM(eq_data[i][1]), CM(eq_data[i][2])), ld3(IN(links_pos), l1), ld4(IN(links_quat), l1));
      const f3 g2 = transform(f3(CM(eq_data[i][3]), CM(eq_data[i][4]), CM(eq_data[i][5])), ld3(IN(links_pos), l2), ld4(IN(links_quat), l2));
      const f3 t1 = g1 - ld3(IN(root_COM), l1), t2 = g2 - ld3(IN(root_COM), l2);
      const unsigned long long m1 = CM(l_dofmask[l1]), m2 = CM(l_dofmask[l2]);
      for (int d = lane; d < nD; d += G) {
        f3 v;
        if (bit64(m1, d)) v += ld3(IN(cdof_vel), d) - cross(t1, ld3(IN(cdof_ang), d));
        if (bit64(m2, d)) v -= ld3(IN(cdof_vel), d) - cross(t2, ld3(IN(cdof_ang), d));
        J[row * JS + d] = v.x; J[(row + 1) * JS + d] = v.y; J[(row + 2) * JS + d] = v.z;
      }
      C.sync();
      if (lane < 3) {
        float jac_qvel = 0.f;
        for (int d = 0; d < nD; d++) jac_qvel += J[(row + lane) * JS + d] * vel[d];
        const f3 pd = g1 - g2;
        const float invweight = CM(l_invweight[l1]) + CM(l_invweight[l2]);
        float imp, aref;
        imp_aref(CM(eq_sol[i]), -norm(pd), jac_qvel, comp(pd, lane), imp, aref);
        const float diag = fmaxf(invweight * (1.0f - imp) / imp, EPS);
        R.R[row + lane] = diag; R.aref[row + lane] = aref; R.D[row + lane] = 1.0f / diag; R.floss[row + lane] = 0.f;
      }
      row += 3;
    } else if (CM(eq_type[i]) == EQ_WELD) {
      weld_rows(C, R, row, CM(eq_obj1[i]), CM(eq_obj2[i]), CM(eq_data[i]), CM(eq_sol[i]));
      row += 6;
    }
  }
  // dynamic welds of this environment (add_weld_constraint), after the static equalities like in the reference's table
  for (int k = 0; k < n_dyn; k++) {
    const float* rec_eq = C.rec + c_h.dyn_eq + k * GSR_DYN_EQ_WORDS;
    weld_rows(C, R, row, __float_as_int(rec_eq[0]), __float_as_int(rec_eq[1]), rec_eq + 2, rec_eq + 13);
    row += 6;
  }
  // ---- frictionloss (CS:1236-1286), static rows in dof order
  if (n_floss > 0 && lane == 0) {
    int r = n_eq;
    for (int d = 0; d < nD; d++)
      if (dof_frictionloss(C, d) > EPS) {
        float imp, aref;
        { const int jd_ = CM(d_joint[d]); imp_aref(CM(j_sol[jd_]), 0.f, vel[d], 0.f, imp, aref); }
        const float diag = fmaxf(dof_invweight(C, d) * (1.0f - imp) / imp, EPS);
        R.R[r] = diag; R.aref[r] = aref; R.D[r] = 1.0f / diag; R.floss[r] = dof_frictionloss(C, d);
        J[r * JS + d] = 1.0f;
        r++;
      }
  }
  // ---- contacts (CS:579-685): pyramidal friction, 4 rows per contact; one lane per (contact, dof) entry
  const int c_base = R.nef;
  // task t = ic * nD + d, walked without integer division
  const int ntask = n_contacts * nD;
  int ic = 0, d = lane;
  while (d >= nD && ic < n_contacts) { d -= nD; ic++; }
  for (int t = lane; t < ntask; t += G) {
    const float* ct = R.ct + (size_t)ic * GSR_CONTACT_WORDS;
    const int la = __float_as_int(ct[CT_LA]), lb = __float_as_int(ct[CT_LB]);
    const f3 cpos(ct[CT_POS], ct[CT_POS + 1], ct[CT_POS + 2]), cn(ct[CT_NRM], ct[CT_NRM + 1], ct[CT_NRM + 2]);
    const float mu = ct[CT_MU];
    f3 d1, d2;
    orthogonals(cn, d1, d2);
    f3 diff;
    const f3 cv = ld3(IN(cdof_vel), d), ca = ld3(IN(cdof_ang), d);
    if (bit64(GM(l_dofmask[la]), d)) diff -= cv - cross(cpos - ld3(IN(root_COM), la), ca);
    if (lb > -1 && bit64(GM(l_dofmask[lb]), d)) diff += cv - cross(cpos - ld3(IN(root_COM), lb), ca);
    const int c0 = c_base + 4 * ic;
    J[c0 * JS + d] = dot(diff, d1 * (-mu) - cn);
    J[(c0 + 1) * JS + d] = dot(diff, d1 * mu - cn);
    J[(c0 + 2) * JS + d] = dot(diff, d2 * (-mu) - cn);
    J[(c0 + 3) * JS + d] = dot(diff, d2 * mu - cn);
    d += G;
    while (d >= nD) { d -= nD; ic++; }
  }
  C.sync();
  for (int r = lane; r < 4 * n_contacts; r += G) {
    const int ic = r >> 2, c = c_base + r;
    const float* ct = R.ct + (size_t)ic * GSR_CONTACT_WORDS;
    const int la = __float_as_int(ct[CT_LA]), lb = __float_as_int(ct[CT_LB]);
    const int ga = __float_as_int(ct[CT_GA]), gb = __float_as_int(ct[CT_GB]);
    const float mu = ct[CT_MU], pen = ct[CT_PEN];
    float jac_qvel = 0.f;
    for (int d = 0; d < nD; d++) jac_qvel += J[c * JS + d] * vel[d];
    float sol[7];
#pragma unroll
    for (int k = 0; k < 7; k++) sol[k] = 0.5f * (GM(g_sol[ga][k]) + GM(g_sol[gb][k]));
    float invweight = GM(l_invweight[la]);
    if (lb > -1) invweight += GM(l_invweight[lb]);
    float imp, aref;
    imp_aref(sol, -pen, jac_qvel, -pen, imp, aref);
    float diag = invweight + mu * mu * invweight;
    diag *= 2.f * mu * mu * (1.0f - imp) / imp;
    diag = fmaxf(diag, EPS);
    R.R[c] = diag; R.aref[c] = aref; R.D[c] = 1.0f / diag; R.floss[c] = 0.f;
  }
  // ---- joint limits (CS:1176-1233): ordered compaction of the violated joints
  if (CM(enable_joint_limit) && n_lim > 0) {
    int r0 = lim_base;
    for (int base = 0; base < nJ; base += G) {
      const int j = base + lane;
      bool viol = false;
      int d = 0;
      float pmin = 0.f, pmax = 0.f;
      if (j < nJ) {
        const int type = GM(j_type[j]);
        if (type == JOINT_REVOLUTE || type == JOINT_PRISMATIC) {
          const int q = GM(j_q_start[j]);
          d = GM(j_dof_start[j]);
          const float qv = qpos[q];
          pmin = qv - dof_limit(C, d, 0); pmax = dof_limit(C, d, 1) - qv;
          viol = fminf(pmin, pmax) < 0.f;
        }
      }
      const unsigned bal = C.ballot(viol);
      if (viol) {
        const int c = r0 + __popc(bal & ((1u << lane) - 1u));
        const float pd = fminf(pmin, pmax);
        const float jac = pmin < pmax ? 1.f : -1.f;
        float sol[7];
#pragma unroll
        for (int k = 0; k < 7; k++) sol[k] = GM(j_sol[j][k]);
        float imp, aref;
        imp_aref(sol, pd, jac * vel[d], pd, imp, aref);
        const float diag = fmaxf(dof_invweight(C, d) * (1.0f - imp) / imp, EPS);
        R.R[c] = diag; R.aref[c] = aref; R.D[c] = 1.0f / diag; R.floss[c] = 0.f;
        J[c * JS + d] = jac;
      }
      r0 += __popc(bal);
    }
  }
  C.sync();
}

// ---------------------------------------------------------------------------------------------------------------------
// Solver pieces
// ---------------------------------------------------------------------------------------------------------------------
// CS:2619-2694 func_update_constraint_batch
template <int G>
__device__ GSR_UPD_NOINLINE void update_constraint(const Ctx<G>& C, const Rows& R, SolverState& X, bool track_prev) {
  const int nD = CM(nD), lane = C.lane;
  X.prev_cost = X.cost;
  float cost = 0.f;
  for (int c = lane; c < R.n_con; c += G) {
    const float Ja = R.Jaref[c], D = R.D[c];
    const int prev = R.active[c] & 1;
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
    R.active[c] = act | (track_prev ? (prev << 1) : (act << 1));
    R.force[c] = floss_force + (-Ja * D * (float)act);
    cost += 0.5f * (Ja * Ja * D * (float)act);
  }
  C.sync();
  float gauss = 0.f;
  float* qfrc = WK(qfrc_c);
  for (int d = lane; d < nD; d += G) {
    float q = 0.f;
    for (int c = 0; c < R.n_con; c++) q += R.J[c * R.JS + d] * R.force[c];
    qfrc[d] = q;
    gauss += 0.5f * (WK(Ma)[d] - WK(force)[d]) * (WK(qacc)[d] - WK(acc_smooth)[d]);
  }
  X.gauss = C.sum(gauss);
  X.cost = C.sum(cost) + X.gauss;
  C.sync();
}

// CS:1394-1452 + CS:1594-1621: H = M + J^T (D active) J (lower), then in-place Cholesky (1 / L_ii kept in WK(Hinv))
template <int G>
__device__ GSR_HESS_NOINLINE void hessian_cholesky(const Ctx<G>& C, const Rows& R, long long* t_split = nullptr) {
  const int nD = CM(nD), MS = c_s.MS, lane = C.lane;
  const float EPS = CM(eps);
  float* H = WKS(H);
  float* Hinv = WKS(Hinv);
  const float* M = WKS(M);
  const int ntri = nD * (nD + 1) / 2;
  // every lane accumulates up to NE lower-triangle entries at once, so one sweep over the rows (with the row weight
  // D_c * active_c loaded once, inactive rows skipped by a group-uniform branch) serves NE * G entries
  constexpr int NE = 6;
  float* w = R.jv;  // scratch: the line search recomputes jv before it reads it
  for (int c = lane; c < R.n_con; c += G) w[c] = (R.active[c] & 1) ? R.D[c] : 0.f;
  C.sync();
  for (int t0 = 0; t0 < ntri; t0 += NE * G) {
    int oi[NE], oj[NE];
    float acc[NE];
#pragma unroll
    for (int e = 0; e < NE; e++) {
      const int t = t0 + e * G + lane;
      int i = 0, j = 0;
      if (t < ntri) tri_decode(t, i, j);
      oi[e] = i; oj[e] = j; acc[e] = 0.f;
    }
    // two rows per trip: 4 NE independent loads in flight before the first multiply (a row of weight 0 adds exact zeros, the
    // additions keep the row order), and the skip of inactive rows is decided per pair
    int c = 0;
    for (; c + 1 < R.n_con; c += 2) {
      const float w0 = w[c], w1 = w[c + 1];
      if (w0 == 0.f && w1 == 0.f) continue;
      const float *J0 = R.J + c * R.JS, *J1 = J0 + R.JS;
      float a0[NE], b0[NE], a1[NE], b1[NE];
#pragma unroll
      for (int e = 0; e < NE; e++) { a0[e] = J0[oj[e]]; b0[e] = J0[oi[e]]; a1[e] = J1[oj[e]]; b1[e] = J1[oi[e]]; }
#pragma unroll
      for (int e = 0; e < NE; e++) { acc[e] += a0[e] * (b0[e] * w0); acc[e] += a1[e] * (b1[e] * w1); }
    }
    if (c < R.n_con) {
      const float wc = w[c];
      const float* Jc = R.J + c * R.JS;
      if (wc != 0.f) {
#pragma unroll
        for (int e = 0; e < NE; e++) acc[e] += Jc[oj[e]] * (Jc[oi[e]] * wc);
      }
    }
#pragma unroll
    for (int e = 0; e < NE; e++) {
      const int t = t0 + e * G + lane;
      if (t < ntri) H[oi[e] * MS + oj[e]] = acc[e] + M[t];  // packed M in task order
    }
  }
  C.sync();
  if (t_split) *t_split = clock64();
  for (int i = 0; i < nD; i++) {
    // Column i: every lane owning a row j >= i forms dot_j = sum_{k<i} H[j][k] H[i][k] and, from the same pivot-row loads,
    // the pivot's own dot, so that L_ii and 1 / L_ii (one rsqrt) are known to every lane without a shuffle and the column
    // is scaled in the same pass; the sums run four terms at a time with a predicated tail (the additions keep their order).
    const int j0 = i + ((lane - i) & (G - 1));
    const unsigned busy = __ballot_sync(C.gmask, j0 < nD);
    if (j0 < nD) {
      const float* Hi = H + i * MS;
      const float hii = Hi[i];
      float dii = 0.f, dt0 = 0.f, dt1 = 0.f;
      const int j1 = j0 + G;
      const float *Hj0 = H + j0 * MS, *Hj1 = H + (j1 < nD ? j1 : j0) * MS;
      for (int k = 0; k < i; k += 4) {
        const bool p1 = k + 1 < i, p2 = k + 2 < i, p3 = k + 3 < i;
        const float b0 = Hi[k], b1 = p1 ? Hi[k + 1] : 0.f, b2 = p2 ? Hi[k + 2] : 0.f, b3 = p3 ? Hi[k + 3] : 0.f;
        const float a0 = Hj0[k], a1 = p1 ? Hj0[k + 1] : 0.f, a2 = p2 ? Hj0[k + 2] : 0.f, a3 = p3 ? Hj0[k + 3] : 0.f;
        dii += b0 * b0; dii += b1 * b1; dii += b2 * b2; dii += b3 * b3;
        dt0 += a0 * b0; dt0 += a1 * b1; dt0 += a2 * b2; dt0 += a3 * b3;
        if (G < 32 && j1 < nD) {
          const float c0 = Hj1[k], c1 = p1 ? Hj1[k + 1] : 0.f, c2 = p2 ? Hj1[k + 2] : 0.f, c3 = p3 ? Hj1[k + 3] : 0.f;
          dt1 += c0 * b0; dt1 += c1 * b1; dt1 += c2 * b2; dt1 += c3 * b3;
        }
      }
      const float x = fmaxf(hii - dii, EPS);
      const float inv = rsqrtf(x), Lii = x * inv;
      __syncwarp(busy);  // every lane has read the pivot before its owner overwrites it
      if (j0 == i) { H[i * MS + i] = Lii; Hinv[i] = inv; }
      else H[j0 * MS + i] = (H[j0 * MS + i] - dt0) * inv;
      for (int j = j1; j < nD; j += G) {  // further rows of this lane (nD > G): the second one is already summed
        float dt = dt1;
        if (j != j1 || G == 32) {
          dt = 0.f;
          for (int k = 0; k < i; k++) dt += H[j * MS + k] * Hi[k];
        }
        H[j * MS + i] = (H[j * MS + i] - dt) * inv;
      }
    }
    C.sync();
  }
}

// CS:1763-1822 rank-1 Cholesky up / down-dates for the rows whose active flag flipped; true when degenerate
template <int G>
__device__ __noinline__ bool hessian_incremental(const Ctx<G>& C, const Rows& R) {
  const int nD = CM(nD), MS = c_s.MS, lane = C.lane;
  const float EPS = CM(eps);
  float* H = WKS(H);
  float* vec = WK(prev_grad);  // scratch (the CG history is unused by the Newton solver)
  if (nD <= G) {
    // one matrix row per lane: lane i keeps vec[i] in a register and is the only one to touch row i of the factor, the
    // pivot values (vec[k], L_kk) travel by shuffle -- no shared-memory round trip and no barrier per column (the same
    // operations in the same order as the general loop below)
    float* Hinv = WKS(Hinv);
    const bool own = lane < nD;
    for (int base = 0; base < R.n_con; base += G) {
      const int cl = base + lane;
      const int a = cl < R.n_con ? R.active[cl] : 0;
      unsigned bal = C.ballot(((a ^ (a >> 1)) & 1) != 0);
      while (bal) {
        const int b = __ffs(bal) - 1;
        bal &= bal - 1;
        const int c = base + b;
        const float sign = (R.active[c] & 1) ? 1.f : -1.f;
        const float sq = sqrtf(R.D[c]);
        float v = own ? R.J[c * R.JS + lane] * sq : 0.f;
        // the loads of column k + 1 (own entry, pivot and its cached reciprocal; column k's update does not touch them) are
        // issued one column ahead, and sqrt / divisions collapse into one rsqrt, so that the dependent chain of a column is
        // shuffle -> fma -> rsqrt -> fma instead of shuffle -> load -> shuffle -> sqrt -> three divisions
        C.sync();  // the pivots written by other lanes in the previous update
        float hik = own ? H[lane * MS] : 0.f, Lkk = H[0], Linv = Hinv[0];
        bool bad = false;
        for (int k = 0; k < nD; k++) {
          const int kn = k + 1 < nD ? k + 1 : k;
          const float hik_n = (own && lane >= kn) ? H[lane * MS + kn] : 0.f, Lkk_n = H[kn * MS + kn], Linv_n = Hinv[kn];
          const float vk = __shfl_sync(C.gmask, v, k, G);
          const float tmp = Lkk * Lkk + sign * vk * vk;
          bad |= fabsf(vk) > EPS && tmp < EPS;
          const float rinv = rsqrtf(fmaxf(tmp, EPS)), r = tmp * rinv;
          const float cc = r * Linv, cinv = Lkk * rinv, s_ = vk * Linv;
          // every lane has read the pivot of column k + 1 (one column ahead, above) before its owner may overwrite it in the next
          // trip: the shuffle orders the instruction streams, this barrier makes the order part of the memory model (racecheck)
          C.sync();
          if (fabsf(vk) > EPS) {
            if (own && lane > k) {
              const float Lik = (hik + s_ * v * sign) * cinv;
              H[lane * MS + k] = Lik;
              v = v * cc - s_ * Lik;
            } else if (lane == k) {
              H[k * MS + k] = r;
              Hinv[k] = rinv;
            }
          }
          hik = hik_n; Lkk = Lkk_n; Linv = Linv_n;
        }
        if (bad) return true;  // degenerate down-date (group-uniform: tmp derives from shuffled values): the caller refactors
      }
    }
    C.sync();
    return false;
  }
  for (int base = 0; base < R.n_con; base += G) {
    const int cl = base + lane;
    const int a = cl < R.n_con ? R.active[cl] : 0;
    unsigned bal = C.ballot(((a ^ (a >> 1)) & 1) != 0);
    while (bal) {
      const int b = __ffs(bal) - 1;
      bal &= bal - 1;
      const int c = base + b;
      const float sign = (R.active[c] & 1) ? 1.f : -1.f;
      const float sq = sqrtf(R.D[c]);
      C.sync();
      for (int d = lane; d < nD; d += G) vec[d] = R.J[c * R.JS + d] * sq;
      for (int k = 0; k < nD; k++) {
        C.sync();
        const float vk = vec[k];
        if (fabsf(vk) > EPS) {
          const float Lkk = H[k * MS + k];
          const float tmp = Lkk * Lkk + sign * vk * vk;
          if (tmp < EPS) return true;
          const float r = sqrtf(tmp), cc = r / Lkk, cinv = 1.0f / cc, s = vk / Lkk;
          for (int i = k + 1 + lane; i < nD; i += G) {
            const float Lik = (H[i * MS + k] + s * vec[i] * sign) * cinv;
            H[i * MS + k] = Lik;
            vec[i] = vec[i] * cc - s * Lik;
          }
          C.sync();
          if (lane == 0) { H[k * MS + k] = r; WK(Hinv)[k] = 1.0f / r; }
        }
      }
    }
  }
  C.sync();
  return false;
}

// CS:2721-2750 gradient and preconditioned gradient (Newton: Cholesky solve CS:1892-1910; CG: M^-1 via L^T D L)
template <int G>
__device__ GSR_GRAD_NOINLINE void update_gradient(const Ctx<G>& C, bool solve = true) {
  const int nD = CM(nD), MS = c_s.MS, lane = C.lane;
  float* grad = WK(grad);
  float* Mgrad = WK(Mgrad);
  for (int d = lane; d < nD; d += G) { const float g = WK(Ma)[d] - WK(force)[d] - WK(qfrc_c)[d]; grad[d] = g; Mgrad[d] = g; }
  C.sync();
  if (!solve) return;  // the caller only needs the gradient (termination test); the preconditioned one follows if it goes on
  if (CM(solver_type) == SOLVER_CG) {
    mass_solve(C, Mgrad);
  } else {
    const float* H = WKS(H);
    const float* Hinv = WKS(Hinv);
    if (nD <= G) {
      // one lane per unknown with the running right-hand side in a register and the pivot broadcast by shuffle: no
      // shared-memory round trip or barrier per column (same products in the same order as the column sweeps below)
      const bool own = lane < nD;
      float s = own ? Mgrad[lane] : 0.f;
      const float hinv = own ? Hinv[lane] : 0.f;
      for (int i = 0; i < nD; i++) {  // L y = g
        const float yi = __shfl_sync(C.gmask, s * hinv, i, G);
        if (lane == i) s = yi;
        if (own && lane > i) s -= H[lane * MS + i] * yi;
      }
      for (int i = nD - 1; i >= 0; i--) {  // L^T x = y
        const float xi = __shfl_sync(C.gmask, s * hinv, i, G);
        if (lane == i) s = xi;
        if (lane < i) s -= H[i * MS + lane] * xi;
      }
      if (own) Mgrad[lane] = s;
      C.sync();
      return;
    }
    if (nD <= 2 * G) {
      // two unknowns per lane (rows lane and lane + G), G < nD <= 2 G
      const int r1 = lane + G;
      const bool own1 = r1 < nD;
      float s0 = Mgrad[lane], s1 = own1 ? Mgrad[r1] : 0.f;
      const float h0 = Hinv[lane], h1 = own1 ? Hinv[r1] : 0.f;
      for (int i = 0; i < nD; i++) {  // L y = g
        const float yi = __shfl_sync(C.gmask, i < G ? s0 * h0 : s1 * h1, i < G ? i : i - G, G);
        if (lane == i) s0 = yi;
        if (r1 == i) s1 = yi;
        if (lane > i) s0 -= H[lane * MS + i] * yi;
        if (own1 && r1 > i) s1 -= H[r1 * MS + i] * yi;
      }
      for (int i = nD - 1; i >= 0; i--) {  // L^T x = y
        const float xi = __shfl_sync(C.gmask, i < G ? s0 * h0 : s1 * h1, i < G ? i : i - G, G);
        if (lane == i) s0 = xi;
        if (r1 == i) s1 = xi;
        if (lane < i) s0 -= H[i * MS + lane] * xi;
        if (r1 < i) s1 -= H[i * MS + r1] * xi;
      }
      Mgrad[lane] = s0;
      if (own1) Mgrad[r1] = s1;
      C.sync();
      return;
    }
    float* y = WK(tmpv);
    for (int i = 0; i < nD; i++) {  // forward substitution, column oriented: y_i final once columns < i are eliminated
      const float yi = Mgrad[i] * Hinv[i];
      if (lane == 0) y[i] = yi;
      for (int j = i + 1 + lane; j < nD; j += G) Mgrad[j] -= H[j * MS + i] * yi;
      C.sync();
    }
    for (int i = nD - 1; i >= 0; i--) {  // backward substitution on y, result back into Mgrad
      const float xi = y[i] * Hinv[i];
      if (lane == 0) Mgrad[i] = xi;
      for (int j = lane; j < i; j += G) y[j] -= H[i * MS + j] * xi;
      C.sync();
    }
  }
}

struct LsPoint { float alpha, cost, d0, d1; };

// CS:2154-2354 the piecewise-quadratic cost and its derivatives at step `alpha`
template <int G>
__device__ __noinline__ LsPoint ls_eval(const Ctx<G>& C, const Rows& R, const SolverState& X, float alpha) {
  float t0 = 0.f, t1 = 0.f, t2 = 0.f;
  for (int c = R.ne + C.lane; c < R.n_con; c += G) {
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
  t0 = C.sum(t0) + X.qg0 + X.es0; t1 = C.sum(t1) + X.qg1 + X.es1; t2 = C.sum(t2) + X.qg2 + X.es2;
  LsPoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * t2 + alpha * t1 + t0;
  p.d0 = 2.f * alpha * t2 + t1;
  p.d1 = 2.f * t2;
  if (p.d1 <= 0.f) p.d1 = CM(eps);
  return p;
}
// the three candidates of a bracketing round (CS:2226-2354 func_ls_point_fn_3alphas_opt) in ONE sweep over the rows: the row
// data are loaded once and the nine butterfly reductions interleave, instead of three dependent evaluations back to back
// (the same additions in the same order per candidate, so the result is bit-identical to three ls_eval calls)
template <int G>
__device__ __noinline__ void ls_eval3(const Ctx<G>& C, const Rows& R, const SolverState& X, float a0, float a1, float a2, LsPoint* out) {
  const float al[3] = {a0, a1, a2};
  float t0[3] = {0.f, 0.f, 0.f}, t1[3] = {0.f, 0.f, 0.f}, t2[3] = {0.f, 0.f, 0.f};
  for (int c = R.ne + C.lane; c < R.n_con; c += G) {
    const float Ja = R.Jaref[c], jv = R.jv[c], D = R.D[c];
    const float h0 = 0.5f * Ja * Ja, h1 = jv * Ja, h2 = 0.5f * jv * jv;
    if (c < R.nef) {
      const float f = R.floss[c], rf = R.R[c] * f;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float x = Ja + al[k] * jv;
        const bool ln = x <= -rf, lp = x >= rf;
        if (ln || lp) {
          t0[k] += (ln ? f * (-0.5f * rf - Ja) : 0.f) + (lp ? f * (-0.5f * rf + Ja) : 0.f);
          t1[k] += (ln ? -f * jv : 0.f) + (lp ? f * jv : 0.f);
        } else {
          t0[k] += D * h0; t1[k] += D * h1; t2[k] += D * h2;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float x = Ja + al[k] * jv;
        if (x < 0.f) { t0[k] += D * h0; t1[k] += D * h1; t2[k] += D * h2; }
      }
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      t0[k] += __shfl_xor_sync(C.gmask, t0[k], o, G);
      t1[k] += __shfl_xor_sync(C.gmask, t1[k], o, G);
      t2[k] += __shfl_xor_sync(C.gmask, t2[k], o, G);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float s0 = t0[k] + X.qg0 + X.es0, s1 = t1[k] + X.qg1 + X.es1, s2 = t2[k] + X.qg2 + X.es2;
    LsPoint p;
    p.alpha = al[k];
    p.cost = al[k] * al[k] * s2 + al[k] * s1 + s0;
    p.d0 = 2.f * al[k] * s2 + s1;
    p.d1 = 2.f * s2;
    if (p.d1 <= 0.f) p.d1 = CM(eps);
    out[k] = p;
  }
}
// CS:2357-2388
__device__ __forceinline__ int update_bracket(LsPoint& p, const LsPoint* cand, float& next_alpha) {
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
template <int G>
__device__ float linesearch(const Ctx<G>& C, const Rows& R, SolverState& X) {
  const int nD = CM(nD), MS = c_s.MS, lane = C.lane;
  const float EPS = CM(eps);
  const float* search = WK(search);
  const float* M = WK(M);
  float* mv = WK(mv);
  float sn = 0.f, qg1 = 0.f, qg2 = 0.f;
  for (int i = lane; i < nD; i += G) {
    const float s = search[i];
    sn += s * s;
    const float m = sym_row_dot(M, i, search, nD);
    mv[i] = m;
    qg1 += s * WK(Ma)[i] - s * WK(force)[i];
    qg2 += 0.5f * s * m;
  }
  const float snorm = sqrtf(C.sum(sn));
  const float gtol = CM(tolerance) * CM(ls_tolerance) * snorm * (CM(meaninertia) * (float)max(1, nD));
  X.ls_it = 0;
  if (snorm < EPS) return 0.f;
  X.qg0 = X.gauss; X.qg1 = C.sum(qg1); X.qg2 = C.sum(qg2);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < R.n_con; c += G) {
    float jv = 0.f;
    for (int d = 0; d < nD; d++) jv += R.J[c * R.JS + d] * search[d];
    R.jv[c] = jv;
    if (c < R.ne) {
      const float Ja = R.Jaref[c], D = R.D[c];
      s0 += D * (0.5f * Ja * Ja); s1 += D * (jv * Ja); s2 += D * (0.5f * jv * jv);
    }
  }
  X.es0 = C.sum(s0); X.es1 = C.sum(s1); X.es2 = C.sum(s2);
  C.sync();
  const int ls_iterations = CM(ls_iterations);
  const LsPoint p0 = ls_eval(C, R, X, 0.f);
  X.ls_it = 1;
  LsPoint p1 = ls_eval(C, R, X, p0.alpha - p0.d0 / p0.d1);
  X.ls_it++;
  if (p0.cost < p1.cost) p1 = p0;
  if (fabsf(p1.d0) < gtol) return p1.alpha;
  const float direction = p1.d0 < 0.f ? 1.f : -1.f;
  bool p2update = false;
  LsPoint p2 = p1;
  while (p1.d0 * direction <= -gtol && X.ls_it < ls_iterations) {
    p2 = p1; p2update = true;
    p1 = ls_eval(C, R, X, p1.alpha - p1.d0 / p1.d1);
    X.ls_it++;
    if (fabsf(p1.d0) < gtol) return p1.alpha;
  }
  if (X.ls_it >= ls_iterations) return p1.alpha;
  if (!p2update) return p1.alpha;
  float a0 = p1.alpha - p1.d0 / p1.d1, a1 = p1.alpha, a2 = (p1.alpha + p2.alpha) * 0.5f;
  while (X.ls_it < ls_iterations) {
    LsPoint cand[3];
    ls_eval3(C, R, X, a0, a1, a2, cand);
    X.ls_it += 3;
    float best_alpha = 0.f, best_cost = 0.f;
    bool found = false;
#pragma unroll
    for (int i = 0; i < 3; i++)
      if (fabsf(cand[i].d0) < gtol && (!found || cand[i].cost < best_cost)) { best_alpha = cand[i].alpha; best_cost = cand[i].cost; found = true; }
    if (found) return best_alpha;
    float p1n, p2n;
    const int b1 = update_bracket(p1, cand, p1n), b2 = update_bracket(p2, cand, p2n);
    if (b1 == 0 && b2 == 0) return a2;
    a0 = p1n; a1 = p2n; a2 = (p1.alpha + p2.alpha) * 0.5f;
  }
  if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
  if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
  return 0.f;
}

// ---------------------------------------------------------------------------------------------------------------------
// constraint/noslip.py: post-solver that removes residual slip.  Projected Gauss-Seidel over the friction-loss rows and over
// the two opposing pyramid edges of each tangent direction with the regulariser dropped (A = J M^-1 J^T).  The reference
// materialises A (n_con^2 words per env); here only the rows M^-1 J^T are kept (SPILL record) and q = J^T f is updated
// incrementally, so A f = (M^-1 J^T)^T q and a 2x2 block of A is four dot products.
// ---------------------------------------------------------------------------------------------------------------------
template <int G>
__device__ __noinline__ void phase_noslip(const Ctx<G>& C, const Rows& R, int n_contacts) {
  const int nD = CM(nD), MS = c_s.MS, JS = R.JS, n_con = R.n_con, lane = C.lane;
  const float EPS = CM(eps);
  float* MJ = C.spill + c_h.spill_minvjt;
  float* b = R.jv;         // dead after the primal solve
  float* q = WK(grad);     // J^T f
  const float* L = WK(L);
  const float* Dinv = WK(Dinv);
  const float* J = R.J;
  // noslip.py:9-58: rows of M^-1 J^T (one lane per row: serial L^T D L substitution on its own row) and b = J acc_smooth - aref
  for (int r = lane; r < n_con; r += G) {
    float* x = MJ + (size_t)r * JS;
    for (int d = 0; d < nD; d++) x[d] = J[r * JS + d];
    for (int j = nD - 1; j > 0; j--) {
      const float xj = x[j];
      for (int i = 0; i < j; i++) x[i] -= L[j * MS + i] * xj;
    }
    for (int i = 0; i < nD; i++) x[i] *= Dinv[i];
    for (int j = 0; j < nD - 1; j++) {
      const float xj = x[j];
      for (int i = j + 1; i < nD; i++) x[i] -= L[i * MS + j] * xj;
    }
    float v = -R.aref[r];
    for (int d = 0; d < nD; d++) v += J[r * JS + d] * WK(acc_smooth)[d];
    b[r] = v;
  }
  for (int d = lane; d < nD; d += G) q[d] = WK(qfrc_c)[d];
  __threadfence_block();
  C.sync();
  const float scale = 1.0f / (CM(meaninertia) * (float)max(1, nD));
  const int const_start = R.nef;
  for (int it = 0; it < CM(noslip_iterations); it++) {  // noslip.py:166-283
    float improvement = 0.f;
    if (it == 0) {
      float s = 0.f;
      for (int c = lane; c < n_con; c += G) s += 0.5f * R.force[c] * R.force[c] * R.R[c];
      improvement = C.sum(s);
    }
    for (int c = R.ne; c < R.nef; c++) {
      float r0 = 0.f, a0 = 0.f;
      for (int d = lane; d < nD; d += G) { const float mj = MJ[(size_t)c * JS + d]; r0 += mj * q[d]; a0 += J[c * JS + d] * mj; }
      const float res = b[c] + C.sum(r0), Acc = C.sum(a0);
      const float old = R.force[c], fl = R.floss[c];
      float f = old - res / Acc;
      f = f < -fl ? -fl : (f > fl ? fl : f);
      const float delta = f - old;
      improvement -= 0.5f * delta * delta / Acc + delta * res;
      C.sync();
      if (lane == 0) R.force[c] = f;
      for (int d = lane; d < nD; d += G) q[d] += J[c * JS + d] * delta;
      C.sync();
    }
    for (int ic = 0; ic < n_contacts; ic++) {
      for (int j2 = 0; j2 < 2; j2++) {
        const int j = const_start + 4 * ic + 2 * j2;
        float r0 = 0.f, r1 = 0.f, a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
        for (int d = lane; d < nD; d += G) {
          const float m0 = MJ[(size_t)j * JS + d], m1 = MJ[(size_t)(j + 1) * JS + d], j0 = J[j * JS + d], j1 = J[(j + 1) * JS + d], qd = q[d];
          r0 += m0 * qd; r1 += m1 * qd; a00 += j0 * m0; a01 += j1 * m0; a10 += j0 * m1; a11 += j1 * m1;
        }
        const float res0 = b[j] + C.sum(r0), res1 = b[j + 1] + C.sum(r1);
        a00 = C.sum(a00); a01 = C.sum(a01); a10 = C.sum(a10); a11 = C.sum(a11);
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
        // noslip.py:358-385 func_cost_change: keep the update only if it does not raise the cost
        float d0 = f0 - o0, d1 = f1 - o1;
        float change = 0.5f * (a00 * d0 * d0 + a01 * d0 * d1 + a10 * d1 * d0 + a11 * d1 * d1) + d0 * res0 + d1 * res1;
        if (change > EPS) { f0 = o0; f1 = o1; d0 = 0.f; d1 = 0.f; change = 0.f; }
        improvement -= change;
        C.sync();
        if (lane == 0) { R.force[j] = f0; R.force[j + 1] = f1; }
        for (int d = lane; d < nD; d += G) q[d] += J[j * JS + d] * d0 + J[(j + 1) * JS + d] * d1;
        C.sync();
      }
    }
    improvement *= scale;
    if (improvement < CM(noslip_tolerance)) break;
  }
  // noslip.py:286-323 kernel_dual_finish: qacc = acc_smooth + M^-1 J^T f
  float* x = WK(tmpv);
  for (int d = lane; d < nD; d += G) { const float v = q[d]; WK(qfrc_c)[d] = v; x[d] = v; }
  C.sync();
  mass_solve(C, x);
  for (int d = lane; d < nD; d += G) { const float a = x[d] + WK(acc_smooth)[d]; WK(qacc)[d] = a; WK(acc)[d] = a; }
  C.sync();
}

// CS:210-245 resolve(): init (CS:2935-3059), monolith body (CS:3163-3190), update_qacc (CS:3240-3261),
// contact forces (CS:3198-3237)
#ifdef GSR_SOLVER_TRACE  // developer build (make trace): SM cycles per solver stage and counters, per environment
__device__ unsigned long long g_solver_dbg[16 * 65536];
#define DBG_T(k) { const long long t_ = clock64(); dbg[k] += t_ - dbg_t; dbg_t = t_; }
#define DBG_N(k, n) dbg[k] += (n)
#else
#define DBG_T(k)
#define DBG_N(k, n)
#endif
template <int G>
__device__ void phase_constraints(const Ctx<G>& C) {
#ifdef GSR_SOLVER_TRACE
  long long dbg[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, dbg_t = clock64();
#endif
  const int nD = CM(nD), nL = CM(nL), MS = c_s.MS, lane = C.lane;
  const float EPS = CM(eps);
  Rows R;
  build_rows(C, R);
  DBG_T(0);
  const int n_con = R.n_con;
  int* SI = C.SI();
  SolverState X;
  X.cost = 0.f; X.prev_cost = 0.f; X.gauss = 0.f; X.ls_it = 0;
  float* qacc = WK(qacc);
  float* Ma = WK(Ma);
  float* search = WK(search);
  const bool warm = n_con > 0 && SI[c_h.is_warmstart];
  for (int d = lane; d < nD; d += G) qacc[d] = warm ? IN(qacc_ws)[d] : WK(acc_smooth)[d];
  C.sync();
  int iters = 0;
  if (n_con > 0) {
    const float* M = WK(M);
    for (int i = lane; i < nD; i += G) {  // CS:2910-2929
      float s = 0.f;
      Ma[i] = sym_row_dot(M, i, qacc, nD);
    }
    for (int c = lane; c < n_con; c += G) {  // CS:2887-2907
      float s = -R.aref[c];
      for (int d = 0; d < nD; d++) s += R.J[c * R.JS + d] * qacc[d];
      R.Jaref[c] = s;
      R.active[c] = 0;
    }
    C.sync();
    DBG_T(13);
    update_constraint(C, R, X, false);
    DBG_T(14);
#ifdef GSR_SOLVER_TRACE
    long long t_split = 0;
    if (CM(solver_type) == SOLVER_NEWTON) hessian_cholesky(C, R, &t_split);
    dbg[15] += t_split - dbg_t; dbg_t = t_split;  // assembly; the factorisation goes to [5] below
    DBG_T(5);
#else
    if (CM(solver_type) == SOLVER_NEWTON) hessian_cholesky(C, R);
#endif
    update_gradient(C);
    for (int d = lane; d < nD; d += G) search[d] = -WK(Mgrad)[d];
    C.sync();
    DBG_T(1);
    const float tol_scaled = (CM(meaninertia) * (float)max(1, nD)) * CM(tolerance);
    for (int it = 0; it < CM(iterations); it++) {  // CS:3062-3139 func_solve_iter
      iters++;
      const float alpha = linesearch(C, R, X);
      DBG_T(2); DBG_N(8, X.ls_it); DBG_N(9, 1);
      if (fabsf(alpha) < EPS) break;
      for (int d = lane; d < nD; d += G) {
        qacc[d] += search[d] * alpha;
        Ma[d] += WK(mv)[d] * alpha;
        if (CM(solver_type) == SOLVER_CG) { WK(prev_grad)[d] = WK(grad)[d]; WK(prev_Mgrad)[d] = WK(Mgrad)[d]; }
      }
      for (int c = lane; c < n_con; c += G) R.Jaref[c] += R.jv[c] * alpha;
      C.sync();
      update_constraint(C, R, X, true);
      DBG_T(3);
      // CS:2841-2884 termination test first: it needs the gradient only (Ma, force, J^T f) -- not the Hessian.  The reference
      // updates the factor and solves for the preconditioned gradient before it tests; on the last iteration of every solve
      // (for most environments: the only one) that work is never used, so it is done here only when the iteration goes on.
      update_gradient(C, false);
      float gn = 0.f;
      for (int d = lane; d < nD; d += G) { const float g = WK(grad)[d]; gn += g * g; }
      gn = sqrtf(C.sum(gn));
      const bool improved = gn > tol_scaled && (X.prev_cost - X.cost) > tol_scaled;
      DBG_T(6);
      if (!improved) break;
      if (CM(solver_type) == SOLVER_NEWTON) {
#ifdef GSR_SOLVER_TRACE
        for (int c = 0; c < n_con; c++) { const int a = R.active[c]; DBG_N(10, (a ^ (a >> 1)) & 1); }
#endif
        if (hessian_incremental(C, R)) { DBG_T(4); DBG_N(11, 1); C.sync(); hessian_cholesky(C, R); DBG_T(5); }
        else DBG_T(4);
      }
      update_gradient(C);
      if (CM(solver_type) == SOLVER_NEWTON) {
        for (int d = lane; d < nD; d += G) search[d] = -WK(Mgrad)[d];
      } else {
        float beta = 0.f, pg = 0.f;
        for (int d = lane; d < nD; d += G) {
          beta += WK(grad)[d] * (WK(Mgrad)[d] - WK(prev_Mgrad)[d]);
          pg += WK(prev_Mgrad)[d] * WK(prev_grad)[d];
        }
        beta = fmaxf(C.sum(beta) / fmaxf(EPS, C.sum(pg)), 0.f);
        for (int d = lane; d < nD; d += G) search[d] = -WK(Mgrad)[d] + beta * search[d];
      }
      C.sync();
    }
  } else {
    for (int d = lane; d < nD; d += G) WK(qfrc_c)[d] = 0.f;
  }
  C.sync();
#ifdef GSR_SOLVER_TRACE
  DBG_T(7);
  dbg[12] = n_con;
  if (lane == 0 && C.env < 65536) for (int k = 0; k < 16; k++) g_solver_dbg[16 * C.env + k] = (unsigned long long)dbg[k];
#endif
  // ---- CS:3240-3261 func_update_qacc
  bool has_nan = false;
  float* acc = WK(acc);
  for (int d = lane; d < nD; d += G) {
    const float a = qacc[d];
    acc[d] = a;
    IN(qacc_ws)[d] = a;
    has_nan |= isnan(a);
  }
  if (C.any(has_nan) && lane == 0) SI[c_h.errno_] |= ERR_NAN_FORCE;
  if (lane == 0) {
    SI[c_h.is_warmstart] = 1;
    int* out = reinterpret_cast<int*>(C.rec);
    out[c_h.n_constraints] = n_con; out[c_h.n_constraints_equality] = R.ne; out[c_h.n_constraints_frictionloss] = R.nef - R.ne;
    out[c_h.solver_iters] = iters;
  }
  const int nc = (CM(enable_collision) && !CM(disable_constraint)) ? SI[c_h.n_contacts] : 0;
  if (CM(noslip_iterations) > 0 && n_con > 0) { C.sync(); phase_noslip(C, R, nc); }  // CS:237-238
  // ---- CS:3198-3237 func_update_contact_force: per-contact force (global record) and per-link totals
  for (int ic = lane; ic < nc; ic += G) {
    const float* ct = R.ct + (size_t)ic * GSR_CONTACT_WORDS;
    const f3 cn(ct[CT_NRM], ct[CT_NRM + 1], ct[CT_NRM + 2]);
    const float mu = ct[CT_MU];
    f3 d1, d2;
    orthogonals(cn, d1, d2);
    const int r0 = R.nef + 4 * ic;
    const f3 force = (d1 * (-mu) - cn) * R.force[r0] + (d1 * mu - cn) * R.force[r0 + 1] + (d2 * (-mu) - cn) * R.force[r0 + 2] + (d2 * mu - cn) * R.force[r0 + 3];
    float* g = C.contacts + (size_t)ic * GSR_CONTACT_WORDS;
    g[CT_FRC] = force.x; g[CT_FRC + 1] = force.y; g[CT_FRC + 2] = force.z;
    if (R.ct != C.contacts) { float* s = const_cast<float*>(ct); s[CT_FRC] = force.x; s[CT_FRC + 1] = force.y; s[CT_FRC + 2] = force.z; }
  }
  C.sync();
  float* lcf = C.rec + c_h.links_contact_force;
  for (int l = lane; l < nL; l += G) {
    f3 s;
    for (int ic = 0; ic < nc; ic++) {
      const float* ct = R.ct + (size_t)ic * GSR_CONTACT_WORDS;
      const int la = __float_as_int(ct[CT_LA]), lb = __float_as_int(ct[CT_LB]);
      const f3 f(ct[CT_FRC], ct[CT_FRC + 1], ct[CT_FRC + 2]);
      if (la == l) s -= f;
      if (lb == l) s += f;
    }
    st3(lcf, l, s);
  }
  C.sync();
}
