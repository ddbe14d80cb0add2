// =====================================================================================================
// TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the reference's batched rigid-body step.
//
// This file restates, one environment at a time and in the reference's own order of operations, the
// algorithm of Genesis' rigid solver (reference = /root/reference @ genesis-world 0.4.5).  It is compiled
// twice (float / double) by oracle/Makefile into oracle/_build/liboracle.so and is used ONLY by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker or as the
// timed CPU baseline.  The product path (genesis_b200/csrc) never includes, links or calls it.
//
// PARITY STATUS: **parity unpinned by reference data** -- the reference stores no golden vectors for this
// path (its tests consult MuJoCo live) and neither Quadrants nor MuJoCo can be installed here, so this
// restatement is pinned only by the reference's *analytic* test assertions re-expressed in tests/
// (mass-matrix factorisation identity, closed-form pendulum acceleration, contact force == weight, PD
// tracking, quaternion identities); see DESIGN.md "Oracle".
//
// Citations are file:line below /root/reference/genesis/ :
//   FK  = engine/solvers/rigid/abd/forward_kinematics.py      FD = engine/solvers/rigid/abd/forward_dynamics.py
//   CS  = engine/solvers/rigid/constraint/solver.py            BP = engine/solvers/rigid/collider/broadphase.py
//   NP  = engine/solvers/rigid/collider/narrowphase.py         MPR= engine/solvers/rigid/collider/mpr.py
//   SF  = engine/solvers/rigid/collider/support_field.py       CT = engine/solvers/rigid/collider/contact.py
//   BOX = engine/solvers/rigid/collider/box_contact.py         RS = engine/solvers/rigid/rigid_solver.py
// =====================================================================================================
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <limits>
#include <thread>
#include "vecmath.hpp"
#include "box_box.hpp"

namespace orc {

enum { GEOM_PLANE = 0, GEOM_SPHERE = 1, GEOM_ELLIPSOID = 2, GEOM_CYLINDER = 3, GEOM_CAPSULE = 4, GEOM_BOX = 5, GEOM_MESH = 6, GEOM_TERRAIN = 7 };
enum { JOINT_FIXED = 0, JOINT_REVOLUTE = 1, JOINT_PRISMATIC = 2, JOINT_SPHERICAL = 3, JOINT_FREE = 4 };
enum { EQ_CONNECT = 0, EQ_WELD = 1, EQ_JOINT = 2 };
enum { CTRL_POSITION = 0, CTRL_VELOCITY = 1, CTRL_FORCE = 2 };
enum { INTEGRATOR_EULER = 0, INTEGRATOR_IMPLICITFAST = 1, INTEGRATOR_APPROX_IMPLICITFAST = 2 };
enum { SOLVER_CG = 0, SOLVER_NEWTON = 1 };
enum { ERR_BROAD_OVERFLOW = 1, ERR_CONTACT_OVERFLOW = 2, ERR_NAN_FORCE = 8, ERR_NAN_ACC = 16 };

struct RawModel {
  std::map<std::string, std::vector<double>> f;
  std::map<std::string, std::vector<int>> i;
  std::map<std::string, double> s;
};

template <typename T>
struct Model {
  int nL = 0, nJ = 0, nD = 0, nQ = 0, nG = 0, nE = 0, nEq = 0, nP = 0, nV = 0;
  int max_contact_pairs = 0, max_broad_pairs = 0, n_contacts_per_pair = 5, len_constraints = 0;
  std::vector<int> links_parent, links_root, links_entity, links_is_fixed, links_joint_start, links_joint_end, links_dof_start,
      links_dof_end, links_q_start, links_q_end, links_geom_start, links_geom_end;
  std::vector<T> links_pos, links_quat, links_inertial_pos, links_inertial_quat, links_inertial_i, links_inertial_mass, links_invweight;
  std::vector<int> joints_type, joints_link, joints_q_start, joints_dof_start, joints_n_dofs;
  std::vector<T> joints_pos, joints_sol_params;
  std::vector<T> dofs_motion_ang, dofs_motion_vel, dofs_limit, dofs_stiffness, dofs_damping, dofs_armature, dofs_frictionloss, dofs_kp,
      dofs_act_bias, dofs_force_range, dofs_invweight, qpos0, mass_parent_mask;
  std::vector<int> dofs_link, dofs_joint, dofs_entity;
  std::vector<int> entities_link_start, entities_link_end, entities_dof_start, entities_dof_end, entities_geom_start, entities_geom_end;
  std::vector<T> entities_gravity_compensation;
  std::vector<int> geoms_link, geoms_type, geoms_is_fixed, geoms_vert_start, geoms_vert_end, geoms_support_start, support_vid;
  std::vector<T> geoms_pos, geoms_quat, geoms_data, geoms_friction, geoms_sol_params, geoms_center, geoms_init_aabb, verts, support_v;
  std::vector<int> eq_type, eq_obj1, eq_obj2;
  std::vector<T> eq_data, eq_sol_params;
  std::vector<int> pairs, pair_idx;
  // height-field terrain (collider.py:541-561): heights already scaled (rows x cols), grid shape, (horizontal, vertical) scale,
  // [x_max, y_max, z_max, x_min, y_min, z_min - 1] in the terrain's frame
  std::vector<T> terrain_hf, terrain_scale, terrain_xyz_maxmin;
  std::vector<int> terrain_rc;
  T meaninertia = 1;

  static std::vector<T> cv(const std::vector<double>& v) { return std::vector<T>(v.begin(), v.end()); }
  void load(const RawModel& r) {
    auto F = [&](const char* k) { auto it = r.f.find(k); return it == r.f.end() ? std::vector<T>() : cv(it->second); };
    auto I = [&](const char* k) { auto it = r.i.find(k); return it == r.i.end() ? std::vector<int>() : it->second; };
    auto S = [&](const char* k, double d) { auto it = r.s.find(k); return it == r.s.end() ? d : it->second; };
    nL = (int)S("n_links", 0); nJ = (int)S("n_joints", 0); nD = (int)S("n_dofs", 0); nQ = (int)S("n_qs", 0);
    nG = (int)S("n_geoms", 0); nE = (int)S("n_entities", 0); nEq = (int)S("n_equalities", 0); nP = (int)S("n_pairs", 0);
    nV = (int)S("n_verts", 0);
    max_contact_pairs = (int)S("max_contact_pairs", 0); max_broad_pairs = (int)S("max_broad_pairs", 0);
    n_contacts_per_pair = (int)S("n_contacts_per_pair", 5); len_constraints = (int)S("len_constraints", 0);
    meaninertia = (T)S("meaninertia", 1.0);
#define LI(x) x = I(#x)
#define LF(x) x = F(#x)
    LI(links_parent); LI(links_root); LI(links_entity); LI(links_is_fixed); LI(links_joint_start); LI(links_joint_end);
    LI(links_dof_start); LI(links_dof_end); LI(links_q_start); LI(links_q_end); LI(links_geom_start); LI(links_geom_end);
    LF(links_pos); LF(links_quat); LF(links_inertial_pos); LF(links_inertial_quat); LF(links_inertial_i); LF(links_inertial_mass);
    LF(links_invweight);
    LI(joints_type); LI(joints_link); LI(joints_q_start); LI(joints_dof_start); LI(joints_n_dofs); LF(joints_pos); LF(joints_sol_params);
    LF(dofs_motion_ang); LF(dofs_motion_vel); LF(dofs_limit); LF(dofs_stiffness); LF(dofs_damping); LF(dofs_armature);
    LF(dofs_frictionloss); LF(dofs_kp); LF(dofs_act_bias); LF(dofs_force_range); LF(dofs_invweight); LF(qpos0); LF(mass_parent_mask);
    LI(dofs_link); LI(dofs_joint); LI(dofs_entity);
    LI(entities_link_start); LI(entities_link_end); LI(entities_dof_start); LI(entities_dof_end); LI(entities_geom_start);
    LI(entities_geom_end); LF(entities_gravity_compensation);
    LI(geoms_link); LI(geoms_type); LI(geoms_is_fixed); LI(geoms_vert_start); LI(geoms_vert_end); LI(geoms_support_start);
    LF(geoms_pos); LF(geoms_quat); LF(geoms_data); LF(geoms_friction); LF(geoms_sol_params); LF(geoms_center); LF(geoms_init_aabb);
    LF(verts); LF(support_v);
    if (r.i.count("support_vid")) support_vid = r.i.at("support_vid");
    LI(eq_type); LI(eq_obj1); LI(eq_obj2); LF(eq_data); LF(eq_sol_params);
    LI(pairs); LI(pair_idx);
    LF(terrain_hf); LF(terrain_scale); LF(terrain_xyz_maxmin); LI(terrain_rc);
#undef LI
#undef LF
  }
  V3<T> v3(const std::vector<T>& a, int i) const { return V3<T>(a[3 * i], a[3 * i + 1], a[3 * i + 2]); }
  Q4<T> q4(const std::vector<T>& a, int i) const { return Q4<T>(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]); }
};

struct Config {
  double dt = 0.01;
  double gravity[3] = {0, 0, -9.81};
  int iterations = 50, ls_iterations = 50;
  double tolerance = 1e-6, ls_tolerance = 1e-2;
  int integrator = INTEGRATOR_APPROX_IMPLICITFAST, solver_type = SOLVER_NEWTON;
  int enable_collision = 1, enable_joint_limit = 1, enable_multi_contact = 1, sort_contacts = 1;
  double eps = 1.1920929e-07;  // gs.EPS for fp32 (genesis/__init__.py:229)
  double mc_perturbation = 1e-2, mc_tolerance = 1e-2, mpr_to_gjk_overlap_ratio = 0.25;  // collider.py:85-87
  double ccd_eps = 1e-9, ccd_tolerance = 1e-6;  // mpr.py:13-19
  int ccd_iterations = 50;
  int broadphase_sap = 0;     // 1: sweep-and-prune traversal of the CPU backend (BP:142-396); 0: all-vs-all over the valid-pair table (BP:400-460)
  int gjk_fallback = 1;       // exact-penetration fallback after MPR (NP:726-733); 0 keeps MPR's answer (A/B tests)
  int use_gjk_collision = 0;  // options/solvers.py use_gjk_collision: convex pairs through GJK + EPA, no MPR (NP:604-607, 736-850)
  int box_box_detection = 0;  // options/solvers.py:511: box-box pairs through func_box_box_contact (BOX:100-785) instead of MPR
  int noslip_iterations = 0;  // options/solvers.py:506-507
  double noslip_tolerance = 1e-6;
};

template <typename T>
struct Contact {
  int geom_a, geom_b, link_a, link_b, pair_idx;
  V3<T> normal, pos, force;
  T penetration, friction;
  T sol_params[7];
};

template <typename T>
struct Env {
  typedef V3<T> V;
  // generalized state
  std::vector<T> qpos, qpos_next, vel, vel_next, acc, acc_smooth, force, qf_bias, qf_passive, qf_applied, qf_smooth, qf_constraint, dofs_pos;
  std::vector<int> ctrl_mode;
  std::vector<T> ctrl_pos, ctrl_vel, ctrl_force;
  std::vector<V> cdof_ang, cdof_vel, cdofvel_ang, cdofvel_vel, cdofd_ang, cdofd_vel, f_ang, f_vel;
  // links
  std::vector<V> l_pos, i_pos, i_pos_bw, root_COM, root_COM_bw, j_pos, cinr_pos, crb_pos, cd_vel, cd_ang, cdd_vel, cdd_ang, cacc_lin,
      cacc_ang, cfrc_vel, cfrc_ang, cfrc_applied_vel, cfrc_applied_ang, contact_force;
  std::vector<Q4<T>> l_quat, i_quat, j_quat;
  std::vector<M3<T>> cinr_I, crb_I;
  std::vector<T> mass_sum, cinr_mass, crb_mass;
  struct DynWeld { int l1, l2; T data[11]; T sol[7]; };
  std::vector<DynWeld> dyn_welds;   // add_weld_constraint (solver.py:1293-1340): per-env equalities after the static ones
  std::vector<T> mass_shift;        // links_state.mass_shift (domain randomisation, accessor.py:415)
  std::vector<V3<T>> i_pos_shift;   // links_state.i_pos_shift (accessor.py:429)
  std::vector<V> xanchor, xaxis;
  // geoms
  std::vector<V> g_pos, aabb_min, aabb_max;
  std::vector<Q4<T>> g_quat;
  std::vector<T> friction_ratio;
  // mass matrix
  std::vector<T> M, L, Dinv;
  // collider
  std::vector<V> contact_cache_normal;
  std::vector<int> broad_pairs;
  int n_broad = 0;
  std::vector<T> sap_value; std::vector<int> sap_geom; std::vector<uint8_t> sap_is_max;  // persistent SAP sort buffer
  std::vector<Contact<T>> contacts;
  int n_contacts = 0;
  // constraints
  std::vector<T> jac, diag, aref, efc_D, efc_force, efc_frictionloss, Jaref, jv;
  std::vector<uint8_t> active, prev_active;
  int n_con = 0, n_eq = 0, n_floss = 0;
  std::vector<T> qacc, qacc_ws, Ma, grad, Mgrad, search, mv, H, nt_vec, cg_prev_grad, cg_prev_Mgrad;
  int is_warmstart = 0, improved = 0, ls_it = 0, ls_result = 0, n_iter = 0;
  T cost = 0, prev_cost = 0, gauss = 0, quad_gauss[3] = {0, 0, 0}, eq_sum[3] = {0, 0, 0};
  int errno_ = 0;
};

template <typename T>
struct Oracle {
  typedef V3<T> V;
  typedef Q4<T> Q;
  Model<T> m;
  Config cfg;
  std::vector<Env<T>> envs;
  T EPS, dt;
  V gravity;

  void init(int B) {
    EPS = (T)cfg.eps;
    dt = (T)cfg.dt;
    gravity = V((T)cfg.gravity[0], (T)cfg.gravity[1], (T)cfg.gravity[2]);
    envs.assign(B, Env<T>());
    for (auto& e : envs) alloc(e);
    for (auto& e : envs) {
      e.qpos = m.qpos0;
      reset_env(e);
      init_fixed_geoms(e);
      forward_all(e);
    }
  }

  void alloc(Env<T>& e) {
    int nD = m.nD, nL = m.nL, nG = m.nG, nC = std::max(1, m.len_constraints);
    for (auto* v : {&e.qpos, &e.qpos_next}) v->assign(m.nQ, 0);
    for (auto* v : {&e.vel, &e.vel_next, &e.acc, &e.acc_smooth, &e.force, &e.qf_bias, &e.qf_passive, &e.qf_applied, &e.qf_smooth,
                    &e.qf_constraint, &e.dofs_pos, &e.ctrl_pos, &e.ctrl_vel, &e.ctrl_force, &e.Dinv, &e.qacc, &e.qacc_ws, &e.Ma, &e.grad,
                    &e.Mgrad, &e.search, &e.mv, &e.nt_vec, &e.cg_prev_grad, &e.cg_prev_Mgrad})
      v->assign(nD, 0);
    e.ctrl_mode.assign(nD, CTRL_FORCE);  // abd/misc.py:170-173
    for (auto* v : {&e.cdof_ang, &e.cdof_vel, &e.cdofvel_ang, &e.cdofvel_vel, &e.cdofd_ang, &e.cdofd_vel, &e.f_ang, &e.f_vel}) v->assign(nD, V());
    for (auto* v : {&e.l_pos, &e.i_pos, &e.i_pos_bw, &e.root_COM, &e.root_COM_bw, &e.j_pos, &e.cinr_pos, &e.crb_pos, &e.cd_vel, &e.cd_ang,
                    &e.cdd_vel, &e.cdd_ang, &e.cacc_lin, &e.cacc_ang, &e.cfrc_vel, &e.cfrc_ang, &e.cfrc_applied_vel, &e.cfrc_applied_ang,
                    &e.contact_force})
      v->assign(nL, V());
    for (auto* v : {&e.l_quat, &e.i_quat, &e.j_quat}) v->assign(nL, Q());
    e.cinr_I.assign(nL, M3<T>()); e.crb_I.assign(nL, M3<T>());
    e.mass_sum.assign(nL, 0); e.cinr_mass.assign(nL, 0); e.crb_mass.assign(nL, 0);
    e.mass_shift.assign(nL, 0); e.i_pos_shift.assign(nL, V());
    e.xanchor.assign(m.nJ, V()); e.xaxis.assign(m.nJ, V());
    e.g_pos.assign(nG, V()); e.aabb_min.assign(nG, V()); e.aabb_max.assign(nG, V()); e.g_quat.assign(nG, Q());
    e.friction_ratio.assign(nG, 1);
    e.M.assign(nD * nD, 0); e.L.assign(nD * nD, 0); e.H.assign(nD * nD, 0);
    e.contact_cache_normal.assign(std::max(1, m.nP), V());
    e.broad_pairs.assign(2 * std::max(1, m.max_broad_pairs), 0);
    e.contacts.assign(std::max(1, m.max_contact_pairs), Contact<T>());
    e.jac.assign((size_t)nC * nD, 0);
    for (auto* v : {&e.diag, &e.aref, &e.efc_D, &e.efc_force, &e.efc_frictionloss, &e.Jaref, &e.jv}) v->assign(nC, 0);
    e.active.assign(nC, 0); e.prev_active.assign(nC, 0);
    for (int l = 0; l < nL; l++) {  // fixed root links keep their model pose (FK:732-736)
      e.l_pos[l] = m.v3(m.links_pos, l);
      e.l_quat[l] = m.q4(m.links_quat, l);
    }
  }

  // RS.set_state / collider.clear / constraint_solver.reset semantics for one env
  void reset_env(Env<T>& e) {
    std::fill(e.vel.begin(), e.vel.end(), T(0));
    std::fill(e.acc.begin(), e.acc.end(), T(0));
    std::fill(e.qacc_ws.begin(), e.qacc_ws.end(), T(0));
    e.is_warmstart = 0;
    for (auto& n : e.contact_cache_normal) n = V();
    e.n_contacts = 0; e.n_con = 0; e.n_eq = 0; e.n_floss = 0;
    e.errno_ = 0;
    e.sap_value.clear();  // collider.clear: first_time = True (the next SAP sweep rebuilds its sort buffer)
  }

  void init_fixed_geoms(Env<T>& e) {  // RS.build: _func_update_geoms(force_update_fixed_geoms=True), RS:348
    for (int g = 0; g < m.nG; g++) {
      int l = m.geoms_link[g];
      e.g_pos[g] = e.l_pos[l] + transform_by_quat(m.v3(m.geoms_pos, g), e.l_quat[l]);
      e.g_quat[g] = transform_quat_by_quat(m.q4(m.geoms_quat, g), e.l_quat[l]);
    }
  }

  void forward_all(Env<T>& e) {
    update_cartesian_space(e);
    forward_velocity(e);
  }

  // ------------------------------------------------------------------ kinematics
  // FK:580-736 func_forward_kinematics_entity
  void fk_entity(Env<T>& e, int i_e) {
    for (int l = m.entities_link_start[i_e]; l < m.entities_link_end[i_e]; l++) {
      V pos = m.v3(m.links_pos, l);
      Q quat = m.q4(m.links_quat, l);
      int p = m.links_parent[l];
      if (p != -1) {
        pos = e.l_pos[p] + transform_by_quat(m.v3(m.links_pos, l), e.l_quat[p]);
        quat = transform_quat_by_quat(m.q4(m.links_quat, l), e.l_quat[p]);
      }
      for (int j = m.links_joint_start[l]; j < m.links_joint_end[l]; j++) {
        int type = m.joints_type[j], qs = m.joints_q_start[j], ds = m.joints_dof_start[j];
        V jpos = m.v3(m.joints_pos, j);
        if (type == JOINT_FREE) {
          e.xanchor[j] = V(e.qpos[qs], e.qpos[qs + 1], e.qpos[qs + 2]);
          e.xaxis[j] = V(0, 0, 1);
        } else if (type == JOINT_FIXED) {
        } else {
          V axis(0, 0, 1);
          if (type == JOINT_REVOLUTE) axis = m.v3(m.dofs_motion_ang, ds);
          else if (type == JOINT_PRISMATIC) axis = m.v3(m.dofs_motion_vel, ds);
          e.xanchor[j] = transform_by_quat(jpos, quat) + pos;
          e.xaxis[j] = transform_by_quat(axis, quat);
        }
        if (type == JOINT_FREE) {
          pos = V(e.qpos[qs], e.qpos[qs + 1], e.qpos[qs + 2]);
          Q q(e.qpos[qs + 3], e.qpos[qs + 4], e.qpos[qs + 5], e.qpos[qs + 6]);
          T n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
          quat = Q(q.w / n, q.x / n, q.y / n, q.z / n);
          V xyz = quat_to_xyz(quat, EPS);
          for (int k = 0; k < 3; k++) { e.dofs_pos[ds + k] = pos[k]; e.dofs_pos[ds + 3 + k] = xyz[k]; }
        } else if (type == JOINT_FIXED) {
        } else if (type == JOINT_SPHERICAL) {
          Q qloc(e.qpos[qs], e.qpos[qs + 1], e.qpos[qs + 2], e.qpos[qs + 3]);
          V xyz = quat_to_xyz(qloc, EPS);
          for (int k = 0; k < 3; k++) e.dofs_pos[ds + k] = xyz[k];
          quat = transform_quat_by_quat(qloc, quat);
          pos = e.xanchor[j] - transform_by_quat(jpos, quat);
        } else if (type == JOINT_REVOLUTE) {
          V axis = m.v3(m.dofs_motion_ang, ds);
          e.dofs_pos[ds] = e.qpos[qs] - m.qpos0[qs];
          Q qloc = rotvec_to_quat(axis * e.dofs_pos[ds], EPS);
          quat = transform_quat_by_quat(qloc, quat);
          pos = e.xanchor[j] - transform_by_quat(jpos, quat);
        } else {  // PRISMATIC
          e.dofs_pos[ds] = e.qpos[qs] - m.qpos0[qs];
          pos = pos + e.xaxis[j] * e.dofs_pos[ds];
        }
      }
      if (!(m.links_parent[l] == -1 && m.links_is_fixed[l])) {
        e.l_pos[l] = pos;
        e.l_quat[l] = quat;
      }
    }
  }

  // geom.py:320-336
  void transform_inertia(const M3<T>& Iin, T mass, const V& trans, const Q& quat, M3<T>& I, V& first) {
    T x = trans.x, y = trans.y, z = trans.z;
    T xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
    M3<T> hhT;
    hhT.m[0][0] = yy + zz; hhT.m[0][1] = -xy; hhT.m[0][2] = -xz;
    hhT.m[1][0] = -xy; hhT.m[1][1] = xx + zz; hhT.m[1][2] = -yz;
    hhT.m[2][0] = -xz; hhT.m[2][1] = -yz; hhT.m[2][2] = xx + yy;
    M3<T> R = quat_to_R(quat, EPS);
    M3<T> RI;  // R @ I
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { T s = 0; for (int k = 0; k < 3; k++) s += R.m[i][k] * Iin.m[k][j]; RI.m[i][j] = s; }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { T s = 0; for (int k = 0; k < 3; k++) s += RI.m[i][k] * R.m[j][k]; I.m[i][j] = s + hhT.m[i][j] * mass; }
    first = trans * mass;
  }

  // FK:337-576 func_COM_links_entity
  void com_entity(Env<T>& e, int i_e) {
    int ls = m.entities_link_start[i_e], le = m.entities_link_end[i_e];
    for (int l = ls; l < le; l++) { e.root_COM_bw[l] = V(); e.mass_sum[l] = 0; }
    for (int l = ls; l < le; l++) {
      T mass = m.links_inertial_mass[l] + e.mass_shift[l];  // FK:377
      e.i_pos_bw[l] = e.l_pos[l] + transform_by_quat(m.v3(m.links_inertial_pos, l) + e.i_pos_shift[l], e.l_quat[l]);  // FK:382
      e.i_quat[l] = transform_quat_by_quat(m.q4(m.links_inertial_quat, l), e.l_quat[l]);
      int r = m.links_root[l];
      e.mass_sum[r] += mass;
      e.root_COM_bw[r] += e.i_pos_bw[l] * mass;
    }
    for (int l = ls; l < le; l++) {
      if (m.links_root[l] == l) {
        if (e.mass_sum[l] > EPS) e.root_COM[l] = e.root_COM_bw[l] / e.mass_sum[l];
        else e.root_COM[l] = e.i_pos_bw[l];
      }
    }
    for (int l = ls; l < le; l++) e.root_COM[l] = e.root_COM[m.links_root[l]];
    for (int l = ls; l < le; l++) {
      e.i_pos[l] = e.i_pos_bw[l] - e.root_COM[l];
      M3<T> Iin;
      for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Iin.m[a][b] = m.links_inertial_i[9 * l + 3 * a + b];
      const T i_mass = m.links_inertial_mass[l] + e.mass_shift[l];  // FK:437
      transform_inertia(Iin, i_mass, e.i_pos[l], e.i_quat[l], e.cinr_I[l], e.cinr_pos[l]);
      e.cinr_mass[l] = i_mass;
    }
    // j_pos/j_quat (FK:451-515) are only consumed by accessors; omitted.
    for (int l = ls; l < le; l++) {
      if (m.links_dof_end[l] - m.links_dof_start[l] <= 0) continue;
      for (int j = m.links_joint_start[l]; j < m.links_joint_end[l]; j++) {
        V off = e.root_COM[l] - e.xanchor[j];
        int type = m.joints_type[j], ds = m.joints_dof_start[j];
        if (type == JOINT_REVOLUTE) {
          e.cdof_ang[ds] = e.xaxis[j];
          e.cdof_vel[ds] = e.xaxis[j].cross(off);
        } else if (type == JOINT_PRISMATIC) {
          e.cdof_ang[ds] = V();
          e.cdof_vel[ds] = e.xaxis[j];
        } else if (type == JOINT_SPHERICAL) {
          M3<T> R = quat_to_R(e.l_quat[l], EPS);
          for (int i = 0; i < 3; i++) { e.cdof_ang[ds + i] = R.col(i); e.cdof_vel[ds + i] = R.col(i).cross(off); }
        } else if (type == JOINT_FREE) {
          for (int i = 0; i < 3; i++) { e.cdof_ang[ds + i] = V(); e.cdof_vel[ds + i] = V(); e.cdof_vel[ds + i][i] = 1; }
          M3<T> R = quat_to_R(e.l_quat[l], EPS);
          for (int i = 0; i < 3; i++) { e.cdof_ang[ds + 3 + i] = R.col(i); e.cdof_vel[ds + 3 + i] = R.col(i).cross(off); }
        }
        for (int d = ds; d < ds + m.joints_n_dofs[j]; d++) {
          e.cdofvel_ang[d] = e.cdof_ang[d] * e.vel[d];
          e.cdofvel_vel[d] = e.cdof_vel[d] * e.vel[d];
        }
      }
    }
  }

  // FK:828-864 func_update_geoms_entity (force_update_fixed_geoms=False)
  void geoms_entity(Env<T>& e, int i_e) {
    for (int g = m.entities_geom_start[i_e]; g < m.entities_geom_end[i_e]; g++) {
      if (m.geoms_is_fixed[g]) continue;
      int l = m.geoms_link[g];
      e.g_pos[g] = e.l_pos[l] + transform_by_quat(m.v3(m.geoms_pos, g), e.l_quat[l]);
      e.g_quat[g] = transform_quat_by_quat(m.q4(m.geoms_quat, g), e.l_quat[l]);
    }
  }

  // FK:1622-1698 func_update_cartesian_space
  void update_cartesian_space(Env<T>& e) {
    for (int i_e = 0; i_e < m.nE; i_e++) {
      fk_entity(e, i_e);
      com_entity(e, i_e);
      geoms_entity(e, i_e);
    }
  }

  // FK:993-1116 func_forward_velocity_entity
  void forward_velocity(Env<T>& e) {
    for (int l = 0; l < m.nL; l++) {
      V cvel_vel, cvel_ang;
      int p = m.links_parent[l];
      if (p != -1) { cvel_vel = e.cd_vel[p]; cvel_ang = e.cd_ang[p]; }
      for (int j = m.links_joint_start[l]; j < m.links_joint_end[l]; j++) {
        int type = m.joints_type[j], ds = m.joints_dof_start[j], nd = m.joints_n_dofs[j];
        if (type == JOINT_FREE) {
          for (int k = 0; k < 3; k++) {
            cvel_vel = cvel_vel + e.cdof_vel[ds + k] * e.vel[ds + k];
            cvel_ang = cvel_ang + e.cdof_ang[ds + k] * e.vel[ds + k];
          }
          for (int k = 0; k < 3; k++) {
            e.cdofd_ang[ds + k] = V(); e.cdofd_vel[ds + k] = V();
            motion_cross_motion(cvel_ang, cvel_vel, e.cdof_ang[ds + k + 3], e.cdof_vel[ds + k + 3], e.cdofd_ang[ds + k + 3], e.cdofd_vel[ds + k + 3]);
          }
          for (int k = 0; k < 3; k++) {
            cvel_vel = cvel_vel + e.cdof_vel[ds + k + 3] * e.vel[ds + k + 3];
            cvel_ang = cvel_ang + e.cdof_ang[ds + k + 3] * e.vel[ds + k + 3];
          }
        } else {
          for (int d = ds; d < ds + nd; d++) motion_cross_motion(cvel_ang, cvel_vel, e.cdof_ang[d], e.cdof_vel[d], e.cdofd_ang[d], e.cdofd_vel[d]);
          for (int d = ds; d < ds + nd; d++) {
            cvel_vel = cvel_vel + e.cdof_vel[d] * e.vel[d];
            cvel_ang = cvel_ang + e.cdof_ang[d] * e.vel[d];
          }
        }
      }
      e.cd_vel[l] = cvel_vel;
      e.cd_ang[l] = cvel_ang;
    }
  }

  // ------------------------------------------------------------------ smooth dynamics
  // FD:292-541 func_compute_mass_matrix
  void mass_matrix(Env<T>& e, bool implicit_damping) {
    int nD = m.nD;
    for (int l = 0; l < m.nL; l++) { e.crb_I[l] = e.cinr_I[l]; e.crb_pos[l] = e.cinr_pos[l]; e.crb_mass[l] = e.cinr_mass[l]; }
    for (int i_e = 0; i_e < m.nE; i_e++) {
      for (int l = m.entities_link_end[i_e] - 1; l >= m.entities_link_start[i_e]; l--) {
        int p = m.links_parent[l];
        if (p != -1) { e.crb_I[p] += e.crb_I[l]; e.crb_mass[p] += e.crb_mass[l]; e.crb_pos[p] += e.crb_pos[l]; }
      }
    }
    for (int l = 0; l < m.nL; l++)
      for (int d = m.links_dof_start[l]; d < m.links_dof_end[l]; d++)
        inertial_mul(e.crb_pos[l], e.crb_I[l], e.crb_mass[l], e.cdof_vel[d], e.cdof_ang[d], e.f_ang[d], e.f_vel[d]);
    for (int i_e = 0; i_e < m.nE; i_e++) {
      int ds = m.entities_dof_start[i_e], de = m.entities_dof_end[i_e];
      for (int i = ds; i < de; i++)
        for (int j = ds; j < de; j++)
          e.M[i * nD + j] = (e.f_ang[i].dot(e.cdof_ang[j]) + e.f_vel[i].dot(e.cdof_vel[j])) * m.mass_parent_mask[i * nD + j];
      for (int i = ds; i < de; i++)
        for (int j = i + 1; j < de; j++) e.M[i * nD + j] = e.M[j * nD + i];
    }
    for (int d = 0; d < nD; d++) e.M[d * nD + d] += m.dofs_armature[d];
    if (implicit_damping) {
      for (int d = 0; d < nD; d++) {
        e.M[d * nD + d] += m.dofs_damping[d] * dt;
        if (e.ctrl_mode[d] <= CTRL_VELOCITY) e.M[d * nD + d] -= m.dofs_act_bias[3 * d + 2] * dt;
      }
    }
  }

  // FD:545-602 func_factor_mass (plain branch)
  void factor_mass(Env<T>& e, bool implicit_damping, const std::vector<uint8_t>* mask = nullptr) {
    int nD = m.nD;
    for (int i_e = 0; i_e < m.nE; i_e++) {
      if (mask && !(*mask)[i_e]) continue;
      int ds = m.entities_dof_start[i_e], de = m.entities_dof_end[i_e], n = de - ds;
      for (int i = ds; i < de; i++) {
        for (int j = ds; j <= i; j++) e.L[i * nD + j] = e.M[i * nD + j];
        if (implicit_damping) {
          e.L[i * nD + i] += m.dofs_damping[i] * dt;
          if (cfg.integrator == INTEGRATOR_IMPLICITFAST && e.ctrl_mode[i] <= CTRL_VELOCITY) e.L[i * nD + i] -= m.dofs_act_bias[3 * i + 2] * dt;
        }
      }
      for (int i_ = 0; i_ < n; i_++) {
        int i = de - i_ - 1;
        T Dinv = T(1) / e.L[i * nD + i];
        e.Dinv[i] = Dinv;
        for (int j_ = 0; j_ < i - ds; j_++) {
          int j = i - j_ - 1;
          T a = e.L[i * nD + j] * Dinv;
          for (int k = ds; k <= j; k++) e.L[j * nD + k] -= a * e.L[i * nD + k];
          e.L[i * nD + j] = a;
        }
        e.L[i * nD + i] = 1;
      }
    }
  }

  // FD:813-894 func_solve_mass_entity
  void solve_mass(Env<T>& e, const std::vector<T>& vec, std::vector<T>& out, const std::vector<uint8_t>* mask = nullptr) {
    int nD = m.nD;
    for (int i_e = 0; i_e < m.nE; i_e++) {
      if (mask && !(*mask)[i_e]) continue;
      int ds = m.entities_dof_start[i_e], de = m.entities_dof_end[i_e];
      for (int i = de - 1; i >= ds; i--) {
        T c = vec[i];
        for (int j = i + 1; j < de; j++) c -= e.L[j * nD + i] * out[j];
        out[i] = c;
      }
      for (int i = ds; i < de; i++) out[i] *= e.Dinv[i];
      for (int i = ds; i < de; i++) {
        T c = out[i];
        for (int j = ds; j < i; j++) c -= e.L[i * nD + j] * out[j];
        out[i] = c;
      }
    }
  }

  // FD:956-1182 func_torque_and_passive_force
  void torque_and_passive(Env<T>& e) {
    for (int l = 0; l < m.nL; l++) {
      int ds = m.links_dof_start[l], de = m.links_dof_end[l];
      if (de - ds <= 0) continue;
      int jt = m.joints_type[m.links_joint_start[l]];
      for (int d = ds; d < de; d++) {
        T force = 0;
        if (e.ctrl_mode[d] == CTRL_FORCE) force = e.ctrl_force[d];
        else if (e.ctrl_mode[d] == CTRL_VELOCITY) force = -m.dofs_act_bias[3 * d + 2] * (e.ctrl_vel[d] - e.vel[d]);
        else if (e.ctrl_mode[d] == CTRL_POSITION && !(jt == JOINT_FREE && d >= ds + 3))
          force = m.dofs_kp[d] * (e.ctrl_pos[d] - e.dofs_pos[d]) + m.dofs_act_bias[3 * d] + (m.dofs_kp[d] + m.dofs_act_bias[3 * d + 1]) * e.dofs_pos[d] +
                  m.dofs_act_bias[3 * d + 2] * (e.vel[d] - e.ctrl_vel[d]);
        e.qf_applied[d] = std::min(std::max(force, m.dofs_force_range[2 * d]), m.dofs_force_range[2 * d + 1]);
      }
      if (jt == JOINT_FREE && (e.ctrl_mode[ds + 3] == CTRL_POSITION || e.ctrl_mode[ds + 4] == CTRL_POSITION || e.ctrl_mode[ds + 5] == CTRL_POSITION)) {
        V xyz(e.dofs_pos[ds + 3], e.dofs_pos[ds + 4], e.dofs_pos[ds + 5]);
        V cxyz(e.ctrl_pos[ds + 3], e.ctrl_pos[ds + 4], e.ctrl_pos[ds + 5]);
        Q qd = transform_quat_by_quat(xyz_to_quat(cxyz), inv_quat(xyz_to_quat(xyz)));
        V rv = quat_to_rotvec(qd, EPS);
        for (int k = 0; k < 3; k++) {
          int d = ds + 3 + k;
          T force = m.dofs_kp[d] * rv[k] + m.dofs_act_bias[3 * d] + (m.dofs_kp[d] + m.dofs_act_bias[3 * d + 1]) * e.dofs_pos[d] +
                    m.dofs_act_bias[3 * d + 2] * (e.vel[d] - e.ctrl_vel[d]);
          e.qf_applied[d] = std::min(std::max(force, m.dofs_force_range[2 * d]), m.dofs_force_range[2 * d + 1]);
        }
      }
    }
    for (int d = 0; d < m.nD; d++) e.qf_passive[d] = -m.dofs_damping[d] * e.vel[d];
    for (int l = 0; l < m.nL; l++) {
      int ds = m.links_dof_start[l], de = m.links_dof_end[l];
      if (de - ds <= 0) continue;
      int jt = m.joints_type[m.links_joint_start[l]];
      if (jt != JOINT_FREE && jt != JOINT_FIXED)
        for (int d = ds; d < de; d++) e.qf_passive[d] += -e.dofs_pos[d] * m.dofs_stiffness[d];
    }
  }

  // FD:1186-1285 func_update_acc
  void update_acc(Env<T>& e, bool update_cacc) {
    for (int i_e = 0; i_e < m.nE; i_e++) {
      for (int l = m.entities_link_start[i_e]; l < m.entities_link_end[i_e]; l++) {
        int p = m.links_parent[l];
        if (p == -1) {
          e.cdd_vel[l] = -gravity * (T(1) - m.entities_gravity_compensation[i_e]);
          e.cdd_ang[l] = V();
          if (update_cacc) { e.cacc_lin[l] = V(); e.cacc_ang[l] = V(); }
        } else {
          e.cdd_vel[l] = e.cdd_vel[p]; e.cdd_ang[l] = e.cdd_ang[p];
          if (update_cacc) { e.cacc_lin[l] = e.cacc_lin[p]; e.cacc_ang[l] = e.cacc_ang[p]; }
        }
        for (int d = m.links_dof_start[l]; d < m.links_dof_end[l]; d++) {
          V lv = e.cdofd_vel[d] * e.vel[d], la = e.cdofd_ang[d] * e.vel[d];
          e.cdd_vel[l] += lv; e.cdd_ang[l] += la;
          if (update_cacc) { e.cacc_lin[l] += lv + e.cdof_vel[d] * e.acc[d]; e.cacc_ang[l] += la + e.cdof_ang[d] * e.acc[d]; }
        }
      }
    }
  }

  // FD:1288-1400 func_update_force
  void update_force(Env<T>& e) {
    for (int l = 0; l < m.nL; l++) {
      V f1a, f1v, f2a, f2v, f3a, f3v;
      inertial_mul(e.cinr_pos[l], e.cinr_I[l], e.cinr_mass[l], e.cdd_vel[l], e.cdd_ang[l], f1a, f1v);
      inertial_mul(e.cinr_pos[l], e.cinr_I[l], e.cinr_mass[l], e.cd_vel[l], e.cd_ang[l], f2a, f2v);
      motion_cross_force(e.cd_ang[l], e.cd_vel[l], f2a, f2v, f3a, f3v);
      e.cfrc_vel[l] = f1v + f3v + e.cfrc_applied_vel[l];
      e.cfrc_ang[l] = f1a + f3a + e.cfrc_applied_ang[l];
    }
    for (int i_e = 0; i_e < m.nE; i_e++)
      for (int l = m.entities_link_end[i_e] - 1; l >= m.entities_link_start[i_e]; l--) {
        int p = m.links_parent[l];
        if (p != -1) { e.cfrc_vel[p] += e.cfrc_vel[l]; e.cfrc_ang[p] += e.cfrc_ang[l]; }
      }
  }

  // FD:1427-1486 func_bias_force
  void bias_force(Env<T>& e) {
    for (int l = 0; l < m.nL; l++)
      for (int d = m.links_dof_start[l]; d < m.links_dof_end[l]; d++) {
        e.qf_bias[d] = e.cdof_ang[d].dot(e.cfrc_ang[l]) + e.cdof_vel[d].dot(e.cfrc_vel[l]);
        e.force[d] = e.qf_passive[d] - e.qf_bias[d] + e.qf_applied[d];
        e.qf_smooth[d] = e.force[d];
      }
  }

  // FD:146-227 func_forward_dynamics
  void forward_dynamics(Env<T>& e) {
    mass_matrix(e, cfg.integrator == INTEGRATOR_APPROX_IMPLICITFAST);
    factor_mass(e, false);
    torque_and_passive(e);
    update_acc(e, false);
    update_force(e);
    bias_force(e);
    solve_mass(e, e.force, e.acc_smooth);
    e.acc = e.acc_smooth;
  }

  // ------------------------------------------------------------------ collision detection
  // FK:1297-1319 kernel_update_geom_aabbs
  void update_aabbs(Env<T>& e) {
    for (int g = 0; g < m.nG; g++) {
      T inf = std::numeric_limits<T>::infinity();
      V lo(inf, inf, inf), hi(-inf, -inf, -inf);
      for (int c = 0; c < 8; c++) {
        V p = transform_by_trans_quat(V(m.geoms_init_aabb[(g * 8 + c) * 3], m.geoms_init_aabb[(g * 8 + c) * 3 + 1], m.geoms_init_aabb[(g * 8 + c) * 3 + 2]), e.g_pos[g], e.g_quat[g]);
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
      }
      e.aabb_min[g] = lo; e.aabb_max[g] = hi;
    }
  }

  // BP:74-139 (clear) + BP:400-460 (all-vs-all traversal, visited in valid-pair order => deterministic)
  // BP:33-70 func_check_collision_valid: a valid static pair that is not tied by a dynamically registered weld
  bool pair_valid(const Env<T>& e, int ga, int gb) const {
    if (m.pair_idx[ga * m.nG + gb] == -1) return false;
    const int la = m.geoms_link[ga], lb = m.geoms_link[gb];
    for (const auto& w : e.dyn_welds)
      if ((w.l1 == la && w.l2 == lb) || (w.l1 == lb && w.l2 == la)) return false;
    return true;
  }
  // BP:142-396 _func_broad_phase_sap (the traversal gs.cpu uses, rigid_solver.py:370-372): geoms sorted along x by an insertion
  // sort that is warm-started from the previous step's order, then one sweep with an active list.  Same candidate SET as the
  // all-vs-all traversal of the GPU backend (BP:400-460), in sweep order.
  void broad_phase_sap(Env<T>& e) {
    for (int c = 0; c < e.n_contacts; c++) {
      Contact<T>& ct = e.contacts[c];
      ct.link_a = ct.link_b = ct.geom_a = ct.geom_b = -1;
      ct.penetration = 0; ct.pos = V(); ct.normal = V(); ct.force = V();
    }
    e.n_contacts = 0;
    const int nG = m.nG, axis = 0;
    if (e.sap_value.empty()) {  // first_time: [min, max] of every geom in geom order
      e.sap_value.resize(2 * nG); e.sap_geom.resize(2 * nG); e.sap_is_max.resize(2 * nG);
      for (int g = 0; g < nG; g++) {
        e.sap_value[2 * g] = e.aabb_min[g][axis]; e.sap_geom[2 * g] = g; e.sap_is_max[2 * g] = 0;
        e.sap_value[2 * g + 1] = e.aabb_max[g][axis]; e.sap_geom[2 * g + 1] = g; e.sap_is_max[2 * g + 1] = 1;
      }
    } else {
      for (int i = 0; i < 2 * nG; i++) e.sap_value[i] = e.sap_is_max[i] ? e.aabb_max[e.sap_geom[i]][axis] : e.aabb_min[e.sap_geom[i]][axis];
    }
    for (int i = 1; i < 2 * nG; i++) {  // insertion sort: near O(n) on a nearly sorted buffer
      const T kv = e.sap_value[i];
      const uint8_t km = e.sap_is_max[i];
      const int kg = e.sap_geom[i];
      int j = i - 1;
      while (j >= 0 && kv < e.sap_value[j]) {
        e.sap_value[j + 1] = e.sap_value[j]; e.sap_is_max[j + 1] = e.sap_is_max[j]; e.sap_geom[j + 1] = e.sap_geom[j];
        j--;
      }
      e.sap_value[j + 1] = kv; e.sap_is_max[j + 1] = km; e.sap_geom[j + 1] = kg;
    }
    e.n_broad = 0;
    std::vector<int> active;
    for (int i = 0; i < 2 * nG; i++) {
      if (!e.sap_is_max[i]) {
        bool overflow = false;
        for (size_t j = 0; j < active.size(); j++) {
          int ga = active[j], gb = e.sap_geom[i];
          if (ga > gb) std::swap(ga, gb);
          if (!pair_valid(e, ga, gb)) continue;
          bool sep = false;
          for (int k = 0; k < 3; k++) if (e.aabb_max[ga][k] <= e.aabb_min[gb][k] || e.aabb_min[ga][k] >= e.aabb_max[gb][k]) sep = true;
          if (sep) { e.contact_cache_normal[m.pair_idx[ga * m.nG + gb]] = V(); continue; }
          if (e.n_broad == m.max_broad_pairs) { e.errno_ |= ERR_BROAD_OVERFLOW; overflow = true; break; }
          e.broad_pairs[2 * e.n_broad] = ga; e.broad_pairs[2 * e.n_broad + 1] = gb; e.n_broad++;
        }
        (void)overflow;
        active.push_back(e.sap_geom[i]);
      } else {
        for (size_t j = 0; j < active.size(); j++)
          if (active[j] == e.sap_geom[i]) { active.erase(active.begin() + j); break; }
      }
    }
  }

  void broad_phase(Env<T>& e) {
    if (cfg.broadphase_sap) { broad_phase_sap(e); return; }
    for (int c = 0; c < e.n_contacts; c++) {
      Contact<T>& ct = e.contacts[c];
      ct.link_a = ct.link_b = ct.geom_a = ct.geom_b = -1;
      ct.penetration = 0; ct.pos = V(); ct.normal = V(); ct.force = V();
    }
    e.n_contacts = 0;
    e.n_broad = 0;
    for (int p = 0; p < m.nP; p++) {
      int ga = m.pairs[2 * p], gb = m.pairs[2 * p + 1];
      if (!e.dyn_welds.empty() && !pair_valid(e, ga, gb)) continue;  // BP:46-56: pairs tied by a dynamically registered weld
      bool sep = false;  // collider/utils.py func_is_geom_aabbs_overlap
      for (int k = 0; k < 3; k++) if (e.aabb_max[ga][k] <= e.aabb_min[gb][k] || e.aabb_min[ga][k] >= e.aabb_max[gb][k]) sep = true;
      if (sep) { e.contact_cache_normal[p] = V(); continue; }
      if (e.n_broad < m.max_broad_pairs) {
        e.broad_pairs[2 * e.n_broad] = ga; e.broad_pairs[2 * e.n_broad + 1] = gb; e.n_broad++;
      } else {
        e.errno_ |= ERR_BROAD_OVERFLOW;
      }
    }
  }

  // float -> int conversion with CUDA semantics (NaN -> 0, saturating): the reference's GPU backend inherits these from
  // cvt.rzi.s32.f32, and acos() of a |z| one ulp above 1 does produce NaN here (SF:143-150 has no clamp)
  static int f2i(T x) {
    if (std::isnan(x)) return 0;
    if (x >= T(2147483647.0)) return 2147483647;
    if (x <= T(-2147483648.0)) return (-2147483647 - 1);
    return (int)x;
  }
  // SF:138-180 _func_support_mesh
  V support_mesh(const V& d, int g) {
    const T PI = T(3.14159265358979323846);
    T theta = std::atan2(d.y, d.x), phi = std::acos(d.z);
    const int res = 180;
    T ii = (theta + PI) / PI / 2 * res, jj = phi / PI * res;
    T dot_max = T(-1e20);
    V best;
    for (int i4 = 0; i4 < 4; i4++) {
      int i, j;
      if (i4 % 2) i = f2i(std::ceil(ii)) % res; else i = f2i(std::floor(ii)) % res;
      if (i4 / 2 > 0) { j = f2i(std::fmin(std::fmax(std::ceil(jj), T(0)), T(res - 1))); if (j == res - 1) j = res - 2; }
      else { j = f2i(std::fmin(std::fmax(std::floor(jj), T(0)), T(res - 1))); if (j == 0) j = 1; }
      int idx = m.geoms_support_start[g] + i * res + j;
      V p(m.support_v[3 * idx], m.support_v[3 * idx + 1], m.support_v[3 * idx + 2]);
      T dot = p.dot(d);
      if (dot > dot_max) { best = p; dot_max = dot; }
    }
    return best;
  }

  // the prism of the height field under test (collider_state.prism): bottom triangle [0..3), top triangle [3..6), terrain frame
  static V* terrain_prism() { static thread_local V prism[6]; return prism; }
  // SF:268-285 _func_support_prism
  V support_prism(const V& d) {
    const V* prism = terrain_prism();
    int istart = d.z < 0 ? 0 : 3, ibest = istart;
    T best = prism[istart].dot(d);
    for (int i = istart + 1; i < istart + 3; i++) { T dot = prism[i].dot(d); if (dot > best) { ibest = i; best = dot; } }
    return prism[ibest];
  }

  // MPR:146-176 support_driver (+ SF:183-309)
  V support_driver(const V& d, int g, const V& pos, const Q& quat) {
    int type = m.geoms_type[g];
    if (type == GEOM_TERRAIN) return support_prism(d);  // MPR:168-172: does not depend on pos / quat
    if (type == GEOM_SPHERE) return pos + d * m.geoms_data[7 * g];
    if (type == GEOM_ELLIPSOID) {
      T a = m.geoms_data[7 * g], b = m.geoms_data[7 * g + 1], c = m.geoms_data[7 * g + 2];
      V dl = inv_transform_by_quat(d, quat);
      T n = std::sqrt(a * a * dl.x * dl.x + b * b * dl.y * dl.y + c * c * dl.z * dl.z);
      return pos + transform_by_quat(V(a * a * dl.x / n, b * b * dl.y / n, c * c * dl.z / n), quat);
    }
    if (type == GEOM_CAPSULE) {
      T r = m.geoms_data[7 * g], hl = T(0.5) * m.geoms_data[7 * g + 1];
      V axis = transform_by_quat(V(0, 0, 1), quat);
      T side = d.dot(axis) < 0 ? T(-1) : T(1);
      return pos + axis * (hl * side) + d * r;
    }
    if (type == GEOM_BOX) {
      V db = inv_transform_by_quat(d, quat);
      V v((db.x < 0 ? T(-1) : T(1)) * m.geoms_data[7 * g] * T(0.5), (db.y < 0 ? T(-1) : T(1)) * m.geoms_data[7 * g + 1] * T(0.5),
          (db.z < 0 ? T(-1) : T(1)) * m.geoms_data[7 * g + 2] * T(0.5));
      return transform_by_trans_quat(v, pos, quat);
    }
    V dm = transform_by_quat(d, inv_quat(quat));  // SF:121-135 _func_support_world
    return transform_by_trans_quat(support_mesh(dm, g), pos, quat);
  }

  // ---- MPR (mpr.py) -------------------------------------------------------------------------------
  struct Simplex { V v[4], v1[4], v2[4]; int size; };
  void compute_support(const V& d, int ga, int gb, const V& pa, const Q& qa, const V& pb, const Q& qb, V& v, V& v1, V& v2) {
    v1 = support_driver(d, ga, pa, qa);
    v2 = support_driver(-d, gb, pb, qb);
    v = v1 - v2;
  }
  V portal_dir(const Simplex& s) { return (s.v[2] - s.v[1]).cross(s.v[3] - s.v[1]).normalized(); }
  bool reach_tolerance(const Simplex& s, const V& v, const V& dir) {  // MPR:134-143
    T dv1 = s.v[1].dot(dir), dv2 = s.v[2].dot(dir), dv3 = s.v[3].dot(dir), dv4 = v.dot(dir);
    T dot1 = std::min(std::min(dv4 - dv1, dv4 - dv2), dv4 - dv3);
    return dot1 < (T)cfg.ccd_tolerance + (T)cfg.ccd_eps * std::max(T(1), dot1);
  }
  void expand_portal(Simplex& s, const V& v, const V& v1, const V& v2) {  // MPR:426-442
    V v4v0 = v.cross(s.v[0]);
    T dot = s.v[1].dot(v4v0);
    int is;
    if (dot > 0) { dot = s.v[2].dot(v4v0); is = dot > 0 ? 1 : 3; }
    else { dot = s.v[3].dot(v4v0); is = dot > 0 ? 2 : 1; }
    s.v1[is] = v1; s.v2[is] = v2; s.v[is] = v;
  }
  int discover_portal(Simplex& s, int ga, int gb, const V& ca, const V& cb, const V& pa, const Q& qa, const V& pb, const Q& qb) {  // MPR:445-598
    const T CE = (T)cfg.ccd_eps;
    s.v1[0] = ca; s.v2[0] = cb; s.v[0] = ca - cb; s.size = 1;
    if (std::abs(s.v[0].x) < CE && std::abs(s.v[0].y) < CE && std::abs(s.v[0].z) < CE) s.v[0].x += T(10) * CE;
    V dir = -s.v[0].normalized();
    V v, v1, v2;
    compute_support(dir, ga, gb, pa, qa, pb, qb, v, v1, v2);
    s.v1[1] = v1; s.v2[1] = v2; s.v[1] = v; s.size = 2;
    T dot = v.dot(dir);
    if (dot < CE) return -1;
    dir = s.v[0].cross(s.v[1]);
    if (dir.dot(dir) < CE) {
      if (std::abs(s.v[1].x) < CE && std::abs(s.v[1].y) < CE && std::abs(s.v[1].z) < CE) return 1;
      return 2;
    }
    dir = dir.normalized();
    compute_support(dir, ga, gb, pa, qa, pb, qb, v, v1, v2);
    dot = v.dot(dir);
    if (dot < CE) return -1;
    s.v1[2] = v1; s.v2[2] = v2; s.v[2] = v; s.size = 3;
    V va = s.v[1] - s.v[0], vb = s.v[2] - s.v[0];
    dir = va.cross(vb).normalized();
    dot = dir.dot(s.v[0]);
    if (dot > 0) {
      std::swap(s.v[1], s.v[2]); std::swap(s.v1[1], s.v1[2]); std::swap(s.v2[1], s.v2[2]);
      dir = -dir;
    }
    int num_trials = 0;
    while (s.size < 4) {
      compute_support(dir, ga, gb, pa, qa, pb, qb, v, v1, v2);
      dot = v.dot(dir);
      if (dot < CE) return -1;
      bool cont = false;
      va = s.v[1].cross(v);
      dot = va.dot(s.v[0]);
      if (dot < -CE) { s.v1[2] = v1; s.v2[2] = v2; s.v[2] = v; cont = true; }
      if (!cont) {
        va = v.cross(s.v[2]);
        dot = va.dot(s.v[0]);
        if (dot < -CE) { s.v1[1] = v1; s.v2[1] = v2; s.v[1] = v; cont = true; }
      }
      if (cont) {
        va = s.v[1] - s.v[0]; vb = s.v[2] - s.v[0];
        dir = va.cross(vb).normalized();
        if (++num_trials == 15) return -1;
      } else {
        s.v1[3] = v1; s.v2[3] = v2; s.v[3] = v; s.size = 4;
      }
    }
    return 0;
  }
  int refine_portal(Simplex& s, int ga, int gb, const V& pa, const Q& qa, const V& pb, const Q& qb) {  // MPR:232-278
    const T CE = (T)cfg.ccd_eps;
    while (true) {
      V dir = portal_dir(s);
      if (s.v[1].dot(dir) > -CE) return 0;
      V v, v1, v2;
      compute_support(dir, ga, gb, pa, qa, pb, qb, v, v1, v2);
      if (!(v.dot(dir) > -CE) || reach_tolerance(s, v, dir)) return -1;
      expand_portal(s, v, v1, v2);
    }
  }
  V find_pos(const Simplex& s) {  // MPR:281-316 (non-mujoco-compat: b = 0 initially)
    T b[4] = {0, 0, 0, 0};
    T sum = 0;
    if (sum < (T)cfg.ccd_eps) {
      V dir = portal_dir(s);
      b[0] = 0;
      for (int i = 1; i < 4; i++) {
        int i1 = i % 3 + 1, i2 = (i + 1) % 3 + 1;
        b[i] = s.v[i1].cross(s.v[i2]).dot(dir);
      }
      sum = b[0] + b[1] + b[2] + b[3];
    }
    V p1, p2;
    for (int i = 0; i < 4; i++) { p1 += s.v1[i] * b[i]; p2 += s.v2[i] * b[i]; }
    return (p1 + p2) * (T(0.5) / sum);
  }
  void find_penetration(Simplex& s, int ga, int gb, const V& pa, const Q& qa, const V& pb, const Q& qb, bool& is_col, V& normal, T& pen, V& pos) {  // MPR:338-423
    int iterations = 0;
    while (true) {
      V dir = portal_dir(s);
      V v, v1, v2;
      compute_support(dir, ga, gb, pa, qa, pb, qb, v, v1, v2);
      if (reach_tolerance(s, v, dir) || iterations > cfg.ccd_iterations) {
        pen = dir.dot(s.v[1]);
        normal = -dir;
        is_col = true;
        pos = find_pos(s);
        return;
      }
      expand_portal(s, v, v1, v2);
      iterations++;
    }
  }
  // MPR:601-683 guess_geoms_center
  void guess_centers(int ga, int gb, const V& pa, const Q& qa, const V& pb, const Q& qb, const V& normal_ws, V& ca, V& cb) {
    ca = transform_by_trans_quat(m.v3(m.geoms_center, ga), pa, qa);
    cb = transform_by_trans_quat(m.v3(m.geoms_center, gb), pb, qb);
    const T CE = (T)cfg.ccd_eps;
    if (std::abs(normal_ws.x) > CE || std::abs(normal_ws.y) > CE || std::abs(normal_ws.z) > CE) {
      auto corner = [&](int g, int c) { return V(m.geoms_init_aabb[(g * 8 + c) * 3], m.geoms_init_aabb[(g * 8 + c) * 3 + 1], m.geoms_init_aabb[(g * 8 + c) * 3 + 2]); };
      ca = transform_by_trans_quat((corner(ga, 7) + corner(ga, 0)) * T(0.5), pa, qa);
      cb = transform_by_trans_quat((corner(gb, 7) + corner(gb, 0)) * T(0.5), pb, qb);
      V delta = ca - cb;
      V normal = delta.normalized();
      if (normal_ws.cross(normal).norm() > T(0.01)) {
        V offset = normal_ws * delta.dot(normal_ws) - delta;
        T offset_norm = offset.norm();
        if (offset_norm > EPS) {
          V dir_offset = offset / offset_norm;
          V dla = inv_transform_by_quat(dir_offset, qa), dlb = inv_transform_by_quat(dir_offset, qb);
          V bsa = corner(ga, 7) - corner(ga, 0), bsb = corner(gb, 7) - corner(gb, 0);
          T length_a = bsa.x * std::abs(dla.x) + bsa.y * std::abs(dla.y) + bsa.z * std::abs(dla.z);
          T length_b = bsb.x * std::abs(dlb.x) + bsb.y * std::abs(dlb.y) + bsb.z * std::abs(dlb.z);
          T ratio = std::min(offset_norm / (length_a + length_b), T(0.5));
          ca = ca + dir_offset * (length_a * ratio);
          cb = cb - dir_offset * (length_b * ratio);
        }
      }
    }
  }
  // MPR:686-819 func_mpr_contact
  void mpr_contact(int ga, int gb, const V& normal_ws, const V& pa, const Q& qa, const V& pb, const Q& qb, bool& is_col, V& normal, T& pen, V& pos,
                   Simplex& last_simplex, bool& last_simplex_valid) {
    V ca, cb;
    guess_centers(ga, gb, pa, qa, pb, qb, normal_ws, ca, cb);
    mpr_contact_from_centers(ga, gb, ca, cb, pa, qa, pb, qb, is_col, normal, pen, pos, last_simplex, last_simplex_valid);
  }
  // MPR:686-765 func_mpr_contact_from_centers
  void mpr_contact_from_centers(int ga, int gb, const V& ca, const V& cb, const V& pa, const Q& qa, const V& pb, const Q& qb, bool& is_col,
                                V& normal, T& pen, V& pos, Simplex& last_simplex, bool& last_simplex_valid) {
    Simplex s;
    int res = discover_portal(s, ga, gb, ca, cb, pa, qa, pb, qb);
    is_col = false; pos = V(); normal = V(); pen = 0;
    last_simplex_valid = false;
    if (res == 1) { is_col = true; pen = 0; normal = -s.v[0].normalized(); pos = (s.v1[1] + s.v2[1]) * T(0.5); }
    else if (res == 2) { is_col = true; pen = s.v[1].norm(); normal = -s.v[1].normalized(); pos = (s.v1[1] + s.v2[1]) * T(0.5); }
    else if (res == 0) {
      res = refine_portal(s, ga, gb, pa, qa, pb, qb);
      if (res >= 0) {
        find_penetration(s, ga, gb, pa, qa, pb, qb, is_col, normal, pen, pos);
        last_simplex = s; last_simplex_valid = is_col;  // v[0] interior point + final portal: a tetrahedron around the origin
      }
    }
  }

  // ---- a17: exact penetration for deep / un-cached convex pairs (NP:726-733 decides when; gjk.py:170-446 func_gjk_contact,
  // epa.py:970-1183 func_safe_epa define the result: witness points w1, w2 of the minimum-translation face of the Minkowski
  // difference, pos = (w1 + w2) / 2, normal = (w2 - w1) / |w2 - w1|, penetration = face distance).  The reference seeds EPA
  // with a GJK tetrahedron; this restatement seeds it with the tetrahedron MPR has just built around the origin (interior
  // point + final portal), which the expanding polytope algorithm (van den Bergen 2001) accepts as well: the minimum over the
  // boundary of a convex set does not depend on the seed.  Limits as in gjk.py:53-56, 89-90 (50 iterations, tolerance 1e-6).
  struct EpaFace { int a, b, c; V n; T d; bool alive; };
  bool epa_make_face(const std::vector<V>& pv, const V& inside, int a, int b, int c, EpaFace& f) {
    V n = (pv[b] - pv[a]).cross(pv[c] - pv[a]);
    T len = n.norm();
    if (!(len > T(1e-20))) return false;
    n = n / len;
    if (n.dot(inside - pv[a]) > 0) { std::swap(b, c); n = -n; }
    f.a = a; f.b = b; f.c = c; f.n = n; f.d = n.dot(pv[a]); f.alive = true;
    return true;
  }
  bool epa_penetration(const Simplex& s0, int ga, int gb, const V& pa, const Q& qa, const V& pb, const Q& qb, V& normal, T& pen, V& pos) {
    const int MAXV = 40, MAXF = 128, MAXIT = 50;
    const T TOL = T(1e-6);
    std::vector<V> pv, p1, p2;
    for (int i = 0; i < 4; i++) { pv.push_back(s0.v[i]); p1.push_back(s0.v1[i]); p2.push_back(s0.v2[i]); }
    V inside = (pv[0] + pv[1] + pv[2] + pv[3]) * T(0.25);
    std::vector<EpaFace> faces;
    const int tri[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};
    for (int i = 0; i < 4; i++) {
      EpaFace f;
      if (!epa_make_face(pv, inside, tri[i][0], tri[i][1], tri[i][2], f)) return false;
      if (f.d < -TOL) return false;  // the origin is not inside the seed
      faces.push_back(f);
    }
    int best = -1;
    for (int it = 0; it < MAXIT; it++) {
      best = -1;
      for (int i = 0; i < (int)faces.size(); i++)
        if (faces[i].alive && (best < 0 || faces[i].d < faces[best].d)) best = i;
      if (best < 0) return false;
      const EpaFace fb = faces[best];
      V w, w1, w2;
      compute_support(fb.n, ga, gb, pa, qa, pb, qb, w, w1, w2);
      if (w.dot(fb.n) - fb.d < TOL) break;
      bool dup = false;
      for (size_t i = 0; i < pv.size(); i++)
        if ((pv[i] - w).dot(pv[i] - w) < T(1e-16)) dup = true;
      if (dup || (int)pv.size() >= MAXV) break;
      const int iw = (int)pv.size();
      pv.push_back(w); p1.push_back(w1); p2.push_back(w2);
      // faces seen from w go away; the boundary of the hole (edges used by exactly one removed face) is the horizon
      const size_t MAXE = 48;
      std::vector<std::pair<int, int>> edges;
      bool overflow = false;
      for (auto& f : faces) {
        if (!f.alive || !(f.n.dot(w) - f.d > T(1e-12))) continue;
        f.alive = false;
        const int e[3][2] = {{f.a, f.b}, {f.b, f.c}, {f.c, f.a}};
        for (int k = 0; k < 3; k++) {
          bool found = false;
          for (size_t q = 0; q < edges.size(); q++)
            if (edges[q].first == e[k][1] && edges[q].second == e[k][0]) { edges[q] = edges.back(); edges.pop_back(); found = true; break; }
          if (!found) {
            if (edges.size() >= MAXE) { overflow = true; break; }
            edges.push_back({e[k][0], e[k][1]});
          }
        }
        if (overflow) break;
      }
      if (overflow || edges.size() < 3 || (int)faces.size() + (int)edges.size() > MAXF) { faces[best].alive = true; break; }
      bool ok = true;
      for (auto& ed : edges) {
        EpaFace f;
        if (!epa_make_face(pv, inside, ed.first, ed.second, iw, f)) { ok = false; break; }
        faces.push_back(f);
      }
      if (!ok) return false;
    }
    if (best < 0 || !faces[best].alive) {  // the loop ended right after a removal: take the nearest surviving face
      best = -1;
      for (int i = 0; i < (int)faces.size(); i++)
        if (faces[i].alive && (best < 0 || faces[i].d < faces[best].d)) best = i;
      if (best < 0) return false;
    }
    const EpaFace& f = faces[best];
    const V p = f.n * f.d, e0 = pv[f.b] - pv[f.a], e1 = pv[f.c] - pv[f.a], e2 = p - pv[f.a];
    const T d00 = e0.dot(e0), d01 = e0.dot(e1), d11 = e1.dot(e1), d20 = e2.dot(e0), d21 = e2.dot(e1);
    const T den = d00 * d11 - d01 * d01;
    if (!(std::abs(den) > T(1e-30))) return false;
    const T lb = (d11 * d20 - d01 * d21) / den, lc = (d00 * d21 - d01 * d20) / den, la = T(1) - lb - lc;
    const V w1 = p1[f.a] * la + p1[f.b] * lb + p1[f.c] * lc, w2 = p2[f.a] * la + p2[f.b] * lb + p2[f.c] * lc;
    const V dlt = w2 - w1;
    const T len = dlt.norm();
    normal = len > T(1e-10) ? dlt / len : -f.n;
    pen = f.d;
    pos = (w1 + w2) * T(0.5);
    return true;
  }


  // ---- a17 proper: the reference's safe GJK + safe EPA (gjk.py:170-446 func_gjk_contact non-MuJoCo branch, :1209-1425
  // func_safe_gjk, :1425-1900 helpers; epa.py:970-1183 func_safe_epa, :1184-1249 witness, :1251-1387 init / attach, :274-433
  // horizon; gjk_utils.py:45-235; gjk_support.py).  Constants of gjk.py:53-110.
  struct GjkVert { V obj1, obj2, mink; int id1, id2; };
  struct GjkFace { int v[3], adj[3]; V normal; T dist2; int map_idx; };
  static constexpr int GJK_MAX_ITER = 50, EPA_MAX_ITER = 50, POLY_MAX_FACES = 300, POLY_MAX_VERTS = 64;
  struct GjkState {
    GjkVert sv[4]; int nverts = 0;            // simplex
    GjkVert pv[POLY_MAX_VERTS]; int npv = 0;  // polytope vertices
    GjkFace pf[POLY_MAX_FACES]; int nfaces = 0;
    int fmap[POLY_MAX_FACES]; int nfmap = 0;
    int hz_face[POLY_MAX_FACES], hz_edge[POLY_MAX_FACES]; int nhz = 0;
    int st_face[3 * POLY_MAX_FACES], st_edge[3 * POLY_MAX_FACES];
    int last_searched = 0;
    V w1, w2; int n_witness = 0; T distance = 0;
  };
  T gjk_FLOAT_MIN() const { return T(1e-15); }
  T gjk_FLOAT_MAX() const { return T(1e15); }
  bool is_discrete(int g) const { return m.geoms_type[g] == GEOM_MESH || m.geoms_type[g] == GEOM_BOX; }
  // gjk_support.py:66-112 support_driver (non-MuJoCo): support point in world frame + vertex identity (-1 for smooth shapes)
  void gjk_support_driver(const V& d, int g, const V& pos, const Q& quat, V& v, int& vid) {
    const int type = m.geoms_type[g];
    vid = -1;
    if (type == GEOM_BOX) {
      V db = inv_transform_by_quat(d, quat);
      V vl((db.x < 0 ? T(-1) : T(1)) * m.geoms_data[7 * g] * T(0.5), (db.y < 0 ? T(-1) : T(1)) * m.geoms_data[7 * g + 1] * T(0.5),
           (db.z < 0 ? T(-1) : T(1)) * m.geoms_data[7 * g + 2] * T(0.5));
      vid = (vl.x > 0 ? 1 : 0) + (vl.y > 0 ? 2 : 0) + (vl.z > 0 ? 4 : 0) + m.geoms_vert_start[g];  // SF:298-309
      v = transform_by_trans_quat(vl, pos, quat);
    } else if (type == GEOM_SPHERE || type == GEOM_ELLIPSOID || type == GEOM_CAPSULE) {
      v = support_driver(d, g, pos, quat);
    } else {  // support field (mesh / cylinder): SF:121-180, identity = support_vid of the winning cell
      V dm = transform_by_quat(d, inv_quat(quat));
      const T PI = T(3.14159265358979323846);
      T theta = std::atan2(dm.y, dm.x), phi = std::acos(dm.z);
      const int res = 180;
      T ii = (theta + PI) / PI / 2 * res, jj = phi / PI * res;
      T dot_max = T(-1e20);
      V best;
      int bid = 0;
      for (int i4 = 0; i4 < 4; i4++) {
        int i, j;
        if (i4 % 2) i = f2i(std::ceil(ii)) % res; else i = f2i(std::floor(ii)) % res;
        if (i4 / 2 > 0) { j = f2i(std::fmin(std::fmax(std::ceil(jj), T(0)), T(res - 1))); if (j == res - 1) j = res - 2; }
        else { j = f2i(std::fmin(std::fmax(std::floor(jj), T(0)), T(res - 1))); if (j == 0) j = 1; }
        int idx = m.geoms_support_start[g] + i * res + j;
        V p(m.support_v[3 * idx], m.support_v[3 * idx + 1], m.support_v[3 * idx + 2]);
        T dot = p.dot(dm);
        if (dot > dot_max) { best = p; dot_max = dot; bid = m.support_vid.empty() ? idx : m.support_vid[idx]; }
      }
      vid = bid;
      v = transform_by_trans_quat(best, pos, quat);
    }
  }
  // gjk.py:1870-1900 func_count_support: boxes have 2^(zero components) candidate vertices; _func_count_supports_mesh never
  // updates its running maximum, so it always answers 1 (SF:330-376) -- restated as is
  int gjk_count_support(const V& dir, int ga, int gb, const Q& qa, const Q& qb) {
    int count = 1;
    for (int i = 0; i < 2; i++) {
      const int g = i == 0 ? ga : gb;
      if (m.geoms_type[g] == GEOM_BOX) {
        V db = inv_transform_by_quat(i == 0 ? dir : -dir, i == 0 ? qa : qb);
        int z = (db.x == 0 ? 1 : 0) + (db.y == 0 ? 1 : 0) + (db.z == 0 ? 1 : 0);
        count *= (1 << z);
      }
    }
    return count;
  }
  bool gjk_vertex_duplicate(const GjkState& S, int id1, int id2) {  // gjk.py:1451-1470
    for (int i = 0; i < S.nverts; i++) {
      if (id1 == -1 || S.sv[i].id1 != id1) continue;
      if (id2 == -1 || S.sv[i].id2 != id2) continue;
      return true;
    }
    return false;
  }
  bool gjk_vertex_degenerate(const GjkState& S, const V& mink) {  // gjk.py:1473-1514
    const T deg2 = T(1e-5) * T(1e-5);
    for (int i = 0; i < S.nverts; i++) { V d = S.sv[i].mink - mink; if (d.dot(d) < deg2) return true; }
    if (S.nverts == 2) {
      V e1 = S.sv[1].mink - S.sv[0].mink, e2 = mink - S.sv[0].mink, n = e1.cross(e2);
      return n.dot(n) < deg2 * e1.dot(e1) * e2.dot(e2);
    }
    if (S.nverts == 3) {
      V e1 = (S.sv[1].mink - S.sv[0].mink).normalized(), e2 = (S.sv[2].mink - S.sv[0].mink).normalized(), n = e1.cross(e2), diff = mink - S.sv[0].mink;
      T nd = n.dot(diff);
      return nd * nd < deg2 * n.dot(n) * diff.dot(diff);
    }
    return false;
  }
  bool gjk_vertex_valid(const GjkState& S, int id1, int id2, const V& mink) { return !gjk_vertex_duplicate(S, id1, id2) && !gjk_vertex_degenerate(S, mink); }
  // gjk.py:1733-1838 func_safe_gjk_support
  GjkVert gjk_safe_support(GjkState& S, int ga, int gb, const V& pa, const Q& qa, const V& pb, const Q& qb, const V& dir) {
    GjkVert r; r.id1 = r.id2 = -1;
    for (int i = 0; i < 9; i++) {
      V nd = dir;
      if (i > 0) {
        const int j = i - 1;
        nd.x += -(T(1) - T(2) * T(j & 1)) * EPS;
        nd.y += -(T(1) - T(2) * T(j & 2)) * EPS;
        nd.z += -(T(1) - T(2) * T(j & 4)) * EPS;
      }
      nd = nd * (T(2) - nd.dot(dir));
      const int num = gjk_count_support(nd, ga, gb, qa, qb);
      if (i > 0 && num > 1) continue;
      gjk_support_driver(nd, ga, pa, qa, r.obj1, r.id1);
      gjk_support_driver(-nd, gb, pb, qb, r.obj2, r.id2);
      r.mink = r.obj1 - r.obj2;
      if (i == 0) { if (num > 1) continue; else break; }
      if (i == 8) break;
      if (gjk_vertex_valid(S, r.id1, r.id2, r.mink)) break;
    }
    return r;
  }
  // gjk.py:1660-1698 func_get_discrete_geom_vertex
  void gjk_discrete_vertex(int g, const V& pos, const Q& quat, int iv, V& v) {
    V vl;
    if (m.geoms_type[g] == GEOM_BOX)
      vl = V(((iv & 1) ? T(1) : T(-1)) * m.geoms_data[7 * g] * T(0.5), ((iv & 2) ? T(1) : T(-1)) * m.geoms_data[7 * g + 1] * T(0.5),
             ((iv & 4) ? T(1) : T(-1)) * m.geoms_data[7 * g + 2] * T(0.5));
    else { const int k = m.geoms_vert_start[g] + iv; vl = V(m.verts[3 * k], m.verts[3 * k + 1], m.verts[3 * k + 2]); }
    v = transform_by_trans_quat(vl, pos, quat);
  }
  // gjk.py:1545-1645 func_search_valid_simplex_vertex
  bool gjk_search_valid_vertex(GjkState& S, int ga, int gb, const V& pa, const Q& qa, const V& pb, const Q& qb, GjkVert& r) {
    r = GjkVert(); r.id1 = r.id2 = -1;
    if (is_discrete(ga) && is_discrete(gb)) {
      const int n0 = m.geoms_vert_end[ga] - m.geoms_vert_start[ga], n1 = m.geoms_vert_end[gb] - m.geoms_vert_start[gb];
      const int num = n0 * n1;
      for (int k = 0; k < num; k++) {
        const int mm = (k + S.last_searched) % num, i = mm / n1, j = mm % n1;
        r.id1 = m.geoms_vert_start[ga] + i; r.id2 = m.geoms_vert_start[gb] + j;
        gjk_discrete_vertex(ga, pa, qa, i, r.obj1);
        gjk_discrete_vertex(gb, pb, qb, j, r.obj2);
        r.mink = r.obj1 - r.obj2;
        if (gjk_vertex_valid(S, r.id1, r.id2, r.mink)) { S.last_searched = (mm + 1) % num; return true; }
      }
      return false;
    }
    if (S.nverts == 3) {
      V dir = (S.sv[2].mink - S.sv[0].mink).cross(S.sv[1].mink - S.sv[0].mink).normalized();
      for (int i = 0; i < 2; i++) {
        r = gjk_safe_support(S, ga, gb, pa, qa, pb, qb, i == 0 ? dir : -dir);
        if (gjk_vertex_valid(S, r.id1, r.id2, r.mink)) return true;
      }
    }
    return false;
  }
  // gjk.py:1209-1425 func_safe_gjk: true when the origin is inside the final tetrahedron
  bool safe_gjk(GjkState& S, int ga, int gb, const V& pa, const Q& qa, const V& pb, const Q& qb) {
    S.nverts = 0;
    bool init_ok = true;
    for (int i = 0; i < 4; i++) {
      V dir; dir[2 - i / 2] = T(1) - T(2) * T(i % 2);
      GjkVert r = gjk_safe_support(S, ga, gb, pa, qa, pb, qb, dir);
      if (!gjk_vertex_valid(S, r.id1, r.id2, r.mink)) {
        if (!gjk_search_valid_vertex(S, ga, gb, pa, qa, pb, qb, r)) { init_ok = false; break; }
      }
      S.sv[i] = r;
      S.nverts++;
    }
    bool intersect = false;
    if (init_ok) {
      int si[4] = {0, 1, 2, 3};
      for (int it = 0; it < GJK_MAX_ITER; it++) {
        V nrm[4]; T sd[4];
        for (int j = 0; j < 4; j++) {
          int s0 = si[2], s1 = si[1], s2 = si[3], ap = si[0];
          if (j == 1) { s0 = si[0]; s1 = si[2]; s2 = si[3]; ap = si[1]; }
          else if (j == 2) { s0 = si[1]; s1 = si[0]; s2 = si[3]; ap = si[2]; }
          else if (j == 3) { s0 = si[0]; s1 = si[1]; s2 = si[2]; ap = si[3]; }
          const V v1 = S.sv[s0].mink, v2 = S.sv[s1].mink, v3 = S.sv[s2].mink, apex = S.sv[ap].mink;  // gjk.py:1701-1730
          V n = (v3 - v1).cross(v2 - v1).normalized();
          if (n.dot(apex - v1) > 0) n = -n;
          nrm[j] = n; sd[j] = n.dot(v1);
        }
        int min_i = 0;
        for (int j = 1; j < 4; j++) if (sd[j] < sd[min_i]) min_i = j;
        const int min_si = si[min_i];
        const V min_normal = nrm[min_i];
        if (sd[min_i] >= 0) { intersect = true; break; }
        S.nverts = 3;
        if (min_si != 3) S.sv[min_si] = S.sv[3];
        GjkVert r = gjk_safe_support(S, ga, gb, pa, qa, pb, qb, min_normal);
        if (gjk_vertex_duplicate(S, r.id1, r.id2)) break;
        if (gjk_vertex_degenerate(S, r.mink)) break;
        if (r.mink.dot(min_normal) < 0) break;
        S.sv[3] = r;
        S.nverts = 4;
      }
    }
    S.distance = intersect ? T(0) : gjk_FLOAT_MAX();
    return intersect;
  }
  // epa.py:1389-1427 func_plane_normal
  bool epa_plane_normal(const V& v1, const V& v2, const V& v3, V& normal) {
    const V d21 = v2 - v1, d31 = v3 - v1, d32 = v3 - v2;
    for (int i = 0; i < 3; i++) {
      V n = i == 0 ? d32.cross(d21) : (i == 1 ? d21.cross(d31) : d31.cross(d32));
      const T nn = n.norm();
      if (nn == 0) return false;
      if (nn > gjk_FLOAT_MIN()) { normal = n.normalized(); return true; }
    }
    return false;
  }
  // epa.py:1304-1386 func_safe_attach_face_to_polytope
  bool epa_attach_face(GjkState& S, int v1, int v2, int v3, int a1, int a2, int a3) {
    const int n = S.nfaces;
    GjkFace& f = S.pf[n];
    f.v[0] = v1; f.v[1] = v2; f.v[2] = v3; f.adj[0] = a1; f.adj[1] = a2; f.adj[2] = a3;
    S.nfaces++;
    V normal;
    if (!epa_plane_normal(S.pv[v3].mink, S.pv[v2].mink, S.pv[v1].mink, normal)) return false;
    const V center = (S.pv[v1].mink + S.pv[v2].mink + S.pv[v3].mink) / T(3);
    T max_orient = -normal.dot(center), max_abs = std::abs(max_orient);
    for (int iv = 0; iv < S.npv; iv++) {
      if (iv != v1 && iv != v2 && iv != v3) {
        const T orient = normal.dot(S.pv[iv].mink - center);
        if (std::abs(orient) > max_abs) { max_abs = std::abs(orient); max_orient = orient; }
      }
    }
    if (max_orient > 0) normal = -normal;
    f.normal = normal;
    T min_d2 = gjk_FLOAT_MAX();
    const int vs[3] = {v1, v2, v3};
    for (int i = 0; i < 3; i++) { const T d = normal.dot(S.pv[vs[i]].mink); if (d * d < min_d2) min_d2 = d * d; }
    f.dist2 = min_d2;
    f.map_idx = -1;
    return true;
  }
  void epa_delete_face(GjkState& S, int i_f) {  // epa.py:384-405
    const int mi = S.pf[i_f].map_idx;
    if (mi >= 0) {
      const int last = S.fmap[S.nfmap - 1];
      S.fmap[mi] = last;
      S.pf[last].map_idx = mi;
      S.nfmap--;
    }
    S.pf[i_f].map_idx = -2;
  }
  // epa.py:274-341 func_epa_horizon (depth-first walk over the faces visible from w; the horizon edges come out as a loop)
  void epa_horizon(GjkState& S, const V& w, int nearest) {
    S.st_face[0] = nearest; S.st_edge[0] = 0;
    int top = 1;
    bool is_first = true;
    while (top > 0) {
      const int i_f = S.st_face[top - 1], i_e = S.st_edge[top - 1];
      const V v = S.pv[S.pf[i_f].v[0]].mink;
      top--;
      const bool deleted = S.pf[i_f].map_idx == -2;
      if (!is_first && deleted) continue;
      const bool visible = S.pf[i_f].normal.dot(w - v) > gjk_FLOAT_MIN();
      if (visible || is_first) {
        epa_delete_face(S, i_f);
        for (int k = is_first ? 0 : 1; k < 3; k++) {
          const int e2 = (i_e + k) % 3, adj = S.pf[i_f].adj[e2];
          if (S.pf[adj].map_idx != -2) {
            const int start_v = S.pf[i_f].v[(e2 + 1) % 3];
            const int ae = S.pf[adj].v[0] == start_v ? 0 : (S.pf[adj].v[1] == start_v ? 1 : 2);  // epa.py:362-381
            S.st_face[top] = adj; S.st_edge[top] = ae; top++;
          }
        }
      } else {
        S.hz_edge[S.nhz] = i_e; S.hz_face[S.nhz] = i_f; S.nhz++;
      }
      is_first = false;
    }
  }
  // gjk_utils.py:45-100 func_triangle_affine_coords, :183-235 func_project_origin_to_plane
  V tri_affine_coords(const V& point, const V& t1, const V& t2, const V& t3) {
    T ms[3];
    for (int i = 0; i < 3; i++) {
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
      if (i == 1) std::swap(i1, i2);
      ms[i] = t2[i1] * t3[i2] - t2[i2] * t3[i1] - t1[i1] * t3[i2] + t1[i2] * t3[i1] + t1[i1] * t2[i2] - t1[i2] * t2[i1];
    }
    T m_max = 0;
    int ix = 0, iy = 0;
    for (int i = 0; i < 3; i++) {
      if (std::abs(ms[i]) >= std::abs(ms[(i + 1) % 3]) && std::abs(ms[i]) >= std::abs(ms[(i + 2) % 3])) {
        m_max = ms[i]; ix = (i + 1) % 3; iy = (i + 2) % 3;
        if (i == 1) std::swap(ix, iy);
        break;
      }
    }
    V cs;
    for (int i = 0; i < 3; i++) {
      const V& a = i == 0 ? t2 : (i == 1 ? t3 : t1);
      const V& b = i == 0 ? t3 : (i == 1 ? t1 : t2);
      cs[i] = point[ix] * a[iy] + point[iy] * b[ix] + a[ix] * b[iy] - point[ix] * b[iy] - point[iy] * a[ix] - b[ix] * a[iy];
    }
    return cs / m_max;
  }
  bool project_origin_to_plane(const V& v1, const V& v2, const V& v3, V& point) {
    const V d21 = v2 - v1, d31 = v3 - v1, d32 = v3 - v2;
    point = V();
    for (int i = 0; i < 3; i++) {
      V n, v;
      if (i == 0) { n = d32.cross(d21); v = v2; } else if (i == 1) { n = d21.cross(d31); v = v1; } else { n = d31.cross(d32); v = v3; }
      const T nv = n.dot(v), nn = n.dot(n);
      if (nn == 0) return false;
      if (nn > gjk_FLOAT_MIN()) { point = n * (nv / nn); return true; }
      if (i == 2) { if (nn < gjk_FLOAT_MIN()) return false; point = n * (nv / nn); return true; }
    }
    return true;
  }
  // epa.py:970-1183 func_safe_epa (+ :1251-1301 init from the GJK tetrahedron, :1184-1249 witness points)
  void safe_epa(GjkState& S, int ga, int gb, const V& pa, const Q& qa, const V& pb, const Q& qb) {
    S.npv = 0; S.nfaces = 0; S.nfmap = 0; S.nhz = 0;
    for (int i = 0; i < 4; i++) S.pv[S.npv++] = S.sv[i];
    {
      const int fv[4][3] = {{0, 1, 2}, {0, 3, 1}, {0, 2, 3}, {3, 2, 1}}, fa[4][3] = {{1, 3, 2}, {2, 3, 0}, {0, 3, 1}, {2, 0, 1}};
      for (int i = 0; i < 4; i++) epa_attach_face(S, fv[i][0], fv[i][1], fv[i][2], fa[i][0], fa[i][1], fa[i][2]);
      for (int i = 0; i < 4; i++) { S.fmap[i] = i; S.pf[i].map_idx = i; }
      S.nfmap = 4;
    }
    T upper = gjk_FLOAT_MAX(), upper2 = gjk_FLOAT_MAX() * gjk_FLOAT_MAX(), lower = 0;
    T tolerance = T(1e-6);
    int nearest = -1, prev_nearest = -1;
    const bool discrete = is_discrete(ga) && is_discrete(gb);
    if (discrete) tolerance = EPS;
    for (int k = 0; k < EPA_MAX_ITER; k++) {
      prev_nearest = nearest;
      T lower2 = gjk_FLOAT_MAX() * gjk_FLOAT_MAX();
      for (int i = 0; i < S.nfmap; i++) {
        const int i_f = S.fmap[i];
        if (S.pf[i_f].dist2 < lower2) { lower2 = S.pf[i_f].dist2; nearest = i_f; }
      }
      if (lower2 > upper2 || nearest == -1) { nearest = prev_nearest; break; }
      lower = std::sqrt(lower2);
      const V dir = S.pf[nearest].normal;
      // epa.py:810-883 func_epa_support(dir, 1.0): plain support pair (no perturbation), appended to the polytope
      if (S.npv >= POLY_MAX_VERTS) break;
      const int wi = S.npv;
      {
        GjkVert& r = S.pv[wi];
        gjk_support_driver(dir, ga, pa, qa, r.obj1, r.id1);
        gjk_support_driver(-dir, gb, pb, qb, r.obj2, r.id2);
        r.mink = r.obj1 - r.obj2;
        S.npv++;
      }
      const V w = S.pv[wi].mink;
      const T upper_k = w.dot(dir);
      if (upper_k < upper) { upper = upper_k; upper2 = upper * upper; }
      if ((upper - lower) < tolerance) break;
      if (discrete) {
        bool repeated = false;
        for (int i = 0; i < S.npv; i++) {
          if (i == wi) continue;
          if (S.pv[i].id1 == S.pv[wi].id1 && S.pv[i].id2 == S.pv[wi].id2) { repeated = true; break; }
        }
        if (repeated) break;
      }
      epa_horizon(S, w, nearest);
      if (S.nhz < 3) { nearest = -1; break; }
      const int nfaces = S.nfaces, nedges = S.nhz;
      if (nfaces + nedges >= POLY_MAX_FACES) break;
      bool attach_ok = true;
      for (int i = 0; i < nedges; i++) {
        const int i_f0 = nfaces + i, i_f1 = nfaces + (i + 1) % nedges;
        const int hf = S.hz_face[i], he = S.hz_edge[i];
        const int hv1 = S.pf[hf].v[he], hv2 = S.pf[hf].v[(he + 1) % 3];
        S.pf[hf].adj[he] = i_f0;
        const int adj0 = i > 0 ? i_f0 - 1 : nfaces + nedges - 1;
        if (!epa_attach_face(S, wi, hv2, hv1, i_f1, hf, adj0)) { attach_ok = false; break; }
        const T d2 = S.pf[S.nfaces - 1].dist2;
        if (d2 >= lower2 - EPS && d2 <= upper2 + EPS) { S.fmap[S.nfmap] = i_f0; S.pf[i_f0].map_idx = S.nfmap; S.nfmap++; }
      }
      if (!attach_ok) { nearest = -1; break; }
      S.nhz = 0;
      if (S.nfmap == 0 || nearest == -1) { nearest = -1; break; }
    }
    S.n_witness = 0; S.distance = 0;
    if (nearest != -1) {
      const GjkFace& f = S.pf[nearest];
      const V f1 = S.pv[f.v[0]].mink, f2 = S.pv[f.v[1]].mink, f3 = S.pv[f.v[2]].mink;
      V proj;
      project_origin_to_plane(f1, f2, f3, proj);
      const V lam = tri_affine_coords(proj, f1, f2, f3);
      const V back = f1 * lam[0] + f2 * lam[1] + f3 * lam[2];
      const T err = (proj - back).norm();
      const T e12 = (f1 - f2).dot(f1 - f2), e23 = (f2 - f3).dot(f2 - f3), e31 = (f3 - f1).dot(f3 - f1);
      const T inv_len = T(1) / std::sqrt(std::max(std::max(e12, e23), std::max(e31, gjk_FLOAT_MIN() * gjk_FLOAT_MIN())));
      if (!(err * inv_len > T(1e-4) && err > T(1e-4))) {
        S.w1 = S.pv[f.v[0]].obj1 * lam[0] + S.pv[f.v[1]].obj1 * lam[1] + S.pv[f.v[2]].obj1 * lam[2];
        S.w2 = S.pv[f.v[0]].obj2 * lam[0] + S.pv[f.v[1]].obj2 * lam[1] + S.pv[f.v[2]].obj2 * lam[2];
        S.n_witness = 1;
        S.distance = -std::sqrt(f.dist2);
      }
    }
  }
  // gjk.py:170-446 func_gjk_contact (non-MuJoCo branch): is_col, penetration, contact position and normal
  void gjk_contact(int ga, int gb, const V& pa, const Q& qa, const V& pb, const Q& qb, bool& is_col, V& normal, T& pen, V& pos) {
    GjkState* Sp = new GjkState();
    GjkState& S = *Sp;
    if (safe_gjk(S, ga, gb, pa, qa, pb, qb)) safe_epa(S, ga, gb, pa, qa, pb, qb);
    is_col = S.distance < 0;
    pen = is_col ? -S.distance : T(0);
    int n_contacts = 0;
    if (is_col && S.n_witness > 0) {
      const V d = S.w2 - S.w1;
      const T len = d.norm();
      if (!(len < gjk_FLOAT_MIN())) { pos = (S.w1 + S.w2) * T(0.5); normal = d / len; n_contacts = 1; }
    }
    if (n_contacts == 0) { is_col = false; pen = 0; }
    delete Sp;
  }

  // CT:181-222 func_add_contact
  void add_contact(Env<T>& e, int ga, int gb, const V& normal, const V& pos, T pen, int i_pair) {
    int c = e.n_contacts;
    if (c < m.max_contact_pairs) {
      Contact<T>& ct = e.contacts[c];
      T fa = m.geoms_friction[ga] * e.friction_ratio[ga], fb = m.geoms_friction[gb] * e.friction_ratio[gb];
      ct.geom_a = ga; ct.geom_b = gb; ct.normal = normal; ct.pos = pos; ct.penetration = pen;
      ct.friction = std::max(std::max(fa, fb), T(1e-2));
      for (int k = 0; k < 7; k++) ct.sol_params[k] = T(0.5) * (m.geoms_sol_params[7 * ga + k] + m.geoms_sol_params[7 * gb + k]);
      ct.link_a = m.geoms_link[ga]; ct.link_b = m.geoms_link[gb]; ct.pair_idx = i_pair;
      ct.force = V();
      e.n_contacts = c + 1;
    } else {
      e.errno_ |= ERR_CONTACT_OVERFLOW;
    }
  }
  V aabb_corner(int g, int c) const { return V(m.geoms_init_aabb[(g * 8 + c) * 3], m.geoms_init_aabb[(g * 8 + c) * 3 + 1], m.geoms_init_aabb[(g * 8 + c) * 3 + 2]); }
  // CT:317-336 func_compute_tolerance
  T compute_tolerance(int ga, int gb, T tol) {
    T size = (aabb_corner(gb, 7) - aabb_corner(gb, 0)).norm();
    if (m.geoms_type[ga] != GEOM_PLANE) size = std::min((aabb_corner(ga, 7) - aabb_corner(ga, 0)).norm(), size);
    return T(0.5) * tol * size;
  }
  // CT:353-412 func_contact_orthogonals (non-mujoco branch)
  void contact_orthogonals(Env<T>& e, int ga, int gb, const V& normal, V& a0, V& a1) {
    int g = gb;
    if (m.geoms_type[ga] != GEOM_PLANE) {
      V sa = aabb_corner(ga, 7), sb = aabb_corner(gb, 7);
      g = (sa.x * sa.y * sa.z < sb.x * sb.y * sb.z) ? ga : gb;
    }
    M3<T> rot = quat_to_R(e.i_quat[m.geoms_link[g]], EPS);
    int axis_idx = 0;
    T amax = 0;
    for (int i = 0; i < 3; i++) { T a = std::abs(rot.col(i).dot(normal)); if (a > amax) { amax = a; axis_idx = i; } }
    axis_idx = (axis_idx + 1) % 3;
    a0 = rot.col(axis_idx);
    a0 = (a0 - normal * normal.dot(a0)).normalized();
    a1 = normal.cross(a0);
  }
  // CT:415-436 func_rotate_frame
  void rotate_frame(const V& pos, const Q& quat, const V& cpos, const Q& qrot, V& npos, Q& nquat) {
    nquat = transform_quat_by_quat(quat, qrot);
    V rel = cpos - pos;
    V vec = transform_by_quat(rel, qrot) - rel;
    npos = pos - vec;
  }

  // collider/utils.py:9-65 func_closest_points_on_segments (Ericson, Real-Time Collision Detection 5.1.9)
  void closest_points_on_segments(const V& p1, const V& p2, const V& q1, const V& q2, V& ca, V& cb) {
    V da = p2 - p1, db = q2 - q1, r = p1 - q1;
    T a = da.dot(da), b = da.dot(db), c = db.dot(db), d = da.dot(r), f = db.dot(r);
    T denom = a * c - b * b;
    T sp = 0, tp = 0;
    if (denom < EPS) {
      sp = 0;
      tp = c > EPS ? std::min(std::max(f / c, T(0)), T(1)) : T(0);
    } else {
      sp = (b * f - c * d) / denom;
      tp = (a * f - b * d) / denom;
      sp = std::min(std::max(sp, T(0)), T(1));
      tp = std::min(std::max(c > EPS ? (b * sp + f) / c : T(0), T(0)), T(1));
      T s_new = std::min(std::max(a > EPS ? (b * tp - d) / a : T(0), T(0)), T(1));
      if (a > EPS) sp = s_new;
    }
    ca = p1 + da * sp;
    cb = q1 + db * tp;
  }
  // collider/capsule_contact.py:10-98 func_capsule_capsule_contact
  void capsule_capsule_contact(int ga, int gb, const V& pa, const Q& qa, const V& pb, const Q& qb, bool& is_col, V& normal, V& cpos, T& pen) {
    T ra = m.geoms_data[7 * ga], hla = T(0.5) * m.geoms_data[7 * ga + 1];
    T rb = m.geoms_data[7 * gb], hlb = T(0.5) * m.geoms_data[7 * gb + 1];
    V axa = transform_by_quat(V(0, 0, 1), qa), axb = transform_by_quat(V(0, 0, 1), qb);
    V P1 = pa - axa * hla, P2 = pa + axa * hla, Q1 = pb - axb * hlb, Q2 = pb + axb * hlb;
    V Pa, Pb;
    closest_points_on_segments(P1, P2, Q1, Q2, Pa, Pb);
    V diff = Pa - Pb;
    T dist_sq = diff.dot(diff), cr = ra + rb;
    is_col = false; normal = V(1, 0, 0); cpos = V(); pen = 0;
    if (dist_sq < cr * cr) {
      is_col = true;
      T dist = std::sqrt(dist_sq);
      if (dist > EPS) normal = diff / dist;
      else {
        V nd = axa.cross(axb);
        T ndl = nd.dot(nd);
        if (ndl > EPS) normal = nd / ndl;
        else normal = std::abs(axa.x) < T(0.9) ? V(0, -axa.z, axa.y) : V(axa.z, 0, -axa.x);
      }
      pen = cr - dist;
      cpos = Pa - normal * ra + normal * (T(0.5) * pen);
    }
  }
  // collider/capsule_contact.py:102-196 func_sphere_capsule_contact (ga = sphere, gb = capsule after the type sort)
  void sphere_capsule_contact(int ga, int gb, const V& pa, const Q& qa, const V& pb, const Q& qb, bool& is_col, V& normal, V& cpos, T& pen) {
    T dir = 1;
    V sc = pa, cc = pb;
    Q cq = qb;
    if (m.geoms_type[gb] == GEOM_SPHERE) { std::swap(ga, gb); sc = pb; cc = pa; cq = qa; dir = -1; }
    T rs = m.geoms_data[7 * ga], rc = m.geoms_data[7 * gb], hl = T(0.5) * m.geoms_data[7 * gb + 1];
    V ax = transform_by_quat(V(0, 0, 1), cq);
    V P1 = cc - ax * hl, P2 = cc + ax * hl, seg = P2 - P1;
    T l2 = seg.dot(seg), t = T(0.5);
    if (l2 > EPS) t = std::min(std::max((sc - P1).dot(seg) / l2, T(0)), T(1));
    V closest = P1 + seg * t, diff = sc - closest;
    T dist_sq = diff.dot(diff), cr = rs + rc;
    is_col = false; normal = V(1, 0, 0); cpos = V(); pen = 0;
    if (dist_sq < cr * cr) {
      is_col = true;
      T dist = std::sqrt(dist_sq);
      if (dist > EPS) normal = diff / dist;
      else normal = std::abs(ax.x) < T(0.9) ? V(0, -ax.z, ax.y) : V(ax.z, 0, -ax.x);
      pen = cr - dist;
      cpos = sc - normal * (rs - T(0.5) * pen);
    }
    normal = normal * dir;
  }
  // NP:519-970 func_convex_convex_contact (MPR ccd with the exact-penetration fallback of NP:726-733)
  void convex_convex_contact(Env<T>& e, int ga, int gb) {
    if (m.geoms_type[ga] == GEOM_PLANE && m.geoms_type[gb] == GEOM_BOX) return;
    if (cfg.box_box_detection && m.geoms_type[ga] == GEOM_BOX && m.geoms_type[gb] == GEOM_BOX) return;  // NP:2028-2032
    if (m.geoms_type[gb] == GEOM_TERRAIN) return;  // NP:2024-2027: terrain pairs go to func_narrow_phase_any_vs_terrain
    bool multi_contact = cfg.enable_multi_contact && m.geoms_type[ga] != GEOM_SPHERE && m.geoms_type[ga] != GEOM_ELLIPSOID &&
                         m.geoms_type[gb] != GEOM_SPHERE && m.geoms_type[gb] != GEOM_ELLIPSOID;
    T tolerance = compute_tolerance(ga, gb, (T)cfg.mc_tolerance);
    V ga_pos0 = e.g_pos[ga], gb_pos0 = e.g_pos[gb];
    Q ga_quat0 = e.g_quat[ga], gb_quat0 = e.g_quat[gb];
    V ga_pos = ga_pos0, gb_pos = gb_pos0;
    Q ga_quat = ga_quat0, gb_quat = gb_quat0;
    bool is_col_0 = false, is_col = false;
    T pen_0 = 0, pen = 0;
    V normal_0, cpos_0, normal, cpos, axis_0, axis_1;
    Q qrot(0, 0, 0, 0);
    int n_con = 0;
    int i_pair = m.pair_idx[(ga > gb ? gb : ga) * m.nG + (ga > gb ? ga : gb)];
    for (int i_det = 0; i_det < 5; i_det++) {
      if (multi_contact && is_col_0) {
        V axis = axis_0 * T(2 * (i_det % 2) - 1) + axis_1 * T(1 - 2 * ((i_det / 2) % 2));
        qrot = rotvec_to_quat(axis * (T)cfg.mc_perturbation, EPS);
        rotate_frame(ga_pos0, ga_quat0, cpos_0, qrot, ga_pos, ga_quat);
        rotate_frame(gb_pos0, gb_quat0, cpos_0, inv_quat(qrot), gb_pos, gb_quat);
      }
      if ((multi_contact && is_col_0) || i_det == 0) {
        const int ta = m.geoms_type[ga], tb = m.geoms_type[gb];
        if (ta == GEOM_CAPSULE && tb == GEOM_CAPSULE) {  // NP:637-648
          capsule_capsule_contact(ga, gb, ga_pos, ga_quat, gb_pos, gb_quat, is_col, normal, cpos, pen);
        } else if ((ta == GEOM_SPHERE && tb == GEOM_CAPSULE) || (ta == GEOM_CAPSULE && tb == GEOM_SPHERE)) {  // NP:649-661
          sphere_capsule_contact(ga, gb, ga_pos, ga_quat, gb_pos, gb_quat, is_col, normal, cpos, pen);
        } else if (m.geoms_type[ga] == GEOM_PLANE) {
          V pdir(m.geoms_data[7 * ga], m.geoms_data[7 * ga + 1], m.geoms_data[7 * ga + 2]);
          pdir = transform_by_quat(pdir, ga_quat);
          normal = -pdir.normalized();
          V v1 = support_driver(normal, gb, gb_pos, gb_quat);
          pen = normal.dot(v1 - ga_pos);
          cpos = v1 - normal * (T(0.5) * pen);
          is_col = pen > 0;
        } else {
          bool prefer_gjk = cfg.use_gjk_collision != 0;  // NP:604-607
          if (!prefer_gjk) {
            V normal_ws = e.contact_cache_normal[i_pair];
            bool guess = std::abs(normal_ws.x) > EPS || std::abs(normal_ws.y) > EPS || std::abs(normal_ws.z) > EPS;
            bool updated = false;
            Simplex last_simplex;
            bool last_simplex_valid = false;
            for (int i_mpr = 0; i_mpr < 2; i_mpr++) {
              if (i_mpr == 1 && i_det == 0 && !is_col && guess) { normal_ws = V(); guess = false; updated = false; }
              if (!updated) { mpr_contact(ga, gb, normal_ws, ga_pos, ga_quat, gb_pos, gb_quat, is_col, normal, pen, cpos, last_simplex, last_simplex_valid); updated = true; }
            }
            // NP:726-733: MPR's answer is replaced by the exact one when the overlap is large or no direction was cached
            if (cfg.gjk_fallback && is_col && pen > tolerance) {
              prefer_gjk = !guess || (T(cfg.mc_tolerance) * pen >= T(cfg.mpr_to_gjk_overlap_ratio) * tolerance);
              if (prefer_gjk && cfg.gjk_fallback == 2) {  // round-1 approximation kept for A/B: EPA seeded with MPR's final portal
                V n2, c2;
                T p2 = 0;
                if (epa_penetration(last_simplex, ga, gb, ga_pos, ga_quat, gb_pos, gb_quat, n2, p2, c2)) { normal = n2; pen = p2; cpos = c2; is_col = p2 > 0; }
                prefer_gjk = false;
              }
            }
          }
          // the reference's path: a fresh GJK distance query, then EPA from its tetrahedron (NP:736-850)
          if (prefer_gjk) gjk_contact(ga, gb, ga_pos, ga_quat, gb_pos, gb_quat, is_col, normal, pen, cpos);
        }
      }
      if (i_det == 0) {
        is_col_0 = is_col; normal_0 = normal; pen_0 = pen; cpos_0 = cpos;
        if (is_col_0) {
          add_contact(e, ga, gb, normal_0, cpos_0, pen_0, i_pair);
          if (multi_contact) { contact_orthogonals(e, ga, gb, normal, axis_0, axis_1); n_con = 1; }
          e.contact_cache_normal[i_pair] = normal;
        } else {
          e.contact_cache_normal[i_pair] = V();
        }
      } else if (multi_contact && is_col) {
        V cpa = transform_by_quat((cpos - normal * (T(0.5) * pen)) - cpos_0, inv_quat(qrot)) + cpos_0;
        V cpb = transform_by_quat((cpos + normal * (T(0.5) * pen)) - cpos_0, qrot) + cpos_0;
        cpos = (cpa + cpb) * T(0.5);
        V twist = normal.cross(normal_0);
        T mcp = (T)cfg.mc_perturbation;
        for (int k = 0; k < 3; k++) twist[k] = std::min(std::max(twist[k], -mcp), mcp);
        normal = normal + twist.cross(normal);
        pen = normal.dot(cpb - cpa);
        bool repeated = false;
        for (int c = 0; c < n_con; c++) {
          if (!repeated) {
            int idx_prev = e.n_contacts - 1 - c;
            if ((cpos - e.contacts[idx_prev].pos).norm() < tolerance) repeated = true;
          }
        }
        if (!repeated && pen > -tolerance) {
          pen = std::max(pen, T(0));
          add_contact(e, ga, gb, normal, cpos, pen, i_pair);
          n_con++;
        }
      }
    }
  }

  // BOX:24-96 func_plane_box_contact
  void plane_box_contact(Env<T>& e, int ga, int gb, int i_pair) {
    V pdir(m.geoms_data[7 * ga], m.geoms_data[7 * ga + 1], m.geoms_data[7 * ga + 2]);
    pdir = transform_by_quat(pdir, e.g_quat[ga]);
    V normal = -pdir.normalized();
    V v1 = support_driver(normal, gb, e.g_pos[gb], e.g_quat[gb]);
    T pen = normal.dot(v1 - e.g_pos[ga]);
    if (pen > 0) {
      V cpos = v1 - normal * (T(0.5) * pen);
      add_contact(e, ga, gb, normal, cpos, pen, i_pair);
      if (cfg.enable_multi_contact) {
        int n_con = 1;
        V cpos_0 = cpos;
        T tolerance = compute_tolerance(ga, gb, (T)cfg.mc_tolerance);
        for (int iv = m.geoms_vert_start[gb]; iv < m.geoms_vert_end[gb]; iv++) {
          if (n_con < m.n_contacts_per_pair) {
            V pc = transform_by_trans_quat(m.v3(m.verts, iv), e.g_pos[gb], e.g_quat[gb]);
            pen = normal.dot(pc - e.g_pos[ga]);
            if (pen > 0) {
              cpos = pc - normal * (T(0.5) * pen);
              if ((cpos - cpos_0).norm() > tolerance) { add_contact(e, ga, gb, normal, cpos, pen, i_pair); n_con++; }
            }
          }
        }
      }
    }
  }

  // BOX:100-785 func_box_box_contact: candidates from box_box_detect (oracle/box_box.hpp); duplicate filter and contact records here
  void box_box_contact(Env<T>& e, int ga, int gb, int i_pair) {
    BoxBoxContacts<T> bb;
    const V ha(m.geoms_data[7 * ga] / 2, m.geoms_data[7 * ga + 1] / 2, m.geoms_data[7 * ga + 2] / 2);
    const V hb(m.geoms_data[7 * gb] / 2, m.geoms_data[7 * gb + 1] / 2, m.geoms_data[7 * gb + 2] / 2);
    if (!box_box_detect(e.g_pos[ga], e.g_quat[ga], ha, e.g_pos[gb], e.g_quat[gb], hb, EPS, bb)) return;
    int n_added = 0;
    const int n_start = e.n_contacts;
    for (int i = 0; i < bb.n; i++) {
      if (n_added >= m.n_contacts_per_pair) continue;
      bool is_valid = true;
      for (int j_ = 0; j_ < n_added; j_++) {  // BOX:437-441, 763-767
        const int j = n_start + j_;
        if (j >= m.max_contact_pairs) continue;
        const V d = bb.pos[i] - e.contacts[j].pos;
        if (std::abs(d.x) < EPS && std::abs(d.y) < EPS && std::abs(d.z) < EPS) is_valid = false;
      }
      if (is_valid) { add_contact(e, ga, gb, -bb.normal, bb.pos[i], -bb.dist[i], i_pair); n_added++; }
    }
  }

  // NP:346-497 func_contact_mpr_terrain: geom `ga` against the prisms of the height-field cells under its bounding box
  void terrain_contact(Env<T>& e, int ga, int gb) {
    const V ga_pos = e.g_pos[ga], gb_pos = e.g_pos[gb];
    const Q ga_quat = e.g_quat[ga], gb_quat = e.g_quat[gb];
    const T margin = 0;
    const T tolerance = compute_tolerance(ga, gb, (T)cfg.mc_tolerance);
    // into the terrain's frame (gu.qd_transform_pos_quat_by_trans_quat with the inverse terrain rotation)
    const Q inv_b = inv_quat(gb_quat);
    const V pa = transform_by_quat(ga_pos - gb_pos, inv_b);
    const Q qa = transform_quat_by_quat(ga_quat, inv_b);
    const V pb;
    const Q qb;
    const V center_a = transform_by_trans_quat(m.v3(m.geoms_center, ga), pa, qa);
    const std::vector<T>& tmm = m.terrain_xyz_maxmin;
    V* prism = terrain_prism();
    T xyz[6];
    for (int i_axis = 0; i_axis < 3; i_axis++)
      for (int i_m = 0; i_m < 2; i_m++) {
        V direction;
        direction[i_axis] = i_m == 0 ? T(1) : T(-1);
        const V v1 = support_driver(direction, ga, pa, qa);
        xyz[3 * i_m + i_axis] = v1[i_axis];
      }
    bool is_return = false;
    for (int i = 0; i < 3; i++) {
      prism[i].z = tmm[5];
      if (tmm[i] < xyz[i + 3] - margin || tmm[i + 3] > xyz[i] + margin) is_return = true;
    }
    if (is_return) return;
    const T sh = m.terrain_scale[0];
    int r_min = (int)std::floor((xyz[3] - tmm[3]) / sh), r_max = (int)std::ceil((xyz[0] - tmm[3]) / sh);
    int c_min = (int)std::floor((xyz[4] - tmm[4]) / sh), c_max = (int)std::ceil((xyz[1] - tmm[4]) / sh);
    r_min = std::max(0, r_min); c_min = std::max(0, c_min);
    r_max = std::min(m.terrain_rc[0] - 1, r_max); c_max = std::min(m.terrain_rc[1] - 1, c_max);
    const int ncols = m.terrain_rc[1];
    int n_con = 0;
    for (int r = r_min; r < r_max; r++) {
      int nvert = 0;
      for (int c = c_min; c <= c_max; c++)
        for (int i = 0; i < 2; i++) {
          if (n_con >= m.n_contacts_per_pair) continue;
          nvert++;
          {  // NP:499-517 func_add_prism_vert
            const T x = sh * (r + i) + tmm[3], y = sh * c + tmm[4], z = m.terrain_hf[(size_t)(r + i) * ncols + c] + margin;
            prism[0] = prism[1]; prism[1] = prism[2]; prism[3] = prism[4]; prism[4] = prism[5];
            prism[2].x = x; prism[5].x = x; prism[2].y = y; prism[5].y = y; prism[5].z = z;
          }
          if (nvert > 2 && (prism[3].z >= xyz[5] || prism[4].z >= xyz[5] || prism[5].z >= xyz[5])) {
            V center_b;
            for (int k = 0; k < 6; k++) center_b = center_b + prism[k];
            center_b = center_b / T(6);
            bool is_col, sv;
            V normal, cpos;
            T pen;
            Simplex simplex;
            mpr_contact_from_centers(ga, gb, center_a, center_b, pa, qa, pb, qb, is_col, normal, pen, cpos, simplex, sv);
            if (is_col) {
              normal = transform_by_quat(normal, gb_quat);
              cpos = transform_by_quat(cpos, gb_quat) + gb_pos;
              bool valid = true;
              const int i_c = e.n_contacts;
              for (int j = 0; j < n_con; j++)
                if (i_c - j - 1 >= 0 && (cpos - e.contacts[i_c - j - 1].pos).norm() < tolerance) { valid = false; break; }
              if (valid) {
                const int i_pair = m.pair_idx[(ga > gb ? gb : ga) * m.nG + (ga > gb ? ga : gb)];
                add_contact(e, ga, gb, normal, cpos, pen, i_pair);
                n_con++;
              }
            }
          }
        }
    }
  }

  // CT:439-538 func_sort_contacts (stable insertion sort of pair groups by x of the group's first contact)
  void sort_contacts(Env<T>& e) {
    int n = e.n_contacts;
    std::vector<T> key(n);
    std::vector<int> idx(n);
    T gk = 0;
    for (int i = 0; i < n; i++) {
      if (i == 0 || e.contacts[i].geom_a != e.contacts[i - 1].geom_a || e.contacts[i].geom_b != e.contacts[i - 1].geom_b) gk = e.contacts[i].pos.x;
      key[i] = gk; idx[i] = i;
    }
    for (int i = 1; i < n; i++) {
      T ck = key[i];
      if (key[i - 1] <= ck) continue;
      int ci = idx[i], j = i - 1;
      while (j >= 0) { if (key[j] <= ck) break; key[j + 1] = key[j]; idx[j + 1] = idx[j]; j--; }
      key[j + 1] = ck; idx[j + 1] = ci;
    }
    std::vector<Contact<T>> tmp(e.contacts.begin(), e.contacts.begin() + n);
    for (int i = 0; i < n; i++) e.contacts[i] = tmp[idx[i]];
  }

  // collider.py:684-808 detection()
  void detection(Env<T>& e) {
    update_aabbs(e);
    if (m.nP == 0) return;
    broad_phase(e);
    bool has_cc = false;
    for (int p = 0; p < m.nP; p++) {
      int ta = m.geoms_type[m.pairs[2 * p]], tb = m.geoms_type[m.pairs[2 * p + 1]];
      if (ta > tb) std::swap(ta, tb);
      const bool specialized = (ta == GEOM_PLANE && tb == GEOM_BOX) || (cfg.box_box_detection && ta == GEOM_BOX && tb == GEOM_BOX);  // collider.py:456-459
      if (!specialized && tb != GEOM_TERRAIN) has_cc = true;  // terrain is not convex (collider.py:449-459)
    }
    for (int i = 0; i < e.n_broad; i++) {  // NP:1989-2060
      int ga = e.broad_pairs[2 * i], gb = e.broad_pairs[2 * i + 1];
      if (m.geoms_type[ga] > m.geoms_type[gb]) std::swap(ga, gb);
      convex_convex_contact(e, ga, gb);
    }
    for (int i = 0; i < e.n_broad; i++) {  // NP:2135-2173
      int ga = e.broad_pairs[2 * i], gb = e.broad_pairs[2 * i + 1];
      if (m.geoms_type[ga] > m.geoms_type[gb]) std::swap(ga, gb);
      if (m.geoms_type[ga] == GEOM_PLANE && m.geoms_type[gb] == GEOM_BOX) plane_box_contact(e, ga, gb, i);
      if (cfg.box_box_detection && m.geoms_type[ga] == GEOM_BOX && m.geoms_type[gb] == GEOM_BOX) box_box_contact(e, ga, gb, i);  // NP:2175-2189
    }
    for (int i = 0; i < e.n_broad; i++) {  // NP:2193-2239 func_narrow_phase_any_vs_terrain
      int ga = e.broad_pairs[2 * i], gb = e.broad_pairs[2 * i + 1];
      if (m.geoms_type[ga] == GEOM_TERRAIN) std::swap(ga, gb);
      if (m.geoms_type[gb] == GEOM_TERRAIN) terrain_contact(e, ga, gb);
    }
    if (has_cc && cfg.sort_contacts) sort_contacts(e);  // collider.py:804-808 (GPU split path)
  }

  // ------------------------------------------------------------------ constraint rows
  T* jrow(Env<T>& e, int c) { return &e.jac[(size_t)c * m.nD]; }

  // CS:800-877 func_equality_joint
  void equality_joint(Env<T>& e, int i_e) {
    const T* sp = &m.eq_sol_params[7 * i_e];
    int j1 = m.eq_obj1[i_e], j2 = m.eq_obj2[i_e];
    int q1 = m.joints_q_start[j1], q2 = m.joints_q_start[j2], d1 = m.joints_dof_start[j1], d2 = m.joints_dof_start[j2];
    int c = e.n_con++;
    e.n_eq++;
    T* J = jrow(e, c);
    for (int d = 0; d < m.nD; d++) J[d] = 0;
    T pos1 = e.qpos[q1], pos2 = e.qpos[q2], ref1 = m.qpos0[q1], ref2 = m.qpos0[q2];
    T diff = pos2 - ref2, pos = pos1 - ref1, deriv = 0;
    for (int k = 0; k < 5; k++) {
      T dp = std::pow(diff, (T)k);
      if (k == 0) dp = 1;
      pos = pos - dp * m.eq_data[11 * i_e + k];
      if (k < 4) deriv = deriv + m.eq_data[11 * i_e + k + 1] * dp * T(k + 1);
    }
    J[d1] = 1;
    J[d2] = -deriv;
    T jac_qvel = J[d1] * e.vel[d1] + J[d2] * e.vel[d2];
    T invweight = m.dofs_invweight[d1] + m.dofs_invweight[d2];
    T imp, aref;
    imp_aref(sp, -std::abs(pos), jac_qvel, pos, imp, aref);
    T diag = std::max(invweight * (T(1) - imp) / imp, EPS);
    e.diag[c] = diag; e.aref[c] = aref; e.efc_D[c] = T(1) / diag;
  }

  // CS:688-797 func_equality_connect
  void equality_connect(Env<T>& e, int i_e) {
    const T* sp = &m.eq_sol_params[7 * i_e];
    int l1 = m.eq_obj1[i_e], l2 = m.eq_obj2[i_e];
    V a1(m.eq_data[11 * i_e], m.eq_data[11 * i_e + 1], m.eq_data[11 * i_e + 2]), a2(m.eq_data[11 * i_e + 3], m.eq_data[11 * i_e + 4], m.eq_data[11 * i_e + 5]);
    V g1 = transform_by_trans_quat(a1, e.l_pos[l1], e.l_quat[l1]), g2 = transform_by_trans_quat(a2, e.l_pos[l2], e.l_quat[l2]);
    T invweight = m.links_invweight[2 * l1] + m.links_invweight[2 * l2];
    for (int i3 = 0; i3 < 3; i3++) {
      int c = e.n_con++;
      e.n_eq++;
      T* J = jrow(e, c);
      for (int d = 0; d < m.nD; d++) J[d] = 0;
      T jac_qvel = 0;
      for (int ab = 0; ab < 2; ab++) {
        T sign = ab == 0 ? T(1) : T(-1);
        int link = ab == 0 ? l1 : l2;
        V pos = ab == 0 ? g1 : g2;
        while (link > -1) {
          for (int d = m.links_dof_end[link] - 1; d >= m.links_dof_start[link]; d--) {
            V tpos = pos - e.root_COM[link];
            V vel = e.cdof_vel[d] - tpos.cross(e.cdof_ang[d]);
            T jac = sign * vel[i3];
            jac_qvel += jac * e.vel[d];
            J[d] += jac;
          }
          link = m.links_parent[link];
        }
      }
      V pd = g1 - g2;
      T imp, aref;
      imp_aref(sp, -pd.norm(), jac_qvel, pd[i3], imp, aref);
      T diag = std::max(invweight * (T(1) - imp) / imp, EPS);
      e.diag[c] = diag; e.aref[c] = aref; e.efc_D[c] = T(1) / diag;
    }
  }

  // CS:984-1173 func_equality_weld: 3 position rows (anchor1 = data[3:6] in body 1, anchor2 = data[0:3] in body 2) and
  // 3 orientation rows (relpose = data[6:10], torquescale = data[10]); translational / rotational invweight components
  void equality_weld(Env<T>& e, int i_e) { weld_rows(e, m.eq_obj1[i_e], m.eq_obj2[i_e], &m.eq_data[11 * i_e], &m.eq_sol_params[7 * i_e]); }
  void weld_rows(Env<T>& e, int l1, int l2, const T* dat, const T* sp) {
    V a1(dat[3], dat[4], dat[5]), a2(dat[0], dat[1], dat[2]);
    Q relpose(dat[6], dat[7], dat[8], dat[9]);
    T torquescale = dat[10];
    V g1 = transform_by_trans_quat(a1, e.l_pos[l1], e.l_quat[l1]), g2 = transform_by_trans_quat(a2, e.l_pos[l2], e.l_quat[l2]);
    V pos_error = g1 - g2;
    Q q = quat_mul(e.l_quat[l1], relpose);
    Q inv_q2 = inv_quat(e.l_quat[l2]);
    Q eq = quat_mul(inv_q2, q);
    V rot_error = V(eq.x, eq.y, eq.z) * torquescale;
    T pos_imp = std::sqrt(pos_error.dot(pos_error) + rot_error.dot(rot_error));
    T iw0 = m.links_invweight[2 * l1] + m.links_invweight[2 * l2], iw1 = m.links_invweight[2 * l1 + 1] + m.links_invweight[2 * l2 + 1];
    for (int i3 = 0; i3 < 3; i3++) {
      int c = e.n_con++;
      e.n_eq++;
      T* J = jrow(e, c);
      for (int d = 0; d < m.nD; d++) J[d] = 0;
      T jac_qvel = 0;
      for (int ab = 0; ab < 2; ab++) {
        T sign = ab == 0 ? T(1) : T(-1);
        int link = ab == 0 ? l1 : l2;
        V pos = ab == 0 ? g1 : g2;
        while (link > -1) {
          for (int d = m.links_dof_end[link] - 1; d >= m.links_dof_start[link]; d--) {
            V tpos = pos - e.root_COM[link];
            V vel = e.cdof_vel[d] - tpos.cross(e.cdof_ang[d]);
            T jac = sign * vel[i3];
            jac_qvel += jac * e.vel[d];
            J[d] += jac;
          }
          link = m.links_parent[link];
        }
      }
      T imp, aref;
      imp_aref(sp, -pos_imp, jac_qvel, pos_error[i3], imp, aref);
      T diag = std::max(iw0 * (T(1) - imp) / imp, EPS);
      e.diag[c] = diag; e.aref[c] = aref; e.efc_D[c] = T(1) / diag;
    }
    int c0 = e.n_con;
    e.n_con += 3; e.n_eq += 3;
    for (int k = 0; k < 3; k++) { T* J = jrow(e, c0 + k); for (int d = 0; d < m.nD; d++) J[d] = 0; }
    for (int ab = 0; ab < 2; ab++) {
      T sign = ab == 0 ? T(1) : T(-1);
      int link = ab == 0 ? l1 : l2;
      while (link > -1) {
        for (int d = m.links_dof_end[link] - 1; d >= m.links_dof_start[link]; d--)
          for (int k = 0; k < 3; k++) jrow(e, c0 + k)[d] += sign * e.cdof_ang[d][k];
        link = m.links_parent[link];
      }
    }
    T jac_qvel[3] = {0, 0, 0};
    for (int d = 0; d < m.nD; d++) {
      V jd(jrow(e, c0)[d], jrow(e, c0 + 1)[d], jrow(e, c0 + 2)[d]);
      // geom.py:224-232 quat_mul_axis(inv_q2, jd), then * q
      Q quat2(-inv_q2.x * jd.x - inv_q2.y * jd.y - inv_q2.z * jd.z, inv_q2.w * jd.x + inv_q2.y * jd.z - inv_q2.z * jd.y,
              inv_q2.w * jd.y + inv_q2.z * jd.x - inv_q2.x * jd.z, inv_q2.w * jd.z + inv_q2.x * jd.y - inv_q2.y * jd.x);
      Q quat3 = quat_mul(quat2, q);
      T v3[3] = {quat3.x, quat3.y, quat3.z};
      for (int k = 0; k < 3; k++) {
        T j = T(0.5) * v3[k] * torquescale;
        jrow(e, c0 + k)[d] = j;
        jac_qvel[k] += j * e.vel[d];
      }
    }
    for (int k = 0; k < 3; k++) {
      T imp, aref;
      imp_aref(sp, -pos_imp, jac_qvel[k], rot_error[k], imp, aref);
      T diag = std::max(iw1 * (T(1) - imp) / imp, EPS);
      e.diag[c0 + k] = diag; e.aref[c0 + k] = aref; e.efc_D[c0 + k] = T(1) / diag;
    }
  }

  // CS:880-937 add_equality_constraints
  void add_equality_constraints(Env<T>& e) {
    e.n_con = 0; e.n_eq = 0;
    for (int i = 0; i < m.nEq; i++) {
      if (m.eq_type[i] == EQ_JOINT) equality_joint(e, i);
      else if (m.eq_type[i] == EQ_CONNECT) equality_connect(e, i);
      else if (m.eq_type[i] == EQ_WELD) equality_weld(e, i);
    }
    for (auto& w : e.dyn_welds) weld_rows(e, w.l1, w.l2, w.data, w.sol);
  }

  // CS:1236-1286 add_frictionloss_constraints
  void add_frictionloss(Env<T>& e) {
    e.n_floss = 0;
    for (int l = 0; l < m.nL; l++)
      for (int j = m.links_joint_start[l]; j < m.links_joint_end[l]; j++)
        for (int d = m.joints_dof_start[j]; d < m.joints_dof_start[j] + m.joints_n_dofs[j]; d++)
          if (m.dofs_frictionloss[d] > EPS) {
            T imp, aref;
            imp_aref(&m.joints_sol_params[7 * j], T(0), e.vel[d], T(0), imp, aref);
            T diag = std::max(m.dofs_invweight[d] * (T(1) - imp) / imp, EPS);
            int c = e.n_con++;
            e.n_floss++;
            e.diag[c] = diag; e.aref[c] = aref; e.efc_D[c] = T(1) / diag; e.efc_frictionloss[c] = m.dofs_frictionloss[d];
            T* J = jrow(e, c);
            for (int k = 0; k < m.nD; k++) J[k] = 0;
            J[d] = 1;
          }
  }

  // CS:579-685 add_collision_constraints
  void add_collision(Env<T>& e) {
    int start = e.n_con;
    for (int ic = 0; ic < e.n_contacts; ic++) {
      const Contact<T>& ct = e.contacts[ic];
      V d1, d2;
      orthogonals(ct.normal, d1, d2);
      T invweight = m.links_invweight[2 * ct.link_a];
      if (ct.link_b > -1) invweight += m.links_invweight[2 * ct.link_b];
      for (int i = 0; i < 4; i++) {
        V d = (i < 2 ? d1 : d2) * T(2 * (i % 2) - 1);
        V n = d * ct.friction - ct.normal;
        int c = start + ic * 4 + i;
        T* J = jrow(e, c);
        for (int k = 0; k < m.nD; k++) J[k] = 0;
        T jac_qvel = 0;
        for (int ab = 0; ab < 2; ab++) {
          T sign = ab == 0 ? T(-1) : T(1);
          int link = ab == 0 ? ct.link_a : ct.link_b;
          while (link > -1) {
            for (int dd = m.links_dof_end[link] - 1; dd >= m.links_dof_start[link]; dd--) {
              V tpos = ct.pos - e.root_COM[link];
              V vel = e.cdof_vel[dd] - tpos.cross(e.cdof_ang[dd]);
              V diff = vel * sign;
              T jac = diff.dot(n);
              jac_qvel += jac * e.vel[dd];
              J[dd] += jac;
            }
            link = m.links_parent[link];
          }
        }
        T imp, aref;
        imp_aref(ct.sol_params, -ct.penetration, jac_qvel, -ct.penetration, imp, aref);
        T diag = invweight + ct.friction * ct.friction * invweight;
        diag *= 2 * ct.friction * ct.friction * (T(1) - imp) / imp;
        diag = std::max(diag, EPS);
        e.diag[c] = diag; e.aref[c] = aref; e.efc_D[c] = T(1) / diag;
      }
    }
    e.n_con += e.n_contacts * 4;
  }

  // CS:1176-1233 add_joint_limit_constraints
  void add_joint_limits(Env<T>& e) {
    for (int l = 0; l < m.nL; l++)
      for (int j = m.links_joint_start[l]; j < m.links_joint_end[l]; j++) {
        if (m.joints_type[j] != JOINT_REVOLUTE && m.joints_type[j] != JOINT_PRISMATIC) continue;
        int q = m.joints_q_start[j], d = m.joints_dof_start[j];
        T pmin = e.qpos[q] - m.dofs_limit[2 * d], pmax = m.dofs_limit[2 * d + 1] - e.qpos[q];
        T pd = std::min(pmin, pmax);
        if (pd < 0) {
          T jac = T((pmin < pmax) * 2 - 1);
          T imp, aref;
          imp_aref(&m.joints_sol_params[7 * j], pd, jac * e.vel[d], pd, imp, aref);
          T diag = std::max(m.dofs_invweight[d] * (T(1) - imp) / imp, EPS);
          int c = e.n_con++;
          e.diag[c] = diag; e.aref[c] = aref; e.efc_D[c] = T(1) / diag;
          T* J = jrow(e, c);
          for (int k = 0; k < m.nD; k++) J[k] = 0;
          J[d] = jac;
        }
      }
  }

  // ------------------------------------------------------------------ constraint solver (monolith semantics)
  // CS:2619-2694 func_update_constraint_batch
  void update_constraint(Env<T>& e) {
    int nD = m.nD, ne = e.n_eq, nef = ne + e.n_floss;
    e.prev_cost = e.cost;
    T cost = 0, gauss = 0;
    for (int c = 0; c < e.n_con; c++) {
      e.prev_active[c] = e.active[c];
      e.active[c] = 1;
      T floss_force = 0;
      if (ne <= c && c < nef) {
        T f = e.efc_frictionloss[c], r = e.diag[c], rf = r * f;
        bool ln = e.Jaref[c] <= -rf, lp = e.Jaref[c] >= rf;
        e.active[c] = !(ln || lp);
        floss_force = T(ln) * f + T(lp) * -f;
        T fc = T(ln) * f * (T(-0.5) * rf - e.Jaref[c]);
        fc = fc + T(lp) * f * (T(-0.5) * rf + e.Jaref[c]);
        cost = cost + fc;
      } else if (nef <= c) {
        e.active[c] = e.Jaref[c] < 0;
      }
      e.efc_force[c] = floss_force + (-e.Jaref[c] * e.efc_D[c] * T(e.active[c]));
    }
    for (int d = 0; d < nD; d++) {
      T q = 0;
      for (int c = 0; c < e.n_con; c++) q = q + e.jac[(size_t)c * nD + d] * e.efc_force[c];
      e.qf_constraint[d] = q;
    }
    for (int d = 0; d < nD; d++) {
      T v = T(0.5) * (e.Ma[d] - e.force[d]) * (e.qacc[d] - e.acc_smooth[d]);
      gauss = gauss + v;
      cost = cost + v;
    }
    for (int c = 0; c < e.n_con; c++) cost = cost + T(0.5) * (e.Jaref[c] * e.Jaref[c] * e.efc_D[c] * T(e.active[c]));
    e.gauss = gauss;
    e.cost = cost;
  }

  // CS:1394-1452 func_hessian_direct_batch + CS:1594-1621 func_cholesky_factor_direct_batch
  void hessian_and_cholesky_direct(Env<T>& e) {
    int nD = m.nD;
    for (int i = 0; i < nD; i++) for (int j = 0; j <= i; j++) e.H[i * nD + j] = 0;
    for (int i = 0; i < nD; i++)
      for (int c = 0; c < e.n_con; c++) {
        T Jci = e.jac[(size_t)c * nD + i];
        if (std::abs(Jci) > EPS)
          for (int j = 0; j <= i; j++) e.H[i * nD + j] = e.H[i * nD + j] + e.jac[(size_t)c * nD + j] * Jci * e.efc_D[c] * T(e.active[c]);
      }
    for (int i_e = 0; i_e < m.nE; i_e++)
      for (int i = m.entities_dof_start[i_e]; i < m.entities_dof_end[i_e]; i++)
        for (int j = m.entities_dof_start[i_e]; j <= i; j++) e.H[i * nD + j] += e.M[i * nD + j];
    for (int i = 0; i < nD; i++) {
      T tmp = e.H[i * nD + i];
      for (int j = 0; j < i; j++) tmp = tmp - e.H[i * nD + j] * e.H[i * nD + j];
      e.H[i * nD + i] = std::sqrt(std::max(tmp, EPS));
      tmp = T(1) / e.H[i * nD + i];
      for (int j = i + 1; j < nD; j++) {
        T dot = 0;
        for (int k = 0; k < i; k++) dot = dot + e.H[j * nD + k] * e.H[i * nD + k];
        e.H[j * nD + i] = (e.H[j * nD + i] - dot) * tmp;
      }
    }
  }

  // CS:1763-1822 changed list + incremental rank-1 update
  bool hessian_incremental(Env<T>& e) {
    int nD = m.nD;
    bool degenerate = false;
    for (int c = 0; c < e.n_con && !degenerate; c++) {
      if (!(e.active[c] ^ e.prev_active[c])) continue;
      T sign = e.active[c] ? T(1) : T(-1);
      T sq = std::sqrt(e.efc_D[c]);
      for (int d = 0; d < nD; d++) e.nt_vec[d] = e.jac[(size_t)c * nD + d] * sq;
      for (int k = 0; k < nD; k++) {
        if (std::abs(e.nt_vec[k]) > EPS) {
          T Lkk = e.H[k * nD + k];
          T tmp = Lkk * Lkk + sign * e.nt_vec[k] * e.nt_vec[k];
          if (tmp < EPS) { degenerate = true; break; }
          T r = std::sqrt(tmp), cc = r / Lkk, cinv = T(1) / cc, s = e.nt_vec[k] / Lkk;
          e.H[k * nD + k] = r;
          for (int i = k + 1; i < nD; i++) e.H[i * nD + k] = (e.H[i * nD + k] + s * e.nt_vec[i] * sign) * cinv;
          for (int i = k + 1; i < nD; i++) e.nt_vec[i] = e.nt_vec[i] * cc - s * e.H[i * nD + k];
        }
      }
    }
    return degenerate;
  }

  // CS:2721-2750 func_update_gradient_batch (+ CS:1892-1910 cholesky solve)
  void update_gradient(Env<T>& e) {
    int nD = m.nD;
    for (int d = 0; d < nD; d++) e.grad[d] = e.Ma[d] - e.force[d] - e.qf_constraint[d];
    if (cfg.solver_type == SOLVER_CG) {
      solve_mass(e, e.grad, e.Mgrad);
    } else {
      for (int i = 0; i < nD; i++) {
        T c = e.grad[i];
        for (int j = 0; j < i; j++) c -= e.H[i * nD + j] * e.Mgrad[j];
        e.Mgrad[i] = c / e.H[i * nD + i];
      }
      for (int i = nD - 1; i >= 0; i--) {
        T c = e.Mgrad[i];
        for (int j = i + 1; j < nD; j++) c -= e.H[j * nD + i] * e.Mgrad[j];
        e.Mgrad[i] = c / e.H[i * nD + i];
      }
    }
  }

  // CS:2935-3059 func_solve_init (non-mujoco-compat branch)
  void solve_init(Env<T>& e) {
    int nD = m.nD;
    for (int d = 0; d < nD; d++) e.qacc[d] = (e.n_con > 0 && e.is_warmstart) ? e.qacc_ws[d] : e.acc_smooth[d];
    for (int i = 0; i < nD; i++) {  // CS:2910-2929 initialize_Ma
      int i_e = m.dofs_entity[i];
      T s = 0;
      for (int j = m.entities_dof_start[i_e]; j < m.entities_dof_end[i_e]; j++) s = s + e.M[i * nD + j] * e.qacc[j];
      e.Ma[i] = s;
    }
    for (int c = 0; c < e.n_con; c++) {  // CS:2887-2907 initialize_Jaref
      T s = -e.aref[c];
      for (int d = 0; d < nD; d++) s = s + e.jac[(size_t)c * nD + d] * e.qacc[d];
      e.Jaref[c] = s;
    }
    update_constraint(e);
    e.improved = e.n_con > 0;
    if (cfg.solver_type == SOLVER_NEWTON) hessian_and_cholesky_direct(e);
    update_gradient(e);
    for (int d = 0; d < nD; d++) e.search[d] = -e.Mgrad[d];
  }

  // ---- line search (CS:2033-2597) ----
  struct LsPoint { T alpha, cost, d0, d1; };
  LsPoint ls_p0(Env<T>& e) {  // CS:2033-2151
    int nD = m.nD, ne = e.n_eq, nef = ne + e.n_floss, n_con = e.n_con;
    for (int i_e = 0; i_e < m.nE; i_e++)
      for (int i = m.entities_dof_start[i_e]; i < m.entities_dof_end[i_e]; i++) {
        T mv = 0;
        for (int j = m.entities_dof_start[i_e]; j < m.entities_dof_end[i_e]; j++) mv = mv + e.M[i * nD + j] * e.search[j];
        e.mv[i] = mv;
      }
    for (int c = 0; c < n_con; c++) {
      T jv = 0;
      for (int d = 0; d < nD; d++) jv = jv + e.jac[(size_t)c * nD + d] * e.search[d];
      e.jv[c] = jv;
    }
    T qg1 = 0, qg2 = 0;
    for (int d = 0; d < nD; d++) {
      qg1 = qg1 + (e.search[d] * e.Ma[d] - e.search[d] * e.force[d]);
      qg2 = qg2 + T(0.5) * e.search[d] * e.mv[d];
    }
    e.quad_gauss[0] = e.gauss; e.quad_gauss[1] = qg1; e.quad_gauss[2] = qg2;
    T t0 = e.gauss, t1 = qg1, t2 = qg2, s0 = 0, s1 = 0, s2 = 0;
    for (int c = 0; c < n_con; c++) {
      T Ja = e.Jaref[c], jv = e.jv[c], D = e.efc_D[c];
      T q0 = D * (T(0.5) * Ja * Ja), q1 = D * (jv * Ja), q2 = D * (T(0.5) * jv * jv);
      if (c < ne) {
        s0 += q0; s1 += q1; s2 += q2; t0 += q0; t1 += q1; t2 += q2;
      } else if (c < nef) {
        T f = e.efc_frictionloss[c], r = e.diag[c], rf = r * f;
        bool ln = Ja <= -rf, lp = Ja >= rf;
        if (ln || lp) { q0 = T(ln) * f * (T(-0.5) * rf - Ja) + T(lp) * f * (T(-0.5) * rf + Ja); q1 = T(ln) * (-f * jv) + T(lp) * (f * jv); q2 = 0; }
        t0 += q0; t1 += q1; t2 += q2;
      } else {
        T act = T(Ja < 0);
        t0 += q0 * act; t1 += q1 * act; t2 += q2 * act;
      }
    }
    e.eq_sum[0] = s0; e.eq_sum[1] = s1; e.eq_sum[2] = s2;
    T hess = 2 * t2;
    if (hess <= 0) hess = EPS;
    e.ls_it = 1;
    return LsPoint{T(0), t0, t1, hess};
  }
  void ls_accumulate(Env<T>& e, T alpha, T& t0, T& t1, T& t2) {  // body shared by CS:2154-2222 and CS:2225-2354
    int ne = e.n_eq, nef = ne + e.n_floss, n_con = e.n_con;
    t0 = e.quad_gauss[0] + e.eq_sum[0]; t1 = e.quad_gauss[1] + e.eq_sum[1]; t2 = e.quad_gauss[2] + e.eq_sum[2];
    for (int c = ne; c < nef; c++) {
      T Ja = e.Jaref[c], jv = e.jv[c], D = e.efc_D[c], f = e.efc_frictionloss[c], r = e.diag[c];
      T q0 = D * (T(0.5) * Ja * Ja), q1 = D * (jv * Ja), q2 = D * (T(0.5) * jv * jv);
      T x = Ja + alpha * jv, rf = r * f;
      bool ln = x <= -rf, lp = x >= rf;
      if (ln || lp) { q0 = T(ln) * f * (T(-0.5) * rf - Ja) + T(lp) * f * (T(-0.5) * rf + Ja); q1 = T(ln) * (-f * jv) + T(lp) * (f * jv); q2 = 0; }
      t0 += q0; t1 += q1; t2 += q2;
    }
    for (int c = nef; c < n_con; c++) {
      T Ja = e.Jaref[c], jv = e.jv[c], D = e.efc_D[c];
      T x = Ja + alpha * jv;
      T act = T(x < 0);
      t0 += D * (T(0.5) * Ja * Ja) * act; t1 += D * (jv * Ja) * act; t2 += D * (T(0.5) * jv * jv) * act;
    }
  }
  LsPoint ls_point(Env<T>& e, T alpha) {  // CS:2154-2222
    T t0, t1, t2;
    ls_accumulate(e, alpha, t0, t1, t2);
    T cost = alpha * alpha * t2 + alpha * t1 + t0, grad = 2 * alpha * t2 + t1, hess = 2 * t2;
    if (hess <= 0) hess = EPS;
    e.ls_it += 1;
    return LsPoint{alpha, cost, grad, hess};
  }
  // CS:2357-2388 update_bracket_no_eval_local
  int update_bracket(LsPoint& p, const LsPoint* cand, T& next_alpha) {
    int flag = 0;
    for (int i = 0; i < 3; i++) {
      if (p.d0 < 0 && cand[i].d0 < 0 && p.d0 < cand[i].d0) { p = cand[i]; flag = 1; }
      else if (p.d0 > 0 && cand[i].d0 > 0 && p.d0 > cand[i].d0) { p = cand[i]; flag = 2; }
    }
    next_alpha = p.alpha;
    if (flag > 0) next_alpha = p.alpha - p.d0 / p.d1;
    return flag;
  }
  // CS:2426-2597 func_linesearch_batch
  T linesearch(Env<T>& e) {
    int nD = m.nD;
    T snorm = 0;
    for (int d = 0; d < nD; d++) snorm = snorm + e.search[d] * e.search[d];
    snorm = std::sqrt(snorm);
    T scale = m.meaninertia * T(std::max(1, nD));
    T gtol = (T)cfg.tolerance * (T)cfg.ls_tolerance * snorm * scale;
    e.ls_it = 0; e.ls_result = 0;
    T res_alpha = 0;
    bool done = false;
    if (snorm < EPS) { e.ls_result = 1; return 0; }
    LsPoint p0 = ls_p0(e);
    LsPoint p1 = ls_point(e, p0.alpha - p0.d0 / p0.d1);
    if (p0.cost < p1.cost) p1 = p0;
    if (std::abs(p1.d0) < gtol) {
      e.ls_result = std::abs(p1.alpha) < EPS ? 2 : 0;
      return p1.alpha;
    }
    int direction = (p1.d0 < 0) * 2 - 1;
    bool p2update = false;
    LsPoint p2 = p1;
    while (p1.d0 * direction <= -gtol && e.ls_it < cfg.ls_iterations) {
      p2 = p1; p2update = true;
      p1 = ls_point(e, p1.alpha - p1.d0 / p1.d1);
      if (std::abs(p1.d0) < gtol) { res_alpha = p1.alpha; done = true; break; }
    }
    if (done) return res_alpha;
    if (e.ls_it >= cfg.ls_iterations) { e.ls_result = 3; return p1.alpha; }
    if (!p2update) { e.ls_result = 6; return p1.alpha; }
    T a0 = p1.alpha - p1.d0 / p1.d1, a1 = p1.alpha, a2 = (p1.alpha + p2.alpha) * T(0.5);
    while (e.ls_it < cfg.ls_iterations) {
      LsPoint cand[3];
      T al[3] = {a0, a1, a2};
      for (int i = 0; i < 3; i++) {
        T t0, t1, t2;
        ls_accumulate(e, al[i], t0, t1, t2);
        cand[i].alpha = al[i];
        cand[i].cost = al[i] * al[i] * t2 + al[i] * t1 + t0;
        cand[i].d0 = 2 * al[i] * t2 + t1;
        cand[i].d1 = 2 * t2;
        if (cand[i].d1 <= 0) cand[i].d1 = EPS;
      }
      e.ls_it += 3;
      T best_alpha = 0, best_cost = 0;
      bool best_found = false;
      for (int i = 0; i < 3; i++)
        if (std::abs(cand[i].d0) < gtol && (!best_found || cand[i].cost < best_cost)) { best_alpha = cand[i].alpha; best_cost = cand[i].cost; best_found = true; }
      if (best_found) return best_alpha;
      T p1_next, p2_next;
      int b1 = update_bracket(p1, cand, p1_next);
      int b2 = update_bracket(p2, cand, p2_next);
      if (b1 == 0 && b2 == 0) {
        e.ls_result = cand[2].cost < p0.cost ? 0 : 7;
        return a2;
      }
      a0 = p1_next; a1 = p2_next; a2 = (p1.alpha + p2.alpha) * T(0.5);
    }
    if (p1.cost <= p2.cost && p1.cost < p0.cost) { e.ls_result = 4; return p1.alpha; }
    if (p2.cost <= p1.cost && p2.cost < p0.cost) { e.ls_result = 4; return p2.alpha; }
    e.ls_result = 5;
    return 0;
  }

  // CS:3062-3139 func_solve_iter
  void solve_iter(Env<T>& e) {
    int nD = m.nD;
    T alpha = linesearch(e);
    if (std::abs(alpha) < EPS) { e.improved = 0; return; }
    for (int d = 0; d < nD; d++) { e.qacc[d] = e.qacc[d] + e.search[d] * alpha; e.Ma[d] = e.Ma[d] + e.mv[d] * alpha; }
    for (int c = 0; c < e.n_con; c++) e.Jaref[c] = e.Jaref[c] + e.jv[c] * alpha;
    if (cfg.solver_type == SOLVER_CG) { e.cg_prev_grad = e.grad; e.cg_prev_Mgrad = e.Mgrad; }
    update_constraint(e);
    if (cfg.solver_type == SOLVER_NEWTON) {
      if (hessian_incremental(e)) hessian_and_cholesky_direct(e);
    }
    update_gradient(e);
    // CS:2841-2884 func_terminate_or_update_descent_batch
    T tol_scaled = (m.meaninertia * T(std::max(1, nD))) * (T)cfg.tolerance;
    T improvement = e.prev_cost - e.cost;
    T gn = 0;
    for (int d = 0; d < nD; d++) gn = gn + e.grad[d] * e.grad[d];
    gn = std::sqrt(gn);
    e.improved = gn > tol_scaled && improvement > tol_scaled;
    if (e.improved) {
      if (cfg.solver_type == SOLVER_NEWTON) {
        for (int d = 0; d < nD; d++) e.search[d] = -e.Mgrad[d];
      } else {
        T beta = 0, pg = 0;
        for (int d = 0; d < nD; d++) { beta = beta + e.grad[d] * (e.Mgrad[d] - e.cg_prev_Mgrad[d]); pg = pg + e.cg_prev_Mgrad[d] * e.cg_prev_grad[d]; }
        beta = std::max(beta / std::max(EPS, pg), T(0));
        for (int d = 0; d < nD; d++) e.search[d] = -e.Mgrad[d] + beta * e.search[d];
      }
    }
  }

  // CS:210-245 resolve()
  // constraint/noslip.py: post-solver that removes residual slip -- projected Gauss-Seidel on the friction-loss rows and on
  // the two opposing pyramid edges of every tangent direction, with the regulariser R dropped (AR = J M^-1 J^T)
  T noslip_cost_change(const T* Ac, T* force, const T* old_force, const T* res, int dim) {  // noslip.py:358-385
    T change = 0;
    if (dim == 1) {
      T delta = force[0] - old_force[0];
      change = T(0.5) * Ac[0] * delta * delta + delta * res[0];
    } else {
      T delta[2];
      for (int i = 0; i < dim; i++) delta[i] = force[i] - old_force[i];
      for (int i = 0; i < dim; i++) {
        for (int j = 0; j < dim; j++) change += T(0.5) * Ac[i * dim + j] * delta[i] * delta[j];
        change += delta[i] * res[i];
      }
    }
    if (change > EPS) {
      for (int i = 0; i < dim; i++) force[i] = old_force[i];
      change = 0;
    }
    return change;
  }
  void noslip(Env<T>& e) {
    const int nD = m.nD, nefc = e.n_con;
    if (nefc == 0) return;
    // noslip.py:9-58 kernel_build_efc_AR_b
    std::vector<T> AR((size_t)nefc * nefc, 0), b(nefc, 0), tmp(nD), rowv(nD);
    for (int r = 0; r < nefc; r++) {
      for (int d = 0; d < nD; d++) rowv[d] = jrow(e, r)[d];
      tmp = rowv;
      solve_mass(e, rowv, tmp);
      for (int c = 0; c < nefc; c++) {
        T sum = 0;
        for (int d = 0; d < nD; d++) sum += jrow(e, c)[d] * tmp[d];
        AR[(size_t)r * nefc + c] = sum;
      }
    }
    for (int c = 0; c < nefc; c++) {
      T v = -e.aref[c];
      for (int d = 0; d < nD; d++) v += jrow(e, c)[d] * e.acc_smooth[d];
      b[c] = v;
    }
    auto residual = [&](int i_efc, int dim, T* res) {  // noslip.py:343-355
      for (int j = 0; j < dim; j++) {
        res[j] = b[i_efc + j];
        for (int k = 0; k < nefc; k++) res[j] += AR[(size_t)(i_efc + j) * nefc + k] * e.efc_force[k];
      }
    };
    // noslip.py:166-283 kernel_noslip
    const int ne = e.n_eq, nf = e.n_floss, const_start = ne + nf;
    const T scale = T(1) / (m.meaninertia * std::max(T(1), T(nD)));
    for (int it = 0; it < cfg.noslip_iterations; it++) {
      T improvement = 0;
      if (it == 0)
        for (int c = 0; c < nefc; c++) improvement += T(0.5) * e.efc_force[c] * e.efc_force[c] * e.diag[c];
      T res[2], old_force[2], Ac[4], bc[2];
      for (int c = ne; c < ne + nf; c++) {
        residual(c, 1, res);
        old_force[0] = e.efc_force[c];
        T Acc = AR[(size_t)c * nefc + c];
        e.efc_force[c] -= res[0] / Acc;
        if (e.efc_force[c] < -e.efc_frictionloss[c]) e.efc_force[c] = -e.efc_frictionloss[c];
        else if (e.efc_force[c] > e.efc_frictionloss[c]) e.efc_force[c] = e.efc_frictionloss[c];
        T delta = e.efc_force[c] - old_force[0];
        improvement -= T(0.5) * delta * delta / Acc + delta * res[0];
      }
      for (int ic = 0; ic < e.n_contacts; ic++) {
        int base = const_start + ic * 4;
        for (int j2 = 0; j2 < 2; j2++) {
          int j = base + j2 * 2;
          residual(j, 2, res);
          for (int i2 = 0; i2 < 2; i2++) old_force[i2] = e.efc_force[j + i2];
          for (int a = 0; a < 2; a++)
            for (int k = 0; k < 2; k++) Ac[a * 2 + k] = AR[(size_t)(j + a) * nefc + (j + k)];
          for (int a = 0; a < 2; a++) {
            bc[a] = res[a];
            for (int k = 0; k < 2; k++) bc[a] -= Ac[a * 2 + k] * old_force[k];
          }
          T mid = T(0.5) * (e.efc_force[j] + e.efc_force[j + 1]);
          T K1 = Ac[0] + Ac[3] - Ac[1] - Ac[2];
          T K0 = mid * (Ac[0] - Ac[3]) + bc[0] - bc[1];
          if (K1 < EPS) {
            e.efc_force[j] = e.efc_force[j + 1] = mid;
          } else {
            T y = -K0 / K1;
            if (y < -mid) { e.efc_force[j] = 0; e.efc_force[j + 1] = 2 * mid; }
            else if (y > mid) { e.efc_force[j] = 2 * mid; e.efc_force[j + 1] = 0; }
            else { e.efc_force[j] = mid + y; e.efc_force[j + 1] = mid - y; }
          }
          improvement -= noslip_cost_change(Ac, &e.efc_force[j], old_force, res, 2);
        }
      }
      improvement *= scale;
      if (improvement < (T)cfg.noslip_tolerance) break;
    }
    // noslip.py:286-323 kernel_dual_finish
    for (int d = 0; d < nD; d++) {
      T q = 0;
      for (int c = 0; c < nefc; c++) q += jrow(e, c)[d] * e.efc_force[c];
      e.qf_constraint[d] = q;
    }
    std::vector<T> qa(nD, 0);
    solve_mass(e, e.qf_constraint, qa);
    for (int d = 0; d < nD; d++) {
      e.qacc[d] = qa[d] + e.acc_smooth[d];
      e.acc[d] = e.qacc[d];
      e.force[d] = e.qf_smooth[d] + e.qf_constraint[d];
    }
  }

  void resolve(Env<T>& e) {
    int nD = m.nD;
    solve_init(e);
    e.n_iter = 0;
    if (e.n_con > 0) {  // CS:3163-3190 func_solve_body_monolith
      for (int it = 0; it < cfg.iterations; it++) {
        solve_iter(e);
        e.n_iter++;
        if (!e.improved) break;
      }
    } else {
      e.improved = 0;
    }
    for (int d = 0; d < nD; d++) {  // CS:3240-3261 func_update_qacc
      e.acc[d] = e.qacc[d];
      e.force[d] = e.qf_smooth[d] + e.qf_constraint[d];
      e.qacc_ws[d] = e.qacc[d];
      if (std::isnan(e.qacc[d])) e.errno_ |= ERR_NAN_FORCE;
    }
    e.is_warmstart = 1;
    if (cfg.noslip_iterations > 0) noslip(e);  // CS:237-238
    // CS:3198-3237 func_update_contact_force
    for (int l = 0; l < m.nL; l++) e.contact_force[l] = V();
    int cs = e.n_eq + e.n_floss;
    for (int ic = 0; ic < e.n_contacts; ic++) {
      Contact<T>& ct = e.contacts[ic];
      V force, d1, d2;
      orthogonals(ct.normal, d1, d2);
      for (int i = 0; i < 4; i++) {
        V d = (i < 2 ? d1 : d2) * T(2 * (i % 2) - 1);
        V n = d * ct.friction - ct.normal;
        force = force + n * e.efc_force[ic * 4 + i + cs];
      }
      ct.force = force;
      e.contact_force[ct.link_a] = e.contact_force[ct.link_a] - force;
      e.contact_force[ct.link_b] = e.contact_force[ct.link_b] + force;
    }
  }

  // ------------------------------------------------------------------ integration
  // FD:1566-1707 func_integrate + diff.py:25-54 func_copy_next_to_curr
  void integrate(Env<T>& e) {
    for (int d = 0; d < m.nD; d++) e.vel_next[d] = e.vel[d] + e.acc[d] * dt;
    for (int l = 0; l < m.nL; l++) {
      int ds = m.links_dof_start[l], qs = m.links_q_start[l], qe = m.links_q_end[l];
      if (m.links_dof_end[l] - ds <= 0) continue;
      int jt = m.joints_type[m.links_joint_start[l]];
      if (jt == JOINT_FREE) {
        for (int k = 0; k < 3; k++) e.qpos_next[qs + k] = e.qpos[qs + k] + e.vel_next[ds + k] * dt;
      }
      if (jt == JOINT_SPHERICAL || jt == JOINT_FREE) {
        int ro = jt == JOINT_FREE ? 3 : 0;
        Q rot0(e.qpos[qs + ro], e.qpos[qs + ro + 1], e.qpos[qs + ro + 2], e.qpos[qs + ro + 3]);
        V ang = V(e.vel_next[ds + ro], e.vel_next[ds + ro + 1], e.vel_next[ds + ro + 2]) * dt;
        Q qrot = rotvec_to_quat(ang, EPS);
        Q rot = transform_quat_by_quat(qrot, rot0);
        e.qpos_next[qs + ro] = rot.w; e.qpos_next[qs + ro + 1] = rot.x; e.qpos_next[qs + ro + 2] = rot.y; e.qpos_next[qs + ro + 3] = rot.z;
      } else {
        for (int k = 0; k < qe - qs; k++) e.qpos_next[qs + k] = e.qpos[qs + k] + e.vel_next[ds + k] * dt;
      }
    }
    bool valid = true;
    for (int d = 0; d < m.nD; d++) valid &= !std::isnan(e.vel_next[d]);
    for (int q = 0; q < m.nQ; q++) valid &= !std::isnan(e.qpos_next[q]);
    if (valid) { e.vel = e.vel_next; e.qpos = e.qpos_next; }
    else e.errno_ |= ERR_NAN_ACC;
  }

  // FD:1787-1859 func_implicit_damping (Euler / implicitfast integrators)
  void implicit_damping(Env<T>& e) {
    std::vector<uint8_t> mask(m.nE, 0);  // FD:1806-1832: only entities with damping / velocity-dependent actuation
    for (int i_e = 0; i_e < m.nE; i_e++)
      for (int d = m.entities_dof_start[i_e]; d < m.entities_dof_end[i_e]; d++) {
        if (m.dofs_damping[d] > EPS) mask[i_e] = 1;
        if (cfg.integrator != INTEGRATOR_EULER && e.ctrl_mode[d] <= CTRL_VELOCITY && std::abs(m.dofs_act_bias[3 * d + 2]) > EPS) mask[i_e] = 1;
      }
    factor_mass(e, true, &mask);
    solve_mass(e, e.force, e.acc, &mask);
  }

  // RS:912-980 substep = kernel_step_1 + _func_constraint_force + kernel_step_2 (+ RS:2603 clear_external_force)
  void step_env(Env<T>& e) {
    forward_dynamics(e);
    add_equality_constraints(e);
    if (cfg.enable_collision) detection(e);
    add_frictionloss(e);
    if (cfg.enable_collision) add_collision(e);
    if (cfg.enable_joint_limit) add_joint_limits(e);
    resolve(e);
    update_acc(e, true);
    if (cfg.integrator != INTEGRATOR_APPROX_IMPLICITFAST) implicit_damping(e);
    integrate(e);
    update_cartesian_space(e);
    forward_velocity(e);
    for (int l = 0; l < m.nL; l++) { e.cfrc_applied_vel[l] = V(); e.cfrc_applied_ang[l] = V(); }
  }

  int n_threads = 1;
  void step(int n) {
    int B = (int)envs.size();
    int nt = std::max(1, std::min(n_threads, B));
    if (nt == 1) {
      for (int s = 0; s < n; s++)
        for (int b = 0; b < B; b++) step_env(envs[b]);
      return;
    }
    // environments are independent: each host thread advances a contiguous block of envs by n steps
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++)
      pool.emplace_back([this, t, nt, B, n]() {
        int b0 = (int)((long long)B * t / nt), b1 = (int)((long long)B * (t + 1) / nt);
        for (int s = 0; s < n; s++)
          for (int b = b0; b < b1; b++) step_env(envs[b]);
      });
    for (auto& th : pool) th.join();
  }
};

}  // namespace orc
