// gsr_types.h -- device-side model (constant memory) and per-env state layout of libgsrigid.
//
// Data layout in HBM (see DESIGN.md "Data layout"): every per-environment quantity lives in one of two big
// buffers (float32 / int32) as rows of B consecutive environments: element (row r, env b) = buf[r * B + b].
// A warp of 32 consecutive environments therefore touches 128 contiguous bytes per load/store, which is the
// reference's `[item, env]` field layout (genesis/utils/array_class.py:1400-1435) with vec3/quat components
// split into scalar rows.  The model (identical for all envs) sits in __constant__ memory so that the uniform
// per-link / per-dof / per-geom table reads of a thread-per-env kernel are single broadcast transactions.
#pragma once
#include <stdint.h>

#define GSR_MAX_LINKS 40
#define GSR_MAX_JOINTS 40
#define GSR_MAX_DOFS 40
#define GSR_MAX_QS 48
#define GSR_MAX_GEOMS 80
#define GSR_MAX_ENTS 16
#define GSR_MAX_EQ 8

enum { GEOM_PLANE = 0, GEOM_SPHERE = 1, GEOM_ELLIPSOID = 2, GEOM_CYLINDER = 3, GEOM_CAPSULE = 4, GEOM_BOX = 5, GEOM_MESH = 6 };
enum { JOINT_FIXED = 0, JOINT_REVOLUTE = 1, JOINT_PRISMATIC = 2, JOINT_SPHERICAL = 3, JOINT_FREE = 4 };
enum { EQ_CONNECT = 0, EQ_WELD = 1, EQ_JOINT = 2 };
enum { CTRL_POSITION = 0, CTRL_VELOCITY = 1, CTRL_FORCE = 2 };
enum { INTEGRATOR_EULER = 0, INTEGRATOR_IMPLICITFAST = 1, INTEGRATOR_APPROX_IMPLICITFAST = 2 };
enum { SOLVER_CG = 0, SOLVER_NEWTON = 1 };
enum { ERR_BROAD_OVERFLOW = 1, ERR_CONTACT_OVERFLOW = 2, ERR_NAN_FORCE = 8, ERR_NAN_ACC = 16 };

struct DevModel {
  int nL, nJ, nD, nQ, nG, nE, nEq, nP, nV;
  int max_contact_pairs, max_broad_pairs, n_contacts_per_pair, len_constraints;
  int iterations, ls_iterations, integrator, solver_type;
  int enable_collision, enable_joint_limit, enable_multi_contact, sort_contacts, has_cc, n_cache_words;
  float meaninertia, dt, eps, tolerance, ls_tolerance, mc_perturbation, mc_tolerance;
  float gravity[3];
  float ccd_eps, ccd_tolerance;
  int ccd_iterations;
  // links
  int l_parent[GSR_MAX_LINKS], l_root[GSR_MAX_LINKS], l_entity[GSR_MAX_LINKS], l_is_fixed[GSR_MAX_LINKS];
  int l_joint_start[GSR_MAX_LINKS], l_joint_end[GSR_MAX_LINKS], l_dof_start[GSR_MAX_LINKS], l_dof_end[GSR_MAX_LINKS];
  int l_q_start[GSR_MAX_LINKS], l_q_end[GSR_MAX_LINKS];
  float l_pos[GSR_MAX_LINKS][3], l_quat[GSR_MAX_LINKS][4], l_ipos[GSR_MAX_LINKS][3], l_iquat[GSR_MAX_LINKS][4];
  float l_inertia[GSR_MAX_LINKS][9], l_mass[GSR_MAX_LINKS], l_invweight[GSR_MAX_LINKS];
  // joints
  int j_type[GSR_MAX_JOINTS], j_q_start[GSR_MAX_JOINTS], j_dof_start[GSR_MAX_JOINTS], j_n_dofs[GSR_MAX_JOINTS];
  float j_pos[GSR_MAX_JOINTS][3], j_sol[GSR_MAX_JOINTS][7];
  // dofs
  int d_link[GSR_MAX_DOFS], d_entity[GSR_MAX_DOFS];
  float d_motion_ang[GSR_MAX_DOFS][3], d_motion_vel[GSR_MAX_DOFS][3], d_limit[GSR_MAX_DOFS][2];
  float d_stiffness[GSR_MAX_DOFS], d_damping[GSR_MAX_DOFS], d_armature[GSR_MAX_DOFS], d_frictionloss[GSR_MAX_DOFS];
  float d_kp[GSR_MAX_DOFS], d_act_bias[GSR_MAX_DOFS][3], d_force_range[GSR_MAX_DOFS][2], d_invweight[GSR_MAX_DOFS];
  float qpos0[GSR_MAX_QS];
  uint8_t mass_mask[GSR_MAX_DOFS][GSR_MAX_DOFS];
  // entities
  int e_link_start[GSR_MAX_ENTS], e_link_end[GSR_MAX_ENTS], e_dof_start[GSR_MAX_ENTS], e_dof_end[GSR_MAX_ENTS];
  int e_geom_start[GSR_MAX_ENTS], e_geom_end[GSR_MAX_ENTS];
  float e_gravcomp[GSR_MAX_ENTS];
  // geoms
  int g_link[GSR_MAX_GEOMS], g_type[GSR_MAX_GEOMS], g_is_fixed[GSR_MAX_GEOMS], g_vert_start[GSR_MAX_GEOMS];
  int g_vert_end[GSR_MAX_GEOMS], g_support_start[GSR_MAX_GEOMS];
  float g_pos[GSR_MAX_GEOMS][3], g_quat[GSR_MAX_GEOMS][4], g_data[GSR_MAX_GEOMS][3], g_friction[GSR_MAX_GEOMS];
  float g_sol[GSR_MAX_GEOMS][7], g_center[GSR_MAX_GEOMS][3], g_aabb_lo[GSR_MAX_GEOMS][3], g_aabb_hi[GSR_MAX_GEOMS][3];
  // equalities
  int eq_type[GSR_MAX_EQ], eq_obj1[GSR_MAX_EQ], eq_obj2[GSR_MAX_EQ];
  float eq_data[GSR_MAX_EQ][11], eq_sol[GSR_MAX_EQ][7];
  // large read-only tables in global memory
  const float* verts;      // [nV][3]
  const float* support_v;  // [n_support_cells][3]
  const int* pairs;        // [nP][2]
  const int* pair_idx;     // [nG][nG]
};

// ---- per-env state fields: X(name, is_int, rows) with rows expressed in model counts ------------------------
// nC = len_constraints, nCP = max_contact_pairs, nW = n_cache_words
#define GSR_FIELDS(X)                                                                                             \
  /* generalized coordinates (reference: rigid_global_info.qpos, dofs_state.*) */                                 \
  X(qpos, 0, nQ) X(vel, 0, nD) X(acc, 0, nD) X(acc_smooth, 0, nD) X(force, 0, nD) X(qf_smooth, 0, nD)             \
  X(qf_constraint, 0, nD) X(qf_applied, 0, nD) X(qf_passive, 0, nD) X(qf_bias, 0, nD) X(dofs_pos, 0, nD)          \
  X(ctrl_pos, 0, nD) X(ctrl_vel, 0, nD) X(ctrl_force, 0, nD) X(ctrl_mode, 1, nD)                                  \
  X(cdof_ang, 0, 3 * nD) X(cdof_vel, 0, 3 * nD) X(cdofd_ang, 0, 3 * nD) X(cdofd_vel, 0, 3 * nD)                   \
  /* links_state */                                                                                               \
  X(links_pos, 0, 3 * nL) X(links_quat, 0, 4 * nL) X(root_COM, 0, 3 * nL) X(cinr_I, 0, 6 * nL)                    \
  X(cinr_pos, 0, 3 * nL) X(cd_vel, 0, 3 * nL) X(cd_ang, 0, 3 * nL) X(cacc_lin, 0, 3 * nL) X(cacc_ang, 0, 3 * nL)  \
  X(cfrc_applied_vel, 0, 3 * nL) X(cfrc_applied_ang, 0, 3 * nL) X(links_contact_force, 0, 3 * nL)                 \
  X(joints_xanchor, 0, 3 * nJ) X(joints_xaxis, 0, 3 * nJ) X(scratch, 0, 16 * nL)                                  \
  /* geoms_state */                                                                                               \
  X(geoms_pos, 0, 3 * nG) X(geoms_quat, 0, 4 * nG) X(aabb_min, 0, 3 * nG) X(aabb_max, 0, 3 * nG)                  \
  X(friction_ratio, 0, nG)                                                                                        \
  /* mass matrix and its L^T D L factor (rigid_global_info.mass_mat, mass_mat_L, mass_mat_D_inv) */               \
  X(mass_mat, 0, nD * nD) X(mass_L, 0, nD * nD) X(mass_Dinv, 0, nD)                                               \
  /* collider_state */                                                                                            \
  X(n_contacts, 1, 1) X(contact_geom_a, 1, nCP) X(contact_geom_b, 1, nCP) X(contact_link_a, 1, nCP)               \
  X(contact_link_b, 1, nCP) X(contact_normal, 0, 3 * nCP) X(contact_pos, 0, 3 * nCP)                              \
  X(contact_penetration, 0, nCP) X(contact_friction, 0, nCP) X(contact_force, 0, 3 * nCP)                         \
  X(cache_normal, 0, 3 * nP) X(cache_valid, 1, nW) X(sort_key, 0, nCP) X(sort_idx, 1, nCP)                        \
  /* constraint_state */                                                                                          \
  X(n_constraints, 1, 1) X(n_constraints_equality, 1, 1) X(n_constraints_frictionloss, 1, 1)                      \
  X(jac, 0, nC * nD) X(efc_diag, 0, nC) X(efc_aref, 0, nC) X(efc_D, 0, nC) X(efc_force, 0, nC)                    \
  X(efc_frictionloss, 0, nC) X(Jaref, 0, nC) X(jv, 0, nC) X(efc_active, 1, nC)                                    \
  X(qacc, 0, nD) X(qacc_ws, 0, nD) X(Ma, 0, nD) X(grad, 0, nD) X(Mgrad, 0, nD) X(search, 0, nD) X(mv, 0, nD)      \
  X(nt_H, 0, nD * nD) X(nt_vec, 0, nD) X(cg_prev_grad, 0, nD) X(cg_prev_Mgrad, 0, nD)                             \
  X(is_warmstart, 1, 1) X(solver_iters, 1, 1) X(errno_, 1, 1)

struct Layout {
#define X(name, is_int, rows) int name;
  GSR_FIELDS(X)
#undef X
};
