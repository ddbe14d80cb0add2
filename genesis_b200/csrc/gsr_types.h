// gsr_types.h -- device-side model, per-env record layout (HBM) and per-env workspace layout (shared memory) of libgsrigid.
//
// Execution model (DESIGN.md "Kernels"): one *lane group* of G lanes (G = 32: a full warp; smaller power-of-two groups are
// supported by the same code) owns one environment for a whole step.  The group stages the environment's record from HBM
// into shared memory with one bulk copy, runs forward dynamics, collision detection, the constraint solver and the
// integrator + forward kinematics entirely on chip, and writes the record back.  Lanes of a group split links / dofs /
// geoms / collision pairs / constraint rows / matrix entries between them and communicate through shared memory and warp
// shuffles.
//
// Data layout in HBM: env-major records.  The caller-owned state buffer of 32-bit words is split into segments
//   HOT     [B][hot_stride]    : IN part (persistent state + kinematic cache, bulk-loaded every step) followed by the
//                                OUT part (accessor-visible results, written every step)
//   CONTACT [B][max_contacts][16] : contact records (AoS, 64 B each), only the first n_contacts are touched
//   CACHE   [B][cache_stride]  : per-pair warm-start normals of the narrow phase (touched for colliding convex pairs only)
//   SPILL   [B][spill_stride]  : constraint rows of environments whose row count exceeds the shared-memory capacity
// so that everything one group touches is contiguous (full 128-byte lines, bulk-copy friendly).  The reference keeps
// `[item, env]` arrays (genesis/utils/array_class.py:1400-1435) because it maps one thread to one env; the public API is
// `(B, item)` in both (genesis/utils/misc.py:415-536), which here is a plain strided view of the record.
#pragma once
#include <stdint.h>

#define GSR_MAX_LINKS 40
#define GSR_MAX_JOINTS 40
#define GSR_MAX_DOFS 40
#define GSR_MAX_QS 48
#define GSR_MAX_GEOMS 80
#define GSR_MAX_ENTS 16
#define GSR_MAX_EQ 8
#define GSR_CONTACT_WORDS 16
#define GSR_MAX_DYN_EQ 8      /* RigidOptions.max_dynamic_constraints (options/solvers.py) */
#define GSR_DYN_EQ_WORDS 20

enum { GEOM_PLANE = 0, GEOM_SPHERE = 1, GEOM_ELLIPSOID = 2, GEOM_CYLINDER = 3, GEOM_CAPSULE = 4, GEOM_BOX = 5, GEOM_MESH = 6, GEOM_TERRAIN = 7 };
enum { JOINT_FIXED = 0, JOINT_REVOLUTE = 1, JOINT_PRISMATIC = 2, JOINT_SPHERICAL = 3, JOINT_FREE = 4 };
enum { EQ_CONNECT = 0, EQ_WELD = 1, EQ_JOINT = 2 };
enum { CTRL_POSITION = 0, CTRL_VELOCITY = 1, CTRL_FORCE = 2 };
enum { INTEGRATOR_EULER = 0, INTEGRATOR_IMPLICITFAST = 1, INTEGRATOR_APPROX_IMPLICITFAST = 2 };
enum { SOLVER_CG = 0, SOLVER_NEWTON = 1 };
enum { ERR_BROAD_OVERFLOW = 1, ERR_CONTACT_OVERFLOW = 2, ERR_NAN_FORCE = 8, ERR_NAN_ACC = 16 };
// contact record word offsets (CONTACT segment, AoS)
enum { CT_GA = 0, CT_GB = 1, CT_LA = 2, CT_LB = 3, CT_PEN = 4, CT_MU = 5, CT_POS = 6, CT_NRM = 9, CT_FRC = 12, CT_KEY = 15 };

struct DevModel {
  int nL, nJ, nD, nQ, nG, nE, nEq, nP, nV;
  int max_contact_pairs, max_broad_pairs, n_contacts_per_pair, len_constraints;
  int iterations, ls_iterations, integrator, solver_type;
  int noslip_iterations;  // constraint/noslip.py post-solver (0 = off)
  int has_mpr;            // any valid pair that needs MPR / EPA (neither a plane pair nor an analytic capsule pair): selects the kernel variant
  int has_box_box;        // a box-box pair under RigidOptions.box_box_detection or a terrain pair: the "full" variant (func_box_box_contact,
                          // func_contact_mpr_terrain compiled in)
  int box_box_detection;  // options/solvers.py:511
  int use_gjk;            // RigidOptions.use_gjk_collision: convex pairs through GJK + EPA only
  // height-field terrain (collider.py:541-561): scaled heights (rows x cols, device), cell size, [x_max, y_max, z_max, x_min, y_min, z_min - 1]
  const float* terrain_hf;
  int terrain_rows, terrain_cols;
  float terrain_hs, terrain_maxmin[6];
  int phase_barriers;     // bit k: CTA barrier k of k_step (0 after the record load, 1 dynamics, 2 collision, 3 constraints, 4 integration)
  int gjk_fallback;       // exact-penetration fallback after MPR (narrowphase.py:726-733); 0 keeps MPR's answer
  float noslip_tolerance;
  int enable_collision, enable_joint_limit, enable_multi_contact, sort_contacts, has_cc, n_cache_words, diagnostics;
  int batch_dofs;          // RigidOptions.batch_dofs_info: per-environment dof parameters live in the record (bd_* fields)
  int disable_constraint;  // RigidOptions.disable_constraint: no rows at all (rigid_solver.py:1026-1042)
  float meaninertia, dt, eps, tolerance, ls_tolerance, mc_perturbation, mc_tolerance;
  float gravity[3];
  float ccd_eps, ccd_tolerance;
  int ccd_iterations;
  int max_depth;  // longest root-to-leaf chain (number of links)
  // links
  int l_parent[GSR_MAX_LINKS], l_root[GSR_MAX_LINKS], l_entity[GSR_MAX_LINKS], l_is_fixed[GSR_MAX_LINKS];
  int l_joint_start[GSR_MAX_LINKS], l_joint_end[GSR_MAX_LINKS], l_dof_start[GSR_MAX_LINKS], l_dof_end[GSR_MAX_LINKS];
  int l_q_start[GSR_MAX_LINKS], l_q_end[GSR_MAX_LINKS], l_depth[GSR_MAX_LINKS];
  unsigned long long l_dofmask[GSR_MAX_LINKS];   // dofs of the link and of all its ancestors
  unsigned long long l_ancmask[GSR_MAX_LINKS];   // the link and all its ancestors
  unsigned long long l_descmask[GSR_MAX_LINKS];  // the link and all its descendants
  float l_pos[GSR_MAX_LINKS][3], l_quat[GSR_MAX_LINKS][4], l_ipos[GSR_MAX_LINKS][3], l_iquat[GSR_MAX_LINKS][4];
  float l_inertia[GSR_MAX_LINKS][9], l_mass[GSR_MAX_LINKS], l_invweight[GSR_MAX_LINKS], l_root_mass[GSR_MAX_LINKS];
  int n_roots, l_rootslot[GSR_MAX_LINKS];  // compact index of the link's tree root
  float l_submass[GSR_MAX_LINKS];          // mass of the subtree rooted at the link
  float l_invweight_rot[GSR_MAX_LINKS];    // rotational component of links_info.invweight (weld equalities)
  // joints
  int j_type[GSR_MAX_JOINTS], j_q_start[GSR_MAX_JOINTS], j_dof_start[GSR_MAX_JOINTS], j_n_dofs[GSR_MAX_JOINTS];
  float j_pos[GSR_MAX_JOINTS][3], j_sol[GSR_MAX_JOINTS][7];
  // dofs
  int d_link[GSR_MAX_DOFS], d_entity[GSR_MAX_DOFS], d_joint[GSR_MAX_DOFS];
  float d_motion_ang[GSR_MAX_DOFS][3], d_motion_vel[GSR_MAX_DOFS][3], d_limit[GSR_MAX_DOFS][2];
  float d_stiffness[GSR_MAX_DOFS], d_damping[GSR_MAX_DOFS], d_armature[GSR_MAX_DOFS], d_frictionloss[GSR_MAX_DOFS];
  float d_kp[GSR_MAX_DOFS], d_act_bias[GSR_MAX_DOFS][3], d_force_range[GSR_MAX_DOFS][2], d_invweight[GSR_MAX_DOFS];
  float qpos0[GSR_MAX_QS];
  unsigned short tri[GSR_MAX_DOFS * (GSR_MAX_DOFS + 1) / 2];  // lower-triangle task t -> (row << 8) | col
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
  const int* support_vid;  // [n_support_cells] vertex index (inside its geom) of every support cell
  const int* pairs;        // [nP][2]
  const int* pair_idx;     // [nG][nG]
  // L^T D L elimination tasks of pivot i: fac_tasks[fac_start[i] .. fac_start[i + 1]) = (j << 8) | k over the ancestor dofs
  // k <= j < i of dof i (branches of the tree never fill in, so only ancestor pairs are touched)
  const unsigned short* fac_tasks;
  int fac_start[GSR_MAX_DOFS + 1];
};

// ---- HOT segment fields: X(name, is_int, items, comps).  Everything up to `kin_end` is the IN part (same offsets in the
// shared-memory workspace); the rest is OUT.
// nL/nD/... are the model counts; "items x comps" words, item-major.
// The record starts with the PRE block: inputs the step reads once, straight from the record (they are never written by the
// kernel, so mirroring them in shared memory and storing them back would only cost workspace and HBM traffic).
#define GSR_HOT_PRE_FIELDS(X)                                                                                      \
  X(ctrl_vel, 0, nD, 1) X(ctrl_force, 0, nD, 1) X(ctrl_mode, 1, nD, 1) X(fk_dirty, 1, 1, 1)
#define GSR_HOT_STATE_FIELDS(X)                                                                                    \
  /* persistent state (rigid_global_info.qpos, dofs_state.vel / ctrl_pos, constraint_state.qacc_ws) and dofs_state.pos */ \
  X(qpos, 0, nQ, 1) X(vel, 0, nD, 1) X(ctrl_pos, 0, nD, 1) X(qacc_ws, 0, nD, 1)                                    \
  X(is_warmstart, 1, 1, 1) X(errno_, 1, 1, 1) X(n_contacts, 1, 1, 1) X(n_broad, 1, 1, 1)                           \
  X(has_ext_force, 1, 1, 1) X(has_shift, 1, 1, 1) X(n_dyn_eq, 1, 1, 1) X(dofs_pos, 0, nD, 1)
#define GSR_HOT_CACHE_FIELDS(X)                                                                                    \
  /* kinematic cache: a pure function of (qpos, vel, poses of the fixed roots).  Written by the FK at the end of a step and   \
     consumed by the next one (eager), or recomputed at the start of every step and never stored (lazy, gsr_set_kinematics_policy) */ \
  X(root_COM, 0, nL, 3) X(cdof_ang, 0, nD, 3) X(cdof_vel, 0, nD, 3) X(cdofd_ang, 0, nD, 3)                         \
  X(cdofd_vel, 0, nD, 3) X(links_pos, 0, nL, 3) X(links_quat, 0, nL, 4) X(cinr_I, 0, nL, 6) X(cinr_pos, 0, nL, 3)  \
  X(cd_vel, 0, nL, 3) X(cd_ang, 0, nL, 3)
#define GSR_HOT_IN_FIELDS(X) GSR_HOT_STATE_FIELDS(X) GSR_HOT_CACHE_FIELDS(X)
#define GSR_HOT_OUT_FIELDS(X)                                                                                      \
  /* rarely used inputs, read from the record on demand: per-geom friction scale (set_friction_ratio) and the external  \
     wrenches of apply_links_external_force (only read and cleared when has_ext_force is set) */                       \
  X(friction_ratio, 0, nG, 1) X(cfrc_applied_vel, 0, nL, 3) X(cfrc_applied_ang, 0, nL, 3)                          \
  /* domain randomisation (links_state.mass_shift / i_pos_shift), read only when has_shift is set */                  \
  X(mass_shift, 0, nL, 1) X(com_shift, 0, nL, 3)                                                                   \
  /* dynamic weld equalities (add_weld_constraint, solver.py:1293-1340): n_dyn_eq records of GSR_DYN_EQ_WORDS words   \
     [link1, link2, eq_data[11], sol_params[7]] */                                                                    \
  X(dyn_eq, 0, GSR_MAX_DYN_EQ, GSR_DYN_EQ_WORDS)                                                                    \
  /* per-environment dof parameters (RigidOptions.batch_dofs_info, array_class.py:1347-1364): nBD = nD when on, else 0 */ \
  X(bd_kp, 0, nBD, 1) X(bd_bias, 0, nBD, 3) X(bd_armature, 0, nBD, 1) X(bd_damping, 0, nBD, 1) X(bd_stiffness, 0, nBD, 1)   \
  X(bd_frictionloss, 0, nBD, 1) X(bd_limit, 0, nBD, 2) X(bd_force_range, 0, nBD, 2) X(bd_invweight, 0, nBD, 1)              \
  X(acc, 0, nD, 1) X(force, 0, nD, 1) X(qf_applied, 0, nD, 1) X(qf_constraint, 0, nD, 1)                           \
  X(cacc_lin, 0, nL, 3) X(cacc_ang, 0, nL, 3) X(links_contact_force, 0, nL, 3)                                     \
  X(aabb_min, 0, nG, 3) X(aabb_max, 0, nG, 3)                                                                      \
  X(n_constraints, 1, 1, 1) X(n_constraints_equality, 1, 1, 1) X(n_constraints_frictionloss, 1, 1, 1)              \
  X(solver_iters, 1, 1, 1) X(step_cost, 1, 1, 1)                                                                   \
  /* diagnostics (cfg_diagnostics): intermediate quantities the parity tests and get_mass_mat read */               \
  X(acc_smooth, 0, nD, 1) X(qf_smooth, 0, nD, 1) X(qf_passive, 0, nD, 1) X(qf_bias, 0, nD, 1)                      \
  X(mass_mat, 0, nD * nD, 1) X(mass_Dinv, 0, nD, 1) X(geoms_pos, 0, nG, 3) X(geoms_quat, 0, nG, 4)

struct HotLayout {
#define X(name, is_int, items, comps) int name;
  GSR_HOT_PRE_FIELDS(X)
  GSR_HOT_IN_FIELDS(X)
  GSR_HOT_OUT_FIELDS(X)
#undef X
  int in_start;    // first word of the IN part = size of the PRE block (multiple of 4)
  int n_state;     // end of the state block of the IN part (multiple of 4): all a lazy-kinematics step loads and stores
  int n_in;        // end of the IN part (multiple of 4): record words [in_start, n_in) are mirrored in shared memory
  int kin_dead;    // first word of the IN part that is dead after collision detection (links_pos ...)
  int hot_stride;  // words per env in the HOT segment (multiple of 4)
  int contact_stride, cache_stride, spill_stride;
  long long contact_base, cache_base, spill_base;  // word offsets of the segments inside the state buffer
  int cache_valid;                                 // offset of the validity bit words inside a CACHE record
  int spill_minvjt;                                // noslip: offset of the M^-1 J^T rows inside a SPILL record
};

// Shared-memory workspace of one environment.  Offsets are RECORD-RELATIVE: word w of the layout sits at shared-memory word
// w - in_start of the environment's slot (the kernels bias the slot pointer once, make_ctx), so that the IN part keeps the
// record's own offsets.  Words [in_start, n_in) mirror the record's IN
// part.  Regions are overlaid by liveness: the dynamics scratch and the kinematic cache that is dead after collision
// detection are reused by the collision scratch and then by the solver rows.
struct SmemLayout {
  // live for the whole step
  int M, L, Dinv, Hinv, force, acc_smooth, acc, qf_applied, qf_passive, qf_bias, qfrc_c;
  int MS;  // row stride of M / L / H (odd)
  // dynamics scratch
  int crb, fdof, cdd, cfrc;
  // collision scratch
  int gpos, gquat, aabb, queue;
  // contacts kept on chip for row assembly (n <= NCC) + solver
  int contacts, NCC;
  int H, qacc, Ma, grad, Mgrad, search, mv, prev_grad, prev_Mgrad, tmpv;
  int J, JS, NCS;                 // rows: J[NCS][JS]
  int r_aref, r_D, r_R, r_floss, r_Jaref, r_jv, r_force, r_active;
  // integrate / FK scratch
  int xanchor, xaxis, cacc, vel_next, qpos_next, pose_a, pose_b;
  int env_words;  // total words per env (multiple of 4, ≡ 8 mod 32 is not required)
};

// field lists for the model-specialised builds (gsr_model_spec_header writes them out as literals, gsr_spec_matches compares them)
#define GSR_SMEM_FIELDS(X)                                                                                                      \
  X(M) X(L) X(Dinv) X(Hinv) X(force) X(acc_smooth) X(acc) X(qf_applied) X(qf_passive) X(qf_bias) X(qfrc_c) X(MS) X(crb) X(fdof)   \
  X(cdd) X(cfrc) X(gpos) X(gquat) X(aabb) X(queue) X(contacts) X(NCC) X(H) X(qacc) X(Ma) X(grad) X(Mgrad) X(search) X(mv)         \
  X(prev_grad) X(prev_Mgrad) X(tmpv) X(J) X(JS) X(NCS) X(r_aref) X(r_D) X(r_R) X(r_floss) X(r_Jaref) X(r_jv) X(r_force)           \
  X(r_active) X(xanchor) X(xaxis) X(cacc) X(vel_next) X(qpos_next) X(pose_a) X(pose_b) X(env_words)
#define GSR_HOT_SCALAR_FIELDS(X)                                                                                                \
  X(in_start) X(n_state) X(n_in) X(kin_dead) X(hot_stride) X(contact_stride) X(cache_stride) X(spill_stride) X(cache_valid)       \
  X(spill_minvjt)
#define GSR_MODEL_SPEC_SCALARS(X)                                                                                               \
  X(nL) X(nJ) X(nD) X(nQ) X(nG) X(nE) X(nEq) X(nP) X(max_contact_pairs) X(max_broad_pairs) X(n_contacts_per_pair)                 \
  X(len_constraints) X(iterations) X(ls_iterations) X(integrator) X(solver_type) X(noslip_iterations) X(box_box_detection)        \
  X(use_gjk) X(gjk_fallback) X(enable_collision) X(enable_joint_limit) X(enable_multi_contact) X(sort_contacts) X(has_cc)         \
  X(n_cache_words) X(diagnostics) X(batch_dofs) X(disable_constraint) X(n_roots) X(max_depth) X(phase_barriers)

// Workspace of the lane-per-environment kernel (gsr_lane.cuh), word offsets per environment.  Words [0, n_head) mirror the
// head of the record (qpos .. dofs_pos, same offsets as HotLayout).  The pool is overlaid phase by phase.
struct LaneLayout {
  int n_head;
  int lpos, lquat, rcom, cdof;
  int M, L, Dinv, Hinv, H;
  int force, acc_smooth, acc, qf_applied, qfrc_c;
  int pool;
  int crb, cfrc, cvel, cdd;                    // dynamics
  int ct, lcf, aabb, queue, nqueue;            // collision (ct, lcf live through the constraint phase)
  int qacc, Ma, grad, Mgrad, search, mv, prev_grad, prev_Mgrad, tmpv;
  int r_aref, r_D, r_R, r_floss, r_Jaref, r_jv, r_force, r_active, J;
  int NCS, NCC;
  int cacc, out4, cnt, vn, qn;                 // outputs + integration
  int fcvel, rcom_out, cinr;                   // final kinematics
  int env_words;
};
